// Fused segmentation head + weighted binary cross-entropy.
//
// Replaces, in ONE pass over the feature map, what the reference runs as
// Conv2D(1, k=1, 'sigmoid') (unet.py:241-244), tf_get_positive_rate
// (losses.py:87-102), tf_weighted_crossentropy (losses.py:17-37) and their
// gradients (~15 TF elementwise/reduce kernels + a 1x1 conv fwd/bwd):
//   z = f.w + b ; p = sigmoid(z) ; loss ; dz ; df = dz*w (*act'(f)) ; dw ; db.
// Memory-bound: reads f and the label once, writes df (and p/z) once.
#include <string.h>

#include "common.cuh"

namespace dnnca {

__device__ __forceinline__ uint32_t float_key(float f) {  // order-preserving
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void label_stats_init_kernel(dnnca_label_stats_t* s) {
  s->sum = 0.0;
  s->min_key = 0xffffffffu;
  s->max_key = 0u;
}

__global__ void __launch_bounds__(256) label_stats_kernel(const float* __restrict__ label, long long count,
                                                         dnnca_label_stats_t* __restrict__ out) {
  double s = 0.0;
  float mn = INFINITY, mx = -INFINITY;
  const long long nvec = ((reinterpret_cast<uintptr_t>(label) & 15) == 0) ? count / 4 : 0;
  const float4* l4 = reinterpret_cast<const float4*>(label);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = l4[i];
    s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (long long i = nvec * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float v = label[i];
    s += v;
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  s = warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double ss[8];
  __shared__ float smn[8], smx[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { ss[warp] = s; smn[warp] = mn; smx[warp] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { s += ss[i]; mn = fminf(mn, smn[i]); mx = fmaxf(mx, smx[i]); }
    atomicAdd(&out->sum, s);
    if (mn <= mx) {
      atomicMin(&out->min_key, float_key(mn));
      atomicMax(&out->max_key, float_key(mx));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) head_fwd_kernel(View f, const float* __restrict__ w,
                                                      const float* __restrict__ b, float* __restrict__ logits,
                                                      float* __restrict__ probs, long long P) {
  const int F = f.c;
  const float bias = b ? b[0] : 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const T* fp = reinterpret_cast<const T*>(f.data) + p * f.cstride + f.coff;
    float z = bias;
    for (int c = 0; c < F; ++c) z = fmaf(ldf(fp + c), w[c], z);
    if (logits) logits[p] = z;
    if (probs) probs[p] = 1.f / (1.f + expf(-z));
  }
}

// bf16 features, F = 8 * NG with NG a power of two <= 32: NG lanes per pixel, one 16-byte load each, the dot product is
// finished with shuffles.  The scalar kernel above reads a pixel's F features from ONE thread (2-byte loads, 128-byte
// stride between the lanes of a warp): head_fwd[64@256] ran at 0.58 TB/s and cost 8 % of the MultiResUnet forward
// (profiles/r02h_multires_sweep.json).
template <int NG>
__global__ void __launch_bounds__(256) head_fwd_vec_kernel(View f, const float* __restrict__ w, const float* __restrict__ b,
                                                          float* __restrict__ logits, float* __restrict__ probs, long long P) {
  const int g = threadIdx.x % NG;
  float wr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wr[j] = w[8 * g + j];
  const float bias = b ? b[0] : 0.f;
  const long long total = P * NG, padded = (total + 31) / 32 * 32;      // warp-uniform trip count (shuffles inside)
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < padded; e += (long long)gridDim.x * blockDim.x) {
    const bool live = e < total;                       // the NG lanes of a pixel are live or dead together
    const long long p = live ? e / NG : P - 1;
    const uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(f.data) + p * f.cstride + f.coff + 8 * g);
    const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      z = fmaf(__uint_as_float(wd[j] << 16), wr[2 * j], z);
      z = fmaf(__uint_as_float(wd[j] & 0xffff0000u), wr[2 * j + 1], z);
    }
#pragma unroll
    for (int o = 1; o < NG; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (live && g == 0) {
      z += bias;
      if (logits) logits[p] = z;
      if (probs) probs[p] = 1.f / (1.f + expf(-z));
    }
  }
}

// FMAX = compile-time bound on the feature count (register-resident dw partials)
template <typename T, int FMAX>
__global__ void __launch_bounds__(256) head_bce_kernel(View f, const float* __restrict__ w,
                                                      const float* __restrict__ b, const float* __restrict__ label,
                                                      const dnnca_label_stats_t* __restrict__ ls,
                                                      dnnca_loss_config_t cfg, float* __restrict__ logits,
                                                      float* __restrict__ probs, float* __restrict__ per_sample,
                                                      View df, int has_df, int act, float alpha,
                                                      float* __restrict__ dw, float* __restrict__ db) {
  // grid = (chunks over H*W, B): every block stays inside one sample
  const int F = f.c;
  const long long HW = (long long)f.h * f.w;
  const long long total = HW * f.n;
  float weight;
  if (cfg.has_weight) {
    weight = cfg.weight;
  } else {
    const float r = (float)(ls->sum / (double)total);  // losses.py:95
    weight = r > 0.f ? 1.f / r : 1.f;                  // losses.py:27
  }
  weight = cfg.weight_mul * weight + cfg.weight_add;    // losses.py:29
  const float bias = b ? b[0] : 0.f;
  float wr[FMAX];
#pragma unroll
  for (int c = 0; c < FMAX; ++c) wr[c] = c < F ? w[c] : 0.f;
  float dwacc[FMAX];
#pragma unroll
  for (int c = 0; c < FMAX; ++c) dwacc[c] = 0.f;
  float dbacc = 0.f, lossacc = 0.f;
  const long long base = (long long)blockIdx.y * HW;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < HW;
       q += (long long)gridDim.x * blockDim.x) {
    const long long p = base + q;
    const T* fp = reinterpret_cast<const T*>(f.data) + p * f.cstride + f.coff;
    float fv[FMAX];
    float z = bias;
#pragma unroll
    for (int c = 0; c < FMAX; ++c) {
      fv[c] = c < F ? ldf(fp + c) : 0.f;
      z = fmaf(fv[c], wr[c], z);
    }
    const float y = label[p];
    const float mask = y * (weight - 1.f) + 1.f;                          // losses.py:31
    const float bce = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));    // from_logits BCE
    lossacc += bce * mask;
    const float pr = 1.f / (1.f + expf(-z));
    if (logits) logits[p] = z;
    if (probs) probs[p] = pr;
    const float dz = mask * (pr - y) * cfg.grad_scale;
    dbacc += dz;
    T* dp = has_df ? reinterpret_cast<T*>(df.data) + p * df.cstride + df.coff : nullptr;
#pragma unroll
    for (int c = 0; c < FMAX; ++c) {
      if (c < F) {
        dwacc[c] = fmaf(dz, fv[c], dwacc[c]);
        if (has_df) stf(dp + c, dz * wr[c] * act_grad(fv[c], act, alpha));
      }
    }
  }
  // block reduction: warp shuffles then 8 partials through smem
  __shared__ float sm[8][FMAX + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  lossacc = warp_sum(lossacc);
  dbacc = warp_sum(dbacc);
#pragma unroll
  for (int c = 0; c < FMAX; ++c) dwacc[c] = warp_sum(dwacc[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < FMAX; ++c) sm[warp][c] = dwacc[c];
    sm[warp][FMAX] = dbacc;
    sm[warp][FMAX + 1] = lossacc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < FMAX + 2; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][c];
    if (c < F) {
      if (dw) atomicAdd(dw + c, s);
    } else if (c == FMAX) {
      if (db) atomicAdd(db, s);
    } else if (c == FMAX + 1) {
      atomicAdd(per_sample + blockIdx.y, s / (float)HW);                 // losses.py:36
    }
  }
}


// 8 pixels per thread for DENSE bf16 features with few channels (configs/unet.yaml: F = 3): F 16-byte loads of f, two
// float4 of labels; df / probs / logits leave as 16-byte stores.  Same per-pixel arithmetic as head_bce_kernel.
template <int F>
__global__ void __launch_bounds__(256) head_bce_pix8_kernel(const __nv_bfloat16* __restrict__ f, const float* __restrict__ w,
                                                           const float* __restrict__ b, const float* __restrict__ label,
                                                           const dnnca_label_stats_t* __restrict__ ls, dnnca_loss_config_t cfg,
                                                           float* __restrict__ logits, float* __restrict__ probs,
                                                           float* __restrict__ per_sample, __nv_bfloat16* __restrict__ df,
                                                           int act, float alpha, float* __restrict__ dw, float* __restrict__ db,
                                                           long long HW, long long total) {
  float weight;
  if (cfg.has_weight) {
    weight = cfg.weight;
  } else {
    const float r = (float)(ls->sum / (double)total);  // losses.py:95
    weight = r > 0.f ? 1.f / r : 1.f;                  // losses.py:27
  }
  weight = cfg.weight_mul * weight + cfg.weight_add;    // losses.py:29
  const float bias = b ? b[0] : 0.f;
  float wr[F], dwacc[F];
#pragma unroll
  for (int c = 0; c < F; ++c) { wr[c] = w[c]; dwacc[c] = 0.f; }
  float dbacc = 0.f, lossacc = 0.f;
  const long long base = (long long)blockIdx.y * HW;
  const long long ngroups = HW / 8;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
    const long long p0 = base + g * 8;
    float fv[8 * F];
    {
      const uint4* q = reinterpret_cast<const uint4*>(f + p0 * F);
#pragma unroll
      for (int i = 0; i < F; ++i) {
        const uint4 v = __ldg(q + i);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          fv[i * 8 + 2 * j] = __uint_as_float(u[j] << 16);
          fv[i * 8 + 2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
        }
      }
    }
    float yv[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(label + p0)), c = __ldg(reinterpret_cast<const float4*>(label + p0) + 1);
      yv[0] = a.x; yv[1] = a.y; yv[2] = a.z; yv[3] = a.w; yv[4] = c.x; yv[5] = c.y; yv[6] = c.z; yv[7] = c.w;
    }
    float zv[8], pv[8], dfv[8 * F];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float z = bias;
#pragma unroll
      for (int c = 0; c < F; ++c) z = fmaf(fv[k * F + c], wr[c], z);
      const float y = yv[k];
      const float mask = y * (weight - 1.f) + 1.f;                          // losses.py:31
      const float bce = fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)));    // from_logits BCE
      lossacc += bce * mask;
      const float pr = 1.f / (1.f + expf(-z));
      zv[k] = z; pv[k] = pr;
      const float dz = mask * (pr - y) * cfg.grad_scale;
      dbacc += dz;
#pragma unroll
      for (int c = 0; c < F; ++c) {
        dwacc[c] = fmaf(dz, fv[k * F + c], dwacc[c]);
        dfv[k * F + c] = dz * wr[c] * act_grad(fv[k * F + c], act, alpha);
      }
    }
    if (logits) {
      float4* q = reinterpret_cast<float4*>(logits + p0);
      q[0] = make_float4(zv[0], zv[1], zv[2], zv[3]); q[1] = make_float4(zv[4], zv[5], zv[6], zv[7]);
    }
    if (probs) {
      float4* q = reinterpret_cast<float4*>(probs + p0);
      q[0] = make_float4(pv[0], pv[1], pv[2], pv[3]); q[1] = make_float4(pv[4], pv[5], pv[6], pv[7]);
    }
    if (df) {
      uint4* q = reinterpret_cast<uint4*>(df + p0 * F);
#pragma unroll
      for (int i = 0; i < F; ++i) {
        __nv_bfloat162 a0 = __floats2bfloat162_rn(dfv[i * 8], dfv[i * 8 + 1]), a1 = __floats2bfloat162_rn(dfv[i * 8 + 2], dfv[i * 8 + 3]);
        __nv_bfloat162 a2 = __floats2bfloat162_rn(dfv[i * 8 + 4], dfv[i * 8 + 5]), a3 = __floats2bfloat162_rn(dfv[i * 8 + 6], dfv[i * 8 + 7]);
        q[i] = make_uint4(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1), *reinterpret_cast<uint32_t*>(&a2),
                          *reinterpret_cast<uint32_t*>(&a3));
      }
    }
  }
  __shared__ float sm[8][F + 2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  lossacc = warp_sum(lossacc);
  dbacc = warp_sum(dbacc);
#pragma unroll
  for (int c = 0; c < F; ++c) dwacc[c] = warp_sum(dwacc[c]);
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < F; ++c) sm[warp][c] = dwacc[c];
    sm[warp][F] = dbacc;
    sm[warp][F + 1] = lossacc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < F + 2; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sm[i][c];
    if (c < F) {
      if (dw) atomicAdd(dw + c, s);
    } else if (c == F) {
      if (db) atomicAdd(db, s);
    } else {
      atomicAdd(per_sample + blockIdx.y, s / (float)HW);                 // losses.py:36
    }
  }
}


// 128-bit vectorised bf16 variant for F % 8 == 0 (unet_big F=64, mulmo F=16): LPP = F/8 lanes share a pixel, each lane
// owns 8 channels (one uint4 of f, one of df); the logit is a shuffle-reduction over the LPP lanes.
template <int LPP>
__global__ void __launch_bounds__(256) head_bce_vec_kernel(View f, const float* __restrict__ w,
                                                          const float* __restrict__ b, const float* __restrict__ label,
                                                          const dnnca_label_stats_t* __restrict__ ls,
                                                          dnnca_loss_config_t cfg, float* __restrict__ logits,
                                                          float* __restrict__ probs, float* __restrict__ per_sample,
                                                          View df, int has_df, int act, float alpha,
                                                          float* __restrict__ dw, float* __restrict__ db) {
  constexpr int F = LPP * 8;
  constexpr int PPB = 256 / LPP;                    // pixels per block iteration
  __shared__ float red[F + 2];
  for (int i = threadIdx.x; i < F + 2; i += 256) red[i] = 0.f;
  __syncthreads();
  const long long HW = (long long)f.h * f.w;
  const long long total = HW * f.n;
  float weight;
  if (cfg.has_weight) {
    weight = cfg.weight;
  } else {
    const float r = (float)(ls->sum / (double)total);
    weight = r > 0.f ? 1.f / r : 1.f;
  }
  weight = cfg.weight_mul * weight + cfg.weight_add;
  const float bias = b ? b[0] : 0.f;
  const int sub = threadIdx.x % LPP, pslot = threadIdx.x / LPP;
  float wr[8], dwacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { wr[j] = w[sub * 8 + j]; dwacc[j] = 0.f; }
  float dbacc = 0.f, lossacc = 0.f;
  const long long base = (long long)blockIdx.y * HW;
  // the trip count is warp-uniform (the LPP lanes of a pixel meet in shuffles): lanes past the last pixel of a
  // ragged tail (H*W not a multiple of 256/LPP) re-read the last pixel and are masked out of every write / sum
  for (long long q0 = (long long)blockIdx.x * PPB; q0 < HW; q0 += (long long)gridDim.x * PPB) {
    const long long q = q0 + pslot;
    const bool live = q < HW;
    const long long p = base + (live ? q : HW - 1);
    const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(f.data) + p * f.cstride + f.coff + sub * 8);
    const uint32_t wd[4] = {raw.x, raw.y, raw.z, raw.w};
    float fv[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { fv[2 * j] = __uint_as_float(wd[j] << 16); fv[2 * j + 1] = __uint_as_float(wd[j] & 0xffff0000u); }
    float z = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) z = fmaf(fv[j], wr[j], z);
#pragma unroll
    for (int o = 1; o < LPP; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    if (!live) continue;
    z += bias;
    const float y = label[p];
    const float mask = y * (weight - 1.f) + 1.f;
    const float pr = 1.f / (1.f + expf(-z));
    const float dz = mask * (pr - y) * cfg.grad_scale;
    if (sub == 0) {
      lossacc += (fmaxf(z, 0.f) - z * y + log1pf(expf(-fabsf(z)))) * mask;
      dbacc += dz;
      if (logits) logits[p] = z;
      if (probs) probs[p] = pr;
    }
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dwacc[j] = fmaf(dz, fv[j], dwacc[j]);
      o8[j] = dz * wr[j] * act_grad(fv[j], act, alpha);
    }
    if (has_df) {
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 t = __floats2bfloat162_rn(o8[2 * j], o8[2 * j + 1]);
        ow[j] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(df.data) + p * df.cstride + df.coff + sub * 8) =
          make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  // lanes with the same channel group: reduce over the warp first (stride LPP), then one smem atomic per warp
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = dwacc[j];
#pragma unroll
    for (int o = LPP; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) < LPP) atomicAdd(red + sub * 8 + j, v);
  }
  lossacc = warp_sum(lossacc);
  dbacc = warp_sum(dbacc);
  if ((threadIdx.x & 31) == 0) { atomicAdd(red + F, dbacc); atomicAdd(red + F + 1, lossacc); }
  __syncthreads();
  for (int c = threadIdx.x; c < F + 2; c += 256) {
    if (c < F) { if (dw) atomicAdd(dw + c, red[c]); }
    else if (c == F) { if (db) atomicAdd(db, red[c]); }
    else atomicAdd(per_sample + blockIdx.y, red[c] / (float)HW);
  }
}

// d(sum of sigmoid outputs)/d(features): df[p,c] = p(1-p) * w[c] * act'(f[p,c])   (callbacks.py:290-299 takes
// g.gradient(model(x), x); this seeds the dgrad chain at the head)
template <typename T>
__global__ void __launch_bounds__(256) head_input_grad_kernel(View f, const float* __restrict__ w,
                                                             const float* __restrict__ b, View df, int act, float alpha,
                                                             long long P) {
  const int F = f.c;
  const float bias = b ? b[0] : 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const T* fp = reinterpret_cast<const T*>(f.data) + p * f.cstride + f.coff;
    T* dp = reinterpret_cast<T*>(df.data) + p * df.cstride + df.coff;
    float z = bias;
    for (int c = 0; c < F; ++c) z = fmaf(ldf(fp + c), w[c], z);
    const float pr = 1.f / (1.f + expf(-z));
    const float dz = pr * (1.f - pr);
    for (int c = 0; c < F; ++c) stf(dp + c, dz * w[c] * act_grad(ldf(fp + c), act, alpha));
  }
}

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_head_input_grad(void* stream, const dnnca_tensor_t* f, const float* w, const float* b,
                                     const dnnca_tensor_t* df, int act, float alpha) {
  DNNCA_CHECK_ARG(view_ok(f) && view_ok(df) && same_shape(f, df) && f->dtype == df->dtype && w, "head_input_grad: bad arguments");
  const long long P = (long long)f->n * f->h * f->w;
  const int grid = grid_for(P, 256, 8);
  DNNCA_DISPATCH_DTYPE(f->dtype, head_input_grad_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(f), w, b, mk(df), act, alpha, P);)
  DNNCA_LAUNCH_CHECK("head_input_grad");
  return DNNCA_OK;
}

extern "C" int dnnca_label_stats_init(void* stream, dnnca_label_stats_t* lstats) {
  DNNCA_CHECK_ARG(lstats, "label_stats_init: null");
  label_stats_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(lstats);
  DNNCA_LAUNCH_CHECK("label_stats_init");
  return DNNCA_OK;
}

extern "C" int dnnca_label_stats(void* stream, const float* label, int64_t count, dnnca_label_stats_t* lstats) {
  DNNCA_CHECK_ARG(label && lstats && count > 0, "label_stats: bad arguments");
  int grid = grid_for(count / 4 + 1, 256 * 4, 4);
  label_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(label, count, lstats);
  DNNCA_LAUNCH_CHECK("label_stats");
  return DNNCA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Label smoothing (losses.py:62-67): tfa.image.gaussian_filter2d(label[..., None], filter_shape, sigma) with the default
// padding 'REFLECT' -- a separable normalised gaussian on taps range(-size//2 + 1, size//2 + 1) (6 -> -2..3), applied
// along W then along H, (size-1)//2 reflected samples before and size-1-(size-1)//2 after (index -1 -> 1, n -> n-2).
// ---------------------------------------------------------------------------------------------------------------
struct GaussTaps { float g[32]; };
template <bool VERTICAL>
__global__ void __launch_bounds__(256) gauss1d_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total,
                                                      int H, int W, int size, int before, GaussTaps taps) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long r = i / W;
    const int y = (int)(r % H);
    const int n = VERTICAL ? H : W, c = VERTICAL ? y : x;
    const float* line = VERTICAL ? src + (r - y) * W + x : src + r * W;
    const int stride = VERTICAL ? W : 1;
    float acc = 0.f;
    for (int k = 0; k < size; ++k) {
      int j = c + k - before;
      if (j < 0) j = -j;
      if (j >= n) j = 2 * n - 2 - j;
      acc += taps.g[k] * line[(long long)j * stride];
    }
    dst[i] = acc;
  }
}

extern "C" int dnnca_gaussian_filter2d(void* stream, const float* label, int n, int h, int w, int filter_size, float sigma,
                                       float* tmp, float* out) {
  DNNCA_CHECK_ARG(label && tmp && out && n > 0 && h > 0 && w > 0, "gaussian_filter2d: bad arguments");
  DNNCA_CHECK_ARG(filter_size >= 1 && filter_size <= 32 && sigma > 0.f, "gaussian_filter2d: filter_size in 1..32, sigma > 0");
  const int before = (filter_size - 1) / 2, after = filter_size - 1 - before;
  DNNCA_CHECK_ARG(before < h && after < h && before < w && after < w, "gaussian_filter2d: REFLECT padding needs size < image");
  GaussTaps t;
  const int k0 = -(filter_size / 2) + 1 - ((filter_size & 1) ? 1 : 0);      // python: -size // 2 + 1 (floor division)
  float sum = 0.f;
  for (int k = 0; k < filter_size; ++k) {
    const float d = (float)(k0 + k);
    t.g[k] = expf(-(d * d) / (2.0f * sigma * sigma));
    sum += t.g[k];
  }
  for (int k = 0; k < filter_size; ++k) t.g[k] /= sum;
  const long long total = (long long)n * h * w;
  const int grid = grid_for(total, 256 * 4, 8);
  gauss1d_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(label, tmp, total, h, w, filter_size, before, t);
  DNNCA_LAUNCH_CHECK("gaussian_filter2d (rows)");
  gauss1d_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(tmp, out, total, h, w, filter_size, before, t);
  DNNCA_LAUNCH_CHECK("gaussian_filter2d (columns)");
  return DNNCA_OK;
}

extern "C" void dnnca_label_stats_decode(const dnnca_label_stats_t* h, double* sum, float* mn, float* mx) {
  auto dec = [](uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    memcpy(&f, &u, 4);
    return f;
  };
  if (sum) *sum = h->sum;
  if (mn) *mn = dec(h->min_key);
  if (mx) *mx = dec(h->max_key);
}

extern "C" int dnnca_head_fwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* b, float* logits,
                              float* probs) {
  DNNCA_CHECK_ARG(view_ok(f) && w && (logits || probs), "head_fwd: bad arguments");
  long long P = (long long)f->n * f->h * f->w;
  const int ngv = f->c / 8;
  if (f->dtype == DNNCA_BF16 && f->c % 8 == 0 && ngv <= 32 && (ngv & (ngv - 1)) == 0 && f->coff % 8 == 0 && f->cstride % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(f->data) & 15) == 0) {
    const int gv = grid_for(P * ngv, 256 * 4, 8);
    cudaStream_t s = (cudaStream_t)stream;
    switch (ngv) {
      case 1: head_fwd_vec_kernel<1><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
      case 2: head_fwd_vec_kernel<2><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
      case 4: head_fwd_vec_kernel<4><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
      case 8: head_fwd_vec_kernel<8><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
      case 16: head_fwd_vec_kernel<16><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
      default: head_fwd_vec_kernel<32><<<gv, 256, 0, s>>>(mk(f), w, b, logits, probs, P); break;
    }
    DNNCA_LAUNCH_CHECK("head_fwd");
    return DNNCA_OK;
  }
  int grid = grid_for(P, 256, 8);
  DNNCA_DISPATCH_DTYPE(f->dtype, head_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(f), w, b, logits, probs, P);)
  DNNCA_LAUNCH_CHECK("head_fwd");
  return DNNCA_OK;
}

extern "C" int dnnca_head_bce_fwd_bwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* b,
                                      const float* label, const dnnca_label_stats_t* lstats,
                                      const dnnca_loss_config_t* cfg, float* logits, float* probs,
                                      float* per_sample, const dnnca_tensor_t* df, int act, float alpha, float* dw,
                                      float* db) {
  DNNCA_CHECK_ARG(view_ok(f) && w && label && cfg && per_sample, "head_bce: bad arguments");
  DNNCA_CHECK_ARG(cfg->has_weight || lstats, "head_bce: label stats needed when no explicit weight is given");
  DNNCA_CHECK_ARG(!df || (view_ok(df) && same_shape(df, f) && df->dtype == f->dtype), "head_bce: bad df");
  if (f->c > 64) DNNCA_UNSUPPORTED("head_bce: at most 64 head features supported (got %d)", f->c);
  const long long HW = (long long)f->h * f->w;
  int gx = (int)((HW + 256 * 4 - 1) / (256 * 4));
  int cap = (sm_count() * 8 + f->n - 1) / f->n;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  dim3 grid(gx, f->n);
  View vdf = df ? mk(df) : mk(f);
  cudaStream_t s = (cudaStream_t)stream;
  auto al = [](const dnnca_tensor_t* t) {
    return t->dtype == DNNCA_BF16 && t->coff % 8 == 0 && t->cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
  };
  {   // few dense bf16 features: 8 pixels per thread, 16-byte accesses
    auto dense = [](const dnnca_tensor_t* t) {
      return t->dtype == DNNCA_BF16 && t->coff == 0 && t->cstride == t->c && (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
    };
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if ((f->c == 3 || f->c == 4 || f->c == 6) && HW % 8 == 0 && dense(f) && (!df || dense(df)) && al16(label) &&
        (!logits || al16(logits)) && (!probs || al16(probs))) {
      int gp = (int)((HW / 8 + 255) / 256);
      if (gp > cap) gp = cap;
      if (gp < 1) gp = 1;
      dim3 g3(gp, f->n);
      const __nv_bfloat16* fp = reinterpret_cast<const __nv_bfloat16*>(f->data);
      __nv_bfloat16* dfp = df ? reinterpret_cast<__nv_bfloat16*>(df->data) : nullptr;
#define LAUNCH_P8(FF) head_bce_pix8_kernel<FF><<<g3, 256, 0, s>>>(fp, w, b, label, lstats, *cfg, logits, probs, per_sample, dfp, act, \
                                                                  alpha, dw, db, HW, HW * f->n)
      if (f->c == 3) LAUNCH_P8(3); else if (f->c == 4) LAUNCH_P8(4); else LAUNCH_P8(6);
#undef LAUNCH_P8
      DNNCA_LAUNCH_CHECK("head_bce_fwd_bwd");
      return DNNCA_OK;
    }
  }
  if ((f->c == 16 || f->c == 32 || f->c == 64) && al(f) && (!df || al(df))) {
    const int lpp = f->c / 8;
    int gv = (int)((HW + (256 / lpp) * 8 - 1) / ((256 / lpp) * 8));
    if (gv > cap) gv = cap;
    if (gv < 1) gv = 1;
    dim3 g2(gv, f->n);
#define LAUNCH_VEC(L) head_bce_vec_kernel<L><<<g2, 256, 0, s>>>(mk(f), w, b, label, lstats, *cfg, logits, probs, per_sample, vdf, \
                                                                df != nullptr, act, alpha, dw, db)
    if (lpp == 2) LAUNCH_VEC(2); else if (lpp == 4) LAUNCH_VEC(4); else LAUNCH_VEC(8);
#undef LAUNCH_VEC
    DNNCA_LAUNCH_CHECK("head_bce_fwd_bwd");
    return DNNCA_OK;
  }
#define LAUNCH_HEAD(FM)                                                                                          \
  DNNCA_DISPATCH_DTYPE(f->dtype, (head_bce_kernel<T, FM><<<grid, 256, 0, s>>>(mk(f), w, b, label, lstats, *cfg, logits, \
                                                                               probs, per_sample, vdf, df != nullptr, \
                                                                               act, alpha, dw, db));)
  if (f->c <= 4) { LAUNCH_HEAD(4) }
  else if (f->c <= 16) { LAUNCH_HEAD(16) }
  else if (f->c <= 32) { LAUNCH_HEAD(32) }
  else { LAUNCH_HEAD(64) }
#undef LAUNCH_HEAD
  DNNCA_LAUNCH_CHECK("head_bce_fwd_bwd");
  return DNNCA_OK;
}
