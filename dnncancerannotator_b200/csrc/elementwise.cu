// Memory-bound per-channel NHWC kernels: channel statistics, BatchNorm
// (finalize / apply / backward), 2x2 max-pool forward/backward, the MultiRes
// add-relu-affine tail, uint8->/255 input tail and dtype/slice conversion.
//
// Reference call sites: layers.BatchNormalization components.py:57,59,130,131;
// layers.MaxPool2D components.py:54; multiresunet.py:120-124,148-150.
// All kernels use the (pixel-lane, channel-lane) layout of common.cuh: coalesced
// along the contiguous channel axis, per-channel parameters in registers,
// fp64 block reductions -> one atomic per channel per CTA.
#include "common.cuh"

namespace dnnca {

template <typename T>
__device__ __forceinline__ const T* px_ptr(const View& v, long long p) {
  return reinterpret_cast<const T*>(v.data) + p * v.cstride + v.coff;
}
template <typename T>
__device__ __forceinline__ T* px_ptr_w(const View& v, long long p) {
  return reinterpret_cast<T*>(v.data) + p * v.cstride + v.coff;
}

// reduce `val` over the pixel lanes (stride CL in thread index) of a 256-thread block
__device__ __forceinline__ double block_reduce_pl(double val, double* sm, int CL, int PL) {
  const int tid = threadIdx.x;
  sm[tid] = val;
  __syncthreads();
  for (int s = PL >> 1; s > 0; s >>= 1) {
    if (tid < s * CL) sm[tid] += sm[tid + s * CL];
    __syncthreads();
  }
  double r = sm[tid % CL];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) channel_stats_kernel(View x, double* __restrict__ stats, int CL, int PL,
                                                           long long P) {
  __shared__ double sm[256];
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < ((x.c + CL - 1) / CL) * CL; c += CL) {
    double s = 0.0, ss = 0.0;
    if (c < x.c) {
      for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL) {
        float v = ldf(px_ptr<T>(x, p) + c);
        s += v;
        ss += (double)v * v;
      }
    }
    s = block_reduce_pl(s, sm, CL, PL);
    ss = block_reduce_pl(ss, sm, CL, PL);
    if (pl == 0 && c < x.c) {
      atomicAdd(stats + c, s);
      atomicAdd(stats + x.c + c, ss);
    }
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float momentum,
                                   float eps, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean = stats[c] / (double)count;
  double var = stats[C + c] / (double)count - mean * mean;
  if (var < 0.0) var = 0.0;
  float invstd = (float)(1.0 / sqrt(var + (double)eps));
  float g = gamma ? gamma[c] : 1.f;
  float scale = g * invstd;
  scale_shift[c] = scale;
  scale_shift[C + c] = beta[c] - (float)mean * scale;
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = invstd;
  if (moving_mean) {
    double unbiased = count > 1 ? var * ((double)count / (double)(count - 1)) : var;
    moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
    moving_var[c] = moving_var[c] * momentum + (float)unbiased * (1.f - momentum);
  }
}

__global__ void bn_inference_params_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                           float eps, const float* __restrict__ mm, const float* __restrict__ mv,
                                           float* __restrict__ scale_shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float scale = (gamma ? gamma[c] : 1.f) * rsqrtf(mv[c] + eps);
  scale_shift[c] = scale;
  scale_shift[C + c] = beta[c] - mm[c] * scale;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) bn_apply_kernel(View x, const float* __restrict__ ss, View y, int CL, int PL,
                                                      long long P) {
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < x.c; c += CL) {
    const float scale = ss ? ss[c] : 1.f, shift = ss ? ss[x.c + c] : 0.f;
    for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL)
      stf(px_ptr_w<TO>(y, p) + c, ldf(px_ptr<TI>(x, p) + c) * scale + shift);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(View x, View dy, const float* __restrict__ mi,
                                                           double* __restrict__ sums, int CL, int PL, long long P) {
  __shared__ double sm[256];
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < ((x.c + CL - 1) / CL) * CL; c += CL) {
    double s0 = 0.0, s1 = 0.0;
    if (c < x.c) {
      const float mean = mi[c], invstd = mi[x.c + c];
      for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL) {
        float g = ldf(px_ptr<T>(dy, p) + c);
        float xh = (ldf(px_ptr<T>(x, p) + c) - mean) * invstd;
        s0 += g;
        s1 += (double)g * xh;
      }
    }
    s0 = block_reduce_pl(s0, sm, CL, PL);
    s1 = block_reduce_pl(s1, sm, CL, PL);
    if (pl == 0 && c < x.c) {
      atomicAdd(sums + c, s0);
      atomicAdd(sums + x.c + c, s1);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(View x, View dy, const float* __restrict__ mi,
                                                          const float* __restrict__ gamma,
                                                          const double* __restrict__ sums, View dx, int act,
                                                          float alpha, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, int CL, int PL, long long P) {
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  const double invM = 1.0 / (double)P;
  for (int c = cl; c < x.c; c += CL) {
    const float mean = mi[c], invstd = mi[x.c + c];
    const float a = (float)(sums[c] * invM), b = (float)(sums[x.c + c] * invM);
    const float k = (gamma ? gamma[c] : 1.f) * invstd;
    if (blockIdx.x == 0 && pl == 0) {
      if (dgamma) dgamma[c] += (float)sums[x.c + c];
      if (dbeta) dbeta[c] += (float)sums[c];
    }
    for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL) {
      float xv = ldf(px_ptr<T>(x, p) + c);
      float xh = (xv - mean) * invstd;
      float g = k * (ldf(px_ptr<T>(dy, p) + c) - a - xh * b);
      g *= act_grad(xv, act, alpha);
      stf(px_ptr_w<T>(dx, p) + c, g);
    }
  }
}

// ---------------------------------------------------------------------------
// max-pool 2x2/2: thread (pl, cl) over OUTPUT pixels
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(View x, View y, uint8_t* __restrict__ idx,
                                                         double* __restrict__ stats, int CL, int PL, long long PO) {
  __shared__ double sm[256];
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  const int ho = y.h, wo = y.w;
  for (int c = cl; c < ((x.c + CL - 1) / CL) * CL; c += CL) {
    double s = 0.0, ss = 0.0;
    if (c < x.c) {
      for (long long p = (long long)blockIdx.x * PL + pl; p < PO; p += (long long)gridDim.x * PL) {
        int ox = (int)(p % wo);
        long long t = p / wo;
        int oy = (int)(t % ho);
        long long n = t / ho;
        long long p00 = (n * x.h + 2 * oy) * x.w + 2 * ox;
        float best = ldf(px_ptr<T>(x, p00) + c);
        int bi = 0;
        float v = ldf(px_ptr<T>(x, p00 + 1) + c);
        if (v > best) { best = v; bi = 1; }
        v = ldf(px_ptr<T>(x, p00 + x.w) + c);
        if (v > best) { best = v; bi = 2; }
        v = ldf(px_ptr<T>(x, p00 + x.w + 1) + c);
        if (v > best) { best = v; bi = 3; }
        stf(px_ptr_w<T>(y, p) + c, best);
        if (idx) idx[p * x.c + c] = (uint8_t)bi;
        s += best;
        ss += (double)best * best;
      }
    }
    if (stats) {
      s = block_reduce_pl(s, sm, CL, PL);
      ss = block_reduce_pl(ss, sm, CL, PL);
      if (pl == 0 && c < x.c) {
        atomicAdd(stats + c, s);
        atomicAdd(stats + x.c + c, ss);
      }
    }
  }
}

// thread (pl, cl) over pooled pixels; writes the four input-gradient pixels of its window
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(View dy, const uint8_t* __restrict__ idx, View dskip,
                                                         int has_skip, View dx, View mask, int has_mask, int act,
                                                         float alpha, int CL, int PL, long long PO) {
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  const int ho = dy.h, wo = dy.w;
  for (int c = cl; c < dy.c; c += CL) {
    for (long long p = (long long)blockIdx.x * PL + pl; p < PO; p += (long long)gridDim.x * PL) {
      int ox = (int)(p % wo);
      long long t = p / wo;
      int oy = (int)(t % ho);
      long long n = t / ho;
      long long p00 = (n * dx.h + 2 * oy) * dx.w + 2 * ox;
      float g = ldf(px_ptr<T>(dy, p) + c);
      int bi = idx[p * dy.c + c];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        long long q = p00 + (k >> 1) * dx.w + (k & 1);
        float v = (k == bi) ? g : 0.f;
        if (has_skip) v += ldf(px_ptr<T>(dskip, q) + c);
        if (has_mask) v *= act_grad(ldf(px_ptr<T>(mask, q) + c), act, alpha);
        stf(px_ptr_w<T>(dx, q) + c, v);
      }
    }
  }
}

// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_relu_affine_kernel(View a, const float* __restrict__ fa, View b,
                                                             const float* __restrict__ fb,
                                                             const float* __restrict__ fo, View y, int CL, int PL,
                                                             long long P) {
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < a.c; c += CL) {
    const float sa = fa ? fa[c] : 1.f, ta = fa ? fa[a.c + c] : 0.f;
    const float sb = fb ? fb[c] : 1.f, tb = fb ? fb[a.c + c] : 0.f;
    const float so = fo ? fo[c] : 1.f, to = fo ? fo[a.c + c] : 0.f;
    for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL) {
      float v = ldf(px_ptr<T>(a, p) + c) * sa + ta + ldf(px_ptr<T>(b, p) + c) * sb + tb;
      v = v > 0.f ? v : 0.f;
      stf(px_ptr_w<T>(y, p) + c, v * so + to);
    }
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) u8_to_unit_kernel(const uint8_t* __restrict__ src, long long count,
                                                        TO* __restrict__ dst) {
  // 16 source bytes per thread-iteration (128-bit load)
  const long long nvec = count / 16;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    uint4 v = __ldg(s4 + i);
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = (float)((w[k >> 2] >> ((k & 3) * 8)) & 0xffu) / 255.0f;
    if (sizeof(TO) == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {          // 4 x 16-byte stores
      float4* d4 = reinterpret_cast<float4*>(dst + i * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) d4[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    } else if (sizeof(TO) == 2 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {   // 2 x 16-byte stores
      uint4* d4 = reinterpret_cast<uint4*>(dst + i * 16);
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(f[8 * k], f[8 * k + 1]), p1 = __floats2bfloat162_rn(f[8 * k + 2], f[8 * k + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(f[8 * k + 4], f[8 * k + 5]), p3 = __floats2bfloat162_rn(f[8 * k + 6], f[8 * k + 7]);
        d4[k] = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                           *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) stf(dst + i * 16 + k, f[k]);
    }
  }
  for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x)
    stf(dst + i, (float)src[i] / 255.0f);
}


// ---------------------------------------------------------------------------
// 128-bit vectorised bf16 variants (C % 8 == 0, 16-byte aligned slices): thread = (pixel lane, group of 8 channels).
// One uint4 load/store moves 8 channels; per-channel parameters of the group stay in registers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&p);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ const uint4* vptr(const View& v, long long p, int g) {
  return reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(v.data) + p * v.cstride + v.coff + 8 * g);
}
__device__ __forceinline__ uint4* vptr_w(const View& v, long long p, int g) {
  return reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(v.data) + p * v.cstride + v.coff + 8 * g);
}

// 8 channels (one 16-byte access) per thread-iteration; the three affine tables sit in shared memory
__global__ void __launch_bounds__(256) add_relu_affine_vec8_kernel(View a, const float* __restrict__ fa, View b,
                                                                  const float* __restrict__ fb,
                                                                  const float* __restrict__ fo, View y, long long P) {
  extern __shared__ float tab[];                    // [6][C]: sa ta sb tb so to
  const int C = a.c, ng = C / 8;
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    tab[i] = fa ? fa[i] : 1.f;          tab[C + i] = fa ? fa[C + i] : 0.f;
    tab[2 * C + i] = fb ? fb[i] : 1.f;  tab[3 * C + i] = fb ? fb[C + i] : 0.f;
    tab[4 * C + i] = fo ? fo[i] : 1.f;  tab[5 * C + i] = fo ? fo[C + i] : 0.f;
  }
  __syncthreads();
  const long long total = P * ng;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long p = e / ng;
    const int g = (int)(e - p * ng);
    const uint4 ra = __ldg(vptr(a, p, g)), rb = __ldg(vptr(b, p, g));
    const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * g + j;
      const float va = __uint_as_float((j & 1) ? (wa[j >> 1] & 0xffff0000u) : (wa[j >> 1] << 16));
      const float vb = __uint_as_float((j & 1) ? (wb[j >> 1] & 0xffff0000u) : (wb[j >> 1] << 16));
      float v = fmaf(va, tab[c], tab[C + c]) + fmaf(vb, tab[2 * C + c], tab[3 * C + c]);
      v = fmaxf(v, 0.f);
      o[j] = fmaf(v, tab[4 * C + c], tab[5 * C + c]);
    }
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 q = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&q);
    }
    *vptr_w(y, p, g) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// Same operation with the affine tables in REGISTERS: the 256 threads of a block are (pixel lane, channel group) with the
// group fastest, so a thread keeps its group for the whole grid-stride loop and loads its 6 x 8 parameters once.  The
// shared-memory variant above reads 48 table words per 16-byte element; at the MultiRes widths (15 / 28 / 54 / 108 groups,
// no power of two) those reads also conflict: add_relu_affine[120@128] ran at 2.4 TB/s, [224@64] at 1.8 TB/s against
// 6.1 TB/s for [32@256] (profiles/r02h_multires_sweep.json).  ng <= 256.
// HA / HB: affine_a / affine_b present (an absent table costs no registers: 4 resident blocks per SM).
template <bool HA, bool HB>
__global__ void __launch_bounds__(256, 4) add_relu_affine_vec8_reg_kernel(View a, const float* __restrict__ fa, View b,
                                                                         const float* __restrict__ fb,
                                                                         const float* __restrict__ fo, View y, long long P,
                                                                         int ng, int ppb) {
  const int g = threadIdx.x % ng, pl = threadIdx.x / ng;
  if (pl >= ppb) return;                      // 256 - ppb * ng idle threads (no barrier follows)
  const int C = a.c;
  float sa[8], ta[8], sb[8], tb[8], so[8], to[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = 8 * g + j;
    sa[j] = HA ? fa[c] : 1.f; ta[j] = HA ? fa[C + c] : 0.f;
    sb[j] = HB ? fb[c] : 1.f; tb[j] = HB ? fb[C + c] : 0.f;
    so[j] = fo ? fo[c] : 1.f; to[j] = fo ? fo[C + c] : 0.f;
  }
  for (long long p = (long long)blockIdx.x * ppb + pl; p < P; p += (long long)gridDim.x * ppb) {
    const uint4 ra = __ldg(vptr(a, p, g)), rb = __ldg(vptr(b, p, g));
    const uint32_t wa[4] = {ra.x, ra.y, ra.z, ra.w}, wb[4] = {rb.x, rb.y, rb.z, rb.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float va = __uint_as_float((j & 1) ? (wa[j >> 1] & 0xffff0000u) : (wa[j >> 1] << 16));
      const float vb = __uint_as_float((j & 1) ? (wb[j >> 1] & 0xffff0000u) : (wb[j >> 1] << 16));
      float v = (HA ? fmaf(va, sa[j], ta[j]) : va) + (HB ? fmaf(vb, sb[j], tb[j]) : vb);
      v = fmaxf(v, 0.f);
      o[j] = fmaf(v, so[j], to[j]);
    }
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 q = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
      w[j] = *reinterpret_cast<uint32_t*>(&q);
    }
    *vptr_w(y, p, g) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// MODE 0: channel stats (x)            -> acc[0..C) += sum x, acc[C..2C) += sum x^2
// MODE 1: BN backward reduce (x, dy)   -> acc[0..C) += sum dy, acc[C..2C) += sum dy*xhat
template <int MODE>
__global__ void __launch_bounds__(256) reduce_vec8_kernel(View x, View dy, const float* __restrict__ mi,
                                                         double* __restrict__ acc, int GL, int PL, long long P) {
  // per-thread partials of 8 channels x 2 quantities meet in shared memory ONCE per channel group (two barriers)
  __shared__ double sm[256 * 16];
  const int gl = threadIdx.x & (GL - 1), pl = threadIdx.x / GL;
  const int ng = x.c / 8;
  for (int g = gl; g < ((ng + GL - 1) / GL) * GL; g += GL) {
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    double d0[8], d1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d0[j] = d1[j] = 0.0;
    if (g < ng) {
      float mean[8], inv[8];
      if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { mean[j] = mi[8 * g + j]; inv[j] = mi[x.c + 8 * g + j]; }
      }
      int cnt = 0;
      const long long stride = (long long)gridDim.x * PL;
      long long p = (long long)blockIdx.x * PL + pl;
      // four pixels per iteration: all loads are issued before the first use (memory-level parallelism)
      for (; p + 3 * stride < P; p += 4 * stride) {
        uint4 xr[4], gr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xr[u] = __ldg(vptr(x, p + u * stride, g));
          if (MODE == 1) gr[u] = __ldg(vptr(dy, p + u * stride, g));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float xv[8];
          unpack8(xr[u], xv);
          if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { s0[j] += xv[j]; s1[j] = fmaf(xv[j], xv[j], s1[j]); }
          } else {
            float gv[8];
            unpack8(gr[u], gv);
#pragma unroll
            for (int j = 0; j < 8; ++j) { s0[j] += gv[j]; s1[j] = fmaf(gv[j], (xv[j] - mean[j]) * inv[j], s1[j]); }
          }
        }
        if ((cnt += 4) >= 64) {       // flush the fp32 partials into fp64 every 64 pixels
#pragma unroll
          for (int j = 0; j < 8; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; s0[j] = s1[j] = 0.f; }
          cnt = 0;
        }
      }
      for (; p < P; p += stride) {
        float xv[8];
        unpack8(*vptr(x, p, g), xv);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s0[j] += xv[j]; s1[j] = fmaf(xv[j], xv[j], s1[j]); }
        } else {
          float gv[8];
          unpack8(*vptr(dy, p, g), gv);
#pragma unroll
          for (int j = 0; j < 8; ++j) { s0[j] += gv[j]; s1[j] = fmaf(gv[j], (xv[j] - mean[j]) * inv[j], s1[j]); }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; }
    }
    // lanes of a warp that share a channel group (GL < 32: lane = pixel lane * GL + gl) meet through shuffles first, so the
    // shared-memory pass sums 8 warp partials (GL <= 32) or PL <= 4 pixel lanes (GL > 32) instead of up to 128 terms
    int terms = PL, slot = threadIdx.x;
    if (GL <= 32) {
      for (int off = GL; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d0[j] += __shfl_xor_sync(0xffffffffu, d0[j], off);
          d1[j] += __shfl_xor_sync(0xffffffffu, d1[j], off);
        }
      }
      terms = 8;
      slot = (threadIdx.x & 31) < GL ? (threadIdx.x >> 5) * GL + gl : -1;
    }
    // sm[(term*GL + gl)*16 + q]
    if (slot >= 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sm[slot * 16 + j] = d0[j]; sm[slot * 16 + 8 + j] = d1[j]; }
    }
    __syncthreads();
    // 16*GL sums of `terms` partials, spread over the 256 threads
    for (int o = threadIdx.x; o < 16 * GL; o += 256) {
      const int q = o & 15, gg = o >> 4;
      double r = 0.0;
      for (int l = 0; l < terms; ++l) r += sm[(l * GL + gg) * 16 + q];
      const int gch = g - gl + gg;               // channel group of lane gg in this pass
      if (gch < ng) atomicAdd(acc + (q < 8 ? 0 : x.c) + 8 * gch + (q & 7), r);
    }
    __syncthreads();
  }
}

// MODE 0: y = x*scale + shift (BN apply);  MODE 1: BN backward apply (dx = k*(dy - a - xhat*b) * act'(x))
template <int MODE>
__global__ void __launch_bounds__(256) map_vec8_kernel(View x, View dy, const float* __restrict__ p0,
                                                      const float* __restrict__ gamma, const double* __restrict__ sums,
                                                      View out, int act, float alpha, float* __restrict__ dgamma,
                                                      float* __restrict__ dbeta, int GL, int PL, long long P) {
  const int gl = threadIdx.x & (GL - 1), pl = threadIdx.x / GL;
  const int ng = x.c / 8;
  const double invM = 1.0 / (double)P;
  for (int g = gl; g < ng; g += GL) {
    float c0[8], c1[8], c2[8], c3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = 8 * g + j;
      if (MODE == 0) {
        c0[j] = p0[c];               // scale
        c1[j] = p0[x.c + c];         // shift
      } else {
        c0[j] = p0[c];               // mean
        c1[j] = p0[x.c + c];         // invstd
        c2[j] = (float)(sums[c] * invM);
        c3[j] = (float)(sums[x.c + c] * invM);
        if (blockIdx.x == 0 && pl == 0) {
          if (dgamma) dgamma[c] += (float)sums[x.c + c];
          if (dbeta) dbeta[c] += (float)sums[c];
        }
      }
    }
    float kk[8];
    if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) kk[j] = (gamma ? gamma[8 * g + j] : 1.f) * c1[j];
    }
    const long long stride = (long long)gridDim.x * PL;
    long long p = (long long)blockIdx.x * PL + pl;
    // four pixels per iteration: all loads are issued before the first use (memory-level parallelism)
    for (; p + 3 * stride < P; p += 4 * stride) {
      uint4 xr[4], gr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xr[u] = __ldg(vptr(x, p + u * stride, g));
        if (MODE == 1) gr[u] = __ldg(vptr(dy, p + u * stride, g));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float xv[8], o[8];
        unpack8(xr[u], xv);
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[j], c0[j], c1[j]);
        } else {
          float gv[8];
          unpack8(gr[u], gv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xh = (xv[j] - c0[j]) * c1[j];
            o[j] = kk[j] * (gv[j] - c2[j] - xh * c3[j]) * act_grad(xv[j], act, alpha);
          }
        }
        __stcs(vptr_w(out, p + u * stride, g), pack8(o));      // streaming: the result is not re-read by this kernel
      }
    }
    for (; p < P; p += stride) {
      float xv[8], o[8];
      unpack8(*vptr(x, p, g), xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(xv[j], c0[j], c1[j]);
      } else {
        float gv[8];
        unpack8(*vptr(dy, p, g), gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - c0[j]) * c1[j];
          o[j] = kk[j] * (gv[j] - c2[j] - xh * c3[j]) * act_grad(xv[j], act, alpha);
        }
      }
      *vptr_w(out, p, g) = pack8(o);
    }
  }
}

// grid of a grid-stride kernel: enough blocks for the work, at most ONE resident wave (a block count that is not a multiple
// of the resident slots leaves a partial last wave: the 118-register BN backward reduction held 2 blocks per SM and ran
// 1024 blocks = 3.46 waves)
// ---------------------------------------------------------------------------
// Inference-time weight folding with physical channel padding (MultiResUnet: 8/17/26/... channel tensors live in
// buffers padded to multiples of 8 so the tensor-core kernels take them; holes carry zero weights and stay zero)
__global__ void __launch_bounds__(256) fold_weights_kernel(const float* __restrict__ w, int taps, int cin, int cout, int layout,
                                                          const int* __restrict__ in_map, int cin_p,
                                                          const int* __restrict__ out_map, int cout_p,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                                          float* __restrict__ w_out, float* __restrict__ b_out) {
  const long long total = (long long)taps * cin_p * cout_p;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int tap, cip, cop;
    if (layout == 0) { cop = (int)(e % cout_p); const long long t = e / cout_p; cip = (int)(t % cin_p); tap = (int)(t / cin_p); }
    else { cip = (int)(e % cin_p); const long long t = e / cin_p; cop = (int)(t % cout_p); tap = (int)(t / cout_p); }
    const int ci = in_map ? in_map[cip] : cip, co = out_map ? out_map[cop] : cop;
    float v = 0.f;
    if (ci >= 0 && co >= 0 && ci < cin && co < cout) {
      v = layout == 0 ? w[((long long)tap * cin + ci) * cout + co] : w[((long long)tap * cout + co) * cin + ci];
      if (mv) v *= (gamma ? gamma[co] : 1.f) * rsqrtf(mv[co] + eps);
    }
    w_out[e] = v;
    if (b_out && tap == 0 && cip == 0) {
      float b = 0.f;
      if (co >= 0 && co < cout) {
        if (mv) { const float sc = (gamma ? gamma[co] : 1.f) * rsqrtf(mv[co] + eps); b = (beta ? beta[co] : 0.f) - mm[co] * sc; }
        else b = beta ? beta[co] : 0.f;
      }
      b_out[cop] = b;
    }
  }
}

__global__ void bn_inference_params_mapped_kernel(int CP, const int* __restrict__ map, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, float eps, const float* __restrict__ mm,
                                                  const float* __restrict__ mv, float* __restrict__ scale_shift) {
  const int cp = blockIdx.x * blockDim.x + threadIdx.x;
  if (cp >= CP) return;
  const int c = map ? map[cp] : cp;
  float scale = 0.f, shift = 0.f;
  if (c >= 0) {
    scale = (gamma ? gamma[c] : 1.f) * rsqrtf(mv[c] + eps);
    shift = beta[c] - mm[c] * scale;
  }
  scale_shift[cp] = scale;
  scale_shift[CP + cp] = shift;
}

template <typename K>
inline int resident_grid(K kern, long long work_items, int per_block) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * per_sm;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// blocks of a reduction over P pixels with PL pixel lanes per block: 64 pixels per thread when that still fills the
// device, down to 16 for small tensors (a thread's pixels are a serial chain of load round trips)
inline int reduce_blocks(long long P, int pl, int resident) {
  for (int ppt = 64; ppt > 16; ppt >>= 1) {
    const long long b = (P + (long long)pl * ppt - 1) / ((long long)pl * ppt);
    if (b >= resident) return (int)b;
  }
  const long long b = (P + (long long)pl * 16 - 1) / ((long long)pl * 16);
  return (int)(b < 1 ? 1 : b);
}

inline bool vec8_ok(const dnnca_tensor_t* t) {
  return t->dtype == DNNCA_BF16 && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}
inline ChanLayout group_layout(int c) {
  int gl = 1;
  while (gl < c / 8 && gl < 256) gl <<= 1;
  return ChanLayout{gl, 256 / gl};
}

}  // namespace dnnca

using namespace dnnca;

// ============================ C ABI =========================================
namespace dnnca {
int try_maxpool_fwd_vec(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, uint8_t*, double*);
int try_maxpool_bwd_vec(cudaStream_t, const dnnca_tensor_t*, const uint8_t*, const dnnca_tensor_t*, const dnnca_tensor_t*,
                        const dnnca_tensor_t*, int, float);
int try_convert_vec(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*);
}  // namespace dnnca

extern "C" int dnnca_channel_stats(void* stream, const dnnca_tensor_t* x, double* stats) {
  DNNCA_CHECK_ARG(view_ok(x) && stats, "channel_stats: bad arguments");
  long long P = (long long)x->n * x->h * x->w;
  if (vec8_ok(x)) {
    ChanLayout G = group_layout(x->c);
    // >= 64 pixels per thread: every block ends with 2*C fp64 atomics, so small tensors get few blocks
    static const int occ0 = resident_grid(reduce_vec8_kernel<0>, 1LL << 40, 1);      // resident blocks on this device
    const int g0 = reduce_blocks(P, G.pl, occ0);
    reduce_vec8_kernel<0><<<g0 < occ0 ? g0 : occ0, 256, 0, (cudaStream_t)stream>>>(mk(x), mk(x), nullptr, stats, G.cl, G.pl, P);
    DNNCA_LAUNCH_CHECK("channel_stats");
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  int grid = grid_for(P, L.pl * 8);
  DNNCA_DISPATCH_DTYPE(x->dtype, channel_stats_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(x), stats, L.cl, L.pl, P);)
  DNNCA_LAUNCH_CHECK("channel_stats");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_finalize(void* stream, const double* stats, int64_t count, int c, const float* gamma,
                                 const float* beta, float momentum, float eps, float* moving_mean,
                                 float* moving_var, float* scale_shift, float* mean_invstd) {
  DNNCA_CHECK_ARG(stats && beta && scale_shift && mean_invstd && c > 0 && count > 0, "bn_finalize: bad arguments");
  DNNCA_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize: moving stats must come in pairs");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, count, c, gamma, beta, momentum, eps,
                                                                         moving_mean, moving_var, scale_shift,
                                                                         mean_invstd);
  DNNCA_LAUNCH_CHECK("bn_finalize");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_inference_params(void* stream, int c, const float* gamma, const float* beta, float eps,
                                         const float* moving_mean, const float* moving_var, float* scale_shift) {
  DNNCA_CHECK_ARG(c > 0 && beta && moving_mean && moving_var && scale_shift, "bn_inference_params: bad arguments");
  bn_inference_params_kernel<<<(c + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c, gamma, beta, eps, moving_mean,
                                                                                 moving_var, scale_shift);
  DNNCA_LAUNCH_CHECK("bn_inference_params");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_apply(void* stream, const dnnca_tensor_t* x, const float* scale_shift,
                              const dnnca_tensor_t* y) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && same_shape(x, y) && scale_shift, "bn_apply: bad arguments");
  DNNCA_CHECK_ARG(x->dtype == y->dtype, "bn_apply: dtype mismatch");
  long long P = (long long)x->n * x->h * x->w;
  if (vec8_ok(x) && vec8_ok(y)) {
    ChanLayout G = group_layout(x->c);
    map_vec8_kernel<0><<<grid_for(P, G.pl * 4), 256, 0, (cudaStream_t)stream>>>(mk(x), mk(x), scale_shift, nullptr, nullptr, mk(y),
                                                                                 0, 0.f, nullptr, nullptr, G.cl, G.pl, P);
    DNNCA_LAUNCH_CHECK("bn_apply");
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  int grid = grid_for(P, L.pl * 4);
  DNNCA_DISPATCH_DTYPE(x->dtype, (bn_apply_kernel<T, T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(x), scale_shift, mk(y), L.cl, L.pl, P));)
  DNNCA_LAUNCH_CHECK("bn_apply");
  return DNNCA_OK;
}

extern "C" int dnnca_convert(void* stream, const dnnca_tensor_t* src, const dnnca_tensor_t* dst) {
  DNNCA_CHECK_ARG(view_ok(src) && view_ok(dst) && same_shape(src, dst), "convert: bad arguments");
  {
    const int r = try_convert_vec((cudaStream_t)stream, src, dst);
    if (r != 0) return r < 0 ? r : DNNCA_OK;
  }
  ChanLayout L = chan_layout(src->c);
  long long P = (long long)src->n * src->h * src->w;
  int grid = grid_for(P, L.pl * 4);
  cudaStream_t s = (cudaStream_t)stream;
  if (src->dtype == DNNCA_F32 && dst->dtype == DNNCA_F32)
    bn_apply_kernel<float, float><<<grid, 256, 0, s>>>(mk(src), nullptr, mk(dst), L.cl, L.pl, P);
  else if (src->dtype == DNNCA_F32)
    bn_apply_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(mk(src), nullptr, mk(dst), L.cl, L.pl, P);
  else if (dst->dtype == DNNCA_F32)
    bn_apply_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>(mk(src), nullptr, mk(dst), L.cl, L.pl, P);
  else
    bn_apply_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(mk(src), nullptr, mk(dst), L.cl, L.pl, P);
  DNNCA_LAUNCH_CHECK("convert");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_bwd_reduce(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy,
                                   const float* mean_invstd, double* sums) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(dy) && same_shape(x, dy) && mean_invstd && sums, "bn_bwd_reduce: bad arguments");
  DNNCA_CHECK_ARG(x->dtype == dy->dtype, "bn_bwd_reduce: dtype mismatch");
  long long P = (long long)x->n * x->h * x->w;
  if (vec8_ok(x) && vec8_ok(dy)) {
    ChanLayout G = group_layout(x->c);
    static const int occ1 = resident_grid(reduce_vec8_kernel<1>, 1LL << 40, 1);
    const int g1 = reduce_blocks(P, G.pl, occ1);
    reduce_vec8_kernel<1><<<g1 < occ1 ? g1 : occ1, 256, 0, (cudaStream_t)stream>>>(mk(x), mk(dy), mean_invstd, sums, G.cl, G.pl, P);
    DNNCA_LAUNCH_CHECK("bn_bwd_reduce");
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  int grid = grid_for(P, L.pl * 8);
  DNNCA_DISPATCH_DTYPE(x->dtype, bn_bwd_reduce_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(x), mk(dy), mean_invstd, sums, L.cl, L.pl, P);)
  DNNCA_LAUNCH_CHECK("bn_bwd_reduce");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_bwd_apply(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy,
                                  const float* mean_invstd, const float* gamma, const double* sums,
                                  const dnnca_tensor_t* dx, int act, float alpha, float* dgamma, float* dbeta) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(dy) && view_ok(dx) && same_shape(x, dy) && same_shape(x, dx) && mean_invstd && sums,
                  "bn_bwd_apply: bad arguments");
  DNNCA_CHECK_ARG(x->dtype == dy->dtype && x->dtype == dx->dtype, "bn_bwd_apply: dtype mismatch");
  long long P = (long long)x->n * x->h * x->w;
  if (vec8_ok(x) && vec8_ok(dy) && vec8_ok(dx)) {
    ChanLayout G = group_layout(x->c);
    map_vec8_kernel<1><<<grid_for(P, G.pl * 4), 256, 0, (cudaStream_t)stream>>>(mk(x), mk(dy), mean_invstd, gamma, sums, mk(dx), act,
                                                                                 alpha, dgamma, dbeta, G.cl, G.pl, P);
    DNNCA_LAUNCH_CHECK("bn_bwd_apply");
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  int grid = grid_for(P, L.pl * 4);
  DNNCA_DISPATCH_DTYPE(x->dtype, bn_bwd_apply_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
      mk(x), mk(dy), mean_invstd, gamma, sums, mk(dx), act, alpha, dgamma, dbeta, L.cl, L.pl, P);)
  DNNCA_LAUNCH_CHECK("bn_bwd_apply");
  return DNNCA_OK;
}

extern "C" int dnnca_maxpool2x2_fwd(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* y, uint8_t* idx,
                                    double* stats) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y), "maxpool_fwd: bad arguments");
  DNNCA_CHECK_ARG(x->h % 2 == 0 && x->w % 2 == 0 && y->h == x->h / 2 && y->w == x->w / 2 && y->n == x->n && y->c == x->c,
                  "maxpool_fwd: y must be [n,h/2,w/2,c] of an even-sized x");
  DNNCA_CHECK_ARG(x->dtype == y->dtype, "maxpool_fwd: dtype mismatch");
  {
    const int r = try_maxpool_fwd_vec((cudaStream_t)stream, x, y, idx, stats);
    if (r != 0) return r < 0 ? r : DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  long long PO = (long long)y->n * y->h * y->w;
  int grid = grid_for(PO, L.pl * 4);
  DNNCA_DISPATCH_DTYPE(x->dtype, maxpool_fwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(x), mk(y), idx, stats, L.cl, L.pl, PO);)
  DNNCA_LAUNCH_CHECK("maxpool_fwd");
  return DNNCA_OK;
}

extern "C" int dnnca_maxpool2x2_bwd(void* stream, const dnnca_tensor_t* dy, const uint8_t* idx,
                                    const dnnca_tensor_t* dskip, const dnnca_tensor_t* dx,
                                    const dnnca_tensor_t* mask, int act, float alpha) {
  DNNCA_CHECK_ARG(view_ok(dy) && view_ok(dx) && idx, "maxpool_bwd: bad arguments");
  DNNCA_CHECK_ARG(dx->h == 2 * dy->h && dx->w == 2 * dy->w && dx->n == dy->n && dx->c == dy->c, "maxpool_bwd: shape mismatch");
  DNNCA_CHECK_ARG(!dskip || (view_ok(dskip) && same_shape(dskip, dx) && dskip->dtype == dx->dtype), "maxpool_bwd: bad dskip");
  DNNCA_CHECK_ARG(!mask || (view_ok(mask) && same_shape(mask, dx) && mask->dtype == dx->dtype), "maxpool_bwd: bad mask");
  DNNCA_CHECK_ARG(dy->dtype == dx->dtype, "maxpool_bwd: dtype mismatch");
  {
    const int r = try_maxpool_bwd_vec((cudaStream_t)stream, dy, idx, dskip, dx, mask, act, alpha);
    if (r != 0) return r < 0 ? r : DNNCA_OK;
  }
  ChanLayout L = chan_layout(dy->c);
  long long PO = (long long)dy->n * dy->h * dy->w;
  int grid = grid_for(PO, L.pl * 2);
  View vs = dskip ? mk(dskip) : mk(dx), vm = mask ? mk(mask) : mk(dx);
  DNNCA_DISPATCH_DTYPE(dy->dtype, maxpool_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
      mk(dy), idx, vs, dskip != nullptr, mk(dx), vm, mask != nullptr, act, alpha, L.cl, L.pl, PO);)
  DNNCA_LAUNCH_CHECK("maxpool_bwd");
  return DNNCA_OK;
}

extern "C" int dnnca_add_relu_affine(void* stream, const dnnca_tensor_t* a, const float* affine_a,
                                     const dnnca_tensor_t* b, const float* affine_b, const float* affine_out,
                                     const dnnca_tensor_t* y) {
  DNNCA_CHECK_ARG(view_ok(a) && view_ok(b) && view_ok(y) && same_shape(a, b) && same_shape(a, y), "add_relu_affine: bad arguments");
  DNNCA_CHECK_ARG(a->dtype == b->dtype && a->dtype == y->dtype, "add_relu_affine: dtype mismatch");
  long long P = (long long)a->n * a->h * a->w;
  // <= 4 groups (<= 32 channels): the shared-memory tables are read as warp-wide broadcasts there (6.1 TB/s measured)
  if (vec8_ok(a) && vec8_ok(b) && vec8_ok(y) && a->c / 8 > 4 && a->c / 8 <= 256) {
    const int ng = a->c / 8, ppb = 256 / ng;
    const int grid = grid_for(P, ppb * 4, 8);
    cudaStream_t s = (cudaStream_t)stream;
#define LAUNCH_ARA(HA, HB) add_relu_affine_vec8_reg_kernel<HA, HB><<<grid, 256, 0, s>>>(mk(a), affine_a, mk(b), affine_b, affine_out, \
                                                                                     mk(y), P, ng, ppb)
    if (affine_a && affine_b) LAUNCH_ARA(true, true);
    else if (affine_a) LAUNCH_ARA(true, false);
    else if (affine_b) LAUNCH_ARA(false, true);
    else LAUNCH_ARA(false, false);
#undef LAUNCH_ARA
    DNNCA_LAUNCH_CHECK("add_relu_affine");
    return DNNCA_OK;
  }
  if (vec8_ok(a) && vec8_ok(b) && vec8_ok(y) && a->c * 6 * 4 <= 48 * 1024) {
    add_relu_affine_vec8_kernel<<<grid_for(P * (a->c / 8), 256 * 4, 8), 256, (size_t)a->c * 6 * 4, (cudaStream_t)stream>>>(
        mk(a), affine_a, mk(b), affine_b, affine_out, mk(y), P);
    DNNCA_LAUNCH_CHECK("add_relu_affine");
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(a->c);
  int grid = grid_for(P, L.pl * 4);
  DNNCA_DISPATCH_DTYPE(a->dtype, add_relu_affine_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
      mk(a), affine_a, mk(b), affine_b, affine_out, mk(y), L.cl, L.pl, P);)
  DNNCA_LAUNCH_CHECK("add_relu_affine");
  return DNNCA_OK;
}

extern "C" int dnnca_u8_to_unit(void* stream, const uint8_t* src, int64_t count, void* dst, int dtype) {
  DNNCA_CHECK_ARG(src && dst && count > 0 && (dtype == DNNCA_F32 || dtype == DNNCA_BF16), "u8_to_unit: bad arguments");
  DNNCA_CHECK_ARG(((uintptr_t)src & 15) == 0, "u8_to_unit: src must be 16-byte aligned");
  int grid = grid_for(count / 16 + 1, 256);
  if (dtype == DNNCA_F32)
    u8_to_unit_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, count, (float*)dst);
  else
    u8_to_unit_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, count, (__nv_bfloat16*)dst);
  DNNCA_LAUNCH_CHECK("u8_to_unit");
  return DNNCA_OK;
}

extern "C" int dnnca_fold_weights(void* stream, const float* w, int taps, int cin, int cout, int layout,
                                  const int32_t* in_map, int cin_phys, const int32_t* out_map, int cout_phys,
                                  const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                                  float eps, float* w_out, float* b_out) {
  DNNCA_CHECK_ARG(w && w_out && taps > 0 && cin > 0 && cout > 0 && cin_phys > 0 && cout_phys > 0 && (layout == 0 || layout == 1),
                  "fold_weights: bad arguments");
  DNNCA_CHECK_ARG((in_map || cin_phys == cin) && (out_map || cout_phys == cout), "fold_weights: a NULL map needs equal logical / physical counts");
  DNNCA_CHECK_ARG((moving_mean == nullptr) == (moving_var == nullptr), "fold_weights: moving statistics come in pairs");
  const long long total = (long long)taps * cin_phys * cout_phys;
  fold_weights_kernel<<<grid_for(total, 256 * 4, 4), 256, 0, (cudaStream_t)stream>>>(w, taps, cin, cout, layout, in_map, cin_phys,
                                                                                    out_map, cout_phys, gamma, beta, moving_mean,
                                                                                    moving_var, eps, w_out, b_out);
  DNNCA_LAUNCH_CHECK("fold_weights");
  return DNNCA_OK;
}

extern "C" int dnnca_bn_inference_params_mapped(void* stream, int c_phys, const int32_t* map, const float* gamma,
                                                const float* beta, float eps, const float* moving_mean,
                                                const float* moving_var, float* scale_shift) {
  DNNCA_CHECK_ARG(c_phys > 0 && beta && moving_mean && moving_var && scale_shift, "bn_inference_params_mapped: bad arguments");
  bn_inference_params_mapped_kernel<<<(c_phys + 127) / 128, 128, 0, (cudaStream_t)stream>>>(c_phys, map, gamma, beta, eps,
                                                                                           moving_mean, moving_var, scale_shift);
  DNNCA_LAUNCH_CHECK("bn_inference_params_mapped");
  return DNNCA_OK;
}
