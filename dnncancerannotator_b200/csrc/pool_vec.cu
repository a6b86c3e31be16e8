// 128-bit vectorised MaxPool2D([2,2], strides=2) forward / backward and dtype conversion for DENSE bf16 tensors with few
// channels (configs/unet.yaml: 3 / 6 / 12; components.py:54 and its gradient).  The scalar kernels in elementwise.cu
// move 2 bytes per access (1.2 TB/s measured); here a thread owns 24 consecutive output elements (G = 24/C pooled
// pixels x C channels = 48 bytes) and the 2 x 48 input elements above them, all as 16-byte accesses.  Arithmetic and
// tie-breaking (first maximum in row-major window order) are those of the scalar kernels, so results are bit-identical.
#include "common.cuh"

namespace dnnca {

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// 48 bf16 (6 x uint4) -> float[48]
__device__ __forceinline__ void load48(const __nv_bfloat16* p, float* f) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const uint4 v = __ldg(q + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[i * 8 + 2 * j] = bf_lo(w[j]); f[i * 8 + 2 * j + 1] = bf_hi(w[j]); }
  }
}
__device__ __forceinline__ void load48_plain(const __nv_bfloat16* p, float* f) {   // may alias a later store
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const uint4 v = q[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { f[i * 8 + 2 * j] = bf_lo(w[j]); f[i * 8 + 2 * j + 1] = bf_hi(w[j]); }
  }
}
__device__ __forceinline__ void store48(__nv_bfloat16* p, const float* f) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i)
    q[i] = make_uint4(pack2(f[i * 8], f[i * 8 + 1]), pack2(f[i * 8 + 2], f[i * 8 + 3]), pack2(f[i * 8 + 4], f[i * 8 + 5]),
                      pack2(f[i * 8 + 6], f[i * 8 + 7]));
}

// thread = (pooled row r over n*Ho, group j of 24 output elements); groups = Wo*C/24 per row
template <int C>
__global__ void __launch_bounds__(256) maxpool_fwd_vec_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                             uint8_t* __restrict__ idx, int groups, long long total) {
  constexpr int G = 24 / C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / groups;
    const int j = (int)(t - r * groups);
    const long long in_row = (long long)groups * 48;               // input elements per row
    const __nv_bfloat16* x0 = x + (2 * r) * in_row + 48 * j;
    float a[48], b[48];
    load48(x0, a);
    load48(x0 + in_row, b);
    float o[24];
    uint32_t ib[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int e = g * C + c;
        float best = a[(2 * g) * C + c];
        uint32_t bi = 0;
        float v = a[(2 * g + 1) * C + c];
        if (v > best) { best = v; bi = 1; }
        v = b[(2 * g) * C + c];
        if (v > best) { best = v; bi = 2; }
        v = b[(2 * g + 1) * C + c];
        if (v > best) { best = v; bi = 3; }
        o[e] = best;
        ib[e >> 2] |= bi << ((e & 3) * 8);
      }
    uint4* yq = reinterpret_cast<uint4*>(y + r * ((long long)groups * 24) + 24 * j);
#pragma unroll
    for (int i = 0; i < 3; ++i)
      yq[i] = make_uint4(pack2(o[i * 8], o[i * 8 + 1]), pack2(o[i * 8 + 2], o[i * 8 + 3]), pack2(o[i * 8 + 4], o[i * 8 + 5]),
                         pack2(o[i * 8 + 6], o[i * 8 + 7]));
    if (idx) {
      uint2* iq = reinterpret_cast<uint2*>(idx + r * ((long long)groups * 24) + 24 * j);
      iq[0] = make_uint2(ib[0], ib[1]);
      iq[1] = make_uint2(ib[2], ib[3]);
      iq[2] = make_uint2(ib[4], ib[5]);
    }
  }
}

// MODE: 0 = no activation mask, 1 = ReLU mask, 2 = LeakyReLU mask
template <int C, int MODE>
__global__ void __launch_bounds__(256) maxpool_bwd_vec_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                             const __nv_bfloat16* dskip, __nv_bfloat16* dx,
                                                             const __nv_bfloat16* __restrict__ mask, float alpha, int groups,
                                                             long long total) {
  constexpr int G = 24 / C;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / groups;
    const int j = (int)(t - r * groups);
    const long long in_row = (long long)groups * 48;
    const long long out_off = r * ((long long)groups * 24) + 24 * j;
    float g24[24];
    {
      const uint4* q = reinterpret_cast<const uint4*>(dy + out_off);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const uint4 v = __ldg(q + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) { g24[i * 8 + 2 * k] = bf_lo(w[k]); g24[i * 8 + 2 * k + 1] = bf_hi(w[k]); }
      }
    }
    uint32_t ib[6];
    {
      const uint2* iq = reinterpret_cast<const uint2*>(idx + out_off);
#pragma unroll
      for (int i = 0; i < 3; ++i) { const uint2 v = __ldg(iq + i); ib[2 * i] = v.x; ib[2 * i + 1] = v.y; }
    }
#pragma unroll
    for (int ar = 0; ar < 2; ++ar) {
      const long long off = (2 * r + ar) * in_row + 48 * j;
      float f[48];
      if (dskip) load48_plain(dskip + off, f);
      else {
#pragma unroll
        for (int e = 0; e < 48; ++e) f[e] = 0.f;
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int e = g * C + c;
          const uint32_t bi = (ib[e >> 2] >> ((e & 3) * 8)) & 0xffu;
#pragma unroll
          for (int bc = 0; bc < 2; ++bc) {
            const int q = (2 * g + bc) * C + c;
            f[q] = (bi == (uint32_t)(2 * ar + bc) ? g24[e] : 0.f) + f[q];
          }
        }
      if (MODE != 0) {
        float m[48];
        load48(mask + off, m);
#pragma unroll
        for (int e = 0; e < 48; ++e) f[e] *= (m[e] > 0.f ? 1.f : (MODE == 1 ? 0.f : alpha));
      }
      store48(dx + off, f);
    }
  }
}

// flat fp32 -> bf16 (8 elements per thread-iteration)
__global__ void __launch_bounds__(256) f32_to_bf16_vec_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                             long long nvec) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = __ldg(s4 + 2 * i), b = __ldg(s4 + 2 * i + 1);
    d4[i] = make_uint4(pack2(a.x, a.y), pack2(a.z, a.w), pack2(b.x, b.y), pack2(b.z, b.w));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// C % 8 == 0 (configs/unet_big.yaml, mulmo_unet.yaml): thread = (pooled pixel lane, group of 8 channels), one 16-byte
// access per tensor and window position; channel-slice views allowed.  The forward kernel can also accumulate the
// BatchNormalization statistics of its output (components.py:59) like reduce_vec8_kernel does.
// ---------------------------------------------------------------------------------------------------------------
struct PV { const __nv_bfloat16* p; long long cs; };     // base pointer (already at the slice's first channel), pixel stride
__device__ __forceinline__ void unpack8f(const uint4& v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) { f[2 * j] = bf_lo(w[j]); f[2 * j + 1] = bf_hi(w[j]); }
}

// AFF: the input carries a folded BatchNorm affine s*a + t (bn_fold.cu): the window is searched for the max of a where
// s >= 0 and for the min where s < 0 (compare sign*a), the stored output / statistics are those of s*a_sel + t
template <bool STATS, bool AFF = false>
__global__ void __launch_bounds__(256) maxpool_fwd_vec8_kernel(PV x, __nv_bfloat16* y, long long ycs, uint8_t* __restrict__ idx,
                                                              double* __restrict__ stats, int C, int Ho, int Wo, long long PO,
                                                              int GL, int PL, const float* __restrict__ aff = nullptr) {
  __shared__ double sm[STATS ? 256 * 16 : 1];
  const int gl = threadIdx.x & (GL - 1), pl = threadIdx.x / GL;
  const int ng = C / 8;
  const long long W = 2LL * Wo;
  for (int g = gl; g < ((ng + GL - 1) / GL) * GL; g += GL) {
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s0[j] = s1[j] = 0.f;
    double d0[8], d1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) d0[j] = d1[j] = 0.0;
    if (g < ng) {
      int cnt = 0;
      float sc[8], sh[8], sg[8];
      if (AFF) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = aff[8 * g + j]; sh[j] = aff[C + 8 * g + j]; sg[j] = sc[j] < 0.f ? -1.f : 1.f; }
      }
      for (long long p = (long long)blockIdx.x * PL + pl; p < PO; p += (long long)gridDim.x * PL) {
        const int ox = (int)(p % Wo);
        const long long t = p / Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        const long long q00 = (n * 2 * Ho + 2 * oy) * W + 2 * ox;
        uint4 r[4];
        r[0] = __ldg(reinterpret_cast<const uint4*>(x.p + q00 * x.cs + 8 * g));
        r[1] = __ldg(reinterpret_cast<const uint4*>(x.p + (q00 + 1) * x.cs + 8 * g));
        r[2] = __ldg(reinterpret_cast<const uint4*>(x.p + (q00 + W) * x.cs + 8 * g));
        r[3] = __ldg(reinterpret_cast<const uint4*>(x.p + (q00 + W + 1) * x.cs + 8 * g));
        float best[8], v[8];
        uint32_t bi[2] = {0u, 0u};
        unpack8f(r[0], best);
        if (AFF) {
#pragma unroll
          for (int j = 0; j < 8; ++j) best[j] *= sg[j];
        }
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          unpack8f(r[k], v);
          if (AFF) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= sg[j];
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (v[j] > best[j]) { best[j] = v[j]; bi[j >> 2] = (bi[j >> 2] & ~(0xffu << ((j & 3) * 8))) | ((uint32_t)k << ((j & 3) * 8)); }
        }
        if (AFF) {       // affine of the selected element, rounded like the stored tensor (statistics are taken on stored values)
#pragma unroll
          for (int j = 0; j < 8; ++j) best[j] = __bfloat162float(__float2bfloat16_rn(fmaf(sg[j] * best[j], sc[j], sh[j])));
        }
        __stcs(reinterpret_cast<uint4*>(y + p * ycs + 8 * g),
               make_uint4(pack2(best[0], best[1]), pack2(best[2], best[3]), pack2(best[4], best[5]), pack2(best[6], best[7])));
        if (idx) *reinterpret_cast<uint2*>(idx + p * C + 8 * g) = make_uint2(bi[0], bi[1]);
        if (STATS) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { s0[j] += best[j]; s1[j] = fmaf(best[j], best[j], s1[j]); }
          if (++cnt == 64) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; s0[j] = s1[j] = 0.f; }
            cnt = 0;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { d0[j] += s0[j]; d1[j] += s1[j]; }
    }
    if (STATS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { sm[threadIdx.x * 16 + j] = d0[j]; sm[threadIdx.x * 16 + 8 + j] = d1[j]; }
      __syncthreads();
      for (int o = threadIdx.x; o < 16 * GL; o += 256) {
        const int q = o & 15, gg = o >> 4;
        double rsum = 0.0;
        for (int l = 0; l < PL; ++l) rsum += sm[(l * GL + gg) * 16 + q];
        const int gch = g - gl + gg;
        if (gch < ng) atomicAdd(stats + (q < 8 ? 0 : C) + 8 * gch + (q & 7), rsum);
      }
      __syncthreads();
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) maxpool_bwd_vec8_kernel(PV dy, const uint8_t* __restrict__ idx, PV dskip, int has_skip,
                                                              __nv_bfloat16* dx, long long dxcs, PV mask, float alpha, int C, int Ho,
                                                              int Wo, long long total) {
  const int ng = C / 8;
  const long long W = 2LL * Wo;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long p = t / ng;
    const int g = (int)(t - p * ng);
    const int ox = (int)(p % Wo);
    const long long u = p / Wo;
    const int oy = (int)(u % Ho);
    const long long n = u / Ho;
    const long long q00 = (n * 2 * Ho + 2 * oy) * W + 2 * ox;
    float gv[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(dy.p + p * dy.cs + 8 * g)), gv);
    const uint2 iv = __ldg(reinterpret_cast<const uint2*>(idx + p * C + 8 * g));
    const uint32_t ib[2] = {iv.x, iv.y};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long q = q00 + (k >> 1) * W + (k & 1);
      float f[8];
      if (has_skip) unpack8f(*reinterpret_cast<const uint4*>(dskip.p + q * dskip.cs + 8 * g), f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = ((((ib[j >> 2] >> ((j & 3) * 8)) & 0xffu) == (uint32_t)k) ? gv[j] : 0.f) + f[j];
      if (MODE != 0) {
        float m[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(mask.p + q * mask.cs + 8 * g)), m);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= (m[j] > 0.f ? 1.f : (MODE == 1 ? 0.f : alpha));
      }
      __stcs(reinterpret_cast<uint4*>(dx + q * dxcs + 8 * g),
             make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7])));
    }
  }
}

static bool vec8_view(const dnnca_tensor_t* t) {
  return t->dtype == DNNCA_BF16 && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}
static PV pv(const dnnca_tensor_t* t) {
  return PV{reinterpret_cast<const __nv_bfloat16*>(t->data) + t->coff, (long long)t->cstride};
}

static bool dense_bf16(const dnnca_tensor_t* t) {
  return t->dtype == DNNCA_BF16 && t->coff == 0 && t->cstride == t->c && (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}

static int vec_grid(long long threads) {
  long long b = (threads + 255) / 256, cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

// MaxPool of a tensor with a folded BatchNorm affine; 1 handled / 0 not covered
int try_maxpool_fwd_affine_vec(cudaStream_t s, const dnnca_tensor_t* x, const float* aff, const dnnca_tensor_t* y, uint8_t* idx,
                               double* stats) {
  if (!(vec8_view(x) && vec8_view(y) && (!idx || (reinterpret_cast<uintptr_t>(idx) & 7) == 0))) return 0;
  const int C = x->c;
  int gl = 1;
  while (gl < C / 8 && gl < 256) gl <<= 1;
  const int pl = 256 / gl;
  const long long PO = (long long)y->n * y->h * y->w;
  long long b = (PO + (long long)pl * (stats ? 32 : 4) - 1) / ((long long)pl * (stats ? 32 : 4)), cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y->data) + y->coff;
  if (stats) maxpool_fwd_vec8_kernel<true, true><<<(int)b, 256, 0, s>>>(pv(x), yp, y->cstride, idx, stats, C, y->h, y->w, PO, gl, pl, aff);
  else maxpool_fwd_vec8_kernel<false, true><<<(int)b, 256, 0, s>>>(pv(x), yp, y->cstride, idx, stats, C, y->h, y->w, PO, gl, pl, aff);
  DNNCA_LAUNCH_CHECK("maxpool_fwd_affine_vec8");
  return 1;
}

// returns 1 when handled, 0 when the shape is not covered
int try_maxpool_fwd_vec(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* y, uint8_t* idx, double* stats) {
  if (vec8_view(x) && vec8_view(y) && (!idx || (reinterpret_cast<uintptr_t>(idx) & 7) == 0)) {
    const int C = x->c;
    int gl = 1;
    while (gl < C / 8 && gl < 256) gl <<= 1;
    const int pl = 256 / gl;
    const long long PO = (long long)y->n * y->h * y->w;
    long long b = (PO + (long long)pl * (stats ? 32 : 4) - 1) / ((long long)pl * (stats ? 32 : 4)), cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y->data) + y->coff;
    if (stats) maxpool_fwd_vec8_kernel<true><<<(int)b, 256, 0, s>>>(pv(x), yp, y->cstride, idx, stats, C, y->h, y->w, PO, gl, pl);
    else maxpool_fwd_vec8_kernel<false><<<(int)b, 256, 0, s>>>(pv(x), yp, y->cstride, idx, stats, C, y->h, y->w, PO, gl, pl);
    DNNCA_LAUNCH_CHECK("maxpool_fwd_vec8");
    return 1;
  }
  if (stats || !dense_bf16(x) || !dense_bf16(y)) return 0;
  const int C = x->c;
  if ((C != 3 && C != 6 && C != 12) || ((long long)y->w * C) % 24) return 0;
  if (idx && (reinterpret_cast<uintptr_t>(idx) & 7)) return 0;
  const int groups = y->w * C / 24;
  const long long total = (long long)y->n * y->h * groups;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x->data);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y->data);
  const int grid = vec_grid(total);
  if (C == 3) maxpool_fwd_vec_kernel<3><<<grid, 256, 0, s>>>(xp, yp, idx, groups, total);
  else if (C == 6) maxpool_fwd_vec_kernel<6><<<grid, 256, 0, s>>>(xp, yp, idx, groups, total);
  else maxpool_fwd_vec_kernel<12><<<grid, 256, 0, s>>>(xp, yp, idx, groups, total);
  DNNCA_LAUNCH_CHECK("maxpool_fwd_vec");
  return 1;
}

template <int C>
static void launch_pool_bwd(cudaStream_t s, int grid, int mode, const __nv_bfloat16* dy, const uint8_t* idx,
                            const __nv_bfloat16* dskip, __nv_bfloat16* dx, const __nv_bfloat16* mask, float alpha, int groups,
                            long long total) {
  if (mode == 0) maxpool_bwd_vec_kernel<C, 0><<<grid, 256, 0, s>>>(dy, idx, dskip, dx, mask, alpha, groups, total);
  else if (mode == 1) maxpool_bwd_vec_kernel<C, 1><<<grid, 256, 0, s>>>(dy, idx, dskip, dx, mask, alpha, groups, total);
  else maxpool_bwd_vec_kernel<C, 2><<<grid, 256, 0, s>>>(dy, idx, dskip, dx, mask, alpha, groups, total);
}

int try_maxpool_bwd_vec(cudaStream_t s, const dnnca_tensor_t* dy, const uint8_t* idx, const dnnca_tensor_t* dskip,
                        const dnnca_tensor_t* dx, const dnnca_tensor_t* mask, int act, float alpha) {
  if (vec8_view(dy) && vec8_view(dx) && (!dskip || vec8_view(dskip)) && (!mask || vec8_view(mask)) &&
      (reinterpret_cast<uintptr_t>(idx) & 7) == 0) {
    const int C = dy->c;
    const int mode8 = (!mask || act == DNNCA_ACT_NONE) ? 0 : (act == DNNCA_ACT_RELU ? 1 : 2);
    const long long total = (long long)dy->n * dy->h * dy->w * (C / 8);
    const int grid = vec_grid(total);
    const PV sk = dskip ? pv(dskip) : pv(dx), mk8 = mask ? pv(mask) : pv(dx);
    __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx->data) + dx->coff;
    if (mode8 == 0) maxpool_bwd_vec8_kernel<0><<<grid, 256, 0, s>>>(pv(dy), idx, sk, dskip != nullptr, dxp, dx->cstride, mk8, alpha, C, dy->h, dy->w, total);
    else if (mode8 == 1) maxpool_bwd_vec8_kernel<1><<<grid, 256, 0, s>>>(pv(dy), idx, sk, dskip != nullptr, dxp, dx->cstride, mk8, alpha, C, dy->h, dy->w, total);
    else maxpool_bwd_vec8_kernel<2><<<grid, 256, 0, s>>>(pv(dy), idx, sk, dskip != nullptr, dxp, dx->cstride, mk8, alpha, C, dy->h, dy->w, total);
    DNNCA_LAUNCH_CHECK("maxpool_bwd_vec8");
    return 1;
  }
  if (!dense_bf16(dy) || !dense_bf16(dx) || (dskip && !dense_bf16(dskip)) || (mask && !dense_bf16(mask))) return 0;
  const int C = dy->c;
  if ((C != 3 && C != 6 && C != 12) || ((long long)dy->w * C) % 24 || (reinterpret_cast<uintptr_t>(idx) & 7)) return 0;
  const int mode = (!mask || act == DNNCA_ACT_NONE) ? 0 : (act == DNNCA_ACT_RELU ? 1 : 2);
  const int groups = dy->w * C / 24;
  const long long total = (long long)dy->n * dy->h * groups;
  const int grid = vec_grid(total);
  const __nv_bfloat16* dyp = reinterpret_cast<const __nv_bfloat16*>(dy->data);
  const __nv_bfloat16* skp = dskip ? reinterpret_cast<const __nv_bfloat16*>(dskip->data) : nullptr;
  const __nv_bfloat16* mp = mask ? reinterpret_cast<const __nv_bfloat16*>(mask->data) : nullptr;
  __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(dx->data);
  if (C == 3) launch_pool_bwd<3>(s, grid, mode, dyp, idx, skp, dxp, mp, alpha, groups, total);
  else if (C == 6) launch_pool_bwd<6>(s, grid, mode, dyp, idx, skp, dxp, mp, alpha, groups, total);
  else launch_pool_bwd<12>(s, grid, mode, dyp, idx, skp, dxp, mp, alpha, groups, total);
  DNNCA_LAUNCH_CHECK("maxpool_bwd_vec");
  return 1;
}

int try_convert_vec(cudaStream_t s, const dnnca_tensor_t* src, const dnnca_tensor_t* dst) {
  if (src->dtype != DNNCA_F32 || !dense_bf16(dst) || src->coff != 0 || src->cstride != src->c ||
      (reinterpret_cast<uintptr_t>(src->data) & 15))
    return 0;
  const long long count = (long long)src->n * src->h * src->w * src->c;
  if (count % 8) return 0;
  f32_to_bf16_vec_kernel<<<vec_grid(count / 8), 256, 0, s>>>(reinterpret_cast<const float*>(src->data),
                                                           reinterpret_cast<__nv_bfloat16*>(dst->data), count / 8);
  DNNCA_LAUNCH_CHECK("convert_vec");
  return 1;
}

}  // namespace dnnca
