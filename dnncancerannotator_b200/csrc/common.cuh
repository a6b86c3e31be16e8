// Shared device/host helpers for libdnnca (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dnnca.h"

namespace dnnca {

// ---- error plumbing (api.cu owns the thread-local buffer) -------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define DNNCA_CHECK_ARG(cond, ...)              \
  do {                                          \
    if (!(cond)) {                              \
      ::dnnca::set_error(__VA_ARGS__);          \
      return DNNCA_ERR_BAD_ARG;                 \
    }                                           \
  } while (0)

#define DNNCA_UNSUPPORTED(...)                  \
  do {                                          \
    ::dnnca::set_error(__VA_ARGS__);            \
    return DNNCA_ERR_UNSUPPORTED;               \
  } while (0)

void note_launch(int n);
void note_family(int f);   // 0 generic, 1 small-channel TMA/FFMA2, 2 tcgen05 (conv/ConvT kernels only)

// placed after every kernel launch: counts it and surfaces launch errors
#define DNNCA_LAUNCH_CHECK(what)                                 \
  do {                                                           \
    ::dnnca::note_launch(1);                                     \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return ::dnnca::cuda_fail(e__, what); \
  } while (0)

int sm_count();

// ---- tensor views ----------------------------------------------------------
struct View {
  void* data;
  int n, h, w, c, cstride, coff, dtype;
  __host__ __device__ long long pixels() const { return (long long)n * h * w; }
};

inline View mk(const dnnca_tensor_t* t) {
  View v;
  v.data = t->data; v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c;
  v.cstride = t->cstride; v.coff = t->coff; v.dtype = t->dtype;
  return v;
}

inline bool view_ok(const dnnca_tensor_t* t) {
  return t && t->data && t->n > 0 && t->h > 0 && t->w > 0 && t->c > 0 && t->coff >= 0 &&
         t->cstride >= t->coff + t->c && (t->dtype == DNNCA_F32 || t->dtype == DNNCA_BF16);
}
inline bool same_shape(const dnnca_tensor_t* a, const dnnca_tensor_t* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}
inline bool same_nhw(const dnnca_tensor_t* a, const dnnca_tensor_t* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w;
}

// ---- scalar load/store in fp32 ---------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// value as it will read back after storing in T (stats are taken on stored values)
template <typename T> __device__ __forceinline__ float rnd(float v);
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }
template <> __device__ __forceinline__ float rnd<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == DNNCA_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == DNNCA_ACT_LEAKY) return v > 0.f ? v : alpha * v;
  return v;
}
// derivative from the stored output (ReLU / LeakyReLU preserve sign)
__device__ __forceinline__ float act_grad(float y, int act, float alpha) {
  if (act == DNNCA_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == DNNCA_ACT_LEAKY) return y > 0.f ? 1.f : alpha;
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- (pixel-lane, channel-lane) thread layout for per-channel NHWC kernels --
// 256 threads = PL pixel lanes x CL channel lanes, CL = min(pow2ceil(C), 256).
// Thread (pl, cl) owns channels cl, cl+CL, ... and strides over pixels by PL, so
// a warp touches contiguous memory and per-channel parameters stay in registers.
struct ChanLayout {
  int cl, pl;
};
inline ChanLayout chan_layout(int c) {
  int cl = 1;
  while (cl < c && cl < 256) cl <<= 1;
  return ChanLayout{cl, 256 / cl};
}

inline int grid_for(long long work_items, int per_block, int waves = 8) {
  long long b = (work_items + per_block - 1) / per_block;
  long long cap = (long long)sm_count() * waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define DNNCA_DISPATCH_DTYPE(dt, ...)                 \
  if ((dt) == DNNCA_F32) {                            \
    using T = float;                                  \
    __VA_ARGS__                                       \
  } else {                                            \
    using T = __nv_bfloat16;                          \
    __VA_ARGS__                                       \
  }

}  // namespace dnnca
