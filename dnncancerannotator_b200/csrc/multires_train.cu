// MultiResUnet training path: the elementwise / gather kernels the backward pass of
// multiresunet.py:31-60 (conv2d_bn), :89-126 (MultiResBlock), :129-164 (ResPath) and :219 (conv10) needs beyond
// the conv / BatchNorm / pool kernels shared with the U-Nets.
//
//   dnnca_bn_apply_act   y = act(x*scale + shift)                 Conv2D -> BN(scale=False) -> Activation('relu')
//   dnnca_act_bwd        dx = dy * act'(y)                        gradient through an Activation that FOLLOWS a BN / add
//   dnnca_accumulate     dst += src                               tensors with several consumers (block input -> shortcut
//                                                                  conv and 3x3 chain; conv3x3 -> conv5x5 and the concat)
//   dnnca_gather_f32     dst[i] = idx[i] >= 0 ? src[idx[i]] : 0   variables <-> their channel-padded (physical) layout
//   dnnca_head_conv_bwd  df = dz*w*act'(f), dw += sum dz*f        1x1 conv to ONE channel whose output is fp32 (conv10)
//
// All kernels are HBM-bound; they use the (pixel-lane, channel-lane) layout of common.cuh, with a 16-byte path for bf16
// views whose channel geometry is a multiple of 8 (every tensor of the training plan: segments are padded to 16).
#include "common.cuh"

namespace dnnca {

namespace {

template <typename T>
__device__ __forceinline__ const T* pxr(const View& v, long long p) {
  return reinterpret_cast<const T*>(v.data) + p * v.cstride + v.coff;
}
template <typename T>
__device__ __forceinline__ T* pxw(const View& v, long long p) {
  return reinterpret_cast<T*>(v.data) + p * v.cstride + v.coff;
}

inline bool vec8(const dnnca_tensor_t* t) {
  return t->dtype == DNNCA_BF16 && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}

__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// MODE 0: y = act(x*scale+shift)   (a = x, p0 = scale|shift)
// MODE 1: dx = dy * act'(y)        (a = y, b = dy)
// MODE 2: dst += src               (a = src, b = dst == out)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) map_kernel(View a, View b, const float* __restrict__ p0, View out, int act, float alpha,
                                                  int CL, int PL, long long P) {
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < a.c; c += CL) {
    float s = 1.f, t = 0.f;
    if (MODE == 0) { s = p0[c]; t = p0[a.c + c]; }
    for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL) {
      const float av = ldf(pxr<T>(a, p) + c);
      float r;
      if (MODE == 0) r = apply_act(av * s + t, act, alpha);
      else if (MODE == 1) r = ldf(pxr<T>(b, p) + c) * act_grad(av, act, alpha);
      else r = av + ldf(pxr<T>(b, p) + c);
      stf(pxw<T>(out, p) + c, r);
    }
  }
}

// bf16, 8 channels (16 bytes) per thread and step; G = c/8 groups per pixel
template <int MODE>
__global__ void __launch_bounds__(256) map_vec8_kernel2(View a, View b, const float* __restrict__ p0, View out, int act,
                                                        float alpha, long long P) {
  const int G = a.c >> 3;
  const long long total = P * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / G;
    const int c0 = (int)(i - p * G) << 3;
    float av[8], bv[8], r[8];
    unpack8(*reinterpret_cast<const uint4*>(pxr<__nv_bfloat16>(a, p) + c0), av);
    if (MODE != 0) unpack8(*reinterpret_cast<const uint4*>(pxr<__nv_bfloat16>(b, p) + c0), bv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) r[j] = apply_act(av[j] * p0[c0 + j] + p0[a.c + c0 + j], act, alpha);
      else if (MODE == 1) r[j] = bv[j] * act_grad(av[j], act, alpha);
      else r[j] = av[j] + bv[j];
    }
    *reinterpret_cast<uint4*>(pxw<__nv_bfloat16>(out, p) + c0) = pack8(r);
  }
}

template <int MODE>
int launch_map(cudaStream_t s, const dnnca_tensor_t* a, const dnnca_tensor_t* b, const float* p0, const dnnca_tensor_t* out,
               int act, float alpha, const char* what) {
  const long long P = (long long)a->n * a->h * a->w;
  if (vec8(a) && (!b || vec8(b)) && vec8(out)) {
    map_vec8_kernel2<MODE><<<grid_for(P * (a->c / 8), 256 * 4, 8), 256, 0, s>>>(mk(a), mk(b ? b : a), p0, mk(out), act, alpha, P);
    DNNCA_LAUNCH_CHECK(what);
    return DNNCA_OK;
  }
  ChanLayout L = chan_layout(a->c);
  const int grid = grid_for(P, L.pl * 4);
  DNNCA_DISPATCH_DTYPE(a->dtype, (map_kernel<T, MODE><<<grid, 256, 0, s>>>(mk(a), mk(b ? b : a), p0, mk(out), act, alpha, L.cl, L.pl, P));)
  DNNCA_LAUNCH_CHECK(what);
  return DNNCA_OK;
}

__global__ void __launch_bounds__(256) gather_f32_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx,
                                                         long long n, float* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    dst[i] = j >= 0 ? src[j] : 0.f;
  }
}

// df[p,c] = dz[p] * w[c] * act'(f[p,c]);  dw[c] += sum_p dz[p] * f[p,c].  One thread per pixel, the channel loop runs
// over the (few, <= 256) head features; per-channel sums: warp shuffle -> shared-memory atomics -> one global atomic per
// channel and block.
template <typename T>
__global__ void __launch_bounds__(256) head_conv_bwd_kernel(View f, const float* __restrict__ w, const float* __restrict__ dz,
                                                            View df, int has_df, int act, float alpha, float* __restrict__ dw,
                                                            long long P) {
  extern __shared__ float red[];      // [f.c]
  for (int c = threadIdx.x; c < f.c; c += blockDim.x) red[c] = 0.f;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long rounds = (P + stride - 1) / stride;            // block-uniform trip count (shuffles inside)
  for (long long k = 0; k < rounds; ++k) {
    const long long p = k * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < P;
    const float g = live ? dz[p] : 0.f;
    for (int c = 0; c < f.c; ++c) {
      const float fv = live ? ldf(pxr<T>(f, p) + c) : 0.f;
      if (live && has_df) stf(pxw<T>(df, p) + c, g * w[c] * act_grad(fv, act, alpha));
      const float s = warp_sum(g * fv);
      if ((threadIdx.x & 31) == 0) atomicAdd(red + c, s);
    }
  }
  __syncthreads();
  if (dw)
    for (int c = threadIdx.x; c < f.c; c += blockDim.x) atomicAdd(dw + c, red[c]);
}

}  // namespace
}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_bn_apply_act(void* stream, const dnnca_tensor_t* x, const float* scale_shift, const dnnca_tensor_t* y,
                                  int act, float alpha) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && same_shape(x, y) && scale_shift && x->dtype == y->dtype, "bn_apply_act: bad arguments");
  return launch_map<0>((cudaStream_t)stream, x, nullptr, scale_shift, y, act, alpha, "bn_apply_act");
}

extern "C" int dnnca_act_bwd(void* stream, const dnnca_tensor_t* y, const dnnca_tensor_t* dy, const dnnca_tensor_t* dx,
                             int act, float alpha) {
  DNNCA_CHECK_ARG(view_ok(y) && view_ok(dy) && view_ok(dx) && same_shape(y, dy) && same_shape(y, dx) && y->dtype == dy->dtype &&
                      y->dtype == dx->dtype, "act_bwd: bad arguments");
  return launch_map<1>((cudaStream_t)stream, y, dy, nullptr, dx, act, alpha, "act_bwd");
}

extern "C" int dnnca_accumulate(void* stream, const dnnca_tensor_t* src, const dnnca_tensor_t* dst) {
  DNNCA_CHECK_ARG(view_ok(src) && view_ok(dst) && same_shape(src, dst) && src->dtype == dst->dtype, "accumulate: bad arguments");
  return launch_map<2>((cudaStream_t)stream, src, dst, nullptr, dst, 0, 0.f, "accumulate");
}

extern "C" int dnnca_gather_f32(void* stream, const float* src, const int32_t* idx, int64_t count, float* dst) {
  DNNCA_CHECK_ARG(src && idx && dst && count > 0, "gather_f32: bad arguments");
  gather_f32_kernel<<<grid_for(count, 256 * 4, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, count, dst);
  DNNCA_LAUNCH_CHECK("gather_f32");
  return DNNCA_OK;
}

extern "C" int dnnca_head_conv_bwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* dz,
                                   const dnnca_tensor_t* df, int act, float alpha, float* dw) {
  DNNCA_CHECK_ARG(view_ok(f) && w && dz && (df || dw), "head_conv_bwd: bad arguments");
  DNNCA_CHECK_ARG(!df || (view_ok(df) && same_shape(df, f) && df->dtype == f->dtype), "head_conv_bwd: bad df");
  if (f->c > 1024) DNNCA_UNSUPPORTED("head_conv_bwd: at most 1024 features (got %d)", f->c);
  const long long P = (long long)f->n * f->h * f->w;
  const int grid = grid_for(P, 256 * 4, 4);
  View vdf = df ? mk(df) : mk(f);
  DNNCA_DISPATCH_DTYPE(f->dtype, (head_conv_bwd_kernel<T><<<grid, 256, (size_t)f->c * 4, (cudaStream_t)stream>>>(
                                     mk(f), w, dz, vdf, df != nullptr, act, alpha, dw, P));)
  DNNCA_LAUNCH_CHECK("head_conv_bwd");
  return DNNCA_OK;
}
