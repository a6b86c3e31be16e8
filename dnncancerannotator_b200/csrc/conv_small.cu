// Direct 3x3 convolution for tiny channel counts (configs/unet.yaml: 3/6/12/24
// channels; first layers of every model: Cin = 1/3/5).  These layers sit far
// below the tensor-core ridge (4-72 flop/B vs ~258), so they run on the FP32
// pipe from shared-memory tiles instead of being padded into a GEMM:
//   * the input tile (+1-pixel halo = the conv's 'same' zero padding) is staged
//     once per CTA as fp32 planes xs[ci][row][col]; every thread owns 4 adjacent
//     pixels x ALL output channels, so one 6-float window (LDS.128 + LDS.64) and
//     3*COUT broadcast weights (LDS.128) feed 12*COUT FMAs;
//   * fprop and dgrad are the same kernel: dgrad stages rot180/transposed
//     weights and swaps the bias+activation epilogue for the act'(mask) product;
//   * outputs go back through shared memory so global stores are coalesced
//     along NHWC rows even for 3-channel (6-byte) pixels; BatchNorm statistics
//     (sum, sum of squares of the stored values) are reduced per CTA -> fp64 atomics;
//   * wgrad: thread = (4-pixel group, ci); keeps the 3x6 input window of its
//     channel in registers and streams dz (LDS.128): 36 FMAs per load,
//     9*COUT register accumulators, one smem/atomic reduction per CTA.
// Reference call sites: layers.Conv2D components.py:47-50,123-126 and their
// tf.GradientTape gradients.
//
// Compiled once per (dtype, kind) with -DSMALL_DT=0|1 (f32|bf16) and -DSMALL_KIND=0|1|2
// (fprop|dgrad|wgrad) so the template instantiations build in parallel.
#include "common.cuh"

#ifndef SMALL_DT
#error "compile with -DSMALL_DT=0|1 -DSMALL_KIND=0|1|2"
#endif
#if SMALL_DT == 0
#define SMALL_T float
#define SMALL_FN(name) name##_f32
#else
#define SMALL_T __nv_bfloat16
#define SMALL_FN(name) name##_bf16
#endif

namespace dnnca {

constexpr int PX = 4;  // pixels per thread along x

template <int COUT> struct CoPad { static constexpr int v = (COUT + 3) / 4 * 4; };

// ----------------------------------------------------------------------------
// shared staging helpers
// ----------------------------------------------------------------------------
// xs[ck][rows][pitch] <- channels [c0, c0+ck) of the (rows x cols) window whose top-left
// image coordinate is (gy0, gx0); out-of-image -> 0.  Global order (row, col, ci) keeps
// the reads contiguous along NHWC rows.
template <typename T>
__device__ __forceinline__ void stage_planes(float* __restrict__ xs, const View& v, long long n, int gy0, int gx0,
                                             int rows, int cols, int pitch, int c0, int ck) {
  const int total = rows * cols * ck;
  const T* base = reinterpret_cast<const T*>(v.data) + v.coff + c0;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int ci = e % ck;
    const int t = e / ck;
    const int col = t % cols, row = t / cols;
    const int gy = gy0 + row, gx = gx0 + col;
    float val = 0.f;
    if (gy >= 0 && gy < v.h && gx >= 0 && gx < v.w)
      val = ldf(base + ((n * v.h + gy) * v.w + gx) * (long long)v.cstride + ci);
    xs[(ci * rows + row) * pitch + col] = val;
  }
}

// ----------------------------------------------------------------------------
// fprop / dgrad
// ----------------------------------------------------------------------------
template <typename T, int CIN, int COUT, int TXN, bool DGRAD>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(View x, const float* __restrict__ w,
                                                           const float* __restrict__ bias, View y, View mask,
                                                           int has_mask, int act, float alpha,
                                                           double* __restrict__ stats, int tiles_x, int tiles_y) {
  constexpr int TYN = 256 / TXN;
  constexpr int TW = TXN * PX;
  constexpr int CK = CIN < 12 ? CIN : 12;           // channels per smem stage
  constexpr int NST = (CIN + CK - 1) / CK;
  constexpr int ROWS = TYN + 2, COLS = TW + 2, PITCH = TW + 4;
  constexpr int COP = CoPad<COUT>::v;
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                                  // [CIN][3][3][COP]
  float* xs = smem + CIN * 9 * COP;                  // [CK][ROWS][PITCH]   (re-used as the output tile)

  int b = blockIdx.x;
  const int tix = b % tiles_x; b /= tiles_x;
  const int tiy = b % tiles_y;
  const long long n = b / tiles_y;
  const int x0 = tix * TW, y0 = tiy * TYN;
  const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;

  // weights -> smem in [ci][dy][dx][co] order (dgrad: rot180 + in/out transpose)
  for (int e = threadIdx.x; e < CIN * 9 * COP; e += 256) {
    const int co = e % COP;
    int t = e / COP;
    const int dx = t % 3; t /= 3;
    const int dy = t % 3;
    const int ci = t / 3;
    float v = 0.f;
    if (co < COUT) {
      if (!DGRAD) v = w[((dy * 3 + dx) * CIN + ci) * COUT + co];
      else        v = w[(((2 - dy) * 3 + (2 - dx)) * COUT + co) * CIN + ci];  // layer Cin = COUT, layer Cout = CIN
    }
    ws[e] = v;
  }

  float acc[COUT][PX];
#pragma unroll
  for (int co = 0; co < COUT; ++co)
#pragma unroll
    for (int p = 0; p < PX; ++p) acc[co][p] = 0.f;

#pragma unroll 1
  for (int st = 0; st < NST; ++st) {
    const int c0 = st * CK;
    const int ck = (CIN - c0) < CK ? (CIN - c0) : CK;
    if (st > 0) __syncthreads();
    stage_planes<T>(xs, x, n, y0 - 1, x0 - 1, ROWS, COLS, PITCH, c0, ck);
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < ck; ++ci) {
      const float* wrow = ws + (c0 + ci) * 9 * COP;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* xr = xs + (ci * ROWS + ty + dy) * PITCH + tx * PX;
        const float4 w0 = *reinterpret_cast<const float4*>(xr);
        const float2 w1 = *reinterpret_cast<const float2*>(xr + 4);
        const float win[6] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y};
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = wrow + (dy * 3 + dx) * COP;
#pragma unroll
          for (int c4 = 0; c4 < COP / 4; ++c4) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + c4 * 4);
            const float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int co = c4 * 4 + j;
              if (co < COUT) {
#pragma unroll
                for (int p = 0; p < PX; ++p) acc[co][p] = fmaf(win[p + dx], wa[j], acc[co][p]);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();  // everyone is done reading xs -> reuse it as the output tile

  // epilogue into smem tile os[row][col][co] (T), then coalesced row stores
  T* os = reinterpret_cast<T*>(xs);
  const int gy = y0 + ty;
  float ssum[COUT], ssq[COUT];
#pragma unroll
  for (int co = 0; co < COUT; ++co) { ssum[co] = 0.f; ssq[co] = 0.f; }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int gx = x0 + tx * PX + p;
    const bool inside = gy < y.h && gx < y.w;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float v = acc[co][p];
      if (!DGRAD) {
        v = apply_act(v + (bias ? bias[co] : 0.f), act, alpha);
      } else if (has_mask && inside) {
        const T* mp = reinterpret_cast<const T*>(mask.data) + ((n * mask.h + gy) * mask.w + gx) * (long long)mask.cstride + mask.coff;
        v *= act_grad(ldf(mp + co), act, alpha);
      }
      const float r = rnd<T>(v);
      if (inside) { ssum[co] += r; ssq[co] += r * r; }
      stf(os + ((ty * TW) + tx * PX + p) * COUT + co, v);
    }
  }
  __syncthreads();
  {
    // cooperative store: tile row = contiguous TW*COUT elements when the view is dense
    const int wv = min(TW, y.w - x0);                    // valid columns
    const int hv = min(TYN, y.h - y0);
    T* ybase = reinterpret_cast<T*>(y.data) + y.coff;
    const int row_elems = wv * COUT;
    for (int e = threadIdx.x; e < hv * row_elems; e += 256) {
      const int row = e / row_elems, r = e % row_elems;
      const int col = r / COUT, co = r % COUT;
      ybase[((n * y.h + y0 + row) * y.w + x0 + col) * (long long)y.cstride + co] = os[(row * TW + col) * COUT + co];
    }
  }
  if (!DGRAD && stats) {
    __shared__ float red[8][2 * COUT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const float s = warp_sum(ssum[co]), q = warp_sum(ssq[co]);
      if (lane == 0) { red[warp][co] = s; red[warp][COUT + co] = q; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * COUT) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
      atomicAdd(stats + threadIdx.x, s);
    }
  }
}

// ----------------------------------------------------------------------------
// wgrad: dw[a][c][ci][co] += sum_p x[p+(a-1,c-1)][ci] * dz[p][co] ; db[co] += sum_p dz[p][co]
// persistent CTAs loop over tiles; thread = (pixel-group slot, ci)
// ----------------------------------------------------------------------------
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(256) conv3x3_small_wgrad_kernel(View x, View dz, float* __restrict__ dw,
                                                                 float* __restrict__ db, int tiles_x, int tiles_y,
                                                                 long long ntiles) {
  constexpr int TYN = 4, TXN = 16, TW = TXN * PX;      // 4 x 64 pixel tiles
  constexpr int ROWS = TYN + 2, COLS = TW + 2, PITCH = TW + 4;
  constexpr int COP = CoPad<COUT>::v;
  constexpr int G = 256 / CIN;                          // pixel-group slots processed concurrently
  constexpr int NGRP = TYN * TXN;                       // 4-pixel groups per tile
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                                     // [CIN][ROWS][PITCH]
  float* gs = smem + CIN * ROWS * PITCH;                // [COUT][TYN][TW]
  float* red = gs + COUT * TYN * TW;                    // [9*CIN*COUT + COUT] block accumulators

  const int slot = threadIdx.x / CIN, ci = threadIdx.x % CIN;
  const bool active = slot < G;

  float acc[9][COUT];
  float dbacc[COUT];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[t][co] = 0.f;
#pragma unroll
  for (int co = 0; co < COUT; ++co) dbacc[co] = 0.f;

  for (int e = threadIdx.x; e < 9 * CIN * COUT + COUT; e += 256) red[e] = 0.f;

#pragma unroll 1
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long b = tile;
    const int tix = (int)(b % tiles_x); b /= tiles_x;
    const int tiy = (int)(b % tiles_y);
    const long long n = b / tiles_y;
    const int x0 = tix * TW, y0 = tiy * TYN;
    __syncthreads();
    stage_planes<T>(xs, x, n, y0 - 1, x0 - 1, ROWS, COLS, PITCH, 0, CIN);
    stage_planes<T>(gs, dz, n, y0, x0, TYN, TW, TW, 0, COUT);
    __syncthreads();
    if (active) {
#pragma unroll 1
      for (int g = slot; g < NGRP; g += G) {
        const int ty = g / TXN, tx = g % TXN;
        float win[3][6];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float* xr = xs + (ci * ROWS + ty + dy) * PITCH + tx * PX;
          const float4 a = *reinterpret_cast<const float4*>(xr);
          const float2 c = *reinterpret_cast<const float2*>(xr + 4);
          win[dy][0] = a.x; win[dy][1] = a.y; win[dy][2] = a.z; win[dy][3] = a.w; win[dy][4] = c.x; win[dy][5] = c.y;
        }
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float4 gv = *reinterpret_cast<const float4*>(gs + (co * TYN + ty) * TW + tx * PX);
          const float ga[4] = {gv.x, gv.y, gv.z, gv.w};
          if (ci == 0) dbacc[co] += (ga[0] + ga[1]) + (ga[2] + ga[3]);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int p = 0; p < PX; ++p) acc[dy * 3 + dx][co] = fmaf(win[dy][p + dx], ga[p], acc[dy * 3 + dx][co]);
        }
      }
    }
  }
  __syncthreads();
  // CTA reduction in shared memory, then one global atomic per output per CTA
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int co = 0; co < COUT; ++co) atomicAdd(red + (t * CIN + ci) * COUT + co, acc[t][co]);
    if (ci == 0) {
#pragma unroll
      for (int co = 0; co < COUT; ++co) atomicAdd(red + 9 * CIN * COUT + co, dbacc[co]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * CIN * COUT; e += 256) atomicAdd(dw + e, red[e]);
  if (db)
    for (int e = threadIdx.x; e < COUT; e += 256) atomicAdd(db + e, red[9 * CIN * COUT + e]);
  (void)COP;
}

// ----------------------------------------------------------------------------
// host dispatch
// ----------------------------------------------------------------------------
template <typename T, int CIN, int COUT, int TXN, bool DGRAD>
static int launch_small(cudaStream_t s, const dnnca_tensor_t* x, const float* w, const float* bias,
                        const dnnca_tensor_t* y, const dnnca_tensor_t* mask, int act, float alpha, double* stats) {
  constexpr int TYN = 256 / TXN, TW = TXN * PX;
  constexpr int CK = CIN < 12 ? CIN : 12;
  constexpr int COP = CoPad<COUT>::v;
  constexpr size_t xs_bytes = (size_t)CK * (TYN + 2) * (TW + 4) * 4;
  constexpr size_t os_bytes = (size_t)TYN * TW * COUT * sizeof(T);
  constexpr size_t smem = (size_t)CIN * 9 * COP * 4 + (xs_bytes > os_bytes ? xs_bytes : os_bytes);
  auto kern = conv3x3_small_kernel<T, CIN, COUT, TXN, DGRAD>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv3x3_small: cudaFuncSetAttribute");
    attr_done = true;
  }
  const int tiles_x = (x->w + TW - 1) / TW, tiles_y = (x->h + TYN - 1) / TYN;
  const long long nblk = (long long)tiles_x * tiles_y * x->n;
  if (nblk > 0x7fffffffLL) { set_error("conv3x3_small: grid too large"); return DNNCA_ERR_UNSUPPORTED; }
  View vm = mask ? mk(mask) : mk(y);
  kern<<<(unsigned)nblk, 256, smem, s>>>(mk(x), w, bias, mk(y), vm, mask != nullptr, act, alpha, stats, tiles_x, tiles_y);
  DNNCA_LAUNCH_CHECK("conv3x3_small");
  return 1;
}

template <typename T, int CIN, int COUT, bool DGRAD>
static int launch_small_w(cudaStream_t s, const dnnca_tensor_t* x, const float* w, const float* bias,
                          const dnnca_tensor_t* y, const dnnca_tensor_t* mask, int act, float alpha, double* stats) {
  if (x->w > 64) return launch_small<T, CIN, COUT, 32, DGRAD>(s, x, w, bias, y, mask, act, alpha, stats);
  if (x->w > 32) return launch_small<T, CIN, COUT, 16, DGRAD>(s, x, w, bias, y, mask, act, alpha, stats);
  return launch_small<T, CIN, COUT, 8, DGRAD>(s, x, w, bias, y, mask, act, alpha, stats);
}

// (kernel-input channels, kernel-output channels) pairs that occur in configs/unet.yaml
// (C = 3 or 5 modalities), the first layers of mulmo_unet.yaml (1 -> 16) and unet_big.yaml (3 -> 64 is
// left to the generic/tensor path), for fprop and -- with the roles swapped -- dgrad.
#define DNNCA_SMALL_FPROP_SHAPES(X) \
  X(1, 4) X(1, 16) X(3, 3) X(5, 3) X(3, 6) X(6, 3) X(6, 6) X(6, 12) X(12, 6) X(12, 12) X(24, 12) \
  X(3, 4) X(4, 4) X(4, 8) X(8, 8) X(8, 4) X(16, 8)
// dgrad: kernel input = layer Cout, kernel output = layer Cin (first-layer dgrads fall to the generic kernel)
#define DNNCA_SMALL_DGRAD_SHAPES(X) \
  X(3, 3) X(6, 3) X(3, 6) X(6, 6) X(12, 6) X(6, 12) X(12, 12) X(12, 24) X(4, 4) X(8, 4) X(4, 8) X(8, 8) X(8, 16)

#if SMALL_KIND == 0
int SMALL_FN(try_conv_fprop_small)(cudaStream_t s, const dnnca_tensor_t* x, const float* w, const float* bias,
                                   const dnnca_tensor_t* y, int act, float alpha, double* stats) {
#define X(CI, CO) \
  if (x->c == CI && y->c == CO) return launch_small_w<SMALL_T, CI, CO, false>(s, x, w, bias, y, nullptr, act, alpha, stats);
  DNNCA_SMALL_FPROP_SHAPES(X)
#undef X
  return 0;
}
#elif SMALL_KIND == 1
int SMALL_FN(try_conv_dgrad_small)(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                   const dnnca_tensor_t* mask, int act, float alpha) {
#define X(CI, CO) \
  if (dz->c == CI && dx->c == CO) return launch_small_w<SMALL_T, CI, CO, true>(s, dz, w, nullptr, dx, mask, act, alpha, nullptr);
  DNNCA_SMALL_DGRAD_SHAPES(X)
#undef X
  return 0;
}
#endif

template <typename T, int CIN, int COUT>
static int launch_small_wgrad(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dz, float* dw, float* db) {
  constexpr int TYN = 4, TXN = 16, TW = TXN * PX;
  constexpr size_t smem = ((size_t)CIN * (TYN + 2) * (TW + 4) + (size_t)COUT * TYN * TW + 9 * CIN * COUT + COUT) * 4;
  auto kern = conv3x3_small_wgrad_kernel<T, CIN, COUT>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv3x3_small_wgrad: cudaFuncSetAttribute");
    attr_done = true;
  }
  const int tiles_x = (x->w + TW - 1) / TW, tiles_y = (x->h + TYN - 1) / TYN;
  const long long ntiles = (long long)tiles_x * tiles_y * x->n;
  long long grid = (long long)sm_count() * 2;
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, 256, smem, s>>>(mk(x), mk(dz), dw, db, tiles_x, tiles_y, ntiles);
  DNNCA_LAUNCH_CHECK("conv3x3_small_wgrad");
  return 1;
}

#define DNNCA_SMALL_WGRAD_SHAPES(X) \
  X(1, 3) X(1, 4) X(1, 16) X(3, 3) X(5, 3) X(3, 6) X(6, 6) X(6, 12) X(12, 12) X(24, 12) X(12, 6) X(6, 3) X(3, 4) \
  X(4, 4) X(4, 8) X(8, 8) X(16, 8) X(8, 4) X(5, 4)

#if SMALL_KIND == 2
int SMALL_FN(try_conv_wgrad_small)(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dz, float* dw,
                                   float* db) {
#define X(CI, CO) \
  if (x->c == CI && dz->c == CO) return launch_small_wgrad<SMALL_T, CI, CO>(s, x, dz, dw, db);
  DNNCA_SMALL_WGRAD_SHAPES(X)
#undef X
  return 0;
}
#endif

}  // namespace dnnca
