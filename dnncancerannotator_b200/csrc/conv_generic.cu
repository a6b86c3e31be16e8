// Shape-generic CUDA-core convolution kernels (any channel count, fp32 or bf16
// activations, fp32 accumulate).  They are the fp32-mode path used for bit-exact
// mask / pool-index checks, and the fallback for shapes the specialised kernels
// (conv_small.cu: tiny channel counts; conv_umma.cu: tcgen05 implicit GEMM) do
// not cover.  Reference call sites: layers.Conv2D components.py:47-50,123-126;
// layers.Convolution2DTranspose components.py:118-120.
#include "common.cuh"

namespace dnnca {

template <typename T>
__device__ __forceinline__ const T* pxp(const View& v, long long p) {
  return reinterpret_cast<const T*>(v.data) + p * v.cstride + v.coff;
}
template <typename T>
__device__ __forceinline__ T* pxw(const View& v, long long p) {
  return reinterpret_cast<T*>(v.data) + p * v.cstride + v.coff;
}

// ------------------------------------------------------------------ fprop ---
// one thread = one output pixel x 4 output channels
template <typename T>
__global__ void __launch_bounds__(256) conv_fprop_generic_kernel(View x, View x2, int c2, const float* __restrict__ w,
                                                                const float* __restrict__ bias, View y, int k,
                                                                int act, float alpha, long long total, int ncb) {
  const int H = x.h, W = x.w, C1 = x.c, Cin = x.c + c2, Cout = y.c, ph = k / 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(t % ncb);
    const long long p = t / ncb;
    const int px = (int)(p % W);
    const long long r = p / W;
    const int py = (int)(r % H);
    const long long n = r / H;
    const int co0 = cb * 4;
    const int nco = min(4, Cout - co0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < k; ++a) {
      const int yy = py + a - ph;
      if (yy < 0 || yy >= H) continue;
      for (int c = 0; c < k; ++c) {
        const int xx = px + c - ph;
        if (xx < 0 || xx >= W) continue;
        const T* xp = pxp<T>(x, (n * H + yy) * W + xx);
        const T* xp2 = c2 ? pxp<T>(x2, (n * H + yy) * W + xx) : nullptr;
        const float* wp = w + (size_t)((a * k + c) * Cin) * Cout + co0;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = ci < C1 ? ldf(xp + ci) : ldf(xp2 + (ci - C1));
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nco) acc[j] = fmaf(xv, wp[(size_t)ci * Cout + j], acc[j]);
        }
      }
    }
    T* yp = pxw<T>(y, p) + co0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nco) stf(yp + j, apply_act(acc[j] + (bias ? bias[co0 + j] : 0.f), act, alpha));
  }
}

// ------------------------------------------------------------------ dgrad ---
// dx[n,p,q,ci] = sum_{a,c,co} dz[n,p-a+ph,q-c+ph,co] * w[a,c,ci,co]
template <typename T>
__global__ void __launch_bounds__(256) conv_dgrad_generic_kernel(View dz, const float* __restrict__ w, View dx,
                                                                View dx2, int c2, int k, View mask, int has_mask,
                                                                int act, float alpha, long long total, int ncb) {
  const int H = dx.h, W = dx.w, C1 = dx.c, Cin = dx.c + c2, Cout = dz.c, ph = k / 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(t % ncb);
    const long long p = t / ncb;
    const int px = (int)(p % W);
    const long long r = p / W;
    const int py = (int)(r % H);
    const long long n = r / H;
    const int ci0 = cb * 4;
    const int nci = min(4, Cin - ci0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < k; ++a) {
      const int yy = py - a + ph;
      if (yy < 0 || yy >= H) continue;
      for (int c = 0; c < k; ++c) {
        const int xx = px - c + ph;
        if (xx < 0 || xx >= W) continue;
        const T* gp = pxp<T>(dz, (n * H + yy) * W + xx);
        const float* wp = w + ((size_t)(a * k + c) * Cin + ci0) * Cout;
        for (int co = 0; co < Cout; ++co) {
          const float g = ldf(gp + co);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < nci) acc[j] = fmaf(g, wp[(size_t)j * Cout + co], acc[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= nci) continue;
      const int ci = ci0 + j;
      if (ci < C1) {
        const float m = has_mask ? act_grad(ldf(pxp<T>(mask, p) + ci), act, alpha) : 1.f;
        stf(pxw<T>(dx, p) + ci, acc[j] * m);
      } else {
        stf(pxw<T>(dx2, p) + (ci - C1), acc[j]);
      }
    }
  }
}

// ------------------------------------------------------------------ wgrad ---
// Tiled reduction over pixels.  MODE 0 (Conv2D):  out[tap][ci][co] += sum_p x[p shifted by tap][ci] * dz[p][co]
//                               MODE 1 (ConvT2x2): out[tap][co][ci] += sum_p dy[2i+a,2j+b][co] * x[i,j][ci]
// grid = (M tiles, N tiles, taps * ksplit); block 16x16, 2x2 outputs per thread, 32-pixel smem stages.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) wgrad_generic_kernel(View xa, View xa2, int c2, View gb,
                                                           float* __restrict__ out, int k, int ksplit, long long P) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tap = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const int a = tap / k, c = tap % k;
  // rows (M) come from the first view for MODE 0 (x), from the second (dy) for MODE 1
  const View& vm = MODE == 0 ? xa : gb;
  const View& vn = MODE == 0 ? gb : xa;
  const int M = vm.c + (MODE == 0 ? c2 : 0), N = vn.c;
  const int m0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int H = xa.h, W = xa.w;  // pixel grid the reduction runs over (= x's grid in both modes)
  const long long chunk = (P + ksplit - 1) / ksplit;
  const long long pbeg = ks * chunk, pend = min(P, pbeg + chunk);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (long long p0 = pbeg; p0 < pend; p0 += 32) {
    // stage: 32 pixels x 32 channels for both operands; thread loads 4 elements of each
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int lin = threadIdx.x + e * 256;
      const int kk = lin >> 5, ch = lin & 31;
      const long long p = p0 + kk;
      float va = 0.f, vb = 0.f;
      if (p < pend) {
        const int px = (int)(p % W);
        const long long r = p / W;
        const int py = (int)(r % H);
        const long long n = r / H;
        if (MODE == 0) {
          const int yy = py + a - k / 2, xx = px + c - k / 2;
          if (m0 + ch < M && yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const int m = m0 + ch;
            va = m < xa.c ? ldf(pxp<T>(xa, (n * H + yy) * W + xx) + m) : ldf(pxp<T>(xa2, (n * H + yy) * W + xx) + (m - xa.c));
          }
          if (n0 + ch < N) vb = ldf(pxp<T>(gb, p) + n0 + ch);
        } else {
          const long long q = (n * gb.h + 2 * py + a) * gb.w + 2 * px + c;
          if (m0 + ch < M) va = ldf(pxp<T>(gb, q) + m0 + ch);
          if (n0 + ch < N) vb = ldf(pxp<T>(xa, p) + n0 + ch);
        }
      }
      As[kk][ch] = va;
      Bs[kk][ch] = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float a0 = As[kk][ty], a1 = As[kk][ty + 16];
      const float b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < M && n < N) atomicAdd(out + ((size_t)tap * M + m) * N + n, acc[i][j]);
    }
}

// per-channel sum of a view accumulated into fp32 (bias gradients)
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(View g, float* __restrict__ out, int CL, int PL,
                                                         long long P) {
  __shared__ float sm[256];
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  for (int c = cl; c < ((g.c + CL - 1) / CL) * CL; c += CL) {
    float s = 0.f;
    if (c < g.c)
      for (long long p = (long long)blockIdx.x * PL + pl; p < P; p += (long long)gridDim.x * PL)
        s += ldf(pxp<T>(g, p) + c);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int st = PL >> 1; st > 0; st >>= 1) {
      if ((int)threadIdx.x < st * CL) sm[threadIdx.x] += sm[threadIdx.x + st * CL];
      __syncthreads();
    }
    if (pl == 0 && c < g.c) atomicAdd(out + c, sm[cl]);
    __syncthreads();
  }
}

// ------------------------------------------------------- ConvTranspose 2x2 ---
// y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci]*k[a,b,co,ci] + bias[co]; thread = output pixel x 4 co
template <typename T>
__global__ void __launch_bounds__(256) tconv_fprop_generic_kernel(View x, const float* __restrict__ kw,
                                                                 const float* __restrict__ bias, View y,
                                                                 long long total, int ncb) {
  const int Cin = x.c, Cout = y.c, HO = y.h, WO = y.w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(t % ncb);
    const long long p = t / ncb;
    const int ox = (int)(p % WO);
    const long long r = p / WO;
    const int oy = (int)(r % HO);
    const long long n = r / HO;
    const int tap = (oy & 1) * 2 + (ox & 1);
    const T* xp = pxp<T>(x, (n * x.h + (oy >> 1)) * x.w + (ox >> 1));
    const int co0 = cb * 4, nco = min(4, Cout - co0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wp = kw + ((size_t)tap * Cout + co0) * Cin;
    for (int ci = 0; ci < Cin; ++ci) {
      const float xv = ldf(xp + ci);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < nco) acc[j] = fmaf(xv, wp[(size_t)j * Cin + ci], acc[j]);
    }
    T* yp = pxw<T>(y, p) + co0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nco) stf(yp + j, acc[j] + (bias ? bias[co0 + j] : 0.f));
  }
}

// dx[n,i,j,ci] = sum_{a,b,co} dy[n,2i+a,2j+b,co]*k[a,b,co,ci]; thread = input pixel x 4 ci
template <typename T>
__global__ void __launch_bounds__(256) tconv_dgrad_generic_kernel(View dy, const float* __restrict__ kw, View dx,
                                                                 View mask, int has_mask, int act, float alpha,
                                                                 long long total, int ncb) {
  const int Cin = dx.c, Cout = dy.c, H = dx.h, W = dx.w;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(t % ncb);
    const long long p = t / ncb;
    const int px = (int)(p % W);
    const long long r = p / W;
    const int py = (int)(r % H);
    const long long n = r / H;
    const int ci0 = cb * 4, nci = min(4, Cin - ci0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const T* gp = pxp<T>(dy, (n * dy.h + 2 * py + (tap >> 1)) * dy.w + 2 * px + (tap & 1));
      const float* wp = kw + (size_t)tap * Cout * Cin + ci0;
      for (int co = 0; co < Cout; ++co) {
        const float g = ldf(gp + co);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nci) acc[j] = fmaf(g, wp[(size_t)co * Cin + j], acc[j]);
      }
    }
    T* dp = pxw<T>(dx, p) + ci0;
    const T* mp = has_mask ? pxp<T>(mask, p) + ci0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j < nci) stf(dp + j, has_mask ? acc[j] * act_grad(ldf(mp + j), act, alpha) : acc[j]);
  }
}

// ----------------------------------------------------------- host launchers --
int launch_conv_fprop_generic(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                              const float* bias, const dnnca_tensor_t* y, int k, int act, float alpha) {
  const int ncb = (y->c + 3) / 4;
  const long long total = (long long)x->n * x->h * x->w * ncb;
  const int grid = grid_for(total, 256, 16);
  View v2 = x2 ? mk(x2) : mk(x);
  const int c2 = x2 ? x2->c : 0;
  DNNCA_DISPATCH_DTYPE(x->dtype, conv_fprop_generic_kernel<T><<<grid, 256, 0, s>>>(mk(x), v2, c2, w, bias, mk(y), k, act, alpha, total, ncb);)
  DNNCA_LAUNCH_CHECK("conv_fprop_generic");
  note_family(0);
  return DNNCA_OK;
}

int launch_conv_dgrad_generic(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                              const dnnca_tensor_t* dx2, int k, const dnnca_tensor_t* mask, int act, float alpha) {
  const int c2 = dx2 ? dx2->c : 0;
  const int ncb = (dx->c + c2 + 3) / 4;
  const long long total = (long long)dx->n * dx->h * dx->w * ncb;
  const int grid = grid_for(total, 256, 16);
  View vm = mask ? mk(mask) : mk(dx);
  View v2 = dx2 ? mk(dx2) : mk(dx);
  DNNCA_DISPATCH_DTYPE(dx->dtype, conv_dgrad_generic_kernel<T><<<grid, 256, 0, s>>>(mk(dz), w, mk(dx), v2, c2, k, vm, mask != nullptr, act, alpha, total, ncb);)
  DNNCA_LAUNCH_CHECK("conv_dgrad_generic");
  note_family(0);
  return DNNCA_OK;
}

static int pick_ksplit(int tiles, long long P) {
  long long want = (2LL * sm_count() + tiles - 1) / tiles;
  long long maxs = P / 128;
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 4096) want = 4096;
  return (int)want;
}

int launch_channel_sum(cudaStream_t s, const dnnca_tensor_t* g, float* out) {
  ChanLayout L = chan_layout(g->c);
  long long P = (long long)g->n * g->h * g->w;
  int grid = grid_for(P, L.pl * 8, 4);
  DNNCA_DISPATCH_DTYPE(g->dtype, channel_sum_kernel<T><<<grid, 256, 0, s>>>(mk(g), out, L.cl, L.pl, P);)
  DNNCA_LAUNCH_CHECK("channel_sum");
  return DNNCA_OK;
}

int launch_conv_wgrad_generic(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2,
                              const dnnca_tensor_t* dz, float* dw, float* db, int k) {
  const long long P = (long long)x->n * x->h * x->w;
  const int c2 = x2 ? x2->c : 0;
  dim3 grid((x->c + c2 + 31) / 32, (dz->c + 31) / 32, 1);
  const int ksplit = pick_ksplit(grid.x * grid.y * k * k, P);
  grid.z = k * k * ksplit;
  View v2 = x2 ? mk(x2) : mk(x);
  DNNCA_DISPATCH_DTYPE(x->dtype, (wgrad_generic_kernel<T, 0><<<grid, 256, 0, s>>>(mk(x), v2, c2, mk(dz), dw, k, ksplit, P));)
  DNNCA_LAUNCH_CHECK("conv_wgrad_generic");
  note_family(0);
  if (db) return launch_channel_sum(s, dz, db);
  return DNNCA_OK;
}

int launch_tconv_fprop_generic(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const float* bias,
                               const dnnca_tensor_t* y) {
  const int ncb = (y->c + 3) / 4;
  const long long total = (long long)y->n * y->h * y->w * ncb;
  const int grid = grid_for(total, 256, 16);
  DNNCA_DISPATCH_DTYPE(x->dtype, tconv_fprop_generic_kernel<T><<<grid, 256, 0, s>>>(mk(x), kw, bias, mk(y), total, ncb);)
  DNNCA_LAUNCH_CHECK("tconv_fprop_generic");
  note_family(0);
  return DNNCA_OK;
}

int launch_tconv_dgrad_generic(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx,
                               const dnnca_tensor_t* mask, int act, float alpha) {
  const int ncb = (dx->c + 3) / 4;
  const long long total = (long long)dx->n * dx->h * dx->w * ncb;
  const int grid = grid_for(total, 256, 16);
  View vm = mask ? mk(mask) : mk(dx);
  DNNCA_DISPATCH_DTYPE(dx->dtype, tconv_dgrad_generic_kernel<T><<<grid, 256, 0, s>>>(mk(dy), kw, mk(dx), vm, mask != nullptr, act, alpha, total, ncb);)
  DNNCA_LAUNCH_CHECK("tconv_dgrad_generic");
  note_family(0);
  return DNNCA_OK;
}

int launch_tconv_wgrad_generic(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk,
                               float* db) {
  const long long P = (long long)x->n * x->h * x->w;
  dim3 grid((dy->c + 31) / 32, (x->c + 31) / 32, 1);
  const int ksplit = pick_ksplit(grid.x * grid.y * 4, P);
  grid.z = 4 * ksplit;
  DNNCA_DISPATCH_DTYPE(x->dtype, (wgrad_generic_kernel<T, 1><<<grid, 256, 0, s>>>(mk(x), mk(x), 0, mk(dy), dk, 2, ksplit, P));)
  DNNCA_LAUNCH_CHECK("tconv_wgrad_generic");
  note_family(0);
  if (db) return launch_channel_sum(s, dy, db);
  return DNNCA_OK;
}

}  // namespace dnnca
