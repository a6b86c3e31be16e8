// BatchNormalization folded into its consumer (north star: "BN-stats/affine ... fused"; SURVEY.md 7.3 option (b)).
//
// The reference's block order is Conv -> act -> BN -> Conv (components.py:46-61, 118-134): the BN output is the next
// conv's input and zero padding is applied AFTER the BN.  With y = s*a + t (per channel) the consumer computes
//     out[p, co] = sum_tap sum_ci w[tap,ci,co] * (s[ci]*a[p+tap, ci] + t[ci]) * inside(p+tap)  + b[co]
//                = sum_tap sum_ci (w[tap,ci,co]*s[ci]) * a[p+tap, ci]                  <- conv of the RAW tensor `a` with scaled weights
//                  + b[co] + sum_{tap : p+tap inside the image} T[tap][co],           T[tap][co] = sum_ci w[tap,ci,co]*t[ci]
// and the second line depends on the pixel only through WHICH taps fall into the padding: 3 row classes x 3 column
// classes = nine bias vectors (exact, not an approximation).  So the BN output never has to be materialised: the
// `bn_apply` pass (read a, write y: 4 B/elem) disappears and one bf16 rounding with it.
// Backward through the fold:
//   dgrad is unchanged (raw weights) and yields d(BN output);
//   wgrad against the BN output x = s*a + t:   dW[tap,ci,co] = s[ci] * sum_p a[p+tap,ci]*dz[p,co]  +  t[ci] * S[tap][co],
//     S[tap][co] = sum of dz over the pixels whose tap stays inside the image = total - excluded row - excluded column
//     + excluded corner: the plain wgrad kernel runs on `a`, then a fix-up pass over dW applies s, t and S.
// MaxPool of a folded tensor: max(s*a+t) = s*max(a)+t for s >= 0 and s*min(a)+t for s < 0 (maxpool_fwd_affine).
#include "common.cuh"

namespace dnnca {

// Tpart[slice][tap][co] = sum_{ci in this block's slice} w[tap,ci,co] * shift[ci]      grid (taps, co tiles, ci slices)
// (partial sums per slice, added in slice order by bias9_kernel: no atomics, so a step is reproducible bit for bit --
// a last-bit difference in a bias flips bf16 roundings downstream and these BN nets amplify that to 1e-3 in the logits)
__global__ void __launch_bounds__(128) fold_shift_kernel(const float* __restrict__ w, int cin, int cout,
                                                        const float* __restrict__ ta, int ca, const float* __restrict__ tb,
                                                        float* __restrict__ T) {
  const int tap = blockIdx.x, co = blockIdx.y * 128 + threadIdx.x;
  const int per = (cin + gridDim.z - 1) / gridDim.z;
  const int c0 = blockIdx.z * per, c1 = min(cin, c0 + per);
  __shared__ float st[64];                       // this slice's shifts (per <= 64)
  for (int i = threadIdx.x; i < c1 - c0; i += blockDim.x) {
    const int ci = c0 + i;
    st[i] = ci < ca ? (ta ? ta[ci] : 0.f) : (tb ? tb[ci - ca] : 0.f);
  }
  __syncthreads();
  if (co >= cout) return;
  const float* wp = w + ((long long)tap * cin + c0) * cout + co;
  float acc = 0.f;
#pragma unroll 8
  for (int i = 0; i < c1 - c0; ++i) acc = fmaf(__ldg(wp + (long long)i * cout), st[i], acc);
  T[((long long)blockIdx.z * 9 + tap) * cout + co] = acc;
}

// bias9[cls][co] = bias[co] + sum over the taps that stay inside the image for border class cls = (row class)*3 + (column
// class), class 0 = first row/column (tap offset -1 is outside), 1 = interior, 2 = last (offset +1 is outside)
__global__ void __launch_bounds__(256) bias9_kernel(const float* __restrict__ T, int slices, const float* __restrict__ bias,
                                                   int cout, float* __restrict__ bias9) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * cout; i += gridDim.x * blockDim.x) {
    const int cls = i / cout, co = i - cls * cout;
    const int ry = cls / 3, rx = cls % 3;
    float b = bias ? bias[co] : 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const bool out_y = (ry == 0 && dy == 0) || (ry == 2 && dy == 2);
        const bool out_x = (rx == 0 && dx == 0) || (rx == 2 && dx == 2);
        if (!out_y && !out_x) {
          float t = 0.f;
          for (int z = 0; z < slices; ++z) t += T[((long long)z * 9 + dy * 3 + dx) * cout + co];
          b += t;
        }
      }
    bias9[i] = b;
  }
}

// border sums of dz [n,h,w,c] (bf16 view): E[0] first row, E[1] last row, E[2] first column, E[3] last column,
// E[4..7] corners (0,0) (0,w-1) (h-1,0) (h-1,w-1), each [c] fp32, accumulated with atomics (caller zeroes).
// grid (n, 4): block (n, k) sums border line k of image n
template <typename T>
__global__ void __launch_bounds__(256) dz_border_sums_kernel(View g, float* __restrict__ E) {
  const int n = blockIdx.x, kind = blockIdx.y;
  const int len = kind < 2 ? g.w : g.h;
  const T* base = reinterpret_cast<const T*>(g.data) + (long long)n * g.h * g.w * g.cstride + g.coff;
  for (int c = threadIdx.x; c < g.c; c += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < len; ++i) {
      const int y = kind == 0 ? 0 : (kind == 1 ? g.h - 1 : i), x = kind == 2 ? 0 : (kind == 3 ? g.w - 1 : i);
      acc += ldf(base + ((long long)y * g.w + x) * g.cstride + c);
    }
    atomicAdd(E + kind * g.c + c, acc);
    if (kind == 0) {       // corners once per image
      atomicAdd(E + 4 * g.c + c, ldf(base + c));
      atomicAdd(E + 5 * g.c + c, ldf(base + (long long)(g.w - 1) * g.cstride + c));
      atomicAdd(E + 6 * g.c + c, ldf(base + (long long)(g.h - 1) * g.w * g.cstride + c));
      atomicAdd(E + 7 * g.c + c, ldf(base + ((long long)(g.h - 1) * g.w + g.w - 1) * g.cstride + c));
    }
  }
}

// bf16, C % 8 == 0: 16-byte loads, (pixel lane, 8-channel group) threads, shared-memory partial sums.  grid (n, 4)
__global__ void __launch_bounds__(256) dz_border_sums_vec8_kernel(const __nv_bfloat16* __restrict__ g, int h, int w, int c,
                                                                 long long cstride, float* __restrict__ E) {
  extern __shared__ float sacc[];                // [5][c]: this line | four corners (kind 0 only)
  const int n = blockIdx.x, kind = blockIdx.y;
  const int len = kind < 2 ? w : h, ng = c / 8;
  for (int i = threadIdx.x; i < 5 * c; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const __nv_bfloat16* base = g + (long long)n * h * w * cstride;
  const int lanes = blockDim.x / ng > 0 ? blockDim.x / ng : 1;
  const int grp = threadIdx.x % ng, lane = threadIdx.x / ng;
  if (lane < lanes) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int i = lane; i < len; i += lanes) {
      const int y = kind == 0 ? 0 : (kind == 1 ? h - 1 : i), x = kind == 2 ? 0 : (kind == 3 ? w - 1 : i);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + ((long long)y * w + x) * cstride + 8 * grp));
      const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[2 * j] += __uint_as_float(wd[j] << 16); acc[2 * j + 1] += __uint_as_float(wd[j] & 0xffff0000u); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(sacc + 8 * grp + j, acc[j]);
  }
  if (kind == 0)
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      sacc[c + ch] = __bfloat162float(base[ch]);
      sacc[2 * c + ch] = __bfloat162float(base[(long long)(w - 1) * cstride + ch]);
      sacc[3 * c + ch] = __bfloat162float(base[(long long)(h - 1) * w * cstride + ch]);
      sacc[4 * c + ch] = __bfloat162float(base[((long long)(h - 1) * w + w - 1) * cstride + ch]);
    }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    atomicAdd(E + kind * c + ch, sacc[ch]);
    if (kind == 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(E + (4 + k) * c + ch, sacc[(1 + k) * c + ch]);
    }
  }
}

// dW[tap,ci,co] = s[ci]*dW[tap,ci,co] + t[ci]*S[tap][co],   S = db - row term - column term + corner term
__global__ void __launch_bounds__(256) wgrad_fixup_kernel(float* __restrict__ dw, const float* __restrict__ db,
                                                         const float* __restrict__ E, int cin, int cout,
                                                         const float* __restrict__ aff_a, int ca,
                                                         const float* __restrict__ aff_b, int cb) {
  const long long total = 9LL * cin * cout;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(e % cout);
    const long long r = e / cout;
    const int ci = (int)(r % cin), tap = (int)(r / cin);
    float s = 1.f, t = 0.f;
    if (ci < ca) { if (aff_a) { s = aff_a[ci]; t = aff_a[ca + ci]; } }
    else if (aff_b) { s = aff_b[ci - ca]; t = aff_b[cb + ci - ca]; }
    float v = s * dw[e];
    if (t != 0.f) {
      // tap (dy,dx) in {0,1,2}^2 reads x[p + (dy-1, dx-1)]: for dy = 0 the pixels of the FIRST row have it outside, for
      // dy = 2 those of the LAST row; columns alike
      const int dy = tap / 3, dx = tap % 3;
      float S = db[co];
      const int rk = dy == 0 ? 0 : (dy == 2 ? 1 : -1), ck = dx == 0 ? 2 : (dx == 2 ? 3 : -1);
      if (rk >= 0) S -= E[rk * cout + co];
      if (ck >= 0) S -= E[ck * cout + co];
      if (rk >= 0 && ck >= 0) S += E[(4 + rk * 2 + (ck - 2)) * cout + co];
      v = fmaf(t, S, v);
    }
    dw[e] = v;
  }
}

// MaxPool 2x2/2 of the folded tensor s*a + t: the window's max of `a` where s >= 0, its min where s < 0 (first such
// element in row-major window order wins, like the plain kernel), then the affine; statistics of the stored output.
template <typename T>
__global__ void __launch_bounds__(256) maxpool_fwd_affine_kernel(View x, const float* __restrict__ aff, View y,
                                                                uint8_t* __restrict__ idx, double* __restrict__ stats, int CL,
                                                                int PL, long long PO) {
  __shared__ double sm[256];
  const int cl = threadIdx.x & (CL - 1), pl = threadIdx.x / CL;
  const int Ho = y.h, Wo = y.w;
  for (int c = cl; c < ((x.c + CL - 1) / CL) * CL; c += CL) {
    double s0 = 0.0, s1 = 0.0;
    if (c < x.c) {
      const float sc = aff[c], sh = aff[x.c + c];
      const float sign = sc < 0.f ? -1.f : 1.f;
      for (long long p = (long long)blockIdx.x * PL + pl; p < PO; p += (long long)gridDim.x * PL) {
        const int ox = (int)(p % Wo);
        const long long t = p / Wo;
        const int oy = (int)(t % Ho);
        const long long n = t / Ho;
        const T* xp = reinterpret_cast<const T*>(x.data) + ((n * x.h + 2 * oy) * x.w + 2 * ox) * (long long)x.cstride + x.coff + c;
        float best = sign * ldf(xp);
        int bi = 0;
        float v = sign * ldf(xp + x.cstride);
        if (v > best) { best = v; bi = 1; }
        v = sign * ldf(xp + (long long)x.w * x.cstride);
        if (v > best) { best = v; bi = 2; }
        v = sign * ldf(xp + (long long)(x.w + 1) * x.cstride);
        if (v > best) { best = v; bi = 3; }
        const float o = rnd<T>(fmaf(sign * best, sc, sh));
        stf(reinterpret_cast<T*>(y.data) + p * y.cstride + y.coff + c, o);
        if (idx) idx[p * x.c + c] = (uint8_t)bi;
        s0 += o;
        s1 += (double)o * o;
      }
    }
    if (stats) {
      // reduce over the PL pixel lanes of this channel lane
      sm[threadIdx.x] = s0;
      __syncthreads();
      if (pl == 0 && c < x.c) { double r = 0.0; for (int l = 0; l < PL; ++l) r += sm[l * CL + cl]; atomicAdd(stats + c, r); }
      __syncthreads();
      sm[threadIdx.x] = s1;
      __syncthreads();
      if (pl == 0 && c < x.c) { double r = 0.0; for (int l = 0; l < PL; ++l) r += sm[l * CL + cl]; atomicAdd(stats + x.c + c, r); }
      __syncthreads();
    }
  }
}

int fprop_umma_affine(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w, const dnnca_tensor_t* y,
                      int act, float alpha, void* ws, size_t ws_bytes, double* stats, const float* scale_a,
                      const float* scale_b, const float* bias9);
int fprop_umma_affine_supported(const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* y);
int try_maxpool_fwd_affine_vec(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, uint8_t*, double*);

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_conv2d_fold_supported(const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* y, int ksize) {
  if (!view_ok(x) || !view_ok(y) || (x2 && !view_ok(x2)) || ksize != 3) return 0;
  if (x->dtype != DNNCA_BF16 || y->dtype != DNNCA_BF16 || !same_nhw(x, y) || (x2 && !same_nhw(x, x2))) return 0;
  return fprop_umma_affine_supported(x, x2, y);
}

static inline int fold_slices(int cin) { return (cin + 31) / 32; }   // 32 input channels per block (<= 64: fold_shift_kernel's smem slice)

extern "C" size_t dnnca_conv2d_fold_scratch_bytes(int cin, int cout) {
  return (cin > 0 && cout > 0) ? (size_t)(9 + 8 + 9 * fold_slices(cin)) * cout * sizeof(float) : 0;
}

// scratch layout (floats): bias9 [9*cout] | E [8*cout] | Tpart [slices][9*cout]
extern "C" int dnnca_conv2d_fprop_affine(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* affine_x,
                                         const float* affine_x2, const float* w, const float* bias, const dnnca_tensor_t* y,
                                         int act, float alpha, double* stats, void* workspace, size_t workspace_bytes,
                                         float* scratch) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && w && workspace && scratch, "conv2d_fprop_affine: bad arguments");
  DNNCA_CHECK_ARG(!x2 || (view_ok(x2) && same_nhw(x, x2)), "conv2d_fprop_affine: bad x2");
  if (!dnnca_conv2d_fold_supported(x, x2, y, 3))
    DNNCA_UNSUPPORTED("conv2d_fprop_affine: shape not served by the folded tensor-core kernel (query dnnca_conv2d_fold_supported)");
  cudaStream_t s = (cudaStream_t)stream;
  const int ca = x->c, cb = x2 ? x2->c : 0, cin = ca + cb, cout = y->c;
  float* bias9 = scratch;
  float* T = scratch + 17 * cout;
  const int zs = fold_slices(cin);
  dim3 grid(9, (cout + 127) / 128, zs);
  fold_shift_kernel<<<grid, 128, 0, s>>>(w, cin, cout, affine_x ? affine_x + ca : nullptr, ca, affine_x2 ? affine_x2 + cb : nullptr, T);
  DNNCA_LAUNCH_CHECK("fold_shift");
  bias9_kernel<<<(9 * cout + 255) / 256, 256, 0, s>>>(T, zs, bias, cout, bias9);
  DNNCA_LAUNCH_CHECK("bias9");
  const int r = fprop_umma_affine(s, x, x2, w, y, act, alpha, workspace, workspace_bytes, stats, affine_x, affine_x2, bias9);
  if (r < 0) return r;
  if (r == 0) DNNCA_UNSUPPORTED("conv2d_fprop_affine: shape not served by the folded tensor-core kernel");
  if (r == 1 && stats) return dnnca_channel_stats(stream, y, stats);
  return DNNCA_OK;
}

extern "C" int dnnca_conv2d_wgrad_affine(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* affine_x,
                                         const float* affine_x2, const dnnca_tensor_t* dz, float* dw, float* db, float* scratch) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(dz) && dw && db && scratch, "conv2d_wgrad_affine: bad arguments (db is required)");
  int r = dnnca_conv2d_wgrad(stream, x, x2, dz, dw, db, 3);           // raw sums against the pre-BN tensor(s)
  if (r != DNNCA_OK) return r;
  cudaStream_t s = (cudaStream_t)stream;
  const int ca = x->c, cb = x2 ? x2->c : 0, cin = ca + cb, cout = dz->c;
  float* E = scratch + 9 * cout;
  cudaError_t e = cudaMemsetAsync(E, 0, sizeof(float) * 8 * cout, s);
  if (e != cudaSuccess) return cuda_fail(e, "conv2d_wgrad_affine: memset");
  note_launch(1);
  dim3 grid(dz->n, 4);
  if (dz->dtype == DNNCA_BF16 && cout % 8 == 0 && cout / 8 <= 256 && dz->coff % 8 == 0 && dz->cstride % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(dz->data) & 15) == 0) {
    dz_border_sums_vec8_kernel<<<grid, 256, sizeof(float) * 5 * cout, s>>>(
        reinterpret_cast<const __nv_bfloat16*>(dz->data) + dz->coff, dz->h, dz->w, cout, dz->cstride, E);
  } else {
    DNNCA_DISPATCH_DTYPE(dz->dtype, dz_border_sums_kernel<T><<<grid, 256, 0, s>>>(mk(dz), E);)
  }
  DNNCA_LAUNCH_CHECK("dz_border_sums");
  wgrad_fixup_kernel<<<grid_for(9LL * cin * cout, 256 * 4, 4), 256, 0, s>>>(dw, db, E, cin, cout, affine_x, ca, affine_x2, cb);
  DNNCA_LAUNCH_CHECK("wgrad_fixup");
  return DNNCA_OK;
}

extern "C" int dnnca_maxpool2x2_fwd_affine(void* stream, const dnnca_tensor_t* x, const float* affine, const dnnca_tensor_t* y,
                                           uint8_t* idx, double* stats) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && affine, "maxpool2x2_fwd_affine: bad arguments");
  DNNCA_CHECK_ARG(y->n == x->n && y->h == x->h / 2 && y->w == x->w / 2 && y->c == x->c && x->dtype == y->dtype,
                  "maxpool2x2_fwd_affine: y must be [n,h/2,w/2,c]");
  {
    const int r = try_maxpool_fwd_affine_vec((cudaStream_t)stream, x, affine, y, idx, stats);
    if (r != 0) return r < 0 ? r : DNNCA_OK;
  }
  ChanLayout L = chan_layout(x->c);
  const long long PO = (long long)y->n * y->h * y->w;
  const int grid = grid_for(PO, L.pl * 8);
  DNNCA_DISPATCH_DTYPE(x->dtype, maxpool_fwd_affine_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(mk(x), affine, mk(y), idx, stats,
                                                                                                      L.cl, L.pl, PO);)
  DNNCA_LAUNCH_CHECK("maxpool2x2_fwd_affine");
  return DNNCA_OK;
}
