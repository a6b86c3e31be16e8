// Thin-plate-spline warp augmentation on the device (SURVEY.md 8f "later" row): random_warp (annotator/data.py:718-763)
// = tfa.image.sparse_image_warp(image, source points, dest points) with its defaults (interpolation_order = 2,
// regularization_weight = 0, num_boundary_points = 0).  The random control points stay a host decision (like the crop
// offsets and flip flags of dnnca_input_tail); the device does the three steps tfa does with dense tensors:
//
//   1. polyharmonic-spline fit (interpolate_spline._solve_interpolation): [[A, B], [B^T, 0]] [w; v] = [f; 0] with
//      A_ij = phi(|c_i - c_j|^2), B = [c, 1], f = dest - source, c = dest, phi(r) = 0.5 r log(max(r, 1e-10));
//   2. dense flow at every pixel (interpolate_spline._apply_interpolation): phi(|q - c_i|^2) w + [q, 1] v;
//   3. bilinear resampling at q - flow (dense_image_warp / interpolate_bilinear: floor clamped to [0, size-2], weights
//      clamped to [0, 1]).
//
// B200 formulation:
//   * `tps_solve_kernel`: one CTA per image, the (P+3) x (P+5) augmented system in FP64 in shared memory (P <= 165; L2-
//     resident scratch beyond), Gaussian elimination with partial pivoting (symmetric indefinite: zero trailing block).
//   * coordinates are divided by the image extent before the fit.  Under the side conditions B^T w = 0 the interpolant
//     is invariant to that scaling (the extra r^2 log s^2 terms collapse into the affine part), and phi stays O(1)
//     instead of O(1e6).
//   * step 2 runs in FP64 as well.  Random control points land within a pixel of each other often enough (0.5 px in
//     the test set) and then carry weights of 1e5 with opposite signs: the flow is a difference of nearly equal phi
//     terms.  Measured: fp32 phi leaves 2e-2 px of error there, fp32 coefficients alone 1e-3 px; the reference's own
//     float32 graph (distances by the |x|^2 - 2xy + |y|^2 expansion in pixel units) sits 0.1-0.3 px from the exact
//     interpolant.  The FP64 log is table driven (`tps_log`: 2 shared-memory words + 8 DFMA).
//   * `tps_warp_kernel` fuses steps 2 and 3: the control points of an image sit in shared memory (y, x, w_y, w_x as
//     doubles), every thread evaluates the flow of its pixel and resamples all channels at once; the dense flow never
//     goes through HBM unless the caller asks for it.  HBM traffic = image in (gathered, within max_diff pixels of
//     the output position) + image out.
#include "common.cuh"

namespace dnnca {

constexpr double TPS_EPSILON = 1e-10;          // interpolate_spline.EPSILON (on squared PIXEL distances)

__device__ __forceinline__ double tps_phi(double r2_norm, double eps_norm) {
  return 0.5 * r2_norm * log(fmax(r2_norm, eps_norm));
}

// ---- FP64 log from a 128-entry table: x = 2^e * m, m in [1, 2); m = m_j (1 + t) with m_j the midpoint of the j-th of 128
// mantissa intervals, |t| <= 2^-8; log x = e ln 2 - log(inv_j) + log1p(t), inv_j = fl(1 / m_j), log1p by a degree-6
// polynomial (t^7 / 7 < 2e-18).  2 shared-memory words + 8 DFMA instead of the ~100 instruction slots of log().
struct TpsLogTab {
  double inv[128];
  double nlg[128];       // -log(inv[j])
};
__device__ __forceinline__ void tps_log_tab_init(TpsLogTab* t) {
  for (int j = threadIdx.x; j < 128; j += blockDim.x) {
    const double inv = 1.0 / (1.0 + ((double)j + 0.5) / 128.0);
    t->inv[j] = inv;
    t->nlg[j] = -log(inv);
  }
}
__device__ __forceinline__ double tps_log(double x, const TpsLogTab* t) {      // x > 0 and normal
  const long long bits = __double_as_longlong(x);
  const int e = (int)(bits >> 52) - 1023;
  const int j = (int)(bits >> 45) & 127;
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double u = fma(m, t->inv[j], -1.0);
  double p = fma(u, -1.0 / 6.0, 0.2);
  p = fma(u, p, -0.25);
  p = fma(u, p, 1.0 / 3.0);
  p = fma(u, p, -0.5);
  p = fma(u, p, 1.0);
  return fma((double)e, 0.69314718055994530942, fma(u, p, t->nlg[j]));
}

// one CTA per image.  The (m x (m + 2)) augmented system, m = P + 3, lives in shared memory when it fits (P <= 165), else
// in the caller's scratch (L2).  coef per image: [m][2] doubles (w rows then v rows, normalised space)
constexpr int TPS_SOLVE_THREADS = 512;

template <bool IN_SMEM>
__global__ void __launch_bounds__(TPS_SOLVE_THREADS) tps_solve_kernel(const float* __restrict__ src, const float* __restrict__ dst,
                                                                      int npts, float inv_extent, double* __restrict__ scratch,
                                                                      double* __restrict__ coef, int* __restrict__ singular) {
  extern __shared__ double sm_mat[];
  const int b = blockIdx.x, m = npts + 3, ld = m + 2;
  double* a = IN_SMEM ? sm_mat : scratch + (size_t)b * m * ld;
  const float* s = src + (size_t)b * npts * 2;
  const float* d = dst + (size_t)b * npts * 2;
  const double inv = (double)inv_extent, eps = TPS_EPSILON * inv * inv;
  // ---- build
  for (int e = threadIdx.x; e < m * ld; e += blockDim.x) {
    const int i = e / ld, j = e % ld;
    double v = 0.0;
    if (i < npts && j < npts) {
      const double dy = ((double)d[2 * i] - (double)d[2 * j]) * inv, dx = ((double)d[2 * i + 1] - (double)d[2 * j + 1]) * inv;
      v = tps_phi(dy * dy + dx * dx, eps);
    } else if (i < npts && j < m) {            // B = [c_y, c_x, 1]
      v = j == npts + 2 ? 1.0 : (double)d[2 * i + (j - npts)] * inv;
    } else if (i >= npts && j < npts) {        // B^T
      v = i == npts + 2 ? 1.0 : (double)d[2 * j + (i - npts)] * inv;
    } else if (i < npts && j >= m) {           // right-hand sides: the control-point flows (pixels)
      v = (double)d[2 * i + (j - m)] - (double)s[2 * i + (j - m)];
    }
    a[e] = v;
  }
  __syncthreads();
  // ---- elimination with partial pivoting
  __shared__ double red_v[TPS_SOLVE_THREADS / 32];
  __shared__ int red_i[TPS_SOLVE_THREADS / 32];
  __shared__ int piv_s, bad;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) bad = 0;
  for (int k = 0; k < m; ++k) {
    double best = -1.0;
    int bi = k;
    for (int i = k + threadIdx.x; i < m; i += blockDim.x) {
      const double v = fabs(a[(size_t)i * ld + k]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
      if (v2 > best || (v2 == best && i2 < bi)) { best = v2; bi = i2; }
    }
    if (lane == 0) { red_v[wid] = best; red_i[wid] = bi; }
    __syncthreads();
    if (wid == 0) {
      best = lane < TPS_SOLVE_THREADS / 32 ? red_v[lane] : -1.0;
      bi = lane < TPS_SOLVE_THREADS / 32 ? red_i[lane] : k;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, best, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
        if (v2 > best || (v2 == best && i2 < bi)) { best = v2; bi = i2; }
      }
      if (lane == 0) {
        piv_s = bi;
        if (!(best > 1e-13)) bad = 1;        // entries are O(1) in normalised coordinates: coincident control points
      }
    }
    __syncthreads();
    const int piv = piv_s;
    if (piv != k) {
      for (int j = k + threadIdx.x; j < ld; j += blockDim.x) {
        const double t = a[(size_t)k * ld + j];
        a[(size_t)k * ld + j] = a[(size_t)piv * ld + j];
        a[(size_t)piv * ld + j] = t;
      }
      __syncthreads();
    }
    const double pk = a[(size_t)k * ld + k];
    const int rows = m - k - 1, cols = ld - k - 1;
    if (pk != 0.0) {
      const double rp = 1.0 / pk;
      for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int i = k + 1 + e / cols, j = k + 1 + e % cols;
        a[(size_t)i * ld + j] = fma(-(a[(size_t)i * ld + k] * rp), a[(size_t)k * ld + j], a[(size_t)i * ld + j]);
      }
    }
    __syncthreads();
  }
  // ---- back substitution (both right-hand sides), column oriented
  for (int k = m - 1; k >= 0; --k) {
    const double pk = a[(size_t)k * ld + k];
    if (threadIdx.x < 2) a[(size_t)k * ld + m + threadIdx.x] = pk != 0.0 ? a[(size_t)k * ld + m + threadIdx.x] / pk : 0.0;
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * k; e += blockDim.x) {
      const int i = e >> 1, r = e & 1;
      a[(size_t)i * ld + m + r] -= a[(size_t)i * ld + k] * a[(size_t)k * ld + m + r];
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) coef[(size_t)b * m * 2 + e] = a[(size_t)(e >> 1) * ld + m + (e & 1)];
  if (threadIdx.x == 0 && bad) atomicExch(singular, 1);
}

// grid = (pixel chunks, n); dynamic smem: npts x 4 doubles
template <int C>
__global__ void __launch_bounds__(256) tps_warp_kernel(const float* __restrict__ image, int h, int w, int c_rt,
                                                       const float* __restrict__ dst, const double* __restrict__ coef, int npts,
                                                       float inv_extent, float* __restrict__ out, float* __restrict__ flow_out) {
  extern __shared__ double sm_pts[];            // [npts][4]: c_y, c_x (pixels), w_y, w_x
  __shared__ TpsLogTab tab;
  tps_log_tab_init(&tab);
  const int b = blockIdx.y, m = npts + 3;
  const int c = C > 0 ? C : c_rt;
  const float* d = dst + (size_t)b * npts * 2;
  const double* cf = coef + (size_t)b * m * 2;
  for (int i = threadIdx.x; i < npts; i += blockDim.x) {
    sm_pts[4 * i] = (double)d[2 * i]; sm_pts[4 * i + 1] = (double)d[2 * i + 1];
    sm_pts[4 * i + 2] = cf[2 * i]; sm_pts[4 * i + 3] = cf[2 * i + 1];
  }
  __syncthreads();
  const double vy0 = cf[2 * npts], vy1 = cf[2 * npts + 1], vx0 = cf[2 * npts + 2], vx1 = cf[2 * npts + 3];
  const double v10 = cf[2 * npts + 4], v11 = cf[2 * npts + 5];
  const double inv = (double)inv_extent, eps = TPS_EPSILON * inv * inv;
  const long long hw = (long long)h * w;
  const float* img = image + (size_t)b * hw * c;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < hw; p += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(p / w), x = (int)(p % w);
    const double py = (double)y, px = (double)x;
    double f0 = 0.0, f1 = 0.0, g0 = 0.0, g1 = 0.0;            // two partial sums per component: shorter dependency chains
    int i = 0;
    for (; i + 1 < npts; i += 2) {
      const double ay = (py - sm_pts[4 * i]) * inv, ax = (px - sm_pts[4 * i + 1]) * inv;
      const double by = (py - sm_pts[4 * i + 4]) * inv, bx = (px - sm_pts[4 * i + 5]) * inv;
      const double ra = fma(ay, ay, ax * ax), rb = fma(by, by, bx * bx);
      const double pa = 0.5 * ra * tps_log(fmax(ra, eps), &tab), pb = 0.5 * rb * tps_log(fmax(rb, eps), &tab);
      f0 = fma(pa, sm_pts[4 * i + 2], f0); f1 = fma(pa, sm_pts[4 * i + 3], f1);
      g0 = fma(pb, sm_pts[4 * i + 6], g0); g1 = fma(pb, sm_pts[4 * i + 7], g1);
    }
    if (i < npts) {
      const double ay = (py - sm_pts[4 * i]) * inv, ax = (px - sm_pts[4 * i + 1]) * inv;
      const double ra = fma(ay, ay, ax * ax);
      const double pa = 0.5 * ra * tps_log(fmax(ra, eps), &tab);
      f0 = fma(pa, sm_pts[4 * i + 2], f0); f1 = fma(pa, sm_pts[4 * i + 3], f1);
    }
    const double qy = py * inv, qx = px * inv;
    const float flow_y = (float)((f0 + g0) + fma(qy, vy0, fma(qx, vx0, v10)));
    const float flow_x = (float)((f1 + g1) + fma(qy, vy1, fma(qx, vx1, v11)));
    if (flow_out) {
      flow_out[((size_t)b * hw + p) * 2] = flow_y;
      flow_out[((size_t)b * hw + p) * 2 + 1] = flow_x;
    }
    // dense_image_warp: sample at grid - flow
    const float sy = (float)y - flow_y, sx = (float)x - flow_x;
    const float fy = fminf(fmaxf(0.f, floorf(sy)), (float)(h - 2)), fx = fminf(fmaxf(0.f, floorf(sx)), (float)(w - 2));
    const float al_y = fminf(fmaxf(0.f, sy - fy), 1.f), al_x = fminf(fmaxf(0.f, sx - fx), 1.f);
    const int iy = (int)fy, ix = (int)fx;
    const float* tl = img + ((size_t)iy * w + ix) * c;
    const float* bl = tl + (size_t)w * c;
    float* o = out + ((size_t)b * hw + p) * c;
#pragma unroll
    for (int ch = 0; ch < c; ++ch) {
      const float vtl = tl[ch], vtr = tl[c + ch], vbl = bl[ch], vbr = bl[c + ch];
      const float top = al_x * (vtr - vtl) + vtl, bot = al_x * (vbr - vbl) + vbl;
      o[ch] = al_y * (bot - top) + top;
    }
  }
}

}  // namespace dnnca

using namespace dnnca;

extern "C" size_t dnnca_tps_workspace_bytes(int n, int npoints) {
  if (n <= 0 || npoints <= 0) return 0;
  const size_t m = (size_t)npoints + 3;
  return (size_t)n * m * (m + 2) * sizeof(double);
}

extern "C" int dnnca_tps_fit(void* stream, const float* source_points, const float* dest_points, int n, int npoints, float extent,
                             void* workspace, size_t workspace_bytes, double* coef, int32_t* singular) {
  DNNCA_CHECK_ARG(source_points && dest_points && workspace && coef && singular, "tps_fit: bad arguments");
  DNNCA_CHECK_ARG(n > 0 && npoints >= 3 && npoints <= 4096 && extent > 0.f, "tps_fit: n > 0, 3 <= npoints <= 4096, extent > 0");
  DNNCA_CHECK_ARG(workspace_bytes >= dnnca_tps_workspace_bytes(n, npoints), "tps_fit: workspace of %zu bytes needed (got %zu)",
                  dnnca_tps_workspace_bytes(n, npoints), workspace_bytes);
  DNNCA_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "tps_fit: workspace must be 8-byte aligned");
  const size_t m = (size_t)npoints + 3, mat_bytes = m * (m + 2) * sizeof(double);
  if (mat_bytes <= 220 * 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(tps_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) return cuda_fail(e, "tps_fit: shared memory attribute");
      attr_set = true;
    }
    tps_solve_kernel<true><<<n, TPS_SOLVE_THREADS, mat_bytes, (cudaStream_t)stream>>>(source_points, dest_points, npoints, 1.0f / extent,
                                                                                  nullptr, coef, singular);
  } else {
    tps_solve_kernel<false><<<n, TPS_SOLVE_THREADS, 0, (cudaStream_t)stream>>>(source_points, dest_points, npoints, 1.0f / extent,
                                                                           reinterpret_cast<double*>(workspace), coef, singular);
  }
  DNNCA_LAUNCH_CHECK("tps_fit");
  return DNNCA_OK;
}

extern "C" int dnnca_tps_warp(void* stream, const float* image, int n, int h, int w, int c, const float* dest_points,
                              const double* coef, int npoints, float extent, float* out, float* flow_out) {
  DNNCA_CHECK_ARG(image && dest_points && coef && out && image != out, "tps_warp: bad arguments");
  DNNCA_CHECK_ARG(n > 0 && n <= 65535 && h >= 2 && w >= 2 && c >= 1 && c <= 64, "tps_warp: 1 <= n <= 65535, h, w >= 2, 1 <= c <= 64");
  DNNCA_CHECK_ARG(npoints >= 3 && npoints <= 1024 && extent > 0.f, "tps_warp: 3 <= npoints <= 1024");
  const long long hw = (long long)h * w;
  int chunks = (int)((hw + 255) / 256);
  const int cap = (sm_count() * 8 + n - 1) / n;
  if (chunks > cap) chunks = cap < 1 ? 1 : cap;
  const dim3 grid(chunks, n);
  const size_t smem = (size_t)npoints * 4 * sizeof(double);
  const float inv = 1.0f / extent;
  cudaStream_t st = (cudaStream_t)stream;
#define TPS_LAUNCH(CC) tps_warp_kernel<CC><<<grid, 256, smem, st>>>(image, h, w, c, dest_points, coef, npoints, inv, out, flow_out)
  switch (c) {
    case 1: TPS_LAUNCH(1); break;
    case 2: TPS_LAUNCH(2); break;
    case 3: TPS_LAUNCH(3); break;
    case 4: TPS_LAUNCH(4); break;
    case 6: TPS_LAUNCH(6); break;
    default: TPS_LAUNCH(0); break;
  }
#undef TPS_LAUNCH
  DNNCA_LAUNCH_CHECK("tps_warp");
  return DNNCA_OK;
}
