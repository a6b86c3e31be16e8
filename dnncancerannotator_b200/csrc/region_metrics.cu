// Region-based detection counters on the device (SURVEY.md 8f "later" row): the update_state of the reference's
// RegionBased{Recall,Precision,TruePositives,FalsePositives,FalseNegatives,FBetaScore,ConfusionMatrix}
// (annotator/utils/metrics.py:80-520) and of the Visualizer's region PR curve (callbacks.py:226-230).
//
// The reference, per slice and per threshold t (metrics.py:108-192, 254-288):
//   label regions  = connected components (4-neighbourhood, tfa.image.connected_components) of  label > 0.5
//   pred regions   = connected components of  morph_open(pred >= t, 5x5)          (image.py:12-29: erosion2d, dilation2d)
//   IoU[i,j]       = |L_i & P_j| / |L_i | P_j|  from one-hot masks broadcast to [labels, preds, H, W, thresholds]
//   tp = #labels with some IoU > IoU_threshold, fn = the other labels, fp = #preds with no IoU > IoU_threshold
// (optionally after a bilinear resize of both images, metrics.py:194-204).
//
// B200 formulation -- nothing is one-hot, every pixel is touched a constant number of times:
//   * thresholding commutes with min / max filters, so the morphological opening is done ONCE on the probabilities
//     (grey opening, `grey_open_kernel`: one shared-memory tile pass) and every threshold plane is  open(p) >= t;
//   * all n*(T+1) binary planes (label plane + T prediction planes per slice) are labelled together by a lock-free
//     union-find in global memory / L2 (init with warp-ballot row runs, merge with atomicMin on the roots, flatten);
//     a component is named by its root = the smallest row-major pixel index in it (the order tfa numbers them in);
//   * regions of one image are disjoint, so |L_i & P_j| is a pair histogram: every pixel lying in both a label region
//     and a prediction region adds one to the key (root_i, root_j) of a small open-addressing table per (slice,
//     threshold) (warp-aggregated); |L_i | P_j| = area_i + area_j - |L_i & P_j|;
//   * `region_eval_kernel` walks the tables, marks detected label roots / hitting prediction roots,
//     `region_count_kernel` counts roots.
// Integer work, bit-exact against the oracle; the IoU comparison is evaluated in fp32 exactly like the reference
// (float32 intersection / float32 union > float32 threshold).
#include "common.cuh"

namespace dnnca {

// ---- bilinear resize, TF2 semantics (half-pixel centres, no antialias) --------------------------------------------
__global__ void __launch_bounds__(256) resize_bilinear_kernel(const float* __restrict__ src, int n, int h, int w,
                                                              float* __restrict__ dst, int oh, int ow) {
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
  const long long total = (long long)n * oh * ow;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % ow);
    const long long r = p / ow;
    const int oy = (int)(r % oh), b = (int)(r / oh);
    const float fy = __fsub_rn(__fmul_rn((float)oy + 0.5f, sy), 0.5f), fx = __fsub_rn(__fmul_rn((float)ox + 0.5f, sx), 0.5f);
    const float fyf = floorf(fy), fxf = floorf(fx);
    const int y0 = max((int)fyf, 0), y1 = min((int)ceilf(fy), h - 1);
    const int x0 = max((int)fxf, 0), x1 = min((int)ceilf(fx), w - 1);
    const float ly = __fsub_rn(fy, fyf), lx = __fsub_rn(fx, fxf);
    const float* s = src + (long long)b * h * w;
    const float tl = s[(long long)y0 * w + x0], tr = s[(long long)y0 * w + x1];
    const float bl = s[(long long)y1 * w + x0], br = s[(long long)y1 * w + x1];
    // no fused multiply-add: the reference rounds the product and the sum separately
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), lx));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), lx));
    dst[p] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ly));
  }
}

// ---- grey opening: max_{k x k}( min_{k x k}( p ) ), windows clipped to the image --------------------------------
// One CTA = one 32x32 output tile.  Shared tile of (32 + 2(k-1))^2 probabilities, +inf outside the image (neutral for
// the erosion); rows-min, columns-min give the eroded tile of (32 + k-1)^2, set to -inf outside the image (neutral
// for the dilation); rows-max, columns-max give the output.
constexpr int OPEN_TILE = 32;
constexpr int OPEN_MAX_K = 15;

__global__ void __launch_bounds__(256) grey_open_kernel(const float* __restrict__ src, int n, int h, int w, int k,
                                                        float* __restrict__ dst) {
  extern __shared__ float sm_open[];
  const int halo = k - 1, before = (k - 1) / 2;
  const int S0 = OPEN_TILE + 2 * halo;      // input tile edge
  const int S1 = OPEN_TILE + halo;          // eroded tile edge
  float* a = sm_open;                       // [S0][S0]
  float* bbuf = sm_open + S0 * S0;          // [S0][S1] then reused
  const int tiles_x = (w + OPEN_TILE - 1) / OPEN_TILE, tiles_y = (h + OPEN_TILE - 1) / OPEN_TILE;
  const float INF = __int_as_float(0x7f800000);
  for (long long tile = blockIdx.x; tile < (long long)n * tiles_x * tiles_y; tile += gridDim.x) {
    const int tx = (int)(tile % tiles_x);
    const long long r = tile / tiles_x;
    const int ty = (int)(r % tiles_y), b = (int)(r / tiles_y);
    const int oy0 = ty * OPEN_TILE, ox0 = tx * OPEN_TILE;
    // eroded tile covers rows oy0 - before .. ; input tile covers rows oy0 - 2*before ..
    const int ey0 = oy0 - before, ex0 = ox0 - before;
    const int iy0 = ey0 - before, ix0 = ex0 - before;
    const float* s = src + (long long)b * h * w;
    for (int i = threadIdx.x; i < S0 * S0; i += blockDim.x) {
      const int yy = iy0 + i / S0, xx = ix0 + i % S0;
      a[i] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? s[(long long)yy * w + xx] : INF;
    }
    __syncthreads();
    // horizontal min: bbuf[y][x] = min_{d<k} a[y][x+d], x < S1
    for (int i = threadIdx.x; i < S0 * S1; i += blockDim.x) {
      const int yy = i / S1, xx = i % S1;
      float m = INF;
      for (int d = 0; d < k; ++d) m = fminf(m, a[yy * S0 + xx + d]);
      bbuf[i] = m;
    }
    __syncthreads();
    // vertical min -> eroded tile in a[y][x], y,x < S1; outside the image -> -inf
    for (int i = threadIdx.x; i < S1 * S1; i += blockDim.x) {
      const int yy = i / S1, xx = i % S1;
      float m = INF;
      for (int d = 0; d < k; ++d) m = fminf(m, bbuf[(yy + d) * S1 + xx]);
      const int gy = ey0 + yy, gx = ex0 + xx;
      a[i] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? m : -INF;
    }
    __syncthreads();
    // horizontal max: bbuf[y][x] = max_{d<k} a[y][x+d], y < S1, x < 32
    for (int i = threadIdx.x; i < S1 * OPEN_TILE; i += blockDim.x) {
      const int yy = i / OPEN_TILE, xx = i % OPEN_TILE;
      float m = -INF;
      for (int d = 0; d < k; ++d) m = fmaxf(m, a[yy * S1 + xx + d]);
      bbuf[i] = m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OPEN_TILE * OPEN_TILE; i += blockDim.x) {
      const int yy = i / OPEN_TILE, xx = i % OPEN_TILE;
      const int gy = oy0 + yy, gx = ox0 + xx;
      if (gy < h && gx < w) {
        float m = -INF;
        for (int d = 0; d < k; ++d) m = fmaxf(m, bbuf[(yy + d) * OPEN_TILE + xx]);
        dst[((long long)b * h + gy) * w + gx] = m;
      }
    }
    __syncthreads();
  }
}

// ---- union-find over pixels -------------------------------------------------------------------------------------
__device__ __forceinline__ int uf_find(const int* L, int i) {
  int p;
  while ((p = __ldcg(L + i)) != i) i = p;
  return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
  for (;;) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(L + a, b);        // a was a root iff old == a
    if (old == a) return;
    a = old;                                    // somebody re-parented a meanwhile: join its new parent with b
  }
}

struct RegionGeom {
  int n, h, w, nthr;
  long long hw;
};

// foreground test of plane (b, k): k == 0 the label (label > 0.5, metrics.py:126), k >= 1 opened prediction >= t_{k-1}
// (metrics.py:135; the comparison is `>=`, unlike the pixel metrics' `>`)
__device__ __forceinline__ bool plane_fg(const float* __restrict__ labels, const float* __restrict__ popen,
                                         const float* __restrict__ thr, const RegionGeom& g, int b, int k, long long i) {
  return k == 0 ? labels[(long long)b * g.hw + i] > 0.5f : popen[(long long)b * g.hw + i] >= thr[k - 1];
}

// L[plane][i] = first pixel of the horizontal run i belongs to within its 32-pixel warp segment, -1 for background.
// FROM_MASK: planes are the images of a uint8 mask (standalone connected components), else the label / prediction planes.
template <bool FROM_MASK>
__global__ void __launch_bounds__(256) ccl_init_kernel(const float* __restrict__ labels, const float* __restrict__ popen,
                                                       const float* __restrict__ thr, const uint8_t* __restrict__ mask,
                                                       RegionGeom g, int* __restrict__ L) {
  const long long planes = (long long)g.n * (g.nthr + 1);
  const long long total = planes * g.hw;
  const long long padded = (total + 31) & ~31LL;
  const int lane = threadIdx.x & 31;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < padded; gi += (long long)gridDim.x * blockDim.x) {
    bool fg = false;
    long long i = 0;
    if (gi < total) {
      const long long q = gi / g.hw;
      i = gi % g.hw;
      fg = FROM_MASK ? mask[gi] != 0 : plane_fg(labels, popen, thr, g, (int)(q / (g.nthr + 1)), (int)(q % (g.nthr + 1)), i);
    }
    const unsigned bits = __ballot_sync(0xffffffffu, fg);
    if (gi < total) {
      int v = -1;
      if (fg) {
        const int x = (int)(i % g.w);
        const int lo = lane - min(lane, x);                       // first lane of this image row inside the warp
        const unsigned below = (lane ? (0xffffffffu >> (32 - lane)) : 0u) & ~(lo ? (0xffffffffu >> (32 - lo)) : 0u);
        const unsigned holes = ~bits & below;                     // background lanes in [lo, lane)
        const int start = holes ? (32 - __clz(holes)) : lo;
        v = (int)i - (lane - start);
      }
      L[gi] = v;
    }
  }
}

__global__ void __launch_bounds__(256) ccl_merge_kernel(RegionGeom g, int* __restrict__ L) {
  const long long total = (long long)g.n * (g.nthr + 1) * g.hw;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
    const long long q = gi / g.hw;
    const int i = (int)(gi % g.hw);
    int* P = L + q * g.hw;
    if (__ldcg(P + i) < 0) continue;
    const int x = i % g.w, y = i / g.w;
    const bool left = x > 0 && __ldcg(P + i - 1) >= 0;
    // runs were joined inside a warp segment by the init kernel (global index = lane mod 32): the only horizontal link
    // still missing is the one across a segment boundary
    if (left && (gi & 31) == 0) uf_union(P, i, i - 1);
    if (y > 0 && __ldcg(P + i - g.w) >= 0) {
      // the pixel above is foreground; if left and upper-left are too, the left neighbour already made this link
      const bool upleft = x > 0 && __ldcg(P + i - g.w - 1) >= 0;
      if (!(left && upleft)) uf_union(P, i, i - g.w);
    }
  }
}

// 64-bit mix (splitmix finaliser) for the pair table
__device__ __forceinline__ unsigned hash_pair(unsigned long long k) {
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ULL;
  k ^= k >> 27; k *= 0x94d049bb133111ebULL;
  k ^= k >> 31;
  return (unsigned)k;
}

// flatten every plane (L[i] = root), accumulate areas at the roots, and insert (label root, prediction root) pairs
__global__ void __launch_bounds__(256) ccl_flatten_pairs_kernel(RegionGeom g, int* __restrict__ L, int* __restrict__ area,
                                                                unsigned long long* __restrict__ keys,
                                                                unsigned* __restrict__ vals, int slots,
                                                                int* __restrict__ overflow) {
  const long long total = (long long)g.n * (g.nthr + 1) * g.hw;
  const long long padded = (total + 31) & ~31LL;
  const int lane = threadIdx.x & 31;
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < padded; gi += (long long)gridDim.x * blockDim.x) {
    int root = -1, lroot = -1;
    long long q = 0;
    int k = 0, b = 0;
    if (gi < total) {
      q = gi / g.hw;
      const int i = (int)(gi % g.hw);
      b = (int)(q / (g.nthr + 1)); k = (int)(q % (g.nthr + 1));
      int* P = L + q * g.hw;
      if (__ldcg(P + i) >= 0) {
        root = uf_find(P, i);
        P[i] = root;
        if (k > 0) {
          const int* PL = L + (long long)b * (g.nthr + 1) * g.hw;   // label plane of this slice
          if (__ldcg(PL + i) >= 0) lroot = uf_find(PL, i);
        }
      }
    }
    // area: one atomic per (warp, root)
    const unsigned long long akey = root >= 0 ? (((unsigned long long)q << 32) | (unsigned)root) + 1ULL : 0ULL;
    const unsigned apeers = __match_any_sync(0xffffffffu, akey);
    if (akey && (int)(__ffs(apeers) - 1) == lane) atomicAdd(area + q * g.hw + root, __popc(apeers));
    // pair histogram: one insertion per (warp, pair)
    const bool both = root >= 0 && lroot >= 0;
    const unsigned long long pkey = both ? (((unsigned long long)(unsigned)lroot << 32) | (unsigned)root) + 1ULL : 0ULL;
    // lanes of one warp may sit in different planes (H*W not a multiple of 32): equal pair AND equal plane
    const unsigned ppeers = __match_any_sync(0xffffffffu, pkey) & __match_any_sync(0xffffffffu, (unsigned long long)q);
    if (both && (int)(__ffs(ppeers) - 1) == lane) {
      const long long tbl = ((long long)b * g.nthr + (k - 1)) * slots;
      unsigned hsh = hash_pair(pkey) & (unsigned)(slots - 1);
      int probe = 0;
      for (; probe < slots; ++probe) {
        const unsigned long long old = atomicCAS(keys + tbl + hsh, 0ULL, pkey);
        if (old == 0ULL || old == pkey) {
          atomicAdd(vals + tbl + hsh, (unsigned)__popc(ppeers));
          break;
        }
        hsh = (hsh + 1) & (unsigned)(slots - 1);
      }
      if (probe == slots) atomicExch(overflow, 1);
    }
  }
}

// one thread per table slot: IoU of the pair against the threshold, marks on both roots
__global__ void __launch_bounds__(256) region_eval_kernel(RegionGeom g, const int* __restrict__ area,
                                                          const unsigned long long* __restrict__ keys,
                                                          const unsigned* __restrict__ vals, int slots, float iou_threshold,
                                                          uint8_t* __restrict__ det_label, uint8_t* __restrict__ det_pred) {
  const long long total = (long long)g.n * g.nthr * slots;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[s];
    if (!key) continue;
    const long long bt = s / slots;
    const int b = (int)(bt / g.nthr), t = (int)(bt % g.nthr);
    const int lroot = (int)((key - 1) >> 32), proot = (int)((key - 1) & 0xffffffffu);
    const int inter = (int)vals[s];
    const int al = area[((long long)b * (g.nthr + 1)) * g.hw + lroot];
    const int ap = area[((long long)b * (g.nthr + 1) + t + 1) * g.hw + proot];
    const float iou = __fdiv_rn((float)inter, (float)(al + ap - inter));     // metrics.py:189-191, float32
    if (iou > iou_threshold) {                                                // metrics.py:218, 235, 279, 283
      det_label[bt * g.hw + lroot] = 1;
      det_pred[bt * g.hw + proot] = 1;
    }
  }
}

// counts roots: [0] labels detected, [1] labels missed, [2] predictions without a hit, [3] predictions with a hit
__global__ void __launch_bounds__(256) region_count_kernel(RegionGeom g, const int* __restrict__ L,
                                                           const uint8_t* __restrict__ det_label,
                                                           const uint8_t* __restrict__ det_pred, int* __restrict__ per_slice,
                                                           unsigned long long* __restrict__ totals) {
  // grid = (chunks, n * nthr): a block stays inside one (slice, threshold)
  const long long bt = blockIdx.y;
  const int b = (int)(bt / g.nthr), t = (int)(bt % g.nthr);
  const int* PL = L + ((long long)b * (g.nthr + 1)) * g.hw;
  const int* PP = L + ((long long)b * (g.nthr + 1) + t + 1) * g.hw;
  const uint8_t* dl = det_label + bt * g.hw;
  const uint8_t* dp = det_pred + bt * g.hw;
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < g.hw; i += (long long)gridDim.x * blockDim.x) {
    if (PL[i] == (int)i) { if (dl[i]) ++c0; else ++c1; }
    if (PP[i] == (int)i) { if (dp[i]) ++c3; else ++c2; }
  }
  __shared__ int sh[4];
  if (threadIdx.x < 4) sh[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(sh + 0, c0);
    if (c1) atomicAdd(sh + 1, c1);
    if (c2) atomicAdd(sh + 2, c2);
    if (c3) atomicAdd(sh + 3, c3);
  }
  __syncthreads();
  if (threadIdx.x < 4 && sh[threadIdx.x]) {
    if (per_slice) atomicAdd(per_slice + ((long long)b * 4 + threadIdx.x) * g.nthr + t, sh[threadIdx.x]);
    if (totals) atomicAdd(totals + (long long)threadIdx.x * g.nthr + t, (unsigned long long)sh[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(long long total, long long hw, int* __restrict__ L) {
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += (long long)gridDim.x * blockDim.x) {
    int* P = L + (gi / hw) * hw;
    const int i = (int)(gi % hw);
    if (__ldcg(P + i) >= 0) P[i] = uf_find(P, i);
  }
}

struct RegionWorkspace {
  float* popen;
  int* L;
  int* area;                      // ---- zeroed from here on
  uint8_t* det_label;
  uint8_t* det_pred;
  unsigned long long* keys;
  unsigned* vals;
  size_t zero_bytes, total_bytes;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static RegionWorkspace carve(void* base, long long n, long long hw, int nthr, int slots) {
  RegionWorkspace r{};
  char* p = reinterpret_cast<char*>(base);
  size_t off = 0;
  r.popen = reinterpret_cast<float*>(p + off); off += align256((size_t)n * hw * 4);
  r.L = reinterpret_cast<int*>(p + off); off += align256((size_t)n * (nthr + 1) * hw * 4);
  const size_t zero_from = off;
  r.area = reinterpret_cast<int*>(p + off); off += align256((size_t)n * (nthr + 1) * hw * 4);
  r.det_label = reinterpret_cast<uint8_t*>(p + off); off += align256((size_t)n * nthr * hw);
  r.det_pred = reinterpret_cast<uint8_t*>(p + off); off += align256((size_t)n * nthr * hw);
  r.keys = reinterpret_cast<unsigned long long*>(p + off); off += align256((size_t)n * nthr * slots * 8);
  r.vals = reinterpret_cast<unsigned*>(p + off); off += align256((size_t)n * nthr * slots * 4);
  r.zero_bytes = off - zero_from;
  r.total_bytes = off;
  return r;
}

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_resize_bilinear(void* stream, const float* src, int n, int h, int w, float* dst, int oh, int ow) {
  DNNCA_CHECK_ARG(src && dst && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "resize_bilinear: bad arguments");
  resize_bilinear_kernel<<<grid_for((long long)n * oh * ow, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, n, h, w, dst, oh, ow);
  DNNCA_LAUNCH_CHECK("resize_bilinear");
  return DNNCA_OK;
}

extern "C" int dnnca_grey_open(void* stream, const float* src, int n, int h, int w, int filter_size, float* dst) {
  DNNCA_CHECK_ARG(src && dst && src != dst && n > 0 && h > 0 && w > 0, "grey_open: bad arguments");
  DNNCA_CHECK_ARG(filter_size >= 1 && filter_size <= OPEN_MAX_K, "grey_open: filter size 1..%d supported (got %d)", OPEN_MAX_K,
                  filter_size);
  const int halo = filter_size - 1, S0 = OPEN_TILE + 2 * halo, S1 = OPEN_TILE + halo;
  const size_t smem = (size_t)(S0 * S0 + S0 * S1) * 4;
  const long long tiles = (long long)n * ((h + OPEN_TILE - 1) / OPEN_TILE) * ((w + OPEN_TILE - 1) / OPEN_TILE);
  const long long cap = (long long)sm_count() * 16;
  grey_open_kernel<<<(int)(tiles < cap ? tiles : cap), 256, smem, (cudaStream_t)stream>>>(src, n, h, w, filter_size, dst);
  DNNCA_LAUNCH_CHECK("grey_open");
  return DNNCA_OK;
}

extern "C" int dnnca_connected_components(void* stream, const uint8_t* mask, int n, int h, int w, int32_t* roots) {
  DNNCA_CHECK_ARG(mask && roots && n > 0 && h > 0 && w > 0, "connected_components: bad arguments");
  const long long hw = (long long)h * w, total = hw * n;
  DNNCA_CHECK_ARG(hw < (1LL << 31), "connected_components: image too large");
  const int grid = grid_for(total, 256, 16);
  RegionGeom g{n, h, w, 0, hw};
  ccl_init_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, nullptr, mask, g, roots);
  DNNCA_LAUNCH_CHECK("ccl_init_mask");
  ccl_merge_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, roots);
  DNNCA_LAUNCH_CHECK("ccl_merge");
  ccl_flatten_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(total, hw, roots);
  DNNCA_LAUNCH_CHECK("ccl_flatten");
  return DNNCA_OK;
}

extern "C" size_t dnnca_region_workspace_bytes(int n, int h, int w, int nthr, int table_slots) {
  if (n <= 0 || h <= 0 || w <= 0 || nthr <= 0 || table_slots <= 0) return 0;
  return carve(nullptr, n, (long long)h * w, nthr, table_slots).total_bytes;
}

extern "C" int dnnca_region_confusion(void* stream, const float* labels, const float* probs, int n, int h, int w,
                                      const float* thresholds, int nthr, float iou_threshold, int morph_filter_size,
                                      void* workspace, size_t workspace_bytes, int table_slots, int32_t* per_slice,
                                      uint64_t* totals, int32_t* overflow) {
  DNNCA_CHECK_ARG(labels && probs && thresholds && workspace && overflow && (per_slice || totals), "region_confusion: bad arguments");
  DNNCA_CHECK_ARG(n > 0 && h > 0 && w > 0 && nthr > 0, "region_confusion: bad shape");
  DNNCA_CHECK_ARG(table_slots >= 64 && (table_slots & (table_slots - 1)) == 0, "region_confusion: table_slots must be a power of two >= 64");
  DNNCA_CHECK_ARG(morph_filter_size >= 1 && morph_filter_size <= OPEN_MAX_K, "region_confusion: morph filter 1..%d (got %d)", OPEN_MAX_K,
                  morph_filter_size);
  const long long hw = (long long)h * w;
  DNNCA_CHECK_ARG(hw < (1LL << 31) && (long long)n * (nthr + 1) < (1LL << 20), "region_confusion: problem too large for one call");
  DNNCA_CHECK_ARG((long long)n * nthr <= 65535, "region_confusion: n * nthr <= 65535 per call");
  RegionWorkspace ws = carve(workspace, n, hw, nthr, table_slots);
  DNNCA_CHECK_ARG(workspace_bytes >= ws.total_bytes, "region_confusion: workspace of %zu bytes needed (got %zu)", ws.total_bytes,
                  workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(ws.area, 0, ws.zero_bytes, st);
  if (e != cudaSuccess) return cuda_fail(e, "region_confusion memset");
  const float* popen = probs;
  if (morph_filter_size > 1) {
    int rc = dnnca_grey_open(stream, probs, n, h, w, morph_filter_size, ws.popen);
    if (rc != DNNCA_OK) return rc;
    popen = ws.popen;
  }
  RegionGeom g{n, h, w, nthr, hw};
  const long long total = (long long)n * (nthr + 1) * hw;
  const int grid = grid_for(total, 256, 16);
  ccl_init_kernel<false><<<grid, 256, 0, st>>>(labels, popen, thresholds, nullptr, g, ws.L);
  DNNCA_LAUNCH_CHECK("ccl_init");
  ccl_merge_kernel<<<grid, 256, 0, st>>>(g, ws.L);
  DNNCA_LAUNCH_CHECK("ccl_merge");
  ccl_flatten_pairs_kernel<<<grid, 256, 0, st>>>(g, ws.L, ws.area, ws.keys, ws.vals, table_slots, overflow);
  DNNCA_LAUNCH_CHECK("ccl_flatten_pairs");
  region_eval_kernel<<<grid_for((long long)n * nthr * table_slots, 256, 16), 256, 0, st>>>(g, ws.area, ws.keys, ws.vals, table_slots,
                                                                                       iou_threshold, ws.det_label, ws.det_pred);
  DNNCA_LAUNCH_CHECK("region_eval");
  int chunks = (int)((hw + 256 * 8 - 1) / (256 * 8));
  if (chunks < 1) chunks = 1;
  region_count_kernel<<<dim3(chunks, n * nthr), 256, 0, st>>>(g, ws.L, ws.det_label, ws.det_pred, per_slice,
                                                            reinterpret_cast<unsigned long long*>(totals));
  DNNCA_LAUNCH_CHECK("region_count");
  return DNNCA_OK;
}
