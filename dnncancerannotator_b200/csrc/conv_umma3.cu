// tcgen05 weight gradient of 3x3 convolutions with >= 64 channels, second generation: ONE halo tile of x per pixel
// tile serves several filter taps through row-shifted MN-major descriptors (K = pixels: a tap is the same buffer read
// from a different first pixel row), instead of one TMA load per tap.
//
// Why: wgrad_umma_kernel (conv_umma.cu) gives every (M tile, N tile, tap) its own CTA, so x and dz travel L2 -> SMEM
// nine times; at 64..128 channels that is ~77 B/clk/SM against ~40 B/clk/SM of L2 bandwidth: 208-245 TFLOP/s
// (profiles/r01b_bench_big.json).  Replaces the gradient of layers.Conv2D (components.py:47-50, :123-126) w.r.t. its kernel.
//
// Geometry: pixel tile = 8 rows x 16 px.  x box {64 ch, 18 px, rows} lands as pixel rows of 128 bytes (SWIZZLE_128B),
// pixel (r, c) at row r*18 + c; a K = 16 step is one tile row, for tap (dy, dx) it starts at row (r+dy)*18 + dx.
//   FULL   (Cin multiple of 128): M = 128 input channels = two 64-channel atoms (LBO = atom stride).  A CTA owns one dy
//          (its x box is loaded already shifted by dy) and keeps the three dx accumulators [128 x BN] in TMEM.
//   PAIRED (64-channel M tiles):  M = 128 = TWO TAPS of the same 64 channels -- the second 64-row block of the MMA is the
//          same atom LBO bytes further ((dy'-dy)*18 + dx'-dx pixel rows) -- so a CTA keeps all nine taps as 4 paired
//          accumulators + 1 single and x is loaded once per pixel tile.
//   TCONV  (ConvT 2x2/s2, Cin multiple of 128): M = 128 input channels of the x tile itself (no halo); the four filter
//          taps are four accumulators whose N operand is dy read through a stride-2 map at (2y + ty, 2x + tx): every
//          byte of x and dy enters shared memory once per (M tile, N tile) pair.
// dz box {64 ch, 16 px, 8 rows} is the N operand (MN-major, LBO = next 64 output channels).  Partial sums leave with
// red.global.add like the first-generation kernel; the bias gradient stays a separate channel-sum pass.
#include <math.h>
#include <stdlib.h>

#include "umma_common.cuh"

namespace dnnca {

constexpr int WH_R = 8, WH_PW = 18;
constexpr int WH_ZATOM = WH_R * 16 * 128;                                   // 16 KB

struct WHArgs {
  int c_a, c_b, cout;
  int tiles_x, tiles_y, nimg;
  int ksplit, mt_a;            // pixel-tile slices; M tiles that belong to x (the rest to x2)
  float* dw;
  float* db;                   // PAIRED: bias gradient from the all-ones block paired with the ninth tap (or NULL)
  int acc_major;               // experiment switch: MMA order accumulator-major (the first version) instead of K-step-major
};

__device__ __forceinline__ bool wh_elect() {
  uint32_t p;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}"
      : "=r"(p));
  return p != 0;
}
constexpr uint32_t WH_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t wh_desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__device__ __forceinline__ void wh_umma(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %6, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(WH_DESC_HI), "r"(leader)
      : "memory");
}

enum { WH_FULL = 0, WH_PAIRED = 1, WH_TCONV = 2 };

template <int BN, int MODE>
struct WHGeom {
  static constexpr bool PAIRED = MODE == WH_PAIRED, TCONV = MODE == WH_TCONV;
  static constexpr int PW = TCONV ? 16 : WH_PW;                // pixel rows of 128 bytes per image row of the x box
  static constexpr int XROWS = (PAIRED ? WH_R + 2 : WH_R) * PW;
  static constexpr int XATOM = (XROWS * 128 + 1023) & ~1023;
  static constexpr int NXA = PAIRED ? 1 : 2, NZA = BN / 64;
  static constexpr int ZTAPS = TCONV ? 4 : 1;                  // dz tiles per stage (ConvT: one per filter tap)
  static constexpr int STAGE = NXA * XATOM + ZTAPS * NZA * WH_ZATOM;
  static constexpr int NACC = PAIRED ? 5 : (TCONV ? 4 : 3);
  static constexpr int STAGES = (196 * 1024 - (PAIRED ? XATOM : 0)) / STAGE > 4 ? 4 : (196 * 1024 - (PAIRED ? XATOM : 0)) / STAGE;
  static constexpr int ONES = PAIRED ? XATOM : 0;              // all-ones block paired with the ninth tap: sum(dz) = db
  static constexpr int SMEM = 1024 + STAGES * STAGE + ONES + 1024;
  static constexpr int TCOLS = NACC * BN <= 256 ? 256 : 512;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(192) wgrad_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB,
                                                        const __grid_constant__ CUtensorMap mapG, const WHArgs a) {
  using G = WHGeom<BN, MODE>;
  constexpr bool PAIRED = MODE == WH_PAIRED, TCONV = MODE == WH_TCONV;
  constexpr int STAGES = G::STAGES, NACC = G::NACC;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + 8;
  uint64_t* accum = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  unsigned char* ring = smem + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool second = (int)blockIdx.x >= a.mt_a;
  constexpr int MCH = PAIRED ? 64 : 128;                                 // input channels per M tile
  const int m0 = (second ? (int)blockIdx.x - a.mt_a : (int)blockIdx.x) * MCH;
  const int n0 = blockIdx.y * BN;
  const int dy = MODE != WH_FULL ? 0 : (int)blockIdx.z / a.ksplit;       // FULL: this CTA's filter row (0..2)
  const int ks = MODE != WH_FULL ? (int)blockIdx.z : (int)blockIdx.z % a.ksplit;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  const int per = (ntiles + a.ksplit - 1) / a.ksplit;
  const int t_beg = ks * per, t_end = min(ntiles, t_beg + per);
  const int kiters = t_end - t_beg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TCOLS);
  unsigned char* ones = ring + STAGES * G::STAGE;
  if (PAIRED) {
    for (int i = threadIdx.x; i < G::ONES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (kiters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapG);
        int s = 0;
        uint32_t ph = 1;
        for (int it = 0; it < kiters; ++it) {
          int b = t_beg + it;
          const int tix = b % a.tiles_x; b /= a.tiles_x;
          const int tiy = b % a.tiles_y;
          const int n = b / a.tiles_y;
          const int x0 = tix * 16, y0 = tiy * WH_R;
          mbar_wait(empty + s, ph);
          unsigned char* st = ring + s * G::STAGE;
          mbar_expect_tx(full + s, (uint32_t)(G::NXA * G::XROWS * 128 + G::ZTAPS * G::NZA * WH_ZATOM));
          // x: pixels x0-1..x0+16; rows y0-1..y0+8 (PAIRED: all taps) or y0+dy-1..+7 (FULL: this CTA's filter row)
          if (TCONV) {
#pragma unroll
            for (int h = 0; h < G::NXA; ++h) tma_load_4d(st + h * G::XATOM, &mapA, full + s, m0 + 64 * h, x0, y0, n);
#pragma unroll
            for (int tap = 0; tap < 4; ++tap)
#pragma unroll
              for (int h = 0; h < G::NZA; ++h)
                tma_load_4d(st + G::NXA * G::XATOM + (tap * G::NZA + h) * WH_ZATOM, &mapG, full + s, n0 + 64 * h, 2 * x0 + (tap & 1),
                            2 * y0 + (tap >> 1), n);
          } else {
#pragma unroll
            for (int h = 0; h < G::NXA; ++h)
              tma_load_4d(st + h * G::XATOM, second ? &mapB : &mapA, full + s, m0 + 64 * h, x0 - 1, y0 - 1 + dy, n);
#pragma unroll
            for (int h = 0; h < G::NZA; ++h) tma_load_4d(st + G::NXA * G::XATOM + h * WH_ZATOM, &mapG, full + s, n0 + 64 * h, x0, y0, n);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    } else if (warp == 1) {
      const uint32_t leader = wh_elect() ? 1u : 0u;
      const bool committer = wh_elect();
      const uint32_t idesc128 = make_idesc(128, BN, 1, 1);
      const uint32_t ring_addr = smem_u32(ring);
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t xaddr = ring_addr + (uint32_t)(s * G::STAGE);
        const uint32_t z_lo0 = wh_desc_lo(xaddr + G::NXA * G::XATOM, WH_ZATOM);
        const uint32_t accf = it ? 1u : 0u;
        // per accumulator: first pixel row of its (first) tap and the distance to its second 64-row block
        uint32_t x_lo0[NACC], z_lo[NACC];
#pragma unroll
        for (int j = 0; j < NACC; ++j) {
          uint32_t row0, lbo;
          if (PAIRED) {
            const int ta = 2 * j, tb = 2 * j + 1;
            row0 = (uint32_t)((ta / 3) * WH_PW + ta % 3);
            if (tb < 9) lbo = (uint32_t)((((tb / 3) - (ta / 3)) * WH_PW + (tb % 3) - (ta % 3)) * 128);
            else lbo = smem_u32(ones) - (xaddr + row0 * 128);   // ninth tap | all-ones block: rows 64.. = sum(dz)
          } else {
            row0 = TCONV ? 0u : (uint32_t)j;          // FULL: dx = j (the box is already shifted by dy)
            lbo = G::XATOM;
          }
          x_lo0[j] = wh_desc_lo(xaddr + row0 * 128, lbo);
          z_lo[j] = TCONV ? z_lo0 + (uint32_t)(j * G::NZA * (WH_ZATOM >> 4)) : z_lo0;     // ConvT: tap j's dy tile
        }
        // K step outermost, accumulators innermost: consecutive MMAs go to DIFFERENT accumulators (DNNCA_WGRAD_ORDER=1
        // restores accumulator-major order, 8 dependent MMAs in a row)
        if (!a.acc_major) {
#pragma unroll
          for (int r = 0; r < WH_R; ++r)               // one tile row = 16 pixels = one K step
#pragma unroll
            for (int j = 0; j < NACC; ++j)
              wh_umma(tmem_base + (uint32_t)(j * BN), x_lo0[j] + (uint32_t)(r * G::PW * 8), z_lo[j] + (uint32_t)(r * 128), idesc128,
                      r ? 1u : accf, leader);
        } else {
#pragma unroll
          for (int j = 0; j < NACC; ++j)
#pragma unroll
            for (int r = 0; r < WH_R; ++r)
              wh_umma(tmem_base + (uint32_t)(j * BN), x_lo0[j] + (uint32_t)(r * G::PW * 8), z_lo[j] + (uint32_t)(r * 128), idesc128,
                      r ? 1u : accf, leader);
        }
        if (committer) umma_commit(empty + s);
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (committer) umma_commit(accum);
    } else {
      // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4).. =====
      const int lg = warp & 3;
      const int cin = a.c_a + a.c_b;
      const int ctens = second ? a.c_b : a.c_a;               // channels of the tensor this M tile belongs to
      const int coff = second ? a.c_a : 0;
      mbar_wait(accum, 0);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < NACC; ++j) {
        int tap, ch;
        bool live;
        bool db_row = false;
        if (PAIRED) {
          const int row = lg * 32 + lane;
          tap = 2 * j + (row >> 6); ch = row & 63; live = tap < 9;
          db_row = tap == 9 && ch == 0 && a.db != nullptr && blockIdx.x == 0;      // first row of the all-ones half
          if (tap > 8) tap = 8;
        } else {
          tap = TCONV ? j : dy * 3 + j; ch = lg * 32 + lane; live = true;
        }
        live = live && (m0 + ch) < ctens;
        // conv: dw[tap][ci][co] (a thread's row is contiguous); ConvT: dk[tap][co][ci] (the warp's lanes are contiguous)
        float* dst_row = TCONV ? a.dw + ((size_t)tap * a.cout + n0) * cin + m0 + ch : a.dw + ((size_t)tap * cin + coff + m0 + ch) * a.cout + n0;
        const size_t estep = TCONV ? (size_t)cin : 1;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(j * BN + c0), v);
          tmem_ld_wait();
          if (db_row) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (n0 + c0 + e < a.cout) atomicAdd(a.db + n0 + c0 + e, __uint_as_float(v[e]));
          }
          if (!live) continue;
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (n0 + c0 + e < a.cout) atomicAdd(dst_row + (size_t)(c0 + e) * estep, __uint_as_float(v[e]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, G::TCOLS);
}

// 4-D {C, W, H, N} view; box {64 channels, px, rows, 1}, SWIZZLE_128B; channels / pixels outside read as zero
static bool wh_map(CUtensorMap* m, const dnnca_tensor_t* t, int px, int rows, int estride = 1) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  char* base = reinterpret_cast<char*>(t->data) + (size_t)t->coff * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (t->cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
  // with a traversal stride the box extent is given in tensor elements (pixels * stride)
  cuuint32_t box[4] = {64, (cuuint32_t)(px * estride), (cuuint32_t)(rows * estride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int MODE>
static int launch_wgrad_halo(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mG, WHArgs a, int mt,
                             int nt) {
  using G = WHGeom<BN, MODE>;
  auto kern = wgrad_halo_kernel<BN, MODE>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "wgrad_halo: cudaFuncSetAttribute");
    done = true;
  }
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  const int groups = MODE == WH_FULL ? 3 : 1;
  // pixel slices so that the grid is ONE resident wave (1 CTA per SM: TMEM and smem): every extra slice costs a full
  // red.global.add flush of the accumulators (measured: 2 waves = 1.4x slower at 64 channels)
  long long want = (long long)sm_count() / ((long long)mt * nt * groups);
  if (want > ntiles / 4) want = ntiles / 4;
  // small problems (mulmo_unet's 64x64 / 32x32 levels): with k slices the CTAs spend ~ntiles/k tile times computing and
  // the flush pushes k * (CTAs per slice) * (accumulator elements) fp32 atomics through L2 (~28 per clock measured:
  // 64->64@64, B=32 with 148 slices = 111 us, of which 100 us flush), so k* = sqrt(compute / flush-per-slice)
  {
    const double t_tile = MODE == WH_PAIRED ? 2500.0 : (BN > 64 ? 3600.0 : 1900.0);        // MMA cycles per pixel tile
    const double flush_per_slice = (double)mt * nt * groups * (G::NACC * 128.0 * BN) / 28.0;
    const long long kopt = (long long)(sqrt((double)ntiles * t_tile / flush_per_slice) + 0.5);
    if (want > kopt) want = kopt;
  }
  if (want < 1) want = 1;
  if (groups * want > 65535) want = 65535 / groups;
  a.ksplit = (int)want;
  a.acc_major = getenv("DNNCA_WGRAD_ORDER") ? atoi(getenv("DNNCA_WGRAD_ORDER")) : 0;
  dim3 grid((unsigned)mt, (unsigned)nt, (unsigned)(groups * a.ksplit));
  kern<<<grid, 192, G::SMEM, s>>>(mA, mB, mG, a);
  DNNCA_LAUNCH_CHECK("wgrad_halo");
  note_family(2);
  return 1;
}

// Conv2D 3x3 wgrad for bf16 views whose channel counts are multiples of 64; returns 1 / 0 (not covered) / <0
// returns 2 when the bias gradient was accumulated into `db` as well (PAIRED variant)
int try_conv3x3_wgrad_halo(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* g, float* dw,
                           float* db) {
  const int ca = x->c, cb = x2 ? x2->c : 0, cout = g->c;
  // tensors with fewer than 64 channels (first layers: 1/3/5 modalities in 16-byte pixels; the 16/32-channel layers of
  // mulmo_unet.yaml) are ONE PAIRED M tile each whose missing channels the TMA zero-fills; likewise the last N tile
  const bool narrow_a = ca < 64, narrow_b = x2 && cb < 64;
  if ((!narrow_a && ca % 64) || (x2 && !narrow_b && cb % 64) || cout % 16) return 0;
  // PAIRED (64-channel M tiles, all nine taps per CTA) measured faster up to 128 input channels per tensor; above
  // that the FULL variant (128-channel M tiles, one filter row per CTA) wins (tools/wgrad_microbench.py)
  static int paired_max = -1;
  if (paired_max < 0) paired_max = getenv("DNNCA_WGRAD_PAIRED_MAX") ? atoi(getenv("DNNCA_WGRAD_PAIRED_MAX")) : 128;
  const bool paired = (ca % 128 != 0) || (cb % 128 != 0) || (ca <= paired_max && cb <= paired_max) || cout % 64 != 0;
  WHArgs a{};
  a.c_a = ca; a.c_b = cb; a.cout = cout; a.dw = dw;
  a.tiles_x = (x->w + 15) / 16; a.tiles_y = (x->h + WH_R - 1) / WH_R; a.nimg = x->n;
  const int mch = paired ? 64 : 128;
  a.mt_a = (ca + mch - 1) / mch;
  const int mt = a.mt_a + (cb + mch - 1) / mch;
  CUtensorMap mA, mB, mG;
  if (!wh_map(&mA, x, WH_PW, paired ? WH_R + 2 : WH_R)) return 0;
  mB = mA;
  if (x2 && !wh_map(&mB, x2, WH_PW, paired ? WH_R + 2 : WH_R)) return 0;
  if (!wh_map(&mG, g, 16, WH_R)) return 0;
  if (paired) {
    a.db = db;
    const int r = launch_wgrad_halo<64, WH_PAIRED>(s, mA, mB, mG, a, mt, (cout + 63) / 64);
    return (r == 1 && db) ? 2 : r;
  }
  if (cout % 128 == 0) return launch_wgrad_halo<128, WH_FULL>(s, mA, mB, mG, a, mt, cout / 128);
  return launch_wgrad_halo<64, WH_FULL>(s, mA, mB, mG, a, mt, cout / 64);
}

// ConvT 2x2/s2 wgrad dk[tap][co][ci] for bf16 views with Cin % 128 == 0 and Cout % 64 == 0; returns 1 / 0 / <0
int try_tconv_wgrad_halo(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk) {
  const int cin = x->c, cout = dy->c;
  if (cin % 128 || cout % 64) return 0;
  WHArgs a{};
  a.c_a = cin; a.c_b = 0; a.cout = cout; a.dw = dk;
  a.tiles_x = (x->w + 15) / 16; a.tiles_y = (x->h + WH_R - 1) / WH_R; a.nimg = x->n;
  a.mt_a = cin / 128;
  CUtensorMap mA, mG;
  if (!wh_map(&mA, x, 16, WH_R)) return 0;
  if (!wh_map(&mG, dy, 16, WH_R, 2)) return 0;
  return launch_wgrad_halo<64, WH_TCONV>(s, mA, mA, mG, a, a.mt_a, cout / 64);
}

}  // namespace dnnca
