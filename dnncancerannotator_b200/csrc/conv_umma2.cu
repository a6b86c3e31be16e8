// tcgen05 implicit-GEMM 3x3 convolution, second generation: persistent CTAs, ONE halo tile per channel chunk
// instead of nine shifted tiles, weights resident in shared memory when they fit, double-buffered TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.  Serves Conv2D fprop and dgrad for inputs whose channel
// counts are multiples of 64 (configs/unet_big.yaml: every layer but the first).
//
// Why: the first-generation kernel (conv_umma.cu) re-loads the activation tile for each of the 9 taps and the
// weight tile for every pixel tile.  At 64-channel layers that is 216 KB of L2->SMEM traffic per 1152 tensor-pipe
// cycles (187 B/clk/SM against ~45 B/clk/SM of L2 bandwidth): ncu shows 4-20 % tensor-pipe activity there
// (profiles/r01_ncu_full_umma_unet_big_b8.csv).
//
// Geometry: M tile = 16 rows x 8 pixels.  The halo box {64 ch, 10 px, 18 rows} lands as pixel rows of 128 bytes
// (SWIZZLE_128B), 10 pixels per image row, so the A operand of tap (dy,dx) is the SAME buffer read through a
// descriptor whose start address is shifted by (dy*10 + dx) rows: the 8 pixels of a tile row form one 8-row core
// group, consecutive tile rows are SBO = 10*128 = 1280 bytes apart (the 128-byte swizzle is a function of the
// shared-memory address, for the TMA write and the MMA read alike, so neither the group start nor the group stride
// needs to be a multiple of the 1024-byte swizzle period).
#include <stdlib.h>

#include "umma_common.cuh"

// -DDNNCA_HALO_TRACE: per-role wait counters printed by CTA (0,0) when DNNCA_HALO_DBG has bit 16 (clock reads are
// long-scoreboard instructions: they cost the epilogue ~10 % when compiled in)
#ifdef DNNCA_HALO_TRACE
#define HALO_CLOCK() clock64()
#define HALO_PRINT(...) printf(__VA_ARGS__)
#else
#define HALO_CLOCK() 0ll
#define HALO_PRINT(...) ((void)0)
#endif

namespace dnnca {

// TAPS = 9: 3x3 convolution over a halo tile.  TAPS = 1: a plain GEMM over the 16x8 pixel tile itself (box {64 ch, 8 px,
// 16 rows}, no halo), K = a.taps * c_a in 64-channel chunks:
//   ConvT 2x2/s2 fprop (a.taps = 1): the four filter taps are N columns [tap*Cout, (tap+1)*Cout) scattered by the
//     epilogue to output pixel (2y + tap/2, 2x + tap%2);
//   ConvT 2x2/s2 dgrad (a.taps = 4, a.sx = 2): chunk (tap, kc) reads dy through a map with traversal stride 2 at
//     origin (2x0 + tap%2, 2y0 + tap/2) and the weight block [tap][Cin tile][kc].
// With so little K per tile the activation ring is four slots deep so that loads run two or more tiles ahead.
template <int BN, int TAPS = 9>
struct HGeom {
  static constexpr int HPX = TAPS == 9 ? 10 : 8;           // pixels per image row of the activation box
  static constexpr int A_BYTES = (TAPS == 9 ? 18 : 16) * HPX * 128;         // bytes one box load delivers (22.5 / 16 KB)
  static constexpr int A_SLOT = (A_BYTES + 1023) & ~1023;  // ring slot (1024-byte aligned for the swizzle period)
  static constexpr int A_SBO = HPX * 128;                  // bytes between consecutive tile rows (8-pixel core groups)
  static constexpr int B_TAP = BN * 128;                   // one tap of one 64-channel chunk
  static constexpr int A_SLOTS = TAPS == 9 ? 3 : 4;
  // control block: barriers + TMEM slot (256 B) | bias table: 9 border classes x BN floats (3x3 fprop with a folded
  // BatchNorm on its input: the bias depends on which taps fall into the zero padding) | BN-statistics accumulators
  // (2*BN doubles)
  static constexpr int BIAS_OFF = 256;
  static constexpr int STAT_OFF = (BIAS_OFF + 9 * BN * 4 + 15) & ~15;
  static constexpr int CTRL = (STAT_OFF + 2 * BN * 8 + 1023) & ~1023;
  // resident: all TAPS * (Cin/64) weight blocks stay in shared memory for the CTA's lifetime
  static constexpr int smem_resident(int kchunks) { return CTRL + A_SLOTS * A_SLOT + TAPS * kchunks * B_TAP + 1024; }
  static constexpr int B_STAGES = BN > 128 ? 3 : 4;
  static constexpr int SMEM_STREAM = CTRL + A_SLOTS * A_SLOT + B_STAGES * B_TAP + 1024;
};

// K-major SWIZZLE_128B descriptor with an explicit stride between 8-row groups and a base offset (rows the start
// address is displaced inside the 1024-byte swizzle period)
__device__ __forceinline__ uint64_t kmajor128_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, bool RESIDENT, int TAPS, int EPI>
__global__ void __launch_bounds__(320) conv_umma_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const __grid_constant__ CUtensorMap mapW, UArgs a) {
  using G = HGeom<BN, TAPS>;
  constexpr int B_STAGES = RESIDENT ? 1 : G::B_STAGES;
  constexpr int A_SLOT = G::A_SLOT, A_SLOTS = G::A_SLOTS;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);      // [A_SLOTS]
  uint64_t* emptyA = fullA + 4;                              // [A_SLOTS]
  uint64_t* fullB = emptyA + 4;                              // [B_STAGES] (resident: [1], completes once)
  uint64_t* emptyB = fullB + 8;                              // [B_STAGES]
  uint64_t* tfull = emptyB + 8;                              // [2] accumulator ready
  uint64_t* tempty = tfull + 2;                              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // bias slice of this CTA's N tile per border class (row class * 3 + column class; class 0 = first row / column,
  // 1 = interior, 2 = last): all nine rows equal `bias` unless a BatchNorm is folded into the input (UArgs::bias9)
  constexpr int NCLS = (EPI == EPI_FPROP && TAPS == 9) ? 9 : 1;
  float* sbias = reinterpret_cast<float*>(smem + G::BIAS_OFF);
  double* sstat = reinterpret_cast<double*>(smem + G::STAT_OFF);   // [2*BN] per-channel sum | sum of squares (fprop + stats)
  const bool do_stats = EPI == EPI_FPROP && a.stats != nullptr;
  unsigned char* aring = smem + G::CTRL;
  unsigned char* bring = aring + G::A_SLOTS * A_SLOT;

  int tid;                                  // volatile: the compiler otherwise re-reads SR_TID.X inside the tile loops
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5, lane = tid & 31;
  const int kca = a.c_a / 64, kcb = a.c_b / 64, kchunks = TAPS == 9 ? kca + kcb : a.taps * kca;
  const int n0 = blockIdx.y * BN;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_SLOTS; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(fullB + s, 1); mbar_init(emptyB + s, 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  for (int i = threadIdx.x; i < NCLS * BN; i += blockDim.x) {
    const int cls = i / BN, col = i - cls * BN;
    float b = 0.f;
    if (EPI != EPI_DGRAD && n0 + col < a.n_total) {
      if (NCLS == 9 && a.bias9) b = a.bias9[cls * a.n_total + n0 + col];
      else if (a.bias) b = a.bias[EPI == EPI_TCONV ? (n0 + col) % a.cout_t : n0 + col];
    }
    sbias[i] = b;
  }
  for (int i = threadIdx.x; i < 2 * BN; i += blockDim.x) sstat[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapW);
      if (RESIDENT) {                       // every weight block once: [tap][kc] blocks of BN x 64
        mbar_expect_tx(fullB, (uint32_t)(TAPS * kchunks * G::B_TAP));
        for (int kc = 0; kc < kchunks; ++kc)
          for (int tap = 0; tap < TAPS; ++tap)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(bring + (kc * TAPS + tap) * G::B_TAP)), "l"(reinterpret_cast<uint64_t>(&mapW)), "r"(smem_u32(fullB)),
                "r"((TAPS == 9 ? kc : kc % kca) * 64), "r"(n0), "r"(TAPS == 9 ? tap : kc / kca)
                : "memory");
      }
      int ai = 0, bi = 0;
      long long tw = 0, t00 = HALO_CLOCK();
      TileWalk tw_(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, tw_.next()) {
        const int n = tw_.n;
        const int x0 = tw_.tix * 8, y0 = tw_.tiy * 16;
        int ktap = 0, kcc = 0;                // TAPS == 1: chunk kc = (filter tap, 64-channel chunk inside it)
        for (int kc = 0; kc < kchunks; ++kc, ++ai) {
          const int s = ai % A_SLOTS;
          const long long tq = HALO_CLOCK();
          if (ai >= A_SLOTS) mbar_wait(emptyA + s, ((ai / A_SLOTS) - 1) & 1);
          tw += HALO_CLOCK() - tq;
          mbar_expect_tx(fullA + s, G::A_BYTES);
          if (TAPS == 9) {
            const bool second = kc >= kca;
            tma_load_4d(aring + s * A_SLOT, second ? &mapB : &mapA, fullA + s, (second ? kc - kca : kc) * 64, x0 - 1, y0 - 1, n);
          } else {
            tma_load_4d(aring + s * A_SLOT, &mapA, fullA + s, kcc * 64, a.sx * x0 + (ktap & 1), a.sx * y0 + (ktap >> 1), n);
          }
          if (!RESIDENT) {
            for (int tap = 0; tap < TAPS; ++tap, ++bi) {
              const int sb = bi % B_STAGES;
              if (bi >= B_STAGES) mbar_wait(emptyB + sb, ((bi / B_STAGES) - 1) & 1);
              mbar_expect_tx(fullB + sb, G::B_TAP);
              asm volatile(
                  "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  ::"r"(smem_u32(bring + sb * G::B_TAP)), "l"(reinterpret_cast<uint64_t>(&mapW)), "r"(smem_u32(fullB + sb)),
                  "r"((TAPS == 9 ? kc : kcc) * 64), "r"(n0), "r"(TAPS == 9 ? tap : ktap)
                  : "memory");
            }
          }
          if (++kcc == kca) { kcc = 0; ++ktap; }
        }
      }
      if ((a.dbg & 16) && blockIdx.x == 0 && blockIdx.y == 0) HALO_PRINT("producer: total %lld wait emptyA %lld loads %d\n", HALO_CLOCK() - t00, tw, ai);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (descriptors in uniform registers) =====
    {
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      constexpr uint32_t a_hi = kmajor128_desc_hi(G::A_SBO), b_hi = kmajor128_desc_hi(1024);
      const uint32_t leader = (umma_elect() && !(a.dbg & 2)) ? 1u : 0u;
      const bool committer = umma_elect();
      const uint32_t aring_lo = kmajor128_desc_lo(smem_u32(aring)), bring_lo = kmajor128_desc_lo(smem_u32(bring));
      if (RESIDENT) { mbar_wait(fullB, 0); }
      int ai = 0, bi = 0, ti = 0;
      long long twa = 0, twt = 0, t00 = HALO_CLOCK();
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
        const int buf = ti & 1;
        const long long tq = HALO_CLOCK();
        if (ti >= 2) mbar_wait(tempty + buf, ((ti >> 1) - 1) & 1);    // epilogue has drained this accumulator
        twt += HALO_CLOCK() - tq;
        tc_fence_after();
        const uint32_t dtm = tmem_base + (uint32_t)(buf * BN);
        for (int kc = 0; kc < kchunks; ++kc, ++ai) {
          const int s = ai % A_SLOTS;
          const long long tq2 = HALO_CLOCK();
          mbar_wait(fullA + s, (ai / A_SLOTS) & 1);
          twa += HALO_CLOCK() - tq2;
          tc_fence_after();
          // descriptors differ only in the 14-bit start-address field: one base per operand plus immediates
          const uint32_t a_base = aring_lo + (uint32_t)(s * (A_SLOT >> 4));
          if (RESIDENT) {
            const uint32_t b_base = bring_lo + (uint32_t)(kc * TAPS * (G::B_TAP >> 4));
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
              const uint32_t da0 = a_base + (uint32_t)(TAPS == 9 ? ((tap / 3) * G::HPX + (tap % 3)) * 8 : 0);      // (dy*HPX+dx) rows of 128 B, >>4
              const uint32_t db0 = b_base + (uint32_t)(tap * (G::B_TAP >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_uniform(dtm, da0 + 2 * k, a_hi, db0 + 2 * k, b_hi, idesc, (kc | tap | k) ? 1u : 0u, leader);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap, ++bi) {
              const int st = bi % B_STAGES;
              mbar_wait(fullB + st, (bi / B_STAGES) & 1);
              tc_fence_after();
              const uint32_t db0 = bring_lo + (uint32_t)(st * (G::B_TAP >> 4));
              const uint32_t da0 = a_base + (uint32_t)(TAPS == 9 ? ((tap / 3) * G::HPX + (tap % 3)) * 8 : 0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_uniform(dtm, da0 + 2 * k, a_hi, db0 + 2 * k, b_hi, idesc, (kc | tap | k) ? 1u : 0u, leader);
              if (committer) umma_commit(emptyB + st);
              __syncwarp();
            }
          }
          if (committer) umma_commit(emptyA + s);          // tile slot free once these MMAs retire
          __syncwarp();
        }
        if (committer) umma_commit(tfull + buf);           // accumulator of this tile complete
        __syncwarp();
      }
      if ((a.dbg & 16) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
        HALO_PRINT("mma: total %lld wait fullA %lld wait tempty %lld tiles %d\n", HALO_CLOCK() - t00, twa, twt, ti);
    }
  } else {
    // ===== epilogue: warps 2..9.  A warp may only touch TMEM lanes 32*(warp%4)..+31, so two warps share each lane
    // quarter and split the columns; the mask row (dgrad) is prefetched before the accumulator is ready =====
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;       // 0: columns [0, BN/2), 1: [BN/2, BN)
    const int r = lg * 32 + lane;           // row of the 16x8 pixel tile
    const int ty = r >> 3, tx = r & 7;
    constexpr int NCHUNK = BN / 64;         // 32-column chunks per warp
    // BN statistics: for BN = 64 every thread keeps per-column partial sums of ITS tile row in registers over all of
    // the CTA's tiles and the warp reduction runs once at the end; wider tiles reduce per tile (register budget)
    constexpr bool REG_STATS = BN <= 64 && EPI == EPI_FPROP;
    constexpr int NMASK = EPI == EPI_DGRAD ? NCHUNK : 1;     // mask rows are prefetched only by dgrad
    float acc1[REG_STATS ? NCHUNK : 1][32], acc2[REG_STATS ? NCHUNK : 1][32];
#pragma unroll
    for (int c = 0; c < (REG_STATS ? NCHUNK : 1); ++c)
#pragma unroll
      for (int j = 0; j < 32; ++j) { acc1[c][j] = 0.f; acc2[c][j] = 0.f; }
    int ti = 0;
    long long twf = 0, t00 = HALO_CLOCK();
    TileWalk tw_(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti, tw_.next()) {
      const int n = tw_.n;
      const int gy = tw_.tiy * 16 + ty, gx = tw_.tix * 8 + tx;
      const bool inside = gy < a.H && gx < a.W;
      const int buf = ti & 1;
      const float* sb = sbias;
      if (NCLS == 9) sb += ((gy == 0 ? 0 : (gy == a.H - 1 ? 2 : 1)) * 3 + (gx == 0 ? 0 : (gx == a.W - 1 ? 2 : 1))) * BN;
      ChunkAddr ca[NCHUNK];
      uint4 m[NMASK][4];
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col = half * (BN / 2) + c * 32;
        ca[c] = chunk_addr<EPI>(a, n, inside ? gy : 0, inside ? gx : 0, n0 + col);
        if (EPI == EPI_DGRAD && ca[c].msk && inside && n0 + col < a.n_total) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m[EPI == EPI_DGRAD ? c : 0][q] = reinterpret_cast<const uint4*>(ca[c].msk)[q];
        }
      }
      const long long tq = HALO_CLOCK();
      mbar_wait(tfull + buf, (ti >> 1) & 1);
      twf += HALO_CLOCK() - tq;
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col = half * (BN / 2) + c * 32;
        uint32_t v[32];
        if (!(a.dbg & 4)) {
          tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN + col), v);
          tmem_ld_wait();
        }
        if (EPI != EPI_FPROP || !do_stats) {
          if (inside && n0 + col < a.n_total)
            epilogue_chunk32<false, EPI>(a, v, ca[c], m[EPI == EPI_DGRAD ? c : 0], sb + col, !(a.dbg & 1),
                                         a.n_total - (n0 + col));
        } else {
          // BatchNormalization statistics of the stored tensor (components.py:57-58,130-132) in the epilogue: rows of
          // the tile are lanes, so a column sum is a 31-shuffle warp reduction; one shared fp64 atomic per lane
          epilogue_chunk32<true, EPI>(a, v, ca[c], m[0], sb + col, inside && n0 + col < a.n_total, a.n_total - (n0 + col));
          if (REG_STATS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float r = inside ? __uint_as_float(v[j]) : 0.f;
              acc1[REG_STATS ? c : 0][j] += r;
              acc2[REG_STATS ? c : 0][j] = fmaf(r, r, acc2[REG_STATS ? c : 0][j]);
            }
          } else {
            float r1[32], r2[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              r1[j] = inside ? __uint_as_float(v[j]) : 0.f;
              r2[j] = r1[j] * r1[j];
            }
            warp_colsum32(r1, lane);
            warp_colsum32(r2, lane);
            if (n0 + col + lane < a.n_total) {
              atomicAdd(sstat + col + lane, (double)r1[0]);
              atomicAdd(sstat + BN + col + lane, (double)r2[0]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + buf);
    }
    if ((a.dbg & 16) && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 2 || warp == 9))
      HALO_PRINT("epi warp %d: total %lld wait tfull %lld\n", warp, HALO_CLOCK() - t00, twf);
    if (REG_STATS && do_stats) {
#pragma unroll
      for (int c = 0; c < (REG_STATS ? NCHUNK : 1); ++c) {
        const int col = half * (BN / 2) + c * 32;
        warp_colsum32(acc1[c], lane);
        warp_colsum32(acc2[c], lane);
        if (n0 + col + lane < a.n_total) {
          atomicAdd(sstat + col + lane, (double)acc1[c][0]);
          atomicAdd(sstat + BN + col + lane, (double)acc2[c][0]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
  if (do_stats)
    for (int i = threadIdx.x; i < BN; i += blockDim.x)
      if (n0 + i < a.n_total) {
        atomicAdd(a.stats + n0 + i, sstat[i]);
        atomicAdd(a.stats + a.n_total + n0 + i, sstat[BN + i]);
      }
}

// ---------------------------------------------------------------- host side
static bool halo_map(CUtensorMap* m, const dnnca_tensor_t* t, int box_w = 10, int box_h = 18, int estride = 1) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  char* base = reinterpret_cast<char*>(t->data) + (size_t)t->coff * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (t->cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
  // with a traversal stride the box extent is given in tensor elements (pixels * stride)
  cuuint32_t box[4] = {64, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool weight_map64(CUtensorMap* m, const void* wp, int ktot, int ntot, int bn, int taps = 9) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)ntot, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * ntot * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wp), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool RESIDENT, int TAPS, int EPI>
static int launch_halo_epi(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                           int kchunks) {
  using G = HGeom<BN, TAPS>;
  const int smem = RESIDENT ? G::smem_resident(kchunks) : G::SMEM_STREAM;
  auto kern = conv_umma_halo_kernel<BN, RESIDENT, TAPS, EPI>;
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_umma_halo: cudaFuncSetAttribute");
    smem_set = smem;
  }
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  const int nt = (a.n_total + BN - 1) / BN;
  int per = sm_count() / nt;                 // persistent CTAs per N tile (1 CTA per SM: smem + 2*BN TMEM columns)
  if (per < 1) per = 1;
  if (per > ntiles) per = ntiles;
  dim3 grid((unsigned)per, (unsigned)nt);
  kern<<<grid, 320, smem, s>>>(mA, mB, mW, a);
  DNNCA_LAUNCH_CHECK("conv_umma_halo");
  note_family(2);
  return 1;
}

// the epilogue kind is a template parameter (the fprop instantiation carries BatchNorm-statistics registers, the dgrad one
// the prefetched mask rows; together they do not fit the 168 registers 320 threads leave)
template <int BN, bool RESIDENT, int TAPS = 9>
static int launch_halo(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                       int kchunks) {
  if (a.epi == EPI_DGRAD) return launch_halo_epi<BN, RESIDENT, TAPS, EPI_DGRAD>(s, mA, mB, mW, a, kchunks);
  if (TAPS == 9 || a.epi == EPI_FPROP) return launch_halo_epi<BN, RESIDENT, TAPS, EPI_FPROP>(s, mA, mB, mW, a, kchunks);   // TAPS 1 + FPROP: 1x1 conv
  if (BN == 256) return launch_halo_epi<256, RESIDENT, TAPS, EPI_TCONV>(s, mA, mB, mW, a, kchunks);
  return 0;
}

// Conv2D 3x3 fprop (pack mode 0 already in `wpack`: [tap][N][K]) or dgrad (pack mode 1); returns 1 / 0 / <0
int try_conv3x3_halo(cudaStream_t s, const dnnca_tensor_t* xa, const dnnca_tensor_t* xb, const void* wpack, int ktot, int ntot,
                     UArgs a) {
  // a single input with fewer than 64 channels (first layers) is one K chunk whose missing channels are zero-filled by
  // the TMA (activations and packed weights alike)
  const bool narrow = !xb && xa->c < 64;
  // K: two inputs need whole 64-channel chunks each; a single input of any width runs ceil(C/64) chunks whose tail
  // channels the TMA zero-fills (activations out of bounds, weights packed with K padded to `ktot`)
  if (xb && (xa->c % 64 || xb->c % 64)) return 0;
  if (!narrow && ktot % 64) return 0;
  // N (output channels) in multiples of 8: the last N tile may be partly empty (the weight TMA zero-fills the missing
  // rows, the epilogue stores 8-column groups below n_total only); dgrad routes whole 32-column chunks to its two
  // destinations and keeps the multiple-of-32 rule
  if (ntot % 8) return 0;
  if (a.epi == EPI_DGRAD && (ntot % 32 || a.split % 32)) return 0;
  const int bn = ntot % 256 == 0 ? 256 : (ntot % 128 == 0 ? 128 : (ntot <= 64 ? 64 : 128));
  CUtensorMap mA, mB, mW;
  if (!halo_map(&mA, xa)) return 0;
  mB = mA;
  if (xb && !halo_map(&mB, xb)) return 0;
  if (!weight_map64(&mW, wpack, ktot, ntot, bn)) return 0;
  a.tiles_x = (a.W + 7) / 8;
  a.tiles_y = (a.H + 15) / 16;
  a.dbg = getenv("DNNCA_HALO_DBG") ? atoi(getenv("DNNCA_HALO_DBG")) : 0;
  const int kchunks = narrow ? 1 : ktot / 64;
  if (narrow) a.c_a = 64;
  const size_t limit = 226 * 1024;
  if (bn == 64) {
    if ((size_t)HGeom<64>::smem_resident(kchunks) <= limit) return launch_halo<64, true>(s, mA, mB, mW, a, kchunks);
    return launch_halo<64, false>(s, mA, mB, mW, a, kchunks);
  }
  if (bn == 128) {
    if ((size_t)HGeom<128>::smem_resident(kchunks) <= limit) return launch_halo<128, true>(s, mA, mB, mW, a, kchunks);
    return launch_halo<128, false>(s, mA, mB, mW, a, kchunks);
  }
  return launch_halo<256, false>(s, mA, mB, mW, a, kchunks);
}

// Conv2D 1x1 fprop (pack mode 0 with one tap already in `wpack`: [N][K]) as a plain GEMM over the 16x8 pixel tile on the
// persistent kernel (TAPS = 1, fprop epilogue); single input of any width.  returns 1 / 0 / <0
int try_conv1x1_halo(cudaStream_t s, const dnnca_tensor_t* x, const void* wpack, int ktot, int ntot, UArgs a) {
  if (ktot % 64 || ntot % 8) return 0;
  const int bn = ntot % 256 == 0 ? 256 : (ntot % 128 == 0 ? 128 : (ntot <= 64 ? 64 : 128));
  CUtensorMap mA, mW;
  if (!halo_map(&mA, x, 8, 16)) return 0;
  if (!weight_map64(&mW, wpack, ktot, ntot, bn, 1)) return 0;
  a.tiles_x = (a.W + 7) / 8;
  a.tiles_y = (a.H + 15) / 16;
  a.nimg = x->n;
  a.taps = 1; a.sx = 1; a.c_a = ktot; a.c_b = 0;
  a.dbg = getenv("DNNCA_HALO_DBG") ? atoi(getenv("DNNCA_HALO_DBG")) : 0;
  const int kchunks = ktot / 64;
  const size_t limit = 226 * 1024;
  if (bn == 64) {
    if ((size_t)HGeom<64, 1>::smem_resident(kchunks) <= limit) return launch_halo<64, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<64, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (bn == 128) {
    if ((size_t)HGeom<128, 1>::smem_resident(kchunks) <= limit) return launch_halo<128, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<128, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if ((size_t)HGeom<256, 1>::smem_resident(kchunks) <= limit) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
  return launch_halo<256, false, 1>(s, mA, mA, mW, a, kchunks);
}

// ConvT 2x2/s2 fprop as ONE GEMM with N = 4*Cout (pack mode 2 already in `wpack`: [tap*Cout + co][Cin]) on the persistent
// kernel; Cin and Cout multiples of 64.  returns 1 / 0 / <0
int try_tconv_fprop_halo(cudaStream_t s, const dnnca_tensor_t* x, const void* wpack, int cin, int cout, UArgs a) {
  if (cin % 64 || cout % 64) return 0;
  CUtensorMap mA, mW;
  if (!halo_map(&mA, x, 8, 16)) return 0;
  if (!weight_map64(&mW, wpack, cin, 4 * cout, 256, 1)) return 0;
  a.tiles_x = (a.W + 7) / 8;
  a.tiles_y = (a.H + 15) / 16;
  a.nimg = x->n;
  a.taps = 1; a.sx = 1;
  a.dbg = getenv("DNNCA_HALO_DBG") ? atoi(getenv("DNNCA_HALO_DBG")) : 0;
  const int kchunks = cin / 64;
  if ((size_t)HGeom<256, 1>::smem_resident(kchunks) <= 226 * 1024) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
  return launch_halo<256, false, 1>(s, mA, mA, mW, a, kchunks);
}

// ConvT 2x2/s2 dgrad (pack mode 3 already in `wpack`: [tap][Cin][Cout]): GEMM N = layer Cin, K = 4 taps x Cout, the
// A operand of tap (ty, tx) = dy pixels (2y + ty, 2x + tx) through a stride-2 map.  returns 1 / 0 / <0
int try_tconv_dgrad_halo(cudaStream_t s, const dnnca_tensor_t* dy, const void* wpack, int cin, int cout, UArgs a) {
  if (cin % 64 || cout % 64) return 0;
  const int bn = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
  CUtensorMap mA, mW;
  if (!halo_map(&mA, dy, 8, 16, 2)) return 0;
  if (!weight_map64(&mW, wpack, cout, cin, bn, 4)) return 0;
  a.tiles_x = (a.W + 7) / 8;
  a.tiles_y = (a.H + 15) / 16;
  a.taps = 4; a.sx = 2; a.c_a = cout; a.c_b = 0;
  a.dbg = getenv("DNNCA_HALO_DBG") ? atoi(getenv("DNNCA_HALO_DBG")) : 0;
  const int kchunks = 4 * (cout / 64);
  const size_t limit = 226 * 1024;
  if (bn == 64) {
    if ((size_t)HGeom<64, 1>::smem_resident(kchunks) <= limit) return launch_halo<64, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<64, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (bn == 128) {
    if ((size_t)HGeom<128, 1>::smem_resident(kchunks) <= limit) return launch_halo<128, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<128, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if ((size_t)HGeom<256, 1>::smem_resident(kchunks) <= limit) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
  return launch_halo<256, false, 1>(s, mA, mA, mW, a, kchunks);
}

}  // namespace dnnca
