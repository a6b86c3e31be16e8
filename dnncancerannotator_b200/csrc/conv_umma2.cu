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
  // staged epilogue: every epilogue warp owns STG_BUFS swizzled [32 px][64 ch] bf16 blocks (4 KB each) that leave through
  // its own TMA stores
  static constexpr int STG_BUFS = BN > 128 ? 1 : 2;
  static constexpr int STG_BYTES = 8 * STG_BUFS * 4096;
  static constexpr int smem_resident(int kchunks, bool staged = false) {
    return CTRL + A_SLOTS * A_SLOT + TAPS * kchunks * B_TAP + (staged ? STG_BYTES : 0) + 1024;
  }
  static constexpr int B_STAGES = BN > 128 ? 3 : 4;
  static constexpr int smem_stream(bool staged) { return CTRL + A_SLOTS * A_SLOT + B_STAGES * B_TAP + (staged ? STG_BYTES : 0) + 1024; }
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

__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}

// Staged epilogue of one 32-column chunk of one accumulator row: bias + activation (fprop) or act'(mask) (dgrad), bf16
// rounding, then four 16-byte stores into the warp's swizzled staging block: pixel row `lane` (128 bytes), 16-byte
// chunk j lands at position j ^ (lane & 7) -- the SWIZZLE_128B pattern the TMA store expects; conflict-free.
template <int EPI>
__device__ __forceinline__ void stage_chunk32(const UArgs& a, const uint32_t (&v)[32], const float* sbias,
                                              const __nv_bfloat16* msk, uint32_t srow, int chunk0, uint32_t sw) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  const float slope = act_slope(a.act, a.alpha);
  if (EPI == EPI_DGRAD) {
    if (msk) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 m = reinterpret_cast<const uint4*>(msk)[q];
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[q * 8 + 2 * j] *= __uint_as_float(mw[j] << 16) > 0.f ? 1.f : slope;
          f[q * 8 + 2 * j + 1] *= __uint_as_float(mw[j] & 0xffff0000u) > 0.f ? 1.f : slope;
        }
      }
    }
  } else if (EPI == EPI_FPROP || EPI == EPI_TCONV) {
    const float4* sb4 = reinterpret_cast<const float4*>(sbias);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = sb4[j];
      f[4 * j + 0] = act_slope_apply(f[4 * j + 0] + b4.x, slope);
      f[4 * j + 1] = act_slope_apply(f[4 * j + 1] + b4.y, slope);
      f[4 * j + 2] = act_slope_apply(f[4 * j + 2] + b4.z, slope);
      f[4 * j + 3] = act_slope_apply(f[4 * j + 3] + b4.w, slope);
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[q * 8 + 0], f[q * 8 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(f[q * 8 + 2], f[q * 8 + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[q * 8 + 4], f[q * 8 + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(f[q * 8 + 6], f[q * 8 + 7]);
    sts_v4(srow + ((((uint32_t)(chunk0 + q)) ^ sw) << 4), *reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
           *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  }
}

// STAGED: epilogue through shared memory + TMA stores (fprop / dgrad kinds); the two epilogue warp groups take alternate
// tiles, each warp owns its 32 accumulator rows over ALL columns of the tile.  Measured motive (B200, B = 32): with the
// per-thread 64-byte global stores removed 256->256@64 fprop ran 115 -> 96 us, 128->128@128 136 -> 124 us; BatchNorm
// statistics as per-tile warp shuffles cost another 37 us at 128->128@128.
template <int BN, bool RESIDENT, int TAPS, int EPI, bool STAGED>
__global__ void __launch_bounds__(320) conv_umma_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const __grid_constant__ CUtensorMap mapW,
                                                            const __grid_constant__ CUtensorMap mapYa,
                                                            const __grid_constant__ CUtensorMap mapYb, UArgs a) {
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
  constexpr bool BNR = EPI == EPI_DGRAD_BNR, IS_DGRAD = EPI == EPI_DGRAD || BNR;
  constexpr int EPI_ADDR = IS_DGRAD ? EPI_DGRAD : EPI;     // addressing / store code of the epilogue helpers
  float* sbias = reinterpret_cast<float*>(smem + G::BIAS_OFF);
  double* sstat = reinterpret_cast<double*>(smem + G::STAT_OFF);   // [2*BN] per-channel sum | sum of squares (fprop + stats)
  const bool do_stats = (EPI == EPI_FPROP || BNR || (EPI == EPI_TCONV && STAGED)) && a.stats != nullptr;
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
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, STAGED ? 4 : 8); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(fullB + s, 1); mbar_init(emptyB + s, 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  for (int i = threadIdx.x; i < NCLS * BN; i += blockDim.x) {
    const int cls = i / BN, col = i - cls * BN;
    float b = 0.f;
    if (BNR) {                              // the BatchNorm's batch mean of this column (destination A only)
      if (n0 + col < a.split) b = a.bnr_mi[n0 + col];
    } else if (!IS_DGRAD && n0 + col < a.n_total) {
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
  } else if (STAGED) {
    // ===== staged epilogue: warps 2..5 take the even tiles of this CTA, warps 6..9 the odd ones (TMEM buffer = tile parity);
    // a warp reads its 32 accumulator rows (TMEM lanes 32*(warp%4)..+31) 64 columns at a time, writes them as a swizzled
    // [32 px][64 ch] bf16 block and hands the block to the TMA (box {64 ch, 8 px, 4 rows}); BatchNorm sums are taken from
    // the staged block with lanes mapped to channel pairs =====
    const int grp = (warp - 2) >> 2, lg = warp & 3;
    const int r = lg * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    constexpr int NBLK = BN / 64;
    constexpr int STG_BUFS = G::STG_BUFS;
    unsigned char* stg = bring + (RESIDENT ? TAPS * kchunks * G::B_TAP : B_STAGES * G::B_TAP) + (warp - 2) * (STG_BUFS * 4096);
    const uint32_t stg_u = smem_u32(stg);
    const uint32_t sw = (uint32_t)(lane & 7);
    // column pass: lane l owns channels 2l, 2l+1 of a 64-column block = one 4-byte word per pixel row
    const uint32_t cword = (uint32_t)((lane & 3) << 2);
    const uint32_t cchunk = (uint32_t)(lane >> 2);
    float s0[NBLK][2], s1[NBLK][2], mean2[NBLK][2];
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      s0[b][0] = s0[b][1] = s1[b][0] = s1[b][1] = 0.f;
      mean2[b][0] = mean2[b][1] = 0.f;
      if (BNR && do_stats) {
        const int ch = n0 + b * 64 + 2 * lane;
        if (ch < a.split) { mean2[b][0] = a.bnr_mi[ch]; mean2[b][1] = a.bnr_mi[ch + 1]; }
      }
    }
    int sbuf = 0;
    int ti = grp;
    TileWalk tw_(blockIdx.x + grp * gridDim.x, 2 * gridDim.x, a.tiles_x, a.tiles_y);
    for (int t = blockIdx.x + grp * gridDim.x; t < ntiles; t += 2 * gridDim.x, ti += 2, tw_.next()) {
      const int n = tw_.n;
      const int gy = tw_.tiy * 16 + ty, gx = tw_.tix * 8 + tx;
      const bool inside = gy < a.H && gx < a.W;
      const int by0 = tw_.tiy * 16 + lg * 4, bx0 = tw_.tix * 8;       // this warp's 4 rows x 8 pixels of the tile
      const int buf = ti & 1;
      const float* sb = sbias;
      if (NCLS == 9) sb += ((gy == 0 ? 0 : (gy == a.H - 1 ? 2 : 1)) * 3 + (gx == 0 ? 0 : (gx == a.W - 1 ? 2 : 1))) * BN;
      const long long pix = ((long long)n * a.H + (inside ? gy : 0)) * a.W + (inside ? gx : 0);
      mbar_wait(tfull + buf, (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int blk = 0; blk < NBLK; ++blk) {
        const int colb = blk * 64, ncol = n0 + colb;
        const bool blk_live = ncol < a.n_total;
        const bool in_a = ncol < a.split;
        uint32_t v0[32], v1[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN + colb), v0);
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN + colb + 32), v1);
        // BatchNorm backward sums: the BN input at this block's pixels, lanes over channel pairs (coalesced 128-byte rows)
        uint32_t xw[(BNR ? 32 : 1)];
        if (BNR && do_stats && blk_live && in_a) {
          const __nv_bfloat16* xb = a.mask + ncol + 2 * lane;
          const bool lane_ok = ncol + 2 * lane < a.split;          // a partly filled block: channels beyond the tensor
#pragma unroll
          for (int p = 0; p < 32; ++p) {
            const int py = by0 + (p >> 3), px = bx0 + (p & 7);
            xw[BNR ? p : 0] = (lane_ok && py < a.H && px < a.W)
                                  ? __ldg(reinterpret_cast<const unsigned int*>(xb + (((long long)n * a.H + py) * a.W + px) * a.mask_cs))
                                  : 0u;
          }
        }
        tmem_ld_wait();
        if (blk == NBLK - 1) {                 // the accumulator now lives in registers: the MMA warp may reuse it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty + buf);
        }
        if (lane == 0) {                       // staging block free again? (its previous TMA store has read it)
          if (STG_BUFS == 2) tma_store_wait_read1();
          else tma_store_wait_read();
        }
        __syncwarp();
        const uint32_t sblock = stg_u + (uint32_t)(sbuf * 4096);
        const uint32_t srow = sblock + (uint32_t)(lane * 128);
        const __nv_bfloat16* mk = nullptr;
        if (EPI == EPI_DGRAD && a.mask && in_a && inside && blk_live) mk = a.mask + pix * a.mask_cs + ncol;
        stage_chunk32<EPI_ADDR>(a, v0, sb + colb, mk, srow, 0, sw);
        stage_chunk32<EPI_ADDR>(a, v1, sb + colb + 32, (mk && ncol + 32 < a.n_total) ? mk + 32 : nullptr, srow, 4, sw);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && blk_live) {
          if (EPI == EPI_TCONV) {
            // ConvT 2x2/s2: column block = (filter tap, 64 output channels); tap (ty, tx) of input pixel (y, x) is output
            // pixel (2y + ty, 2x + tx): the output is described to the TMA as {C, 2, W, 2, H*N} (mapYa), one box per tap
            const int tap = ncol / a.cout_t, ch = ncol - tap * a.cout_t;
            if (by0 < a.H) tma_store_5d(&mapYa, stg + sbuf * 4096, ch, tap & 1, bx0, tap >> 1, n * a.H + by0);
          } else if (in_a) {
            tma_store_4d(&mapYa, stg + sbuf * 4096, ncol, bx0, by0, n);
          } else {
            tma_store_4d(&mapYb, stg + sbuf * 4096, ncol - a.split, bx0, by0, n);
          }
          tma_store_commit();
        }
        if (do_stats && blk_live && (!BNR || in_a)) {
          // sums of the values as they read back from the bf16 tensor, from the staged block
#pragma unroll
          for (int p = 0; p < 32; ++p) {
            const bool pv = (by0 + (p >> 3) < a.H) && (bx0 + (p & 7) < a.W);
            const uint32_t w = lds_u32(sblock + (uint32_t)(p * 128) + ((cchunk ^ (uint32_t)(p & 7)) << 4) + cword);
            const float lo = pv ? __uint_as_float(w << 16) : 0.f, hi = pv ? __uint_as_float(w & 0xffff0000u) : 0.f;
            if (BNR) {
              const uint32_t xv = xw[BNR ? p : 0];
              s0[blk][0] += lo;
              s0[blk][1] += hi;
              s1[blk][0] = fmaf(lo, __uint_as_float(xv << 16) - mean2[blk][0], s1[blk][0]);
              s1[blk][1] = fmaf(hi, __uint_as_float(xv & 0xffff0000u) - mean2[blk][1], s1[blk][1]);
            } else {
              s0[blk][0] += lo;
              s0[blk][1] += hi;
              s1[blk][0] = fmaf(lo, lo, s1[blk][0]);
              s1[blk][1] = fmaf(hi, hi, s1[blk][1]);
            }
          }
        }
        if (STG_BUFS == 2) sbuf ^= 1;
      }
    }
    if (do_stats) {
#pragma unroll
      for (int b = 0; b < NBLK; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          atomicAdd(sstat + b * 64 + 2 * lane + e, (double)s0[b][e]);
          atomicAdd(sstat + BN + b * 64 + 2 * lane + e, (double)s1[b][e]);
        }
    }
    if (lane == 0) tma_store_wait_all();       // the staging blocks must outlive their stores
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
    constexpr bool REG_STATS = BN <= 64 && (EPI == EPI_FPROP || BNR);
    constexpr int NMASK = IS_DGRAD ? NCHUNK : 1;             // mask rows are prefetched only by dgrad
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
        ca[c] = chunk_addr<EPI_ADDR>(a, n, inside ? gy : 0, inside ? gx : 0, n0 + col);
        if (IS_DGRAD && ca[c].msk && inside && n0 + col < a.n_total) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m[IS_DGRAD ? c : 0][q] = reinterpret_cast<const uint4*>(ca[c].msk)[q];
        } else if (BNR) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m[IS_DGRAD ? c : 0][q] = make_uint4(0u, 0u, 0u, 0u);
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
        if (BNR && do_stats) {
          // BatchNormalization backward sums of the gradient as it reads back from the bf16 tensor (components.py:57-59,
          // 130-131 under GradientTape): r1 = dy, r2 = dy * (x - mean); invstd is applied once per CTA at the end
          const bool live = inside && ca[c].msk != nullptr && n0 + col < a.n_total;
          epilogue_chunk32<true, EPI_DGRAD>(a, v, ca[c], m[IS_DGRAD ? c : 0], sb + col, inside && n0 + col < a.n_total,
                                            a.n_total - (n0 + col));
          float r1[32], r2[32];
          const float4* mean4 = reinterpret_cast<const float4*>(sbias + col);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 mq = m[IS_DGRAD ? c : 0][q];
            const uint32_t mw[4] = {mq.x, mq.y, mq.z, mq.w};
            const float4 ma = mean4[2 * q], mb = mean4[2 * q + 1];
            const float mean8[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float g0 = live ? __uint_as_float(v[q * 8 + 2 * j]) : 0.f;
              const float g1 = live ? __uint_as_float(v[q * 8 + 2 * j + 1]) : 0.f;
              r1[q * 8 + 2 * j] = g0;
              r1[q * 8 + 2 * j + 1] = g1;
              r2[q * 8 + 2 * j] = g0 * (__uint_as_float(mw[j] << 16) - mean8[2 * j]);
              r2[q * 8 + 2 * j + 1] = g1 * (__uint_as_float(mw[j] & 0xffff0000u) - mean8[2 * j + 1]);
            }
          }
          if (REG_STATS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { acc1[REG_STATS ? c : 0][j] += r1[j]; acc2[REG_STATS ? c : 0][j] += r2[j]; }
          } else {
            warp_colsum32(r1, lane);
            warp_colsum32(r2, lane);
            if (n0 + col + lane < a.split) {
              atomicAdd(sstat + col + lane, (double)r1[0]);
              atomicAdd(sstat + BN + col + lane, (double)r2[0]);
            }
          }
        } else if (EPI != EPI_FPROP || !do_stats) {
          if (inside && n0 + col < a.n_total)
            epilogue_chunk32<false, EPI_ADDR>(a, v, ca[c], m[IS_DGRAD ? c : 0], sb + col, !(a.dbg & 1),
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
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      if (BNR) {
        if (n0 + i < a.split) {               // sum dy | sum dy*xhat = invstd * sum dy*(x - mean)
          atomicAdd(a.stats + n0 + i, sstat[i]);
          atomicAdd(a.stats + a.split + n0 + i, sstat[BN + i] * (double)a.bnr_mi[a.split + n0 + i]);
        }
      } else if (EPI == EPI_TCONV) {
        if (n0 + i < a.n_total) {             // the four filter taps of a channel are four column blocks
          atomicAdd(a.stats + (n0 + i) % a.cout_t, sstat[i]);
          atomicAdd(a.stats + a.cout_t + (n0 + i) % a.cout_t, sstat[BN + i]);
        }
      } else if (n0 + i < a.n_total) {
        atomicAdd(a.stats + n0 + i, sstat[i]);
        atomicAdd(a.stats + a.n_total + n0 + i, sstat[BN + i]);
      }
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

// ConvT 2x2/s2 output [N, 2H, 2W, C] seen from the INPUT pixel grid: {C, 2 (tx), W, 2 (ty), H*N}; box {64 ch, 1, 8 px, 1, 4 rows}.
// (H and N share one dimension: a 4-row box never straddles two images when H % 4 == 0.)
static bool tconv_out_map(CUtensorMap* m, __nv_bfloat16* base, int c, long long cstride, int w, int h, int n) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc || c <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (cstride * 2) % 16) return false;
  cuuint64_t dims[5] = {(cuuint64_t)c, 2, (cuuint64_t)w, 2, (cuuint64_t)h * n};
  cuuint64_t strides[4] = {(cuuint64_t)cstride * 2, (cuuint64_t)2 * cstride * 2, (cuuint64_t)2 * w * cstride * 2,
                           (cuuint64_t)4 * w * cstride * 2};
  cuuint32_t box[5] = {64, 1, 8, 1, 4};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool halo_staged_enabled() {
  static int v = -1;
  if (v < 0) v = (getenv("DNNCA_HALO_STAGED") && atoi(getenv("DNNCA_HALO_STAGED")) == 0) ? 0 : 1;      // A/B switch
  return v == 1;
}

// do the resident weights (+ the staging blocks of a staged epilogue) fit into shared memory?
template <int BN, int TAPS>
static bool resident_fits(int kchunks, const UArgs& a) {
  // (resident weights win over the staged epilogue where only one of them fits: [64+64]->64@256 fprop measured 290 us
  // resident + per-thread stores against 444 us streamed + staged)
  (void)a;
  return (size_t)HGeom<BN, TAPS>::smem_resident(kchunks, false) <= 226 * 1024;
}

// output views of a staged epilogue: 4-D {C, W, H, N}, box {64 ch, 8 px, 4 rows, 1}, SWIZZLE_128B
static bool out_map(CUtensorMap* m, __nv_bfloat16* base, int c, long long cstride, int w, int h, int n) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc || c <= 0) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)cstride * 2, (cuuint64_t)w * cstride * 2, (cuuint64_t)h * w * cstride * 2};
  cuuint32_t box[4] = {64, 8, 4, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool RESIDENT, int TAPS, int EPI, bool STAGED>
static int launch_halo_st(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const CUtensorMap& mYa,
                          const CUtensorMap& mYb, const UArgs& a, int kchunks) {
  using G = HGeom<BN, TAPS>;
  const int smem = RESIDENT ? G::smem_resident(kchunks, STAGED) : G::smem_stream(STAGED);
  auto kern = conv_umma_halo_kernel<BN, RESIDENT, TAPS, EPI, STAGED>;
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
  kern<<<grid, 320, smem, s>>>(mA, mB, mW, mYa, mYb, a);
  DNNCA_LAUNCH_CHECK("conv_umma_halo");
  note_family(2);
  return 1;
}

template <int BN, bool RESIDENT, int TAPS, int EPI>
static int launch_halo_epi(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                           int kchunks) {
  using G = HGeom<BN, TAPS>;
  if (EPI == EPI_TCONV && halo_staged_enabled() && a.cout_t % 64 == 0 && a.H % 4 == 0 &&
      (size_t)(RESIDENT ? G::smem_resident(kchunks, true) : G::smem_stream(true)) <= 226 * 1024) {
    CUtensorMap mY;
    if (tconv_out_map(&mY, a.ya, a.cout_t, a.ya_cs, a.W, a.H, a.nimg)) {
      const int r = launch_halo_st<BN, RESIDENT, TAPS, EPI, true>(s, mA, mB, mW, mY, mY, a, kchunks);
      return (r == 1 && a.stats) ? 2 : r;      // 2: the BatchNorm statistics of the output were taken as well
    }
  }
  // staged epilogue (fprop / dgrad kinds): needs whole 64-column blocks per destination and room for the staging blocks
  if (EPI != EPI_TCONV && halo_staged_enabled() && (a.split % 64 == 0 || a.split >= a.n_total) &&
      (size_t)(RESIDENT ? G::smem_resident(kchunks, true) : G::smem_stream(true)) <= 226 * 1024) {
    CUtensorMap mYa, mYb;
    const int ca = a.split < a.n_total ? a.split : a.n_total, cb = a.n_total - ca;
    bool ok = out_map(&mYa, a.ya, ca, a.ya_cs, a.W, a.H, a.nimg);
    mYb = mYa;
    if (ok && cb > 0) ok = out_map(&mYb, a.yb, cb, a.yb_cs, a.W, a.H, a.nimg);
    if (ok) return launch_halo_st<BN, RESIDENT, TAPS, EPI == EPI_TCONV ? EPI_FPROP : EPI, true>(s, mA, mB, mW, mYa, mYb, a, kchunks);
  }
  return launch_halo_st<BN, RESIDENT, TAPS, EPI, false>(s, mA, mB, mW, mA, mA, a, kchunks);
}

// the epilogue kind is a template parameter (the fprop instantiation carries BatchNorm-statistics registers, the dgrad one
// the prefetched mask rows; together they do not fit the 168 registers 320 threads leave)
template <int BN, bool RESIDENT, int TAPS = 9>
static int launch_halo(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                       int kchunks) {
  if (a.epi == EPI_DGRAD) return launch_halo_epi<BN, RESIDENT, TAPS, EPI_DGRAD>(s, mA, mB, mW, a, kchunks);
  if (a.epi == EPI_DGRAD_BNR) return launch_halo_epi<BN, RESIDENT, TAPS, EPI_DGRAD_BNR>(s, mA, mB, mW, a, kchunks);
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
  if ((a.epi == EPI_DGRAD || a.epi == EPI_DGRAD_BNR) && (ntot % 32 || a.split % 32)) return 0;
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
    if (resident_fits<64, 9>(kchunks, a)) return launch_halo<64, true>(s, mA, mB, mW, a, kchunks);
    return launch_halo<64, false>(s, mA, mB, mW, a, kchunks);
  }
  if (bn == 128) {
    if (resident_fits<128, 9>(kchunks, a)) return launch_halo<128, true>(s, mA, mB, mW, a, kchunks);
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
    if (resident_fits<64, 1>(kchunks, a)) return launch_halo<64, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<64, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (bn == 128) {
    if (resident_fits<128, 1>(kchunks, a)) return launch_halo<128, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<128, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (resident_fits<256, 1>(kchunks, a)) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
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
  if (resident_fits<256, 1>(kchunks, a)) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
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
    if (resident_fits<64, 1>(kchunks, a)) return launch_halo<64, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<64, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (bn == 128) {
    if (resident_fits<128, 1>(kchunks, a)) return launch_halo<128, true, 1>(s, mA, mA, mW, a, kchunks);
    return launch_halo<128, false, 1>(s, mA, mA, mW, a, kchunks);
  }
  if (resident_fits<256, 1>(kchunks, a)) return launch_halo<256, true, 1>(s, mA, mA, mW, a, kchunks);
  return launch_halo<256, false, 1>(s, mA, mA, mW, a, kchunks);
}

}  // namespace dnnca
