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
// Geometry: M tile = 16 rows x 8 pixels.  The halo box {64 ch, 16 px, 18 rows} lands as pixel rows of 128 bytes
// (SWIZZLE_128B), 16 pixels per image row, so the A operand of tap (dy,dx) is the SAME buffer read through a
// descriptor whose start address is shifted by (dy*16 + dx) rows: the 8 pixels of a tile row form one 8-row core
// group, consecutive tile rows are SBO = 16*128 = 2048 bytes apart.  Only pixels x0-1..x0+8 of the 16 are used.
#include <stdlib.h>

#include "umma_common.cuh"

namespace dnnca {

constexpr int A_SLOT = 18 * 16 * 128;      // halo tile of one 64-channel chunk (36 KB)

template <int BN>
struct HGeom {
  static constexpr int B_TAP = BN * 128;                   // one tap of one 64-channel chunk
  static constexpr int A_SLOTS = 2;
  static constexpr int CTRL = 6144;                        // barriers + TMEM slot + bias slice (BN floats at +256) + BN-statistics accumulators (2*BN doubles at +2048)
  // resident: all 9 * (Cin/64) weight blocks stay in shared memory for the CTA's lifetime
  static constexpr int smem_resident(int kchunks) { return CTRL + A_SLOTS * A_SLOT + 9 * kchunks * B_TAP + 1024; }
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

template <int BN, bool RESIDENT>
__global__ void __launch_bounds__(320) conv_umma_halo_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const __grid_constant__ CUtensorMap mapW, UArgs a) {
  using G = HGeom<BN>;
  constexpr int B_STAGES = RESIDENT ? 1 : G::B_STAGES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* fullA = reinterpret_cast<uint64_t*>(smem);      // [2]
  uint64_t* emptyA = fullA + 2;                              // [2]
  uint64_t* fullB = emptyA + 2;                              // [B_STAGES] (resident: [1], completes once)
  uint64_t* emptyB = fullB + 8;                              // [B_STAGES]
  uint64_t* tfull = emptyB + 8;                              // [2] accumulator ready
  uint64_t* tempty = tfull + 2;                              // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* sbias = reinterpret_cast<float*>(smem + 256);      // [BN] bias slice of this CTA's N tile (fprop)
  double* sstat = reinterpret_cast<double*>(smem + 2048);   // [2*BN] per-channel sum | sum of squares (fprop + stats)
  const bool do_stats = a.stats != nullptr && a.epi == EPI_FPROP;
  unsigned char* aring = smem + G::CTRL;
  unsigned char* bring = aring + G::A_SLOTS * A_SLOT;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kca = a.c_a / 64, kcb = a.c_b / 64, kchunks = kca + kcb;
  const int n0 = blockIdx.y * BN;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(fullB + s, 1); mbar_init(emptyB + s, 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  for (int i = threadIdx.x; i < BN; i += blockDim.x)
    sbias[i] = (a.bias && a.epi != EPI_DGRAD && n0 + i < a.n_total) ? a.bias[n0 + i] : 0.f;
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
        mbar_expect_tx(fullB, (uint32_t)(9 * kchunks * G::B_TAP));
        for (int kc = 0; kc < kchunks; ++kc)
          for (int tap = 0; tap < 9; ++tap)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                ::"r"(smem_u32(bring + (kc * 9 + tap) * G::B_TAP)), "l"(reinterpret_cast<uint64_t>(&mapW)), "r"(smem_u32(fullB)),
                "r"(kc * 64), "r"(n0), "r"(tap)
                : "memory");
      }
      int ai = 0, bi = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int b = t;
        const int tix = b % a.tiles_x; b /= a.tiles_x;
        const int tiy = b % a.tiles_y;
        const int n = b / a.tiles_y;
        const int x0 = tix * 8, y0 = tiy * 16;
        for (int kc = 0; kc < kchunks; ++kc, ++ai) {
          const int s = ai & 1;
          if (ai >= 2) mbar_wait(emptyA + s, ((ai >> 1) - 1) & 1);
          mbar_expect_tx(fullA + s, A_SLOT);
          const bool second = kc >= kca;
          tma_load_4d(aring + s * A_SLOT, second ? &mapB : &mapA, fullA + s, (second ? kc - kca : kc) * 64, x0 - 1, y0 - 1, n);
          if (!RESIDENT) {
            for (int tap = 0; tap < 9; ++tap, ++bi) {
              const int sb = bi % B_STAGES;
              if (bi >= B_STAGES) mbar_wait(emptyB + sb, ((bi / B_STAGES) - 1) & 1);
              mbar_expect_tx(fullB + sb, G::B_TAP);
              asm volatile(
                  "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                  ::"r"(smem_u32(bring + sb * G::B_TAP)), "l"(reinterpret_cast<uint64_t>(&mapW)), "r"(smem_u32(fullB + sb)),
                  "r"(kc * 64), "r"(n0), "r"(tap)
                  : "memory");
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
      if (RESIDENT) { mbar_wait(fullB, 0); }
      int ai = 0, bi = 0, ti = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
        const int buf = ti & 1;
        if (ti >= 2) mbar_wait(tempty + buf, ((ti >> 1) - 1) & 1);    // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t dtm = tmem_base + (uint32_t)(buf * BN);
        for (int kc = 0; kc < kchunks; ++kc, ++ai) {
          const int s = ai & 1;
          mbar_wait(fullA + s, (ai >> 1) & 1);
          tc_fence_after();
          // descriptors differ only in the 14-bit start-address field: build one per operand and add immediates
          // (the MMA-issuing thread is a serial instruction stream; at N = 64 an MMA retires every ~32 cycles)
          const uint64_t a_base = kmajor128_desc(smem_u32(aring + s * A_SLOT), 2048, 0);
          if (RESIDENT) {
            const uint64_t b_base = kmajor128_desc(smem_u32(bring + kc * 9 * G::B_TAP), 1024, 0);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint64_t da0 = a_base + (uint64_t)(((tap / 3) * 16 + (tap % 3)) * 8);      // (dy*16+dx) rows of 128 B, >>4
              const uint64_t db0 = b_base + (uint64_t)(tap * (G::B_TAP >> 4));
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(dtm, da0 + 2 * k, db0 + 2 * k, idesc, (kc | tap | k) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, ++bi) {
              const int st = bi % B_STAGES;
              mbar_wait(fullB + st, (bi / B_STAGES) & 1);
              tc_fence_after();
              const uint64_t db0 = kmajor128_desc(smem_u32(bring + st * G::B_TAP), 1024, 0);
              const uint64_t da0 = a_base + (uint64_t)(((tap / 3) * 16 + (tap % 3)) * 8);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16(dtm, da0 + 2 * k, db0 + 2 * k, idesc, (kc | tap | k) ? 1u : 0u);
              umma_commit(emptyB + st);
            }
          }
          umma_commit(emptyA + s);          // halo slot free once these MMAs retire
        }
        umma_commit(tfull + buf);           // accumulator of this tile complete
      }
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
    constexpr bool REG_STATS = BN <= 64;
    float acc1[REG_STATS ? NCHUNK : 1][32], acc2[REG_STATS ? NCHUNK : 1][32];
#pragma unroll
    for (int c = 0; c < (REG_STATS ? NCHUNK : 1); ++c)
#pragma unroll
      for (int j = 0; j < 32; ++j) { acc1[c][j] = 0.f; acc2[c][j] = 0.f; }
    int ti = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++ti) {
      int b = t;
      const int tix = b % a.tiles_x; b /= a.tiles_x;
      const int tiy = b % a.tiles_y;
      const int n = b / a.tiles_y;
      const int gy = tiy * 16 + ty, gx = tix * 8 + tx;
      const bool inside = gy < a.H && gx < a.W;
      const int buf = ti & 1;
      ChunkAddr ca[NCHUNK];
      uint4 m[NCHUNK][4];
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col = half * (BN / 2) + c * 32;
        ca[c] = chunk_addr(a, n, inside ? gy : 0, inside ? gx : 0, n0 + col);
        if (ca[c].msk && inside) {
#pragma unroll
          for (int q = 0; q < 4; ++q) m[c][q] = reinterpret_cast<const uint4*>(ca[c].msk)[q];
        }
      }
      mbar_wait(tfull + buf, (ti >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        const int col = half * (BN / 2) + c * 32;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * BN + col), v);
        tmem_ld_wait();
        if (!do_stats) {
          if (inside && n0 + col < a.n_total) epilogue_chunk32(a, v, ca[c], m[c], sbias + col);
        } else {
          // BatchNormalization statistics of the stored tensor (components.py:57-58,130-132) in the epilogue: rows of
          // the tile are lanes, so a column sum is a 31-shuffle warp reduction; one shared fp64 atomic per lane
          float r1[32], r2[32];
          epilogue_chunk32<true>(a, v, ca[c], m[c], sbias + col, r1, inside && n0 + col < a.n_total);
          if (REG_STATS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float r = inside ? r1[j] : 0.f;
              acc1[REG_STATS ? c : 0][j] += r;
              acc2[REG_STATS ? c : 0][j] = fmaf(r, r, acc2[REG_STATS ? c : 0][j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              r1[j] = inside ? r1[j] : 0.f;
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
static bool halo_map(CUtensorMap* m, const dnnca_tensor_t* t) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  char* base = reinterpret_cast<char*>(t->data) + (size_t)t->coff * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (t->cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
  cuuint32_t box[4] = {64, 16, 18, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool weight_map64(CUtensorMap* m, const void* wp, int ktot, int ntot, int bn) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)ntot, 9};
  cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * ntot * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wp), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, bool RESIDENT>
static int launch_halo(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                       int kchunks) {
  using G = HGeom<BN>;
  const int smem = RESIDENT ? G::smem_resident(kchunks) : G::SMEM_STREAM;
  auto kern = conv_umma_halo_kernel<BN, RESIDENT>;
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

// Conv2D 3x3 fprop (pack mode 0 already in `wpack`: [tap][N][K]) or dgrad (pack mode 1); returns 1 / 0 / <0
int try_conv3x3_halo(cudaStream_t s, const dnnca_tensor_t* xa, const dnnca_tensor_t* xb, const void* wpack, int ktot, int ntot,
                     UArgs a) {
  // a single input with fewer than 64 channels (first layers) is one K chunk whose missing channels are zero-filled by
  // the TMA (activations and packed weights alike)
  const bool narrow = !xb && xa->c < 64;
  if ((!narrow && (xa->c % 64 || (xb && xb->c % 64))) || ntot % 64) return 0;
  const int bn = ntot % 256 == 0 ? 256 : (ntot % 128 == 0 ? 128 : 64);
  if (a.split % bn) return 0;               // an N tile must not straddle the two dgrad destinations
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
  const size_t limit = 220 * 1024;
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

}  // namespace dnnca
