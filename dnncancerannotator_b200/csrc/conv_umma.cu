// tcgen05 implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// One kernel serves Conv2D fprop, Conv2D dgrad, ConvTranspose2x2 fprop and ConvTranspose2x2 dgrad:
//     D[128 pixels, BN channels] = sum over (tap, channel chunk)  A_tap[128 px, KC] * Wp_tap[BN, KC]^T
//   * A tile  : an 8x16 pixel rectangle of the NHWC activation (4-D TMA box {KC, 16, 8, 1}); a filter tap
//               is the same box at shifted coordinates and the TMA's out-of-bounds zero fill IS the
//               'same' padding.  ConvT dgrad uses traversal stride 2 (one 2x2 tap per box).
//               A second input tensor continues the K loop (virtual tf.concat, components.py:164).
//   * B tile  : bf16 weights pre-packed K-major [tap][N][K] by pack_weights_kernel (3-D TMA box).
//   * both land in 128/64/32-byte swizzled K-major shared memory = the canonical UMMA operand layout,
//     so the shared-memory matrix descriptors need only (start address, SBO, swizzle mode).
//   * warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected lane) + TMEM allocator,
//     warps 2..5 = epilogue: tcgen05.ld 32 lanes x 32 columns -> bias/activation (fprop) or act'(mask)
//     (dgrad) -> bf16 -> 16-byte global stores straight into the (possibly channel-sliced) NHWC output;
//     ConvT fprop scatters each 2x2 tap to its pixel-shuffled position.
//   * full/empty mbarrier ring between TMA and MMA (tcgen05.commit frees a stage), one mbarrier hands
//     the finished accumulator to the epilogue.
// wgrad (both operands pixel-major = MN-major) is wgrad_umma_kernel below.
// Reference call sites: layers.Conv2D components.py:47-50,123-126; Convolution2DTranspose :118-120.
#include <stdlib.h>

#include "umma_common.cuh"

namespace dnnca {

// ---------------------------------------------------------------- weight packing
// Conv2D  w[k,k,Cin,Cout] (HWIO) -> fprop pack  [tap][Cout][Cin]           (MODE 0)
//                                 -> dgrad pack  [tap'][Cin][Cout], tap' = rot180(tap)   (MODE 1)
// ConvT   k[2,2,Cout,Cin]        -> fprop pack  [1][tap*Cout + co][Cin]    (MODE 2: plain bf16 cast)
//                                 -> dgrad pack  [tap][Cin][Cout]           (MODE 3)
// K (the contiguous axis) is padded to a multiple of 16 with zeros in modes 0/2 (`kp` >= cin): layers whose input
// has fewer than 16 channels (first convs: 1, 3 or 5 modalities) still feed whole K=16 MMAs.
// `sa` / `sb` (mode 0 only, may be NULL): per-input-channel scale of a BatchNorm folded into the conv's input
// (channels [0, ca) of the first tensor, [ca, cin) of the second): the packed weight is bf16(w * scale[ci])
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                          int mode, int taps, int cin, int cout, int kp,
                                                          const float* __restrict__ sa, int ca,
                                                          const float* __restrict__ sb) {
  const long long total = (mode == 0 || mode == 2) ? (long long)taps * cout * kp : (long long)taps * cin * cout;
  const int kk = taps == 9 ? 3 : (taps == 4 ? 2 : 1);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    float v;
    if (mode == 0) {            // e = (tap, co, ci padded)
      const int ci = (int)(e % kp);
      const long long t = e / kp;
      const int co = (int)(t % cout), tap = (int)(t / cout);
      v = ci < cin ? w[((long long)tap * cin + ci) * cout + co] : 0.f;
      if (ci < ca) { if (sa) v *= sa[ci]; }
      else if (ci < cin && sb) v *= sb[ci - ca];
    } else if (mode == 1) {     // e = (tap', ci, co), source tap = rot180
      const int co = (int)(e % cout);
      const long long t = e / cout;
      const int ci = (int)(t % cin), tp = (int)(t / cin);
      const int a = kk - 1 - tp / kk, c = kk - 1 - tp % kk;
      v = w[((long long)(a * kk + c) * cin + ci) * cout + co];
    } else if (mode == 2) {     // e = (tap*cout + co, ci padded) from k[tap][co][ci]
      const int ci = (int)(e % kp);
      const long long t = e / kp;
      v = ci < cin ? w[t * cin + ci] : 0.f;
    } else {                    // e = (tap, ci, co) from k[tap][co][ci]
      const int co = (int)(e % cout);
      const long long t = e / cout;
      const int ci = (int)(t % cin), tap = (int)(t / cin);
      v = w[((long long)tap * cout + co) * cin + ci];
    }
    out[e] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------- the kernel
// short K loops (<= 8 iterations: ConvT fprop, 1x1 convs) run with a 2-deep ring so that 3-4 CTAs fit an SM and overlap
// each other's load / MMA / epilogue phases (the one-tile-per-CTA kernel has no intra-CTA overlap of those)
__host__ __device__ inline int umma_ring_depth(int kiters, int max_stages) {
  const int want = kiters <= 8 ? (kiters < 2 ? kiters : 2) : max_stages;
  return want < max_stages ? want : max_stages;
}

template <int BN, int KC>
struct UGeom {
  static constexpr int SWB = KC * 2;
  static constexpr int A_BYTES = 128 * SWB, B_BYTES = BN * SWB, STAGE = A_BYTES + B_BYTES;
  static constexpr int BUDGET = BN > 128 ? 196 * 1024 : 100 * 1024;
  static constexpr int STAGES_ = (BUDGET - 2048) / STAGE;
  static constexpr int STAGES = STAGES_ > 8 ? 8 : STAGES_;
  static constexpr int SMEM = 1024 + STAGES * STAGE + 1024;     // + alignment slack
  static constexpr int TCOLS = BN < 32 ? 32 : BN;
};

template <int BN, int KC>
__global__ void __launch_bounds__(192) conv_umma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                       const __grid_constant__ CUtensorMap mapB,
                                                       const __grid_constant__ CUtensorMap mapW, UArgs a) {
  using G = UGeom<BN, KC>;
  constexpr int SWB = G::SWB, STAGES = G::STAGES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned stage ring (128B swizzle atoms), control block in the first 1024 bytes
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);            // [STAGES]
  uint64_t* empty = full + STAGES;                               // [STAGES]
  uint64_t* accum = empty + STAGES;                              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  unsigned char* ring = smem + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int b = blockIdx.x;
  const int tix = b % a.tiles_x; b /= a.tiles_x;
  const int tiy = b % a.tiles_y;
  const int n = b / a.tiles_y;
  const int x0 = tix * 16, y0 = tiy * 8;
  const int n0 = blockIdx.y * BN;

  const int kca = a.c_a / KC, kcb = a.c_b / KC;
  const int kiters = a.taps * (kca + kcb);
  // ring depth = min(STAGES, K iterations): short-K launches (ConvT fprop, 1x1) then need little shared memory and
  // several CTAs share an SM, overlapping each other's load / MMA / epilogue phases
  const int nst = umma_ring_depth(kiters, STAGES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, G::TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapW);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % nst;
        if (it >= nst) mbar_wait(empty + s, ((it / nst) - 1) & 1);
        // iteration order: input (A then B) -> tap -> channel chunk
        int r = it;
        const bool second = r >= a.taps * kca;
        if (second) r -= a.taps * kca;
        const int kcn = second ? kcb : kca;
        const int tap = r / kcn, kc = r % kcn;
        const int ox = (tap % a.ktap) + a.offbase, oy = (tap / a.ktap) + a.offbase;
        unsigned char* sa = ring + s * G::STAGE;
        mbar_expect_tx(full + s, G::STAGE);
        tma_load_4d(sa, second ? &mapB : &mapA, full + s, kc * KC, x0 * a.sx + ox, y0 * a.sx + oy, n);
        // weights: {K (all inputs), N, tap}
        const int kglob = (second ? a.c_a : 0) + kc * KC;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(sa + G::A_BYTES)), "l"(reinterpret_cast<uint64_t>(&mapW)), "r"(smem_u32(full + s)), "r"(kglob),
            "r"(n0), "r"(tap)
            : "memory");
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (descriptors in uniform registers; an
    // `if (lane == 0)` body costs ~100 cycles of issue per MMA, more than these N <= 64 MMAs take) =====
    {
      constexpr uint32_t idesc = make_idesc(128, BN < 16 ? 16 : BN, 0, 0);
      // descriptor high word: SBO = 8 rows of SWB bytes, version 1, swizzle mode (see kmajor_desc)
      constexpr uint32_t dhi = ((8u * SWB) >> 4) | (1u << 14) | ((SWB == 128 ? 2u : (SWB == 64 ? 4u : 6u)) << 29);
      const uint32_t leader = umma_elect() ? 1u : 0u;
      const bool committer = umma_elect();
      for (int it = 0; it < kiters; ++it) {
        const int s = it % nst;
        mbar_wait(full + s, (it / nst) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring + s * G::STAGE), sb = sa + G::A_BYTES;
        const uint32_t da = kmajor128_desc_lo(sa), db = kmajor128_desc_lo(sb);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k)   // +32 bytes per K=16 step inside the swizzle row (start-address field is >>4)
          umma_bf16_uniform(tmem_base, da + (uint32_t)(2 * k), dhi, db + (uint32_t)(2 * k), dhi, idesc, (it | k) ? 1u : 0u, leader);
        if (committer) umma_commit(empty + s);             // frees the stage when these MMAs retire
        __syncwarp();
      }
      if (committer) umma_commit(accum);                   // accumulator complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32*(warp%4).. =====
    const int lg = warp & 3;
    const int r = lg * 32 + lane;           // row of the 128-pixel tile
    const int ty = r >> 4, tx = r & 15;
    const int gy = y0 + ty, gx = x0 + tx;
    const bool inside = gy < a.H && gx < a.W;
    mbar_wait(accum, 0);
    tc_fence_after();
    constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CH) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
      if (CH == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
      tmem_ld_wait();
      const int ncol = n0 + c0;               // first GEMM column of this chunk
      if (!inside || ncol >= a.n_total) continue;
      epilogue_chunk<CH>(a, v, n, gy, gx, ncol, a.n_total - ncol);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, G::TCOLS);
}

static bool bf16_view16(const dnnca_tensor_t* t);
static bool bf16_view8(const dnnca_tensor_t* t);
static bool bf16_view_in(const dnnca_tensor_t* t);

// ---------------------------------------------------------------- wgrad
// dW[tap][ci][co] += sum_pixels xs_tap[p][ci] * dz[p][co]        (Conv2D, MODE 0; ConvT uses xs = x, dz at 2p+tap)
// GEMM with M = 128 input channels, N = BN output channels, K = pixels.  Both operands are "MN-major":
// a TMA box {64 channels, 16 px, 4 rows} lands as 64 pixel rows of 128 bytes = eight 1024-byte SWIZZLE_128B
// atoms (8 pixels x 64 channels), exactly the canonical MN-major UMMA layout with SBO = 1024 (next 8 pixels)
// and LBO = 8192 (next 64 channels).  Channels beyond the tensor are zero-filled by the TMA, so any channel
// count that is a multiple of 16 works (at the price of idle MMA rows).  A CTA owns one (M tile, N tile, tap)
// and a slice of the pixel tiles; partial sums are added to the fp32 gradient with red.global.add.
struct WArgs {
  int taps, ktap, offbase, sx;     // conv: 9,3,-1,1 ; ConvT: 4,2,0,2 (dz sampled with traversal stride 2)
  int c_a, c_b, cout;              // channels of x, x2; output channels
  int tiles_x, tiles_y, nimg;      // 16x4 pixel tiles over the x grid
  int ksplit;
  int tconv;                       // output index order: conv [tap][ci][co], ConvT [tap][co][ci]
  float* dw;
};

template <int BN>
__global__ void __launch_bounds__(192) wgrad_umma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                        const __grid_constant__ CUtensorMap mapB,
                                                        const __grid_constant__ CUtensorMap mapG, WArgs a) {
  constexpr int ATOM = 8192;                       // 64 px x 128 B
  constexpr int NB = BN / 64;
  constexpr int STAGE = (2 + NB) * ATOM;
  constexpr int STAGES = BN > 128 ? 4 : 3;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + STAGES;
  uint64_t* accum = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  unsigned char* ring = smem + 1024;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // M tiles cover x first, then x2 (each tensor padded to whole 128-row tiles by the TMA's zero fill)
  const int mta = (a.c_a + 127) / 128;
  const bool second = (int)blockIdx.x >= mta;
  const int m0 = (second ? (int)blockIdx.x - mta : (int)blockIdx.x) * 128, n0 = blockIdx.y * BN;
  const int tap = blockIdx.z / a.ksplit, ks = blockIdx.z % a.ksplit;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  const int per = (ntiles + a.ksplit - 1) / a.ksplit;
  const int t_beg = ks * per, t_end = min(ntiles, t_beg + per);
  const int kiters = t_end - t_beg;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (kiters > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int ox = (tap % a.ktap) + a.offbase, oy = (tap / a.ktap) + a.offbase;
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          if (it >= STAGES) mbar_wait(empty + s, ((it / STAGES) - 1) & 1);
          int b = t_beg + it;
          const int tix = b % a.tiles_x; b /= a.tiles_x;
          const int tiy = b % a.tiles_y;
          const int n = b / a.tiles_y;
          const int x0 = tix * 16, y0 = tiy * 4;
          unsigned char* st = ring + s * STAGE;
          mbar_expect_tx(full + s, STAGE);
          // A: two 64-channel atoms of the (virtually concatenated) input, shifted by the tap for Conv2D
          const int xa = a.tconv ? x0 : x0 + ox, ya = a.tconv ? y0 : y0 + oy;
#pragma unroll
          for (int h = 0; h < 2; ++h)      // channels >= C read zeros
            tma_load_4d(st + h * ATOM, second ? &mapB : &mapA, full + s, m0 + 64 * h, xa, ya, n);
          const int xg = a.tconv ? 2 * x0 + ox : x0, yg = a.tconv ? 2 * y0 + oy : y0;
#pragma unroll
          for (int h = 0; h < NB; ++h) tma_load_4d(st + (2 + h) * ATOM, &mapG, full + s, n0 + 64 * h, xg, yg, n);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(128, BN, 1, 1);
        for (int it = 0; it < kiters; ++it) {
          const int s = it % STAGES;
          mbar_wait(full + s, (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + s * STAGE), sb = sa + 2 * ATOM;
          // MN-major SWIZZLE_128B: LBO = next 64-channel atom, SBO = next 8 pixels
          auto mn_desc = [](uint32_t addr) {
            uint64_t d = 0;
            d |= (uint64_t)((addr & 0x3FFFF) >> 4);
            d |= (uint64_t)(ATOM >> 4) << 16;
            d |= (uint64_t)(1024 >> 4) << 32;
            d |= (uint64_t)1 << 46;
            d |= (uint64_t)2 << 61;
            return d;
          };
#pragma unroll
          for (int k = 0; k < 4; ++k)       // 16 pixels (2 atoms along K) per MMA: +2048 bytes
            umma_bf16(tmem_base, mn_desc(sa + k * 2048), mn_desc(sb + k * 2048), idesc, (it | k) ? 1u : 0u);
          umma_commit(empty + s);
        }
        umma_commit(accum);
      }
    } else {
      const int lg = warp & 3;
      const int row = m0 + lg * 32 + lane;      // channel inside its tensor (row of the accumulator)
      const bool live = row < (second ? a.c_b : a.c_a);
      const int m = row + (second ? a.c_a : 0); // input channel of the (virtually concatenated) layer
      mbar_wait(accum, 0);
      tc_fence_after();
      const int cin = a.c_a + a.c_b;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!live) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int co = n0 + c0 + j;
          if (co < a.cout) {
            float* dst = a.tconv ? a.dw + ((size_t)tap * a.cout + co) * cin + m : a.dw + ((size_t)tap * cin + m) * a.cout + co;
            atomicAdd(dst, __uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

static bool act_map64(CUtensorMap* m, const dnnca_tensor_t* t, int estride) {
  // 4-D {C, W, H, N}; box {64 channels, 16 px, 4 rows, 1}, SWIZZLE_128B; channels >= C read as zero
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  char* base = reinterpret_cast<char*>(t->data) + (size_t)t->coff * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (t->cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(16 * estride), (cuuint32_t)(4 * estride), 1};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
static int launch_wgrad_umma(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mG, WArgs a) {
  constexpr int STAGE = (2 + BN / 64) * 8192, STAGES = BN > 128 ? 4 : 3, SMEM = 1024 + STAGES * STAGE + 1024;
  auto kern = wgrad_umma_kernel<BN>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "wgrad_umma: cudaFuncSetAttribute");
    done = true;
  }
  const int mt = (a.c_a + 127) / 128 + (a.c_b + 127) / 128, nt = (a.cout + BN - 1) / BN;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  long long want = (3LL * sm_count() + (long long)mt * nt * a.taps - 1) / ((long long)mt * nt * a.taps);
  if (want > ntiles / 4) want = ntiles / 4;
  if (want < 1) want = 1;
  if ((long long)a.taps * want > 65535) want = 65535 / a.taps;
  a.ksplit = (int)want;
  dim3 grid(mt, nt, a.taps * a.ksplit);
  kern<<<grid, 192, SMEM, s>>>(mA, mB, mG, a);
  DNNCA_LAUNCH_CHECK("wgrad_umma");
  note_family(2);
  return 1;
}

int launch_channel_sum(cudaStream_t s, const dnnca_tensor_t* g, float* out);
int try_conv3x3_wgrad_halo(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* g, float* dw,
                           float* db);
int try_tconv_wgrad_halo(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk);
static bool wgrad_halo_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("DNNCA_DISABLE_WGRAD_HALO") ? 0 : 1;
  return on == 1;
}

static int wgrad_umma_common(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* g,
                             float* dw, float* db, int k, int tconv) {
  if (!bf16_view_in(x) || (x2 && !bf16_view_in(x2)) || !bf16_view_in(g)) return 0;
  const int ca = x->c, cb = x2 ? x2->c : 0;
  if (!tconv && k == 3 && wgrad_halo_enabled()) {     // halo-tile kernel (conv_umma3.cu) for channel counts % 64 == 0
    int r = try_conv3x3_wgrad_halo(s, x, x2, g, dw, db);
    if (r < 0) return r;
    if (r == 2) return 1;                                  // bias gradient taken inside the kernel
    if (r == 1) {
      if (db) {
        int e = launch_channel_sum(s, g, db);
        if (e != DNNCA_OK) return e;
      }
      return 1;
    }
  }
  if (tconv && wgrad_halo_enabled()) {                  // ConvT on the halo-tile kernel: Cin % 128 == 0, Cout % 64 == 0
    int r = try_tconv_wgrad_halo(s, x, g, dw);
    if (r < 0) return r;
    if (r == 1) {
      if (db) {
        int e = launch_channel_sum(s, g, db);
        if (e != DNNCA_OK) return e;
      }
      return 1;
    }
  }
  // partial 16x4 pixel tiles are fine: the TMA zero-fills x and dz outside the image, so they add nothing
  CUtensorMap mA, mB, mG;
  if (!act_map64(&mA, x, 1)) return 0;
  mB = mA;
  if (x2 && !act_map64(&mB, x2, 1)) return 0;
  if (!act_map64(&mG, g, tconv ? 2 : 1)) return 0;
  WArgs a{};
  a.taps = tconv ? 4 : k * k; a.ktap = tconv ? 2 : k; a.offbase = tconv ? 0 : -(k / 2); a.sx = tconv ? 2 : 1;
  a.c_a = ca; a.c_b = cb; a.cout = g->c; a.tiles_x = (x->w + 15) / 16; a.tiles_y = (x->h + 3) / 4; a.nimg = x->n; a.tconv = tconv; a.dw = dw;
  int r;
  const int cout = g->c;
  if (cout > 128) r = launch_wgrad_umma<256>(s, mA, mB, mG, a);
  else if (cout > 64) r = launch_wgrad_umma<128>(s, mA, mB, mG, a);
  else r = launch_wgrad_umma<64>(s, mA, mB, mG, a);
  if (r != 1) return r;
  if (db) {
    int e = launch_channel_sum(s, g, db);
    if (e != DNNCA_OK) return e;
  }
  return 1;
}

int try_conv_wgrad_umma(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* dz,
                        float* dw, float* db, int k) {
  return wgrad_umma_common(s, x, x2, dz, dw, db, k, 0);
}
int try_tconv_wgrad_umma(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk, float* db) {
  return wgrad_umma_common(s, x, nullptr, dy, dk, db, 2, 1);
}

// ---------------------------------------------------------------- host side
static bool act_map(CUtensorMap* m, const dnnca_tensor_t* t, int kc, int estride) {
  // 4-D {C, W, H, N} view of a (channel-sliced) NHWC bf16 tensor; box {kc, 16, 8, 1}
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  char* base = reinterpret_cast<char*>(t->data) + (size_t)t->coff * 2;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (t->cstride * 2) % 16) return false;
  cuuint64_t dims[4] = {(cuuint64_t)t->c, (cuuint64_t)t->w, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {(cuuint64_t)t->cstride * 2, (cuuint64_t)t->w * t->cstride * 2, (cuuint64_t)t->h * t->w * t->cstride * 2};
  cuuint32_t box[4] = {(cuuint32_t)kc, 16, 8, 1};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  // with a traversal stride the box extent is given in tensor elements (16 px * stride)
  box[1] = 16 * estride;
  box[2] = 8 * estride;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool weight_map(CUtensorMap* m, const void* wp, int ktot, int ntot, int taps, int kc, int bn) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)ktot, (cuuint64_t)ntot, (cuuint64_t)taps};
  cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)ktot * ntot * 2};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)bn, 1};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wp), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int KC>
static int launch_umma(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                       int nimg) {
  using G = UGeom<BN, KC>;
  auto kern = conv_umma_kernel<BN, KC>;
  static bool done = false;
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return cuda_fail(e, "conv_umma: cudaFuncSetAttribute");
    done = true;
  }
  dim3 grid((unsigned)((long long)a.tiles_x * a.tiles_y * nimg), (unsigned)((a.n_total + BN - 1) / BN));
  const int kiters = a.taps * (a.c_a / KC + a.c_b / KC);
  const int nst = umma_ring_depth(kiters, G::STAGES);
  kern<<<grid, 192, 1024 + nst * G::STAGE + 1024, s>>>(mA, mB, mW, a);
  DNNCA_LAUNCH_CHECK("conv_umma");
  note_family(2);
  return 1;
}

static int pick_kc(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }
static int gcd_kc(int a, int b) { return a < b ? a : b; }

template <int KC>
static int dispatch_bn(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mW, const UArgs& a,
                       int nimg, int bn) {
  switch (bn) {
    case 256: return launch_umma<256, KC>(s, mA, mB, mW, a, nimg);
    case 128: return launch_umma<128, KC>(s, mA, mB, mW, a, nimg);
    case 64: return launch_umma<64, KC>(s, mA, mB, mW, a, nimg);
    case 32: return launch_umma<32, KC>(s, mA, mB, mW, a, nimg);
    case 16: return launch_umma<16, KC>(s, mA, mB, mW, a, nimg);
  }
  return 0;
}

// N tile: the largest of 256/128/64/32/16 dividing `unit` (columns that must not straddle a destination boundary)
static int pick_bn(int unit, int ntotal) {
  for (int bn : {256, 128, 64, 32, 16})
    if (unit % bn == 0 && bn <= ntotal) return bn;
  return 0;
}

static int pad16(int c) { return (c + 15) / 16 * 16; }
static int pad64(int c) { return (c + 63) / 64 * 64; }
// room for either K padding (multiples of 16, or of 64 when pick_kchunk prefers 64-channel chunks)
size_t umma_pack_bytes(int taps, int cin, int cout) { return (size_t)taps * pad64(cin) * pad16(cout) * 2; }

// K chunk (TMA box / swizzle width) of a single-input operand with ANY channel count: channels beyond the tensor are
// zero-filled by the TMA (activations) and by the pack kernel (weights), so K may be padded up to whole chunks.
// cost = chunks * (fixed per-chunk pipeline step ~ 32 channels of work + chunk width); MultiResUnet's 8/24/40/72/120/144
// /216/432/864-channel tensors otherwise fall to 16-channel chunks (9 of them for 144 channels)
static int pick_kchunk(int c, int* kp) {
  int best = 0, bestcost = 1 << 30;
  for (int kc : {64, 32, 16}) {
    const int n = (c + kc - 1) / kc, cost = n * (32 + kc);
    if (cost < bestcost) { bestcost = cost; best = kc; *kp = n * kc; }
  }
  return best;
}
// N tile of fprop / ConvT fprop for any N that is a multiple of 8: the smallest of 16..256 covering it, 256-column tiles
// beyond (the weight TMA zero-fills rows >= N, the epilogue stores 8-column groups below N only)
static int pick_bn_cover(int n) {
  for (int bn : {16, 32, 64, 128, 256})
    if (n <= bn) return bn;
  return 256;
}

static int pack(cudaStream_t s, const float* w, void* out, int mode, int taps, int cin, int cout, int kp = 0,
                const float* sa = nullptr, int ca = 0, const float* sb = nullptr) {
  if (!w) return DNNCA_OK;          // workspace still holds the packing of an earlier call (dnnca.h: inference with unchanged weights)
  if (kp <= 0) kp = pad16(cin);
  const long long total = (long long)taps * kp * cout;
  pack_weights_kernel<<<grid_for(total, 256 * 4, 4), 256, 0, s>>>(w, reinterpret_cast<__nv_bfloat16*>(out), mode, taps, cin, cout, kp,
                                                                  sa, sa || sb ? ca : cin, sb);
  DNNCA_LAUNCH_CHECK("pack_weights");
  return DNNCA_OK;
}

static bool bf16_view16(const dnnca_tensor_t* t) {
  return t && t->dtype == DNNCA_BF16 && t->c % 16 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}
// output written in 16-byte groups of 8 channels
static bool bf16_view8(const dnnca_tensor_t* t) {
  return t && t->dtype == DNNCA_BF16 && t->c % 8 == 0 && t->coff % 8 == 0 && t->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}
// operand that is only READ through the TMA: any channel count (missing channels are zero-filled out of bounds)
static bool bf16_view_in(const dnnca_tensor_t* t) {
  return t && t->dtype == DNNCA_BF16 && t->coff % 8 == 0 && t->cstride % 8 == 0 && (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}

int try_conv3x3_halo(cudaStream_t s, const dnnca_tensor_t* xa, const dnnca_tensor_t* xb, const void* wpack, int ktot, int ntot,
                     UArgs a);
int try_tconv_fprop_halo(cudaStream_t s, const dnnca_tensor_t* x, const void* wpack, int cin, int cout, UArgs a);
int try_conv1x1_halo(cudaStream_t s, const dnnca_tensor_t* x, const void* wpack, int ktot, int ntot, UArgs a);
int try_tconv_dgrad_halo(cudaStream_t s, const dnnca_tensor_t* dy, const void* wpack, int cin, int cout, UArgs a);
static bool halo_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("DNNCA_NO_HALO") ? 0 : 1;     // A/B switch for profiling the first-generation kernel
  return v == 1;
}

static thread_local bool g_pack_only = false;
// BatchNorm folded into the conv input (dnnca_conv2d_fprop_affine): scale vectors for the weight packing and the
// per-border-class bias table for the epilogue; set by fprop_umma_affine around one try_conv_fprop_umma call.  With a
// fold active only the persistent halo kernel may serve the layer (it alone carries the class-bias epilogue).
struct FoldCtx { const float* sa; const float* sb; const float* bias9; };
static thread_local const FoldCtx* g_fold = nullptr;

// BatchNorm backward reduction fused into a dgrad epilogue (dnnca_conv2d_dgrad_bnreduce / dnnca_convtranspose2x2_dgrad_bnreduce):
// the BN's input tensor, its [2C] mean|invstd and the [2C] fp64 sums; `done` is set when the persistent halo kernel took it
struct BnrCtx { const dnnca_tensor_t* x; const float* mi; double* sums; bool done; };
static thread_local BnrCtx* g_bnr = nullptr;
static void bnr_arm(UArgs& a, const dnnca_tensor_t* dx) {
  a.epi = EPI_DGRAD_BNR; a.act = DNNCA_ACT_NONE; a.alpha = 0.f;
  a.mask = reinterpret_cast<const __nv_bfloat16*>(g_bnr->x->data) + g_bnr->x->coff; a.mask_cs = g_bnr->x->cstride;
  a.stats = g_bnr->sums; a.bnr_mi = g_bnr->mi;
  (void)dx;
}
static void bnr_disarm(UArgs& a) { a.epi = EPI_DGRAD; a.mask = nullptr; a.mask_cs = 0; a.stats = nullptr; a.bnr_mi = nullptr; }
static bool bnr_usable(const dnnca_tensor_t* dx, const dnnca_tensor_t* mask) {
  return g_bnr && !mask && g_bnr->x->dtype == DNNCA_BF16 && g_bnr->x->c == dx->c && g_bnr->x->n == dx->n && g_bnr->x->h == dx->h &&
         g_bnr->x->w == dx->w && g_bnr->x->coff % 8 == 0 && g_bnr->x->cstride % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(g_bnr->x->data) & 15) == 0;
}

// returns 1 handled / 0 not covered / <0 error
// returns 2 when the kernel also accumulated the BatchNorm statistics into `stats` (else the caller runs channel_stats)
int try_conv_fprop_umma(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w, const float* bias,
                        const dnnca_tensor_t* y, int k, int act, float alpha, void* ws, size_t ws_bytes, double* stats) {
  if (!ws || !bf16_view8(y)) return 0;
  if (x2 ? (!bf16_view16(x) || !bf16_view16(x2) || !bf16_view16(y)) : !bf16_view_in(x)) return 0;
  const int ca = x->c, cb = x2 ? x2->c : 0, cin = ca + cb, cout = y->c, taps = k * k;
  if (ws_bytes < umma_pack_bytes(taps, cin, cout)) return 0;
  int kp = cin;
  const int kc = cb ? gcd_kc(pick_kc(ca), pick_kc(cb)) : pick_kchunk(ca, &kp);
  const int exact = pick_bn(cout, cout);
  const int bn = (exact >= 32 || exact == cout) ? exact : pick_bn_cover(cout);      // a dividing tile of >= 32 columns, else one covering tile
  if (!kc || !bn) return 0;
  CUtensorMap mA, mB, mW;
  if (!act_map(&mA, x, kc, 1)) return 0;
  mB = mA;
  if (x2 && !act_map(&mB, x2, kc, 1)) return 0;
  if (!weight_map(&mW, ws, kp, cout, taps, kc, bn)) return 0;
  int r = g_fold ? pack(s, w, ws, 0, taps, cin, cout, kp, g_fold->sa, ca, g_fold->sb) : pack(s, w, ws, 0, taps, cin, cout, kp);
  if (r != DNNCA_OK) return r;
  if (g_pack_only) return 1;            // dnnca_conv2d_prepack: the packing a later w == NULL call of this layer expects
  UArgs a{};
  a.taps = taps; a.ktap = k; a.sx = 1; a.offbase = -(k / 2); a.c_a = cb ? ca : kp; a.c_b = cb; a.H = x->h; a.W = x->w;
  a.tiles_x = (x->w + 15) / 16; a.tiles_y = (x->h + 7) / 8; a.epi = EPI_FPROP; a.act = act; a.alpha = alpha; a.bias = bias;
  a.ya = reinterpret_cast<__nv_bfloat16*>(y->data) + y->coff; a.ya_cs = y->cstride; a.split = cout; a.yb = a.ya; a.yb_cs = a.ya_cs;
  a.mask = nullptr; a.n_total = cout; a.cout_t = cout; a.nimg = x->n;
  a.bias9 = g_fold ? g_fold->bias9 : nullptr;
  if (k == 3 && (kc == 64 || (!x2 && ca < 64)) && halo_enabled()) {
    a.stats = stats;
    r = try_conv3x3_halo(s, x, x2, ws, kp, cout, a);
    if (r != 0) return (r == 1 && stats) ? 2 : r;
    a.stats = nullptr;
  }
  if (g_fold) return 0;                 // no other kernel applies the folded input affine
  if (k == 1 && !x2 && kc == 64 && halo_enabled()) {      // 1x1 convs (MultiResUnet shortcuts) as a persistent plain GEMM
    r = try_conv1x1_halo(s, x, ws, kp, cout, a);
    if (r != 0) return r;
  }
  if (kc == 64) return dispatch_bn<64>(s, mA, mB, mW, a, x->n, bn);
  if (kc == 32) return dispatch_bn<32>(s, mA, mB, mW, a, x->n, bn);
  return dispatch_bn<16>(s, mA, mB, mW, a, x->n, bn);
}

int try_conv_dgrad_umma(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                        const dnnca_tensor_t* dx2, int k, const dnnca_tensor_t* mask, int act, float alpha, void* ws,
                        size_t ws_bytes) {
  if (!ws || !bf16_view16(dz) || !bf16_view16(dx) || (dx2 && !bf16_view16(dx2)) || (mask && !bf16_view16(mask))) return 0;
  const int cout = dz->c, ca = dx->c, cb = dx2 ? dx2->c : 0, cin = ca + cb, taps = k * k;
  if (ws_bytes < umma_pack_bytes(taps, cin, cout)) return 0;
  const int kc = pick_kc(cout);
  const int bn = pick_bn(cb ? (ca < cb ? ca : cb) : ca, cin);
  if (!kc || !bn || (cb && (ca % bn || cb % bn))) return 0;
  int r = pack(s, w, ws, 1, taps, cin, cout);     // [tap'][Cin][Cout]: GEMM N = layer Cin, K = layer Cout
  if (r != DNNCA_OK) return r;
  CUtensorMap mA, mW;
  if (!act_map(&mA, dz, kc, 1)) return 0;
  if (!weight_map(&mW, ws, cout, cin, taps, kc, bn)) return 0;
  UArgs a{};
  a.taps = taps; a.ktap = k; a.sx = 1; a.offbase = -(k / 2); a.c_a = cout; a.c_b = 0; a.H = dx->h; a.W = dx->w;
  a.tiles_x = (dx->w + 15) / 16; a.tiles_y = (dx->h + 7) / 8; a.epi = EPI_DGRAD; a.act = act; a.alpha = alpha; a.bias = nullptr;
  a.ya = reinterpret_cast<__nv_bfloat16*>(dx->data) + dx->coff; a.ya_cs = dx->cstride; a.split = ca;
  a.yb = dx2 ? reinterpret_cast<__nv_bfloat16*>(dx2->data) + dx2->coff : a.ya; a.yb_cs = dx2 ? dx2->cstride : a.ya_cs;
  a.mask = mask ? reinterpret_cast<const __nv_bfloat16*>(mask->data) + mask->coff : nullptr; a.mask_cs = mask ? mask->cstride : 0;
  a.n_total = cin; a.cout_t = cin; a.nimg = dx->n;
  if (k == 3 && (kc == 64 || cout < 64) && halo_enabled()) {     // cout < 64: one K chunk, missing channels zero-filled
    const bool bnr = bnr_usable(dx, mask);
    if (bnr) bnr_arm(a, dx);
    r = try_conv3x3_halo(s, dz, nullptr, ws, cout, cin, a);
    if (bnr) { g_bnr->done = r == 1; bnr_disarm(a); }
    if (r != 0) return r;
  }
  if (kc == 64) return dispatch_bn<64>(s, mA, mA, mW, a, dx->n, bn);
  if (kc == 32) return dispatch_bn<32>(s, mA, mA, mW, a, dx->n, bn);
  return dispatch_bn<16>(s, mA, mA, mW, a, dx->n, bn);
}

// returns 2 when the kernel also accumulated the BatchNorm statistics of y into `stats`
int try_tconv_fprop_umma(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const float* bias, const dnnca_tensor_t* y,
                         void* ws, size_t ws_bytes, double* stats) {
  if (!ws || !bf16_view_in(x) || !bf16_view16(y)) return 0;
  const int cin = x->c, cout = y->c;
  if (ws_bytes < umma_pack_bytes(4, cin, cout)) return 0;
  int kp = cin;
  const int kc = cin % 16 == 0 ? pick_kc(cin) : pick_kchunk(cin, &kp);      // any input width: K padded with zeros
  const int bn = pick_bn(cout, 4 * cout);
  if (!kc || !bn) return 0;
  CUtensorMap mA, mW;
  if (!act_map(&mA, x, kc, 1)) return 0;
  if (!weight_map(&mW, ws, kp, 4 * cout, 1, kc, bn)) return 0;
  int r = pack(s, kw, ws, 2, 4, cin, cout, kp);   // [tap*Cout + co][Cin]: one GEMM with N = 4*Cout
  if (r != DNNCA_OK) return r;
  if (g_pack_only) return 1;
  UArgs a{};
  a.taps = 1; a.ktap = 1; a.sx = 1; a.offbase = 0; a.c_a = kp; a.c_b = 0; a.H = x->h; a.W = x->w;
  a.tiles_x = (x->w + 15) / 16; a.tiles_y = (x->h + 7) / 8; a.epi = EPI_TCONV; a.act = DNNCA_ACT_NONE; a.alpha = 0.f; a.bias = bias;
  a.ya = reinterpret_cast<__nv_bfloat16*>(y->data) + y->coff; a.ya_cs = y->cstride; a.split = 4 * cout; a.yb = a.ya; a.yb_cs = a.ya_cs;
  a.mask = nullptr; a.n_total = 4 * cout; a.cout_t = cout;
  if (kc == 64 && kp == cin && halo_enabled()) {
    a.stats = stats;
    r = try_tconv_fprop_halo(s, x, ws, cin, cout, a);
    if (r != 0) return r;
    a.stats = nullptr;
  }
  if (kc == 64) return dispatch_bn<64>(s, mA, mA, mW, a, x->n, bn);
  if (kc == 32) return dispatch_bn<32>(s, mA, mA, mW, a, x->n, bn);
  return dispatch_bn<16>(s, mA, mA, mW, a, x->n, bn);
}

int try_tconv_dgrad_umma(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx,
                         const dnnca_tensor_t* mask, int act, float alpha, void* ws, size_t ws_bytes) {
  if (!ws || !bf16_view16(dy) || !bf16_view16(dx) || (mask && !bf16_view16(mask))) return 0;
  const int cout = dy->c, cin = dx->c;
  if (ws_bytes < umma_pack_bytes(4, cin, cout)) return 0;
  const int kc = pick_kc(cout);
  const int bn = pick_bn(cin, cin);
  if (!kc || !bn) return 0;
  int r = pack(s, kw, ws, 3, 4, cin, cout);       // [tap][Cin][Cout]
  if (r != DNNCA_OK) return r;
  CUtensorMap mA, mW;
  if (!act_map(&mA, dy, kc, 2)) return 0;          // traversal stride 2: one 2x2 tap of dy per box
  if (!weight_map(&mW, ws, cout, cin, 4, kc, bn)) return 0;
  UArgs a{};
  a.taps = 4; a.ktap = 2; a.sx = 2; a.offbase = 0; a.c_a = cout; a.c_b = 0; a.H = dx->h; a.W = dx->w;
  a.tiles_x = (dx->w + 15) / 16; a.tiles_y = (dx->h + 7) / 8; a.epi = EPI_DGRAD; a.act = act; a.alpha = alpha; a.bias = nullptr;
  a.ya = reinterpret_cast<__nv_bfloat16*>(dx->data) + dx->coff; a.ya_cs = dx->cstride; a.split = cin; a.yb = a.ya; a.yb_cs = a.ya_cs;
  a.mask = mask ? reinterpret_cast<const __nv_bfloat16*>(mask->data) + mask->coff : nullptr; a.mask_cs = mask ? mask->cstride : 0;
  a.n_total = cin; a.cout_t = cin; a.nimg = dx->n;
  if (kc == 64 && halo_enabled()) {
    const bool bnr = bnr_usable(dx, mask);
    if (bnr) bnr_arm(a, dx);
    r = try_tconv_dgrad_halo(s, dy, ws, cin, cout, a);
    if (bnr) { g_bnr->done = r == 1; bnr_disarm(a); }
    if (r != 0) return r;
  }
  if (kc == 64) return dispatch_bn<64>(s, mA, mA, mW, a, dx->n, bn);
  if (kc == 32) return dispatch_bn<32>(s, mA, mA, mW, a, dx->n, bn);
  return dispatch_bn<16>(s, mA, mA, mW, a, dx->n, bn);
}

// Conv2D 3x3 fprop whose input(s) carry a folded BatchNorm affine; 1 / 2 (statistics taken) handled, 0 not served
int fprop_umma_affine(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w, const dnnca_tensor_t* y,
                      int act, float alpha, void* ws, size_t ws_bytes, double* stats, const float* scale_a,
                      const float* scale_b, const float* bias9) {
  FoldCtx f{scale_a, scale_b, bias9};
  g_fold = &f;
  const int r = try_conv_fprop_umma(s, x, x2, w, nullptr, y, 3, act, alpha, ws, ws_bytes, stats);
  g_fold = nullptr;
  return r;
}
// dgrad with the BatchNorm backward reduction of destination A in its epilogue: returns what the plain try_* returns and
// sets *fused when the sums were taken (else the caller runs bn_bwd_reduce)
int conv_dgrad_umma_bnreduce(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                             const dnnca_tensor_t* dx2, int k, void* ws, size_t ws_bytes, const dnnca_tensor_t* bn_x,
                             const float* mi, double* sums, bool* fused) {
  BnrCtx c{bn_x, mi, sums, false};
  g_bnr = &c;
  const int r = try_conv_dgrad_umma(s, dz, w, dx, dx2, k, nullptr, DNNCA_ACT_NONE, 0.f, ws, ws_bytes);
  g_bnr = nullptr;
  *fused = c.done;
  return r;
}
int tconv_dgrad_umma_bnreduce(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx, void* ws,
                              size_t ws_bytes, const dnnca_tensor_t* bn_x, const float* mi, double* sums, bool* fused) {
  BnrCtx c{bn_x, mi, sums, false};
  g_bnr = &c;
  const int r = try_tconv_dgrad_umma(s, dy, kw, dx, nullptr, DNNCA_ACT_NONE, 0.f, ws, ws_bytes);
  g_bnr = nullptr;
  *fused = c.done;
  return r;
}

// would fprop_umma_affine serve this shape? (host-side checks only; mirrors try_conv_fprop_umma + try_conv3x3_halo)
int fprop_umma_affine_supported(const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* y) {
  if (!halo_enabled() || !bf16_view8(y)) return 0;
  if (x2 ? (!bf16_view16(x) || !bf16_view16(x2) || !bf16_view16(y) || x->c % 64 || x2->c % 64) : !bf16_view_in(x)) return 0;
  if (x->h < 2 || x->w < 2) return 0;
  int kp = x->c;
  const int kc = x2 ? 64 : pick_kchunk(x->c, &kp);
  if (!(kc == 64 || (!x2 && x->c < 64))) return 0;
  return (x2 || x->c < 64 || kp % 64 == 0) ? 1 : 0;
}

// packing only: what dnnca_conv2d_prepack / dnnca_convtranspose2x2_prepack run (1 packed / 0 shape not served / <0 error)
int prepack_conv_fprop_umma(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                            const dnnca_tensor_t* y, int k, void* ws, size_t ws_bytes) {
  g_pack_only = true;
  const int r = try_conv_fprop_umma(s, x, x2, w, nullptr, y, k, DNNCA_ACT_NONE, 0.f, ws, ws_bytes, nullptr);
  g_pack_only = false;
  return r;
}
int prepack_tconv_fprop_umma(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const dnnca_tensor_t* y, void* ws,
                             size_t ws_bytes) {
  g_pack_only = true;
  const int r = try_tconv_fprop_umma(s, x, kw, nullptr, y, ws, ws_bytes, nullptr);
  g_pack_only = false;
  return r;
}

}  // namespace dnnca
