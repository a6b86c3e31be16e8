// tcgen05 / TMEM PTX wrappers and UMMA descriptor helpers shared by conv_umma.cu and conv_umma2.cu
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace dnnca {

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// Warp-uniform issue: the WHOLE warp runs the issue loop and one elected lane (`leader`) executes the MMA, so the
// descriptors stay in uniform registers (an `if (lane == 0)` body keeps them in vector registers and pays an R2UR
// round trip per MMA: ~100 cycles of issue per MMA measured, which exposes N <= 128 MMAs).  Descriptors are passed as
// 32-bit halves: `*_lo` = start address >> 4 | LBO field, `*_hi` = SBO field | version | swizzle mode.
__device__ __forceinline__ bool umma_elect() {
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
__host__ __device__ constexpr uint32_t kmajor128_desc_hi(uint32_t sbo_bytes) { return (sbo_bytes >> 4) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t kmajor128_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFF) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_bf16_uniform(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                  uint32_t idesc, uint32_t acc, uint32_t leader) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t gets row (lane) t, v[j] = column j
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major swizzled operand: rows of SWB bytes, 8-row groups SBO = 8*SWB apart (mma_sm100_desc.hpp SmemDescriptor)
template <int SWB>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr) {
  constexpr uint64_t layout = SWB == 128 ? 2 : (SWB == 64 ? 4 : 6);   // SWIZZLE_128B / 64B / 32B
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                 // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * SWB) >> 4) << 32;                   // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
  d |= layout << 61;                                       // layout type, bits [61,64)
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);                       // F32 accum, BF16 A/B
}


// EPI_DGRAD_BNR (persistent halo kernels only): dgrad whose destination A is the output of a BatchNormalization -- `mask`
// points at the BN's INPUT tensor and the epilogue also accumulates the BN backward sums (sum dy, sum dy*xhat) of the
// gradient it stores, so the separate bn_bwd_reduce pass over (x, dy) disappears
enum { EPI_FPROP = 0, EPI_DGRAD = 1, EPI_TCONV = 2, EPI_DGRAD_BNR = 3 };

struct UArgs {
  int taps, ktap;            // filter taps; ktap = sqrt(taps)
  int sx, offbase;           // A coordinate = pixel*sx + (tap offset) + offbase   (conv: sx 1, offbase -pad)
  int c_a, c_b;              // channels of input A / B (K extents)
  int H, W;                  // pixel grid of the A tile space (= output grid except for ConvT fprop)
  int tiles_x, tiles_y;
  int epi, act;
  float alpha;
  const float* bias;
  const float* bias9;        // 3x3 fprop with a BatchNorm folded into the input: [9][n_total] bias per border class, or NULL
  // outputs: channel-slice views (bf16).  ya: columns [0, split); yb: columns [split, N)
  __nv_bfloat16* ya; long long ya_cs; int split;
  __nv_bfloat16* yb; long long yb_cs;
  const __nv_bfloat16* mask; long long mask_cs;   // dgrad: post-activation output of the layer that produced dx
  int n_total;               // total GEMM N
  int cout_t;                // ConvT fprop: channels per tap
  int nimg;                  // images (persistent kernels)
  int dbg;                   // experiment switch (descriptor variants)
  double* stats;             // fprop (persistent halo kernel): per-channel sum | sum of squares of the STORED output, or NULL
                             // EPI_DGRAD_BNR: [2*split] sum dy | sum dy*xhat of the BatchNorm whose output gradient is destination A
  const float* bnr_mi;       // EPI_DGRAD_BNR: that BatchNorm's [2*split] mean | invstd
};

// epilogue of one 32-(or 16-)column chunk of one accumulator row: bias+act (fprop / ConvT) or act'(mask) (dgrad),
// bf16 conversion, 16-byte stores into the (possibly channel-sliced) NHWC destination
// destination / mask addressing of one accumulator-row chunk
struct ChunkAddr {
  __nv_bfloat16* dst;
  const __nv_bfloat16* msk;   // nullptr when no mask applies to this chunk
  int bch;                    // bias channel
};
// EPI >= 0: the epilogue kind is a compile-time constant (persistent kernels); -1: read a.epi
template <int EPI = -1>
__device__ __forceinline__ ChunkAddr chunk_addr(const UArgs& a, int n, int gy, int gx, int ncol) {
  ChunkAddr r;
  long long pix;
  int ch;
  const int epi = EPI >= 0 ? EPI : a.epi;
  if (epi == EPI_TCONV) {
    const int tap = ncol / a.cout_t;
    ch = ncol - tap * a.cout_t;
    pix = ((long long)n * (2 * a.H) + 2 * gy + (tap >> 1)) * (2 * a.W) + 2 * gx + (tap & 1);
    r.dst = a.ya + pix * a.ya_cs + ch;
    r.bch = ch;
  } else {
    pix = ((long long)n * a.H + gy) * a.W + gx;
    if (ncol < a.split) { ch = ncol; r.dst = a.ya + pix * a.ya_cs + ch; }
    else { ch = ncol - a.split; r.dst = a.yb + pix * a.yb_cs + ch; }
    r.bch = ncol;
  }
  r.msk = (epi == EPI_DGRAD && a.mask && ncol < a.split) ? a.mask + pix * a.mask_cs + ch : nullptr;
  return r;
}

__device__ __forceinline__ float act_slope(int act, float alpha) {
  return act == DNNCA_ACT_RELU ? 0.f : (act == DNNCA_ACT_LEAKY ? alpha : 1.f);
}
__device__ __forceinline__ float act_slope_apply(float v, float slope) { return fmaf(slope, fminf(v, 0.f), fmaxf(v, 0.f)); }

// tile index -> (tile column, tile row, image) of a persistent CTA walking tiles blockIdx.x, +gridDim.x, ...: the two
// divisions run once, every step is adds and compares (a division by a run-time value costs ~100 cycles)
struct TileWalk {
  int tix, tiy, n;
  int sx, sy, sn, tiles_x, tiles_y;
  __device__ __forceinline__ TileWalk(int first, int step, int tx_, int ty_) : tiles_x(tx_), tiles_y(ty_) {
    tix = first % tx_; int r = first / tx_; tiy = r % ty_; n = r / ty_;
    sx = step % tx_; r = step / tx_; sy = r % ty_; sn = r / ty_;
  }
  __device__ __forceinline__ void next() {
    tix += sx;
    if (tix >= tiles_x) { tix -= tiles_x; ++tiy; }
    tiy += sy;
    if (tiy >= tiles_y) { tiy -= tiles_y; ++n; }
    n += sn;
  }
};

// persistent-kernel epilogue of a 32-column chunk: `m` holds the mask row prefetched before the accumulator was ready,
// `sbias` is the CTA's bias slice in shared memory
// after the call v[0] = sum over the 32 lanes of column `lane` (31 shuffles: lanes trade halves of the column set)
__device__ __forceinline__ void warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
}

// STATS: on return v[] holds the values as they read back from the bf16 tensor (fp32 bit patterns)
// `nvalid`: columns of this chunk below N (a multiple of 8): 8-column groups at or beyond it belong to the zero-padded
// tail of a partially filled N tile and are not stored
template <bool STATS = false, int EPI = -1>
__device__ __forceinline__ void epilogue_chunk32(const UArgs& a, uint32_t (&v)[32], const ChunkAddr& ca, const uint4 (&m)[4],
                                                 const float* sbias, bool store = true, int nvalid = 32) {
  const int epi = EPI >= 0 ? EPI : a.epi;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  // branch-free activation: act(v) = max(v,0) + slope*min(v,0), act'(y) = y > 0 ? 1 : slope with slope = 0 (ReLU),
  // alpha (LeakyReLU), 1 (none) -- a per-element switch on a.act compiles to two uniform branches per element
  const float slope = act_slope(a.act, a.alpha);
  if (epi == EPI_DGRAD) {
    if (ca.msk) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t mw[4] = {m[q].x, m[q].y, m[q].z, m[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[q * 8 + 2 * j] *= __uint_as_float(mw[j] << 16) > 0.f ? 1.f : slope;
          f[q * 8 + 2 * j + 1] *= __uint_as_float(mw[j] & 0xffff0000u) > 0.f ? 1.f : slope;
        }
      }
    }
  } else {
    const float4* sb4 = reinterpret_cast<const float4*>(sbias);      // 128-byte aligned: chunks start at multiples of 32
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b4 = sb4[j];
      f[4 * j + 0] = act_slope_apply(f[4 * j + 0] + b4.x, slope);
      f[4 * j + 1] = act_slope_apply(f[4 * j + 1] + b4.y, slope);
      f[4 * j + 2] = act_slope_apply(f[4 * j + 2] + b4.z, slope);
      f[4 * j + 3] = act_slope_apply(f[4 * j + 3] + b4.w, slope);
    }
  }
  uint4* d4 = reinterpret_cast<uint4*>(ca.dst);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[q * 8 + 0], f[q * 8 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(f[q * 8 + 2], f[q * 8 + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[q * 8 + 4], f[q * 8 + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(f[q * 8 + 6], f[q * 8 + 7]);
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2);
    o.w = *reinterpret_cast<uint32_t*>(&p3);
    if (store && q * 8 < nvalid) d4[q] = o;   // (a streaming st.global.cs here measured 3 % slower: the consumer kernel
                                              // that follows finds part of this output in L2)
    if (STATS) {                              // the values as they read back from the bf16 tensor
      const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[q * 8 + 2 * j] = ow[j] << 16;
        v[q * 8 + 2 * j + 1] = ow[j] & 0xffff0000u;
      }
    }
  }
}

// `nvalid`: GEMM columns of this chunk below N (a multiple of 8; >= CH for full chunks): columns at or beyond it belong
// to the zero-padded part of a covering N tile and are neither biased nor stored
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const UArgs& a, const uint32_t (&v)[32], int n, int gy, int gx, int ncol,
                                               int nvalid = CH) {
  __nv_bfloat16* dst;
  int ch;
  long long pix;
  if (a.epi == EPI_TCONV) {
    const int tap = ncol / a.cout_t;
    ch = ncol - tap * a.cout_t;
    pix = ((long long)n * (2 * a.H) + 2 * gy + (tap >> 1)) * (2 * a.W) + 2 * gx + (tap & 1);
    dst = a.ya + pix * a.ya_cs + ch;
  } else {
    pix = ((long long)n * a.H + gy) * a.W + gx;
    if (ncol < a.split) { ch = ncol; dst = a.ya + pix * a.ya_cs + ch; }
    else { ch = ncol - a.split; dst = a.yb + pix * a.yb_cs + ch; }
  }
  float f[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) f[j] = __uint_as_float(v[j]);
  const float slope = act_slope(a.act, a.alpha);
  if (a.epi == EPI_DGRAD) {
    if (a.mask && ncol < a.split) {
      const uint4* mp = reinterpret_cast<const uint4*>(a.mask + pix * a.mask_cs + ch);
#pragma unroll
      for (int q = 0; q < CH / 8; ++q) {
        const uint4 m = mp[q];
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          f[q * 8 + 2 * j] *= __uint_as_float(mw[j] << 16) > 0.f ? 1.f : slope;
          f[q * 8 + 2 * j + 1] *= __uint_as_float(mw[j] & 0xffff0000u) > 0.f ? 1.f : slope;
        }
      }
    }
  } else {
    const int bch = a.epi == EPI_TCONV ? ch : ncol;
    if (a.bias) {
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = act_slope_apply(f[j] + (j < nvalid ? __ldg(a.bias + bch + j) : 0.f), slope);
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j) f[j] = act_slope_apply(f[j], slope);
    }
  }
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int q = 0; q < CH / 8; ++q) {
    uint4 o;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[q * 8 + 0], f[q * 8 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(f[q * 8 + 2], f[q * 8 + 3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[q * 8 + 4], f[q * 8 + 5]);
    __nv_bfloat162 p3 = __floats2bfloat162_rn(f[q * 8 + 6], f[q * 8 + 7]);
    o.x = *reinterpret_cast<uint32_t*>(&p0);
    o.y = *reinterpret_cast<uint32_t*>(&p1);
    o.z = *reinterpret_cast<uint32_t*>(&p2);
    o.w = *reinterpret_cast<uint32_t*>(&p3);
    if (q * 8 < nvalid) d4[q] = o;
  }
}

}  // namespace dnnca
