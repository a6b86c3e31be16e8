// tcgen05 "row-Toeplitz" implicit GEMM: 3x3 convolutions whose tensors have FEW channels (configs/unet.yaml: 3/6/12,
// SURVEY D4) on the tensor cores.  Replaces layers.Conv2D at components.py:47-50 / :123-126 for those shapes.
//
// Why: a 6-byte NHWC pixel cannot be a K-major UMMA operand row (16-byte granularity), so conv_umma*.cu cannot take these
// layers and conv_small.cu runs them on the FP32 pipe, where they are issue-bound at 20-25 TFLOP/s (12-20 % of the HBM
// roofline, profiles/r01_bench.json).  Here the GEMM is laid along image ROWS instead of pixels:
//
//   M = 128 image rows y,   K = a window of (P+2 pixels) x C contiguous elements of input row y+dy,
//   N = P output pixels x Cout,   B_dy[k][n] = banded (block-Toeplitz) expansion of w[dy][.][.][.]
//
//   out[y, x0+p, co] = sum_dy  A[y+dy, window(x0)] . B_dy[:, (p,co)]
//
// In NHWC a row segment of P+2 pixels IS contiguous, so one TMA box {64 elements, Hs+2 rows} of the tensor viewed as
// {W*C, H, N} lands as 128-byte rows in SWIZZLE_128B K-major form, and the three dy taps are the SAME buffer read
// through descriptors whose start address is shifted by one 128-byte row (the swizzle is address-based, see
// conv_umma2.cu).  The TMA's out-of-bounds zero fill is the 'same' padding in both directions.  Only 3 of every P+2
// window pixels carry non-zero weights, so ~15-25 % of the MMA flops are useful -- which still leaves every layer of
// unet.yaml far from the tensor roofline; a tile is paced by the MMA time of these small-N shapes (~8 + 1.1*N cycles per
// K=16 step measured) and by the TMEM read of the epilogue (64 B/clk/SM), both close to the tile's HBM time.
//
// Weight precision: fprop expands every band twice, as hi = bf16(w) and lo = bf16(w - hi) (the activations are exact
// bf16 already), side by side along N (columns [0,N) and [N,2N) of the same MMA; the epilogue adds them), so the result
// equals the FP32-pipe kernels' up to accumulation order; that doubles the tensor and TMEM-read time.  dgrad uses a single
// bf16 band (DNNCA_ROW_EXACT_DGRAD=1: hi|lo as well); DNNCA_ROW_BF16_WEIGHTS=1 issues a single band everywhere (weights
// rounded to bf16, like conv_umma*.cu).  A mixed bf16 x fp16 MMA (fp16 bands) is not an option: kind::f16 with different
// A and B formats is an illegal instruction on sm_100a.
//
// Kernels: conv_row_umma_kernel (fprop and dgrad: dgrad is the same GEMM over dz with rot180/transposed bands and an
// act'(mask) epilogue) and conv_row_wgrad_kernel (M = window elements, N = P x Cout, K = image rows; both operands
// MN-major; per-CTA accumulation in TMEM over all of its tiles, band extraction once at the end).
#include <stdlib.h>
#include <string.h>

#include <utility>

#include "umma_common.cuh"

namespace dnnca {

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// programmatic dependent launch: the kernel may start while its predecessor in the stream is still draining; everything
// that reads the predecessor's output (the TMA loads of activations / gradients) sits behind pdl_wait()
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
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
// named barrier of one epilogue group (256 threads); ids 1 and 2
__device__ __forceinline__ void group_bar_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// SWIZZLE_128B shared-memory descriptor split into its two 32-bit words.  The high word (SBO = 1024 bytes between
// 8-row groups, descriptor version, layout type) is the same for every operand of these kernels; the low word holds
// the start address (>>4) and the leading byte offset (>>4; MN-major: bytes between 64-element blocks; K-major: unused).
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
// loads / stores of a row tile: {row elements, rows, image} maps, or {row elements, image in group, rows, image group}
// when ipt > 1 images are interleaved row by row in one tile (see row_common_geometry)
__device__ __forceinline__ void row_tma_load(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int n, int ipt) {
  if (ipt > 1) tma_load_4d(dst, map, bar, x, 0, y, n);
  else tma_load_3d(dst, map, bar, x, y, n);
}
__device__ __forceinline__ void row_tma_store(const CUtensorMap* map, uint32_t src, int x, int y, int n, int ipt) {
  if (ipt > 1)
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(x), "r"(0), "r"(y), "r"(n)
                 : "memory");
  else
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(map)),
                 "r"(src), "r"(x), "r"(y), "r"(n)
                 : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr & 0x3FFFF) >> 4) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
// issued only when `leader` != 0: no C++ branch around the MMA, so an unrolled issue loop stays straight-line while the
// whole warp runs it (operands then live in uniform registers)
__device__ __forceinline__ void umma_lo_if(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc,
                                           uint32_t leader) {
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
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI), "r"(leader)
      : "memory");
}

constexpr int ROW_MAX_UNITS = 12;     // (operand, row tap, window atom) band blocks
constexpr int ROW_MAX_MMA = 48;
constexpr int ROW_MAX_ACC = 8;        // wgrad accumulators

struct RowUnit {
  unsigned char op, tap, atom, block, kk0, nk, pad0, pad1;   // band block: smem block, first 32-byte slot, K steps
};
// wgrad: one accumulator = one MMA per K step; its M = 128 rows are two 64-element blocks (or one two-atom window)
enum { WB_NONE = 0, WB_WINDOW = 1, WB_ONES = 2, WB_WINDOW128 = 3 };
struct RowAcc {
  unsigned char kind[2], op[2], tap[2], m64, zi;     // zi: which dZ tap buffer (ConvT wgrad: rows 2i+zi)
  int first, second;             // byte offsets from the stage start; -1 = the all-ones block (second unused when m64)
};

struct RowArgs {
  int nops;                       // input operands (2 = the two producers of a virtual concat)
  int C[2], halo[2], atoms[2], ksteps[2], coff[2];
  int P, N, nsplit;               // pixels per strip; GEMM N = P*(oa+ob); columns [0,nsplit) go to destination a
  int parts, NP;                  // hi/lo weight parts; MMA N = parts * N
  int oa, ob;                     // channels of destination a / b
  int Hs, tiles_x, tiles_y, nimg; // rows per tile (<= 128), strips per row, row tiles per image, image groups
  int obufs;                      // output staging buffers per epilogue group: 2 (a TMA store may still read one while the
                                  // next tile is written) or 1 when shared memory is short (wide N: 1->16 first layers)
  int ipt_max;                    // host search: 1 = do not interleave images
  int ipt, hrows;                 // images interleaved in one tile (tile row v = image row v / ipt of image v % ipt); Hs / ipt
  int abuf, nstage;               // bytes of one window-atom buffer; ring depth
  // Conv2DTranspose k=s=2 variants (kernel [2][2][Cout][Cin], components.py:118-120):
  //   tconv 1 = fprop: window = P input pixels (no halo, one tap); columns = (a, 2P output pixels, co), output rows 2i+a
  //   tconv 2 = dgrad: taps a = rows 2i+a of dy (loaded as separate buffers through a 4-D map), window = 2P dy pixels
  //   tconv 3 = wgrad: x window against the dy rows 2i+a
  int tconv, pad, ltaps, ntaps;   // pad = halo TILE rows above/below a tile (padr image rows x ipt); ltaps = separately loaded tap buffers
  int padr;
  int win_step[2];                // window start element = strip * win_step - halo
  int tap_stride;                 // bytes between the A operands of consecutive taps inside a stage
  int box_w, box_h;               // output TMA box (elements of destination a, rows)
  int dgrad, act, has_mask;
  float alpha;
  const float* w;                 // [3][3][cin_tot][cout] fp32 (HWIO)
  const float* bias;
  int cin_tot, cout;
  // fprop / dgrad: band blocks and the per-tile MMA list (offsets in 16-byte units: A from the stage, B from the blocks)
  int nunits, nblocks, nmma;
  RowUnit unit[ROW_MAX_UNITS];
  unsigned short mma_a[ROW_MAX_MMA], mma_b[ROW_MAX_MMA];
  // wgrad only
  int zatoms, zbuf, nacc, ztaps, nwt;   // nwt = weight-gradient elements (9*cin*cout or 4*cin*cout)
  RowAcc acc[ROW_MAX_ACC];
  float* dw;
  float* db;
  int dbg;                        // DNNCA_ROW_DBG experiment bits: 1 no TMA store, 2 no TMA loads after the first ring fill,
                                  // 4 no MMA issue, 8 no epilogue arithmetic, 16 print CTA 0's per-role clock trace
};

__host__ __device__ inline int row_stage_bytes(const RowArgs& a) {
  return a.ltaps * (a.atoms[0] + (a.nops > 1 ? a.atoms[1] : 0)) * a.abuf;
}
__host__ __device__ inline int row_out_bytes(const RowArgs& a) { return a.Hs * a.N * 2; }
__host__ __device__ inline int row_mask_bytes(const RowArgs& a) { return a.has_mask ? a.Hs * a.nsplit * 2 : 0; }
constexpr int ROW_CTRL = 256 + ROW_MAX_MMA * 8 + 1024 + 1024;      // barriers, MMA table, bias slice, debug trace + lut
__host__ __device__ inline int row_smem_bytes(const RowArgs& a) {
  return a.nstage * row_stage_bytes(a) + a.nblocks * a.NP * 128 + 2 * a.obufs * row_out_bytes(a) + 2 * row_mask_bytes(a) + ROW_CTRL + 1024;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

enum { REPI_NONE = 0, REPI_RELU = 1, REPI_LEAKY = 2, REPI_DGRAD = 3, REPI_DGRAD_RELU = 4, REPI_DGRAD_LEAKY = 5 };

// one 8-column chunk of one accumulator row -> 8 bf16 (16 bytes); sb / mrow are shared-space addresses (mrow 0 = no mask)
template <int EPI>
__device__ __forceinline__ uint4 row_epilogue8(float* f, uint32_t sb, uint32_t mrow, float alpha) {
  uint4 o;
  if (EPI <= REPI_LEAKY) {
    const uint4 b0 = lds128(sb), b1 = lds128(sb + 16);
    f[0] += __uint_as_float(b0.x); f[1] += __uint_as_float(b0.y); f[2] += __uint_as_float(b0.z); f[3] += __uint_as_float(b0.w);
    f[4] += __uint_as_float(b1.x); f[5] += __uint_as_float(b1.y); f[6] += __uint_as_float(b1.z); f[7] += __uint_as_float(b1.w);
    if (EPI == REPI_LEAKY) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = f[e] > 0.f ? f[e] : alpha * f[e];
    }
    if (EPI == REPI_RELU) {
      o.x = pack_relu_bf16x2(f[0], f[1]); o.y = pack_relu_bf16x2(f[2], f[3]);
      o.z = pack_relu_bf16x2(f[4], f[5]); o.w = pack_relu_bf16x2(f[6], f[7]);
    } else {
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    }
  } else {
    if (EPI != REPI_DGRAD && mrow) {
      const uint4 m = lds128(mrow);
      const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m0 = __uint_as_float(mw[e] << 16), m1 = __uint_as_float(mw[e] & 0xffff0000u);
        if (EPI == REPI_DGRAD_RELU) {
          f[2 * e] = m0 > 0.f ? f[2 * e] : 0.f;
          f[2 * e + 1] = m1 > 0.f ? f[2 * e + 1] : 0.f;
        } else {
          f[2 * e] = m0 > 0.f ? f[2 * e] : alpha * f[2 * e];
          f[2 * e + 1] = m1 > 0.f ? f[2 * e + 1] : alpha * f[2 * e + 1];
        }
      }
    }
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  }
  return o;
}

template <int EPI>
__global__ void __launch_bounds__(608) conv_row_umma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB,
                                                           const __grid_constant__ CUtensorMap mapOA,
                                                           const __grid_constant__ CUtensorMap mapOB,
                                                           const __grid_constant__ CUtensorMap mapM,
                                                           const __grid_constant__ RowArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = row_stage_bytes(a);
  const int bblk = a.NP * 128;
  unsigned char* aring = smem;                                   // first: UMMA over-reads of short tiles stay inside
  unsigned char* bbase = aring + a.nstage * stage_bytes;
  unsigned char* obase = bbase + a.nblocks * bblk;
  unsigned char* mbase = obase + 2 * a.obufs * row_out_bytes(a);
  unsigned char* ctrl = mbase + 2 * row_mask_bytes(a);
  uint64_t* fullA = reinterpret_cast<uint64_t*>(ctrl);          // [8]
  uint64_t* emptyA = fullA + 8;                                  // [8]
  uint64_t* tfull = emptyA + 8;                                  // [4] accumulator buffers (NBUF = 4 when they fit TMEM)
  uint64_t* tempty = tfull + 4;                                  // [4]
  uint64_t* mfull = tempty + 4;                                  // [2]
  uint64_t* mempty = mfull + 2;                                  // [2]
  uint64_t* bready = mempty + 2;                                 // [1] band blocks built (18 warps)
  uint64_t* turn = bready + 1;                                   // [2] hand-over between the two MMA-issuing warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(turn + 2);
  uint2* mma_tab = reinterpret_cast<uint2*>(ctrl + 256);         // [nmma] {A offset from the stage, B descriptor low word}
  float* sbias = reinterpret_cast<float*>(ctrl + 256 + ROW_MAX_MMA * 8);
  constexpr bool HAS_MASK = EPI == REPI_DGRAD_RELU || EPI == REPI_DGRAD_LEAKY;
  uint32_t* trace = reinterpret_cast<uint32_t*>(ctrl + 256 + ROW_MAX_MMA * 8 + 1024);   // [5][48] clock stamps (dbg & 16)
  unsigned short* ulut = reinterpret_cast<unsigned short*>(trace + 240);
  const bool tracing = (a.dbg & 16) && blockIdx.x == 0;
  const bool etrace = (a.dbg & 32) != 0;               // epilogue sub-steps of group 0 instead of the role trace
#define ROW_TRACE(role, i) do { if (tracing && !etrace && (i) < 48) trace[(role) * 48 + (i)] = (uint32_t)clock(); } while (0)
#define ROW_ETRACE(step, i) do { if (tracing && etrace && (i) < 48) trace[(step) * 48 + (i)] = (uint32_t)clock(); } while (0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  const int nbuf = 4 * a.NP <= 512 ? 4 : 2;          // accumulator buffers: tile it uses buffer it % nbuf
  uint32_t tcols = 32;
  while (tcols < (uint32_t)(nbuf * a.NP)) tcols <<= 1;
  const uint32_t t_entry = (uint32_t)clock();

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstage; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, 8); }
    for (int s = 0; s < 2; ++s) { mbar_init(mfull + s, 1); mbar_init(mempty + s, 8); }
    mbar_init(bready, 18);
    mbar_init(turn, 1); mbar_init(turn + 1, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tcols);
  for (int i = threadIdx.x; i < a.N; i += blockDim.x) sbias[i] = (!a.dgrad && a.bias) ? a.bias[i % a.cout] : 0.f;
  {
    const uint32_t b_lo0 = desc_lo(smem_u32(bbase), 16);
    for (int i = threadIdx.x; i < a.nmma; i += blockDim.x) mma_tab[i] = make_uint2(a.mma_a[i], b_lo0 + a.mma_b[i]);
  }
  // band blocks (NP rows x 128 bytes, SWIZZLE_128B K-major; rows [0,N) = hi part, [N,2N) = lo part): zero fill, then
  // scatter the 3*C non-zeros of every (operand, row tap, column).  ulut: [op][tap][atom] -> block | first slot << 8
  {
    uint4* z = reinterpret_cast<uint4*>(bbase);
    const int nz = a.nblocks * bblk / 16;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int u = threadIdx.x; u < a.nunits; u += blockDim.x) {
      const RowUnit un = a.unit[u];
      ulut[(un.op * 3 + un.tap) * 2 + un.atom] = (unsigned short)(un.block | (un.kk0 << 8));
    }
  }
  // the fp32 weights, staged once in the (still idle) output staging area: the scatter below then reads shared memory
  // instead of issuing one dependent global load per band element
  float* wsm = reinterpret_cast<float*>(obase);
  {
    const int nwt = (a.tconv ? 4 : 9) * a.cin_tot * a.cout;
    for (int i = threadIdx.x; i < nwt; i += blockDim.x) wsm[i] = __ldg(a.w + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_setup = (uint32_t)clock();
  // the TMA producer (warp 0) starts streaming tiles now; warps 1..18 expand the bands and meet at named barrier 3
  if (warp != 0) {
  auto put_band = [&](int op, int tap, int n, int k, float v) {
    const int atom = k >> 6, kin = k & 63;
    const int e = ulut[(op * 3 + tap) * 2 + atom];
    const int ch = (e >> 8) * 2 + (kin >> 3);
    unsigned char* blk = bbase + (e & 0xff) * bblk + (kin & 7) * 2;
    if (a.parts > 1) {                                      // bf16 hi | lo parts
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const int n2 = n + a.N;
      *reinterpret_cast<__nv_bfloat16*>(blk + n * 128 + ((ch ^ (n & 7)) << 4)) = hi;
      *reinterpret_cast<__nv_bfloat16*>(blk + n2 * 128 + ((ch ^ (n2 & 7)) << 4)) = __float2bfloat16_rn(v - __bfloat162float(hi));
    } else {                                                // single bf16 band
      *reinterpret_cast<__nv_bfloat16*>(blk + n * 128 + ((ch ^ (n & 7)) << 4)) = __float2bfloat16_rn(v);
    }
  };
  // one (operand, tap, column) per thread-iteration; its non-zeros are walked with running indices (no divisions
  // in the inner loops: they dominated the prologue of the 12-channel layers)
  const int bt = threadIdx.x - 32, nbt = blockDim.x - 32;
  if (a.tconv == 1) {
    // ConvT fprop: column n = (a, pp = 2p+b, co); non-zeros k = p*Cin + ci with K[a][b][co][ci]
    const int cin = a.C[0], nh = a.N >> 1;
    for (int n = bt; n < a.N; n += nbt) {
      const int ar = n / nh, rem = n - ar * nh;
      const int pp = rem / a.cout, co = rem - pp * a.cout;
      const float* wrow = wsm + ((ar * 2 + (pp & 1)) * a.cout + co) * cin;
      for (int ci = 0; ci < cin; ++ci) put_band(0, 0, n, (pp >> 1) * cin + ci, wrow[ci]);
    }
  } else if (a.tconv == 2) {
    // ConvT dgrad: tap = a; column n = (p, ci); non-zeros k = (2p+b)*Cout + co with K[a][b][co][ci]
    const int cz = a.C[0], cin = a.oa;
    for (int pi = bt; pi < 2 * a.N; pi += nbt) {
      const int ar = pi / a.N, n = pi - ar * a.N;
      const int p = n / cin, ci = n - p * cin;
      for (int b = 0; b < 2; ++b)
        for (int co = 0; co < cz; ++co)
          put_band(0, ar, n, (2 * p + b) * cz + co, wsm[((ar * 2 + b) * cz + co) * cin + ci]);
    }
  } else {
    // work item = (operand, tap, column, dx): C band elements each, so even the 12-channel layers keep every thread busy
    const int per_op = 9 * a.N;
    for (int wi = bt; wi < a.nops * per_op; wi += nbt) {
      const int op = wi >= per_op ? 1 : 0;
      const int r = wi - op * per_op;
      const int tap = r / (3 * a.N), r2 = r - tap * 3 * a.N;
      const int n = r2 / 3, dxi = r2 - n * 3;
      const int C = a.C[op];
      int p, widx, wstep_c;                                 // weight index = widx + c*wstep_c
      if (!a.dgrad) {
        p = n / a.cout;
        const int co = n - p * a.cout;
        widx = ((tap * 3 + dxi) * a.cin_tot + a.coff[op]) * a.cout + co;
        wstep_c = a.cout;
      } else {                                              // column = (p, ci) of dx / dx2; window channel = forward co
        int ci;
        if (n < a.nsplit) { p = n / a.oa; ci = n - p * a.oa; }
        else { const int m = n - a.nsplit; p = m / a.ob; ci = a.oa + m - p * a.ob; }
        widx = (((2 - tap) * 3 + (2 - dxi)) * a.cin_tot + ci) * a.cout;
        wstep_c = 1;
      }
      int k = a.halo[op] + (p + dxi - 1) * C;               // window element of pixel p + dx, channel 0
      for (int c = 0; c < C; ++c, ++k) put_band(op, tap, n, k, wsm[widx + c * wstep_c]);
    }
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) mbar_arrive(bready);
  asm volatile("bar.sync 3, 576;" ::: "memory");
  }

  pdl_launch_dependents();                 // the next kernel may be scheduled as soon as SMs free up
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      pdl_wait();                          // the producer of our inputs has completed and flushed
      const int natoms = a.atoms[0] + (a.nops > 1 ? a.atoms[1] : 0);
      const uint32_t abytes = (uint32_t)(a.ltaps * natoms * (a.Hs + 2 * a.pad) * 128);
      const uint32_t mbytes = (uint32_t)(a.Hs * a.nsplit * 2);
      int it = 0, s = 0;
      uint32_t ph = 1;                                       // parity of the "previous" phase: passes on a fresh barrier
      int tix = blockIdx.x % a.tiles_x, rest = blockIdx.x / a.tiles_x;
      const int step_x = gridDim.x % a.tiles_x, step_r = gridDim.x / a.tiles_x;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int tiy = rest % a.tiles_y, n = rest / a.tiles_y;
        const int y0 = tiy * a.hrows;
        mbar_wait(emptyA + s, ph);
        ROW_TRACE(0, it);
        if ((a.dbg & 2) && it >= a.nstage) {
          mbar_arrive(fullA + s);
        } else {
          mbar_expect_tx(fullA + s, abytes);
          unsigned char* dst = aring + s * stage_bytes;
          if (a.tconv == 2) {                                  // dy rows 2i+tap through the {row, parity, i, n} map
            for (int tap = 0; tap < 2; ++tap)
              for (int at = 0; at < a.atoms[0]; ++at, dst += a.abuf)
                tma_load_4d(dst, &mapA, fullA + s, tix * a.win_step[0] + at * 64, tap, y0, n);
          } else {
            for (int op = 0; op < a.nops; ++op)
              for (int at = 0; at < a.atoms[op]; ++at, dst += a.abuf)
                row_tma_load(dst, op ? &mapB : &mapA, fullA + s, tix * a.win_step[op] - a.halo[op] + at * 64, y0 - a.padr, n, a.ipt);
          }
        }
        if (HAS_MASK) {
          const int mb = it & 1;
          if (it == 0) mbar_wait(bready, 0);
          if (it >= 2) mbar_wait(mempty + mb, ((it >> 1) - 1) & 1);
          mbar_expect_tx(mfull + mb, mbytes);
          row_tma_load(mbase + mb * row_mask_bytes(a), &mapM, mfull + mb, tix * a.box_w, y0, n, a.ipt);
        }
        if (++s == a.nstage) { s = 0; ph ^= 1u; }
        tix += step_x; rest += step_r;
        if (tix >= a.tiles_x) { tix -= a.tiles_x; ++rest; }
      }
    }
  } else if (warp == 1 || warp == 18) {
    // ===== MMA issuers: two warps take alternate tiles (warp <-> accumulator buffer).  A warp takes its barrier waits
    // (accumulator drained, operands landed) while the other one is issuing, then waits for its TURN, so MMAs enter the
    // tensor pipe in tile order without the ~480-cycle bubble a single issuer leaves between tiles.  The per-tile MMA
    // list is a table; the whole warp runs the loop (addresses stay in uniform registers), one elected lane issues =====
    {
      const int mw = warp == 1 ? 0 : 1;
      const uint32_t leader = (elect_one() && !(a.dbg & 4)) ? 1u : 0u;
      const bool committer = elect_one();
      const uint32_t idesc = make_idesc(128, a.NP, 0, 0);
      const uint32_t a_ring_lo = desc_lo(smem_u32(aring), 16);
      const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
      const int nmma = a.nmma;
      int s = mw % a.nstage;
      uint32_t ph = (uint32_t)(mw / a.nstage) & 1u;
      uint32_t tph = mw ? 0u : 1u;                            // warp 0's first turn is free (fresh barrier, parity 1)
      for (int t = blockIdx.x + mw * gridDim.x, it = mw; t < ntiles; t += 2 * gridDim.x, it += 2) {
        const int buf = nbuf == 4 ? (it & 3) : (it & 1), use = nbuf == 4 ? (it >> 2) : (it >> 1);
        if (use >= 1) mbar_wait(tempty + buf, (use - 1) & 1);
        mbar_wait(fullA + s, ph);
        mbar_wait(turn + mw, tph);
        tph ^= 1u;
        tc_fence_after();
        if (lane == 0) ROW_TRACE(1, it);
        const uint32_t dtm = tmem_base + (uint32_t)(buf * a.NP);
        const uint32_t a_lo0 = a_ring_lo + (uint32_t)s * stage16;
        {
          const uint2 e = mma_tab[0];
          umma_lo_if(dtm, a_lo0 + e.x, e.y, idesc, 0u, leader);
        }
#pragma unroll 4
        for (int i = 1; i < nmma; ++i) {
          const uint2 e = mma_tab[i];
          umma_lo_if(dtm, a_lo0 + e.x, e.y, idesc, 1u, leader);
        }
        if (committer) {
          mbar_arrive(turn + (mw ^ 1));                       // the other warp may issue the next tile
          umma_commit(emptyA + s);
          umma_commit(tfull + buf);
        }
        __syncwarp();
        if (lane == 0) ROW_TRACE(2, it);
        s += 2;                                               // two tiles ahead in the ring
        while (s >= a.nstage) { s -= a.nstage; ph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: two groups of 8 warps take alternate tiles (group g <-> accumulator buffer g), so the TMEM-read
    // -> smem -> TMA-store latency chain of one tile overlaps the next tile's; each group double-buffers its staging
    // tile so a store may still be reading while the next tile is written.  A warp owns TMEM lanes 32*(warp%4).. =
    // image rows; the two warps of a lane quarter split the columns =====
    const int ew = warp - 2;                        // warps 2..17
    const int g = ew >> 3;
    const int lg = warp & 3;
    const int half = (ew & 7) >> 2;
    const int row = lg * 32 + lane;
    const bool warp_live = lg * 32 < a.Hs;
    const bool live = row < a.Hs;
    const bool issuer = threadIdx.x == 64 + g * 256;
    const int nchunk = a.N >> 4;                    // 8-column chunks per warp
    const int c_first = half * nchunk;
    const int csplit = a.nsplit >> 3;               // chunks [0, csplit) belong to destination a
    const int nb = a.N - a.nsplit;
    const int offA = row * a.nsplit * 2;                              // + chunk*16
    const int offB = a.Hs * a.nsplit * 2 + row * nb * 2 - a.nsplit * 2;   // + chunk*16
    const float alpha = a.alpha;
    const uint32_t out_s = smem_u32(obase) + (uint32_t)(a.obufs * g * row_out_bytes(a));
    const uint32_t msk_s = smem_u32(mbase) + (uint32_t)(g * row_mask_bytes(a) + offA);
    const uint32_t bias_s = smem_u32(sbias);
    const uint32_t tlane = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c_first * 8);
    const bool two = a.parts > 1;
    // tile coordinates advance incrementally (integer divisions by run-time values cost ~100 cycles each, serially)
    const int t0 = blockIdx.x + g * gridDim.x, tstep = 2 * gridDim.x;
    int tix = t0 % a.tiles_x, rest = t0 / a.tiles_x;
    int tiy = rest % a.tiles_y, n = rest / a.tiles_y;
    const int step_x = tstep % a.tiles_x, step_r = tstep / a.tiles_x;
    const int step_y = step_r % a.tiles_y, step_n = step_r / a.tiles_y;
    const uint32_t out_bytes = (uint32_t)row_out_bytes(a);
    const bool nbuf4 = nbuf == 4;
    int it = g, git = 0;
    for (int t = t0; t < ntiles; t += tstep, it += 2, ++git) {
      const uint32_t par = (uint32_t)(it >> 1) & 1u;                     // mask stage of this group
      const int buf = nbuf4 ? (it & 3) : (it & 1);
      const uint32_t tpar = (uint32_t)(nbuf4 ? (it >> 2) : (it >> 1)) & 1u;
      const uint32_t tbase = tlane + (uint32_t)(buf * a.NP);
      const uint32_t outb = out_s + (a.obufs > 1 ? (uint32_t)(git & 1) * out_bytes : 0u);
      if (issuer) {                                          // the store that last read THIS staging buffer is done
        if (a.obufs > 1) tma_store_wait_read1();
        else tma_store_wait_read0();
      }
      group_bar_sync(g);
      if (threadIdx.x == 64) ROW_ETRACE(0, it >> 1);
      if (HAS_MASK) mbar_wait(mfull + g, par);
      mbar_wait(tfull + buf, tpar);
      tc_fence_after();
      if (issuer) ROW_TRACE(3, it);
      if (threadIdx.x == 64) ROW_ETRACE(1, it >> 1);
      if (warp_live && !(a.dbg & 8)) {
        for (int c0 = 0; c0 < nchunk; c0 += 3) {
          uint32_t v[24], u[24];
          const int cnt = min(3, nchunk - c0);
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (j < cnt) {
              tmem_ld8(tbase + (uint32_t)((c0 + j) * 8), v + 8 * j);
              if (two) tmem_ld8(tbase + (uint32_t)(a.N + (c0 + j) * 8), u + 8 * j);
            }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j < cnt) {
              const int c = c_first + c0 + j;
              const bool to_a = c < csplit;
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * j + e]) + (two ? __uint_as_float(u[8 * j + e]) : 0.f);
              const uint4 o = row_epilogue8<EPI>(f, bias_s + c * 32, (HAS_MASK && to_a && live) ? msk_s + c * 16 : 0u, alpha);
              if (live) sts128(outb + (uint32_t)((to_a ? offA : offB) + c * 16), o);
            }
          }
        }
      }
      if (threadIdx.x == 64) ROW_ETRACE(2, it >> 1);
      tc_fence_before();
      fence_proxy_async();                                   // staging writes -> visible to the TMA store
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(tempty + buf);
        if (HAS_MASK) mbar_arrive(mempty + g);
      }
      if (threadIdx.x == 64) ROW_ETRACE(3, it >> 1);
      group_bar_sync(g);
      if (issuer) ROW_TRACE(4, it);
      if (threadIdx.x == 64) ROW_ETRACE(4, it >> 1);
      if (issuer && !(a.dbg & 1)) {
        row_tma_store(&mapOA, outb, tix * a.box_w, tiy * a.box_h, n, a.ipt);
        if (a.ob) row_tma_store(&mapOB, outb + (uint32_t)(a.Hs * a.nsplit * 2), tix * a.P * a.ob, tiy * a.box_h, n, a.ipt);
        tma_store_commit();
      }
      tix += step_x; tiy += step_y; n += step_n;
      if (tix >= a.tiles_x) { tix -= a.tiles_x; ++tiy; }
      if (tiy >= a.tiles_y) { tiy -= a.tiles_y; ++n; }
    }
    if (issuer) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tcols);
  if (tracing && threadIdx.x == 0) {
    const uint32_t t0 = trace[0];
    const uint32_t t_end = (uint32_t)clock();
    const int last = min(47, (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x);
    for (int i = 0; i < 24; ++i)
      printf("tile %2d  tma %7u  mma_start %7u  mma_done %7u  epi_start %7u  epi_done %7u\n", i, trace[i] - t0, trace[48 + i] - t0,
             trace[96 + i] - t0, trace[144 + i] - t0, trace[192 + i] - t0);
    if (!etrace)
      printf("entry %d setup %d first_tma 0 last tile(%d): mma_done %u epi_done %u end %u\n", (int)(t_entry - t0), (int)(t_setup - t0),
             last, trace[96 + last] - t0, trace[192 + last] - t0, t_end - t0);
  }
#undef ROW_TRACE
#undef ROW_ETRACE
}

// ---------------------------------------------------------------------------------------------------------------
// wgrad: per CTA, D[k][n] = sum over its tiles and rows of  X[row+tap-1][window k] * dZ[row][n]  in TMEM (M = window
// elements, N = P*Cout, K = rows; both operands MN-major).  The M = 128 rows of one MMA are two 64-element blocks LBO
// bytes apart, and LBO is free: one-atom windows are PAIRED -- (tap 0 | tap 1) is the same buffer one 128-byte row
// further, (tap 2 | all-ones block) yields the bias gradient sum(dZ) in rows 64.. for free -- so a K step of one operand
// costs 2 MMAs instead of 4.  Two-atom windows take M = 128 per tap plus an M = 64 all-ones MMA.  The 3-wide band of D is
// folded into dw[3][3][cin][cout] once at the end.
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int roww_stage_bytes(const RowArgs& a) { return row_stage_bytes(a) + a.ztaps * a.zatoms * a.zbuf; }
__host__ __device__ inline int roww_nw(const RowArgs& a) { return a.nwt + a.cout; }
constexpr int ROWW_CTRL = 512 + 1024 + 8 * ROW_MAX_ACC * 8;       // barriers, trace, per-stage descriptor table
__host__ __device__ inline int roww_smem_bytes(const RowArgs& a) {
  return a.nstage * roww_stage_bytes(a) + a.abuf + ((roww_nw(a) * 4 + 127) & ~127) + ROWW_CTRL + 1024;
}

__global__ void __launch_bounds__(192) conv_row_wgrad_kernel(const __grid_constant__ CUtensorMap mapA,
                                                            const __grid_constant__ CUtensorMap mapB,
                                                            const __grid_constant__ CUtensorMap mapZ,
                                                            const __grid_constant__ RowArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = roww_stage_bytes(a);
  const int xbytes = row_stage_bytes(a);
  unsigned char* ring = smem;
  unsigned char* ones = ring + a.nstage * stage_bytes;           // (Hs+2) rows x 128 B of bf16 1.0
  float* sW = reinterpret_cast<float*>(ones + a.abuf);
  unsigned char* ctrl = reinterpret_cast<unsigned char*>(sW) + ((roww_nw(a) * 4 + 127) & ~127);
  uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);            // [8]
  uint64_t* empty = full + 8;                                    // [8]
  uint64_t* accum = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  uint32_t* trace = reinterpret_cast<uint32_t*>(ctrl + 512);    // [3][48] clock stamps (dbg & 16)
  uint2* wtab = reinterpret_cast<uint2*>(ctrl + 512 + 1024);    // [stage][acc] {A descriptor low word at K step 0, idesc}
  uint32_t* zoff = reinterpret_cast<uint32_t*>(ctrl + 512 + 960);   // [acc] dZ tap-buffer offset, 16-byte units
  const bool tracing = (a.dbg & 16) && blockIdx.x == 0;
#define ROW_TRACE(role, i) do { if (tracing && (i) < 48) trace[(role) * 48 + (i)] = (uint32_t)clock(); } while (0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  uint32_t tcols = 32;
  while (tcols < (uint32_t)(a.nacc * a.N)) tcols <<= 1;
  const uint32_t t_entry = (uint32_t)clock();

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(accum, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tcols);
  for (int i = threadIdx.x; i < a.abuf / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  for (int i = threadIdx.x; i < roww_nw(a); i += blockDim.x) sW[i] = 0.f;
  for (int i = threadIdx.x; i < a.nstage * a.nacc; i += blockDim.x) {
    const int s = i / a.nacc, j = i - s * a.nacc;
    const RowAcc ac = a.acc[j];
    const uint32_t stage = smem_u32(ring + s * stage_bytes), ones_addr = smem_u32(ones);
    const uint32_t first = ac.first < 0 ? ones_addr : stage + (uint32_t)ac.first;
    const uint32_t second = ac.m64 ? first + 1024 : (ac.second < 0 ? ones_addr : stage + (uint32_t)ac.second);
    wtab[s * ROW_MAX_ACC + j] = make_uint2(desc_lo(first, second - first), make_idesc(ac.m64 ? 64 : 128, a.N, 1, 1));
    if (s == 0) zoff[j] = (uint32_t)(ac.zi * a.zatoms * a.zbuf) >> 4;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_setup = (uint32_t)clock();
  uint32_t t_acc = 0, t_ext = 0;

  pdl_launch_dependents();
  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&mapA);
      tma_prefetch_desc(&mapZ);
      pdl_wait();
      const int natoms = a.atoms[0] + (a.nops > 1 ? a.atoms[1] : 0);
      const uint32_t bytes = (uint32_t)(natoms * (a.Hs + 2 * a.pad) * 128 + a.ztaps * a.zatoms * a.Hs * 128);
      int it = 0, s = 0;
      uint32_t ph = 1;
      int tix = blockIdx.x % a.tiles_x, rest = blockIdx.x / a.tiles_x;
      const int step_x = gridDim.x % a.tiles_x, step_r = gridDim.x / a.tiles_x;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int tiy = rest % a.tiles_y, n = rest / a.tiles_y;
        const int y0 = tiy * a.hrows;
        mbar_wait(empty + s, ph);
        ROW_TRACE(0, it);
        mbar_expect_tx(full + s, bytes);
        unsigned char* dst = ring + s * stage_bytes;
        for (int op = 0; op < a.nops; ++op)
          for (int at = 0; at < a.atoms[op]; ++at, dst += a.abuf)
            row_tma_load(dst, op ? &mapB : &mapA, full + s, tix * a.win_step[op] - a.halo[op] + at * 64, y0 - a.padr, n, a.ipt);
        if (a.tconv) {                                        // dy rows 2i+tap
          for (int tap = 0; tap < 2; ++tap)
            for (int z = 0; z < a.zatoms; ++z, dst += a.zbuf) tma_load_4d(dst, &mapZ, full + s, tix * a.N + z * 64, tap, y0, n);
        } else {
          for (int z = 0; z < a.zatoms; ++z, dst += a.zbuf) row_tma_load(dst, &mapZ, full + s, tix * a.N + z * 64, y0, n, a.ipt);
        }
        if (++s == a.nstage) { s = 0; ph ^= 1u; }
        tix += step_x; rest += step_r;
        if (tix >= a.tiles_x) { tix -= a.tiles_x; ++rest; }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t leader = (elect_one() && !(a.dbg & 4)) ? 1u : 0u;
      const bool committer = elect_one();
      const int ksteps = a.Hs >> 4;
      const uint32_t ring_addr = smem_u32(ring);
      const int nacc = a.nacc;
      int it = 0, s = 0;
      uint32_t ph = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        mbar_wait(full + s, ph);
        tc_fence_after();
        if (lane == 0) ROW_TRACE(1, it);
        const uint32_t z_lo0 = desc_lo(ring_addr + (uint32_t)(s * stage_bytes + xbytes), a.zbuf);
        const uint32_t accf = it ? 1u : 0u;
        uint32_t dcol = tmem_base;
        for (int j = 0; j < nacc; ++j, dcol += a.N) {
          const uint2 e = wtab[s * ROW_MAX_ACC + j];
          const uint32_t z_lo = z_lo0 + zoff[j];
          umma_lo_if(dcol, e.x, z_lo, e.y, accf, leader);
#pragma unroll 8
          for (int ks = 1; ks < ksteps; ++ks)                // 16 rows = 2048 bytes = 128 descriptor units
            umma_lo_if(dcol, e.x + ks * 128, z_lo + ks * 128, e.y, 1u, leader);
        }
        if (committer) umma_commit(empty + s);
        __syncwarp();
        if (lane == 0) ROW_TRACE(2, it);
        if (++s == a.nstage) { s = 0; ph ^= 1u; }
      }
      if (committer) umma_commit(accum);
    }
  } else {
    // ===== band extraction: warps 2..5, TMEM lane quarter = warp % 4; accumulator row = lane index.  Each accumulator
    // is dumped to shared memory (the operand ring is idle once the last MMA has retired) with a padded row stride,
    // then every gradient element is summed by ONE thread: no atomics, no divergent selects =====
    const int lg = warp & 3;
    const int et = threadIdx.x - 64;                          // 0..127
    mbar_wait(accum, 0);
    tc_fence_after();
    t_acc = (uint32_t)clock();
    const int k = lg * 32 + lane;
    float* dump = reinterpret_cast<float*>(ring);
    const int ld = a.N + 1;
    float* sdb = sW + a.nwt;
    uint32_t col = 0;
    for (int j = 0; j < a.nacc; ++j) {
      const RowAcc ac = a.acc[j];
      for (int c0 = 0; c0 < a.N; c0 += 8, col += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_base + ((uint32_t)(lg * 32) << 16) + col, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e) dump[k * ld + c0 + e] = __uint_as_float(v[e]);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const bool full128 = ac.kind[0] == WB_WINDOW128;
      const int nhalf = (full128 || ac.m64) ? 1 : 2;
      for (int h = 0; h < nhalf; ++h) {
        const int kind = ac.kind[h], op = ac.op[h], tap = ac.tap[h];
        const float* blk = dump + (h * 64) * ld;               // rows of this 64-element block (or the whole window)
        if (kind == WB_ONES) {
          for (int co = et; co < a.cout; co += 128) {
            float sum = 0.f;
            for (int n = co; n < a.N; n += a.cout) sum += blk[n];
            sdb[co] += sum;
          }
        } else if (a.tconv) {                                   // dk[tap][b][co][ci] = sum_p D[(p,ci)][(2p+b,co)]
          const int C = a.C[op], nout = 2 * a.cout * C;
          for (int o = et; o < nout; o += 128) {
            const int ci = o % C, r = o / C;
            const int co = r % a.cout, b = r / a.cout;
            float sum = 0.f;
            for (int p = 0; p < a.P; ++p) sum += blk[(p * C + ci) * ld + (2 * p + b) * a.cout + co];
            sW[((tap * 2 + b) * a.cout + co) * a.cin_tot + ci] = sum;
          }
        } else {                                                // dw[tap][dx][ci][co] = sum_p D[(p+dx,ci)][(p,co)]
          const int C = a.C[op], nout = 3 * C * a.cout;
          for (int o = et; o < nout; o += 128) {
            const int co = o % a.cout, r = o / a.cout;
            const int ci = r % C, dxi = r / C;
            const float* src = blk + (a.halo[op] + (dxi - 1) * C + ci) * ld + co;
            float sum = 0.f;
            for (int p = 0; p < a.P; ++p) sum += src[p * (C * ld + a.cout)];
            sW[((tap * 3 + dxi) * a.cin_tot + a.coff[op] + ci) * a.cout + co] = sum;
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  t_ext = (uint32_t)clock();
  tc_fence_before();
  __syncthreads();
  // flush: CTAs start at different offsets so that they do not all hit the same L2 lines at the same time
  const int nw = a.nwt;
  const int rot = (int)(((long long)blockIdx.x * nw) / gridDim.x);
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    int e = i + rot;
    if (e >= nw) e -= nw;
    atomicAdd(a.dw + e, sW[e]);
  }
  if (a.db)
    for (int i = threadIdx.x; i < a.cout; i += blockDim.x) atomicAdd(a.db + i, sW[nw + i]);
  if (warp == 1) tmem_dealloc(tmem_base, tcols);
  if (tracing && threadIdx.x == 64) {
    __threadfence();
    const uint32_t t_end = (uint32_t)clock();
    const uint32_t t0 = trace[0];
    for (int i = 0; i < 4; ++i)
      printf("tile %2d  tma %7u  mma_start %7u  mma_done %7u\n", i, trace[i] - t0, trace[48 + i] - t0, trace[96 + i] - t0);
    const int last = min(47, (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x);
    printf("entry %d setup %d first_tma 0 last_mma_issue(%d) %u accum_done %u extracted %u end %u\n", (int)(t_entry - t0),
           (int)(t_setup - t0), last, trace[96 + last] - t0, t_acc - t0, t_ext - t0, t_end - t0);
  }
#undef ROW_TRACE
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static bool row_dense_bf16(const dnnca_tensor_t* t) {
  return t->dtype == DNNCA_BF16 && t->coff == 0 && t->cstride == t->c && ((long long)t->w * t->c * 2) % 16 == 0 &&
         (reinterpret_cast<uintptr_t>(t->data) & 15) == 0;
}

// tensor viewed as {W*C, H, N}; box {box_e, box_rows, 1}.  ipt > 1: {W*C, ipt, H, N/ipt} with box {box_e, ipt, box_rows, 1},
// which lands (and leaves) as rows ordered (image row, image in group): the images of a group interleaved row by row
static bool row_map(CUtensorMap* m, const dnnca_tensor_t* t, int box_e, int box_rows, bool swz, int ipt = 1) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc || box_e > 256 || box_rows > 256 || (box_e * 2) % 16) return false;
  const cuuint64_t rowe = (cuuint64_t)t->w * t->c;
  if (ipt > 1) {
    if (t->n % ipt) return false;
    cuuint64_t dims[4] = {rowe, (cuuint64_t)ipt, (cuuint64_t)t->h, (cuuint64_t)(t->n / ipt)};
    cuuint64_t strides[3] = {rowe * 2 * t->h, rowe * 2, rowe * 2 * t->h * ipt};
    cuuint32_t box[4] = {(cuuint32_t)box_e, (cuuint32_t)ipt, (cuuint32_t)box_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  cuuint64_t dims[3] = {rowe, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[2] = {rowe * 2, rowe * 2 * t->h};
  cuuint32_t box[3] = {(cuuint32_t)box_e, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, t->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static const int ROW_SMEM_LIMIT = 226 * 1024;
static const int ROW_MAX_C = 16;

static bool row_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("DNNCA_DISABLE_ROW_UMMA") ? 0 : 1;
  return on == 1;
}
// default: bf16 hi|lo bands along N; DNNCA_ROW_BF16_WEIGHTS=1: one bf16 band (half the MMA N)
static int row_parts() {
  static int parts = 0;
  if (!parts) parts = getenv("DNNCA_ROW_BF16_WEIGHTS") ? 1 : 2;
  return parts;
}
// measured tcgen05 time of one K=16 step at M=128 with these operand layouts (cycles)
static double row_mma_cycles(int n) { return 8.0 + 1.1 * n; }
static int row_dbg() {
  const char* e = getenv("DNNCA_ROW_DBG");
  return e ? atoi(e) : 0;
}

// window geometry of one input operand for P pixels per strip; false when it does not fit two 64-element atoms
static bool row_operand(RowArgs& a, int op, int C, int P) {
  if ((P * C) % 8) return false;
  // conv: P pixels + one halo pixel each side, start aligned down to 16 bytes; ConvT: P (fprop, wgrad) or 2P (dgrad)
  // pixels, no halo
  const int halo = a.tconv ? 0 : (C + 7) / 8 * 8;
  const int need = a.tconv ? (a.tconv == 2 ? 2 * P * C : P * C) : halo + P * C + C;
  if (need > 128) return false;
  a.C[op] = C; a.halo[op] = halo; a.ksteps[op] = (need + 15) / 16; a.atoms[op] = (a.ksteps[op] + 3) / 4;
  a.win_step[op] = a.tconv == 2 ? 2 * P * C : P * C;
  return true;
}

static bool row_common_geometry(RowArgs& a, const dnnca_tensor_t* x) {
  const int H = x->h, W = x->w;
  if (W % a.P) return false;
  a.hrows = 0;
  for (int d = 128; d >= 16; d -= 16)        // tallest row tile that divides the image
    if (H % d == 0) { a.hrows = d; break; }
  if (!a.hrows) return false;
  // images of 64 rows or fewer would leave half (or more) of the 128 MMA rows empty at the same tensor time: interleave
  // ipt images row by row in one tile (tile row v = row v / ipt of image v % ipt), each with its own zero halo rows, so
  // the dy taps stay a plain shift by ipt tile rows.  Conv only (the ConvT maps already spend the fourth dimension).
  a.ipt = 1;
  static int no_ipt = -1;
  if (no_ipt < 0) no_ipt = getenv("DNNCA_ROW_NO_INTERLEAVE") ? 1 : 0;
  if (!a.tconv && !no_ipt)
    while (a.ipt < a.ipt_max && 2 * a.ipt * a.hrows <= 128 && x->n % (2 * a.ipt) == 0) a.ipt *= 2;
  a.Hs = a.hrows * a.ipt;
  a.tiles_x = W / a.P; a.tiles_y = H / a.hrows; a.nimg = x->n / a.ipt;
  if ((long long)a.tiles_x * a.tiles_y * a.nimg > 0x7fffffffLL) return false;
  a.padr = a.tconv ? 0 : 1;                  // halo image rows above / below
  a.pad = a.padr * a.ipt;                    // the same in tile rows
  a.abuf = ((a.Hs + 2 * a.pad) * 128 + 1023) & ~1023;
  return true;
}

// band blocks and the per-tile MMA list.  A band unit with a single K step (the 16-element tail of a two-atom window)
// shares a 128-byte-row block with up to three others, one 32-byte slot each.
static bool row_plan_mmas(RowArgs& a) {
  const int bblk16 = a.NP * 128 / 16;
  int nblocks = 0, nunits = 0, nmma = 0;
  int narrow_block = -1, narrow_used = 4;
  int a_off = 0;                                        // operand offset inside the stage, 16-byte units
  a.ntaps = a.tconv == 1 ? 1 : (a.tconv == 2 ? 2 : 3);
  a.ltaps = a.tconv == 2 ? 2 : 1;
  a.tap_stride = a.tconv == 2 ? a.atoms[0] * a.abuf : 128 * a.ipt;     // conv: the next image row of the same buffer
  for (int op = 0; op < a.nops; ++op) {
    for (int tap = 0; tap < a.ntaps; ++tap)
      for (int at = 0; at < a.atoms[op]; ++at) {
        const int nk = a.ksteps[op] - 4 * at < 4 ? a.ksteps[op] - 4 * at : 4;
        if (nunits >= ROW_MAX_UNITS) return false;
        RowUnit& u = a.unit[nunits++];
        u.op = (unsigned char)op; u.tap = (unsigned char)tap; u.atom = (unsigned char)at;
        u.nk = (unsigned char)nk; u.pad0 = u.pad1 = 0;
        if (nk == 1) {
          if (narrow_used == 4) { narrow_block = nblocks++; narrow_used = 0; }
          u.block = (unsigned char)narrow_block; u.kk0 = (unsigned char)narrow_used++;
        } else {
          u.block = (unsigned char)nblocks++; u.kk0 = 0;
        }
        for (int kk = 0; kk < nk; ++kk) {
          if (nmma >= ROW_MAX_MMA) return false;
          const int ao = a_off + (at * a.abuf + tap * a.tap_stride + kk * 32) / 16;
          const int bo = u.block * bblk16 + (u.kk0 + kk) * 2;
          if (ao > 0xFFFF || bo > 0xFFFF) return false;
          a.mma_a[nmma] = (unsigned short)ao; a.mma_b[nmma] = (unsigned short)bo;
          ++nmma;
        }
      }
    a_off += a.ltaps * a.atoms[op] * a.abuf / 16;
  }
  a.nblocks = nblocks; a.nunits = nunits; a.nmma = nmma;
  return true;
}

// launch with the programmatic-stream-serialization attribute (DNNCA_NO_PDL=1: plain launch)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, int smem, cudaStream_t s, Args&&... args) {
  static int pdl = -1;
  if (pdl < 0) pdl = getenv("DNNCA_NO_PDL") ? 0 : 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

static int persistent_grid(int ntiles) {
  int g = sm_count();
  return g < ntiles ? g : ntiles;
}

template <int EPI>
static int launch_row_epi(cudaStream_t s, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mOA,
                          const CUtensorMap& mOB, const CUtensorMap& mM, const RowArgs& a) {
  const int smem = row_smem_bytes(a);
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_row_umma_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_row_umma: cudaFuncSetAttribute");
    smem_set = smem;
  }
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  {
    cudaError_t e = launch_pdl(conv_row_umma_kernel<EPI>, persistent_grid(ntiles), 608, smem, s, mA, mB, mOA, mOB, mM, a);
    if (e != cudaSuccess) return cuda_fail(e, "conv_row_umma: launch");
  }
  DNNCA_LAUNCH_CHECK("conv_row_umma");
  note_family(2);
  return 1;
}

// 4-D view of a [n, 2h, 2w, c] tensor as {row elements, row parity, h, n}: box {64, 1, rows, 1} = rows 2i+parity
static bool row_map_parity(CUtensorMap* m, const dnnca_tensor_t* t, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc || box_rows > 256 || (t->h & 1)) return false;
  const cuuint64_t rowe = (cuuint64_t)t->w * t->c;
  cuuint64_t dims[4] = {rowe, 2, (cuuint64_t)t->h / 2, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {rowe * 2, rowe * 4, rowe * 2 * t->h};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// conv fprop: inputs ina [, inb] -> outa ; conv dgrad: ina = dz -> outa = dx [, outb = dx2]
// tconv 1 (ConvT fprop): ina = x [n,h,w,cin] -> outa = y [n,2h,2w,cout] ; tconv 2 (ConvT dgrad): ina = dy -> outa = dx
static int launch_row(cudaStream_t s, bool dgrad, int tconv, const dnnca_tensor_t* ina, const dnnca_tensor_t* inb, const float* w,
                      const float* bias, const dnnca_tensor_t* outa, const dnnca_tensor_t* outb, const dnnca_tensor_t* mask,
                      int act, float alpha) {
  if (!row_enabled()) return 0;
  if (dgrad && (act == DNNCA_ACT_NONE)) mask = nullptr;
  if (!row_dense_bf16(ina) || (inb && !row_dense_bf16(inb)) || !row_dense_bf16(outa) || (outb && !row_dense_bf16(outb)) ||
      (mask && !row_dense_bf16(mask)))
    return 0;
  const int ca = ina->c, cb = inb ? inb->c : 0, oa = outa->c, ob = outb ? outb->c : 0;
  if (ca > ROW_MAX_C || cb > ROW_MAX_C || oa > ROW_MAX_C || ob > ROW_MAX_C) return 0;
  RowArgs a;
  memset(&a, 0, sizeof(a));
  a.nops = inb ? 2 : 1;
  a.tconv = tconv;
  a.dgrad = dgrad ? 1 : 0; a.act = act; a.alpha = alpha; a.has_mask = mask ? 1 : 0;
  a.w = w; a.bias = bias; a.oa = oa; a.ob = ob;
  a.dbg = row_dbg();
  // fprop keeps fp32-accurate weights (hi|lo bands: the logits bound of BASELINE.json is 1e-2); dgrad / ConvT dgrad use a
  // single bf16 band like every other tensor-core path here (gradient bound 2e-2; halves their MMA columns and TMEM
  // reads).  DNNCA_ROW_EXACT_DGRAD=1 restores hi|lo bands for the gradients.
  static int exact_dgrad = -1;
  if (exact_dgrad < 0) exact_dgrad = getenv("DNNCA_ROW_EXACT_DGRAD") ? 1 : 0;
  a.parts = (dgrad && !exact_dgrad) ? 1 : row_parts();
  if (dgrad) { a.cin_tot = oa + ob; a.cout = ca; }
  else { a.cin_tot = ca + cb; a.cout = oa; a.coff[0] = 0; a.coff[1] = ca; }
  const dnnca_tensor_t* grid_t = tconv == 2 ? outa : ina;        // tensor whose rows are the GEMM M
  // strip width P: cheapest tensor time per pixel (MMAs per tile x their N) among the widths whose plan fits
  bool found = false;
  RowArgs best;
  double best_cost = 0.0;
  // every strip width, with short images interleaved up to 8 per tile and without: cheapest tensor time per pixel
  for (int pass = 0; pass < 2; ++pass)
  for (int P = 16; P >= 2; P >>= 1) {
    a.ipt_max = pass ? 1 : 8;
    a.P = P;
    if (!row_operand(a, 0, ca, P) || (inb && !row_operand(a, 1, cb, P))) continue;
    if (tconv == 1) { a.N = 4 * P * oa; a.nsplit = a.N; a.box_w = 2 * P * oa; }
    else { a.N = P * (oa + ob); a.nsplit = P * oa; a.box_w = P * oa; }
    if (a.nsplit % 8 || (a.N - a.nsplit) % 8 || a.box_w % 8) continue;
    a.NP = a.parts * a.N;
    if (a.N % 16 || a.NP > 256 || a.N < 16) continue;
    if (!row_common_geometry(a, grid_t)) continue;
    a.box_h = tconv == 1 ? 2 * a.Hs : a.hrows;
    if (!row_plan_mmas(a)) continue;
    bool fits = false;
    for (a.obufs = 2; a.obufs >= 1 && !fits; --a.obufs)
      for (a.nstage = 4; a.nstage >= 2; --a.nstage)
        if (row_smem_bytes(a) <= ROW_SMEM_LIMIT) { fits = true; break; }
    if (!fits) continue;
    ++a.obufs;                                              // the loop's decrement after the fitting pass
    if ((tconv ? 4 : 9) * a.cin_tot * a.cout * 4 > 2 * a.obufs * row_out_bytes(a) + 2 * row_mask_bytes(a)) continue;   // weight staging
    const double cost = (a.nmma * row_mma_cycles(a.NP) + 300.0) / ((double)P * a.Hs);
    if (!found || cost < best_cost) { best = a; best_cost = cost; found = true; }
  }
  if (found) a = best;
  if (!found) return 0;
  CUtensorMap mA, mB, mOA, mOB, mM;
  if (tconv == 2) {
    if (!row_map_parity(&mA, ina, a.Hs)) return 0;
  } else if (!row_map(&mA, ina, 64, a.hrows + 2 * a.padr, true, a.ipt)) return 0;
  mB = mA;
  if (inb && !row_map(&mB, inb, 64, a.hrows + 2 * a.padr, true, a.ipt)) return 0;
  if (!row_map(&mOA, outa, a.box_w, a.box_h, false, a.ipt)) return 0;
  mOB = mOA;
  if (outb && !row_map(&mOB, outb, a.P * ob, a.hrows, false, a.ipt)) return 0;
  mM = mOA;
  if (mask && !row_map(&mM, mask, a.box_w, a.hrows, false, a.ipt)) return 0;
  if (!dgrad) {
    if (act == DNNCA_ACT_RELU) return launch_row_epi<REPI_RELU>(s, mA, mB, mOA, mOB, mM, a);
    if (act == DNNCA_ACT_LEAKY) return launch_row_epi<REPI_LEAKY>(s, mA, mB, mOA, mOB, mM, a);
    return launch_row_epi<REPI_NONE>(s, mA, mB, mOA, mOB, mM, a);
  }
  if (!mask) return launch_row_epi<REPI_DGRAD>(s, mA, mB, mOA, mOB, mM, a);
  if (act == DNNCA_ACT_RELU) return launch_row_epi<REPI_DGRAD_RELU>(s, mA, mB, mOA, mOB, mM, a);
  return launch_row_epi<REPI_DGRAD_LEAKY>(s, mA, mB, mOA, mOB, mM, a);
}

int try_conv_fprop_row(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w, const float* bias,
                       const dnnca_tensor_t* y, int act, float alpha) {
  return launch_row(s, false, 0, x, x2, w, bias, y, nullptr, nullptr, act, alpha);
}

int try_conv_dgrad_row(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                       const dnnca_tensor_t* dx2, const dnnca_tensor_t* mask, int act, float alpha) {
  return launch_row(s, true, 0, dz, nullptr, w, nullptr, dx, dx2, mask, act, alpha);
}

// Conv2DTranspose k=s=2 (components.py:118-120): y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci] k[a,b,co,ci] + bias[co]
int try_tconv_fprop_row(cudaStream_t s, const dnnca_tensor_t* x, const float* k, const float* bias, const dnnca_tensor_t* y) {
  return launch_row(s, false, 1, x, nullptr, k, bias, y, nullptr, nullptr, DNNCA_ACT_NONE, 0.f);
}
int try_tconv_dgrad_row(cudaStream_t s, const dnnca_tensor_t* dy, const float* k, const dnnca_tensor_t* dx,
                        const dnnca_tensor_t* mask, int act, float alpha) {
  return launch_row(s, true, 2, dy, nullptr, k, nullptr, dx, nullptr, mask, act, alpha);
}

// accumulators of one K step: pair the 64-element blocks (taps of one-atom windows, then the all-ones block)
static bool roww_plan(RowArgs& a) {
  int n = 0;
  if (a.tconv) {                 // (x window | ones) against the dy rows 2i and 2i+1
    if (a.atoms[0] > 1) return false;
    for (int ar = 0; ar < 2; ++ar) {
      RowAcc& c = a.acc[n++];
      memset(&c, 0, sizeof(c));
      c.kind[0] = WB_WINDOW; c.tap[0] = (unsigned char)ar; c.kind[1] = WB_ONES; c.first = 0; c.second = -1; c.zi = (unsigned char)ar;
    }
    a.nacc = n;
    return n * a.N <= 512;
  }
  auto window_off = [&](int op, int tap) { return (op ? a.atoms[0] * a.abuf : 0) + tap * 128 * a.ipt; };
  if (a.atoms[0] > 1) {
    for (int op = 0; op < a.nops; ++op)
      for (int tap = 0; tap < 3; ++tap) {
        RowAcc& c = a.acc[n++];
        memset(&c, 0, sizeof(c));
        c.kind[0] = c.kind[1] = WB_WINDOW128; c.op[0] = c.op[1] = (unsigned char)op; c.tap[0] = c.tap[1] = (unsigned char)tap;
        c.first = window_off(op, tap); c.second = c.first + a.abuf; c.m64 = 0;
      }
    RowAcc& c = a.acc[n++];
    memset(&c, 0, sizeof(c));
    c.kind[0] = WB_ONES; c.first = -1; c.m64 = 1;
  } else {
    struct Blk { int kind, op, tap, off; } blk[8];
    int nb = 0;
    for (int op = 0; op < a.nops; ++op)
      for (int tap = 0; tap < 3; ++tap) blk[nb++] = Blk{WB_WINDOW, op, tap, window_off(op, tap)};
    blk[nb++] = Blk{WB_ONES, 0, 0, -1};
    for (int i = 0; i < nb; i += 2) {
      RowAcc& c = a.acc[n++];
      memset(&c, 0, sizeof(c));
      c.kind[0] = (unsigned char)blk[i].kind; c.op[0] = (unsigned char)blk[i].op; c.tap[0] = (unsigned char)blk[i].tap;
      c.first = blk[i].off;
      if (i + 1 < nb) {
        c.kind[1] = (unsigned char)blk[i + 1].kind; c.op[1] = (unsigned char)blk[i + 1].op; c.tap[1] = (unsigned char)blk[i + 1].tap;
        c.second = blk[i + 1].off; c.m64 = 0;
      } else {
        c.m64 = 1;
      }
    }
  }
  a.nacc = n;
  return n <= ROW_MAX_ACC && n * a.N <= 512;
}

static int launch_row_wgrad(cudaStream_t s, int tconv, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* dz,
                            float* dw, float* db) {
  if (!row_enabled()) return 0;
  if (!row_dense_bf16(x) || (x2 && !row_dense_bf16(x2)) || !row_dense_bf16(dz)) return 0;
  const int ca = x->c, cb = x2 ? x2->c : 0, co = dz->c;
  if (ca > ROW_MAX_C || cb > ROW_MAX_C || co > ROW_MAX_C) return 0;
  RowArgs a;
  memset(&a, 0, sizeof(a));
  a.nops = x2 ? 2 : 1;
  a.tconv = tconv ? 3 : 0;
  a.cin_tot = ca + cb; a.cout = co; a.coff[0] = 0; a.coff[1] = ca; a.dw = dw; a.db = db;
  a.nwt = (tconv ? 4 : 9) * a.cin_tot * co;
  a.ztaps = tconv ? 2 : 1;
  a.ltaps = 1;
  a.dbg = row_dbg();
  bool found = false;
  RowArgs best;
  double best_cost = 0.0;
  for (int pass = 0; pass < 2; ++pass)
  for (int P = 16; P >= 2; P >>= 1) {
    a.ipt_max = pass ? 1 : 8;
    a.P = P;
    if (!row_operand(a, 0, ca, P) || (x2 && !row_operand(a, 1, cb, P))) continue;
    if (x2 && a.atoms[0] != a.atoms[1]) continue;
    a.N = (tconv ? 2 : 1) * P * co;
    if (a.N % 16 || a.N > 256 || a.N < 16) continue;
    if (!row_common_geometry(a, x)) continue;
    if (!roww_plan(a)) continue;
    a.zatoms = (a.N + 63) / 64; a.zbuf = a.Hs * 128;
    bool fits = false;
    for (a.nstage = 6; a.nstage >= 2; --a.nstage)
      if (roww_smem_bytes(a) <= ROW_SMEM_LIMIT) { fits = true; break; }
    if (!fits || (a.N + 1) * 512 > a.nstage * roww_stage_bytes(a)) continue;     // the ring also hosts the accumulator dump
    const double cost = (a.nacc * (a.Hs / 16) * row_mma_cycles(a.N) + 200.0) / ((double)P * a.Hs);
    if (!found || cost < best_cost) { best = a; best_cost = cost; found = true; }
  }
  if (found) a = best;
  if (!found) return 0;
  CUtensorMap mA, mB, mZ;
  if (!row_map(&mA, x, 64, a.hrows + 2 * a.padr, true, a.ipt)) return 0;
  mB = mA;
  if (x2 && !row_map(&mB, x2, 64, a.hrows + 2 * a.padr, true, a.ipt)) return 0;
  if (tconv) {
    if (!row_map_parity(&mZ, dz, a.Hs)) return 0;
  } else if (!row_map(&mZ, dz, 64, a.hrows, true, a.ipt)) return 0;
  const int smem = roww_smem_bytes(a);
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_row_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "conv_row_wgrad: cudaFuncSetAttribute");
    smem_set = smem;
  }
  const int ntiles = a.tiles_x * a.tiles_y * a.nimg;
  {
    cudaError_t e = launch_pdl(conv_row_wgrad_kernel, persistent_grid(ntiles), 192, smem, s, mA, mB, mZ, a);
    if (e != cudaSuccess) return cuda_fail(e, "conv_row_wgrad: launch");
  }
  DNNCA_LAUNCH_CHECK("conv_row_wgrad");
  note_family(2);
  return 1;
}

int try_conv_wgrad_row(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* dz, float* dw,
                       float* db) {
  return launch_row_wgrad(s, 0, x, x2, dz, dw, db);
}
// dk[a][b][co][ci] += sum x[n,i,j,ci] dy[n,2i+a,2j+b,co] ; db[co] += sum dy
int try_tconv_wgrad_row(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk, float* db) {
  return launch_row_wgrad(s, 1, x, nullptr, dy, dk, db);
}

}  // namespace dnnca
