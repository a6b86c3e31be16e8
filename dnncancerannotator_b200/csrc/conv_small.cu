// Direct 3x3 convolution for tiny channel counts (configs/unet.yaml: 3/6/12 channels and the
// 12+12 / 6+6 / 3+3 decoder concats; first layers of the other models: Cin = 1/3/5).
// These layers sit far below the tensor-core ridge (4-72 flop/B vs ~258) and their im2col rows
// (6-byte pixels) cannot form UMMA/TMA-im2col operands, so they run on the FP32 pipe:
//
//   * TMA-staged NHWC tiles: a dense NHWC tensor is described to the TMA as rows of 16-byte
//     chunks (tma.cuh); ONE cp.async.bulk.tensor box brings a (rows+halo) x (pixels+halo) tile
//     into shared memory, zero-filling everything outside the image = the conv's 'same' padding.
//     No per-element address arithmetic is spent on staging.
//   * the raw tile is de-interleaved once per CTA into fp32 channel planes xs[ci][row][col];
//     every thread owns 4 adjacent pixels x ALL output channels.
//   * packed FFMA2 (fma.rn.f32x2): accumulators are pixel pairs, weights sit in shared memory
//     as duplicated (w,w) pairs -> one LDS.128 feeds 4 FFMA2 = 8 FMA.  Measured on B200: FFMA2
//     does not raise the FMA-pipe peak (71 TFLOP/s either way) but halves the issue slots the
//     FMAs need, so the LDS / address / unpack instructions issue in their shadow.
//   * a second input tensor (x2) is consumed in the same K loop: tf.concat([tconv, skip])
//     (components.py:164) is never materialised; dgrad symmetrically writes two outputs.
//   * fprop and dgrad are the same kernel: dgrad stages rot180/transposed weights and swaps
//     the bias+activation epilogue for the act'(mask) product.
//   * results leave through shared memory + one TMA store per output tensor (coalesced even
//     for 6-byte pixels); BatchNorm statistics of the stored values are reduced per CTA into
//     fp64 atomics.
//   * wgrad: persistent CTAs; thread = (4-pixel group slot, ci, co-block) keeps the 3x6 input
//     window of its channel as register pairs and streams dz pairs: 18 FFMA2 per LDS.128.
//
// Compiled once per (dtype, kind) with -DSMALL_DT=0|1 (f32|bf16) and -DSMALL_KIND=0|1|2
// (fprop|dgrad|wgrad) so the template instantiations build in parallel.
// Reference call sites: layers.Conv2D components.py:47-50,123-126 and their gradients.
#include "small_common.cuh"

#ifndef SMALL_DT
#error "compile with -DSMALL_DT=0|1 -DSMALL_KIND=0|1|2"
#endif
#if SMALL_DT == 0
#define SMALL_T float
#define SMALL_FN(name) name##_f32
#else
#define SMALL_T __nv_bfloat16
#define SMALL_FN(name) name##_bf16
#endif

namespace dnnca {

// ----------------------------------------------------------------------------
// fprop / dgrad
// ----------------------------------------------------------------------------
template <typename T, int CA, int CB, int OA, int OB, int TXN>
struct FGeom {
  static constexpr int TYN = 256 / TXN, TW = TXN * PX;
  static constexpr int ROWS = TYN + 2, PITCH = TW + 4;
  static constexpr int CIN = CA + CB, COT = OA + OB, COTP = ru(COT, 2);
  static constexpr int NCH = cmax(Raw<T, CA, TW, 1>::NCH, Raw<T, CB, TW, 1>::NCH);
  static constexpr int CK = cmax(CA, CB);
  static constexpr int RAW_BYTES = ru(ROWS * NCH * 16, 128);
  static constexpr int WS_BYTES = ru(CIN * 9 * COTP * 8, 128);
  static constexpr int OSA_BYTES = ru(TYN * TW * OA * (int)sizeof(T), 128);
  static constexpr int OSB_BYTES = ru(TYN * TW * OB * (int)sizeof(T), 128);
  static constexpr int PLANES = CK * ROWS * PITCH;        // floats
  // a second, one-pixel-shifted copy of the planes (aligned odd pixel pairs without register moves) where shared
  // memory is cheap: the 1/3/4-channel layers at full resolution, whose few FFMA2 per window amortise moves worst
  static constexpr bool SHIFT = false;     // measured: the extra STS/LDS cost more than the moves they save here
  static constexpr int XS_BYTES = cmax(ru((SHIFT ? 2 : 1) * PLANES * 4, 128), OSA_BYTES + OSB_BYTES);
  static constexpr int OFF_RAW = 128, OFF_WS = OFF_RAW + RAW_BYTES, OFF_XS = OFF_WS + WS_BYTES;
  static constexpr int SMEM = OFF_XS + XS_BYTES;
  static constexpr int EPC = 16 / (int)sizeof(T);
  static constexpr bool FITS = NCH <= 256 && TW * OA / EPC <= 256 && TW * OB / EPC <= 256 && SMEM <= 200 * 1024;
};

struct FArgs {
  const float* w;
  const float* bias;
  View ya, yb, mask;       // outputs (manual-store fallback / geometry) and the dgrad mask
  int has_mask, act;
  float alpha;
  double* stats;
  int tiles_x, tiles_y, tma_out;
};

template <typename T, int CA, int CB, int OA, int OB, int TXN, bool DGRAD>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB,
                                                           const __grid_constant__ CUtensorMap mapOA,
                                                           const __grid_constant__ CUtensorMap mapOB, FArgs a) {
  using G = FGeom<T, CA, CB, OA, OB, TXN>;
  constexpr int TYN = G::TYN, TW = G::TW, ROWS = G::ROWS, PITCH = G::PITCH;
  constexpr int CIN = G::CIN, COT = G::COT, COTP = G::COTP;
  constexpr int CBS = CB ? CB : 1;                            // keeps dead template arguments well-formed
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  T* raw = reinterpret_cast<T*>(smem + G::OFF_RAW);
  float* ws2 = reinterpret_cast<float*>(smem + G::OFF_WS);   // [CIN][3][3][COTP] of (w,w) pairs
  float* xs = reinterpret_cast<float*>(smem + G::OFF_XS);    // [CK][ROWS][PITCH], later the output tiles

  int b = blockIdx.x;
  const int tix = b % a.tiles_x; b /= a.tiles_x;
  const int tiy = b % a.tiles_y;
  const int n = b / a.tiles_y;
  const int x0 = tix * TW, y0 = tiy * TYN;
  const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, ROWS * Raw<T, CA, TW, 1>::NCH * 16);
    tma_load_4d(raw, &mapA, bar, 0, Raw<T, CA, TW, 1>::chunk_start(x0), y0 - 1, n);
  }
  // weights -> smem as duplicated pairs, [ci][dy][dx][co] (dgrad: rot180 + in/out transpose)
  for (int e = threadIdx.x; e < CIN * 9 * COTP; e += 256) {
    const int co = e % COTP;
    const int t = e / COTP;
    const int tap = t % 9;
    const int ci = t / 9;
    const int dy = tap / 3, dx = tap % 3;
    float v = 0.f;
    if (co < COT) {
      if (!DGRAD) v = a.w[((dy * 3 + dx) * CIN + ci) * COT + co];
      else        v = a.w[(((2 - dy) * 3 + (2 - dx)) * COT + co) * CIN + ci];  // layer Cin = COT, layer Cout = CIN
    }
    ws2[2 * e] = v;
    ws2[2 * e + 1] = v;
  }

  u64 acc[COTP][2];
#pragma unroll
  for (int co = 0; co < COTP; ++co) acc[co][0] = acc[co][1] = 0ull;

#pragma unroll
  for (int s = 0; s < (CB ? 2 : 1); ++s) {
    const int cs = s == 0 ? CA : CB;
    const int cbase = s == 0 ? 0 : CA;
    mbar_wait(bar, s);
    if (s == 0) deinterleave<T, CA, TW, 1>(raw, xs, ROWS, PITCH, G::SHIFT ? G::PLANES : 0);
    else        deinterleave<T, CBS, TW, 1>(raw, xs, ROWS, PITCH, G::SHIFT ? G::PLANES : 0);
    __syncthreads();                                   // planes ready, raw buffer free, weights visible
    if (CB && s == 0 && threadIdx.x == 0) {            // second input streams in behind the first one's math
      mbar_expect_tx(bar, ROWS * Raw<T, CBS, TW, 1>::NCH * 16);
      tma_load_4d(raw, &mapB, bar, 0, Raw<T, CBS, TW, 1>::chunk_start(x0), y0 - 1, n);
    }
#pragma unroll 1
    for (int ci = 0; ci < cs; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* xr = xs + (ci * ROWS + ty + dy) * PITCH + tx * PX;
        const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(xr);                 // (c0,c1) (c2,c3)
        const u64 p45 = *reinterpret_cast<const u64*>(xr + 4);                         // (c4,c5)
        const u64 p01 = q.x, p23 = q.y;
        u64 p12, p34;
        if (G::SHIFT) {
          const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(xr + G::PLANES);  // shifted copy: (c1,c2) (c3,c4)
          p12 = q1.x; p34 = q1.y;
        } else {
          p12 = pack2(hi32(p01), lo32(p23)); p34 = pack2(hi32(p23), lo32(p45));
        }
        const float* wrow = ws2 + ((cbase + ci) * 9 + dy * 3) * COTP * 2;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const u64 lo = dx == 0 ? p01 : (dx == 1 ? p12 : p23);
          const u64 hi = dx == 0 ? p23 : (dx == 1 ? p34 : p45);
#pragma unroll
          for (int cp = 0; cp < COTP / 2; ++cp) {
            const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wrow + (dx * COTP + 2 * cp) * 2);
            acc[2 * cp][0] = ffma2(lo, wv.x, acc[2 * cp][0]);
            acc[2 * cp][1] = ffma2(hi, wv.x, acc[2 * cp][1]);
            acc[2 * cp + 1][0] = ffma2(lo, wv.y, acc[2 * cp + 1][0]);
            acc[2 * cp + 1][1] = ffma2(hi, wv.y, acc[2 * cp + 1][1]);
          }
        }
      }
    }
    __syncthreads();                                   // everyone done with the planes
  }

  // ---- epilogue: output tile(s) [row][col][c] in T, then TMA store (or guarded manual store) ----
  T* osA = reinterpret_cast<T*>(xs);
  T* osB = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(xs) + G::OSA_BYTES);
  const int gy = y0 + ty;
  float ssum[COT], ssq[COT];
#pragma unroll
  for (int co = 0; co < COT; ++co) { ssum[co] = 0.f; ssq[co] = 0.f; }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int gx = x0 + tx * PX + p;
    const bool inside = gy < a.ya.h && gx < a.ya.w;
#pragma unroll
    for (int co = 0; co < COT; ++co) {
      float v = (p & 1) ? hi32(acc[co][p >> 1]) : lo32(acc[co][p >> 1]);
      if (!DGRAD) {
        v = apply_act(v + (a.bias ? a.bias[co] : 0.f), a.act, a.alpha);
      } else if (a.has_mask && co < OA && inside) {
        const T* mp = reinterpret_cast<const T*>(a.mask.data) +
                      (((long long)n * a.mask.h + gy) * a.mask.w + gx) * a.mask.cstride + a.mask.coff;
        v *= act_grad(ldf(mp + co), a.act, a.alpha);
      }
      if (!DGRAD && inside) {
        const float r = rnd<T>(v);
        ssum[co] += r;
        ssq[co] += r * r;
      }
      if (co < OA) stf(osA + (ty * TW + tx * PX + p) * OA + co, v);
      else         stf(osB + (ty * TW + tx * PX + p) * (OB ? OB : 1) + (co - OA), v);
    }
  }
  if (a.tma_out) {
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
      constexpr int EPC = 16 / (int)sizeof(T);
      tma_store_4d(&mapOA, osA, 0, x0 * OA / EPC, y0, n);
      if (OB) tma_store_4d(&mapOB, osB, 0, x0 * OB / EPC, y0, n);
      tma_store_commit();
      tma_store_wait_read();
    }
  } else {
    __syncthreads();
    const int wv = min(TW, a.ya.w - x0), hv = min(TYN, a.ya.h - y0);
#pragma unroll
    for (int o = 0; o < (OB ? 2 : 1); ++o) {
      const View& yv = o ? a.yb : a.ya;
      const T* os = o ? osB : osA;
      const int oc = o ? OB : OA;
      T* ybase = reinterpret_cast<T*>(yv.data) + yv.coff;
      const int row_elems = wv * oc;
      for (int e = threadIdx.x; e < hv * row_elems; e += 256) {
        const int row = e / row_elems, r = e % row_elems;
        const int col = r / oc, co = r % oc;
        ybase[(((long long)n * yv.h + y0 + row) * yv.w + x0 + col) * yv.cstride + co] = os[(row * TW + col) * oc + co];
      }
    }
  }
  if (!DGRAD && a.stats) {
    __shared__ float red[8][2 * COT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int co = 0; co < COT; ++co) {
      const float s = warp_sum(ssum[co]), q = warp_sum(ssq[co]);
      if (lane == 0) { red[warp][co] = s; red[warp][COT + co] = q; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * COT) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
      atomicAdd(a.stats + threadIdx.x, s);
    }
  }
}

// ----------------------------------------------------------------------------
// wgrad: dw[a][c][ci][co] += sum_p [x|x2][p+(a-1,c-1)][ci] * dz[p][co] ; db[co] += sum_p dz[p][co]
// ----------------------------------------------------------------------------
template <typename T, int CA, int CB, int COUT>
struct WGeom {
  static constexpr int CIN = CA + CB;
  // 8-row tiles halve the per-tile fixed cost; 4 rows when the planes of a wide layer would not let 2 CTAs share an SM
  static constexpr int TYN = (CIN + COUT <= 24) ? 8 : 4, TXN = 16, TW = TXN * PX, ROWS = TYN + 2, PITCH = TW + 4;
  // output channels per thread: 9*COB accumulator pairs must leave room for 2 CTAs (<= 128 registers/thread)
  static constexpr int COB = COUT % 3 == 0 ? 3 : (COUT % 4 == 0 ? 4 : COUT);
  static constexpr int NCB = COUT / COB;
  static constexpr int TPS = CIN * NCB;                                    // threads per pixel-group slot
  static constexpr int G = 256 / TPS;      // (a power-of-two slot count was measured slower: fewer active threads)
  static constexpr int NCHA = Raw<T, CA, TW, 1>::NCH, NCHB = Raw<T, CB, TW, 1>::NCH, NCHG = Raw<T, COUT, TW, 0>::NCH;
  static constexpr int RAWA = ru(ROWS * NCHA * 16, 128), RAWB = ru(ROWS * NCHB * 16, 128), RAWG = ru(TYN * NCHG * 16, 128);
  static constexpr int PLANES = CIN * ROWS * PITCH;       // floats
  static constexpr bool SHIFT = CIN <= 6;                 // + one-pixel-shifted copy of the planes (see FGeom)
  static constexpr int NS = (RAWA + RAWB + RAWG) <= 16 * 1024 ? 3 : 2;   // TMA stages (raw tiles in flight)
  static constexpr int XS = ru((SHIFT ? 2 : 1) * PLANES * 4, 128), GS = ru(COUT * TYN * TW * 4, 128);
  static constexpr int RED = ru((9 * CIN * COUT + COUT) * 4, 128);
  static constexpr int RAWS = RAWA + RAWB + RAWG;         // one stage of raw tiles
  static constexpr int OFF_RA = 128, OFF_XS = OFF_RA + NS * RAWS, OFF_GS = OFF_XS + XS, OFF_RED = OFF_GS + GS,
                       SMEM = OFF_RED + RED;
  static constexpr bool FITS = COUT % COB == 0 && TPS <= 256 && NCHA <= 256 && NCHB <= 256 && NCHG <= 256 && SMEM <= 200 * 1024;
};

template <typename T, int CA, int CB, int COUT>
__global__ void __launch_bounds__(256, 2) conv3x3_small_wgrad_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                 const __grid_constant__ CUtensorMap mapB,
                                                                 const __grid_constant__ CUtensorMap mapG,
                                                                 float* __restrict__ dw, float* __restrict__ db,
                                                                 int tiles_x, int tiles_y, int ntiles) {
  using G = WGeom<T, CA, CB, COUT>;
  constexpr int TYN = G::TYN, TXN = G::TXN, TW = G::TW, ROWS = G::ROWS, PITCH = G::PITCH;
  constexpr int CIN = G::CIN, COB = G::COB, NCB = G::NCB;
  constexpr int NGRP = TYN * TXN;
  constexpr int CBS = CB ? CB : 1;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);        // [NS]
  unsigned char* rawbase = smem + G::OFF_RA;
  float* xs = reinterpret_cast<float*>(smem + G::OFF_XS);   // [CIN][ROWS][PITCH] + shifted copy
  float* gs = reinterpret_cast<float*>(smem + G::OFF_GS);   // [COUT][TYN][TW]
  float* red = reinterpret_cast<float*>(smem + G::OFF_RED);
  constexpr int NS = G::NS;

  const int slot = threadIdx.x / G::TPS;
  const int rem = threadIdx.x % G::TPS;
  const int ci = rem / NCB, cb = rem % NCB;
  const bool active = slot < G::G;

  u64 acc[9][COB];
  float dbacc[COB];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < COB; ++c) acc[t][c] = 0ull;
#pragma unroll
  for (int c = 0; c < COB; ++c) dbacc[c] = 0.f;
  for (int e = threadIdx.x; e < 9 * CIN * COUT + COUT; e += 256) red[e] = 0.f;

  constexpr uint32_t TX_BYTES = ROWS * G::NCHA * 16 + ROWS * G::NCHB * 16 + TYN * G::NCHG * 16;
  auto issue = [&](int tile, int st) {
    int b = tile;
    const int tix = b % tiles_x; b /= tiles_x;
    const int tiy = b % tiles_y;
    const int n = b / tiles_y;
    const int x0 = tix * TW, y0 = tiy * TYN;
    unsigned char* rb = rawbase + st * G::RAWS;
    mbar_expect_tx(bar + st, TX_BYTES);
    tma_load_4d(rb, &mapA, bar + st, 0, Raw<T, CA, TW, 1>::chunk_start(x0), y0 - 1, n);
    if (CB) tma_load_4d(rb + G::RAWA, &mapB, bar + st, 0, Raw<T, CBS, TW, 1>::chunk_start(x0), y0 - 1, n);
    tma_load_4d(rb + G::RAWA + G::RAWB, &mapG, bar + st, 0, Raw<T, COUT, TW, 0>::chunk_start(x0), y0, n);
  };
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(bar + i, 1);
    fence_mbar_init();
    for (int i = 0; i < NS - 1; ++i)                        // NS-1 tiles in flight before the first wait
      if (blockIdx.x + i * (int)gridDim.x < ntiles) issue(blockIdx.x + i * gridDim.x, i);
  }
  __syncthreads();

  int it = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int st = it % NS;
    mbar_wait(bar + st, (it / NS) & 1);
    const unsigned char* rb = rawbase + st * G::RAWS;
    deinterleave<T, CA, TW, 1>(reinterpret_cast<const T*>(rb), xs, ROWS, PITCH, G::SHIFT ? G::PLANES : 0);
    if (CB) deinterleave<T, CBS, TW, 1>(reinterpret_cast<const T*>(rb + G::RAWA), xs + CA * ROWS * PITCH, ROWS, PITCH, G::SHIFT ? G::PLANES : 0);
    deinterleave<T, COUT, TW, 0>(reinterpret_cast<const T*>(rb + G::RAWA + G::RAWB), gs, TYN, TW);
    __syncthreads();                                       // planes ready, this raw stage free
    {
      const int nt = tile + (NS - 1) * (int)gridDim.x;     // refill the stage consumed in the previous iteration
      if (threadIdx.x == 0 && nt < ntiles) issue(nt, (it + NS - 1) % NS);
    }
    if (active) {
#pragma unroll 1
      for (int g = slot; g < NGRP; g += G::G) {
        const int ty = g / TXN, tx = g % TXN;
        u64 pr[3][5];                                       // per dy: (0,1) (1,2) (2,3) (3,4) (4,5)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float* xr = xs + (ci * ROWS + ty + dy) * PITCH + tx * PX;
          const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(xr);
          pr[dy][0] = q.x;
          pr[dy][2] = q.y;
          pr[dy][4] = *reinterpret_cast<const u64*>(xr + 4);
          if (G::SHIFT) {
            const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(xr + G::PLANES);
            pr[dy][1] = q1.x;
            pr[dy][3] = q1.y;
          } else {
            pr[dy][1] = pack2(hi32(q.x), lo32(q.y));
            pr[dy][3] = pack2(hi32(q.y), lo32(pr[dy][4]));
          }
        }
#pragma unroll
        for (int c = 0; c < COB; ++c) {
          const int co = cb * COB + c;
          const ulonglong2 gv = *reinterpret_cast<const ulonglong2*>(gs + (co * TYN + ty) * TW + tx * PX);
          if (ci == 0) dbacc[c] += (lo32(gv.x) + hi32(gv.x)) + (lo32(gv.y) + hi32(gv.y));
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              acc[dy * 3 + dx][c] = ffma2(pr[dy][dx], gv.x, acc[dy * 3 + dx][c]);
              acc[dy * 3 + dx][c] = ffma2(pr[dy][dx + 2], gv.y, acc[dy * 3 + dx][c]);
            }
        }
      }
    }
    __syncthreads();                                       // before the next tile overwrites the planes
  }
  // CTA reduction in shared memory, then one global atomic per output per CTA
  if (active) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int c = 0; c < COB; ++c)
        atomicAdd(red + (t * CIN + ci) * COUT + cb * COB + c, lo32(acc[t][c]) + hi32(acc[t][c]));
    if (ci == 0) {
#pragma unroll
      for (int c = 0; c < COB; ++c) atomicAdd(red + 9 * CIN * COUT + cb * COB + c, dbacc[c]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * CIN * COUT; e += 256) atomicAdd(dw + e, red[e]);
  if (db)
    for (int e = threadIdx.x; e < COUT; e += 256) atomicAdd(db + e, red[9 * CIN * COUT + e]);
}

// ----------------------------------------------------------------------------
// host dispatch
// ----------------------------------------------------------------------------
static bool views_ok(const dnnca_tensor_t* t) { return t == nullptr || tma_row_ok(t); }

template <typename T, int CA, int CB, int OA, int OB, int TXN, bool DGRAD>
static int launch_small(cudaStream_t s, const dnnca_tensor_t* xa, const dnnca_tensor_t* xb, const float* w,
                        const float* bias, const dnnca_tensor_t* ya, const dnnca_tensor_t* yb,
                        const dnnca_tensor_t* mask, int act, float alpha, double* stats) {
  using G = FGeom<T, CA, CB, OA, OB, TXN>;
  if constexpr (!G::FITS) {
    return 0;
  } else {
    constexpr int EPC = 16 / (int)sizeof(T);
    constexpr int CBS = CB ? CB : 1;
    auto kern = conv3x3_small_kernel<T, CA, CB, OA, OB, TXN, DGRAD>;
    static bool attr_done = false;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
      if (e != cudaSuccess) return cuda_fail(e, "conv3x3_small: cudaFuncSetAttribute");
      attr_done = true;
    }
    CUtensorMap mA, mB, mOA, mOB;
    if (!make_row_map(&mA, xa, Raw<T, CA, G::TW, 1>::NCH, G::ROWS)) return 0;
    mB = mA;
    if (CB && !make_row_map(&mB, xb, Raw<T, CBS, G::TW, 1>::NCH, G::ROWS)) return 0;
    const bool tma_out = tma_row_ok(ya) && (!OB || tma_row_ok(yb));
    mOA = mA;
    mOB = mA;
    if (tma_out) {
      if (!make_row_map(&mOA, ya, G::TW * OA / EPC, G::TYN)) return 0;
      if (OB && !make_row_map(&mOB, yb, G::TW * OB / EPC, G::TYN)) return 0;
    }
    FArgs a;
    a.w = w; a.bias = bias; a.ya = mk(ya); a.yb = yb ? mk(yb) : mk(ya); a.mask = mask ? mk(mask) : mk(ya);
    a.has_mask = mask != nullptr; a.act = act; a.alpha = alpha; a.stats = stats;
    a.tiles_x = (xa->w + G::TW - 1) / G::TW; a.tiles_y = (xa->h + G::TYN - 1) / G::TYN; a.tma_out = tma_out;
    const long long nblk = (long long)a.tiles_x * a.tiles_y * xa->n;
    if (nblk > 0x7fffffffLL) return 0;
    kern<<<(unsigned)nblk, 256, G::SMEM, s>>>(mA, mB, mOA, mOB, a);
    DNNCA_LAUNCH_CHECK("conv3x3_small");
    note_family(1);
    return 1;
  }
}

template <typename T, int CA, int CB, int OA, int OB, bool DGRAD>
static int launch_small_w(cudaStream_t s, const dnnca_tensor_t* xa, const dnnca_tensor_t* xb, const float* w,
                          const float* bias, const dnnca_tensor_t* ya, const dnnca_tensor_t* yb,
                          const dnnca_tensor_t* mask, int act, float alpha, double* stats) {
  // widest tile that fits the image width and the TMA box limits
  int r = 0;
  if (xa->w > 64) r = launch_small<T, CA, CB, OA, OB, 32, DGRAD>(s, xa, xb, w, bias, ya, yb, mask, act, alpha, stats);
  if (r == 0 && xa->w > 32) r = launch_small<T, CA, CB, OA, OB, 16, DGRAD>(s, xa, xb, w, bias, ya, yb, mask, act, alpha, stats);
  if (r == 0) r = launch_small<T, CA, CB, OA, OB, 8, DGRAD>(s, xa, xb, w, bias, ya, yb, mask, act, alpha, stats);
  return r;
}

// (input channels of x, of x2, output channels) occurring in configs/unet.yaml (C = 3 or 5 modalities), the
// first layers of mulmo_unet.yaml (1 -> 16) and the small golden-test nets (F = 4)
#define DNNCA_SMALL_FPROP_SHAPES(X) \
  X(1, 0, 4) X(1, 0, 16) X(3, 0, 3) X(5, 0, 3) X(3, 0, 6) X(6, 0, 6) X(6, 0, 12) X(12, 0, 12) X(3, 3, 3) X(6, 6, 6) \
  X(12, 12, 12) X(3, 0, 4) X(4, 0, 4) X(4, 0, 8) X(8, 0, 8) X(4, 4, 4) X(8, 8, 8)
// dgrad: kernel input = dz (layer Cout); outputs = dx (+ dx2); first-layer dgrads go to the generic kernel
#define DNNCA_SMALL_DGRAD_SHAPES(X) \
  X(3, 3, 0) X(6, 3, 0) X(6, 6, 0) X(12, 6, 0) X(12, 12, 0) X(3, 3, 3) X(6, 6, 6) X(12, 12, 12) X(4, 4, 0) X(8, 4, 0) \
  X(8, 8, 0) X(4, 4, 4) X(8, 8, 8)

#if SMALL_KIND == 0
int SMALL_FN(try_conv_fprop_small)(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                                   const float* bias, const dnnca_tensor_t* y, int act, float alpha, double* stats) {
  if (!tma_row_ok(x) || !views_ok(x2)) return 0;
  const int c2 = x2 ? x2->c : 0;
#define X(CA, CB, CO) \
  if (x->c == CA && c2 == CB && y->c == CO) \
    return launch_small_w<SMALL_T, CA, CB, CO, 0, false>(s, x, x2, w, bias, y, nullptr, nullptr, act, alpha, stats);
  DNNCA_SMALL_FPROP_SHAPES(X)
#undef X
  return 0;
}
#elif SMALL_KIND == 1
int SMALL_FN(try_conv_dgrad_small)(cudaStream_t s, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                   const dnnca_tensor_t* dx2, const dnnca_tensor_t* mask, int act, float alpha) {
  if (!tma_row_ok(dz)) return 0;
  const int c2 = dx2 ? dx2->c : 0;
#define X(CI, OA, OB) \
  if (dz->c == CI && dx->c == OA && c2 == OB) \
    return launch_small_w<SMALL_T, CI, 0, OA, OB, true>(s, dz, nullptr, w, nullptr, dx, dx2, mask, act, alpha, nullptr);
  DNNCA_SMALL_DGRAD_SHAPES(X)
#undef X
  return 0;
}
#endif

template <typename T, int CA, int CB, int COUT>
static int launch_small_wgrad(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2,
                              const dnnca_tensor_t* dz, float* dw, float* db) {
  using G = WGeom<T, CA, CB, COUT>;
  if constexpr (!G::FITS) {
    return 0;
  } else {
    auto kern = conv3x3_small_wgrad_kernel<T, CA, CB, COUT>;
    static bool attr_done = false;
    if (!attr_done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
      if (e != cudaSuccess) return cuda_fail(e, "conv3x3_small_wgrad: cudaFuncSetAttribute");
      attr_done = true;
    }
    CUtensorMap mA, mB, mG;
    if (!make_row_map(&mA, x, G::NCHA, G::ROWS)) return 0;
    mB = mA;
    if (CB && !make_row_map(&mB, x2, G::NCHB, G::ROWS)) return 0;
    if (!make_row_map(&mG, dz, G::NCHG, G::TYN)) return 0;
    const int tiles_x = (x->w + G::TW - 1) / G::TW, tiles_y = (x->h + G::TYN - 1) / G::TYN;
    const long long ntiles = (long long)tiles_x * tiles_y * x->n;
    if (ntiles > 0x7fffffffLL) return 0;
    static int per_sm = 0;
    if (per_sm == 0) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, G::SMEM) != cudaSuccess || per_sm < 1) per_sm = 1;
    }
    long long grid = (long long)sm_count() * per_sm;     // persistent: one resident wave
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, 256, G::SMEM, s>>>(mA, mB, mG, dw, db, tiles_x, tiles_y, (int)ntiles);
    DNNCA_LAUNCH_CHECK("conv3x3_small_wgrad");
    note_family(1);
    return 1;
  }
}

#if SMALL_KIND == 2
int SMALL_FN(try_conv_wgrad_small)(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* x2,
                                   const dnnca_tensor_t* dz, float* dw, float* db) {
  if (!tma_row_ok(x) || !views_ok(x2) || !tma_row_ok(dz)) return 0;
  const int c2 = x2 ? x2->c : 0;
#define X(CA, CB, CO) \
  if (x->c == CA && c2 == CB && dz->c == CO) return launch_small_wgrad<SMALL_T, CA, CB, CO>(s, x, x2, dz, dw, db);
  DNNCA_SMALL_FPROP_SHAPES(X)
#undef X
  return 0;
}
#endif

}  // namespace dnnca
