// Conv2DTranspose k=s=2 for tiny channel counts (configs/unet.yaml decoder: 12->12, 12->6, 6->3),
// same machinery as conv_small.cu: TMA-staged dense NHWC tiles, fp32 planes in shared memory,
// FFMA2 pixel pairs, TMA store.  The non-overlapping 2x2 taps make the op a per-pixel
// [Cin] x [Cin, 4*Cout] product followed by a pixel shuffle:
//   fprop : thread = (4 input pixels, tap) -> 4 output pixels x Cout
//   dgrad : thread = (4 input pixels, tap) partial sums over co, butterfly-reduced over the 4 taps
//           (dy is de-interleaved "space-to-depth" so a tap's 4 pixels are contiguous)
//   wgrad : persistent CTAs, thread = (pixel-group slot, ci): dK[tap][co][ci] += x * dy pairs
// Reference call site: layers.Convolution2DTranspose components.py:118-120.
#include "small_common.cuh"

namespace dnnca {

constexpr int TG = 64;   // 4-pixel groups per tile (256 input pixels)

template <typename T, int CIN, int COUT, int TXN>
struct TGeom {
  static constexpr int TIY = TG / TXN, TWI = TXN * PX;              // input tile rows x cols
  static constexpr int EPC = 16 / (int)sizeof(T);
  static constexpr int NCHX = TWI * CIN / EPC + ((TWI * CIN) % EPC ? 1 : 0);
  static constexpr int NCHY = 2 * TWI * COUT / EPC;                  // output / dy tile row in chunks
  static constexpr int COUTP = ru(COUT, 2), CINP = ru(CIN, 2);
  static constexpr int RAWX = ru(TIY * NCHX * 16, 128), RAWY = ru(2 * TIY * NCHY * 16, 128);
  static constexpr int XS = ru(CIN * TIY * TWI * 4, 128);            // planes [ci][TIY][TWI]
  static constexpr int GS = ru(4 * COUT * TIY * TWI * 4, 128);       // planes [tap][co][TIY][TWI]
  static constexpr bool ALIGNED = (TWI * CIN) % EPC == 0 && (2 * TWI * COUT) % EPC == 0;
};

// dy raw tile [2*TIY][2*TWI*COUT] -> gs[tap][co][TIY][TWI] (space-to-depth); one thread per dy pixel
template <typename T, int COUT, int TIY, int TWI>
__device__ __forceinline__ void deinterleave_s2d(const T* __restrict__ raw, float* __restrict__ gs) {
  constexpr int RP = 2 * TWI * COUT;
  for (int e = threadIdx.x; e < 4 * TIY * TWI; e += 256) {
    const int r = e / (2 * TWI), c = e - r * (2 * TWI);
    const int tap = (r & 1) * 2 + (c & 1);
    const T* src = raw + r * RP + c * COUT;
    float* d = gs + ((tap * COUT) * TIY + (r >> 1)) * TWI + (c >> 1);
#pragma unroll
    for (int j = 0; j < COUT; ++j) d[j * TIY * TWI] = ldf(src + j);
  }
}

// ------------------------------------------------------------------ fprop ---
template <typename T, int CIN, int COUT, int TXN>
__global__ void __launch_bounds__(256) tconv_small_fprop_kernel(const __grid_constant__ CUtensorMap mapX,
                                                               const __grid_constant__ CUtensorMap mapY,
                                                               const float* __restrict__ kw,
                                                               const float* __restrict__ bias, int tiles_x,
                                                               int tiles_y) {
  using G = TGeom<T, CIN, COUT, TXN>;
  constexpr int TIY = G::TIY, TWI = G::TWI, COUTP = G::COUTP;
  constexpr int WS = ru(4 * CIN * COUTP * 8, 128);
  constexpr int OFF_RAW = 128, OFF_WS = OFF_RAW + G::RAWX, OFF_XS = OFF_WS + WS, OFF_OS = OFF_XS + G::XS;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  T* raw = reinterpret_cast<T*>(smem + OFF_RAW);
  float* ws2 = reinterpret_cast<float*>(smem + OFF_WS);      // [tap][ci][COUTP] (w,w) pairs
  float* xs = reinterpret_cast<float*>(smem + OFF_XS);
  T* os = reinterpret_cast<T*>(smem + OFF_OS);               // [2*TIY][2*TWI][COUT]

  int b = blockIdx.x;
  const int tix = b % tiles_x; b /= tiles_x;
  const int tiy = b % tiles_y;
  const int n = b / tiles_y;
  const int x0 = tix * TWI, y0 = tiy * TIY;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, TIY * G::NCHX * 16);
    tma_load_4d(raw, &mapX, bar, 0, x0 * CIN / G::EPC, y0, n);
  }
  for (int e = threadIdx.x; e < 4 * CIN * COUTP; e += 256) {
    const int co = e % COUTP;
    const int t = e / COUTP;
    const int ci = t % CIN, tap = t / CIN;
    const float v = co < COUT ? kw[(tap * COUT + co) * CIN + ci] : 0.f;
    ws2[2 * e] = v;
    ws2[2 * e + 1] = v;
  }
  mbar_wait(bar, 0);
  deinterleave<T, CIN, TWI, 0>(raw, xs, TIY, TWI);
  __syncthreads();

  const int g = threadIdx.x >> 2, tap = threadIdx.x & 3;
  const int gy = g / TXN, gx = g % TXN;
  u64 acc[COUTP][2];
#pragma unroll
  for (int co = 0; co < COUTP; ++co) acc[co][0] = acc[co][1] = 0ull;
#pragma unroll 2
  for (int ci = 0; ci < CIN; ++ci) {
    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(xs + (ci * TIY + gy) * TWI + gx * PX);
    const float* wrow = ws2 + ((tap * CIN + ci) * COUTP) * 2;
#pragma unroll
    for (int cp = 0; cp < COUTP / 2; ++cp) {
      const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wrow + 4 * cp);
      acc[2 * cp][0] = ffma2(q.x, wv.x, acc[2 * cp][0]);
      acc[2 * cp][1] = ffma2(q.y, wv.x, acc[2 * cp][1]);
      acc[2 * cp + 1][0] = ffma2(q.x, wv.y, acc[2 * cp + 1][0]);
      acc[2 * cp + 1][1] = ffma2(q.y, wv.y, acc[2 * cp + 1][1]);
    }
  }
  const int orow = 2 * gy + (tap >> 1);
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ocol = 2 * (gx * PX + p) + (tap & 1);
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const float v = ((p & 1) ? hi32(acc[co][p >> 1]) : lo32(acc[co][p >> 1])) + (bias ? bias[co] : 0.f);
      stf(os + (orow * 2 * TWI + ocol) * COUT + co, v);
    }
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_4d(&mapY, os, 0, 2 * x0 * COUT / G::EPC, 2 * y0, n);
    tma_store_commit();
    tma_store_wait_read();
  }
}

// ------------------------------------------------------------------ dgrad ---
template <typename T, int CIN, int COUT, int TXN>
__global__ void __launch_bounds__(256) tconv_small_dgrad_kernel(const __grid_constant__ CUtensorMap mapDY,
                                                               const __grid_constant__ CUtensorMap mapDX,
                                                               const float* __restrict__ kw, View mask, int has_mask,
                                                               int act, float alpha, int H, int W, int tiles_x,
                                                               int tiles_y) {
  using G = TGeom<T, CIN, COUT, TXN>;
  constexpr int TIY = G::TIY, TWI = G::TWI, CINP = G::CINP;
  constexpr int WS = ru(4 * COUT * CINP * 8, 128);
  constexpr int OSB = ru(TIY * TWI * CIN * (int)sizeof(T), 128);
  constexpr int OFF_RAW = 128, OFF_WS = OFF_RAW + G::RAWY, OFF_GS = OFF_WS + WS, OFF_OS = OFF_GS + G::GS;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  T* raw = reinterpret_cast<T*>(smem + OFF_RAW);
  float* ws2 = reinterpret_cast<float*>(smem + OFF_WS);      // [tap][co][CINP] (w,w) pairs
  float* gs = reinterpret_cast<float*>(smem + OFF_GS);
  T* os = reinterpret_cast<T*>(smem + OFF_OS);               // [TIY][TWI][CIN]
  (void)OSB;

  int b = blockIdx.x;
  const int tix = b % tiles_x; b /= tiles_x;
  const int tiy = b % tiles_y;
  const int n = b / tiles_y;
  const int x0 = tix * TWI, y0 = tiy * TIY;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 2 * TIY * G::NCHY * 16);
    tma_load_4d(raw, &mapDY, bar, 0, 2 * x0 * COUT / G::EPC, 2 * y0, n);
  }
  for (int e = threadIdx.x; e < 4 * COUT * CINP; e += 256) {
    const int ci = e % CINP;
    const int t = e / CINP;                                   // tap*COUT + co
    const float v = ci < CIN ? kw[t * CIN + ci] : 0.f;
    ws2[2 * e] = v;
    ws2[2 * e + 1] = v;
  }
  mbar_wait(bar, 0);
  deinterleave_s2d<T, COUT, TIY, TWI>(raw, gs);
  __syncthreads();

  const int g = threadIdx.x >> 2, tap = threadIdx.x & 3;
  const int gy = g / TXN, gx = g % TXN;
  u64 acc[CINP][2];
#pragma unroll
  for (int ci = 0; ci < CINP; ++ci) acc[ci][0] = acc[ci][1] = 0ull;
#pragma unroll 2
  for (int co = 0; co < COUT; ++co) {
    const ulonglong2 q = *reinterpret_cast<const ulonglong2*>(gs + ((tap * COUT + co) * TIY + gy) * TWI + gx * PX);
    const float* wrow = ws2 + ((tap * COUT + co) * CINP) * 2;
#pragma unroll
    for (int cp = 0; cp < CINP / 2; ++cp) {
      const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wrow + 4 * cp);
      acc[2 * cp][0] = ffma2(q.x, wv.x, acc[2 * cp][0]);
      acc[2 * cp][1] = ffma2(q.y, wv.x, acc[2 * cp][1]);
      acc[2 * cp + 1][0] = ffma2(q.x, wv.y, acc[2 * cp + 1][0]);
      acc[2 * cp + 1][1] = ffma2(q.y, wv.y, acc[2 * cp + 1][1]);
    }
  }
  // sum the four taps (adjacent lanes), then lane `tap` stores pixel p = tap of the group
  const int gyy = y0 + gy, gxx = x0 + gx * PX + tap;
  const bool inside = gyy < H && gxx < W;
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
    float v[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float t = (p & 1) ? hi32(acc[ci][p >> 1]) : lo32(acc[ci][p >> 1]);
      t += __shfl_xor_sync(0xffffffffu, t, 1);
      t += __shfl_xor_sync(0xffffffffu, t, 2);
      v[p] = t;
    }
    float out = tap == 0 ? v[0] : (tap == 1 ? v[1] : (tap == 2 ? v[2] : v[3]));
    if (has_mask && inside) {
      const T* mp = reinterpret_cast<const T*>(mask.data) + (((long long)n * H + gyy) * W + gxx) * mask.cstride + mask.coff;
      out *= act_grad(ldf(mp + ci), act, alpha);
    }
    stf(os + (gy * TWI + gx * PX + tap) * CIN + ci, out);
  }
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    tma_store_4d(&mapDX, os, 0, x0 * CIN / G::EPC, y0, n);
    tma_store_commit();
    tma_store_wait_read();
  }
}

// ------------------------------------------------------------------ wgrad ---
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(256) tconv_small_wgrad_kernel(const __grid_constant__ CUtensorMap mapX,
                                                               const __grid_constant__ CUtensorMap mapDY,
                                                               float* __restrict__ dk, float* __restrict__ db,
                                                               int tiles_x, int tiles_y, int ntiles) {
  constexpr int TXN = 16;
  using G = TGeom<T, CIN, COUT, TXN>;
  constexpr int TIY = G::TIY, TWI = G::TWI;
  constexpr int SLOTS = 256 / CIN;
  constexpr int RED = ru((4 * COUT * CIN + COUT) * 4, 128);
  constexpr int OFF_RX = 128, OFF_RY = OFF_RX + G::RAWX, OFF_XS = OFF_RY + G::RAWY, OFF_GS = OFF_XS + G::XS,
                OFF_RED = OFF_GS + G::GS;
  (void)RED;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  T* rawX = reinterpret_cast<T*>(smem + OFF_RX);
  T* rawY = reinterpret_cast<T*>(smem + OFF_RY);
  float* xs = reinterpret_cast<float*>(smem + OFF_XS);
  float* gs = reinterpret_cast<float*>(smem + OFF_GS);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);

  const int slot = threadIdx.x / CIN, ci = threadIdx.x % CIN;
  const bool active = slot < SLOTS;
  u64 acc[4][COUT];
  float dbacc[COUT];
#pragma unroll
  for (int t = 0; t < 4; ++t)
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[t][c] = 0ull;
#pragma unroll
  for (int c = 0; c < COUT; ++c) dbacc[c] = 0.f;
  for (int e = threadIdx.x; e < 4 * COUT * CIN + COUT; e += 256) red[e] = 0.f;

  auto issue = [&](int tile) {
    int b = tile;
    const int tix = b % tiles_x; b /= tiles_x;
    const int tiy = b % tiles_y;
    const int n = b / tiles_y;
    const int x0 = tix * TWI, y0 = tiy * TIY;
    mbar_expect_tx(bar, TIY * G::NCHX * 16 + 2 * TIY * G::NCHY * 16);
    tma_load_4d(rawX, &mapX, bar, 0, x0 * CIN / G::EPC, y0, n);
    tma_load_4d(rawY, &mapDY, bar, 0, 2 * x0 * COUT / G::EPC, 2 * y0, n);
  };
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    issue(blockIdx.x);
  }
  __syncthreads();
  uint32_t phase = 0;
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, phase);
    phase ^= 1;
    deinterleave<T, CIN, TWI, 0>(rawX, xs, TIY, TWI);
    deinterleave_s2d<T, COUT, TIY, TWI>(rawY, gs);
    __syncthreads();
    if (threadIdx.x == 0 && tile + (int)gridDim.x < ntiles) issue(tile + gridDim.x);
    if (active) {
#pragma unroll 1
      for (int g = slot; g < TG; g += SLOTS) {
        const int gy = g / TXN, gx = g % TXN;
        const ulonglong2 xq = *reinterpret_cast<const ulonglong2*>(xs + (ci * TIY + gy) * TWI + gx * PX);
#pragma unroll
        for (int tap = 0; tap < 4; ++tap)
#pragma unroll
          for (int co = 0; co < COUT; ++co) {
            const ulonglong2 gq = *reinterpret_cast<const ulonglong2*>(gs + ((tap * COUT + co) * TIY + gy) * TWI + gx * PX);
            if (ci == 0) dbacc[co] += (lo32(gq.x) + hi32(gq.x)) + (lo32(gq.y) + hi32(gq.y));
            acc[tap][co] = ffma2(xq.x, gq.x, acc[tap][co]);
            acc[tap][co] = ffma2(xq.y, gq.y, acc[tap][co]);
          }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int tap = 0; tap < 4; ++tap)
#pragma unroll
      for (int co = 0; co < COUT; ++co)
        atomicAdd(red + (tap * COUT + co) * CIN + ci, lo32(acc[tap][co]) + hi32(acc[tap][co]));
    if (ci == 0) {
#pragma unroll
      for (int co = 0; co < COUT; ++co) atomicAdd(red + 4 * COUT * CIN + co, dbacc[co]);
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 4 * COUT * CIN; e += 256) atomicAdd(dk + e, red[e]);
  if (db)
    for (int e = threadIdx.x; e < COUT; e += 256) atomicAdd(db + e, red[4 * COUT * CIN + e]);
}

// ----------------------------------------------------------- host dispatch ---
template <typename T, int CIN, int COUT, int TXN>
static int launch_tconv_fprop(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const float* bias,
                              const dnnca_tensor_t* y) {
  using G = TGeom<T, CIN, COUT, TXN>;
  constexpr int SMEM = 128 + G::RAWX + ru(4 * CIN * G::COUTP * 8, 128) + G::XS + ru(4 * G::TIY * G::TWI * COUT * (int)sizeof(T), 128);
  if constexpr (!G::ALIGNED || G::NCHX > 256 || G::NCHY > 256 || SMEM > 200 * 1024) {
    return 0;
  } else {
    auto kern = tconv_small_fprop_kernel<T, CIN, COUT, TXN>;
    static bool done = false;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) return cuda_fail(e, "tconv_small_fprop: cudaFuncSetAttribute");
      done = true;
    }
    CUtensorMap mx, my;
    if (!make_row_map(&mx, x, G::NCHX, G::TIY) || !make_row_map(&my, y, G::NCHY, 2 * G::TIY)) return 0;
    const int tiles_x = (x->w + G::TWI - 1) / G::TWI, tiles_y = (x->h + G::TIY - 1) / G::TIY;
    kern<<<(unsigned)((long long)tiles_x * tiles_y * x->n), 256, SMEM, s>>>(mx, my, kw, bias, tiles_x, tiles_y);
    DNNCA_LAUNCH_CHECK("tconv_small_fprop");
    note_family(1);
    return 1;
  }
}

template <typename T, int CIN, int COUT, int TXN>
static int launch_tconv_dgrad(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx,
                              const dnnca_tensor_t* mask, int act, float alpha) {
  using G = TGeom<T, CIN, COUT, TXN>;
  constexpr int SMEM = 128 + G::RAWY + ru(4 * COUT * G::CINP * 8, 128) + G::GS + ru(G::TIY * G::TWI * CIN * (int)sizeof(T), 128);
  if constexpr (!G::ALIGNED || G::NCHX > 256 || G::NCHY > 256 || SMEM > 200 * 1024) {
    return 0;
  } else {
    auto kern = tconv_small_dgrad_kernel<T, CIN, COUT, TXN>;
    static bool done = false;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) return cuda_fail(e, "tconv_small_dgrad: cudaFuncSetAttribute");
      done = true;
    }
    CUtensorMap my, mx;
    if (!make_row_map(&my, dy, G::NCHY, 2 * G::TIY) || !make_row_map(&mx, dx, G::NCHX, G::TIY)) return 0;
    const int tiles_x = (dx->w + G::TWI - 1) / G::TWI, tiles_y = (dx->h + G::TIY - 1) / G::TIY;
    View vm = mask ? mk(mask) : mk(dx);
    kern<<<(unsigned)((long long)tiles_x * tiles_y * dx->n), 256, SMEM, s>>>(my, mx, kw, vm, mask != nullptr, act, alpha,
                                                                             dx->h, dx->w, tiles_x, tiles_y);
    DNNCA_LAUNCH_CHECK("tconv_small_dgrad");
    note_family(1);
    return 1;
  }
}

template <typename T, int CIN, int COUT>
static int launch_tconv_wgrad(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk, float* db) {
  using G = TGeom<T, CIN, COUT, 16>;
  constexpr int SMEM = 128 + G::RAWX + G::RAWY + G::XS + G::GS + ru((4 * COUT * CIN + COUT) * 4, 128);
  if constexpr (!G::ALIGNED || G::NCHX > 256 || G::NCHY > 256 || SMEM > 200 * 1024 || CIN > 256) {
    return 0;
  } else {
    auto kern = tconv_small_wgrad_kernel<T, CIN, COUT>;
    static bool done = false;
    static int per_sm = 1;
    if (!done) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
      if (e != cudaSuccess) return cuda_fail(e, "tconv_small_wgrad: cudaFuncSetAttribute");
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, SMEM) != cudaSuccess || per_sm < 1) per_sm = 1;
      done = true;
    }
    CUtensorMap mx, my;
    if (!make_row_map(&mx, x, G::NCHX, G::TIY) || !make_row_map(&my, dy, G::NCHY, 2 * G::TIY)) return 0;
    const int tiles_x = (x->w + G::TWI - 1) / G::TWI, tiles_y = (x->h + G::TIY - 1) / G::TIY;
    const long long ntiles = (long long)tiles_x * tiles_y * x->n;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > ntiles) grid = ntiles;
    kern<<<(unsigned)grid, 256, SMEM, s>>>(mx, my, dk, db, tiles_x, tiles_y, (int)ntiles);
    DNNCA_LAUNCH_CHECK("tconv_small_wgrad");
    note_family(1);
    return 1;
  }
}

// (Cin, Cout) of the transposed convs in configs/unet.yaml and the small golden-test nets
#define DNNCA_TCONV_SHAPES(X) X(12, 12) X(12, 6) X(6, 3) X(6, 6) X(8, 8) X(8, 4) X(24, 8)

template <typename T>
static int tconv_fprop_t(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const float* bias,
                         const dnnca_tensor_t* y) {
#define X(CI, CO)                                                              \
  if (x->c == CI && y->c == CO) {                                              \
    int r = x->w > 32 ? launch_tconv_fprop<T, CI, CO, 16>(s, x, kw, bias, y) : 0; \
    if (r == 0) r = launch_tconv_fprop<T, CI, CO, 8>(s, x, kw, bias, y);        \
    return r;                                                                  \
  }
  DNNCA_TCONV_SHAPES(X)
#undef X
  return 0;
}
template <typename T>
static int tconv_dgrad_t(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx,
                         const dnnca_tensor_t* mask, int act, float alpha) {
#define X(CI, CO)                                                                            \
  if (dx->c == CI && dy->c == CO) {                                                          \
    int r = dx->w > 32 ? launch_tconv_dgrad<T, CI, CO, 16>(s, dy, kw, dx, mask, act, alpha) : 0; \
    if (r == 0) r = launch_tconv_dgrad<T, CI, CO, 8>(s, dy, kw, dx, mask, act, alpha);        \
    return r;                                                                                \
  }
  DNNCA_TCONV_SHAPES(X)
#undef X
  return 0;
}
template <typename T>
static int tconv_wgrad_t(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk, float* db) {
#define X(CI, CO) \
  if (x->c == CI && dy->c == CO) return launch_tconv_wgrad<T, CI, CO>(s, x, dy, dk, db);
  DNNCA_TCONV_SHAPES(X)
#undef X
  return 0;
}

int try_tconv_fprop_small(cudaStream_t s, const dnnca_tensor_t* x, const float* kw, const float* bias,
                          const dnnca_tensor_t* y) {
  if (!tma_row_ok(x) || !tma_row_ok(y)) return 0;
  return x->dtype == DNNCA_F32 ? tconv_fprop_t<float>(s, x, kw, bias, y) : tconv_fprop_t<__nv_bfloat16>(s, x, kw, bias, y);
}
int try_tconv_dgrad_small(cudaStream_t s, const dnnca_tensor_t* dy, const float* kw, const dnnca_tensor_t* dx,
                          const dnnca_tensor_t* mask, int act, float alpha) {
  if (!tma_row_ok(dy) || !tma_row_ok(dx)) return 0;
  return dx->dtype == DNNCA_F32 ? tconv_dgrad_t<float>(s, dy, kw, dx, mask, act, alpha)
                                : tconv_dgrad_t<__nv_bfloat16>(s, dy, kw, dx, mask, act, alpha);
}
int try_tconv_wgrad_small(cudaStream_t s, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk, float* db) {
  if (!tma_row_ok(x) || !tma_row_ok(dy)) return 0;
  return x->dtype == DNNCA_F32 ? tconv_wgrad_t<float>(s, x, dy, dk, db) : tconv_wgrad_t<__nv_bfloat16>(s, x, dy, dk, db);
}

}  // namespace dnnca
