// Helpers shared by the small-channel TMA/FFMA2 kernels (conv_small.cu, tconv_small.cu).
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace dnnca {

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo32(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi32(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

constexpr int PX = 4;  // pixels per thread along x (two FFMA2 pixel pairs)
__host__ __device__ constexpr int ru(int a, int b) { return (a + b - 1) / b * b; }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// geometry of one raw TMA tile row: C interleaved channels, TW pixels + HALO pixels each side
template <typename T, int C, int TW, int HALO>
struct Raw {
  static constexpr int EPC = 16 / (int)sizeof(T);                         // elements per 16-byte chunk
  static constexpr int OFF = HALO ? (((-C) % EPC + EPC) % EPC) : 0;       // (x0-1)*C mod EPC, x0 % EPC == 0
  static constexpr int NCH = C ? (OFF + (TW + 2 * HALO) * C + EPC - 1) / EPC : 0;
  __device__ static int chunk_start(int x0) { return ((x0 - HALO) * C - OFF) / EPC; }
};

// raw tile [rows][NCH*EPC] (T, interleaved) -> planes dst[ci][rows][pitch] (fp32); one thread per pixel
// `shift` != 0: a second copy of every plane displaced by one pixel (dst1[.. col] = dst[.. col+1]) is written `shift`
// floats behind the first, so that the odd pixel pairs (c1,c2),(c3,c4) of a 3-tap window are naturally aligned
// 64-bit shared-memory loads instead of register moves (IMAD.MOV would compete with FFMA2 for the FMA pipe).
template <typename T, int C, int TW, int HALO>
__device__ __forceinline__ void deinterleave(const T* __restrict__ raw, float* __restrict__ dst, int rows, int pitch,
                                             int shift = 0) {
  using G = Raw<T, C, TW, HALO>;
  constexpr int COLS = TW + 2 * HALO;
  constexpr int RP = G::NCH * G::EPC;
  constexpr bool WORDS = sizeof(T) == 2 && (G::OFF % 2 == 0) && (C % 2 == 0);
  for (int e = threadIdx.x; e < rows * COLS; e += 256) {
    const int row = e / COLS, col = e - row * COLS;
    const T* src = raw + row * RP + G::OFF + col * C;
    float* d = dst + row * pitch + col;
    float v[C];
    if (WORDS) {                                     // bf16 pairs: one 32-bit LDS per two channels
      const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
#pragma unroll
      for (int j = 0; j < C / 2; ++j) {
        const uint32_t wd = s32[j];
        v[2 * j] = __uint_as_float(wd << 16);
        v[2 * j + 1] = __uint_as_float(wd & 0xffff0000u);
      }
    } else {
#pragma unroll
      for (int j = 0; j < C; ++j) v[j] = ldf(src + j);
    }
#pragma unroll
    for (int j = 0; j < C; ++j) d[j * rows * pitch] = v[j];
    if (shift && col > 0) {
#pragma unroll
      for (int j = 0; j < C; ++j) d[shift - 1 + j * rows * pitch] = v[j];
    }
  }
}

}  // namespace dnnca
