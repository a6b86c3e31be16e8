// Micro-benchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed FFMA2 (fma.rn.f32x2),
// with and without interleaved broadcast LDS.128 (the conv inner-loop mix).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 pk(float x, float y) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(x), "f"(y)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  __shared__ __align__(16) float ws[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) ws[i] = in[i];
  __syncthreads();
  float x = in[threadIdx.x], y = in[threadIdx.x + 256];
  if (MODE == 0) {            // scalar FFMA, 32 independent accumulators, register operands
    float a[32];
    for (int j = 0; j < 32; ++j) a[j] = j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = fmaf(a[j], x, y);
    float s = 0; for (int j = 0; j < 32; ++j) s += a[j];
    out[blockIdx.x * 256 + threadIdx.x] = s;
  } else if (MODE == 1) {     // FFMA2, 32 independent 2-wide accumulators
    u64 a[32]; u64 xx = pk(x, x), yy = pk(y, y);
    for (int j = 0; j < 32; ++j) a[j] = pk(j, j + 1);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = ffma2(a[j], xx, yy);
    u64 s = a[0]; for (int j = 1; j < 32; ++j) s = ffma2(s, pk(1.f, 1.f), a[j]);
    out[blockIdx.x * 256 + threadIdx.x] = __uint_as_float((unsigned)s) + __uint_as_float((unsigned)(s >> 32));
  } else if (MODE == 2) {     // scalar FFMA fed by broadcast LDS.128 weights: 16 FFMA per LDS.128 (4 px x 4 co)
    float a[4][4]; for (int i = 0; i < 16; ++i) a[i / 4][i % 4] = i;
    float w4[4] = {x, y, x + 1, y + 1};
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 wv = *reinterpret_cast<const float4*>(ws + ((it * 8 + j) & 255) * 4);
        float wa[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int p = 0; p < 4; ++p) a[c][p] = fmaf(w4[p], wa[c], a[c][p]);
      }
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i / 4][i % 4];
    out[blockIdx.x * 256 + threadIdx.x] = s;
  } else {                    // FFMA2 fed by LDS.128 of duplicated weight pairs: 4 FFMA2 (8 FMA) per LDS.128... x2 co
    u64 a[4][2]; for (int i = 0; i < 8; ++i) a[i / 2][i % 2] = pk(i, i);
    u64 p0 = pk(x, y), p1 = pk(y, x);
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(ws + ((it * 8 + j) & 255) * 4);
        a[0][0] = ffma2(p0, wv.x, a[0][0]); a[0][1] = ffma2(p1, wv.x, a[0][1]);
        a[1][0] = ffma2(p0, wv.y, a[1][0]); a[1][1] = ffma2(p1, wv.y, a[1][1]);
        a[2][0] = ffma2(p1, wv.x, a[2][0]); a[2][1] = ffma2(p0, wv.y, a[2][1]);
        a[3][0] = ffma2(p1, wv.y, a[3][0]); a[3][1] = ffma2(p0, wv.x, a[3][1]);
      }
    u64 s = a[0][0]; for (int i = 1; i < 8; ++i) s = ffma2(s, pk(1.f, 1.f), a[i / 2][i % 2]);
    out[blockIdx.x * 256 + threadIdx.x] = __uint_as_float((unsigned)s) + __uint_as_float((unsigned)(s >> 32));
  }
}

template <int MODE> void run(const char* name, double fma_per_thread_iter) {
  float *out, *in; int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  k<MODE><<<blocks, 256>>>(out, in, 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = fma_per_thread_iter * iters * blocks * 256.0;
  printf("%-44s %8.3f ms  %7.2f TFLOP/s\n", name, ms, 2 * fma / ms / 1e9);
}
int main() {
  run<0>("scalar FFMA (reg operands)", 32);
  run<1>("FFMA2 (reg operands)", 64);
  run<2>("scalar FFMA + LDS.128 broadcast (16:1)", 128);
  run<3>("FFMA2 + LDS.128 broadcast (8 FFMA2:1)", 128);
  return 0;
}
