// Pixel-threshold confusion counters on the device: the statistics behind tf.keras.metrics.Precision /
// Recall / AUC / the reference's FBetaScore (annotator/utils/metrics.py:37-77) as configured by
// configs/additionals/metrics.yaml:1-23 and attached to the model at engine.py:273.
//
// Keras counts, for every threshold t_k, TP_k = #(y != 0 and p > t_k) and FP_k = #(y == 0 and p > t_k).  With ascending
// thresholds this is a suffix sum of a histogram over b(p) = #{k : t_k < p}, so ONE pass over (p, y) with two histograms
// (positives / negatives) of nthr+1 bins serves every threshold-based metric; the host turns bins into TP/FP/FN/TN.
// Memory-bound: 8 bytes per pixel, 16-byte loads, warp-aggregated shared-memory atomics (most pixels of a segmentation
// map fall into the same bin), one 64-bit global atomic per bin and block at the end.
#include "common.cuh"

namespace dnnca {

constexpr int METRIC_MAX_THR = 1023;

__device__ __forceinline__ int metric_bin(float p, const float* __restrict__ thr, int nthr) {
  // number of thresholds strictly below p (binary search; exact float comparisons like `y_pred > thresholds`)
  int lo = 0, hi = nthr;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (thr[mid] < p) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) threshold_hist_kernel(const float* __restrict__ probs, const float* __restrict__ labels,
                                                            long long count, const float* __restrict__ thresholds, int nthr,
                                                            unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];            // [nthr] thresholds (as float) | [2*(nthr+1)] bins
  float* sthr = reinterpret_cast<float*>(sh);
  unsigned int* sbin = sh + nthr;
  const int nb = nthr + 1;
  for (int i = threadIdx.x; i < nthr; i += blockDim.x) sthr[i] = thresholds[i];
  for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) sbin[i] = 0u;
  __syncthreads();
  const long long nvec = count / 4;
  const float4* p4 = reinterpret_cast<const float4*>(probs);
  const float4* y4 = reinterpret_cast<const float4*>(labels);
  auto add = [&](float p, float y) {
    const int slot = (y != 0.f ? 0 : nb) + metric_bin(p, sthr, nthr);
    // warp-aggregated increment: lanes with the same slot elect one leader
    const unsigned peers = __match_any_sync(__activemask(), slot);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(sbin + slot, (unsigned)__popc(peers));
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(p4 + i), y = __ldg(y4 + i);
    add(p.x, y.x); add(p.y, y.y); add(p.z, y.z); add(p.w, y.w);
  }
  for (long long i = nvec * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    add(probs[i], labels[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x)
    if (sbin[i]) atomicAdd(hist + i, (unsigned long long)sbin[i]);
}

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_threshold_hist(void* stream, const float* probs, const float* labels, int64_t count, const float* thresholds,
                                    int nthr, uint64_t* hist) {
  DNNCA_CHECK_ARG(probs && labels && thresholds && hist && count > 0, "threshold_hist: bad arguments");
  DNNCA_CHECK_ARG(nthr >= 1 && nthr <= METRIC_MAX_THR, "threshold_hist: 1..%d thresholds supported (got %d)", METRIC_MAX_THR, nthr);
  DNNCA_CHECK_ARG(((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(labels)) & 15) == 0,
                  "threshold_hist: probs and labels must be 16-byte aligned");
  // a block sees at most 2^32 - 1 pixels (32-bit shared bins): cap the per-block share at 2^31
  long long blocks = (count / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (count / blocks > (1LL << 31)) DNNCA_UNSUPPORTED("threshold_hist: %lld pixels per launch exceed the 32-bit block counters", (long long)count);
  const size_t smem = (size_t)(nthr + 2 * (nthr + 1)) * 4;
  threshold_hist_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(probs, labels, count, thresholds, nthr,
                                                                         reinterpret_cast<unsigned long long*>(hist));
  DNNCA_LAUNCH_CHECK("threshold_hist");
  return DNNCA_OK;
}
