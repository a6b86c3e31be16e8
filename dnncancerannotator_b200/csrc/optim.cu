// Keras-form Adam over the flat parameter buffer (engine.py:276-284).
// One launch updates every variable of the model (the reference runs one
// ResourceApplyAdam per variable); the step counter lives on the device so the
// whole training step can be replayed from a CUDA graph.
#include "common.cuh"

namespace dnnca {

__global__ void adam_tick_kernel(long long* step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ m, float* __restrict__ v, long long count,
                                                  const float* __restrict__ hyper, const long long* __restrict__ step,
                                                  const float* __restrict__ l2) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const double t = (double)(*step);
  // lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)   [TF-semantics: keras Adam._prepare_local]
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    const float pi = p[i];
    if (l2) gi = fmaf(2.f * l2[i], pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - lr_t * mi / (sqrtf(vi) + eps);
  }
}

// loss scalar of one step, on the device (keras Model.train_step: compiled loss + regulariser losses, both taken at
// the weights the forward pass used): out[0] += scale * (mean_b per_sample[b] + sum_i l2[i]*p[i]^2)
__global__ void __launch_bounds__(256) loss_total_kernel(const float* __restrict__ per_sample, int batch,
                                                        const float* __restrict__ p, const float* __restrict__ l2,
                                                        long long count, float scale, float* __restrict__ out) {
  float acc = 0.f;
  if (l2)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
      acc = fmaf(l2[i] * p[i], p[i], acc);
  if (blockIdx.x == 0) {
    float s = 0.f;
    for (int b = threadIdx.x; b < batch; b += blockDim.x) s += per_sample[b];
    acc += s / (float)batch;
  }
  acc = warp_sum(acc);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    if (t != 0.f || blockIdx.x == 0) atomicAdd(out, scale * t);
  }
}

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_loss_total(void* stream, const float* per_sample, int batch, const float* params, const float* l2,
                                int64_t count, float scale, float* out) {
  DNNCA_CHECK_ARG(per_sample && batch > 0 && out && (!l2 || (params && count > 0)), "loss_total: bad arguments");
  const int grid = l2 ? grid_for(count, 256 * 8, 2) : 1;
  loss_total_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(per_sample, batch, params, l2, count, scale, out);
  DNNCA_LAUNCH_CHECK("loss_total");
  return DNNCA_OK;
}

extern "C" int dnnca_adam_step(void* stream, float* params, const float* grads, float* m, float* v, int64_t count,
                               const float* hyper, int64_t* step, const float* l2) {
  DNNCA_CHECK_ARG(params && grads && m && v && hyper && step && count > 0, "adam_step: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(step));
  note_launch(1);
  int grid = grid_for(count, 256 * 4, 4);
  adam_kernel<<<grid, 256, 0, s>>>(params, grads, m, v, count, hyper, reinterpret_cast<const long long*>(step), l2);
  DNNCA_LAUNCH_CHECK("adam_step");
  return DNNCA_OK;
}
