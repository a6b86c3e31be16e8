// Gradient all-reduce FUSED into the Adam step over NVLink peer memory (one box, one process per GPU).
//
// The reference's exchange step is MirroredStrategy's SUM all-reduce of all gradients followed by one Adam update per
// variable (engine.py:260-263, 276-284).  For configs/unet.yaml the whole gradient is 35 KB: a separate NCCL all-reduce is
// pure latency (~45 us exposed at the end of a 1.9 ms step).  Here every rank's flat gradient buffer lives in peer-mapped
// memory (cudaMalloc + CUDA IPC, NVLink 5 / NVSwitch) and ONE kernel per rank
//   1. publishes "my gradients of step e are complete" by storing e into its slot of every peer's flag array
//      (st.release.sys over NVLink) and waits until all peers' slots in its OWN flag array reached e,
//   2. reads element i of every rank's gradient buffer straight from peer memory, sums them in rank order (so all
//      replicas compute bit-identical sums -- mirrored variables stay mirrored), and applies the Keras-form Adam update to
//      its own copy of the parameters (optim.cu's arithmetic),
//   3. after its last read, publishes "done reading step e" the same way; the next step's first kernel (p2p_wait_done)
//      holds the gradient zeroing back until every peer is done reading.
// No NCCL call, no extra pass over the gradients, no staging copy.  One-shot all-to-all reads cost (N-1) x bytes per
// rank, which is the right trade for small models (latency-bound; measured up to mulmo_unet's 6.9 MB: 5.80 vs 6.00 ms/step on
// two GPUs against the overlapped NCCL buckets, whose CTAs take SMs from the persistent conv kernels); models with tens of MB of gradients keep the bucketed
// NCCL path overlapped with the backward pass (parallel.py).  Spin loops carry a timeout and raise a sticky error flag
// instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace dnnca {

struct P2PPeers {
  const float* grads[8];          // every rank's flat gradient buffer (own rank: local pointer)
  unsigned long long* flags[8];   // every rank's flag array: [0..8) "gradients ready" slots, [8..16) "done reading" slots, [16] error
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// waits until slots [base, base+world) of the LOCAL flag array are >= epoch; false on timeout
__device__ __forceinline__ bool wait_slots(unsigned long long* local, int base, int world, unsigned long long epoch) {
  for (int r = 0; r < world; ++r) {
    long long spins = 0;
    while (ld_acquire_sys(local + base + r) < epoch) {
      if (++spins > (1LL << 26)) { atomicExch(local + 16, 1ULL); return false; }     // ~ seconds: a peer died or never launched
      __nanosleep(64);
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) p2p_adam_kernel(P2PPeers peers, int rank, int world, float* __restrict__ p,
                                                      float* __restrict__ m, float* __restrict__ v, long long count,
                                                      long long reduce_count, float* __restrict__ reduced_out,
                                                      const float* __restrict__ hyper, const long long* __restrict__ step,
                                                      const long long* __restrict__ p2p_epoch, const float* __restrict__ l2,
                                                      unsigned int* __restrict__ done_blocks) {
  const unsigned long long epoch = (unsigned long long)(*p2p_epoch); // exchange counter, incremented by the tick kernel just before
  unsigned long long* local = peers.flags[rank];
  __shared__ int ok;
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();                                        // this rank's gradients (earlier kernels) are visible system-wide
      for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + rank, epoch);
    }
    ok = wait_slots(local, 0, world, epoch) ? 1 : 0;
  }
  __syncthreads();
  if (ok) {
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
    const double t = (double)(*step);                                // Adam's own iteration count (survives checkpoint resume)
    const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, t)) / (1.0 - pow((double)b1, t)));
    auto adam1 = [&](long long i, float g) {
      const float pi = p[i];
      float gi = g;
      if (l2) gi = fmaf(2.f * l2[i], pi, gi);
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] = pi - lr_t * mi / (sqrtf(vi) + eps);
    };
    // 16-byte peer reads (every buffer is a cudaMalloc base: aligned); sums in rank order: identical on every replica
    const long long nvec = reduce_count / 4;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nvec; q += (long long)gridDim.x * blockDim.x) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 t = __ldcv(reinterpret_cast<const float4*>(peers.grads[r]) + q);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
      }
      const long long i = q * 4;
      if (i < count) adam1(i, g.x);
      if (i + 1 < count) adam1(i + 1, g.y);
      if (i + 2 < count) adam1(i + 2, g.z);
      if (i + 3 < count) adam1(i + 3, g.w);
      // summed gradients + loss scalar for the host (get_grads, reported loss); a LOCAL buffer: the peers are still
      // reading this rank's raw gradients
      if (reduced_out) reinterpret_cast<float4*>(reduced_out)[q] = g;
    }
    for (long long i = nvec * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < reduce_count; i += (long long)gridDim.x * blockDim.x) {
      float g = 0.f;
      for (int r = 0; r < world; ++r) g += __ldcv(peers.grads[r] + i);
      if (i < count) adam1(i, g);
      if (reduced_out) reduced_out[i] = g;
    }
  }
  // last block out: every read of peer memory by this rank has completed -> tell the peers
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int prev = atomicAdd(done_blocks, 1u);
    if (prev == gridDim.x - 1) {
      *done_blocks = 0;
      for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + 8 + rank, epoch);
    }
  }
}

// first kernel of a step: nobody may still be reading this rank's gradients of the previous step when they are zeroed
__global__ void p2p_wait_done_kernel(unsigned long long* local, int world, const long long* __restrict__ p2p_epoch) {
  const unsigned long long epoch = (unsigned long long)(*p2p_epoch); // value BEFORE this step's tick = the previous exchange
  if (threadIdx.x == 0 && epoch > 0) wait_slots(local, 8, world, epoch);
}

__global__ void adam_tick2_kernel(long long* step, long long* p2p_epoch) { *step += 1; *p2p_epoch += 1; }

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_p2p_alloc(size_t bytes, void** out) {
  DNNCA_CHECK_ARG(out && bytes > 0, "p2p_alloc: bad arguments");
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "p2p_alloc: cudaMalloc");
  e = cudaMemset(*out, 0, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "p2p_alloc: cudaMemset");
  return DNNCA_OK;
}
extern "C" int dnnca_p2p_free(void* p) {
  if (!p) return DNNCA_OK;
  cudaError_t e = cudaFree(p);
  return e == cudaSuccess ? DNNCA_OK : cuda_fail(e, "p2p_free");
}
extern "C" int dnnca_p2p_export(void* p, unsigned char* handle64) {
  DNNCA_CHECK_ARG(p && handle64, "p2p_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return cuda_fail(e, "p2p_export: cudaIpcGetMemHandle");
  memcpy(handle64, &h, 64);
  return DNNCA_OK;
}
extern "C" int dnnca_p2p_import(const unsigned char* handle64, void** out) {
  DNNCA_CHECK_ARG(handle64 && out, "p2p_import: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "p2p_import: cudaIpcOpenMemHandle");
  return DNNCA_OK;
}
extern "C" int dnnca_p2p_close(void* p) {
  if (!p) return DNNCA_OK;
  cudaError_t e = cudaIpcCloseMemHandle(p);
  return e == cudaSuccess ? DNNCA_OK : cuda_fail(e, "p2p_close");
}

extern "C" int dnnca_p2p_wait_done(void* stream, void* local_flags, int world, const int64_t* p2p_epoch) {
  DNNCA_CHECK_ARG(local_flags && p2p_epoch && world >= 1 && world <= 8, "p2p_wait_done: bad arguments");
  p2p_wait_done_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(local_flags), world,
                                                           reinterpret_cast<const long long*>(p2p_epoch));
  DNNCA_LAUNCH_CHECK("p2p_wait_done");
  return DNNCA_OK;
}

extern "C" int dnnca_p2p_adam_step(void* stream, const void* const* peer_grads, void* const* peer_flags, int rank, int world,
                                   float* params, float* m, float* v, int64_t count, int64_t reduce_count, float* reduced_out,
                                   const float* hyper, int64_t* step, int64_t* p2p_epoch, const float* l2,
                                   unsigned int* done_blocks) {
  DNNCA_CHECK_ARG(peer_grads && peer_flags && params && m && v && hyper && step && p2p_epoch && done_blocks,
                  "p2p_adam_step: bad arguments");
  DNNCA_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world && count > 0 && reduce_count >= count,
                  "p2p_adam_step: bad rank / world / counts");
  P2PPeers pp{};
  for (int r = 0; r < world; ++r) {
    DNNCA_CHECK_ARG(peer_grads[r] && peer_flags[r], "p2p_adam_step: missing peer pointer %d", r);
    pp.grads[r] = reinterpret_cast<const float*>(peer_grads[r]);
    pp.flags[r] = reinterpret_cast<unsigned long long*>(peer_flags[r]);
  }
  cudaStream_t s = (cudaStream_t)stream;
  adam_tick2_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(step), reinterpret_cast<long long*>(p2p_epoch));
  note_launch(1);
  int grid = grid_for((reduce_count + 3) / 4, 256, 2);
  if (grid > sm_count()) grid = sm_count();          // every block spins on the flags first: keep the grid one resident wave
  p2p_adam_kernel<<<grid, 256, 0, s>>>(pp, rank, world, params, m, v, count, reduce_count, reduced_out, hyper,
                                       reinterpret_cast<const long long*>(step), reinterpret_cast<const long long*>(p2p_epoch), l2,
                                       done_blocks);
  DNNCA_LAUNCH_CHECK("p2p_adam_step");
  return DNNCA_OK;
}
