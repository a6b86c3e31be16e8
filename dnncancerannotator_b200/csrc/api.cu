// C-ABI entry points that dispatch between kernel families, plus error plumbing.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace dnnca {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return DNNCA_ERR_CUDA;
}

static long long g_launches = 0;
void note_launch(int n) { __atomic_fetch_add(&g_launches, (long long)n, __ATOMIC_RELAXED); }
static long long g_family[3] = {0, 0, 0};
void note_family(int f) { __atomic_fetch_add(&g_family[f], 1LL, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200; not cached so a later call can still pick the real value up
  }
  return cached;
}

// ---- TMA tensor maps ----------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

bool make_row_map(CUtensorMap* map, const dnnca_tensor_t* t, int box_chunks, int box_rows) {
  if (!t || !tma_row_ok(t) || box_chunks < 1 || box_chunks > 256 || box_rows < 1 || box_rows > 256) return false;
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  const int es = t->dtype == DNNCA_F32 ? 4 : 2;
  const cuuint64_t epc = 16 / es;
  const cuuint64_t row_bytes = (cuuint64_t)t->w * t->c * es;
  cuuint64_t dims[4] = {epc, row_bytes / 16, (cuuint64_t)t->h, (cuuint64_t)t->n};
  cuuint64_t strides[3] = {16, row_bytes, row_bytes * t->h};
  cuuint32_t box[4] = {(cuuint32_t)epc, (cuuint32_t)box_chunks, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t->data, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// kernel families (conv_generic.cu / conv_small.cu)
int launch_conv_fprop_generic(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, int, int, float);
int launch_conv_dgrad_generic(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, const dnnca_tensor_t*, int, float);
int launch_conv_wgrad_generic(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*, int);
int launch_tconv_fprop_generic(cudaStream_t, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*);
int launch_tconv_dgrad_generic(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int launch_tconv_wgrad_generic(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
int try_tconv_fprop_small(cudaStream_t, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*);
int try_tconv_dgrad_small(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int try_tconv_wgrad_small(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
// tcgen05 implicit-GEMM family (conv_umma.cu)
size_t umma_pack_bytes(int taps, int cin, int cout);
int try_conv_fprop_umma(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, int, int, float, void*, size_t, double*);
int try_conv_dgrad_umma(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, const dnnca_tensor_t*, int, float, void*, size_t);
int try_tconv_fprop_umma(cudaStream_t, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, void*, size_t, double*);
int try_tconv_dgrad_umma(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float, void*, size_t);
int conv_dgrad_umma_bnreduce(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, void*, size_t,
                             const dnnca_tensor_t*, const float*, double*, bool*);
int tconv_dgrad_umma_bnreduce(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, void*, size_t, const dnnca_tensor_t*,
                              const float*, double*, bool*);
int try_conv_wgrad_umma(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*, int);
int try_tconv_wgrad_umma(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
int prepack_conv_fprop_umma(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, int, void*, size_t);
int prepack_tconv_fprop_umma(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, void*, size_t);
// tcgen05 row-Toeplitz family for few-channel 3x3 convs (conv_row_umma.cu)
int try_conv_fprop_row(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, int, float);
int try_conv_dgrad_row(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int try_conv_wgrad_row(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
int try_tconv_fprop_row(cudaStream_t, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*);
int try_tconv_dgrad_row(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int try_tconv_wgrad_row(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
// return 1 when the shape was handled, 0 when not covered, <0 on error
int try_conv_fprop_small_f32(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, int, float, double*);
int try_conv_fprop_small_bf16(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const float*, const float*, const dnnca_tensor_t*, int, float, double*);
int try_conv_dgrad_small_f32(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int try_conv_dgrad_small_bf16(cudaStream_t, const dnnca_tensor_t*, const float*, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, int, float);
int try_conv_wgrad_small_f32(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);
int try_conv_wgrad_small_bf16(cudaStream_t, const dnnca_tensor_t*, const dnnca_tensor_t*, const dnnca_tensor_t*, float*, float*);

}  // namespace dnnca

using namespace dnnca;

static int g_force_generic = 0;
static int g_no_umma = 0;          // test hook: keep wgrad off the tensor cores (it needs no workspace to opt in)

extern "C" int dnnca_version(void) { return DNNCA_VERSION; }
extern "C" const char* dnnca_last_error(void) { return g_err; }
extern "C" int dnnca_sm_count(int* out) {
  if (!out) return DNNCA_ERR_BAD_ARG;
  *out = sm_count();
  return DNNCA_OK;
}
// kernels launched by this library since load (or since the last reset)
extern "C" long long dnnca_debug_launch_count(int reset) {
  long long v = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_launches, 0LL, __ATOMIC_RELAXED);
  return v;
}
// test hook: route every conv through the shape-generic kernels
extern "C" int dnnca_debug_force_generic(int on) {
  int old = g_force_generic;
  g_force_generic = on;
  return old;
}
// per-family launch counters: 0 = shape-generic, 1 = small-channel TMA/FFMA2, 2 = tcgen05
extern "C" long long dnnca_debug_family_count(int family, int reset) {
  if (family < 0 || family > 2) return -1;
  long long v = __atomic_load_n(&g_family[family], __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_family[family], 0LL, __ATOMIC_RELAXED);
  return v;
}

static bool act_ok(int act) { return act == DNNCA_ACT_NONE || act == DNNCA_ACT_RELU || act == DNNCA_ACT_LEAKY; }

static bool second_ok(const dnnca_tensor_t* a, const dnnca_tensor_t* b) {
  return b == nullptr || (view_ok(b) && same_nhw(a, b) && a->dtype == b->dtype);
}

extern "C" size_t dnnca_conv_workspace_bytes(int taps, int cin, int cout) {
  if (taps <= 0 || cin <= 0 || cout <= 0) return 0;
  return (umma_pack_bytes(taps, cin, cout) + 255) / 256 * 256;
}

extern "C" int dnnca_conv2d_fprop(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                                  const float* bias, const dnnca_tensor_t* y, int ksize, int act, float alpha,
                                  double* stats, void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && (w || workspace), "conv2d_fprop: bad tensor arguments");
  DNNCA_CHECK_ARG(same_nhw(x, y) && x->dtype == y->dtype, "conv2d_fprop: x and y must share n,h,w and dtype ('same' padding, stride 1)");
  DNNCA_CHECK_ARG(second_ok(x, x2), "conv2d_fprop: x2 must share n,h,w and dtype with x");
  DNNCA_CHECK_ARG(act_ok(act), "conv2d_fprop: unknown activation %d", act);
  if (ksize != 1 && ksize != 3) DNNCA_UNSUPPORTED("conv2d_fprop: kernel size %d (only 1 and 3 are used by the reference models)", ksize);
  cudaStream_t s = (cudaStream_t)stream;
  int r = 0;
  if (!w) {     // prepacked: `workspace` still holds the tensor-core packing of an earlier call with the same weights
    r = g_force_generic ? 0 : try_conv_fprop_umma(s, x, x2, nullptr, bias, y, ksize, act, alpha, workspace, workspace_bytes, stats);
    if (r < 0) return r;
    if (r == 0) DNNCA_UNSUPPORTED("conv2d_fprop: w == NULL (prepacked weights) needs a shape the tensor-core kernels serve");
    if (r == 1 && stats) return dnnca_channel_stats(stream, y, stats);
    return DNNCA_OK;
  }
  if (!g_force_generic && ksize == 3 && x->dtype == DNNCA_BF16) {
    r = try_conv_fprop_row(s, x, x2, w, bias, y, act, alpha);
    if (r < 0) return r;
    if (r == 1) return stats ? dnnca_channel_stats(stream, y, stats) : DNNCA_OK;
  }
  if (!g_force_generic && ksize == 3) {
    r = x->dtype == DNNCA_F32 ? try_conv_fprop_small_f32(s, x, x2, w, bias, y, act, alpha, stats)
                              : try_conv_fprop_small_bf16(s, x, x2, w, bias, y, act, alpha, stats);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  if (!g_force_generic) {
    r = try_conv_fprop_umma(s, x, x2, w, bias, y, ksize, act, alpha, workspace, workspace_bytes, stats);
    if (r < 0) return r;
    if (r == 2) return DNNCA_OK;                    // statistics taken in the conv epilogue
    if (r == 1) return stats ? dnnca_channel_stats(stream, y, stats) : DNNCA_OK;
  }
  r = launch_conv_fprop_generic(s, x, x2, w, bias, y, ksize, act, alpha);
  if (r != DNNCA_OK) return r;
  if (stats) return dnnca_channel_stats(stream, y, stats);
  return DNNCA_OK;
}

extern "C" int dnnca_conv2d_dgrad(void* stream, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                  const dnnca_tensor_t* dx2, int ksize, const dnnca_tensor_t* mask, int act,
                                  float alpha, void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(dz) && view_ok(dx) && w, "conv2d_dgrad: bad tensor arguments");
  DNNCA_CHECK_ARG(same_nhw(dz, dx) && dz->dtype == dx->dtype, "conv2d_dgrad: dz and dx must share n,h,w and dtype");
  DNNCA_CHECK_ARG(second_ok(dx, dx2), "conv2d_dgrad: dx2 must share n,h,w and dtype with dx");
  DNNCA_CHECK_ARG(!mask || (view_ok(mask) && same_shape(mask, dx) && mask->dtype == dx->dtype), "conv2d_dgrad: bad mask");
  DNNCA_CHECK_ARG(act_ok(act), "conv2d_dgrad: unknown activation %d", act);
  if (ksize != 1 && ksize != 3) DNNCA_UNSUPPORTED("conv2d_dgrad: kernel size %d", ksize);
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_force_generic && ksize == 3 && dx->dtype == DNNCA_BF16) {
    int r = try_conv_dgrad_row(s, dz, w, dx, dx2, mask, act, alpha);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  if (!g_force_generic && ksize == 3) {
    int r = dx->dtype == DNNCA_F32 ? try_conv_dgrad_small_f32(s, dz, w, dx, dx2, mask, act, alpha)
                                   : try_conv_dgrad_small_bf16(s, dz, w, dx, dx2, mask, act, alpha);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  if (!g_force_generic) {
    int r = try_conv_dgrad_umma(s, dz, w, dx, dx2, ksize, mask, act, alpha, workspace, workspace_bytes);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  return launch_conv_dgrad_generic(s, dz, w, dx, dx2, ksize, mask, act, alpha);
}

// dgrad whose destination `dx` is the output gradient of a BatchNormalization (input tensor `bn_x`): the BN backward
// sums are taken in the dgrad epilogue where the tensor-core halo kernel serves the layer, by bn_bwd_reduce otherwise
extern "C" int dnnca_conv2d_dgrad_bnreduce(void* stream, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                           const dnnca_tensor_t* dx2, int ksize, const dnnca_tensor_t* bn_x,
                                           const float* mean_invstd, double* sums, void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(dz) && view_ok(dx) && w && view_ok(bn_x) && mean_invstd && sums, "conv2d_dgrad_bnreduce: bad arguments");
  DNNCA_CHECK_ARG(same_nhw(dz, dx) && dz->dtype == dx->dtype, "conv2d_dgrad_bnreduce: dz and dx must share n,h,w and dtype");
  DNNCA_CHECK_ARG(second_ok(dx, dx2), "conv2d_dgrad_bnreduce: dx2 must share n,h,w and dtype with dx");
  DNNCA_CHECK_ARG(same_shape(bn_x, dx) && bn_x->dtype == dx->dtype, "conv2d_dgrad_bnreduce: bn_x must have dx's shape and dtype");
  if (ksize != 1 && ksize != 3) DNNCA_UNSUPPORTED("conv2d_dgrad_bnreduce: kernel size %d", ksize);
  cudaStream_t s = (cudaStream_t)stream;
  bool fused = false;
  int r = 0;
  if (!g_force_generic && ksize == 3 && dx->dtype == DNNCA_BF16) r = try_conv_dgrad_row(s, dz, w, dx, dx2, nullptr, DNNCA_ACT_NONE, 0.f);
  if (r == 0 && !g_force_generic && ksize == 3)
    r = dx->dtype == DNNCA_F32 ? try_conv_dgrad_small_f32(s, dz, w, dx, dx2, nullptr, DNNCA_ACT_NONE, 0.f)
                               : try_conv_dgrad_small_bf16(s, dz, w, dx, dx2, nullptr, DNNCA_ACT_NONE, 0.f);
  if (r == 0 && !g_force_generic)
    r = conv_dgrad_umma_bnreduce(s, dz, w, dx, dx2, ksize, workspace, workspace_bytes, bn_x, mean_invstd, sums, &fused);
  if (r < 0) return r;
  if (r == 0) {
    r = launch_conv_dgrad_generic(s, dz, w, dx, dx2, ksize, nullptr, DNNCA_ACT_NONE, 0.f);
    if (r != DNNCA_OK) return r;
  }
  if (!fused) return dnnca_bn_bwd_reduce(stream, bn_x, dx, mean_invstd, sums);
  return DNNCA_OK;
}

extern "C" int dnnca_convtranspose2x2_dgrad_bnreduce(void* stream, const dnnca_tensor_t* dy, const float* k, const dnnca_tensor_t* dx,
                                                     const dnnca_tensor_t* bn_x, const float* mean_invstd, double* sums,
                                                     void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(dy) && view_ok(dx) && k && view_ok(bn_x) && mean_invstd && sums, "convtranspose2x2_dgrad_bnreduce: bad arguments");
  DNNCA_CHECK_ARG(dy->n == dx->n && dy->h == 2 * dx->h && dy->w == 2 * dx->w && dx->dtype == dy->dtype,
                  "convtranspose2x2_dgrad_bnreduce: dy must be [n,2h,2w,cout]");
  DNNCA_CHECK_ARG(same_shape(bn_x, dx) && bn_x->dtype == dx->dtype, "convtranspose2x2_dgrad_bnreduce: bn_x must have dx's shape and dtype");
  cudaStream_t s = (cudaStream_t)stream;
  bool fused = false;
  int r = 0;
  if (!g_force_generic) {
    r = dx->dtype == DNNCA_BF16 ? try_tconv_dgrad_row(s, dy, k, dx, nullptr, DNNCA_ACT_NONE, 0.f) : 0;
    if (r == 0) r = try_tconv_dgrad_small(s, dy, k, dx, nullptr, DNNCA_ACT_NONE, 0.f);
    if (r == 0) r = tconv_dgrad_umma_bnreduce(s, dy, k, dx, workspace, workspace_bytes, bn_x, mean_invstd, sums, &fused);
    if (r < 0) return r;
  }
  if (r == 0) {
    r = launch_tconv_dgrad_generic(s, dy, k, dx, nullptr, DNNCA_ACT_NONE, 0.f);
    if (r != DNNCA_OK) return r;
  }
  if (!fused) return dnnca_bn_bwd_reduce(stream, bn_x, dx, mean_invstd, sums);
  return DNNCA_OK;
}

extern "C" int dnnca_conv2d_wgrad(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2,
                                  const dnnca_tensor_t* dz, float* dw, float* db, int ksize) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(dz) && dw, "conv2d_wgrad: bad tensor arguments");
  DNNCA_CHECK_ARG(same_nhw(x, dz) && x->dtype == dz->dtype, "conv2d_wgrad: x and dz must share n,h,w and dtype");
  DNNCA_CHECK_ARG(second_ok(x, x2), "conv2d_wgrad: x2 must share n,h,w and dtype with x");
  if (ksize != 1 && ksize != 3) DNNCA_UNSUPPORTED("conv2d_wgrad: kernel size %d", ksize);
  cudaStream_t s = (cudaStream_t)stream;
  if (!g_force_generic && !g_no_umma && ksize == 3 && x->dtype == DNNCA_BF16) {
    int r = try_conv_wgrad_row(s, x, x2, dz, dw, db);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  if (!g_force_generic && ksize == 3) {
    int r = x->dtype == DNNCA_F32 ? try_conv_wgrad_small_f32(s, x, x2, dz, dw, db)
                                  : try_conv_wgrad_small_bf16(s, x, x2, dz, dw, db);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  if (!g_force_generic && !g_no_umma) {
    int r = try_conv_wgrad_umma(s, x, x2, dz, dw, db, ksize);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  return launch_conv_wgrad_generic(s, x, x2, dz, dw, db, ksize);
}

extern "C" int dnnca_convtranspose2x2_fprop(void* stream, const dnnca_tensor_t* x, const float* k, const float* bias,
                                            const dnnca_tensor_t* y, double* stats, void* workspace,
                                            size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && (k || workspace), "convtranspose2x2_fprop: bad tensor arguments");
  DNNCA_CHECK_ARG(y->n == x->n && y->h == 2 * x->h && y->w == 2 * x->w && x->dtype == y->dtype,
                  "convtranspose2x2_fprop: y must be [n,2h,2w,cout] with x's dtype");
  if (!k) {     // prepacked (see conv2d_fprop)
    int rp = g_force_generic ? 0 : try_tconv_fprop_umma((cudaStream_t)stream, x, nullptr, bias, y, workspace, workspace_bytes, stats);
    if (rp < 0) return rp;
    if (rp == 0) DNNCA_UNSUPPORTED("convtranspose2x2_fprop: k == NULL (prepacked weights) needs a shape the tensor-core kernels serve");
    return (stats && rp != 2) ? dnnca_channel_stats(stream, y, stats) : DNNCA_OK;
  }
  int r = (g_force_generic || x->dtype != DNNCA_BF16) ? 0 : try_tconv_fprop_row((cudaStream_t)stream, x, k, bias, y);
  if (r == 0 && !g_force_generic) r = try_tconv_fprop_small((cudaStream_t)stream, x, k, bias, y);
  if (r == 0 && !g_force_generic) r = try_tconv_fprop_umma((cudaStream_t)stream, x, k, bias, y, workspace, workspace_bytes, stats);
  if (r < 0) return r;
  const bool stats_taken = r == 2;
  if (r == 0) r = launch_tconv_fprop_generic((cudaStream_t)stream, x, k, bias, y);
  else r = DNNCA_OK;
  if (r != DNNCA_OK) return r;
  if (stats && !stats_taken) return dnnca_channel_stats(stream, y, stats);
  return DNNCA_OK;
}

extern "C" int dnnca_convtranspose2x2_dgrad(void* stream, const dnnca_tensor_t* dy, const float* k,
                                            const dnnca_tensor_t* dx, const dnnca_tensor_t* mask, int act,
                                            float alpha, void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(dy) && view_ok(dx) && k, "convtranspose2x2_dgrad: bad tensor arguments");
  DNNCA_CHECK_ARG(dy->n == dx->n && dy->h == 2 * dx->h && dy->w == 2 * dx->w && dx->dtype == dy->dtype,
                  "convtranspose2x2_dgrad: dy must be [n,2h,2w,cout]");
  DNNCA_CHECK_ARG(!mask || (view_ok(mask) && same_shape(mask, dx) && mask->dtype == dx->dtype), "convtranspose2x2_dgrad: bad mask");
  DNNCA_CHECK_ARG(act_ok(act), "convtranspose2x2_dgrad: unknown activation %d", act);
  if (!g_force_generic) {
    int r = dx->dtype == DNNCA_BF16 ? try_tconv_dgrad_row((cudaStream_t)stream, dy, k, dx, mask, act, alpha) : 0;
    if (r == 0) r = try_tconv_dgrad_small((cudaStream_t)stream, dy, k, dx, mask, act, alpha);
    if (r == 0) r = try_tconv_dgrad_umma((cudaStream_t)stream, dy, k, dx, mask, act, alpha, workspace, workspace_bytes);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  return launch_tconv_dgrad_generic((cudaStream_t)stream, dy, k, dx, mask, act, alpha);
}

extern "C" int dnnca_convtranspose2x2_wgrad(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy,
                                            float* dk, float* db) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(dy) && dk, "convtranspose2x2_wgrad: bad tensor arguments");
  DNNCA_CHECK_ARG(dy->n == x->n && dy->h == 2 * x->h && dy->w == 2 * x->w && x->dtype == dy->dtype,
                  "convtranspose2x2_wgrad: dy must be [n,2h,2w,cout]");
  if (!g_force_generic) {
    int r = (x->dtype == DNNCA_BF16 && !g_no_umma) ? try_tconv_wgrad_row((cudaStream_t)stream, x, dy, dk, db) : 0;
    if (r == 0) r = try_tconv_wgrad_small((cudaStream_t)stream, x, dy, dk, db);
    if (r == 0 && !g_no_umma) r = try_tconv_wgrad_umma((cudaStream_t)stream, x, dy, dk, db);
    if (r < 0) return r;
    if (r == 1) return DNNCA_OK;
  }
  return launch_tconv_wgrad_generic((cudaStream_t)stream, x, dy, dk, db);
}

// ---- pinned host staging buffers (input tail: data.py:110 prefetch -> H2D) ------------------------------------------
extern "C" int dnnca_host_alloc(size_t bytes, int write_combined, void** out) {
  DNNCA_CHECK_ARG(out && bytes > 0, "host_alloc: bad arguments");
  cudaError_t e = cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  if (e != cudaSuccess) return cuda_fail(e, "host_alloc: cudaHostAlloc");
  return DNNCA_OK;
}
extern "C" int dnnca_host_free(void* p) {
  if (!p) return DNNCA_OK;
  cudaError_t e = cudaFreeHost(p);
  if (e != cudaSuccess) return cuda_fail(e, "host_free: cudaFreeHost");
  return DNNCA_OK;
}

// ---- weight packing for prepacked (w == NULL) inference calls ---------------------------------------------------------
extern "C" int dnnca_conv2d_prepack(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                                    const dnnca_tensor_t* y, int ksize, void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && w && workspace && second_ok(x, x2), "conv2d_prepack: bad arguments");
  if (ksize != 1 && ksize != 3) DNNCA_UNSUPPORTED("conv2d_prepack: kernel size %d", ksize);
  const int r = g_force_generic ? 0 : prepack_conv_fprop_umma((cudaStream_t)stream, x, x2, w, y, ksize, workspace, workspace_bytes);
  if (r < 0) return r;
  if (r == 0) DNNCA_UNSUPPORTED("conv2d_prepack: this shape is not served by the tensor-core kernels (pass w to conv2d_fprop instead)");
  return DNNCA_OK;
}
extern "C" int dnnca_convtranspose2x2_prepack(void* stream, const dnnca_tensor_t* x, const float* k, const dnnca_tensor_t* y,
                                              void* workspace, size_t workspace_bytes) {
  DNNCA_CHECK_ARG(view_ok(x) && view_ok(y) && k && workspace, "convtranspose2x2_prepack: bad arguments");
  const int r = g_force_generic ? 0 : prepack_tconv_fprop_umma((cudaStream_t)stream, x, k, y, workspace, workspace_bytes);
  if (r < 0) return r;
  if (r == 0) DNNCA_UNSUPPORTED("convtranspose2x2_prepack: this shape is not served by the tensor-core kernels");
  return DNNCA_OK;
}
