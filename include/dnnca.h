/*
 * dnnca.h -- C ABI of libdnnca.so: the B200 (sm_100a) kernels behind the
 * DNNCancerAnnotator conv-stack hot path.
 *
 * The reference (yoshihikoueno/DNNCancerAnnotator) has no FFI of its own: it is
 * pure Python on tf.keras, and every entry point below replaces the TensorFlow
 * kernel(s) reached through one Keras call site of the reference.  Each
 * declaration cites that call site (file:line under the reference root).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *  - Every function returns 0 on success or a negative dnnca_status_t;
 *    dnnca_last_error() returns a thread-local message for the last failure.
 *  - Every launch is asynchronous on the cudaStream_t passed as `stream`
 *    (a void*; NULL = legacy default stream).  The library allocates no device
 *    memory and never synchronises; the caller owns all buffers and workspaces.
 *  - There is no CPU fallback and no backend dispatch: without a CUDA device the
 *    launches fail with DNNCA_ERR_CUDA.
 *  - Activations are NHWC.  A dnnca_tensor_t is a *channel-slice view* of an
 *    NHWC buffer: element (n,y,x,ch) lives at
 *        data + ((n*h + y)*w + x)*cstride + coff + ch     (in elements of dtype)
 *    so a producer can write its channels straight into a concat buffer
 *    (tf.concat at components.py:164 and unet.py:187 needs no copy kernel).
 *  - Weights keep the reference's TensorFlow layouts, fp32 masters:
 *    Conv2D kernel [kh,kw,Cin,Cout] (HWIO), Conv2DTranspose kernel [kh,kw,Cout,Cin].
 *  - Parameter gradients, statistics and the loss are fp32 (fp64 accumulators
 *    where noted); max-pool argmax is one uint8 (0..3, row-major in the window).
 */
#ifndef DNNCA_H_
#define DNNCA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DNNCA_VERSION 100 /* 0.1.0 */
#define DNNCA_API __attribute__((visibility("default")))

typedef enum {
  DNNCA_OK = 0,
  DNNCA_ERR_BAD_ARG = -1,
  DNNCA_ERR_UNSUPPORTED = -2,
  DNNCA_ERR_CUDA = -3,
  DNNCA_ERR_NCCL = -4
} dnnca_status_t;

typedef enum { DNNCA_F32 = 0, DNNCA_BF16 = 1 } dnnca_dtype_t;

/* components.py:323-335 solve_activation: 'relu' and LeakyReLU(alpha) (leakyReLU.yaml:1-4) */
typedef enum { DNNCA_ACT_NONE = 0, DNNCA_ACT_RELU = 1, DNNCA_ACT_LEAKY = 2 } dnnca_act_t;

typedef struct dnnca_tensor {
  void* data;      /* device pointer to element (0,0,0, channel 0 of the buffer) */
  int32_t n, h, w; /* logical extent of the view */
  int32_t c;       /* channels in the view */
  int32_t cstride; /* channels per pixel in the underlying buffer (>= coff + c) */
  int32_t coff;    /* first channel of the view inside the buffer */
  int32_t dtype;   /* dnnca_dtype_t */
} dnnca_tensor_t;

DNNCA_API int dnnca_version(void);
DNNCA_API const char* dnnca_last_error(void);
/* Number of SMs of the current device (grid sizing is a multiple of it). */
DNNCA_API int dnnca_sm_count(int* out);
/* test hook: non-zero routes every conv through the shape-generic kernels; returns the old value */
DNNCA_API int dnnca_debug_force_generic(int on);
/* number of kernels this library has launched since load (reset != 0 zeroes the counter afterwards) */
DNNCA_API long long dnnca_debug_launch_count(int reset);
/* conv / ConvT launches per kernel family since load: 0 = shape-generic CUDA-core, 1 = small-channel
 * TMA + FFMA2, 2 = tcgen05 implicit GEMM (tests assert which family served a shape) */
DNNCA_API long long dnnca_debug_family_count(int family, int reset);

/* ---------------------------------------------------------------------------
 * Workspace of the tensor-core (tcgen05) kernels: room for the bf16 K-major repack of one layer's
 * weights, `taps * cin * cout * 2` bytes (taps = k*k, or 4 for ConvT).  The caller owns it (one per
 * layer or one shared scratch on a single stream); every fprop/dgrad call re-packs from the fp32
 * masters, so nothing in it is state.  workspace == NULL selects the CUDA-core kernels.
 * Inference with unchanged weights (model.evaluate / predict sweeps, engine.py:198-203, 222): an fprop call with
 * w == NULL (k == NULL for ConvT) re-uses the packing the SAME layer's previous fprop call left in `workspace`
 * (caller's contract: a per-layer workspace nothing else wrote); it fails with DNNCA_ERR_UNSUPPORTED when the
 * shape is not served by the tensor-core kernels.  dnnca_conv2d_prepack / dnnca_convtranspose2x2_prepack write
 * exactly that packing (no convolution is launched): call them once per weight version, then fprop with w == NULL.
 * ------------------------------------------------------------------------- */
DNNCA_API size_t dnnca_conv_workspace_bytes(int taps, int cin, int cout);
DNNCA_API int dnnca_conv2d_prepack(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                                   const dnnca_tensor_t* y, int ksize, void* workspace, size_t workspace_bytes);
DNNCA_API int dnnca_convtranspose2x2_prepack(void* stream, const dnnca_tensor_t* x, const float* k, const dnnca_tensor_t* y,
                                             void* workspace, size_t workspace_bytes);

/* ---------------------------------------------------------------------------
 * Conv2D, stride 1, 'same' zero padding, k in {1,3}
 *   replaces layers.Conv2D at components.py:47-50, components.py:123-126,
 *   multiresunet.py:51-52 (use_bias=False -> bias == NULL).
 * y = act([x | x2] (*) w + bias)  (cross-correlation).  `x2` (may be NULL) is a
 * second input whose channels follow x's: the kernel reads the two producers of
 * tf.concat([tconv0, cropped], -1) (components.py:164) directly, so the concat is
 * never materialised; w is [k,k,Cx+Cx2,Cout].  If `stats` != NULL the per-channel
 * sum and sum of squares of the *stored* output are ACCUMULATED into
 * stats[0..C) and stats[C..2C) (fp64; caller zeroes) for the BatchNormalization
 * that follows (components.py:57-58, 130-132).
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_conv2d_fprop(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* w,
                                 const float* bias, const dnnca_tensor_t* y, int ksize, int act, float alpha,
                                 double* stats, void* workspace, size_t workspace_bytes);

/* Gradient w.r.t. the conv input(s) (tf.GradientTape over the Conv2D above):
 *   [dx | dx2] = dz (*) rot180(w)^T ; dx2 (may be NULL) receives the channels of the
 *   second input.  If `mask` != NULL, dx *= act'(mask) where `mask` is the stored
 *   (post-activation) output of the layer that produced x -- so the kernel emits
 *   that layer's dz directly (ReLU/LeakyReLU preserve sign: act'(y) is decided by
 *   y > 0).  dx2 is never masked (it is a skip gradient that max-pool backward
 *   merges and masks). */
DNNCA_API int dnnca_conv2d_dgrad(void* stream, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                 const dnnca_tensor_t* dx2, int ksize, const dnnca_tensor_t* mask, int act,
                                 float alpha, void* workspace, size_t workspace_bytes);

/* Gradient w.r.t. kernel and bias: dw[k,k,Cx+Cx2,Cout] and db[Cout] are
 * ACCUMULATED (+=) in fp32 (caller zeroes the flat gradient buffer once per
 * step); x2 and db may be NULL. */
DNNCA_API int dnnca_conv2d_wgrad(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2,
                                 const dnnca_tensor_t* dz, float* dw, float* db, int ksize);

/* ---------------------------------------------------------------------------
 * Conv2DTranspose, k = s = 2 (non-overlapping), activation=None
 *   replaces layers.Convolution2DTranspose at components.py:118-120 and
 *   Conv2DTranspose at multiresunet.py:200-215.   kernel [2,2,Cout,Cin].
 *   y[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci]*k[a,b,co,ci] + bias[co]
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_convtranspose2x2_fprop(void* stream, const dnnca_tensor_t* x, const float* k, const float* bias,
                                 const dnnca_tensor_t* y, double* stats, void* workspace, size_t workspace_bytes);
DNNCA_API int dnnca_convtranspose2x2_dgrad(void* stream, const dnnca_tensor_t* dy, const float* k, const dnnca_tensor_t* dx,
                                 const dnnca_tensor_t* mask, int act, float alpha, void* workspace,
                                 size_t workspace_bytes);
DNNCA_API int dnnca_convtranspose2x2_wgrad(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, float* dk,
                                 float* db);

/* ---------------------------------------------------------------------------
 * MaxPool2D([2,2], strides=2), VALID
 *   replaces layers.MaxPool2D at components.py:54 / MaxPooling2D at
 *   multiresunet.py:183-195.  First maximum in row-major window order wins.
 *   idx: uint8 [n, h/2, w/2, c] dense (may be NULL for inference).
 * bwd: dx = scatter(dy, idx) (+ dskip if != NULL: the gradient arriving over the
 *   skip connection, components.py:162-164) and then (* act'(mask)) if mask != NULL.
 *   dx may alias dskip (in-place accumulate into the concat-gradient slice).
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_maxpool2x2_fwd(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* y, uint8_t* idx,
                         double* stats);
DNNCA_API int dnnca_maxpool2x2_bwd(void* stream, const dnnca_tensor_t* dy, const uint8_t* idx, const dnnca_tensor_t* dskip,
                         const dnnca_tensor_t* dx, const dnnca_tensor_t* mask, int act, float alpha);

/* ---------------------------------------------------------------------------
 * BatchNormalization(axis=-1), keras defaults momentum .99 / eps 1e-3
 *   replaces layers.BatchNormalization at components.py:57,59,130,131 and
 *   multiresunet.py:53,120,124,150,162.
 * channel_stats : stats[0..C) += sum x ; stats[C..2C) += sum x^2   (fp64)
 * bn_finalize   : training.  mean = S/M, var = SS/M - mean^2 (biased),
 *                 scale_shift[0..C) = gamma*rsqrt(var+eps) (gamma NULL -> 1),
 *                 scale_shift[C..2C) = beta - mean*scale,
 *                 mean_invstd[0..C) = mean, [C..2C) = rsqrt(var+eps),
 *                 moving_mean = moving_mean*momentum + mean*(1-momentum),
 *                 moving_var  = moving_var*momentum + var*M/(M-1)*(1-momentum).
 * bn_inference_params : scale/shift from the moving statistics (training=False).
 * bn_apply      : y = x*scale + shift  (y may be a concat slice; may alias x)
 * bn_bwd_reduce : sums[0..C) += sum dy ; sums[C..2C) += sum dy*xhat   (fp64)
 * bn_bwd_apply  : dx = gamma*invstd*(dy - sums0/M - xhat*sums1/M), then
 *                 (* act'(x)) if act != NONE (x is the activation output feeding
 *                 the BN: Conv -> act -> BN order of components.py:46-61);
 *                 dgamma += sums1 (if != NULL), dbeta += sums0.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_channel_stats(void* stream, const dnnca_tensor_t* x, double* stats);
DNNCA_API int dnnca_bn_finalize(void* stream, const double* stats, int64_t count, int c, const float* gamma,
                      const float* beta, float momentum, float eps, float* moving_mean, float* moving_var,
                      float* scale_shift, float* mean_invstd);
DNNCA_API int dnnca_bn_inference_params(void* stream, int c, const float* gamma, const float* beta, float eps,
                              const float* moving_mean, const float* moving_var, float* scale_shift);
DNNCA_API int dnnca_bn_apply(void* stream, const dnnca_tensor_t* x, const float* scale_shift, const dnnca_tensor_t* y);
DNNCA_API int dnnca_bn_bwd_reduce(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, const float* mean_invstd,
                        double* sums);
DNNCA_API int dnnca_bn_bwd_apply(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* dy, const float* mean_invstd,
                       const float* gamma, const double* sums, const dnnca_tensor_t* dx, int act, float alpha,
                       float* dgamma, float* dbeta);

/* ---------------------------------------------------------------------------
 * BatchNormalization folded into its consumer (the reference's block order is Conv -> act -> BN -> Conv,
 * components.py:46-61 and 118-134, BN -> MaxPool at :54-59): instead of materialising y = s*a + t (bn_apply), the
 * consumer reads the pre-BN tensor `a` together with the [2C] scale|shift array bn_finalize / bn_inference_params
 * wrote.  Exact, including the zero padding that follows the BN (nine border-class bias vectors).
 *   conv2d_fold_supported : 1 when conv2d_fprop_affine serves this (x, x2, y) shape (3x3, bf16, tensor-core kernel).
 *   conv2d_fprop_affine   : y = act(conv3x3([s*x+t | s2*x2+t2] zero-padded, w) + bias); affine_* may be NULL
 *                           (identity).  `scratch`: dnnca_conv2d_fold_scratch_bytes(cin, cout) bytes owned by the layer,
 *                           shared between the fprop and the wgrad call of a step.  stats as in conv2d_fprop.
 *   conv2d_wgrad_affine   : gradients w.r.t. w and bias of that layer: dW = s*dW_raw + t*S (S = sums of dz over the
 *                           pixels whose tap stays inside the image); db is required.
 *   maxpool2x2_fwd_affine : MaxPool2D of s*x + t (max of x where s >= 0, min where s < 0), y / idx / stats as in
 *                           maxpool2x2_fwd; the backward pass is maxpool2x2_bwd unchanged.
 * ------------------------------------------------------------------------- */
/* dgrad whose destination dx is the gradient of a BatchNormalization OUTPUT (the reference's Conv -> act -> BN -> Conv
 * chain, components.py:46-61,118-134, differentiated by GradientTape): besides dx (and dx2) the call leaves the BN's
 * backward sums (sum dy | sum dy*xhat over n,h,w, fp64 [2C], accumulated: zero them per step) in `sums`, taken in the
 * dgrad epilogue on the tensor-core path and by dnnca_bn_bwd_reduce otherwise.  bn_x: the BN's input; mean_invstd: what
 * dnnca_bn_finalize wrote.  dnnca_bn_bwd_apply follows unchanged. */
DNNCA_API int dnnca_conv2d_dgrad_bnreduce(void* stream, const dnnca_tensor_t* dz, const float* w, const dnnca_tensor_t* dx,
                                          const dnnca_tensor_t* dx2, int ksize, const dnnca_tensor_t* bn_x,
                                          const float* mean_invstd, double* sums, void* workspace, size_t workspace_bytes);
DNNCA_API int dnnca_convtranspose2x2_dgrad_bnreduce(void* stream, const dnnca_tensor_t* dy, const float* k, const dnnca_tensor_t* dx,
                                                    const dnnca_tensor_t* bn_x, const float* mean_invstd, double* sums,
                                                    void* workspace, size_t workspace_bytes);
DNNCA_API int dnnca_conv2d_fold_supported(const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const dnnca_tensor_t* y, int ksize);
DNNCA_API size_t dnnca_conv2d_fold_scratch_bytes(int cin, int cout);
DNNCA_API int dnnca_conv2d_fprop_affine(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* affine_x,
                                        const float* affine_x2, const float* w, const float* bias, const dnnca_tensor_t* y,
                                        int act, float alpha, double* stats, void* workspace, size_t workspace_bytes,
                                        float* scratch);
DNNCA_API int dnnca_conv2d_wgrad_affine(void* stream, const dnnca_tensor_t* x, const dnnca_tensor_t* x2, const float* affine_x,
                                        const float* affine_x2, const dnnca_tensor_t* dz, float* dw, float* db, float* scratch);
DNNCA_API int dnnca_maxpool2x2_fwd_affine(void* stream, const dnnca_tensor_t* x, const float* affine, const dnnca_tensor_t* y,
                                          uint8_t* idx, double* stats);

/* ---------------------------------------------------------------------------
 * Head Conv2D(filters=1, k=1, 'sigmoid') + WeightedCrossentropy
 *   replaces layers.Conv2D at unet.py:241-244 and losses.py:17-37, 60-72, 87-102.
 * label_stats: lstats = {sum(label) fp64, min, max} over the whole per-replica
 *   batch (tf_get_positive_rate, losses.py:87-102); 16-byte device struct,
 *   caller zero-fills it before the call (dnnca_label_stats_init does that).
 * head_fwd: logits/probs only (training=False paths: evaluate, Visualizer).
 * head_bce_fwd_bwd: z = f.w + b ; p = sigmoid(z) ; weight = cfg.weight if
 *   cfg.has_weight else (1/r if r > 0 else 1) ; weight = weight_mul*weight +
 *   weight_add ; mask = y*(weight-1)+1 ; per_sample[b] += mean_HW(mask *
 *   (max(z,0) - z*y + log1p(exp(-|z|)))) ; dz = mask*(p - y)*grad_scale with
 *   grad_scale = 1/(B*H*W*replicas) ; df = dz*w (* act'(f) if act != NONE) ;
 *   dw += sum dz*f ; db += sum dz.   per_sample/dw/db are accumulated (caller zeroes).
 * ------------------------------------------------------------------------- */
/* Label smoothing of WeightedCrossentropy (losses.py:62-67): out = tfa.image.gaussian_filter2d(label[..., None],
 * filter_shape = filter_size, sigma, padding = 'REFLECT')[..., 0] for label [n,h,w] fp32; `tmp` is a scratch tensor of
 * the same size (rows pass), `out` may not alias `label` or `tmp`. */
DNNCA_API int dnnca_gaussian_filter2d(void* stream, const float* label, int n, int h, int w, int filter_size, float sigma,
                            float* tmp, float* out);

typedef struct dnnca_label_stats {
  double sum;
  uint32_t min_key; /* order-preserving uint encoding of a float */
  uint32_t max_key;
} dnnca_label_stats_t;

typedef struct dnnca_loss_config {
  float weight; /* used when has_weight != 0 (losses.py:25) */
  int32_t has_weight;
  float weight_add; /* losses.py:29 */
  float weight_mul;
  float grad_scale; /* 1/(B*H*W*replicas) */
} dnnca_loss_config_t;

DNNCA_API int dnnca_label_stats_init(void* stream, dnnca_label_stats_t* lstats);
DNNCA_API int dnnca_label_stats(void* stream, const float* label, int64_t count, dnnca_label_stats_t* lstats);
/* host-side decode of a dnnca_label_stats_t copied back from the device */
DNNCA_API void dnnca_label_stats_decode(const dnnca_label_stats_t* host_copy, double* sum, float* min, float* max);

DNNCA_API int dnnca_head_fwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* b, float* logits,
                   float* probs);
/* Seed of the input-gradient chain: df = d(sum_p sigmoid(f.w + b)) / df = p(1-p) * w (* act'(f)).  Replaces the
 * GradientTape of callbacks.py:290-299 (g.gradient(model(x), x), "sensitivity" maps) at the head; the dgrad entry
 * points above carry it down to the network input. */
DNNCA_API int dnnca_head_input_grad(void* stream, const dnnca_tensor_t* f, const float* w, const float* b,
                                    const dnnca_tensor_t* df, int act, float alpha);
DNNCA_API int dnnca_head_bce_fwd_bwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* b,
                           const float* label, const dnnca_label_stats_t* lstats, const dnnca_loss_config_t* cfg,
                           float* logits, float* probs, float* per_sample, const dnnca_tensor_t* df, int act,
                           float alpha, float* dw, float* db);

/* ---------------------------------------------------------------------------
 * Pixel-threshold confusion counters
 *   replaces the update_state of tf.keras.metrics.Precision / Recall / AUC and of the reference's FBetaScore
 *   (annotator/utils/metrics.py:37-77) for the metrics of configs/additionals/metrics.yaml:1-23, which
 *   engine.py:273 attaches to the model: every one of them is a function of
 *       TP_k = #(y != 0 and p > t_k),  FP_k = #(y == 0 and p > t_k),  P = #(y != 0),  N = #(y == 0).
 *   `thresholds` (device, fp32, ASCENDING, nthr <= 1023).  hist: uint64 [2*(nthr+1)], ACCUMULATED (the caller zeroes
 *   it at reset_state): hist[b] += #(y != 0 and b(p) == b), hist[nthr+1+b] += #(y == 0 and b(p) == b) with
 *   b(p) = #{k : t_k < p};  hence TP_k = sum_{b > k} hist[b], FP_k likewise, P = sum_b hist[b].
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_threshold_hist(void* stream, const float* probs, const float* labels, int64_t count,
                                   const float* thresholds, int nthr, uint64_t* hist);

/* ---------------------------------------------------------------------------
 * Gradient exchange through NCCL (SURVEY 8b(ii) `dnnca_nccl_*`)
 *   replaces the cross-replica SUM all-reduce tf.distribute.MirroredStrategy performs on every gradient
 *   (annotator/engine.py:260-263) for hosts that bind this library without torch.distributed: one process per GPU,
 *   rank 0 calls dnnca_nccl_unique_id and hands the 128 bytes to the other ranks by any side channel, every rank calls
 *   dnnca_nccl_comm_init_rank with its device current, then per step dnnca_nccl_allreduce_bucket on slices of the flat
 *   gradient buffer in reverse layer order (in place, SUM; the loss is pre-scaled by 1/replicas through
 *   dnnca_loss_config_t.grad_scale, so the sum IS the average), each on the stream it is given -- the caller
 *   orders it after the kernel that wrote the bucket's last gradient and before dnnca_adam_step.  dnnca_nccl_broadcast
 *   mirrors the variables at start-up (bytes from `root`).  libnccl.so.2 is opened on first use, not at load time;
 *   failures return DNNCA_ERR_NCCL with the NCCL message in dnnca_last_error().
 * ------------------------------------------------------------------------- */
#define DNNCA_NCCL_UNIQUE_ID_BYTES 128
DNNCA_API int dnnca_nccl_unique_id(unsigned char* id128);
DNNCA_API int dnnca_nccl_comm_init_rank(void** comm, int nranks, const unsigned char* id128, int rank);
DNNCA_API int dnnca_nccl_comm_destroy(void* comm);
DNNCA_API int dnnca_nccl_allreduce_bucket(void* comm, void* stream, void* buf, int64_t count, int dtype);
DNNCA_API int dnnca_nccl_broadcast(void* comm, void* stream, void* buf, int64_t bytes, int root);

/* ---------------------------------------------------------------------------
 * Region-based detection counters ("later" row)
 *   replaces the update_state of RegionBasedRecall / Precision / TruePositives / FalsePositives / FalseNegatives /
 *   FBetaScore / ConfusionMatrix (annotator/utils/metrics.py:80-520; get_tp_fn :206-227, get_tp_fp :229-252,
 *   get_tp_fn_fp :254-288) as configured by configs/additionals/metrics.yaml:24-59 and by the Visualizer's region PR
 *   curve (callbacks.py:226-230), together with the third-party ops those call:
 *     dnnca_resize_bilinear        tf.image.resize(image, [oh, ow]) of metrics.py:194-204 (TF2 bilinear, half-pixel
 *                                  centres, no antialias) for single-channel fp32 planes [n,h,w] -> [n,oh,ow];
 *     dnnca_grey_open              annotator/utils/image.py:12-29 morph_open (tf.nn.erosion2d then tf.nn.dilation2d, zero
 *                                  structuring element, 'SAME': out-of-image taps do not take part) applied to the
 *                                  PROBABILITIES: 1[open(p) >= t] == morph_open(1[p >= t]) for every t, so one pass
 *                                  serves all thresholds; filter_size 1..15; dst may not alias src;
 *     dnnca_connected_components   tfa.image.connected_components (4-neighbourhood) of n uint8 masks [n,h,w]:
 *                                  roots[b,y,x] = row-major index (y*w + x) of the FIRST pixel of the component the
 *                                  pixel belongs to, -1 for background -- the rank of a root among the roots of the
 *                                  batch in (image, row, column) order, plus one, is the id tfa assigns;
 *     dnnca_region_confusion       the whole per-batch update for labels [n,h,w] / probabilities [n,h,w] (fp32) and
 *                                  nthr thresholds (device fp32, any order): label regions = components of
 *                                  label > 0.5, prediction regions at threshold t = components of
 *                                  open(p) >= t (metrics.py:126-141), IoU over all (label, prediction) pairs in fp32
 *                                  (metrics.py:189-191), a pair counts when IoU > iou_threshold.  Counters, each
 *                                  [4][nthr] in the order { labels detected (tp of get_tp_fn), labels missed (fn),
 *                                  predictions without a hit (fp), predictions with a hit (tp of get_tp_fp) }:
 *                                  per_slice int32 [n][4][nthr] (may be NULL; the caller zeroes it: it is the
 *                                  return_raw=True form) and totals uint64 [4][nthr] (may be NULL; ACCUMULATED).
 *                                  workspace: dnnca_region_workspace_bytes(n,h,w,nthr,table_slots) bytes, caller-owned;
 *                                  table_slots (power of two >= 64) bounds the number of distinct overlapping
 *                                  (label region, prediction region) pairs per slice and threshold; *overflow
 *                                  (device int32, caller zeroes) is set to 1 if a table filled up -- the counters
 *                                  are then invalid and the call must be repeated with a larger table.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_resize_bilinear(void* stream, const float* src, int n, int h, int w, float* dst, int oh, int ow);
DNNCA_API int dnnca_grey_open(void* stream, const float* src, int n, int h, int w, int filter_size, float* dst);
DNNCA_API int dnnca_connected_components(void* stream, const uint8_t* mask, int n, int h, int w, int32_t* roots);
DNNCA_API size_t dnnca_region_workspace_bytes(int n, int h, int w, int nthr, int table_slots);
DNNCA_API int dnnca_region_confusion(void* stream, const float* labels, const float* probs, int n, int h, int w,
                                     const float* thresholds, int nthr, float iou_threshold, int morph_filter_size,
                                     void* workspace, size_t workspace_bytes, int table_slots, int32_t* per_slice,
                                     uint64_t* totals, int32_t* overflow);

/* ---------------------------------------------------------------------------
 * MultiResUnet elementwise tail (inference): y = s2*relu(a*sa+ta + b*sb+tb)+t2
 *   replaces BatchNormalization -> add -> Activation('relu') -> BatchNormalization
 *   at multiresunet.py:120-124 and add -> relu -> BN at :148-150, :160-162.
 *   affine_a / affine_b / affine_out are [2C] scale|shift arrays (NULL = identity).
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_add_relu_affine(void* stream, const dnnca_tensor_t* a, const float* affine_a, const dnnca_tensor_t* b,
                          const float* affine_b, const float* affine_out, const dnnca_tensor_t* y);

/* ---------------------------------------------------------------------------
 * Inference-time weight folding with physical channel padding (MultiResUnet, multiresunet.py:31-60, 89-126):
 *   conv2d_bn = Conv2D(use_bias=False) -> BatchNormalization(scale=False) [-> activation] evaluated with the moving
 *   statistics is a conv with kernel K*s and bias beta - mean*s, s = gamma*rsqrt(var+eps) (gamma NULL -> 1).
 *   The model's odd channel counts (8/17/26/35/51/53/71/...) are stored in buffers whose concat segments are padded
 *   to multiples of 8 so that the tcgen05 kernels take every layer; `in_map[ci_phys]` / `out_map[co_phys]` give the
 *   logical channel of a physical one or -1 for a hole (NULL = identity).  Holes get zero weights and zero bias.
 *   layout 0: w [taps,Cin,Cout] (Conv2D HWIO) ; layout 1: w [taps,Cout,Cin] (Conv2DTranspose).  moving_var == NULL:
 *   no BatchNorm, `beta` is the layer's bias (or NULL).  w_out is fp32 in the same layout with physical extents.
 * bn_inference_params_mapped: scale|shift of an inference BatchNorm laid out over physical channels (holes: 0 | 0).
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_fold_weights(void* stream, const float* w, int taps, int cin, int cout, int layout,
                                 const int32_t* in_map, int cin_phys, const int32_t* out_map, int cout_phys,
                                 const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                                 float eps, float* w_out, float* b_out);
DNNCA_API int dnnca_bn_inference_params_mapped(void* stream, int c_phys, const int32_t* map, const float* gamma,
                                               const float* beta, float eps, const float* moving_mean,
                                               const float* moving_var, float* scale_shift);

/* ---------------------------------------------------------------------------
 * MultiResUnet TRAINING (multiresunet.yaml is trained by the reference like every other config, engine.py:286):
 * the elementwise steps of its backward pass that the U-Net kernels above do not cover.
 *   bn_apply_act  : y = act(x*scale + shift) -- Conv2D -> BatchNormalization(scale=False) -> Activation('relu')
 *                   (conv2d_bn, multiresunet.py:51-58); scale_shift as written by dnnca_bn_finalize.
 *   act_bwd       : dx = dy * act'(y), y the stored activation OUTPUT -- the gradient through Activation('relu') after
 *                   add([shortcut, out]) in ResPath (multiresunet.py:148-149, 160-161); dx may alias dy.
 *   accumulate    : dst += src -- gradients of tensors with several consumers (MultiResBlock input -> 1x1 shortcut and
 *                   3x3 chain, conv3x3 -> conv5x5 and concatenate, multiresunet.py:103-119).
 *   gather_f32    : dst[i] = idx[i] >= 0 ? src[idx[i]] : 0 -- the variables live in the reference's shapes; the training
 *                   plan computes on channel-padded (physical) copies: one gather scatters every variable into its padded
 *                   layout before the forward pass, one gathers the gradients / moving statistics back.
 *   head_conv_bwd : backward of the 1x1 conv to ONE channel of conv10 (multiresunet.py:219), whose fp32 output feeds a
 *                   BatchNormalization: df = dz * w * act'(f) (f stored post-activation), dw[c] += sum dz * f[.,c];
 *                   dz fp32 [n*h*w]; df or dw may be NULL.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_bn_apply_act(void* stream, const dnnca_tensor_t* x, const float* scale_shift, const dnnca_tensor_t* y,
                                 int act, float alpha);
DNNCA_API int dnnca_act_bwd(void* stream, const dnnca_tensor_t* y, const dnnca_tensor_t* dy, const dnnca_tensor_t* dx,
                            int act, float alpha);
DNNCA_API int dnnca_accumulate(void* stream, const dnnca_tensor_t* src, const dnnca_tensor_t* dst);
DNNCA_API int dnnca_gather_f32(void* stream, const float* src, const int32_t* idx, int64_t count, float* dst);
DNNCA_API int dnnca_head_conv_bwd(void* stream, const dnnca_tensor_t* f, const float* w, const float* dz,
                                  const dnnca_tensor_t* df, int act, float alpha, float* dw);

/* ---------------------------------------------------------------------------
 * Input tail: uint8 -> /255 -> activation dtype (data.py:193-206 `base`, 766-788)
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_u8_to_unit(void* stream, const uint8_t* src, int64_t count, void* dst, int dtype);
/* The whole input tail in one pass over the RAW combined uint8 slices [n,hin,win,s] (all slice types incl. the label
 * as channels, what the TFRecord / PNG decoder yields): crop to [hout,wout] at crop_yx[b] = (row, column) (device int32
 * [n,2]; NULL = the centre crop of data.py:182-197; the random crop of data.py:677-689 is the centre origin plus the
 * host-drawn offset), left-right flip of the cropped window where flip[b] != 0 (tf.image.random_flip_left_right,
 * data.py:620-625; NULL = none), cast and /255 (data.py:198-199), and the feature / label split of to_feature_label
 * (data.py:766-788): x[b,y,x,i] = src[..., feature_idx[i]]/255 in `x_dtype` with `x_cstride` elements per pixel
 * (>= nf: the padded bf16 input buffer of the first conv can be written directly), y[b,y,x] = src[..., label_idx]/255
 * (fp32; y_out may be NULL, label_idx < 0 writes zeros).  feature_idx is a HOST array (nf <= 16). */
DNNCA_API int dnnca_input_tail(void* stream, const uint8_t* combined, int n, int hin, int win, int s,
                               const int32_t* crop_yx, const uint8_t* flip, int hout, int wout,
                               const int32_t* feature_idx, int nf, int label_idx, void* x_out, int x_dtype,
                               int x_cstride, float* y_out);
/* Binary label masks shipped as bits: y[i] = bit i of `bits` in numpy.packbits order (bit 7 of byte 0 first), as fp32
 * 0 / 1 -- the label channel of data.py:193-206 (a PNG mask: label / 255 is 0 or 1) at 1/8 of the host->device bytes of
 * the uint8 form.  count = number of labels, a multiple of 8; y 16-byte aligned. */
DNNCA_API int dnnca_unpack_label_bits(void* stream, const uint8_t* bits, int64_t count, float* y);
/* Thin-plate-spline warp augmentation ("later" row): random_warp (data.py:718-763) =
 * tfa.image.sparse_image_warp(image, source_control_point_locations, dest_control_point_locations) with its defaults
 * (interpolation_order 2, regularization_weight 0, num_boundary_points 0).  Control points are device fp32 [n,npoints,2]
 * in (row, column) pixel coordinates, drawn by the host (data.py:742-746).
 *   dnnca_tps_fit   solves the polyharmonic-spline system of every image (interpolate_spline._solve_interpolation; train
 *                   points = DEST locations, values = dest - source) in FP64 by partial-pivoting elimination, in
 *                   coordinates divided by `extent` (the image side; any positive scale gives the same interpolant).
 *                   coef: FP64 [n][npoints+3][2] (w rows, then v rows for row, column, 1), to be used with the same
 *                   dest_points and extent.  workspace: dnnca_tps_workspace_bytes(n, npoints), 8-byte aligned.
 *                   *singular (device int32, caller zeroes) is set to 1 if a system had no usable pivot (coincident
 *                   control points): the coefficients of that call are invalid.
 *   dnnca_tps_warp  evaluates the dense flow at every pixel (interpolate_spline._apply_interpolation, in FP64: close
 *                   control points carry large weights of opposite sign) and resamples
 *                   image fp32 [n,h,w,c] at (row, column) - flow bilinearly with tfa's clamping
 *                   (dense_image_warp.interpolate_bilinear: floor in [0, size-2], weights in [0, 1]) into out (same
 *                   shape, may not alias image).  flow_out: fp32 [n,h,w,2] or NULL (sparse_image_warp's second result). */
DNNCA_API size_t dnnca_tps_workspace_bytes(int n, int npoints);
DNNCA_API int dnnca_tps_fit(void* stream, const float* source_points, const float* dest_points, int n, int npoints,
                            float extent, void* workspace, size_t workspace_bytes, double* coef, int32_t* singular);
DNNCA_API int dnnca_tps_warp(void* stream, const float* image, int n, int h, int w, int c, const float* dest_points,
                             const double* coef, int npoints, float extent, float* out, float* flow_out);
/* dtype conversion between two views of equal logical shape (fp32 <-> bf16, slice copies) */
DNNCA_API int dnnca_convert(void* stream, const dnnca_tensor_t* src, const dnnca_tensor_t* dst);

/* ---------------------------------------------------------------------------
 * Adam, keras form (engine.py:276-284), one launch over the flat parameter buffer.
 *   hyper (device, fp32): {lr, beta1, beta2, eps}; step (device int64) is
 *   incremented by the kernel (graph-replay safe).  l2 != NULL: per-element L2
 *   coefficient added to the gradient as 2*l2*p (kernel_regularizer.yaml:1-4).
 *   m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ;
 *   p <- p - lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_adam_step(void* stream, float* params, const float* grads, float* m, float* v, int64_t count,
                    const float* hyper, int64_t* step, const float* l2);

/* ---------------------------------------------------------------------------
 * Scalar training loss of one step, on the device (keras Model.train_step as driven by engine.py:126-135: compiled
 * loss, losses.py:60-72 reduced over the batch, plus the kernel_regularizer terms of kernel_regularizer.yaml:1-4,
 * both at the weights the forward pass used):
 *   out[0] += scale * (mean_b per_sample[b] + sum_i l2[i]*params[i]^2)      (l2 may be NULL; caller zeroes out)
 * `scale` = 1/replicas: the SUM all-reduce of engine.py:260-263 then yields the global mean keras reports.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_loss_total(void* stream, const float* per_sample, int batch, const float* params, const float* l2,
                               int64_t count, float scale, float* out);

/* ---------------------------------------------------------------------------
 * Gradient all-reduce fused into the Adam step over NVLink peer memory (replaces, for small models, MirroredStrategy's
 * NCCL all-reduce + the per-variable ResourceApplyAdam that follow the backward pass: engine.py:260-263, 276-284).
 * One process per GPU on one box; every rank's flat gradient buffer and a 17-word flag array are allocated with
 * p2p_alloc (cudaMalloc), exported / imported through CUDA IPC handles (64 bytes each, exchanged by the host side) and
 * p2p_adam_step launches ONE kernel per rank that (1) publishes "gradients complete" to all peers and waits for theirs,
 * (2) sums element i over all ranks' buffers straight from peer memory in rank order and applies the Keras-form Adam
 * update of adam_step to the local parameters, (3) publishes "done reading".  p2p_wait_done is the first kernel of the
 * next step: it holds this rank's gradient zeroing back until every peer has finished reading.
 *   count        = trainable parameters (Adam runs over [0, count));
 *   reduce_count = elements summed (>= count: the 4-float tail of the gradient buffer carries the loss scalar);
 *   reduced_out  = LOCAL buffer [reduce_count] receiving the sums (may be NULL);  p2p_epoch = device int64 exchange
 *   counter (starts at 0, identical on all ranks);  done_blocks = device uint32, zero.
 *   flags[16] becomes non-zero if a wait timed out (a peer died): the step's result is then undefined.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_p2p_alloc(size_t bytes, void** out);
DNNCA_API int dnnca_p2p_free(void* ptr);
DNNCA_API int dnnca_p2p_export(void* ptr, unsigned char* handle64);
DNNCA_API int dnnca_p2p_import(const unsigned char* handle64, void** out);
DNNCA_API int dnnca_p2p_close(void* ptr);
DNNCA_API int dnnca_p2p_wait_done(void* stream, void* local_flags, int world, const int64_t* p2p_epoch);
DNNCA_API int dnnca_p2p_adam_step(void* stream, const void* const* peer_grads, void* const* peer_flags, int rank, int world,
                                  float* params, float* m, float* v, int64_t count, int64_t reduce_count, float* reduced_out,
                                  const float* hyper, int64_t* step, int64_t* p2p_epoch, const float* l2,
                                  unsigned int* done_blocks);

/* ---------------------------------------------------------------------------
 * Page-locked host staging buffers for the input tail (the reference's tf.data pipeline ends with prefetch,
 * data.py:110; here the H2D copy of batch i+1 runs from these buffers while batch i computes).
 * write_combined != 0 -> cudaHostAllocWriteCombined: DMA reads are not snooped through the CPU caches (for buffers
 * the CPU only writes).  The one place where the library allocates: HOST memory, on explicit request.
 * ------------------------------------------------------------------------- */
DNNCA_API int dnnca_host_alloc(size_t bytes, int write_combined, void** out);
DNNCA_API int dnnca_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* DNNCA_H_ */
