// Input tail of the training / evaluation step on the device.
//
// The reference prepares a batch on the host with tf.data: centre crop of the combined slice image to
// `output_size` (data.py:182-197), random crop = centre crop displaced by a clipped normal offset
// (random_crop, data.py:677-689), random left-right flip (augment_random_flip, data.py:620-625), float cast and /255
// (data.py:198-199), then the split into the feature channels and the label channel (to_feature_label,
// data.py:766-788).  Here the host hands over the RAW combined uint8 slices [n, hin, win, s] (every slice type incl.
// the label as one channel, exactly what the TFRecord / PNG decoder yields) plus per-sample crop origins and flip flags
// (the random numbers stay a host decision, like the reference's tf.random ops), and ONE pass writes
//     x [n, hout, wout, nf]  in the activation dtype (bf16 / fp32), optionally into a wider (padded) pixel stride,
//     y [n, hout, wout]      fp32,
// reading every needed source byte once.  HBM-bound: (s + 2*nf + 4) bytes per output pixel for bf16.
#include "common.cuh"

namespace dnnca {

struct TailArgs {
  const uint8_t* src;
  int n, hin, win, s;
  const int32_t* crop_yx;   // [n,2] device (row, column of the crop origin) or NULL -> centre crop
  const uint8_t* flip;      // [n] device, non-zero = mirror left-right, or NULL
  int hout, wout;
  int fidx[16];             // source channel of every feature channel
  int nf, label_idx;
  void* x;
  int x_cstride;            // elements per output pixel (>= nf; extra channels are left untouched)
  float* y;
};

template <typename TO>
__global__ void __launch_bounds__(256) input_tail_kernel(TailArgs a) {
  const long long total = (long long)a.n * a.hout * a.wout;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % a.wout);
    const long long r = p / a.wout;
    const int oy = (int)(r % a.hout), b = (int)(r / a.hout);
    int cy = (a.hin - a.hout) / 2, cx = (a.win - a.wout) / 2;        // data.py:186-187 ((shape - output_size) // 2)
    if (a.crop_yx) { cy = a.crop_yx[2 * b]; cx = a.crop_yx[2 * b + 1]; }
    // tf.image.random_flip_left_right runs AFTER the crop in the reference's augmentation order (data.py:92-101):
    // output column ox of a flipped sample is column wout-1-ox of the cropped window
    const int wx = (a.flip && a.flip[b]) ? a.wout - 1 - ox : ox;
    const uint8_t* sp = a.src + (((long long)b * a.hin + cy + oy) * a.win + cx + wx) * a.s;
    TO* xp = reinterpret_cast<TO*>(a.x) + p * a.x_cstride;
#pragma unroll 4
    for (int c = 0; c < a.nf; ++c) stf(xp + c, (float)sp[a.fidx[c]] / 255.0f);          // data.py:198-199
    if (a.y) a.y[p] = a.label_idx >= 0 ? (float)sp[a.label_idx] / 255.0f : 0.f;
  }
}

// binary label masks shipped as bits (numpy.packbits order: bit 7 of byte 0 is element 0): one thread expands one byte
// into eight fp32 labels with two 16-byte stores
__global__ void __launch_bounds__(256) unpack_label_bits_kernel(const uint8_t* __restrict__ bits, long long nbytes,
                                                                float* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (long long)gridDim.x * blockDim.x) {
    const unsigned b = bits[i];
    float4 lo, hi;
    lo.x = (float)((b >> 7) & 1u); lo.y = (float)((b >> 6) & 1u); lo.z = (float)((b >> 5) & 1u); lo.w = (float)((b >> 4) & 1u);
    hi.x = (float)((b >> 3) & 1u); hi.y = (float)((b >> 2) & 1u); hi.z = (float)((b >> 1) & 1u); hi.w = (float)(b & 1u);
    float4* o = reinterpret_cast<float4*>(y + i * 8);
    o[0] = lo;
    o[1] = hi;
  }
}

}  // namespace dnnca

using namespace dnnca;

extern "C" int dnnca_unpack_label_bits(void* stream, const uint8_t* bits, int64_t count, float* y) {
  DNNCA_CHECK_ARG(bits && y && count > 0 && count % 8 == 0, "unpack_label_bits: count must be a positive multiple of 8");
  DNNCA_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0, "unpack_label_bits: output must be 16-byte aligned");
  const long long nbytes = count / 8;
  unpack_label_bits_kernel<<<grid_for(nbytes, 256, 8), 256, 0, (cudaStream_t)stream>>>(bits, nbytes, y);
  DNNCA_LAUNCH_CHECK("unpack_label_bits");
  return DNNCA_OK;
}

extern "C" int dnnca_input_tail(void* stream, const uint8_t* combined, int n, int hin, int win, int s,
                                const int32_t* crop_yx, const uint8_t* flip, int hout, int wout,
                                const int32_t* feature_idx, int nf, int label_idx, void* x_out, int x_dtype,
                                int x_cstride, float* y_out) {
  DNNCA_CHECK_ARG(combined && x_out && n > 0 && s > 0 && hout > 0 && wout > 0 && hin >= hout && win >= wout,
                  "input_tail: bad arguments");
  DNNCA_CHECK_ARG(feature_idx && nf > 0 && nf <= 16 && x_cstride >= nf && label_idx < s, "input_tail: bad channel selection");
  DNNCA_CHECK_ARG(x_dtype == DNNCA_F32 || x_dtype == DNNCA_BF16, "input_tail: bad output dtype");
  TailArgs a{};
  a.src = combined; a.n = n; a.hin = hin; a.win = win; a.s = s; a.crop_yx = crop_yx; a.flip = flip; a.hout = hout; a.wout = wout;
  for (int i = 0; i < nf; ++i) {
    DNNCA_CHECK_ARG(feature_idx[i] >= 0 && feature_idx[i] < s, "input_tail: feature index %d out of range", feature_idx[i]);
    a.fidx[i] = feature_idx[i];
  }
  a.nf = nf; a.label_idx = label_idx; a.x = x_out; a.x_cstride = x_cstride; a.y = y_out;
  const long long total = (long long)n * hout * wout;
  const int grid = grid_for(total, 256 * 2, 8);
  if (x_dtype == DNNCA_F32) input_tail_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else input_tail_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  DNNCA_LAUNCH_CHECK("input_tail");
  return DNNCA_OK;
}
