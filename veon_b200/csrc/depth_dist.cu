// Depth-distribution producer: the step in front of the lifting path.
//
// Reference: LSSViewTransformerRaw.downsample_depth + get_two_hot_depth
// (view_transformer_raw.py:393-429): view/permute/contiguous + where + min, then repeat to
// D+1 bins, abs, where, softmax, slice, permute -- about ten ATen kernels and a [.., D+1]
// temporary four times the size of the result.  Here one kernel: a thread per output pixel takes
// the minimum of its s x s block (0 = no measurement counts as 1e5), evaluates the D+1 clamped
// gaps three times from registers (max, sum, write) and stores the D probabilities with the
// depth axis outermost, which is the layout the pooling kernels gather from.  Forward only.
#include "common.cuh"

namespace veon {

__global__ void __launch_bounds__(256)
k_two_hot_depth(const float* __restrict__ depths, int64_t n_pix, int H, int W, int s, int D,
                float bin0, float step, float gamma, float* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const int w = (int)(p % W);
  const int64_t r = p / W;
  const int h = (int)(r % H);
  const int64_t bn = r / H;
  float d;
  if (s <= 1) {
    d = __ldg(depths + p);
  } else {
    const float* blk = depths + ((bn * H + h) * s) * (int64_t)(W * s) + (int64_t)w * s;
    d = 1e5f;
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) {
        float v = __ldg(blk + (int64_t)i * (W * s) + j);
        v = (v == 0.0f) ? 1e5f : v;
        d = fminf(d, v);
      }
  }
  // gap_k = max(-|d - c_k| * gamma, -16), c_k = k * step + bin0   (k = 0..D)
  auto gap = [&](int k) {
    const float c = __fadd_rn(__fmul_rn((float)k, step), bin0);
    const float g = -fabsf(d - c) * gamma;
    return g >= -16.0f ? g : -16.0f;
  };
  float m = -16.0f;
  for (int k = 0; k <= D; ++k) m = fmaxf(m, gap(k));
  double acc = 0.0;  // D+1 terms: keep the normaliser free of summation-order effects
  for (int k = 0; k <= D; ++k) acc += (double)expf(gap(k) - m);
  const float sum = (float)acc;
  float* o = out + (bn * D * H + h) * (int64_t)W + w;
  const int64_t plane = (int64_t)H * W;
  for (int k = 0; k < D; ++k) o[k * plane] = expf(gap(k) - m) / sum;
}

}  // namespace veon

using namespace veon;

extern "C" int veon_two_hot_depth(const float* depths, int64_t BN, int H_out, int W_out,
                                  int downsample, int D, float depth_lo, float depth_step,
                                  float gamma, float* out, void* stream_) {
  if (!depths || !out || BN <= 0 || H_out <= 0 || W_out <= 0 || D <= 0 || downsample < 0 ||
      !(depth_step > 0.f))
    return VEON_E_BADARG;
  const int64_t n_pix = BN * H_out * W_out;
  if (n_pix > 0x7fffffffLL * 256) return VEON_E_RANGE;
  // bin centres like the reference: arange(D+1) * step + (lo + step/2), the sum in double
  const float bin0 = (float)((double)depth_lo + (double)depth_step / 2);
  k_two_hot_depth<<<(unsigned)ceil_div64(n_pix, 256), 256, 0, (cudaStream_t)stream_>>>(
      depths, n_pix, H_out, W_out, downsample < 1 ? 1 : downsample, D, bin0, depth_step, gamma, out);
  VEON_LAUNCH_CHECK();
  return 0;
}
