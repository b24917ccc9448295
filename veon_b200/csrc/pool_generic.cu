// Interval-driven pooling kernels with the reference's literal semantics, for
//  (a) the binary-compatible entry points veon_bev_pool_v2 / _grad (what
//      bev_pool.cpp:7-14 declares and bev_pool_cuda.cu:125-140 defines), and
//  (b) rank arrays that veon_pool_plan_build() rejects (unsorted ranks_bev,
//      hand-made intervals, non-canonical ranks_feat).
// Still B200-shaped rather than a copy of the reference kernels: one WARP per
// interval with lanes over channels (coalesced 128-byte feature / gradient
// rows, 64-bit offsets) instead of one thread per (interval, channel) /
// one thread per pixel.
#include "common.cuh"

namespace veon {

__device__ __forceinline__ int64_t vol_index(int32_t rank, int c, int C, int layout,
                                             int64_t V) {
  if (layout == VEON_LAYOUT_BZYXC) return (int64_t)rank * C + c;
  const int64_t b = rank / V, v = rank - b * V;
  return (b * C + c) * V + v;
}

// forward: out[rank(interval)][c] = sum_i depth[rd_i] * feat[rf_i][c]
__global__ void __launch_bounds__(256)
k_generic_fwd(int C, int64_t n_intervals, int layout, int64_t V,
              const float* __restrict__ depth, const float* __restrict__ feat,
              const int32_t* __restrict__ rd, const int32_t* __restrict__ rf,
              const int32_t* __restrict__ rb, const int32_t* __restrict__ istarts,
              const int32_t* __restrict__ ilens, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_intervals) return;
  const int32_t s = istarts[k], l = ilens[k];
  const int32_t rank = rb[s];
  for (int c0 = 0; c0 < C; c0 += 32) {
    const int c = c0 + lane;
    float acc = 0.f;
    for (int32_t i = 0; i < l; ++i) {
      const float d = __ldg(depth + rd[s + i]);
      const float f = c < C ? __ldg(feat + (int64_t)rf[s + i] * C + c) : 0.f;
      acc = fmaf(f, d, acc);
    }
    if (c < C) out[vol_index(rank, c, C, layout, V)] = acc;
  }
}

// literal backward: intervals are runs of equal ranks_feat (bev_pool.py:47-57)
__global__ void __launch_bounds__(256)
k_generic_bwd_intervals(int C, int64_t n_intervals, int layout, int64_t V,
                        const float* __restrict__ out_grad, const float* __restrict__ depth,
                        const float* __restrict__ feat, const int32_t* __restrict__ rd,
                        const int32_t* __restrict__ rf, const int32_t* __restrict__ rb,
                        const int32_t* __restrict__ istarts, const int32_t* __restrict__ ilens,
                        float* __restrict__ depth_grad, float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (k >= n_intervals) return;
  const int32_t s = istarts[k], l = ilens[k];
  // depth_grad[rd_i] = <out_grad[rb_i, :], feat[rf_i, :]>
  for (int32_t i = 0; i < l; ++i) {
    const int32_t rank = rb[s + i];
    const float* frow = feat + (int64_t)rf[s + i] * C;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32)
      dot = fmaf(__ldg(out_grad + vol_index(rank, c, C, layout, V)), __ldg(frow + c), dot);
    dot = warp_sum(dot);
    if (lane == 0) depth_grad[rd[s + i]] = dot;
  }
  // feat_grad[rf_start, c] = sum_i out_grad[rb_i, c] * depth[rd_i]
  float* grow = feat_grad + (int64_t)rf[s] * C;
  for (int c = lane; c < C; c += 32) {
    float acc = 0.f;
    for (int32_t i = 0; i < l; ++i)
      acc = fmaf(__ldg(out_grad + vol_index(rb[s + i], c, C, layout, V)),
                 __ldg(depth + rd[s + i]), acc);
    grow[c] = acc;
  }
}

// point-driven backward for arbitrary rank arrays: one warp per point,
// feat_grad accumulated with float atomics (fallback only; not deterministic)
__global__ void __launch_bounds__(256)
k_generic_bwd_points(int C, int64_t n_points, int layout, int64_t V,
                     const float* __restrict__ out_grad, const float* __restrict__ depth,
                     const float* __restrict__ feat, const int32_t* __restrict__ rd,
                     const int32_t* __restrict__ rf, const int32_t* __restrict__ rb,
                     float* __restrict__ depth_grad, float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_points) return;
  const int32_t rank = rb[i];
  const float d = __ldg(depth + rd[i]);
  const float* frow = feat + (int64_t)rf[i] * C;
  float* grow = feat_grad + (int64_t)rf[i] * C;
  float dot = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float g = __ldg(out_grad + vol_index(rank, c, C, layout, V));
    dot = fmaf(g, __ldg(frow + c), dot);
    atomicAdd(grow + c, g * d);
  }
  dot = warp_sum(dot);
  if (lane == 0) depth_grad[rd[i]] = dot;
}

}  // namespace veon

using namespace veon;

static int generic_fwd(int c, int64_t n, int layout, int64_t V, const float* depth,
                       const float* feat, const int32_t* rd, const int32_t* rf,
                       const int32_t* rb, const int32_t* is, const int32_t* il, float* out,
                       void* stream) {
  if (c <= 0 || n < 0 || !depth || !feat || !rd || !rf || !rb || !is || !il || !out ||
      (layout != VEON_LAYOUT_BZYXC && layout != VEON_LAYOUT_BCZYX) ||
      (layout == VEON_LAYOUT_BCZYX && V <= 0))
    return VEON_E_BADARG;
  if (n == 0) return 0;
  const int64_t blocks = ceil_div64(n * 32, 256);
  if (blocks > 0x7fffffffLL) return VEON_E_RANGE;
  k_generic_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(c, n, layout, V, depth, feat,
                                                                   rd, rf, rb, is, il, out);
  VEON_LAUNCH_CHECK();
  return 0;
}

extern "C" int veon_bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat,
                                const int32_t* ranks_depth, const int32_t* ranks_feat,
                                const int32_t* ranks_bev, const int32_t* interval_starts,
                                const int32_t* interval_lengths, float* out, void* stream) {
  return generic_fwd(c, n_intervals, VEON_LAYOUT_BZYXC, 0, depth, feat, ranks_depth, ranks_feat,
                     ranks_bev, interval_starts, interval_lengths, out, stream);
}

extern "C" int veon_bev_pool_v2_generic(int c, int n_intervals, int layout, int64_t V,
                                        const float* depth, const float* feat,
                                        const int32_t* ranks_depth, const int32_t* ranks_feat,
                                        const int32_t* ranks_bev,
                                        const int32_t* interval_starts,
                                        const int32_t* interval_lengths, float* out,
                                        void* stream) {
  return generic_fwd(c, n_intervals, layout, V, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                     interval_starts, interval_lengths, out, stream);
}

extern "C" int veon_bev_pool_v2_grad(int c, int n_intervals, const float* out_grad,
                                     const float* depth, const float* feat,
                                     const int32_t* ranks_depth, const int32_t* ranks_feat,
                                     const int32_t* ranks_bev, const int32_t* interval_starts,
                                     const int32_t* interval_lengths, float* depth_grad,
                                     float* feat_grad, void* stream) {
  if (c <= 0 || n_intervals < 0 || !out_grad || !depth || !feat || !ranks_depth || !ranks_feat ||
      !ranks_bev || !interval_starts || !interval_lengths || !depth_grad || !feat_grad)
    return VEON_E_BADARG;
  if (n_intervals == 0) return 0;
  const int64_t blocks = ceil_div64((int64_t)n_intervals * 32, 256);
  k_generic_bwd_intervals<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      c, n_intervals, VEON_LAYOUT_BZYXC, 0, out_grad, depth, feat, ranks_depth, ranks_feat,
      ranks_bev, interval_starts, interval_lengths, depth_grad, feat_grad);
  VEON_LAUNCH_CHECK();
  return 0;
}

extern "C" int veon_bev_pool_v2_grad_generic(int c, int64_t n_points, int layout, int64_t V,
                                             const float* out_grad, const float* depth,
                                             const float* feat, const int32_t* ranks_depth,
                                             const int32_t* ranks_feat,
                                             const int32_t* ranks_bev, float* depth_grad,
                                             float* feat_grad, void* stream) {
  if (c <= 0 || n_points < 0 || !out_grad || !depth || !feat || !ranks_depth || !ranks_feat ||
      !ranks_bev || !depth_grad || !feat_grad ||
      (layout != VEON_LAYOUT_BZYXC && layout != VEON_LAYOUT_BCZYX) ||
      (layout == VEON_LAYOUT_BCZYX && V <= 0))
    return VEON_E_BADARG;
  if (n_points == 0) return 0;
  const int64_t blocks = ceil_div64(n_points * 32, 256);
  if (blocks > 0x7fffffffLL) return VEON_E_RANGE;
  k_generic_bwd_points<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      c, n_points, layout, V, out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
      depth_grad, feat_grad);
  VEON_LAUNCH_CHECK();
  return 0;
}
