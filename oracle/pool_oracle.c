/* TEST INFRASTRUCTURE ONLY: plain-C restatement of the reference's two CUDA
 * kernels, one loop nest per kernel, same accumulation order, same
 * channels-last layout, caller-zeroed outputs.
 *
 *   oracle_bev_pool_v2       <- bev_pool_v2_kernel   bev_pool_cuda.cu:21-48
 *   oracle_bev_pool_v2_grad  <- bev_pool_grad_kernel bev_pool_cuda.cu:67-121
 *
 * nvcc contracts the kernels' `psum += a * b` into one FMA, so fmaf() is used
 * here to reproduce the same sequence of roundings.  Offsets are 64-bit (the
 * reference's int32 `rank * c` overflows at C4; irrelevant on the CPU sizes
 * this runs at).  OpenMP over intervals = "all host threads it can use".
 */
#include <math.h>
#include <stdint.h>

void oracle_bev_pool_v2(int c, int n_intervals, const float* depth, const float* feat,
                        const int32_t* ranks_depth, const int32_t* ranks_feat,
                        const int32_t* ranks_bev, const int32_t* interval_starts,
                        const int32_t* interval_lengths, float* out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int index = 0; index < n_intervals; ++index) {
    const int start = interval_starts[index], length = interval_lengths[index];
    float* cur_out = out + (int64_t)ranks_bev[start] * c;
    for (int cur_c = 0; cur_c < c; ++cur_c) {
      float psum = 0.0f;
      for (int i = 0; i < length; ++i)
        psum = fmaf(feat[(int64_t)ranks_feat[start + i] * c + cur_c],
                    depth[ranks_depth[start + i]], psum);
      cur_out[cur_c] = psum;
    }
  }
}

/* intervals here are runs of equal ranks_feat (bev_pool.py:47-57) */
void oracle_bev_pool_v2_grad(int c, int n_intervals, const float* out_grad, const float* depth,
                             const float* feat, const int32_t* ranks_depth,
                             const int32_t* ranks_feat, const int32_t* ranks_bev,
                             const int32_t* interval_starts, const int32_t* interval_lengths,
                             float* depth_grad, float* feat_grad) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int idx = 0; idx < n_intervals; ++idx) {
    const int start = interval_starts[idx], length = interval_lengths[idx];
    for (int i = 0; i < length; ++i) {
      const float* g = out_grad + (int64_t)ranks_bev[start + i] * c;
      const float* f = feat + (int64_t)ranks_feat[start + i] * c;
      float s = 0.0f;
      for (int cur_c = 0; cur_c < c; ++cur_c) s = fmaf(g[cur_c], f[cur_c], s);
      depth_grad[ranks_depth[start + i]] = s;
    }
    float* fg = feat_grad + (int64_t)ranks_feat[start] * c;
    for (int cur_c = 0; cur_c < c; ++cur_c) {
      float s = 0.0f;
      for (int i = 0; i < length; ++i)
        s = fmaf(out_grad[(int64_t)ranks_bev[start + i] * c + cur_c],
                 depth[ranks_depth[start + i]], s);
      fg[cur_c] = s;
    }
  }
}
