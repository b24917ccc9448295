// Shared between geometry.cu (get_lidar_coor) and prepare.cu (geometry fused into the
// classification kernel): the per-camera transform and the per-point formula.  One inline
// function, so both routes execute the same float operations in the same order and the ranks
// of the fused route are the bits of the unfused one.
#pragma once
#include "common.cuh"

namespace veon {

struct CamXform {
  float undo[9];   // inverse(post_rots)
  float c2e[9];    // sensor2ego[:3,:3] @ inverse(cam2imgs)
  float pt[3];     // post_trans
  float t[3];      // sensor2ego[:3,3]
  float bda[9];
};

__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
  return fmaf(m[2], z, fmaf(m[1], y, m[0] * x));
}

// frustum point f3 = (x_img, y_img, depth) of one camera -> ego/lidar frame
// (view_transformer.py:114-152, same operation order)
__device__ __forceinline__ void lidar_point(const float* __restrict__ f3, const CamXform& x,
                                            float& ox, float& oy, float& oz) {
  const float fx = __ldg(f3) - x.pt[0];
  const float fy = __ldg(f3 + 1) - x.pt[1];
  const float fz = __ldg(f3 + 2) - x.pt[2];
  float qx = dot3(x.undo, fx, fy, fz), qy = dot3(x.undo + 3, fx, fy, fz);
  const float qz = dot3(x.undo + 6, fx, fy, fz);
  qx *= qz;
  qy *= qz;
  const float ex = dot3(x.c2e, qx, qy, qz) + x.t[0];
  const float ey = dot3(x.c2e + 3, qx, qy, qz) + x.t[1];
  const float ez = dot3(x.c2e + 6, qx, qy, qz) + x.t[2];
  ox = dot3(x.bda, ex, ey, ez);
  oy = dot3(x.bda + 3, ex, ey, ez);
  oz = dot3(x.bda + 6, ex, ey, ez);
}

// host: fills xf[B*N] (one launch); defined in geometry.cu
int launch_cam_xforms(const float* sensor2ego, const float* cam2imgs, const float* post_rots,
                      const float* post_trans, const float* bda, int B, int N, CamXform* xf,
                      cudaStream_t stream);

}  // namespace veon
