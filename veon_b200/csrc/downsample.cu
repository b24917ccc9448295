// 2x2x2 max-downsample of the pooled volume and its gradient.
//
// Reference: LSSViewTransformerRaw.forward (view_transformer_raw.py:549-553)
//   bev_feat.view(b, c, z/2, 2, y/2, 2, x/2, 2).amax(dim=(3, 5, 7))
// i.e. a strided 8-D reduction in ATen (2.8 ms for the 1.31 GB volume of C2 on B200) and, in
// training, ATen's amax backward: grad * (in == out) / count(in == out), ties sharing equally.
// Both are plain streaming passes here: a thread owns two neighbouring outputs = one 16-byte
// load from each of the four input x-rows, so every access is a full coalesced line.
// NaN propagates like torch.amax (a NaN input makes the output NaN).
#include "common.cuh"

namespace veon {

__device__ __forceinline__ float max_nan(float a, float b) { return (b > a || b != b) ? b : a; }

// volumes: in [BC][Z][Y][X], out [BC][Z/2][Y/2][X/2]; X % 4 == 0, Z and Y even
// MASK: also store, per output, which of its 8 inputs equal the maximum (bit (dz*2+dy)*2+dx):
// all the backward needs (grad * bit / popcount), so the volume need not be kept for it.
template <bool MASK>
__global__ void __launch_bounds__(256)
k_maxdown2_fwd(const float* __restrict__ in, int64_t n_quads, int Zh, int Yh, int X4, int Y, int X,
               float* __restrict__ out, uint8_t* __restrict__ mask) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(q % X4);
    int64_t r = q / X4;
    const int yo = (int)(r % Yh);
    r /= Yh;
    const int zo = (int)(r % Zh);
    const int64_t bc = r / Zh;
    const float* p = in + ((bc * (2 * Zh) + 2 * zo) * Y + 2 * yo) * (int64_t)X + 4 * x4;
    const float4 a = ld_stream4(p), b = ld_stream4(p + X);
    const float4 c = ld_stream4(p + (int64_t)Y * X), d = ld_stream4(p + (int64_t)Y * X + X);
    float2 o;
    o.x = max_nan(max_nan(max_nan(a.x, a.y), max_nan(b.x, b.y)),
                  max_nan(max_nan(c.x, c.y), max_nan(d.x, d.y)));
    o.y = max_nan(max_nan(max_nan(a.z, a.w), max_nan(b.z, b.w)),
                  max_nan(max_nan(c.z, c.w), max_nan(d.z, d.w)));
    const int64_t oi = ((bc * Zh + zo) * Yh + yo) * (int64_t)(X / 2) + 2 * x4;
    *reinterpret_cast<float2*>(out + oi) = o;
    if (MASK) {
      const uint32_t m0 = (a.x == o.x) | (a.y == o.x) << 1 | (b.x == o.x) << 2 | (b.y == o.x) << 3 |
                          (c.x == o.x) << 4 | (c.y == o.x) << 5 | (d.x == o.x) << 6 | (d.y == o.x) << 7;
      const uint32_t m1 = (a.z == o.y) | (a.w == o.y) << 1 | (b.z == o.y) << 2 | (b.w == o.y) << 3 |
                          (c.z == o.y) << 4 | (c.w == o.y) << 5 | (d.z == o.y) << 6 | (d.w == o.y) << 7;
      *reinterpret_cast<uchar2*>(mask + oi) = make_uchar2((unsigned char)m0, (unsigned char)m1);
    }
  }
}

__global__ void __launch_bounds__(256)
k_maxdown2_bwd(const float* __restrict__ in, const float* __restrict__ out,
               const float* __restrict__ grad_out, int64_t n_quads, int Zh, int Yh, int X4, int Y,
               int X, float* __restrict__ grad_in) {
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(q % X4);
    int64_t r = q / X4;
    const int yo = (int)(r % Yh);
    r /= Yh;
    const int zo = (int)(r % Zh);
    const int64_t bc = r / Zh;
    const int64_t ibase = ((bc * (2 * Zh) + 2 * zo) * Y + 2 * yo) * (int64_t)X + 4 * x4;
    const int64_t obase = ((bc * Zh + zo) * Yh + yo) * (int64_t)(X / 2) + 2 * x4;
    const int64_t step[4] = {0, X, (int64_t)Y * X, (int64_t)Y * X + X};
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream4(in + ibase + step[k]);
    const float2 m = *reinterpret_cast<const float2*>(out + obase);
    const float2 g = *reinterpret_cast<const float2*>(grad_out + obase);
    int n0 = 0, n1 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      n0 += (v[k].x == m.x) + (v[k].y == m.x);
      n1 += (v[k].z == m.y) + (v[k].w == m.y);
    }
    // grad * mask / count, exactly ATen's formula (count == 0 only if the output is NaN)
    const float g0 = g.x / (float)n0, g1 = g.y / (float)n1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 w;
      w.x = v[k].x == m.x ? g0 : 0.f;
      w.y = v[k].y == m.x ? g0 : 0.f;
      w.z = v[k].z == m.y ? g1 : 0.f;
      w.w = v[k].w == m.y ? g1 : 0.f;
      st_stream4(grad_in + ibase + step[k], w);
    }
  }
}

static int check(const void* a, const void* b, int64_t BC, int Z, int Y, int X) {
  if (!a || !b || BC <= 0 || Z <= 0 || Y <= 0 || X <= 0) return VEON_E_BADARG;
  if ((Z & 1) || (Y & 1) || (X & 3) || ((uintptr_t)a & 15) || ((uintptr_t)b & 7))
    return VEON_E_UNSUPPORTED;
  return 0;
}

}  // namespace veon

using namespace veon;

extern "C" int veon_maxdown2_fwd(const float* in, int64_t BC, int Z, int Y, int X, float* out,
                                 void* stream_) {
  int rc = check(in, out, BC, Z, Y, X);
  if (rc) return rc;
  const int64_t n_quads = BC * (Z / 2) * (Y / 2) * (X / 4);
  int64_t blocks = ceil_div64(n_quads, 256);
  if (blocks > 148 * 64) blocks = 148 * 64;
  k_maxdown2_fwd<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(
      in, n_quads, Z / 2, Y / 2, X / 4, Y, X, out, nullptr);
  VEON_LAUNCH_CHECK();
  return 0;
}

extern "C" int veon_maxdown2_fwd_mask(const float* in, int64_t BC, int Z, int Y, int X, float* out,
                                      uint8_t* mask, void* stream_) {
  int rc = check(in, out, BC, Z, Y, X);
  if (rc) return rc;
  if (!mask) return VEON_E_BADARG;
  if ((uintptr_t)mask & 1) return VEON_E_UNSUPPORTED;
  const int64_t n_quads = BC * (Z / 2) * (Y / 2) * (X / 4);
  int64_t blocks = ceil_div64(n_quads, 256);
  if (blocks > 148 * 64) blocks = 148 * 64;
  k_maxdown2_fwd<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(
      in, n_quads, Z / 2, Y / 2, X / 4, Y, X, out, mask);
  VEON_LAUNCH_CHECK();
  return 0;
}

extern "C" int veon_maxdown2_bwd(const float* in, const float* out, const float* grad_out,
                                 int64_t BC, int Z, int Y, int X, float* grad_in, void* stream_) {
  int rc = check(in, out, BC, Z, Y, X);
  if (rc) return rc;
  if (!grad_out || !grad_in) return VEON_E_BADARG;
  if (((uintptr_t)grad_in & 15) || ((uintptr_t)grad_out & 7)) return VEON_E_UNSUPPORTED;
  const int64_t n_quads = BC * (Z / 2) * (Y / 2) * (X / 4);
  int64_t blocks = ceil_div64(n_quads, 256);
  if (blocks > 148 * 64) blocks = 148 * 64;
  k_maxdown2_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(
      in, out, grad_out, n_quads, Z / 2, Y / 2, X / 4, Y, X, grad_in);
  VEON_LAUNCH_CHECK();
  return 0;
}
