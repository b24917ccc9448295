// Read bandwidth of the tail kernel's operand pattern: every CTA step reads SEG bytes from each of
// ROWS channel planes (plane stride = V floats); consecutive CTAs take consecutive segments.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o read_pattern read_pattern.cu && ./read_pattern
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int SEG_F4>   // float4 per row segment per warp-row (32 -> 512 B, 64 -> 1 KB, ...)
__global__ void __launch_bounds__(512, 1) k_read(const float4* __restrict__ x, int64_t V4, int C, int64_t n_seg,
                                                 float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int64_t seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
    for (int c0 = 0; c0 < C; c0 += 32) {
      const float4* p = x + (int64_t)(c0 + 2 * warp) * V4 + seg * SEG_F4 + lane;
      float4 v[2 * SEG_F4 / 32];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int k = 0; k < SEG_F4 / 32; ++k) v[r * (SEG_F4 / 32) + k] = __ldcs(p + r * V4 + k * 32);
#pragma unroll
      for (int i = 0; i < 2 * SEG_F4 / 32; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <int SEG_F4>
void run(const float4* x, int64_t V, int C, float* sink, int ctas_per_sm) {
  const int64_t V4 = V / 4, n_seg = V4 / SEG_F4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) k_read<SEG_F4><<<148 * ctas_per_sm, 512>>>(x, V4, C, n_seg, sink);
  cudaEventRecord(e0);
  const int n = 10;
  for (int i = 0; i < n; ++i) k_read<SEG_F4><<<148 * ctas_per_sm, 512>>>(x, V4, C, n_seg, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  ms /= n;
  printf("segment %5d B, %d CTA/SM: %8.1f us  %7.1f GB/s\n", SEG_F4 * 16, ctas_per_sm, ms * 1e3,
         (double)V * C * 4 / ms / 1e6);
}

int main() {
  const int64_t V = 640000; const int C = 512;
  float4* x; float* sink;
  cudaMalloc(&x, (size_t)V * C * 4); cudaMalloc(&sink, 4);
  cudaMemset(x, 0, (size_t)V * C * 4);
  for (int cps = 1; cps <= 2; ++cps) {
    run<32>(x, V, C, sink, cps);
    run<64>(x, V, C, sink, cps);
    run<128>(x, V, C, sink, cps);
  }
  return 0;
}
