// Micro-benchmark: how fast can persistent warps write the forward's output pattern
// (per tile: 64 channel planes x 128 B, planes 4*V bytes apart) as a function of the
// number of resident warps per SM?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void st4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mode 0: warp per tile (16 x STG.128 per lane-set)   mode 1: same but `spin` dependent ALU ops between tiles
__global__ void k_store(float* out, long long V, int C, int n_tiles_per_sample, int B, int spin) {
  extern __shared__ float pad[];
  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const long long TW = (long long)gridDim.x * warps_per_cta;
  const long long n_items = (long long)B * n_tiles_per_sample;
  const int q4 = (lane & 7) * 4, r = lane >> 3;
  float x = (float)lane;
  for (long long item = (long long)blockIdx.x * warps_per_cta + (threadIdx.x >> 5); item < n_items; item += TW) {
    const long long b = item / n_tiles_per_sample;
    const long long v0 = (item - b * n_tiles_per_sample) * 32;
    float* o = out + (b * C + r) * V + v0 + q4;
    for (int i = 0; i < spin; ++i) x = x * 1.0001f + 0.5f;
#pragma unroll 4
    for (int c = r; c < C; c += 4, o += 4 * V) st4(o, make_float4(x, 0.f, 0.f, 0.f));
  }
  if (x == 123.f) pad[0] = x;
}

int main(int argc, char** argv) {
  const int B = 8, C = 64; const long long V = 640000;
  float* out; cudaMalloc(&out, sizeof(float) * B * C * V);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfgs[][3] = {{256,1,0},{256,2,0},{224,2,0},{256,3,0},{256,4,0},{256,6,0},{256,8,0},{1024,2,0},{128,16,0},
                         {256,2,500},{256,2,2000},{256,2,4000},{256,4,2000},{256,4,4000},{256,8,4000}};
  for (auto& c : cfgs) {
    const int threads = c[0], ctas = c[1], spin = c[2];
    // dynamic smem sized so that exactly `ctas` CTAs fit per SM
    int smem = (227 * 1024) / ctas - 2048; if (smem > 227 * 1024 - 1024) smem = 227 * 1024 - 1024;
    cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 3; ++it) k_store<<<sms * ctas, threads, smem>>>(out, V, C, (int)(V / 32), B, spin);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) k_store<<<sms * ctas, threads, smem>>>(out, V, C, (int)(V / 32), B, spin);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("threads %4d x %d CTAs/SM = %2d warps/SM spin %4d : %7.1f us  %6.0f GB/s  (%s)\n", threads, ctas, threads / 32 * ctas, spin,
           ms * 1e3, sizeof(float) * B * C * V / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
