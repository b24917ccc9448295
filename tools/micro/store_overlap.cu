// Micro-benchmark: does per-tile work overlap with the forward's store stream, and does it
// matter whether that work is ALU-only or shared-memory traffic?  Persistent warps, 2 CTAs x 8
// warps per SM, per item: `work` units of compute, then the tile's 16 x STG.128.
//   kind 0: dependent FMA chain          (issue slots only)
//   kind 1: LDS.128 + STS.128 round trip (goes through the SM's load/store unit)
//   kind 2: like 1, but the stores are skipped (cost of the work alone)
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void st4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(256)
k(float* out, long long V, int C, int tps, int B, int work, int kind) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* mine = reinterpret_cast<float4*>(sm) + warp * 576;  // a 9 KB region per warp
  const long long TW = (long long)gridDim.x * 8, n_items = (long long)B * tps;
  const int q4 = (lane & 7) * 4, r = lane >> 3;
  float4 x = make_float4((float)lane, 1.f, 2.f, 3.f);
  for (int i = lane; i < 576; i += 32) mine[i] = x;
  __syncwarp();
  for (long long item = (long long)blockIdx.x * 8 + warp; item < n_items; item += TW) {
    const long long b = item / tps, v0 = (item - b * tps) * 32;
    float* o = out + (b * C + r) * V + v0 + q4;
    if (kind == 0) {
      for (int i = 0; i < work; ++i) x.x = x.x * 1.0001f + 0.5f;
    } else {
      for (int i = 0; i < work; ++i) {   // 2 x 4 wavefronts per iteration
        float4 y = mine[(lane + 32 * (i & 15))];
        y.x += x.x;
        mine[(lane + 32 * ((i + 1) & 15))] = y;
        x.x = y.y;
      }
      __syncwarp();
    }
    if (kind != 2) {
#pragma unroll 4
      for (int c = r; c < C; c += 4, o += 4 * V) st4(o, x);
    }
  }
  if (x.x == 123.f) out[0] = x.x;
}

int main() {
  const int B = 8, C = 64; const long long V = 640000;
  float* out; cudaMalloc(&out, sizeof(float) * B * C * V);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 8 * 576 * 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfgs[][2] = {{0,0},{500,0},{1000,0},{1500,0},{16,1},{16,2},{32,1},{32,2},{64,1},{64,2},{128,1},{128,2}};
  for (auto& c : cfgs) {
    for (int it = 0; it < 3; ++it) k<<<sms * 2, 256, smem>>>(out, V, C, (int)(V / 32), B, c[0], c[1]);
    cudaEventRecord(e0);
    for (int it = 0; it < 10; ++it) k<<<sms * 2, 256, smem>>>(out, V, C, (int)(V / 32), B, c[0], c[1]);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("kind %d work %5d : %7.1f us (%s)\n", c[1], c[0], ms * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
