// bev_pool_v2 forward, fused with the reference's memset and transpose.
//
// Reference behaviour: QuickCumsumCuda.forward (bev_pool.py:17-41: new_zeros
// + bev_pool_v2_kernel, bev_pool_cuda.cu:21-48) followed by
// `.permute(0,4,1,2,3).contiguous()` (bev_pool.py:91).  The reference runs one
// thread per (interval, channel), writes only occupied voxels of a
// channels-last volume, and needs a memset before and a full transpose after.
//
// Here one warp owns one 32-voxel x-run of the output ("tile") for a chunk of
// 32*KCH channels.  Because points are sorted by voxel, the tile's points are
// one contiguous slice [tile_start[t], tile_start[t+1]) of the rank arrays.
// Lanes run over CHANNELS while accumulating (feature rows are read as full
// 128-byte lines) and over VOXELS while storing (one 16-byte store per lane =
// four full 128-byte lines of four channel planes per instruction); a padded
// shared-memory tile [c][33] does the transposition conflict-free both ways.  Empty voxels are
// written as zeros from an occupancy mask, so the volume is touched exactly
// once: no memset, no permute pass, no atomics.  Accumulation order inside a
// voxel is the rank order, fma(feat, depth, acc) starting from 0 -- the same
// sequence of roundings as the reference kernel's `psum += feat * depth`.
#include "common.cuh"

namespace veon {

constexpr int kFwdWarps = 4;
constexpr int kTilePitch = kTileVoxels + 1;  // 33: conflict-free both ways

template <int KCH>
__global__ void __launch_bounds__(kFwdWarps * 32)
k_pool_fwd(const float* __restrict__ depth, const float* __restrict__ feat,
           const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
           const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ tile_start,
           int64_t n_tiles, int64_t tiles_per_sample, int64_t V, int C, int n_chunks,
           int vec_ok, float* __restrict__ out) {
  constexpr int CC = 32 * KCH;
  constexpr int U = 4;  // points whose feature rows are in flight together
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = smem + warp * (CC * kTilePitch);

  // consecutive CTAs share a tile group and differ in channel chunk, so the
  // rank slice they all read stays in L1/L2
  const int64_t group = blockIdx.x / n_chunks;
  const int cbase = (int)(blockIdx.x - group * n_chunks) * CC;
  const int64_t t = group * kFwdWarps + warp;
  if (t >= n_tiles) return;  // no block-level barrier below
  const int64_t b = t / tiles_per_sample;
  const int64_t v0 = (t - b * tiles_per_sample) * kTileVoxels;
  const int64_t g0 = b * V + v0;

  const int32_t s = __ldg(tile_start + t), e = __ldg(tile_start + t + 1);
  uint32_t occ = 0;
  int32_t prev_rb = -1;
  for (int32_t base = s; base < e; base += 32) {
    const int32_t i = base + lane;
    const bool valid = i < e;
    int32_t my_rb = -1, my_rf = 0;
    float my_d = 0.f;
    if (valid) {
      my_rb = __ldg(ranks_bev + i);
      my_rf = __ldg(ranks_feat + i);
      my_d = __ldg(depth + __ldg(ranks_depth + i));
    }
    int32_t up = __shfl_up_sync(0xffffffffu, my_rb, 1);
    if (lane == 0) up = prev_rb;
    const bool first = valid && (my_rb != up);
    const int32_t vl = (int32_t)(my_rb - g0);  // 0..31 when valid
    const int32_t packed = (vl & 0xff) | (first ? 0x100 : 0);
    prev_rb = __shfl_sync(0xffffffffu, my_rb, 31);
    occ |= __reduce_or_sync(0xffffffffu, first ? (1u << (vl & 31)) : 0u);
    const int cnt = min(32, e - base);
    for (int j0 = 0; j0 < cnt; j0 += U) {
      float f[U][KCH];
      float dj[U];
      int pj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = min(j0 + u, 31);
        pj[u] = __shfl_sync(0xffffffffu, packed, j);
        dj[u] = __shfl_sync(0xffffffffu, my_d, j);
        const int32_t fj = __shfl_sync(0xffffffffu, my_rf, j);
        const float* frow = feat + (int64_t)fj * C + cbase + lane;
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
          const bool in = (j0 + u < cnt) && (cbase + lane + 32 * k < C);
          f[u][k] = in ? __ldg(frow + 32 * k) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j0 + u < cnt) {
          const int vlj = pj[u] & 0xff;
          const bool fst = (pj[u] & 0x100) != 0;
#pragma unroll
          for (int k = 0; k < KCH; ++k) {
            float* a = tile + (lane + 32 * k) * kTilePitch + vlj;
            const float acc = fst ? 0.f : *a;
            *a = fmaf(f[u][k], dj[u], acc);
          }
        }
      }
    }
  }
  __syncwarp();

  // Write-out: lane (r = lane/8, q = lane%8) stores voxels 4q..4q+3 of channel
  // 4*it + r as one 16-byte store, so one instruction covers 4 channel planes x
  // 128 bytes.  The four scalar shared-memory reads behind it hit banks
  // (c + 4q + i) mod 32 = all distinct (pitch 33).
  const int q4 = (lane & 7) * 4, r = lane >> 3;
  const uint32_t occ4 = (occ >> q4) & 0xfu;
  const int cmax = min(CC, C - cbase);
  float* o = out + ((int64_t)b * C + cbase + r) * V + v0 + q4;
  const int64_t ostep = 4 * V;
  const float* trow = tile + r * kTilePitch + q4;
  if (vec_ok && v0 + kTileVoxels <= V) {  // full, 16-byte aligned tile (the usual case)
    if (occ == 0u) {
#pragma unroll 4
      for (int c = r; c < cmax; c += 4, o += ostep) st_stream4(o, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
#pragma unroll 4
      for (int c = r; c < cmax; c += 4, o += ostep, trow += 4 * kTilePitch) {
        float4 v;
        v.x = (occ4 & 1u) ? trow[0] : 0.f;
        v.y = (occ4 & 2u) ? trow[1] : 0.f;
        v.z = (occ4 & 4u) ? trow[2] : 0.f;
        v.w = (occ4 & 8u) ? trow[3] : 0.f;
        st_stream4(o, v);
      }
    }
  } else {  // ragged volume edge: scalar, bounds-checked
    for (int c = r; c < cmax; c += 4, o += ostep, trow += 4 * kTilePitch)
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (v0 + q4 + i < V) st_stream(o + i, ((occ4 >> i) & 1u) ? trow[i] : 0.f);
  }
}

template <int KCH>
static int launch_fwd(const float* depth, const float* feat, const int32_t* rd,
                      const int32_t* rf, const int32_t* rb, const int32_t* tile_start, int B,
                      int C, int64_t V, float* out, cudaStream_t stream) {
  constexpr int CC = 32 * KCH;
  const size_t smem = sizeof(float) * kFwdWarps * CC * kTilePitch;
  static bool attr_set = false;
  if (!attr_set) {
    VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd<KCH>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int64_t tps = ceil_div64(V, kTileVoxels), n_tiles = (int64_t)B * tps;
  const int n_chunks = (C + CC - 1) / CC;
  const int vec_ok = ((V & 3) == 0) && (((uintptr_t)out & 15) == 0);
  const int64_t blocks = ceil_div64(n_tiles, kFwdWarps) * n_chunks;
  if (blocks > 0x7fffffffLL) return VEON_E_RANGE;
  k_pool_fwd<KCH><<<(unsigned)blocks, kFwdWarps * 32, smem, stream>>>(
      depth, feat, rd, rf, rb, tile_start, n_tiles, tps, V, C, n_chunks, vec_ok, out);
  VEON_LAUNCH_CHECK();
  return 0;
}

}  // namespace veon

using namespace veon;

// channel-chunk override for tuning (0 = automatic); read once
static int fwd_kch_override() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VEON_FWD_KCH");
    v = e ? atoi(e) : 0;
  }
  return v;
}

extern "C" int veon_bev_pool_v2_fwd_planar(const float* depth, const float* feat,
                                           const int32_t* ranks_depth,
                                           const int32_t* ranks_feat,
                                           const int32_t* ranks_bev,
                                           const int32_t* tile_start, int B, int C, int64_t V,
                                           float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !tile_start || !out ||
      B <= 0 || C <= 0 || V <= 0)
    return VEON_E_BADARG;
  int kch = fwd_kch_override();
  if (kch == 0) kch = (C <= 32) ? 1 : 2;
  switch (kch) {
    case 1: return launch_fwd<1>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, B, C, V, out, stream);
    case 2: return launch_fwd<2>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, B, C, V, out, stream);
    case 4: return launch_fwd<4>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, B, C, V, out, stream);
    default: return VEON_E_BADARG;
  }
}
