// bev_pool_v2 forward, fused with the reference's memset and transpose.
//
// Reference behaviour: QuickCumsumCuda.forward (bev_pool.py:17-41: new_zeros
// + bev_pool_v2_kernel, bev_pool_cuda.cu:21-48) followed by
// `.permute(0,4,1,2,3).contiguous()` (bev_pool.py:91).  The reference runs one
// thread per (interval, channel), writes only occupied voxels of a
// channels-last volume, and needs a memset before and a full transpose after.
//
// Here one warp handles one 32-voxel x-run of the output ("tile") for a chunk of
// 32*KCH channels at a time (persistent grid, warps stride over the tiles).
// Because points are sorted by voxel, the tile's points are one contiguous
// slice [tile_start[t], tile_start[t+1]) of the rank arrays.  Lanes run over
// CHANNELS while gathering (feature rows are read as full 128-byte lines, the
// per-voxel sum lives in registers) and over VOXELS while storing (one 16-byte
// store per lane = four full 128-byte lines of four channel planes per
// instruction); a padded, zero-filled shared tile [c][36] does the transposition.
// Empty voxels come out as zeros, so the volume is touched exactly once: no
// memset, no permute pass, no atomics.  Accumulation order inside a voxel is the
// rank order, fma(feat, depth, acc) starting from 0 -- the same sequence of
// roundings as the reference kernel's `psum += feat * depth`.
//
// Kernels in this file:
//   k_pool_fwd         the warp-per-tile kernel above (all tiles below the heavy threshold)
//   k_pool_fwd_heavy   a CTA per heavy tile (rows staged by cp.async), queued behind it
//   k_pool_fwd_narrow  C <= 32 (C % 4 == 0): a LANE per voxel instead of a lane per channel, all
//                      channels of the voxel in registers, same fma order; heavy grid first
//   k_pool_fwd_group   opt-in experiment: a CTA per group of tiles, software-pipelined
//   k_pool_ds_fwd      opt-in: pooling fused with the neck's 2x2x2 max-downsample
#include "common.cuh"

#ifndef VEON_FWD_WARPS
#define VEON_FWD_WARPS 8
#endif

namespace veon {

// timing experiments only (profiles/README.md): -DVEON_FWD_EXPERIMENT + env VEON_FWD_DBG
//   1 no zero-fill  4 no global stores  8 no feature-row loads
#ifdef VEON_FWD_EXPERIMENT
#define VEON_DBG(bit) (dbg & (bit))
#else
#define VEON_DBG(bit) false
#endif

// x 2 CTAs/SM.  Keep it at 8: the 8 consecutive tiles of a CTA then cover an aligned 1 KB
// of every channel plane; 7-warp CTAs (896 B) write the same volume 19 % slower
// (tools/micro/store_pattern.cu: 266 vs 224 us).
constexpr int kFwdWarps = VEON_FWD_WARPS;
constexpr int kRowPitch = kTileVoxels + 4;   // 36 floats: rows stay 16-byte aligned

// ---- per-warp prefetch ring in shared memory (cp.async, no registers held) ----
// slot (ints): [0]=s [1]=e [2]=tile [3]=cbase, then one int4 per point (lane j):
//   landing   {ranks_bev, ranks_feat, ranks_depth, depth}
//   fixed up  {ranks_bev, row offset (floats), voxel | first<<8, depth}
#ifndef VEON_FWD_DIST
#define VEON_FWD_DIST 2
#endif
constexpr int kDist = VEON_FWD_DIST;         // prefetch distance in items per stage
constexpr int kRingSlots = 4 * kDist;        // bounds run 3*kDist ahead
constexpr int kSlotInts = 4 + 4 * 32;

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// every group except the (kDist-1) most recent ones has landed
__device__ __forceinline__ void cp_async_wait_dist() {
  asm volatile("cp.async.wait_group %0;" ::"n"(kDist - 1) : "memory");
}

// Persistent: every warp strides over (tile, channel-chunk) items.
//  * index prefetch: tile bounds -> rank triple -> depth value run 3 / 2 / 1 x kDist
//    items ahead as 4-byte cp.async copies into a per-warp ring (no registers held;
//    `wait_group kDist-1` leaves the youngest group in flight, so short items do not
//    expose the latency); a lanes=points fix-up turns the landed ranks into
//    (row byte offset, voxel | first-of-voxel) so that the gather loop needs ONE
//    broadcast LDS.128 per point.  The (tile, chunk, sample) of the next item is kept
//    as running counters: no integer division in the loop;
//  * gather: lanes = channels, up to 16 points' rows in flight (issued in groups of 4,
//    groups past the tile's last point are skipped), per-voxel sums in registers
//    (fma(feat, depth, acc) in rank order), one store per occupied voxel into the
//    zero-filled [c][36] shared tile;
//  * write-out: lanes = voxels, LDS.128 + one 16-byte streaming store covers
//    4 channel planes x 128 B per instruction.
// (A TMA tensor-store write-out was measured slower here: ~17 B/clk/SM for boxes
//  of 128-byte rows vs ~23 B/clk/SM for st.global.v4; see profiles/README.md.)
// Slot header: [0]=s [1]=e [2]=g0 (global voxel index of the tile's first voxel)
//              [3]=sample<<16 | chunk
// FULLC: C is a multiple of the chunk width, so no channel predicate anywhere.
template <int KCH, bool FULLC>
__global__ void __launch_bounds__(kFwdWarps * 32)
k_pool_fwd(const float* __restrict__ depth, const float* __restrict__ feat,
           const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
           const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ tile_start,
           const int32_t* __restrict__ heavy, uint32_t n_items, uint32_t tiles_per_sample,
           int64_t V, int C, uint32_t n_chunks, int vec_ok, float* __restrict__ out, int dbg) {
  constexpr int CC = 32 * KCH;
  constexpr int U = (KCH <= 2) ? 16 : 8;  // feature rows in flight per warp
  constexpr int kTileFloats = CC * kRowPitch;
  constexpr int kWarpFloats = kTileFloats + kRingSlots * kSlotInts;
  extern __shared__ __align__(16) float smem[];
  pdl_launch_dependents();  // the heavy-tile grid may be queued behind this one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* tile = smem + warp * kWarpFloats;
  int32_t* ring = reinterpret_cast<int32_t*>(tile + kTileFloats);
  const uint32_t TW = gridDim.x * kFwdWarps;
  const uint32_t first_item = blockIdx.x * kFwdWarps + warp;
  if (first_item >= n_items) return;
  const uint32_t my_items = (n_items - first_item + TW - 1) / TW;
  // tiles with at least this many points belong to k_pool_fwd_heavy
  const int32_t heavy_thr = heavy ? __ldg(heavy + 1) : 0x7fffffff;
  const uint32_t Vu = (uint32_t)V;                    // B*V < 2^31 (checked by the launcher)
  const bool whole_tiles = (Vu % kTileVoxels) == 0;   // no ragged tile at a sample's end

  // running decode of the next item to enter the ring (advanced by TW items per call)
  uint32_t nx_t = first_item / n_chunks;
  uint32_t nx_chunk = first_item - nx_t * n_chunks;
  uint32_t nx_b = nx_t / tiles_per_sample;
  uint32_t nx_vt = nx_t - nx_b * tiles_per_sample;
  const uint32_t step_t = TW / n_chunks, step_c = TW - step_t * n_chunks;

  auto slot_of = [&](uint32_t m) { return ring + (m & (kRingSlots - 1)) * kSlotInts; };
  auto issue_bounds = [&](uint32_t m) {  // header of item m; must be called for m = 0, 1, 2, ...
    int32_t* sl = slot_of(m);
    if (m < my_items) {
      if (lane < 2) cp_async4(sl + lane, tile_start + nx_t + lane);
      if (lane == 0) {
        const int32_t g0 = (int32_t)(nx_b * Vu + nx_vt * kTileVoxels);
        *reinterpret_cast<int2*>(sl + 2) = make_int2(g0, (int)((nx_b << 16) | nx_chunk));
      }
      uint32_t dt = step_t;
      nx_chunk += step_c;
      if (nx_chunk >= n_chunks) {
        nx_chunk -= n_chunks;
        ++dt;
      }
      nx_t += dt;
      nx_vt += dt;
      while (nx_vt >= tiles_per_sample) {
        nx_vt -= tiles_per_sample;
        ++nx_b;
      }
    } else if (lane < 2) {
      sl[lane] = 0;
    }
  };
  auto issue_ranks = [&](uint32_t m) {  // needs bounds(m)
    int32_t* sl = slot_of(m);
    int32_t* pt = sl + 4 + 4 * lane;
    const int2 se = *reinterpret_cast<const int2*>(sl);
    const int32_t i = se.x + lane;
    if (i < se.y) {
      cp_async4(pt + 0, ranks_bev + i);
      cp_async4(pt + 1, ranks_feat + i);
      cp_async4(pt + 2, ranks_depth + i);
    } else {
      pt[0] = -1;
    }
  };
  // needs ranks(m): depth gather + fix-up of the landed ranks (lanes = points)
  auto issue_depth = [&](uint32_t m) {
    int32_t* sl = slot_of(m);
    int32_t* pt = sl + 4 + 4 * lane;
    const int32_t rb = pt[0];
    if (rb >= 0) {
      cp_async4(pt + 3, depth + pt[2]);
      const int2 h = *reinterpret_cast<const int2*>(sl + 2);
      const int32_t up = lane ? pt[-4] : -1;
      // byte offset of this channel chunk of the point's feature row (< 2^32, checked)
      pt[1] = (int32_t)(((uint32_t)pt[1] * (uint32_t)C + (uint32_t)(h.y & 0xffff) * CC) * 4u);
      pt[2] = (rb - h.x) | ((rb != up) ? 0x100 : 0);  // voxel | first-of-voxel
    }
  };

  // prologue: fill the pipeline (three serialized latencies, once per warp)
#pragma unroll
  for (int i = 0; i < 3 * kDist; ++i) issue_bounds(i);
  cp_async_commit(); cp_async_wait_all(); __syncwarp();
#pragma unroll
  for (int i = 0; i < 2 * kDist; ++i) issue_ranks(i);
  cp_async_commit(); cp_async_wait_all(); __syncwarp();
#pragma unroll
  for (int i = 0; i < kDist; ++i) issue_depth(i);
  cp_async_commit(); cp_async_wait_all(); __syncwarp();

  const int q4 = (lane & 7) * 4, r = lane >> 3;  // write-out role of this lane
  float* const tlane = tile + lane * kRowPitch;   // flush base: channel = lane (+32k)
  const char* const feat_lane = reinterpret_cast<const char*>(feat) + lane * 4;
  const int64_t ostep = 4 * V;

  const uint32_t last_first = blockIdx.x * kFwdWarps + kFwdWarps - 1;
  const uint32_t common_items = last_first < n_items ? (n_items - last_first + TW - 1) / TW : 0;
  for (uint32_t m = 0; m < my_items; ++m) {
    if (VEON_DBG(16) && m < common_items) __syncthreads();
    cp_async_wait_dist();
    __syncwarp();
    int32_t* sl = slot_of(m);
    const int4 h = *reinterpret_cast<const int4*>(sl);  // s, e, g0, sample<<16 | chunk
    issue_bounds(m + 3 * kDist);  // lands in the slot item m-kDist used
    issue_ranks(m + 2 * kDist);
    issue_depth(m + kDist);
    cp_async_commit();
    const int32_t s0 = h.x, e0 = h.y, g0 = h.z;
    if (e0 - s0 >= heavy_thr) continue;

    const uint32_t b = (uint32_t)h.w >> 16;
    const int cbase = (h.w & 0xffff) * CC;
    const int cmax = FULLC ? CC : min(CC, C - cbase);
    // (b*C + cbase + r)*V + v0 + q4  with  v0 = g0 - b*V
    float* o = out + (int64_t)(b * (uint32_t)(C - 1) + (uint32_t)(cbase + r)) * V + g0 + q4;
    const bool fast = vec_ok && (whole_tiles || (uint32_t)g0 - b * Vu + kTileVoxels <= Vu);

    if (e0 <= s0 && fast) {  // empty tile
      if (!VEON_DBG(4)) {
#pragma unroll 4
        for (int c = r; c < cmax; c += 4, o += ostep) st_stream4(o, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }

    __syncwarp();
    if (!VEON_DBG(1)) {  // zero the tile (empty voxels must read as 0)
      float4* t4 = reinterpret_cast<float4*>(tile);
#pragma unroll
      for (int i = 0; i < kTileFloats / 4 / 32; ++i) t4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    float acc[KCH];
    int acc_vl = -1;
    const bool full_chunk = FULLC || (cmax == CC);
    for (int32_t base = s0; base < e0; base += 32) {
      const int cnt = min(32, e0 - base);
      if (base != s0) {  // tile with more than 32 points: later chunks are fetched synchronously
        const int32_t i = base + lane;
        __syncwarp();
        int32_t rb = -1, rf = 0;
        float d = 0.f;
        if (i < e0) {
          rb = __ldg(ranks_bev + i);
          rf = __ldg(ranks_feat + i);
          d = __ldg(depth + __ldg(ranks_depth + i));
        }
        const int32_t up = __shfl_up_sync(0xffffffffu, rb, 1);
        const bool first = (lane > 0) && (rb != up);  // lane 0 continues or starts: see below
        int32_t* pt = sl + 4 + 4 * lane;
        pt[1] = (int32_t)(((uint32_t)rf * (uint32_t)C + (uint32_t)cbase) * 4u);
        pt[2] = (rb - g0) | (first ? 0x100 : 0);
        pt[3] = __float_as_int(d);
        if (lane == 0) {  // first point of the chunk: new voxel iff it differs from the last one
          const int32_t last = sl[4 + 4 * 31 + 0];
          pt[2] = (rb - g0) | ((rb != last) ? 0x100 : 0);
        }
        __syncwarp();
        pt[0] = rb;
        __syncwarp();
      }
      const int4* pts = reinterpret_cast<const int4*>(sl + 4);
      for (int j0 = 0; j0 < cnt; j0 += U) {
        int4 p[U];
        float f[U][KCH];
#pragma unroll
        for (int g = 0; g < U / 4; ++g) {
          if (j0 + 4 * g < cnt) {  // warp-uniform: skip the groups past the last point
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const int u = 4 * g + uu;
              p[u] = pts[min(j0 + u, cnt - 1)];
              const float* row = reinterpret_cast<const float*>(feat_lane + (uint32_t)p[u].y);
#pragma unroll
              for (int k = 0; k < KCH; ++k)
                f[u][k] = VEON_DBG(8) ? (float)p[u].y
                          : (FULLC || full_chunk || lane + 32 * k < cmax) ? __ldg(row + 32 * k) : 0.f;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < U / 4; ++g) {
          if (j0 + 4 * g < cnt) {
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const int u = 4 * g + uu;
              if (j0 + u < cnt) {
                const float dj = __int_as_float(p[u].w);
                if (p[u].z & 0x100) {  // first point of its voxel (warp-uniform)
                  if (acc_vl >= 0) {
#pragma unroll
                    for (int k = 0; k < KCH; ++k) tlane[32 * k * kRowPitch + acc_vl] = acc[k];
                  }
                  acc_vl = p[u].z & 0xff;
#pragma unroll
                  for (int k = 0; k < KCH; ++k) acc[k] = fmaf(f[u][k], dj, 0.f);
                } else {
#pragma unroll
                  for (int k = 0; k < KCH; ++k) acc[k] = fmaf(f[u][k], dj, acc[k]);
                }
              }
            }
          }
        }
      }
    }
    if (acc_vl >= 0) {
#pragma unroll
      for (int k = 0; k < KCH; ++k) tlane[32 * k * kRowPitch + acc_vl] = acc[k];
    }
    __syncwarp();

    // Write-out: lane (r = lane/8, q = lane%8) moves voxels 4q..4q+3 of channel
    // 4*it + r with one LDS.128 + one 16-byte streaming store.
    const float* trow = tile + r * kRowPitch + q4;
    if (fast) {
#pragma unroll 4
      for (int c = r; c < cmax; c += 4, o += ostep, trow += 4 * kRowPitch) {
        const float4 v4 = *reinterpret_cast<const float4*>(trow);
        if (!VEON_DBG(4) || v4.x == 123.456f) st_stream4(o, v4);
      }
    } else {  // ragged volume edge / unaligned volume: scalar, bounds-checked
      const int v0 = (int)((uint32_t)g0 - b * Vu);
      for (int c = r; c < cmax; c += 4, o += ostep, trow += 4 * kRowPitch)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (v0 + q4 + i < V) st_stream(o + i, trow[i]);
    }
    // (the __syncwarp before the next zero-fill protects the tile reuse)
  }
  cp_async_wait_all();
}

// ---- heavy tiles -------------------------------------------------------------
// A tile with hundreds of points (voxels next to the cameras collect a whole
// frustum column each) would keep ONE warp of the kernel above busy for longer
// than an average warp's entire share of the volume, one exposed load latency per
// 16 points.  The plan lists such tiles and the kernel above skips them; here a
// whole CTA takes one: 128 points per round, ALL their feature rows in flight at
// once (cp.async into shared memory, the index records of the next two rounds
// prefetched in registers), then each warp runs the fma chains of its voxels out
// of shared memory.  Only the loads are parallelised -- every (voxel, channel) sum
// is still accumulated point by point in rank order (partial sums are carried in
// the shared tile between rounds), so the result stays bit-identical.
constexpr int kHeavyThreads = 256;
constexpr int kHeavyChunk = 128;  // points per round; one index record per thread < 128

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

template <int KCH>
__global__ void __launch_bounds__(kHeavyThreads)
k_pool_fwd_heavy(const float* __restrict__ depth, const float* __restrict__ feat,
                 const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
                 const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ tile_start,
                 const int32_t* __restrict__ heavy, int heavy_cap, uint32_t tiles_per_sample,
                 int64_t V, int C, uint32_t n_chunks, int vec_ok, float* __restrict__ out,
                 int join, int min_points) {
  constexpr int CC = 32 * KCH;
  constexpr int kSegs = CC / 4;  // 16-byte segments per staged row
  extern __shared__ __align__(16) float hsm[];
  float* rows = hsm;                                            // [kHeavyChunk][CC]
  float* tile = rows + kHeavyChunk * CC;                        // [CC][kRowPitch]
  float* dep = tile + CC * kRowPitch;                           // [kHeavyChunk]
  uint32_t* off = reinterpret_cast<uint32_t*>(dep + kHeavyChunk);  // [kHeavyChunk]
  int32_t* bounds = reinterpret_cast<int32_t*>(off + kHeavyChunk); // [2][start 32 | end 32]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // no griddepcontrol.wait: this grid and k_pool_fwd write disjoint tiles and read only
  // what earlier, normally launched work produced (see launch_fwd_impl)
  pdl_launch_dependents();
  const uint32_t n_heavy = (uint32_t)min(__ldg(heavy), heavy_cap);
  const uint32_t n_work = n_heavy * n_chunks;

  for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint32_t hi = w / n_chunks;
    const int cbase = (int)(w - hi * n_chunks) * CC;
    const uint32_t t = (uint32_t)__ldg(heavy + 2 + hi);
    const int32_t s = __ldg(tile_start + t);
    const int32_t n = __ldg(tile_start + t + 1) - s;
    if (n < min_points) continue;  // left to the main grid (k_pool_fwd_narrow takes more)
    const uint32_t b = t / tiles_per_sample;
    const int v0 = (int)(t - b * tiles_per_sample) * kTileVoxels;
    const int32_t g0 = (int32_t)((int64_t)b * V) + v0;
    const int cmax = min(CC, C - cbase);
    const int n_rounds = (n + kHeavyChunk - 1) / kHeavyChunk;

    // index records, one point per thread < kHeavyChunk, two rounds deep in registers:
    //   stage A(k): ranks of round k            (rf, rb, previous point's rb, rd)
    //   stage B(k): depth[rd] of round k        (needs A(k))
    int32_t rf1 = 0, rb1 = -1, rp1 = -1, rd1 = 0, rf2 = 0, rb2 = -1, rp2 = -1, rd2 = 0;
    float d1 = 0.f;
    auto stage_a = [&](int k, int32_t& rf, int32_t& rb, int32_t& rp, int32_t& rd) {
      const int32_t i = k * kHeavyChunk + tid;
      rb = -1;
      if (tid < kHeavyChunk && i < n) {
        rf = __ldg(ranks_feat + s + i);
        rd = __ldg(ranks_depth + s + i);
        rb = __ldg(ranks_bev + s + i);
        rp = i ? __ldg(ranks_bev + s + i - 1) : -1;
      }
    };
    stage_a(0, rf1, rb1, rp1, rd1);
    if (n_rounds > 1) stage_a(1, rf2, rb2, rp2, rd2);
    if (rb1 >= 0) d1 = __ldg(depth + rd1);

    {  // clear the tile and both boundary tables
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int i = tid; i < CC * kRowPitch / 4; i += kHeavyThreads)
        t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tid < 128) bounds[tid] = 0;
    }
    __syncthreads();

    for (int k = 0; k < n_rounds; ++k) {
      const int cnt = min(kHeavyChunk, n - k * kHeavyChunk);
      int32_t* cstart = bounds + (k & 1) * 64;
      int32_t* cend = cstart + 32;
      // publish round k's records and the [start, end) of every voxel inside the round
      if (rb1 >= 0) {
        const int vox = rb1 - g0, voxp = rp1 - g0;
        off[tid] = (uint32_t)rf1 * (uint32_t)C;
        dep[tid] = d1;
        const bool first = (tid == 0) || (rb1 != rp1);
        if (first) cstart[vox] = tid;
        if (first && tid > 0) cend[voxp] = tid;
        if (tid == cnt - 1) cend[vox] = cnt;
      }
      // advance the register pipeline: B(k+1) and A(k+2) fly during this round's row copy
      rf1 = rf2; rb1 = rb2; rp1 = rp2; rd1 = rd2;
      d1 = 0.f;
      if (rb1 >= 0) d1 = __ldg(depth + rd1);
      rb2 = -1;
      if (k + 2 < n_rounds) stage_a(k + 2, rf2, rb2, rp2, rd2);
      // the other table is free now (last read in round k-1): clear it for round k+1
      if (tid >= 128 && tid < 192) bounds[((k + 1) & 1) * 64 + tid - 128] = 0;
      __syncthreads();

      // every feature row of the round in flight at once
#pragma unroll
      for (int i = 0; i < kHeavyChunk * kSegs / kHeavyThreads; ++i) {
        const int idx = tid + kHeavyThreads * i;
        const int r = idx / kSegs, seg = (idx % kSegs) * 4;
        if (r < cnt && seg < cmax) cp_async16(rows + r * CC + seg, feat + off[r] + cbase + seg);
      }
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();

      // fma chains: warp w owns voxels w, w+8, w+16, w+24; lanes = channels
#pragma unroll 1
      for (int v = warp; v < kTileVoxels; v += kHeavyThreads / 32) {
        const int a = cstart[v], e = cend[v];
        if (e <= a) continue;
        float acc[KCH];
#pragma unroll
        for (int c = 0; c < KCH; ++c) acc[c] = tile[(lane + 32 * c) * kRowPitch + v];
#pragma unroll 4
        for (int j = a; j < e; ++j) {
          const float dj = dep[j];
#pragma unroll
          for (int c = 0; c < KCH; ++c) acc[c] = fmaf(rows[j * CC + lane + 32 * c], dj, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < KCH; ++c) tile[(lane + 32 * c) * kRowPitch + v] = acc[c];
      }
      __syncthreads();
    }

    {  // write-out: thread (row = tid/8, q = tid%8) moves voxels 4q..4q+3 of channel row (+32j)
      const int q4 = (tid & 7) * 4;
      const bool fast = vec_ok && (v0 + kTileVoxels <= V);
      for (int c = tid >> 3; c < cmax; c += kHeavyThreads / 8) {
        float* o = out + ((int64_t)b * C + cbase + c) * V + v0 + q4;
        const float* trow = tile + c * kRowPitch + q4;
        if (fast) {
          st_stream4(o, *reinterpret_cast<const float4*>(trow));
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (v0 + q4 + i < V) st_stream(o + i, trow[i]);
        }
      }
    }
    __syncthreads();
  }
  // join != 0: do not COMPLETE before the grid this one is a programmatic dependent of.  Needed
  // when the launches are captured into a CUDA graph, where the next node would depend on this
  // grid only; on a plain stream the next launch waits for everything before it anyway, and the
  // wait would put the main grid's memory flush on this grid's critical path (+17 us).
  if (join) pdl_wait();
}


// ---- tile-group kernel (experimental: VEON_FWD_GROUP=1) -------------------------------
// A CTA takes 8 consecutive tiles (256 voxels = one aligned 1 KB run of every channel plane)
// at a time: the index records of the group's points are prefetched one group ahead in
// registers (thread per point), ALL their feature rows are gathered by 16-byte cp.async into
// shared memory at once, warp w then runs the rank-ordered fma chains of tile w's voxels out
// of shared memory into a [c][260] staging tile, and the whole CTA writes the 64 x 1 KB rows
// together.  Compared with the warp-per-tile kernel: no per-point address arithmetic in the
// gather, no exposed row-load latency per 16 points, and the store stream leaves the SM in
// aligned 1 KB runs.  Heavy tiles are skipped exactly as in k_pool_fwd.
#ifndef VEON_GROUP_ROUND
#define VEON_GROUP_ROUND 64
#endif
#ifndef VEON_GROUP_TILES
#define VEON_GROUP_TILES 4
#endif
constexpr int kGroupTiles = VEON_GROUP_TILES;    // 4 or 8: one warp per tile
constexpr int kGroupVoxels = kGroupTiles * kTileVoxels;
#ifndef VEON_GROUP_THREADS
#define VEON_GROUP_THREADS (32 * VEON_GROUP_TILES)
#endif
constexpr int kGroupThreads = VEON_GROUP_THREADS;  // >= kGroupRound (thread per point)
constexpr int kGroupRound = VEON_GROUP_ROUND;    // points staged per round (thread per point)
constexpr int kStagePitch = kGroupVoxels + 4;    // floats; rows stay 16-byte aligned

template <int KCH>
__global__ void __launch_bounds__(kGroupThreads)
k_pool_fwd_group(const float* __restrict__ depth, const float* __restrict__ feat,
                 const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
                 const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ tile_start,
                 const int32_t* __restrict__ heavy, uint32_t n_items, uint32_t n_chunks,
                 uint32_t groups_per_sample, int64_t V, int C, float* __restrict__ out) {
  constexpr int CC = 32 * KCH;
  constexpr int kSegs = CC / 4;
  constexpr int kW = kGroupThreads / 32;
  extern __shared__ __align__(16) float gsm[];
  float* rows = gsm;                                                 // [2][kGroupRound][CC]
  float* stage = rows + 2 * kGroupRound * CC;                        // [CC][kStagePitch]
  float* dep = stage + CC * kStagePitch;                             // [2][kGroupRound]
  uint32_t* off = reinterpret_cast<uint32_t*>(dep + 2 * kGroupRound);  // [2][kGroupRound]
  int32_t* b0 = reinterpret_cast<int32_t*>(off + 2 * kGroupRound);   // [3][start | end] round 0
  int32_t* bx = b0 + 6 * kGroupVoxels;                               // [start | end] extra rounds
  int32_t* hdrs = bx + 2 * kGroupVoxels;                             // [4][12]: tile_start[0..T]
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (blockIdx.x >= n_items) return;
  const uint32_t my_items = (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const int32_t heavy_thr = heavy ? __ldg(heavy + 1) : 0x7fffffff;

  auto item_group = [&](uint32_t k, uint32_t& g, uint32_t& chunk) {
    const uint32_t item = blockIdx.x + k * gridDim.x;
    g = item / n_chunks;
    chunk = item - g * n_chunks;
  };
  auto load_hdr = [&](uint32_t k) -> int32_t {   // thread j <= T: tile_start[T g + j] of item k
    if (tid <= kGroupTiles && k < my_items) {
      uint32_t g, chunk;
      item_group(k, g, chunk);
      return __ldg(tile_start + (int64_t)g * kGroupTiles + tid);
    }
    return 0;
  };
  auto hdr_of = [&](uint32_t k) { return hdrs + (k & 3) * 12; };
  // compacted position q of a group (heavy tiles left out) -> point index, index of the
  // previous compacted point (-1 for q == 0); returns the group's number of points
  auto locate = [&](const int32_t* hdr, int q, int32_t& i, int32_t& iprev) -> int {
    int acc = 0;
    int32_t last = -1;
    i = -1;
    iprev = -1;
#pragma unroll
    for (int j = 0; j < kGroupTiles; ++j) {
      const int32_t s = hdr[j];
      int n = hdr[j + 1] - s;
      if (n >= heavy_thr) n = 0;
      if (i < 0 && q < acc + n) {
        i = s + (q - acc);
        iprev = (q == acc) ? last : i - 1;
      }
      if (n > 0) last = s + n - 1;
      acc += n;
    }
    return acc;
  };
  struct Rec { int32_t rb, rp, rf, rd; float d; };
  auto stage_a = [&](const int32_t* hdr, int q, Rec& r) {
    r.rb = -1;
    r.rp = -1;
    r.rd = 0;
    if (tid < kGroupRound) {
      int32_t i, ip;
      locate(hdr, q, i, ip);
      if (i >= 0) {
        r.rb = __ldg(ranks_bev + i);
        r.rf = __ldg(ranks_feat + i);
        r.rd = __ldg(ranks_depth + i);
        if (ip >= 0) r.rp = __ldg(ranks_bev + ip);
      }
    }
  };
  auto stage_b = [&](Rec& r) {
    r.d = 0.f;
    if (r.rb >= 0) r.d = __ldg(depth + r.rd);
  };
  auto publish = [&](const Rec& r, int cnt, int32_t g0, uint32_t chunk, int slot,
                     int32_t* cstart) {
    int32_t* cend = cstart + kGroupVoxels;
    if (r.rb >= 0) {
      const int vox = r.rb - g0;
      off[slot * kGroupRound + tid] = ((uint32_t)r.rf * (uint32_t)C + chunk * CC) * 4u;
      dep[slot * kGroupRound + tid] = r.d;
      const bool starts = (tid == 0) || (r.rb != r.rp);
      if (starts) cstart[vox] = tid;
      if (tid > 0 && r.rb != r.rp) cend[r.rp - g0] = tid;
      if (tid == cnt - 1) cend[vox] = cnt;
    }
  };
  auto gather_rows = [&](int cnt, int slot) {   // every feature row of the round at once
    const char* fbase = reinterpret_cast<const char*>(feat);
    float* dst = rows + slot * kGroupRound * CC;
    const uint32_t* o = off + slot * kGroupRound;
#pragma unroll
    for (int i = 0; i < kGroupRound * kSegs / kGroupThreads; ++i) {
      const int idx = tid + kGroupThreads * i;
      const int r = idx / kSegs, seg = (idx % kSegs) * 4;
      if (r < cnt) cp_async16(dst + r * CC + seg, fbase + o[r] + seg * 4);
    }
    cp_async_commit();
  };
  // warp w takes voxels w, w + W, ...: a tile's points spread over all warps; lanes = channels
  auto accumulate = [&](const int32_t* cstart, int slot) {
    const int32_t* cend = cstart + kGroupVoxels;
    const float* rw = rows + slot * kGroupRound * CC;
    const float* dp = dep + slot * kGroupRound;
    const int mine = kW * lane + warp;
    const bool has = mine < kGroupVoxels;
    const int a_l = has ? cstart[mine] : 0, e_l = has ? cend[mine] : 0;
    uint32_t m = __ballot_sync(0xffffffffu, e_l > a_l);
    while (m) {
      const int l = __ffs(m) - 1;
      m &= m - 1;
      const int a = __shfl_sync(0xffffffffu, a_l, l), e = __shfl_sync(0xffffffffu, e_l, l);
      float* sp = stage + lane * kStagePitch + kW * l + warp;
      float acc[KCH];
#pragma unroll
      for (int c = 0; c < KCH; ++c) acc[c] = sp[32 * c * kStagePitch];
#pragma unroll 2
      for (int j = a; j < e; ++j) {
        const float dj = dp[j];
#pragma unroll
        for (int c = 0; c < KCH; ++c) acc[c] = fmaf(rw[j * CC + lane + 32 * c], dj, acc[c]);
      }
#pragma unroll
      for (int c = 0; c < KCH; ++c) sp[32 * c * kStagePitch] = acc[c];
    }
  };
  auto group_total = [&](const int32_t* hdr) {
    int32_t i, ip;
    return locate(hdr, 0x7fffffff, i, ip);
  };

  // ---- prologue: headers 0..2 published, rows of item 0 in flight, records of item 1
  // complete, ranks of item 2 requested
  {
    const int32_t h0 = load_hdr(0), h1 = load_hdr(1), h2 = load_hdr(2);
    if (tid <= kGroupTiles) {
      hdr_of(0)[tid] = h0;
      hdr_of(1)[tid] = h1;
      hdr_of(2)[tid] = h2;
    }
    for (int i = tid; i < 8 * kGroupVoxels; i += kGroupThreads) b0[i] = 0;   // b0 and bx
  }
  int32_t hreg = load_hdr(3);
  __syncthreads();
  Rec r1, r2, r3;
  {
    uint32_t g, chunk;
    item_group(0, g, chunk);
    Rec r0;
    stage_a(hdr_of(0), tid, r0);
    stage_b(r0);
    const int total0 = group_total(hdr_of(0));
    publish(r0, min(total0, kGroupRound), (int32_t)(g * kGroupVoxels), chunk, 0, b0);
    __syncthreads();
    gather_rows(min(total0, kGroupRound), 0);
  }
  stage_a(hdr_of(1), tid, r1);
  stage_b(r1);
  stage_a(hdr_of(2), tid, r2);

  // per-item scalars are computed once, one item ahead: (group, chunk), point total, heavy mask
  auto describe = [&](uint32_t k, uint32_t& g, uint32_t& chunk, int& total, uint32_t& hmask) {
    g = chunk = 0;
    total = 0;
    hmask = 0;
    if (k >= my_items) return;
    item_group(k, g, chunk);
    const int32_t* hdr = hdr_of(k);
#pragma unroll
    for (int j = 0; j < kGroupTiles; ++j) {
      const int n = hdr[j + 1] - hdr[j];
      if (n >= heavy_thr) hmask |= 1u << j; else total += n;
    }
  };
  uint32_t g, chunk, g1, chunk1, heavy_mask, heavy_mask1;
  int total, total1;
  describe(0, g, chunk, total, heavy_mask);
  int t0 = 0, t1 = 1, t2 = 2;                           // round-0 tables of items k, k+1, k+2

  for (uint32_t k = 0; k < my_items; ++k) {
    const int32_t* hdr = hdr_of(k);
    const uint32_t b = g / groups_per_sample;
    const int32_t g0 = (int32_t)(g * kGroupVoxels);     // global voxel index (V % 32 == 0)
    const int cbase = (int)chunk * CC;
    const int cur = k & 1, nxt = cur ^ 1;

    // T0: publish the records of item k+1 (prefetched), the header of item k+3
    describe(k + 1, g1, chunk1, total1, heavy_mask1);
    if (k + 1 < my_items)
      publish(r1, min(total1, kGroupRound), (int32_t)(g1 * kGroupVoxels), chunk1, nxt,
              b0 + t1 * 2 * kGroupVoxels);
    if (tid <= kGroupTiles) hdr_of(k + 3)[tid] = hreg;
    hreg = load_hdr(k + 4);
    __syncthreads();

    // T1: rows of item k+1 go in flight (consumed one iteration later), ranks of item k+3 and
    // depths of item k+2 are requested, the staging tile is cleared; then wait for the rows of
    // item k, which were requested one iteration ago
    gather_rows(min(total1, kGroupRound), nxt);
    stage_a(hdr_of(k + 3), tid, r3);
    stage_b(r2);
    if (total > 0) {
      float4* s4 = reinterpret_cast<float4*>(stage);
#pragma unroll
      for (int i = 0; i < (CC * kStagePitch / 4 + kGroupThreads - 1) / kGroupThreads; ++i)
        if (tid + i * kGroupThreads < CC * kStagePitch / 4)
          s4[tid + i * kGroupThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // the round-0 table item k-1 used is free: clear it for item k+2
    for (int i = tid; i < 2 * kGroupVoxels; i += kGroupThreads) b0[t2 * 2 * kGroupVoxels + i] = 0;
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();

    // T2: fma chains of item k; rounds beyond the first are fetched synchronously
    if (total > 0) accumulate(b0 + t0 * 2 * kGroupVoxels, cur);
    for (int base = kGroupRound; base < total; base += kGroupRound) {
      const int cnt = min(kGroupRound, total - base);
      Rec r;
      stage_a(hdr, base + tid, r);
      stage_b(r);
      __syncthreads();                                   // previous round fully consumed
      publish(r, cnt, g0, chunk, cur, bx);
      __syncthreads();
      gather_rows(cnt, cur);
      cp_async_wait_all();
      __syncthreads();
      accumulate(bx, cur);
      __syncthreads();
      for (int i = tid; i < 2 * kGroupVoxels; i += kGroupThreads) bx[i] = 0;
    }
    __syncthreads();

    // T3: CC planes x (kGroupVoxels * 4) bytes; warp w takes planes w, w+W, ...; a lane's
    // 16 bytes of half h lie in tile 4 h + lane / 8
    {
      constexpr int kHalves = kGroupVoxels / 128;
      float* o = out + ((int64_t)b * C + cbase + warp) * V + (g0 - (int32_t)(b * (uint32_t)V)) +
                 4 * lane;
      const int64_t ostep = (int64_t)kW * V;
      const float* sp = stage + warp * kStagePitch + 4 * lane;
      bool skip[kHalves];
#pragma unroll
      for (int h = 0; h < kHalves; ++h) skip[h] = (heavy_mask >> (4 * h + (lane >> 3))) & 1u;
      if (total > 0) {
#pragma unroll 4
        for (int c = 0; c < CC / kW; ++c, o += ostep) {
#pragma unroll
          for (int h = 0; h < kHalves; ++h)
            if (!skip[h])
              st_stream4(o + 128 * h,
                         *reinterpret_cast<const float4*>(sp + c * kW * kStagePitch + 128 * h));
        }
      } else {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int c = 0; c < CC / kW; ++c, o += ostep) {
#pragma unroll
          for (int h = 0; h < kHalves; ++h)
            if (!skip[h]) st_stream4(o + 128 * h, z);
        }
      }
    }
    r1 = r2;
    r2 = r3;
    g = g1; chunk = chunk1; total = total1; heavy_mask = heavy_mask1;
    { const int t = t0; t0 = t1; t1 = t2; t2 = t; }
  }
  cp_async_wait_all();
}

// ---- fused pooling + 2x2x2 max-downsample, forward ------------------------
// VEON's neck reduces the pooled volume 8x right away (view_transformer_raw.py:549-553:
// view(b,c,z/2,2,y/2,2,x/2,2).amax).  Here the full-resolution volume is never written: a CTA
// takes one output row (b, z/2, y/2), pools its four input x-rows one after the other into a
// shared [c][X] tile -- points of a row are one contiguous slice of the sorted rank arrays,
// found through the per-voxel prefix `voxel_start` the preparation leaves in its workspace;
// rows staged by cp.async, rank-ordered fma chains as in k_pool_fwd_heavy -- and folds each
// into a running [c][X/2] maximum.  Sums are the same bits as the unfused kernel's and max is
// exact, so the result equals pool + amax bit for bit.  Forward only (inference path).
__device__ __forceinline__ float ds_max(float a, float b) { return (b > a || b != b) ? b : a; }
constexpr int kDsThreads = 256;
constexpr int kDsRound = 96;   // points staged per round (thread per point)

template <int KCH>
__global__ void __launch_bounds__(kDsThreads)
k_pool_ds_fwd(const float* __restrict__ depth, const float* __restrict__ feat,
              const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
              const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ voxel_start,
              uint32_t n_items, uint32_t n_chunks, int Z, int Y, int Xfull, int xsplit,
              int C, float* __restrict__ out) {
  constexpr int CC = 32 * KCH;
  constexpr int kSegs = CC / 4;
  const int X = Xfull / xsplit;              // voxels of the x-run one item covers
  const int Xp = (X + 3) / 4 * 4 + 4;        // pitch of the full-resolution row tile
  const int Xh = X / 2, Xhp = Xh + 1;        // pitch of the running maximum
  extern __shared__ __align__(16) float dsm[];
  float* rows = dsm;                                                // [kDsRound][CC]
  float* stage = rows + kDsRound * CC;                              // [CC][Xp]
  float* outb = stage + CC * Xp;                                    // [CC][Xhp]
  float* dep = outb + CC * Xhp;                                     // [kDsRound]
  uint32_t* off = reinterpret_cast<uint32_t*>(dep + kDsRound);      // [kDsRound]
  int32_t* bounds = reinterpret_cast<int32_t*>(off + kDsRound);     // [2][start Xp | end Xp]
  int32_t* rs = bounds + 4 * Xp;                                    // [4] first point of a row
  int32_t* rn = rs + 4;                                             // [4] points in the row
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Zh = Z / 2, Yh = Y / 2;
  const int64_t V = (int64_t)Z * Y * Xfull;

  struct Rec { int32_t rb, rp, rf, rd; float d; };
  struct Rnd { int sub, base; };
  auto first_round = [&]() -> Rnd {
    for (int q = 0; q < 4; ++q) if (rn[q] > 0) return Rnd{q, 0};
    return Rnd{4, 0};
  };
  auto next_round = [&](Rnd r) -> Rnd {
    if (r.base + kDsRound < rn[r.sub]) return Rnd{r.sub, r.base + kDsRound};
    for (int q = r.sub + 1; q < 4; ++q) if (rn[q] > 0) return Rnd{q, 0};
    return Rnd{4, 0};
  };
  auto stage_a = [&](Rnd r, Rec& c) {
    c.rb = -1;
    c.rp = -1;
    const int q = r.base + tid;
    if (tid < kDsRound && q < rn[r.sub]) {
      const int32_t i = rs[r.sub] + q;
      c.rb = __ldg(ranks_bev + i);
      c.rf = __ldg(ranks_feat + i);
      c.rd = __ldg(ranks_depth + i);
      if (q > 0) c.rp = __ldg(ranks_bev + i - 1);
    }
  };
  auto stage_b = [&](Rec& c) {
    c.d = 0.f;
    if (c.rb >= 0) c.d = __ldg(depth + c.rd);
  };

  for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t piece = item / n_chunks, chunk = item - piece * n_chunks;
    const uint32_t cell = piece / (uint32_t)xsplit;
    const int x0 = (int)(piece - cell * (uint32_t)xsplit) * X;   // first input voxel of the run
    const int yo = (int)(cell % (uint32_t)Yh);
    const int zo = (int)((cell / (uint32_t)Yh) % (uint32_t)Zh);
    const int64_t b = cell / ((uint32_t)Yh * (uint32_t)Zh);
    const int cbase = (int)chunk * CC;
    __syncthreads();   // previous item fully written out
    if (tid < 4) {
      const int64_t g = b * V + ((int64_t)(2 * zo + (tid >> 1)) * Y + 2 * yo + (tid & 1)) * Xfull + x0;
      const int32_t s0 = __ldg(voxel_start + g);
      rs[tid] = s0;
      rn[tid] = __ldg(voxel_start + g + X) - s0;
    }
    for (int i = tid; i < 4 * Xp; i += kDsThreads) bounds[i] = 0;
    __syncthreads();

    Rnd cur = first_round();
    Rec rec;
    rec.rb = -1;
    if (cur.sub < 4) {
      stage_a(cur, rec);
      stage_b(rec);
    }
    int parity = 0;
    for (int sub = 0; sub < 4; ++sub) {
      if (rn[sub] == 0) {   // empty input row: contributes zeros
        for (int c = warp; c < CC; c += kDsThreads / 32)
          for (int xo = lane; xo < Xh; xo += 32)
            outb[c * Xhp + xo] = sub == 0 ? 0.f : ds_max(outb[c * Xhp + xo], 0.f);
        continue;
      }
      const int32_t g_row = (int32_t)(b * V + ((int64_t)(2 * zo + (sub >> 1)) * Y + 2 * yo + (sub & 1)) * Xfull + x0);
      {
        float4* s4 = reinterpret_cast<float4*>(stage);
        for (int i = tid; i < CC * Xp / 4; i += kDsThreads) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      while (cur.sub == sub) {
        const int cnt = min(kDsRound, rn[sub] - cur.base);
        int32_t* cstart = bounds + parity * 2 * Xp;
        int32_t* cend = cstart + Xp;
        if (rec.rb >= 0) {   // publish the round's records and voxel [start, end) ranges
          const int vox = rec.rb - g_row;
          off[tid] = ((uint32_t)rec.rf * (uint32_t)C + chunk * CC) * 4u;
          dep[tid] = rec.d;
          const bool starts = (tid == 0) || (rec.rb != rec.rp);
          if (starts) cstart[vox] = tid;
          if (tid > 0 && rec.rb != rec.rp) cend[rec.rp - g_row] = tid;
          if (tid == cnt - 1) cend[vox] = cnt;
        }
        const Rnd nx = next_round(cur);
        __syncthreads();
        {
          const char* fbase = reinterpret_cast<const char*>(feat);
#pragma unroll
          for (int i = 0; i < kDsRound * kSegs / kDsThreads; ++i) {
            const int idx = tid + kDsThreads * i;
            const int r = idx / kSegs, seg = (idx % kSegs) * 4;
            if (r < cnt) cp_async16(rows + r * CC + seg, fbase + off[r] + seg * 4);
          }
          cp_async_commit();
        }
        Rec nrec;
        nrec.rb = -1;
        if (nx.sub < 4) stage_a(nx, nrec);
        for (int i = tid; i < 2 * Xp; i += kDsThreads) bounds[(parity ^ 1) * 2 * Xp + i] = 0;
        cp_async_wait_all();
        __syncthreads();
        if (nx.sub < 4) stage_b(nrec);
        // fma chains: voxel block k (32 voxels) belongs to warp k % 8; lanes = channels
        for (int vb = warp * 32; vb < X; vb += kDsThreads) {
          const int vv = vb + lane;
          const int a_l = vv < X ? cstart[vv] : 0, e_l = vv < X ? cend[vv] : 0;
          uint32_t m = __ballot_sync(0xffffffffu, e_l > a_l);
          while (m) {
            const int v = __ffs(m) - 1;
            m &= m - 1;
            const int a = __shfl_sync(0xffffffffu, a_l, v), e = __shfl_sync(0xffffffffu, e_l, v);
            float* sp = stage + lane * Xp + vb + v;
            float acc[KCH];
#pragma unroll
            for (int c = 0; c < KCH; ++c) acc[c] = sp[32 * c * Xp];
#pragma unroll 4
            for (int j = a; j < e; ++j) {
              const float dj = dep[j];
#pragma unroll
              for (int c = 0; c < KCH; ++c) acc[c] = fmaf(rows[j * CC + lane + 32 * c], dj, acc[c]);
            }
#pragma unroll
            for (int c = 0; c < KCH; ++c) sp[32 * c * Xp] = acc[c];
          }
        }
        parity ^= 1;
        rec = nrec;
        cur = nx;
        __syncthreads();
      }
      // fold the finished row into the running maximum (pairs along x)
      // (warp w: channels w, w+8, ...; lanes along x -- the same element-to-thread map in
      //  every fold and in the write-out, so outb needs no barrier of its own)
      for (int c = warp; c < CC; c += kDsThreads / 32)
        for (int xo = lane; xo < Xh; xo += 32) {
          const float2 p2 = *reinterpret_cast<const float2*>(stage + c * Xp + 2 * xo);
          const float m2 = ds_max(p2.x, p2.y);
          outb[c * Xhp + xo] = sub == 0 ? m2 : ds_max(outb[c * Xhp + xo], m2);
        }
      __syncthreads();
    }
    // out[b, cbase + c, zo, yo, :]
    for (int c = warp; c < CC; c += kDsThreads / 32) {
      float* o = out + (((b * C + cbase + c) * Zh + zo) * Yh + yo) * (int64_t)(Xfull / 2) + x0 / 2;
      for (int xo = lane; xo < Xh; xo += 32) o[xo] = outb[c * Xhp + xo];
    }
  }
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}


// ---- narrow rows (C <= 32): lane = voxel ---------------------------------------------------
// With few channels a lane-per-channel warp leaves lanes idle and pays one row-load latency
// per point (the logit-space lift pools Q + 2 = 20 channels over 4x the points of C2:
// 437 us for 410 MB).  Here a warp still owns a 32-voxel tile, but a LANE owns a voxel: it
// walks its own contiguous slice of the sorted points, two points per trip with all their
// loads (ranks, depth, the 16-byte pieces of both rows) in flight together, and accumulates
// every channel of its voxel in registers with the same fma(feat, depth, acc) chain in rank
// order, so the result is the same bits as k_pool_fwd's.  The write-out needs no transposition:
// for each channel plane the 32 lanes store one aligned 128-byte run (zeros for empty voxels).
// Segment bounds come from ranks_bev itself (head flags over the tile's points, 32 at a time).
// A lane walks two points per load latency, so this kernel keeps every tile below
// kNarrowHeavyMin points (at C3 density the plan's threshold of 96 would hand half of all
// points to the CTA-per-tile kernel: 252 us of a 437 us call); k_pool_fwd_heavy skips those.
constexpr int kNarrowWarps = 8;
constexpr int kNarrowHeavyMin = 512;
template <int NV>  // 16-byte pieces per feature row: C = 4 * NV
__global__ void __launch_bounds__(kNarrowWarps * 32, (NV <= 5) ? 3 : 2)
k_pool_fwd_narrow(const float* __restrict__ depth, const float* __restrict__ feat,
                  const int32_t* __restrict__ ranks_depth, const int32_t* __restrict__ ranks_feat,
                  const int32_t* __restrict__ ranks_bev, const int32_t* __restrict__ tile_start,
                  const int32_t* __restrict__ heavy, uint32_t n_tiles, uint32_t tiles_per_sample,
                  int64_t V, float* __restrict__ out, uint32_t zero, int heavy_min) {
  constexpr int C = 4 * NV;
  __shared__ int32_t seg_s[kNarrowWarps][32];
  pdl_launch_dependents();  // the heavy-tile grid may be queued behind this one
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t* seg = seg_s[warp];
  const int32_t heavy_thr = heavy ? max(__ldg(heavy + 1), heavy_min) : 0x7fffffff;
  const uint32_t TW = gridDim.x * kNarrowWarps;
  // The per-tile chain (bounds -> ranks_bev -> segment table -> ranks -> rows -> stores) is a
  // string of dependent load latencies, so the head of the chain runs ahead: bounds are read
  // two tiles early, the first 32 ranks_bev of a tile one tile early.
  const uint32_t first = blockIdx.x * kNarrowWarps + warp;
  int32_t ns = 0, ne = 0, as = 0, ae = 0, nrb = 0;  // next tile, the one after; next tile's ranks
  if (first < n_tiles) {
    ns = __ldg(tile_start + first);
    ne = __ldg(tile_start + first + 1);
    if (ns + lane < ne) nrb = __ldg(ranks_bev + ns + lane);
  }
  if (first + TW < n_tiles) {
    as = __ldg(tile_start + first + TW);
    ae = __ldg(tile_start + first + TW + 1);
  }
  for (uint32_t t = first; t < n_tiles; t += TW) {
    const int32_t s0 = ns, e0 = ne, rb_first = nrb;
    ns = as;
    ne = ae;
    as = ae = 0;
    if (t + 2 * TW < n_tiles) {
      as = __ldg(tile_start + t + 2 * TW);
      ae = __ldg(tile_start + t + 2 * TW + 1);
    }
    nrb = (ns + lane < ne) ? __ldg(ranks_bev + ns + lane) : 0;
    if (e0 - s0 >= heavy_thr) continue;  // k_pool_fwd_heavy's
    const uint32_t b = t / tiles_per_sample;
    const uint32_t v0 = (t - b * tiles_per_sample) * kTileVoxels;
    const int32_t g0 = (int32_t)((int64_t)b * V) + (int32_t)v0;  // rank of the tile's first voxel
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    if (e0 > s0) {
      // first point of every occupied voxel
      seg[lane] = -1;
      __syncwarp();
      int32_t prev;
      {
        const int32_t vox = (s0 + lane < e0) ? rb_first - g0 : -2;
        const int32_t up = __shfl_up_sync(0xffffffffu, vox, 1);
        if (s0 + lane < e0 && (lane == 0 || vox != up) && (uint32_t)vox < 32u) seg[vox] = s0 + lane;
        prev = __shfl_sync(0xffffffffu, vox, 31);
      }
      for (int32_t p0 = s0 + 32; p0 < e0; p0 += 128) {  // four rounds' loads in flight together
        int32_t vx[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int32_t i = p0 + 32 * r + lane;
          vx[r] = (i < e0) ? __ldg(ranks_bev + i) - g0 : -2;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int32_t i = p0 + 32 * r + lane;
          int32_t up = __shfl_up_sync(0xffffffffu, vx[r], 1);
          if (lane == 0) up = prev;
          if (i < e0 && vx[r] != up && (uint32_t)vx[r] < 32u) seg[vx[r]] = i;
          prev = __shfl_sync(0xffffffffu, vx[r], 31);
        }
      }
      __syncwarp();
      const int32_t st = seg[lane];
      const uint32_t occ = __ballot_sync(0xffffffffu, st >= 0);
      const uint32_t higher = (lane == 31) ? 0u : (occ >> (lane + 1));
      const int nxt = higher ? lane + __ffs(higher) : 0;
      const int32_t nst = __shfl_sync(0xffffffffu, st, nxt);
      const int32_t en = higher ? nst : e0;
      __syncwarp();  // seg is rewritten for the next tile
      if (st >= 0) {
        int32_t i = st;
        bool two = i + 1 < en;
        int32_t r0 = __ldg(ranks_depth + i), f0 = __ldg(ranks_feat + i);
        int32_t r1 = two ? __ldg(ranks_depth + i + 1) : r0;
        int32_t f1 = two ? __ldg(ranks_feat + i + 1) : f0;
        while (true) {
          // A trip always runs both fma chains (no branch for the scheduler to sink the second
          // point's loads into); a missing second point is (+0) * (-0): acc + (-0) == acc bit for
          // bit, whatever acc is.
          const float d0 = __ldg(depth + r0), d1 = two ? __ldg(depth + r1) : -0.f;
          const float4* row0 = reinterpret_cast<const float4*>(feat + (int64_t)f0 * C);
          const float4* row1 = reinterpret_cast<const float4*>(feat + (int64_t)f1 * C);
          float4 a[NV], q[NV];
#pragma unroll
          for (int k = 0; k < NV; ++k) a[k] = __ldg(row0 + k);
#pragma unroll
          for (int k = 0; k < NV; ++k) q[k] = two ? __ldg(row1 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
          // `zero` is a kernel argument that is always 0: making the first depth value depend
          // on every piece of the second row keeps the assembler from re-using the first row's
          // registers for the second row's loads (it otherwise issues them only after the first
          // chain has consumed its operands: two exposed latencies per trip instead of one).
          uint32_t tie = 0;
#pragma unroll
          for (int k = 0; k < NV; ++k) tie |= __float_as_uint(q[k].x);
          const float d0t = __uint_as_float(__float_as_uint(d0) | (tie & zero));
          i += 2;  // the next trip's ranks travel with this trip's rows
          const bool more = i < en;
          if (more) {
            two = i + 1 < en;
            r0 = __ldg(ranks_depth + i);
            f0 = __ldg(ranks_feat + i);
            r1 = two ? __ldg(ranks_depth + i + 1) : r0;
            f1 = two ? __ldg(ranks_feat + i + 1) : f0;
          }
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            acc[4 * k + 0] = fmaf(a[k].x, d0t, acc[4 * k + 0]);
            acc[4 * k + 1] = fmaf(a[k].y, d0t, acc[4 * k + 1]);
            acc[4 * k + 2] = fmaf(a[k].z, d0t, acc[4 * k + 2]);
            acc[4 * k + 3] = fmaf(a[k].w, d0t, acc[4 * k + 3]);
          }
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            acc[4 * k + 0] = fmaf(q[k].x, d1, acc[4 * k + 0]);
            acc[4 * k + 1] = fmaf(q[k].y, d1, acc[4 * k + 1]);
            acc[4 * k + 2] = fmaf(q[k].z, d1, acc[4 * k + 2]);
            acc[4 * k + 3] = fmaf(q[k].w, d1, acc[4 * k + 3]);
          }
          if (!more) break;
        }
      }
    }
    float* o = out + (int64_t)b * C * V + v0 + lane;
#pragma unroll
    for (int c = 0; c < C; ++c) st_stream(o + (int64_t)c * V, acc[c]);
  }
}

template <int KCH, bool FULLC>
static int launch_fwd_impl(const float* depth, const float* feat, const int32_t* rd,
                      const int32_t* rf, const int32_t* rb, const int32_t* tile_start,
                      const int32_t* heavy, int64_t heavy_ints, int B, int C, int64_t V,
                      bool feat_rows_fit_32bit, float* out, cudaStream_t stream) {
  constexpr int CC = 32 * KCH;
  const size_t smem = sizeof(float) * kFwdWarps * (CC * kRowPitch + kRingSlots * kSlotInts);
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd<KCH, FULLC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_pool_fwd<KCH, FULLC>,
                                                                kFwdWarps * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (env_flag("VEON_FWD_CTAS", 0) > 0 && env_flag("VEON_FWD_CTAS", 0) < ctas_per_sm) ctas_per_sm = env_flag("VEON_FWD_CTAS", 0);
  }
  const int64_t tps = ceil_div64(V, kTileVoxels), n_tiles = (int64_t)B * tps;
  const int n_chunks = (C + CC - 1) / CC;
  if (n_tiles * n_chunks > 0x7fffffffLL || (int64_t)B * V > 0x7fffffffLL) return VEON_E_RANGE;
  // the gather addresses rows with 32-bit float offsets
  if (!feat_rows_fit_32bit) return VEON_E_RANGE;
  const int vec_ok = ((V & 3) == 0) && (((uintptr_t)out & 15) == 0);
  int64_t blocks = ceil_div64(n_tiles * n_chunks, kFwdWarps);
  const int64_t resident = (int64_t)ctas_per_sm * sm_count();  // persistent grid
  if (blocks > resident) blocks = resident;
  // the heavy-tile kernel stages feature rows with 16-byte copies
  if (heavy && ((C & 3) != 0 || ((uintptr_t)feat & 15) != 0)) heavy = nullptr;
  int capturing = 0;   // inside a stream capture the heavy grid joins the main grid explicitly
  {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone)
      capturing = 1;
  }
  // tuning knobs, read once per process
  static const int dbg = env_flag("VEON_FWD_DBG", 0);
  static const int knob_group = env_flag("VEON_FWD_GROUP", 0);
  static const int knob_heavy_ctas = env_flag("VEON_FWD_HEAVY_CTAS", 8);
  static const int knob_order = env_flag("VEON_FWD_ORDER", 1);
  const bool group_ok = FULLC && vec_ok && (V % kTileVoxels) == 0 && (tps % kGroupTiles) == 0 &&
                        (C & 3) == 0 && ((uintptr_t)feat & 15) == 0;
  if (group_ok && knob_group) {
    const size_t gsmem = sizeof(float) * (2 * kGroupRound * CC + CC * kStagePitch +
                                          4 * kGroupRound + 8 * kGroupVoxels + 48);
    static int group_ctas_per_sm = 0;
    if (group_ctas_per_sm == 0) {
      VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_group<KCH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
      VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &group_ctas_per_sm, k_pool_fwd_group<KCH>, kGroupThreads, gsmem));
      if (group_ctas_per_sm < 1) group_ctas_per_sm = 1;
    }
    const int64_t n_gitems = n_tiles / kGroupTiles * n_chunks;
    int64_t gblocks = (int64_t)group_ctas_per_sm * sm_count();
    if (gblocks > n_gitems) gblocks = n_gitems;
    k_pool_fwd_group<KCH><<<(unsigned)gblocks, kGroupThreads, gsmem, stream>>>(
        depth, feat, rd, rf, rb, tile_start, heavy, (uint32_t)n_gitems, (uint32_t)n_chunks,
        (uint32_t)(tps / kGroupTiles), V, C, out);
    VEON_LAUNCH_CHECK();
    if (heavy) {
      const size_t hsmem2 = sizeof(float) * (kHeavyChunk * CC + CC * kRowPitch + 2 * kHeavyChunk + 128);
      static int hc = 0;
      if (hc == 0) {
        VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_heavy<KCH>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem2));
        VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&hc, k_pool_fwd_heavy<KCH>,
                                                                    kHeavyThreads, hsmem2));
        if (hc < 1) hc = 1;
      }
      const int heavy_cap = (int)(heavy_ints - 2);
      int64_t hblocks = (int64_t)heavy_cap * n_chunks;
      if (hblocks > (int64_t)hc * sm_count()) hblocks = (int64_t)hc * sm_count();
      if (hblocks > 0) {
        VEON_CUDA_TRY(launch_pdl(k_pool_fwd_heavy<KCH>, dim3((unsigned)hblocks), dim3(kHeavyThreads),
                                 hsmem2, stream, depth, feat, rd, rf, rb, tile_start, heavy,
                                 heavy_cap, (uint32_t)tps, V, C, (uint32_t)n_chunks, vec_ok, out,
                                 capturing, 0));
        VEON_LAUNCH_CHECK();
      }
    }
    return 0;
  }
  if (heavy) {
    const size_t hsmem = sizeof(float) * (kHeavyChunk * CC + CC * kRowPitch + 2 * kHeavyChunk + 128);
    static int heavy_ctas_per_sm = 0;
    if (heavy_ctas_per_sm == 0) {
      VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_heavy<KCH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
      VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &heavy_ctas_per_sm, k_pool_fwd_heavy<KCH>, kHeavyThreads, hsmem));
      if (heavy_ctas_per_sm < 1) heavy_ctas_per_sm = 1;
    }
    const int heavy_cap = (int)(heavy_ints - 2);
    const int per_sm = min(heavy_ctas_per_sm, max(1, knob_heavy_ctas));
    int64_t hblocks = (int64_t)heavy_cap * n_chunks;
    if (hblocks > (int64_t)per_sm * sm_count()) hblocks = (int64_t)per_sm * sm_count();
    if (hblocks > 0) {
      // The two grids write disjoint tiles and neither waits for the other: the second is a
      // programmatic dependent of the first (both trigger at their first instruction), so
      // its CTAs are scheduled as soon as SM resources free up instead of after the drain.
      // Heavy-tile CTAs last: they fill the SMs as the persistent CTAs of the main grid
      // finish one by one (VEON_FWD_ORDER=0 runs them first instead).
      if (knob_order == 1) {
        k_pool_fwd<KCH, FULLC><<<(unsigned)blocks, kFwdWarps * 32, smem, stream>>>(
            depth, feat, rd, rf, rb, tile_start, heavy, (uint32_t)(n_tiles * n_chunks),
            (uint32_t)tps, V, C, (uint32_t)n_chunks, vec_ok, out, dbg);
        VEON_LAUNCH_CHECK();
        VEON_CUDA_TRY(launch_pdl(k_pool_fwd_heavy<KCH>, dim3((unsigned)hblocks), dim3(kHeavyThreads),
                                 hsmem, stream, depth, feat, rd, rf, rb, tile_start, heavy,
                                 heavy_cap, (uint32_t)tps, V, C, (uint32_t)n_chunks, vec_ok, out,
                                 capturing, 0));
        VEON_LAUNCH_CHECK();
        return 0;
      }
      k_pool_fwd_heavy<KCH><<<(unsigned)hblocks, kHeavyThreads, hsmem, stream>>>(
          depth, feat, rd, rf, rb, tile_start, heavy, heavy_cap, (uint32_t)tps, V, C,
          (uint32_t)n_chunks, vec_ok, out, 0, 0);
      VEON_LAUNCH_CHECK();
      VEON_CUDA_TRY(launch_pdl(k_pool_fwd<KCH, FULLC>, dim3((unsigned)blocks), dim3(kFwdWarps * 32),
                               smem, stream, depth, feat, rd, rf, rb, tile_start, heavy,
                               (uint32_t)(n_tiles * n_chunks), (uint32_t)tps, V, C,
                               (uint32_t)n_chunks, vec_ok, out, dbg));
      VEON_LAUNCH_CHECK();
      return 0;
    }
    heavy = nullptr;
  }
  k_pool_fwd<KCH, FULLC><<<(unsigned)blocks, kFwdWarps * 32, smem, stream>>>(
      depth, feat, rd, rf, rb, tile_start, heavy, (uint32_t)(n_tiles * n_chunks), (uint32_t)tps,
      V, C, (uint32_t)n_chunks, vec_ok, out, dbg);
  VEON_LAUNCH_CHECK();
  return 0;
}

// main grid = k_pool_fwd_narrow, heavy tiles = k_pool_fwd_heavy<1> queued behind it (as in
// launch_fwd_impl)
template <int NV>
static int launch_fwd_narrow(const float* depth, const float* feat, const int32_t* rd,
                             const int32_t* rf, const int32_t* rb, const int32_t* tile_start,
                             const int32_t* heavy, int64_t heavy_ints, int B, int64_t V,
                             float* out, cudaStream_t stream) {
  constexpr int C = 4 * NV;
  const int64_t tps = V / kTileVoxels, n_tiles = (int64_t)B * tps;
  static int ctas_per_sm = 0;
  if (ctas_per_sm == 0) {
    VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_pool_fwd_narrow<NV>,
                                                                kNarrowWarps * 32, 0));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  int64_t blocks = ceil_div64(n_tiles, kNarrowWarps);
  if (blocks > (int64_t)ctas_per_sm * sm_count()) blocks = (int64_t)ctas_per_sm * sm_count();
  static const int heavy_min = env_flag("VEON_NARROW_HEAVY_MIN", kNarrowHeavyMin);  // tuning knob
  int capturing = 0;
  {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone)
      capturing = 1;
  }
  static const int knob_order = env_flag("VEON_NARROW_ORDER", 0);        // 0: heavy grid first (344 -> 295 us at C3 density)
  static const int knob_heavy_ctas = env_flag("VEON_NARROW_HEAVY_CTAS", 8);
  int64_t hblocks = 0;
  size_t hsmem = 0;
  int heavy_cap = 0;
  if (heavy) {
    hsmem = sizeof(float) * (kHeavyChunk * 32 + 32 * kRowPitch + 2 * kHeavyChunk + 128);
    static int heavy_ctas_per_sm = 0;
    if (heavy_ctas_per_sm == 0) {
      VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_fwd_heavy<1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
      VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &heavy_ctas_per_sm, k_pool_fwd_heavy<1>, kHeavyThreads, hsmem));
      if (heavy_ctas_per_sm < 1) heavy_ctas_per_sm = 1;
    }
    heavy_cap = (int)(heavy_ints - 2);
    const int per_sm = min(heavy_ctas_per_sm, max(1, knob_heavy_ctas));
    hblocks = heavy_cap;
    if (hblocks > (int64_t)per_sm * sm_count()) hblocks = (int64_t)per_sm * sm_count();
  }
  if (hblocks > 0 && knob_order == 0 && !capturing) {
    // heavy tiles first (a normal launch), the main grid moves in beside it; neither waits
    k_pool_fwd_heavy<1><<<(unsigned)hblocks, kHeavyThreads, hsmem, stream>>>(
        depth, feat, rd, rf, rb, tile_start, heavy, heavy_cap, (uint32_t)tps, V, C, 1u, 1, out, 0,
        heavy_min);
    VEON_LAUNCH_CHECK();
    VEON_CUDA_TRY(launch_pdl(k_pool_fwd_narrow<NV>, dim3((unsigned)blocks), dim3(kNarrowWarps * 32),
                             0, stream, depth, feat, rd, rf, rb, tile_start, heavy,
                             (uint32_t)n_tiles, (uint32_t)tps, V, out, 0u, heavy_min));
    VEON_LAUNCH_CHECK();
    return 0;
  }
  k_pool_fwd_narrow<NV><<<(unsigned)blocks, kNarrowWarps * 32, 0, stream>>>(
      depth, feat, rd, rf, rb, tile_start, heavy, (uint32_t)n_tiles, (uint32_t)tps, V, out, 0u,
      heavy_min);
  VEON_LAUNCH_CHECK();
  if (hblocks > 0) {
    VEON_CUDA_TRY(launch_pdl(k_pool_fwd_heavy<1>, dim3((unsigned)hblocks), dim3(kHeavyThreads),
                             hsmem, stream, depth, feat, rd, rf, rb, tile_start, heavy,
                             heavy_cap, (uint32_t)tps, V, C, 1u, 1, out, capturing, heavy_min));
    VEON_LAUNCH_CHECK();
  }
  return 0;
}

template <int KCH>
static int launch_fwd(const float* depth, const float* feat, const int32_t* rd,
                      const int32_t* rf, const int32_t* rb, const int32_t* tile_start,
                      const int32_t* heavy, int64_t heavy_ints, int B, int C, int64_t V,
                      bool feat_rows_fit_32bit, float* out, cudaStream_t stream) {
  if (C % (32 * KCH) == 0)
    return launch_fwd_impl<KCH, true>(depth, feat, rd, rf, rb, tile_start, heavy, heavy_ints, B, C,
                                      V, feat_rows_fit_32bit, out, stream);
  return launch_fwd_impl<KCH, false>(depth, feat, rd, rf, rb, tile_start, heavy, heavy_ints, B, C,
                                     V, feat_rows_fit_32bit, out, stream);
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
                                           const int32_t* tile_start,
                                           const int32_t* tile_heavy, int64_t tile_heavy_ints,
                                           int B, int C, int64_t V, int64_t n_feat_rows,
                                           float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !tile_start || !out ||
      B <= 0 || C <= 0 || V <= 0 || n_feat_rows <= 0 || (tile_heavy && tile_heavy_ints < 2))
    return VEON_E_BADARG;
  // the gather addresses feature rows with 32-bit BYTE offsets; the ring header packs
  // sample<<16 | chunk
  const bool fit32 = n_feat_rows * (int64_t)C <= 0x3fffffffLL && B < 65536 &&
                     (int64_t)C <= 65535LL * 32;
  int kch = fwd_kch_override();
  {  // narrow rows: lane-per-voxel kernel (VEON_FWD_NARROW=0 keeps the lane-per-channel one)
    static const int knob_narrow = env_flag("VEON_FWD_NARROW", 1);
    // without a heavy list a single lane would walk arbitrarily long voxels; rank arithmetic
    // is exact (no float32 merging of voxels) only while B*V <= 2^24
    const bool ok = knob_narrow && kch == 0 && C <= 32 && (C & 3) == 0 && tile_heavy && fit32 &&
                    V % kTileVoxels == 0 && (int64_t)B * V <= (1 << 24) &&
                    (((uintptr_t)feat | (uintptr_t)out) & 15) == 0;
    if (ok) {
#define VEON_NARROW_CASE(NV_) case NV_: return launch_fwd_narrow<NV_>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, tile_heavy, tile_heavy_ints, B, V, out, stream);
      switch (C / 4) {
        VEON_NARROW_CASE(1) VEON_NARROW_CASE(2) VEON_NARROW_CASE(3) VEON_NARROW_CASE(4)
        VEON_NARROW_CASE(5) VEON_NARROW_CASE(6) VEON_NARROW_CASE(7) VEON_NARROW_CASE(8)
        default: break;
      }
#undef VEON_NARROW_CASE
    }
  }
  if (kch == 0) kch = (C <= 32) ? 1 : 2;
  switch (kch) {
    case 1: return launch_fwd<1>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, tile_heavy, tile_heavy_ints, B, C, V, fit32, out, stream);
    case 2: return launch_fwd<2>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, tile_heavy, tile_heavy_ints, B, C, V, fit32, out, stream);
    case 4: return launch_fwd<4>(depth, feat, ranks_depth, ranks_feat, ranks_bev, tile_start, tile_heavy, tile_heavy_ints, B, C, V, fit32, out, stream);
    default: return VEON_E_BADARG;
  }
}

extern "C" int veon_bev_pool_v2_ds_fwd(const float* depth, const float* feat,
                                       const int32_t* ranks_depth, const int32_t* ranks_feat,
                                       const int32_t* ranks_bev, const int32_t* voxel_start,
                                       int B, int C, int Z, int Y, int X, int64_t n_feat_rows,
                                       float* out, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !voxel_start || !out ||
      B <= 0 || C <= 0 || Z <= 0 || Y <= 0 || X <= 0 || n_feat_rows <= 0)
    return VEON_E_BADARG;
  constexpr int KCH = 2, CC = 64;
  // even grid, whole 64-channel chunks, 16-byte feature rows, exact float32 voxel ranks
  if ((Z | Y | X) & 1 || C % CC != 0 || ((uintptr_t)feat & 15) != 0) return VEON_E_UNSUPPORTED;
  const int64_t V = (int64_t)Z * Y * X;
  if ((int64_t)B * V >= (1 << 24) || n_feat_rows * (int64_t)C > 0x3fffffffLL) return VEON_E_RANGE;
  // an item covers half an x-row when that keeps the pieces even: smaller tiles, more CTAs/SM
  int xsplit = 1;   // (halving the rows was measured slower: per-item overheads dominate)
  {
    const char* e = getenv("VEON_DS_XSPLIT");
    if (e && atoi(e) >= 1 && X % (2 * atoi(e)) == 0) xsplit = atoi(e);
  }
  const int Xs = X / xsplit;
  const int Xp = (Xs + 3) / 4 * 4 + 4, Xhp = Xs / 2 + 1;
  const size_t smem = sizeof(float) * ((size_t)kDsRound * CC + (size_t)CC * Xp + (size_t)CC * Xhp +
                                       2 * kDsRound + 4 * (size_t)Xp + 8);
  if (smem > 220 * 1024) return VEON_E_UNSUPPORTED;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    VEON_CUDA_TRY(cudaFuncSetAttribute(k_pool_ds_fwd<KCH>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  int per_sm = 1;
  VEON_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pool_ds_fwd<KCH>,
                                                              kDsThreads, smem));
  if (per_sm < 1) per_sm = 1;
  const int n_chunks = C / CC;
  const int64_t n_items = (int64_t)B * (Z / 2) * (Y / 2) * xsplit * n_chunks;
  if (n_items > 0x7fffffffLL) return VEON_E_RANGE;
  int64_t blocks = (int64_t)per_sm * sm_count();
  if (blocks > n_items) blocks = n_items;
  k_pool_ds_fwd<KCH><<<(unsigned)blocks, kDsThreads, smem, stream>>>(
      depth, feat, ranks_depth, ranks_feat, ranks_bev, voxel_start, (uint32_t)n_items,
      (uint32_t)n_chunks, Z, Y, X, xsplit, C, out);
  VEON_LAUNCH_CHECK();
  return 0;
}
