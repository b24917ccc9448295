// Open-vocabulary tail on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   logits[v, q] = sum_c feat_occ[b, c, v] * W[q, c]      san_in_veon_temporal.py:257-259
//   per-class max over prompts, first-index argmax         san_in_veon_entry_temporal.py:273-297
//   occupancy gate, free label, [B,X,Y,Z] uint8            veon_temporal.py:223-229,240
//
// GEMM view per CTA tile: M = 128 voxels, N = padded prompt count, K = C.
//   * A = feat tile.  In memory the voxel index is contiguous, i.e. A is
//     "MN-major"; kind::tf32 accepts that (with the SWIZZLE_128B_BASE32B layout,
//     the only one allowed for MN-major 32-bit operands), so the tile sits in
//     shared memory in the same orientation as in HBM: atoms of 4 channel rows x
//     32 voxels (512 B), 32-byte chunks XOR-swizzled by the row index; descriptor
//     LBO = 512 B between 32-voxel groups, SBO = 2 KB between 4-row groups.
//   * B = W, K-major SW128.
//   * D in TMEM (fp32), two accumulator buffers so the epilogue of tile i
//     overlaps the main loop of tile i+1.
//   * fp32 fidelity: 3xTF32.  a = a_hi + a_lo, w = w_hi + w_lo with *_hi exactly
//     TF32-representable; D = a_hi*[w_hi ; w_lo] (one MMA, N = 2*Npad, the
//     w_lo product lands in its own TMEM columns) + a_lo*w_hi.  Error ~2^-21
//     relative, which is what keeps the labels at >= 99.99 % agreement (plain
//     TF32 would flip ~0.3 % of near-tie voxels).
//   * The split is done IN REGISTERS on the way in (LDG.128 -> hi/lo -> two
//     STS.128): 16 B of shared-memory traffic per element instead of 20-24 B if
//     a TMA-landed tile had to be re-read to be split; at ~22 B/clk/SM of HBM the
//     shared-memory port is the scarce resource of this kernel.
// Warp roles (21 warps): 0-15 producers (the loads of the next stage are in flight while the
// current one is split and stored), 16 MMA issuer (+TMEM alloc), 17-20 epilogue.
#include <math.h>

#include "common.cuh"

namespace veon {
namespace tc {

constexpr int KC = 32;        // channels per pipeline stage = 4 UMMA K-steps of 8 (tf32)
constexpr int TM = 128;       // voxels per tile = UMMA M
constexpr int kProdWarps = 16;  // 2 channel rows of the 32-row stage each
constexpr int kMmaWarp = 16;
constexpr int kEpiWarp0 = 17;   // 17..20: warp % 4 = 1,2,3,0 -> the four TMEM lane quarters
constexpr int kWarps = 21;
constexpr int kMaxStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  do {  // try_wait suspends the thread in hardware for a bounded time, so this is not a hot spin
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// completion of all previously issued MMAs of this thread -> one arrival on `bar`
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// 32 lanes x 16 consecutive columns -> 16 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version 1<<46 | layout type [61,64)
constexpr uint64_t kSw128 = 2;         // SWIZZLE_128B         (B operand, K-major)
constexpr uint64_t kSw128Base32 = 1;   // SWIZZLE_128B_BASE32B (the only MN-major layout for tf32)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint64_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46) | (layout_type << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, A MN-major,
// B K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {  // exactly TF32-representable part
  return __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

struct Params {
  const float* feat;     // [B,C,V]
  const float* w;        // [Q,C]
  const int32_t* cls;    // [Q]
  const float* bin_occ;  // [B,2,V]
  uint8_t* labels;       // [B,X,Y,Z]
  float* logits;         // LOGITS variant: sem_occ [B,Q,V] instead of labels
  int B, C, Q, Z, Y, X, npad, stages, free_label;
  int64_t V;
  uint32_t tmem_cols;
};

// WP = float4 pieces of the W chunk a producer thread owns (1 for npad <= 64, else 2).
// With WP == 1 three named register sets fit (three stages of A in flight per lane).
// LOGITS: the epilogue stores the raw logits (semantic_inference_3d alone) instead of labels.
template <int WP, bool LOGITS>
__global__ void __launch_bounds__(kWarps * 32, 1) k_tail_tc(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int npad = p.npad, stages = p.stages;
  const uint32_t a_bytes = KC * TM * 4;          // 16 KB per A buffer (hi or lo)
  const uint32_t w_bytes = 2 * npad * KC * 4;    // [W_hi ; W_lo] chunk
  const uint32_t stage_bytes = 2 * a_bytes + w_bytes;
  uint8_t* stage_base = smem_raw;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_bytes);
  uint64_t* full = bars;                  // [stages]   producers -> MMA
  uint64_t* empty = bars + kMaxStages;    // [stages]   MMA -> producers
  uint64_t* acc_full = bars + 2 * kMaxStages;   // [2]  MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;           // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full + s, kProdWarps); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {  // TMEM allocation is warp-wide
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t vtiles = (p.V + TM - 1) / TM;
  const int64_t n_tiles = (int64_t)p.B * vtiles;
  const int n_chunks = p.C / KC;

  if (warp < kProdWarps) {
    // ============================ PRODUCERS ============================
    // Everything that does not change from stage to stage is set up once: the lane's
    // two shared-memory offsets, and the (up to two) W pieces this thread owns.
    const int ma = lane >> 3, j = lane & 7;
    const int row0 = 2 * warp;  // this warp's two channel rows inside a 32-row stage
    uint32_t a_off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      // channel row kl of the chunk = row kl%4 of 4-row group kl/4; the lane's 16 bytes are
      // half of 32-byte chunk j/2, which is XOR-swizzled by the row
      const int kl = row0 + i, kg4 = kl >> 2, kin = kl & 3;
      a_off[i] = (uint32_t)(kg4 * 4 + ma) * 512u + (uint32_t)kin * 128u +
                 (uint32_t)(((j >> 1) ^ kin) << 5) + (uint32_t)((j & 1) << 4);
    }
    constexpr int kWPieces = WP;  // npad * 8 float4 pieces over 512 threads (npad <= 64 * WP)
    const float4* w_src[kWPieces];
    uint32_t w_hi_off[kWPieces], w_lo_off[kWPieces];
    bool w_has[kWPieces], w_real[kWPieces];
#pragma unroll
    for (int q = 0; q < kWPieces; ++q) {
      const int idx = threadIdx.x + q * kProdWarps * 32;
      const int n = idx >> 3, j4 = idx & 7, n2 = n + npad;
      w_has[q] = idx < npad * 8;
      w_real[q] = w_has[q] && n < p.Q;
      w_src[q] = reinterpret_cast<const float4*>(p.w + (int64_t)(w_real[q] ? n : 0) * p.C) + j4;
      w_hi_off[q] = 2 * a_bytes + (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u +
                    (uint32_t)((j4 ^ (n & 7)) << 4);
      w_lo_off[q] = 2 * a_bytes + (uint32_t)(n2 >> 3) * 1024u + (uint32_t)(n2 & 7) * 128u +
                    (uint32_t)((j4 ^ (n2 & 7)) << 4);
    }
    const int64_t chunk_stride = (int64_t)KC * p.V;  // floats between consecutive chunks
    auto tile_src = [&](int64_t tile, bool& vin) -> const float* {
      const int64_t b = tile / vtiles;
      const int64_t v = (tile - b * vtiles) * TM + 4 * lane;
      vin = (tile < n_tiles) && (v < p.V);  // V % 4 == 0: the float4 is fully in or out
      return p.feat + ((int64_t)b * p.C + row0) * p.V + v;
    };
    auto load_rows = [&](const float* src, bool vin, float4* a) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        a[i] = vin ? ld_stream4(src + (int64_t)i * p.V) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    uint32_t s = 0, phase = 0;  // stage slot and the parity its `empty` barrier completes next
    bool first_round = true;
    // position of the stage whose data is being LOADED (runs one stage ahead of the stores)
    int64_t ltile = blockIdx.x;
    int lch = 0;
    bool vin;
    const float* src = tile_src(ltile, vin);
    auto load_stage = [&](float4* a, float4* wv) {
      load_rows(src, vin, a);
#pragma unroll
      for (int q = 0; q < kWPieces; ++q)
        wv[q] = w_real[q] ? __ldg(w_src[q] + lch * (KC / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      // advance the load position
      if (++lch == n_chunks) {
        lch = 0;
        ltile += gridDim.x;
        src = tile_src(ltile, vin);
      } else {
        src += chunk_stride;
      }
    };
    auto store_stage = [&](const float4* a, const float4* wv) {
      if (!first_round) mbar_wait(empty + s, phase);
      uint8_t* st = stage_base + (size_t)s * stage_bytes;
#pragma unroll
      for (int q = 0; q < kWPieces; ++q) {
        if (w_has[q]) {
          float4 hi, lo;
          hi.x = tf32_hi(wv[q].x); hi.y = tf32_hi(wv[q].y);
          hi.z = tf32_hi(wv[q].z); hi.w = tf32_hi(wv[q].w);
          lo.x = wv[q].x - hi.x; lo.y = wv[q].y - hi.y;
          lo.z = wv[q].z - hi.z; lo.w = wv[q].w - hi.w;
          *reinterpret_cast<float4*>(st + w_hi_off[q]) = hi;
          *reinterpret_cast<float4*>(st + w_lo_off[q]) = lo;
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 hi, lo;
        hi.x = tf32_hi(a[i].x); hi.y = tf32_hi(a[i].y);
        hi.z = tf32_hi(a[i].z); hi.w = tf32_hi(a[i].w);
        lo.x = a[i].x - hi.x; lo.y = a[i].y - hi.y;
        lo.z = a[i].z - hi.z; lo.w = a[i].w - hi.w;
        *reinterpret_cast<float4*>(st + a_off[i]) = hi;
        *reinterpret_cast<float4*>(st + a_bytes + a_off[i]) = lo;
      }
      fence_async_smem();  // generic-proxy writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
      if (++s == (uint32_t)stages) {
        s = 0;
        if (first_round) first_round = false; else phase ^= 1;
      }
    };
    // Two named register sets, loop unrolled by two: the loads of stage k+1 are issued
    // before stage k is written, and no register is ever copied while its load is in flight
    // (a rotating copy would wait for the load it copies).  When the thread owns a single W
    // piece (npad <= 64) a third set fits: three stages of A in flight per lane.
    const int64_t my_tiles = (n_tiles > (int64_t)blockIdx.x)
                                 ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t n_stages_total = my_tiles * n_chunks;
    if constexpr (WP == 1) {   // (a fourth set spills: 2 319 instead of 3 087 samples/s)
      float4 a0[2], a1[2], a2[2], w0[1], w1[1], w2[1];
      if (n_stages_total > 0) load_stage(a0, w0);
      if (n_stages_total > 1) load_stage(a1, w1);
      for (int64_t k = 0; k < n_stages_total; k += 3) {
        if (k + 2 < n_stages_total) load_stage(a2, w2);
        store_stage(a0, w0);
        if (k + 1 >= n_stages_total) break;
        if (k + 3 < n_stages_total) load_stage(a0, w0);
        store_stage(a1, w1);
        if (k + 2 >= n_stages_total) break;
        if (k + 4 < n_stages_total) load_stage(a1, w1);
        store_stage(a2, w2);
      }
    } else {
      float4 a0[2], a1[2], w0[kWPieces], w1[kWPieces];
      if (n_stages_total > 0) load_stage(a0, w0);
      for (int64_t k = 0; k < n_stages_total; k += 2) {
        if (k + 1 < n_stages_total) load_stage(a1, w1);
        store_stage(a0, w0);
        if (k + 1 >= n_stages_total) break;
        if (k + 2 < n_stages_total) load_stage(a0, w0);
        store_stage(a1, w1);
      }
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA ISSUER ============================
    if (lane == 0) {
      const uint32_t idesc_main = make_idesc(2 * npad), idesc_lo = make_idesc(npad);
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const int acc = tcount & 1;
        const uint32_t around = tcount >> 1;
        if (around > 0) mbar_wait(acc_empty + acc, (around - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * 2 * npad);
        for (int ch = 0; ch < n_chunks; ++ch, ++it) {
          const int s = it % stages;
          mbar_wait(full + s, (it / stages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * stage_bytes);
          const uint32_t sw = sa + 2 * a_bytes;
#pragma unroll
          for (int k = 0; k < KC / 8; ++k)  // a_hi * [w_hi ; w_lo]
            tc_mma_tf32(d, make_desc(sa + k * 4096, 512, 2048, kSw128Base32),
                        make_desc(sw + k * 32, 16, 1024, kSw128), idesc_main,
                        (ch > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < KC / 8; ++k)  // a_lo * w_hi, into the w_hi columns
            tc_mma_tf32(d, make_desc(sa + a_bytes + k * 4096, 512, 2048, kSw128Base32),
                        make_desc(sw + k * 32, 16, 1024, kSw128), idesc_lo, 1u);
          tc_commit(empty + s);  // smem stage reusable once these MMAs have read it
        }
        tc_commit(acc_full + acc);
      }
    }
  } else {
    // ============================ EPILOGUE ============================
    const int quarter = warp & 3;  // TMEM lanes a warp may touch: 32*(warp%4) ..
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int acc = tcount & 1;
      mbar_wait(acc_full + acc, (tcount >> 1) & 1);
      tc_fence_after();
      const int64_t b = tile / vtiles;
      const int64_t v = (tile - b * vtiles) * TM + 32 * quarter + lane;
      const uint32_t taddr = tmem_base + (uint32_t)(acc * 2 * npad) + ((uint32_t)(32 * quarter) << 16);
      if constexpr (LOGITS) {
        float* dst = p.logits + (int64_t)b * p.Q * p.V + v;  // lanes = consecutive voxels
        for (int q0 = 0; q0 < npad; q0 += 16) {
          float m[16], x[16];
          tc_ld16(taddr + q0, m);
          tc_ld16(taddr + npad + q0, x);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (q0 + i < p.Q && v < p.V) dst[(int64_t)(q0 + i) * p.V] = m[i] + x[i];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + acc);
        continue;
      }
      float best = 0.f, cur = 0.f;
      int best_cls = -1, cur_cls = -1;
      bool bad = false;
      for (int q0 = 0; q0 < npad; q0 += 16) {
        float m[16], x[16];
        tc_ld16(taddr + q0, m);          // a_hi*w_hi + a_lo*w_hi
        tc_ld16(taddr + npad + q0, x);   // a_hi*w_lo
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int q = q0 + i;
          if (q < p.Q) {
            const float logit = m[i] + x[i];
            const int cls = __ldg(p.cls + q);
            bad |= !(logit < INFINITY);
            if (cls != cur_cls) {
              if (cur_cls >= 0 && (best_cls < 0 || cur > best)) { best = cur; best_cls = cur_cls; }
              cur_cls = cls;
              cur = logit;
            } else {
              cur = fmaxf(cur, logit);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);  // accumulator drained
      if (cur_cls >= 0 && (best_cls < 0 || cur > best)) { best = cur; best_cls = cur_cls; }
      if (v < p.V) {
        bad |= (best == -INFINITY);
        const float b0 = __ldg(p.bin_occ + ((int64_t)b * 2 + 0) * p.V + v);
        const float b1 = __ldg(p.bin_occ + ((int64_t)b * 2 + 1) * p.V + v);
        const float mx = fmaxf(b0, b1);
        const float e0 = expf(b0 - mx), e1 = expf(b1 - mx);
        const bool occupied = (e0 / (e0 + e1)) > 0.5f;
        const int label = (occupied && !bad) ? best_cls : p.free_label;
        const int xx = (int)(v % p.X);
        const int yy = (int)((v / p.X) % p.Y);
        const int zz = (int)(v / ((int64_t)p.X * p.Y));
        p.labels[(((int64_t)b * p.X + xx) * p.Y + yy) * p.Z + zz] = (uint8_t)label;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols)
                 : "memory");
  }
}

}  // namespace tc
}  // namespace veon

using namespace veon;

// returns 0 when launched, VEON_E_UNSUPPORTED when the shape does not fit this path.
// logits != nullptr: write sem_occ [B,Q,V] (cls / bin_occ / labels unused).
template <int WP, bool LOGITS>
static int launch_variant(const tc::Params& p, unsigned grid, size_t smem, cudaStream_t stream) {
  static size_t attr_smem_dev[kMaxDevices] = {};   // per device: one process may drive several GPUs
  size_t& attr_smem = attr_smem_dev[current_device()];
  if (smem > attr_smem) {
    VEON_CUDA_TRY(cudaFuncSetAttribute(tc::k_tail_tc<WP, LOGITS>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  tc::k_tail_tc<WP, LOGITS><<<grid, tc::kWarps * 32, smem, stream>>>(p);
  VEON_LAUNCH_CHECK();
  return 0;
}

int veon_tail_tc_launch(const float* feat_occ, const float* text_w, const int32_t* cls,
                        const float* bin_occ, int B, int C, int Q, int Z, int Y, int X,
                        int free_label, uint8_t* labels, float* logits, cudaStream_t stream) {
  const int64_t V = (int64_t)Z * Y * X;
  const int npad = ((Q + 15) / 16) * 16;
  if (C % tc::KC != 0 || (V & 3) != 0 || npad > 128 || (((uintptr_t)feat_occ | (uintptr_t)text_w) & 15))
    return VEON_E_UNSUPPORTED;
  const size_t stage_bytes = 2 * (size_t)tc::KC * tc::TM * 4 + 2 * (size_t)npad * tc::KC * 4;
  const size_t tail = 1024;  // barriers + tmem slot
  int stages = (int)((227 * 1024 - tail) / stage_bytes);
  if (stages > tc::kMaxStages) stages = tc::kMaxStages;
  if (stages < 2) return VEON_E_UNSUPPORTED;
  const size_t smem = stages * stage_bytes + tail;
  const bool one_piece = npad <= 64;
  uint32_t cols = 32;
  while (cols < (uint32_t)(4 * npad)) cols <<= 1;  // 2 accumulators x 2*npad columns
  if (cols > 512) return VEON_E_UNSUPPORTED;
  tc::Params p;
  p.feat = feat_occ; p.w = text_w; p.cls = cls; p.bin_occ = bin_occ; p.labels = labels;
  p.logits = logits;
  p.B = B; p.C = C; p.Q = Q; p.Z = Z; p.Y = Y; p.X = X; p.npad = npad; p.stages = stages;
  p.free_label = free_label; p.V = V; p.tmem_cols = cols;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = (int64_t)B * ((V + tc::TM - 1) / tc::TM);
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  if (logits)
    return one_piece ? launch_variant<1, true>(p, grid, smem, stream)
                     : launch_variant<2, true>(p, grid, smem, stream);
  return one_piece ? launch_variant<1, false>(p, grid, smem, stream)
                   : launch_variant<2, false>(p, grid, smem, stream);
}
