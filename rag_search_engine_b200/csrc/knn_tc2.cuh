// knn_tc2.cuh — K4 filter pass, second mapping: QUERIES RESIDENT IN TENSOR MEMORY.
//
// knn_tc_kernel (knn_tc.cuh) streams the 393 KB query block from L2 next to the rows and is bound by
// the TMA fill rate (r01 ncu: 14.7 GB of TMA loads per pass for 7.4 GB of rows).  Here the queries
// never move after the prologue: a CTA PAIR (cta_group::2) holds the 256 queries as the UMMA A
// operand in TMEM (128 per CTA: lane = query, 384 columns = k), the corpus streams once through
// shared memory as the B operand (each CTA loads half of every 128-row tile), and
// D[256 queries x 128 rows] lands in the remaining 128 TMEM columns of both CTAs.  TMA traffic per
// pass = the rows, nothing else.  In the epilogue one thread owns one query (its threshold lives in
// a register) and scans the tile's 128 rows.
//
//   cluster (2 CTAs) x 74 = 148 SMs, persistent over 128-row tiles
//   warp 0  : TMA producer for this CTA's 64 rows of each tile (12-stage ring, 2 k-blocks = 16 KB/stage)
//   warp 1  : leader CTA — single-thread tcgen05.mma.cta_group::2 issuer;
//             peer CTA   — forwards "my stage is full" to the leader's barrier (remote mbarrier arrive)
//   warps 2-5: prologue — tcgen05.st the CTA's 128 queries into TMEM; then epilogue
//   tcgen05.commit ... multicast::cluster frees the stage / publishes D in BOTH CTAs.
// Same survivor output as knn_tc_kernel<1>: cand_pairs[q*cap + slot] = {local row, dot~/|a|}.
#pragma once
#include "knn_tc.cuh"

namespace rse {

constexpr int kT2TileRows = 128;                          // UMMA N
constexpr int kT2HalfRows = 64;                           // rows per CTA per tile
constexpr int kT2KbPerStage = 2;
constexpr int kT2StageBytes = kT2KbPerStage * kT2HalfRows * kTcBK * 4;   // 16,384
constexpr int kT2Stages = 12;                             // 192 KB: two full tiles in flight
constexpr int kT2Threads = 192;
constexpr int kT2ACols = kScanD;                          // 384 TMEM columns hold the queries
constexpr int kT2DCol = kScanD;                           // D starts at column 384 (128 columns)
constexpr int kT2SmemBytes = kT2Stages * kT2StageBytes + 2 * kT2TileRows * 4 + 64 * 8 + 16 + 1024;
// kind::tf32, D=f32, A (TMEM) / B K-major, M=256 (cta pair), N=128
constexpr uint32_t kT2Idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kT2TileRows >> 3) << 17) | ((256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t target_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_bar)), "r"(target_rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tc2_commit_mc(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tc2_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// tmap_b: [n_rows][384] f32, box {32, 64}, SWIZZLE_128B.  q: [256][384] f32 (zero padded).
// thr: per-query threshold on dot~/|a| (+inf for padding).  grid = 2 * n_clusters, cluster = 2.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2Threads, 1)
knn_tc2_filter_kernel(const __grid_constant__ CUtensorMap tmap_b, const float* __restrict__ q,
                      const float* __restrict__ amag, int64_t n_rows, int64_t n_tiles,
                      const float* __restrict__ thr, uint2* __restrict__ cand_pairs,
                      unsigned int* __restrict__ cand_count, int cap) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - smem_u32(smem_raw));
  float* s_inv = reinterpret_cast<float*>(smem + kT2Stages * kT2StageBytes);        // [2][128] 1/|a| per tile row
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * kT2TileRows);
  uint64_t* full = bars;                          // [S] local TMA landed
  uint64_t* empty = bars + kT2Stages;             // [S] both CTAs' MMA reads retired (multicast commit)
  uint64_t* pfull = bars + 2 * kT2Stages;         // [S] leader only: the peer's stage is full
  uint64_t* tfull = bars + 3 * kT2Stages;         // [1] D complete (multicast commit)
  uint64_t* tempty = bars + 3 * kT2Stages + 1;    // [1] leader only: both CTAs' epilogues drained D
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kT2Stages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int64_t cluster_id = blockIdx.x >> 1;
  const int64_t n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kT2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&pfull[s], 1); }
    mbar_init(tfull, 1);
    mbar_init(tempty, 8);                         // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- prologue: this CTA's 128 queries → TMEM columns [0, 384), lane = query
  if (warp >= 2) {
    const int quarter = warp & 3;
    const int qrow = static_cast<int>(rank) * 128 + quarter * 32 + lane;
    const float4* qp = reinterpret_cast<const float4*>(q + static_cast<int64_t>(qrow) * kScanD);
    for (int c0 = 0; c0 < kScanD; c0 += 32) {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 f = __ldg(qp + c0 / 4 + j);
        v[4 * j + 0] = __float_as_uint(f.x); v[4 * j + 1] = __float_as_uint(f.y);
        v[4 * j + 2] = __float_as_uint(f.z); v[4 * j + 3] = __float_as_uint(f.w);
      }
      tc_st32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                              // barriers initialised + queries resident in both CTAs
  tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer: my 64 rows of every tile =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
        const int row0 = static_cast<int>(t * kT2TileRows + rank * kT2HalfRows);
        for (int kb = 0; kb < kTcKBlocks; kb += kT2KbPerStage) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], kT2StageBytes);
          unsigned char* sb = smem + stage * kT2StageBytes;
#pragma unroll
          for (int j = 0; j < kT2KbPerStage; ++j)
            tma_load_2d(sb + j * (kT2HalfRows * kTcBK * 4), &tmap_b, (kb + j) * kTcBK, row0, &full[stage]);
          if (++stage == kT2Stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();                                  // reconverge before the aligned cluster barrier below
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (leader) {
        // ===================== MMA issuer (leader CTA) =====================
        uint32_t it = 0;
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
          mbar_wait(tempty, (it & 1u) ^ 1u);       // both epilogues drained D
          tc_fence_after();
          for (int kb = 0; kb < kTcKBlocks; kb += kT2KbPerStage) {
            mbar_wait(&full[stage], phase);
            mbar_wait(&pfull[stage], phase);
            tc_fence_after();
            const uint32_t b_addr = base + stage * kT2StageBytes;
#pragma unroll
            for (int j = 0; j < kT2KbPerStage; ++j) {
              const uint64_t bdesc = tc_smem_desc(b_addr + j * (kT2HalfRows * kTcBK * 4));
#pragma unroll
              for (int k = 0; k < kTcBK / 8; ++k) {
                const int kk = (kb + j) * (kTcBK / 8) + k;           // global K-step (8 tf32 = 8 TMEM columns)
                tc2_mma_tf32_ts(tmem_base + kT2DCol, tmem_base + static_cast<uint32_t>(kk * 8),
                                bdesc + static_cast<uint64_t>(2 * k), kT2Idesc, kk != 0 ? 1u : 0u);
              }
            }
            tc2_commit_mc(&empty[stage]);          // both CTAs may refill this stage
            if (++stage == kT2Stages) { stage = 0; phase ^= 1u; }
          }
          tc2_commit_mc(tfull);                    // D complete in both CTAs
        }
      } else {
        // ===================== peer CTA: tell the leader when my stage has landed =====================
        for (int64_t t = cluster_id; t < n_tiles; t += n_clusters) {
          for (int kb = 0; kb < kTcKBlocks; kb += kT2KbPerStage) {
            mbar_wait(&full[stage], phase);
            mbar_arrive_remote(&pfull[stage], 0u);
            if (++stage == kT2Stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5): one thread = one query =====================
    const int quarter = warp & 3;
    const int et = quarter * 32 + lane;                               // 0..127: also the tile row this thread prepares
    const int qi = static_cast<int>(rank) * 128 + et;
    const float my_thr = thr[qi];
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(kT2DCol);
    uint32_t it = 0;
    for (int64_t t = cluster_id; t < n_tiles; t += n_clusters, ++it) {
      // 1/|a| of the tile's rows (0 for empty slots / past the end → product 0, handled by `ok` below)
      float* inv = s_inv + (it & 1u) * kT2TileRows;
      const int64_t myrow = t * kT2TileRows + et;
      const float am = (myrow < n_rows) ? __ldg(amag + myrow) : -1.0f;
      inv[et] = am > 0.0f ? rsqrtf(am) : -1.0f;
      asm volatile("bar.sync 1, 128;" ::: "memory");                  // the 128 epilogue threads only
      mbar_wait(tfull, it & 1u);
      tc_fence_after();
      for (int c0 = 0; c0 < kT2TileRows; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr + static_cast<uint32_t>(c0), v);
        const float4* inv4 = reinterpret_cast<const float4*>(inv + c0);
        uint32_t mask = 0u;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 is = inv4[j4];
          mask |= (is.x > 0.0f && __uint_as_float(v[4 * j4 + 0]) * is.x >= my_thr) ? (1u << (4 * j4 + 0)) : 0u;
          mask |= (is.y > 0.0f && __uint_as_float(v[4 * j4 + 1]) * is.y >= my_thr) ? (1u << (4 * j4 + 1)) : 0u;
          mask |= (is.z > 0.0f && __uint_as_float(v[4 * j4 + 2]) * is.z >= my_thr) ? (1u << (4 * j4 + 2)) : 0u;
          mask |= (is.w > 0.0f && __uint_as_float(v[4 * j4 + 3]) * is.w >= my_thr) ? (1u << (4 * j4 + 3)) : 0u;
        }
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1u;
          const unsigned int slot = atomicAdd(&cand_count[qi], 1u);
          if (slot < static_cast<unsigned int>(cap)) {
            float sv = 0.0f;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) if (jj == j) sv = __uint_as_float(v[jj]) * inv[c0 + jj];
            cand_pairs[static_cast<int64_t>(qi) * cap + slot] =
                make_uint2(static_cast<uint32_t>(t * kT2TileRows + c0 + j), __float_as_uint(sv));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(tempty); else mbar_arrive_remote(tempty, 0u);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace rse
