// knn_tc.cuh — K4: tensor-core path for LARGE query batches, with exact re-score.
//
// Same reference computation as K1 (vec0 cosine KNN, semantic_search.py:254-261); this path
// exists because the exact lane-per-row scan is FP32-pipe-bound beyond ~16 queries per pass.
// It NEVER decides a result on approximate numbers:
//
//   1. probe   : TF32 tcgen05 GEMM over a strided sample of 128-row tiles, epilogue stores the
//                approximate distances d~ of the sample; the existing radix select gives each
//                query the K'-th smallest sample d~ (tau_s).
//   2. filter  : the same GEMM over ALL rows; epilogue keeps row r for query q iff
//                  cos~(q, r) >= 1 - tau_s - 2*eps,
//                where eps bounds |d~ - d_exact| (TF32 operand truncation, eps = 2.5e-3 > 2^-9;
//                see DESIGN.md §5).  Every row of the exact top-K' passes:
//                  d_K' <= (K'-th exact d in the sample) <= tau_s + eps, and d~ <= d + eps.
//   3. refine  : per query, s_K = the K'-th largest approximate value among the survivors (exact,
//                since all rows above the filter bound survived); only rows within 2*eps of s_K
//                stay (K' plus a thin band instead of K'·C/S rows).
//   4. re-score: those rows are re-computed with the EXACT sequential fp32 arithmetic of K1 (same
//                code: mac<FMA>, cosine_tail), keyed with the vec0 emit-order key, sorted, and the
//                first K' emitted → the same packed candidates K1+K2 emit.
//   A query whose survivor / refined list overflows is flagged and re-run through K1/K2 by the host.
//
// GEMM mapping: D[128 rows x 256 queries] per tile, K = 384 in 12 k-blocks of 32 floats (one
// 128-byte swizzle atom).  The 393 KB query block cannot stay in shared memory, so it streams from
// L2 with the rows; to halve that re-streaming a CTA works on a PAIR of row tiles per query
// k-block (r01 ncu: with one tile per k-block the kernel moved 22.1 GB through TMA per pass, 3x the
// rows, and sat at the TMA fill rate, ~8.3 TB/s).  Warp-specialised: warp 0 = TMA producer
// (cp.async.bulk.tensor, 3-stage mbarrier ring, 64 KB/stage: 2 x 16 KB of rows + 32 KB of queries),
// warp 1 = single-thread tcgen05.mma.kind::tf32 issuer (UMMA 128x256x8, the two accumulators fill
// the 512 TMEM columns), warps 2-5 = epilogue (tcgen05.ld 32x32b, one TMEM lane = one row per
// thread; the producer keeps prefetching the next pair during the epilogue).
// Persistent, one CTA per SM.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "knn_scan.cuh"

namespace rse {

constexpr int kTcBM = 128;
constexpr int kTcBN = 256;
constexpr int kTcBK = 32;
constexpr int kTcStages = 3;
constexpr int kTcKBlocks = kScanD / kTcBK;                 // 12
constexpr int kTcStageA = kTcBM * kTcBK * 4;               // 16,384 (one 128-row tile, one k-block)
constexpr int kTcStageB = kTcBN * kTcBK * 4;               // 32,768 (256 queries, one k-block)
constexpr int kTcStageBytes = 2 * kTcStageA + kTcStageB;   // 65,536: TWO row tiles share one query k-block
constexpr int kTcThreads = 192;                            // 6 warps
constexpr int kTcSmemBytes = kTcStages * kTcStageBytes + 2 * kTcBN * 4 + 16 * 8 + 16 + 1024;
constexpr float kTcEps = 2.5e-3f;                          // bound on |d~ - d| (see header)
constexpr int kTcCandCap = 8192;                           // survivors kept per query

// UMMA instruction descriptor, kind::tf32, D=f32, A/B K-major, M=128, N=256
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kTcBN >> 3) << 17) | ((kTcBM >> 4) << 24);

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row groups of 1024 B)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                              // LBO (unused for swizzled K-major) = 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                      // SBO = 1024 B
  d |= static_cast<uint64_t>(1) << 46;                              // descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                              // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- per-query constants
// inv_sb[q] = 1/||q|| (f32); used by the probe epilogue to turn cos~ into d~.
__global__ void tc_query_consts_kernel(const double* __restrict__ sb, int nq, float* __restrict__ inv_sb) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= kTcBN) return;
  inv_sb[q] = (q < nq && sb[q] > 0.0) ? static_cast<float>(1.0 / sb[q]) : 0.0f;
}

// thr[q] = (1 - tau_s - 2 eps) * ||q||   (compared against dot~ / ||a||), +inf for padding columns
// or when the sample gave no bound.  tau_s = K'-th smallest sample d~ from the radix-select state:
// resolved bits of the prefix, unresolved low bits set (an upper bound of the bucket).
__global__ void tc_threshold_kernel(const SelState* __restrict__ st, const double* __restrict__ sb, int nq,
                                    unsigned int kprime, float* __restrict__ thr) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= kTcBN) return;
  float t = __int_as_float(0x7F800000);   // +inf → nothing passes
  if (q < nq) {
    const SelState s = st[q];
    const unsigned long long hi_mask = s.mask >> 32, hi_pref = s.prefix >> 32;
    if (s.mask == 0ull) {
      t = -__int_as_float(0x7F800000);    // fewer than K' sample rows: keep everything (host avoids this case)
    } else {
      const uint32_t okey = static_cast<uint32_t>(hi_pref | (~hi_mask & 0xFFFFFFFFull));
      const float tau = __uint_as_float(f32_from_orderable(okey));
      const float cut = 1.0f - tau - 2.0f * kTcEps - 1e-6f;
      t = cut * static_cast<float>(sb[q]);
      if (cut < 0.0f) t = cut * static_cast<float>(sb[q]) * 1.0001f - 1e-30f;
    }
  }
  thr[q] = t;
}

// ---------------------------------------------------------------- the GEMM
// MODE 0 (probe) : tiles t = 0..n_tiles-1 map to row tile t*tile_stride; d~ stored to
//                  dist[q*ld + t*128 + r] (empty vec0 slots get the invalid sentinel).
// MODE 1 (filter): all row tiles; survivors appended to cand_pairs[q*cap + slot] = {local row, dot~/|a|}.
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_q,
              const float* __restrict__ amag, int64_t n_rows, int64_t n_tiles, int64_t tile_stride, int nq,
              const float* __restrict__ thr, const float* __restrict__ inv_sb, uint32_t* __restrict__ dist,
              int64_t ld, uint2* __restrict__ cand_pairs, unsigned int* __restrict__ cand_count, int cap) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - smem_u32(smem_raw));
  float* s_thr = reinterpret_cast<float*>(smem + kTcStages * kTcStageBytes);
  float* s_isb = s_thr + kTcBN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_isb + kTcBN);
  uint64_t* full = bars;                       // [kTcStages]
  uint64_t* empty = bars + kTcStages;          // [kTcStages]
  uint64_t* tfull = bars + 2 * kTcStages;      // [2]
  uint64_t* tempty = bars + 2 * kTcStages + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < kTcBN; i += blockDim.x) {
    s_thr[i] = (MODE == 1) ? thr[i] : 0.0f;
    s_isb[i] = (MODE == 0) ? inv_sb[i] : 0.0f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_pairs = (n_tiles + 1) / 2;
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        const int64_t t0 = 2 * p, t1 = 2 * p + 1;
        const int row0 = static_cast<int>(t0 * tile_stride * kTcBM);
        // an odd tail has no second tile: point it past the end (TMA zero-fills out-of-bounds rows)
        const int row1 = (t1 < n_tiles) ? static_cast<int>(t1 * tile_stride * kTcBM) : static_cast<int>(n_rows);
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], kTcStageBytes);
          unsigned char* sa = smem + stage * kTcStageBytes;
          tma_load_2d(sa, &tmap_a, kb * kTcBK, row0, &full[stage]);
          tma_load_2d(sa + kTcStageA, &tmap_a, kb * kTcBK, row1, &full[stage]);
          tma_load_2d(sa + 2 * kTcStageA, &tmap_q, kb * kTcBK, 0, &full[stage]);
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;
      for (int64_t p = blockIdx.x; p < n_pairs; p += gridDim.x, ++it) {
        mbar_wait(&tempty[0], (it & 1u) ^ 1u);            // epilogue drained both accumulators
        tc_fence_after();
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = base + stage * kTcStageBytes;
          const uint64_t adesc0 = tc_smem_desc(a_addr);
          const uint64_t adesc1 = tc_smem_desc(a_addr + kTcStageA);
          const uint64_t bdesc = tc_smem_desc(a_addr + 2 * kTcStageA);
#pragma unroll
          for (int k = 0; k < kTcBK / 8; ++k) {
            // advance 8 tf32 = 32 B inside the swizzle atom: +2 in the (>>4) start-address field
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            tc_mma_tf32(tmem_base, adesc0 + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), kTcIdesc, accum);
            tc_mma_tf32(tmem_base + kTcBN, adesc1 + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), kTcIdesc, accum);
          }
          tc_commit(&empty[stage]);                       // smem slot free once these MMAs retire
          if (++stage == kTcStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tfull[0]);                             // both accumulators complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                         // TMEM lanes [32*quarter, +32) belong to this warp
    uint32_t it = 0;
    for (int64_t p = blockIdx.x; p < n_pairs; p += gridDim.x, ++it) {
      float amv[2];
#pragma unroll
      for (int buf = 0; buf < 2; ++buf) {
        const int64_t t = 2 * p + buf;
        const int64_t row = t * tile_stride * kTcBM + quarter * 32 + lane;
        amv[buf] = (t < n_tiles && row < n_rows) ? __ldg(amag + row) : -1.0f;
      }
      mbar_wait(&tfull[0], it & 1u);
      tc_fence_after();
#pragma unroll
      for (int buf = 0; buf < 2; ++buf) {
        const int64_t t = 2 * p + buf;
        if (t >= n_tiles) continue;
        const int64_t row = t * tile_stride * kTcBM + quarter * 32 + lane;
        const float am = amv[buf];
        const bool valid = am > 0.0f;
        const float inv_sa = valid ? rsqrtf(am) : 0.0f;
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(buf * kTcBN);
        const int ncol = (MODE == 0) ? ((nq + 31) & ~31) : kTcBN;
        for (int c0 = 0; c0 < ncol; c0 += 32) {
          uint32_t v[32];
          tc_ld32(taddr0 + static_cast<uint32_t>(c0), v);
          if (MODE == 0) {
            const int64_t orow = t * kTcBM + quarter * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int q = c0 + j;
              if (q < nq) {
                uint32_t bits = 0x7FFFFFFFu;              // invalid sentinel (empty slot / past the end)
                if (valid) bits = __float_as_uint(1.0f - __uint_as_float(v[j]) * inv_sa * s_isb[q]);
                dist[static_cast<int64_t>(q) * ld + orow] = bits;
              }
            }
          } else {
            // branch-free compare of 32 columns → bitmask; survivors are rare (≈K'·C/S of C rows)
            const float4* th4 = reinterpret_cast<const float4*>(s_thr + c0);
            uint32_t mask = 0u;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 th = th4[j4];
              mask |= (__uint_as_float(v[4 * j4 + 0]) * inv_sa >= th.x) ? (1u << (4 * j4 + 0)) : 0u;
              mask |= (__uint_as_float(v[4 * j4 + 1]) * inv_sa >= th.y) ? (1u << (4 * j4 + 1)) : 0u;
              mask |= (__uint_as_float(v[4 * j4 + 2]) * inv_sa >= th.z) ? (1u << (4 * j4 + 2)) : 0u;
              mask |= (__uint_as_float(v[4 * j4 + 3]) * inv_sa >= th.w) ? (1u << (4 * j4 + 3)) : 0u;
            }
            if (!valid) mask = 0u;
            while (mask) {
              const int j = __ffs(mask) - 1;
              const int q = c0 + j;
              mask &= mask - 1u;
              const unsigned int slot = atomicAdd(&cand_count[q], 1u);
              if (slot < static_cast<unsigned int>(cap)) {
                // (row, dot~/|a|): the second-level refinement ranks survivors by this value
                float sv = 0.0f;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) if (jj == j) sv = __uint_as_float(v[jj]) * inv_sa;
                cand_pairs[static_cast<int64_t>(q) * cap + slot] = make_uint2(static_cast<uint32_t>(row), __float_as_uint(sv));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[0]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------- refine + exact re-score + finish
// One CTA per query.
//   (a) second-level refinement on the approximate values the filter kept: with s_K = the K'-th
//       largest dot~/|a| among the survivors (= among ALL rows, because every row above the filter
//       threshold survived), only rows with s >= s_K - 2*eps*|q| can be in the exact top-K'
//       (same argument as the filter bound, now with the exact K'-th approximate value);
//   (b) those few rows (K' plus the 2*eps band) are re-computed with the EXACT sequential fp32
//       arithmetic of K1 (mac<FMA>, cosine_tail), lane-per-row straight from global memory;
//   (c) the exact emit-order keys are sorted and the first K' emitted as packed candidates.
// status[q] = 1 (host re-runs the query through K1/K2) when the survivor list or the refined
// list overflowed, or fewer than K' rows survived (a corpus/query the bound does not cover: fewer
// than K' valid rows, zero or non-finite query norm).  `normalized`: the survivor values are cos~
// (knn_tc3, unit-norm operands) instead of dot~/|a| (knn_tc / knn_tc2), so the band is not scaled by |q|.
// smem: cap * 8 bytes (pairs) — reused for the exact keys.
// 256 threads, 256 histogram bins (one per thread): the digit d with  sum(hist[0..d-1]) < need <= sum(hist[0..d]).
// The thread that owns d writes *out_digit = d and *out_before = sum(hist[0..d-1]); *out_digit stays 256 when
// the histogram holds fewer than `need` entries.  s_warp: 8 words of scratch.  Ends with a barrier.
__device__ __forceinline__ void block_pick_digit(const unsigned int* s_hist, unsigned int need, unsigned int* s_warp,
                                                 unsigned int* out_digit, unsigned int* out_before) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int c = s_hist[threadIdx.x];
  unsigned int incl = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  if (threadIdx.x == 0) *out_digit = 256u;
  __syncthreads();
  unsigned int base = 0u;
#pragma unroll
  for (int w = 0; w < 8; ++w) base += (w < warp) ? s_warp[w] : 0u;
  incl += base;
  if (incl >= need && incl - c < need) { *out_digit = threadIdx.x; *out_before = incl - c; }
  __syncthreads();
}

constexpr int kTcRefineCap = 2048;                         // rows re-scored exactly per query at most

template <bool FMA>
__global__ void __launch_bounds__(kSelThreads, 2)
knn_refine_kernel(const float* __restrict__ emb, const float* __restrict__ amag, const float* __restrict__ q,
                  const double* __restrict__ sb, const uint2* __restrict__ cand_pairs,
                  const unsigned int* __restrict__ cand_count, int cap, int kprime, uint64_t pos_base,
                  const int64_t* __restrict__ rowid, const int32_t* __restrict__ movie_idx,
                  long long* __restrict__ cand, int* __restrict__ status, int normalized) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint2* pairs = reinterpret_cast<uint2*>(smem_raw);                       // [cap]
  __shared__ float s_q[kScanD];
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned int s_prefix, s_need, s_m, s_digit, s_before;
  __shared__ unsigned int s_warp[8];
  __shared__ uint32_t s_rows[kTcRefineCap];
  const int qi = blockIdx.x;
  const unsigned int cnt = cand_count[qi];
  long long* out = cand + static_cast<int64_t>(qi) * kprime * 3;
  if (cnt > static_cast<unsigned int>(cap)) {                              // survivor overflow
    if (threadIdx.x == 0) status[qi] = 1;
    for (int i = threadIdx.x; i < kprime * 3; i += blockDim.x) out[i] = -1ll;
    return;
  }
  const int n = static_cast<int>(cnt);
  // (knn_tc3 never emits an empty slot or a zero row: its thresholds are positive and those rows read cos~ = 0;
  //  knn_tc / knn_tc2 check |a|^2 in their epilogues)
  for (int i = threadIdx.x; i < n; i += blockDim.x) pairs[i] = cand_pairs[static_cast<int64_t>(qi) * cap + i];
  for (int i = threadIdx.x; i < kScanD; i += blockDim.x) s_q[i] = q[static_cast<int64_t>(qi) * kScanD + i];
  if (threadIdx.x == 0) { s_prefix = 0u; s_need = static_cast<unsigned int>(kprime < n ? kprime : n); s_m = 0u; }
  __syncthreads();

  // ---- (a) K'-th largest approximate value: MSD radix select (4 x 8 bits) on ~orderable(s)
  //      (descending s == ascending ~orderable)
  uint32_t resolved_mask = 0u;
  if (n > kprime) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0u;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t key = ~f32_orderable(pairs[i].y);
        if ((key & resolved_mask) == prefix) atomicAdd(&s_hist[(key >> shift) & 0xFFu], 1u);
      }
      __syncthreads();
      block_pick_digit(s_hist, s_need, s_warp, &s_digit, &s_before);     // a thread-0 loop over the bins cost 10 % of the kernel
      if (threadIdx.x == 0) {
        s_prefix = prefix | ((s_digit & 0xFFu) << shift);
        s_need = s_need - s_before;
      }
      resolved_mask |= 0xFFu << shift;
      __syncthreads();
    }
  }
  // s_K (exact K'-th largest approximate value), cut = s_K - 2 eps |q|  (slack for f32 rounding)
  float cut = -__int_as_float(0x7F800000);
  if (n > kprime) {
    const float s_k = __uint_as_float(f32_from_orderable(~s_prefix));
    const float qn = normalized ? 1.0f : static_cast<float>(sb[qi]);
    cut = s_k - (2.0f * kTcEps + 1e-6f) * qn - 1e-30f;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint2 pr = pairs[i];
    if (__uint_as_float(pr.y) >= cut && pr.y != 0xFF800000u) {
      const unsigned int slot = atomicAdd(&s_m, 1u);
      if (slot < kTcRefineCap) s_rows[slot] = pr.x;
    }
  }
  __syncthreads();
  const unsigned int m = s_m;
  if (m > kTcRefineCap || m < static_cast<unsigned int>(kprime)) {         // refined list overflow (mass ties) / too few rows
    if (threadIdx.x == 0) status[qi] = 1;
    for (int i = threadIdx.x; i < kprime * 3; i += blockDim.x) out[i] = -1ll;
    return;
  }
  if (threadIdx.x == 0) status[qi] = 0;

  // ---- (b) exact distances, lane-per-row from global memory (the reference's sequential sum)
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);   // pairs are dead
  __syncthreads();
  int m2 = 2;
  while (m2 < static_cast<int>(m)) m2 <<= 1;
  const double sbq = sb[qi];
  for (int i = threadIdx.x; i < m2; i += blockDim.x) {
    unsigned long long key = ~0ull;
    if (i < static_cast<int>(m)) {
      const uint32_t row = s_rows[i];
      const float4* rp = reinterpret_cast<const float4*>(emb + static_cast<int64_t>(row) * kScanD);
      float acc = 0.0f;
#pragma unroll 8
      for (int c = 0; c < kScanD / 4; ++c) {
        const float4 a = __ldg(rp + c);
        acc = mac<FMA>(acc, a.x, s_q[4 * c + 0]);
        acc = mac<FMA>(acc, a.y, s_q[4 * c + 1]);
        acc = mac<FMA>(acc, a.z, s_q[4 * c + 2]);
        acc = mac<FMA>(acc, a.w, s_q[4 * c + 3]);
      }
      const float d = cosine_tail(acc, sqrt(static_cast<double>(__ldg(amag + row))), sbq);
      key = knn_key(f32_orderable(__float_as_uint(d)), pos_base + static_cast<uint64_t>(row));
    }
    keys[i] = key;
  }
  // ---- (c) keys-only bitonic sort, emit the first K'
  for (int size = 2; size <= m2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (m2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kprime; i += blockDim.x) {
    long long* c = out + static_cast<int64_t>(i) * 3;
    if (i < static_cast<int>(m)) {
      const uint64_t key = keys[i];
      const uint64_t pos = knn_key_pos(key);
      const int64_t local = static_cast<int64_t>(pos - pos_base);
      c[0] = static_cast<long long>(key);
      c[1] = rowid ? rowid[local] : static_cast<long long>(pos);
      c[2] = movie_idx ? movie_idx[local] : -1;
    } else {
      c[0] = -1ll; c[1] = -1ll; c[2] = -1ll;
    }
  }
}

}  // namespace rse
