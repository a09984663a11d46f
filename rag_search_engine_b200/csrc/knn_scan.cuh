// knn_scan.cuh — K1: exact vec0-cosine scan of the chunk-embedding matrix.
//
// Replaces the per-row `distance_cosine_float` loop that sqlite-vec runs inside
// `embedding MATCH :q AND k = :k` (reference call site
// rag_search_engine/utils/semantic_search.py:254-261).
//
// The kernel does not approximate: every row's distance is computed with the
// reference's own association — ONE sequential fp32 accumulator over the 384
// products (separate multiply and add, or fused when the index was loaded in
// aarch64 mode), double sqrt/mul/div/sub tail, narrowed to f32 — so the f32
// distance of every row is bit-identical to the oracle and no shortlist/guard
// band is needed for the streaming path.
//
// B200 mapping (HBM-bound, 1536 B/row):
//   * one warp = one 32-row tile, LANE-PER-ROW (the only mapping that keeps the
//     reference's summation order); rows are staged in shared memory by the TMA
//     engine as 32 bulk copies (cp.async.bulk, one 1536 B row per lane) landing
//     on a row stride of 1552 B (≡16 mod 128), which makes the lane-per-row
//     128-bit LDS conflict-free;
//   * 4 warps per CTA, one 49.7 KB stage each (199 KB of the 227 KB), one CTA
//     per SM, persistent over tiles; a warp re-arms its stage as soon as its
//     last LDS retired, so the other three stages are in flight while it does
//     the 384-step dependent add chain — ≥100 KB in flight per SM, well above
//     the ~35 KB latency×bandwidth product per SM;
//   * up to QB=16 queries share one pass over the corpus (query vectors in
//     shared memory, read as broadcast LDS.128), which is what the hybrid batch
//     path uses.  For QB >= 2 the arithmetic is PACKED: two queries per
//     instruction with Blackwell's f32x2 FMA pipe (FFMA2).  ptxas contracts
//     mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under --fmad=false, which
//     would change the reference's unfused rounding, so the unfused step is
//     written as p = fma2(a, b, -0.0); acc = fma2(p, 1.0, acc) with -0.0 / 1.0
//     passed as kernel arguments (unknown to ptxas, held in uniform registers):
//     each fma2 then rounds exactly once on an exact product / sum, i.e.
//     rn(a*b) and rn(p+acc) — bit-identical to FMUL + FADD at half the
//     FMA-pipe cycles (r01 ncu: the scalar QB=8 kernel sat at 88 % of the
//     3-register FMA-pipe rate).
#pragma once
#include "common.cuh"

namespace rse {

constexpr int kScanD = 384;
constexpr int kScanRowStride = 388;                     // floats: 1552 B ≡ 16 (mod 128)
constexpr int kScanTileRows = 32;
constexpr int kScanWarps = 4;
constexpr int kScanStageBytes = kScanTileRows * kScanRowStride * 4;   // 49,664
constexpr int kScanMaxQB = 16;

__host__ __device__ constexpr int scan_smem_bytes(int qb) {
  return kScanWarps * kScanStageBytes + qb * kScanD * 4 + kScanWarps * 8 + 16;
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

template <bool FMA>
__device__ __forceinline__ float mac(float acc, float a, float b) {
  if (FMA) return __fmaf_rn(a, b, acc);
  return __fadd_rn(acc, __fmul_rn(a, b));   // never contracted
}

// ---- packed (two queries per instruction) arithmetic
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long dup2(float a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a));
  return r;
}
template <bool FMA>
__device__ __forceinline__ unsigned long long mac2(unsigned long long acc, unsigned long long a2, unsigned long long b2,
                                                   unsigned long long negz2, unsigned long long one2) {
  if (FMA) return fma2(a2, b2, acc);
  return fma2(fma2(a2, b2, negz2), one2, acc);   // rn(rn(a*b) + acc): the unfused reference step, never contracted
}

// vec0 tail: 1 - dot / (sqrt((double)aMag) * sqrt((double)bMag)), narrowed to f32.
__device__ __forceinline__ float cosine_tail(float dot, double sa, double sb) {
  double den = __dmul_rn(sa, sb);
  double d = __dsub_rn(1.0, __ddiv_rn(static_cast<double>(dot), den));
  return __double2float_rn(d);
}

// ---------------------------------------------------------------- query prep
// sb[j] = sqrt((double) Σ q_i^2) with the sequential fp32 sum (the reference
// recomputes bMag for every row; it is row-independent).
template <bool FMA>
__global__ void knn_query_prep_kernel(const float* __restrict__ q, int nq, int dim, double* __restrict__ sb) {
  // one warp per query: the row is fetched coalesced into shared memory, then ONE lane adds the squares in
  // index order (the reference's single sequential accumulator).  A thread-per-query loop over global memory
  // took 29 us for 256 queries (384 dependent scalar loads each).
  constexpr int kMaxStaged = 1024;
  __shared__ float s_v[4][kMaxStaged];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + w;
  if (j >= nq) return;
  const float* v = q + static_cast<int64_t>(j) * dim;
  float m = 0.0f;
  if (dim <= kMaxStaged) {
    for (int i = lane; i < dim; i += 32) s_v[w][i] = v[i];
    __syncwarp();
    if (lane == 0) {
#pragma unroll 8
      for (int i = 0; i < dim; ++i) m = mac<FMA>(m, s_v[w][i], s_v[w][i]);
    }
  } else if (lane == 0) {
    for (int i = 0; i < dim; ++i) m = mac<FMA>(m, v[i], v[i]);
  }
  if (lane == 0) sb[j] = sqrt(static_cast<double>(m));
}

// amag[row] = Σ a_i^2 (sequential fp32), -1 for empty vec0 slots.  Load time only.
template <bool FMA>
__global__ void knn_row_sqmag_kernel(const float* __restrict__ emb, int64_t n_rows, int dim,
                                     const uint8_t* __restrict__ valid, float* __restrict__ amag) {
  int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  if (valid && !valid[r]) { amag[r] = -1.0f; return; }
  const float* v = emb + r * dim;
  float m = 0.0f;
  if ((dim & 3) == 0) {
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (int i = 0; i < dim / 4; ++i) {
      float4 a = v4[i];
      m = mac<FMA>(m, a.x, a.x); m = mac<FMA>(m, a.y, a.y);
      m = mac<FMA>(m, a.z, a.z); m = mac<FMA>(m, a.w, a.w);
    }
  } else {
    for (int i = 0; i < dim; ++i) m = mac<FMA>(m, v[i], v[i]);
  }
  amag[r] = m;
}

// ---------------------------------------------------------------- K1 fast path (dim = 384)
template <int QB, bool FMA>
__global__ void __launch_bounds__(kScanWarps * 32, 1)
knn_scan384_kernel(const float* __restrict__ emb, const float* __restrict__ amag, int64_t n_rows,
                   const float* __restrict__ q, const double* __restrict__ sb, int nq,
                   float* __restrict__ dist, int64_t ld, unsigned long long negz2, unsigned long long one2) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* stage = reinterpret_cast<float*>(smem + warp * kScanStageBytes);
  float* qs = reinterpret_cast<float*>(smem + kScanWarps * kScanStageBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kScanWarps * kScanStageBytes + QB * kScanD * 4) + warp;

  // queries -> smem (zero-fill unused slots so the unrolled QB loop stays finite).
  // QB == 1: plain [384].  QB >= 2: pair-interleaved, qs[(jp*384 + i)*2 + {0,1}] = q[2jp + {0,1}][i],
  // so one LDS.128 yields {q0[i], q1[i], q0[i+1], q1[i+1]} = two packed operands.
  for (int i = threadIdx.x; i < QB * kScanD; i += blockDim.x) {
    int j, e;
    if (QB == 1) { j = 0; e = i; }
    else { const int jp = i / (2 * kScanD); const int r = i - jp * 2 * kScanD; e = r >> 1; j = 2 * jp + (r & 1); }
    qs[i] = (j < nq) ? q[j * kScanD + e] : 0.0f;
  }
  if (lane == 0) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int64_t n_tiles = (n_rows + kScanTileRows - 1) / kScanTileRows;
  const int64_t tstride = static_cast<int64_t>(gridDim.x) * kScanWarps;
  int64_t tile = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
  uint32_t parity = 0;

  auto issue = [&](int64_t t) {
    const int64_t row0 = t * kScanTileRows;
    const int64_t rem = n_rows - row0;
    const int nvalid = rem < kScanTileRows ? static_cast<int>(rem) : kScanTileRows;
    if (lane == 0) mbar_arrive_expect_tx(bar, static_cast<uint32_t>(nvalid) * (kScanD * 4));
    __syncwarp();
    if (lane < nvalid)
      bulk_g2s(stage + lane * kScanRowStride, emb + (row0 + lane) * kScanD, kScanD * 4, bar);
  };

  if (tile < n_tiles) issue(tile);

  double sbr[QB];
#pragma unroll
  for (int j = 0; j < QB; ++j) sbr[j] = (j < nq) ? sb[j] : 1.0;

  const float4* rowp = reinterpret_cast<const float4*>(stage + lane * kScanRowStride);
  const float4* qp = reinterpret_cast<const float4*>(qs);

  for (; tile < n_tiles; tile += tstride) {
    const int64_t row = tile * kScanTileRows + lane;
    const bool in_range = row < n_rows;
    const float am = in_range ? __ldg(amag + row) : -1.0f;   // issued before the wait

    mbar_wait(bar, parity);
    parity ^= 1u;

    float acc[QB];
    if constexpr (QB == 1) {
      acc[0] = 0.0f;
#pragma unroll 4
      for (int c = 0; c < kScanD / 4; ++c) {
        const float4 a = rowp[c];
        const float4 b = qp[c];
        acc[0] = mac<FMA>(acc[0], a.x, b.x);
        acc[0] = mac<FMA>(acc[0], a.y, b.y);
        acc[0] = mac<FMA>(acc[0], a.z, b.z);
        acc[0] = mac<FMA>(acc[0], a.w, b.w);
      }
    } else {
      constexpr int NP = QB / 2;
      unsigned long long acc2[NP];
#pragma unroll
      for (int jp = 0; jp < NP; ++jp) acc2[jp] = 0ull;                 // {+0.0f, +0.0f}
      const ulonglong2* qp2 = reinterpret_cast<const ulonglong2*>(qs);  // [NP][192]
#pragma unroll 2
      for (int c = 0; c < kScanD / 4; ++c) {
        const float4 a = rowp[c];
        const unsigned long long ax = dup2(a.x), ay = dup2(a.y), az = dup2(a.z), aw = dup2(a.w);
#pragma unroll
        for (int jp = 0; jp < NP; ++jp) {
          const ulonglong2 b01 = qp2[jp * (kScanD / 2) + 2 * c];
          const ulonglong2 b23 = qp2[jp * (kScanD / 2) + 2 * c + 1];
          acc2[jp] = mac2<FMA>(acc2[jp], ax, b01.x, negz2, one2);
          acc2[jp] = mac2<FMA>(acc2[jp], ay, b01.y, negz2, one2);
          acc2[jp] = mac2<FMA>(acc2[jp], az, b23.x, negz2, one2);
          acc2[jp] = mac2<FMA>(acc2[jp], aw, b23.y, negz2, one2);
        }
      }
#pragma unroll
      for (int jp = 0; jp < NP; ++jp) {
        acc[2 * jp] = __uint_as_float(static_cast<unsigned int>(acc2[jp]));
        acc[2 * jp + 1] = __uint_as_float(static_cast<unsigned int>(acc2[jp] >> 32));
      }
    }
    __syncwarp();   // every lane's LDS of this stage has retired (values consumed above)
    if (tile + tstride < n_tiles) issue(tile + tstride);

    if (in_range) {
      if (am >= 0.0f) {
        const double sa = sqrt(static_cast<double>(am));
#pragma unroll
        for (int j = 0; j < QB; ++j)
          if (j < nq) dist[j * ld + row] = cosine_tail(acc[j], sa, sbr[j]);
      } else {
#pragma unroll
        for (int j = 0; j < QB; ++j)
          if (j < nq) dist[j * ld + row] = __uint_as_float(0x7FFFFFFFu);   // empty slot → okey 0xFFFFFFFF
      }
    }
  }
}

// Any other embedding width (the reference's own unit test uses dim = 4, tests/test_semantic_search.py:34-44; a
// different sentence-transformers model is 768 wide; the CLIP text embeddings of llm/multimodal.py are 512 wide).
// Same arithmetic as the 384 kernel — one SEQUENTIAL fp32 accumulator per (row, query), unfused unless FMA — so the
// rows are staged through shared memory: a warp loads 32 rows x 64 floats with coalesced 128-byte reads into a
// [32][65] tile (stride 65: lane-per-row reads are conflict-free), then every lane walks its own row's 64 values
// in order for up to 16 queries; the accumulators stay in registers across the chunks.  (r01's version read
// global memory lane-per-row — 32 different 2 KB-strided rows per load instruction — and reached ~3 % of the HBM
// bandwidth; it was "not a performance path" until the image search of §8 f4 started using it.)
constexpr int kGenChunk = 64;
constexpr int kGenWarps = 4;
template <bool FMA>
__global__ void __launch_bounds__(kGenWarps * 32)
knn_scan_generic_kernel(const float* __restrict__ emb, const float* __restrict__ amag,
                        int64_t n_rows, int dim, const float* __restrict__ q,
                        const double* __restrict__ sb, int nq, float* __restrict__ dist,
                        int64_t ld) {
  __shared__ float s_tile[kGenWarps][32][kGenChunk + 1];
  __shared__ float s_q[kScanMaxQB][kGenChunk];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * kGenWarps + warp) * 32;
  const int64_t row = row0 + lane;
  float acc[kScanMaxQB];
#pragma unroll
  for (int j = 0; j < kScanMaxQB; ++j) acc[j] = 0.0f;
  for (int c0 = 0; c0 < dim; c0 += kGenChunk) {
    const int nc = min(kGenChunk, dim - c0);
    __syncthreads();                                           // the previous chunk's queries / tile are consumed
    for (int i = threadIdx.x; i < nq * kGenChunk; i += blockDim.x) {
      const int j = i / kGenChunk, k = i % kGenChunk;
      s_q[j][k] = k < nc ? __ldg(q + static_cast<int64_t>(j) * dim + c0 + k) : 0.0f;
    }
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const int64_t gr = row0 + r;
      const float* src = emb + gr * dim + c0;
      float v0 = 0.0f, v1 = 0.0f;
      if (gr < n_rows) {
        if (lane < nc) v0 = __ldg(src + lane);
        if (lane + 32 < nc) v1 = __ldg(src + lane + 32);
      }
      s_tile[warp][r][lane] = v0;
      s_tile[warp][r][lane + 32] = v1;
    }
    __syncthreads();
    const float* mine = s_tile[warp][lane];
    for (int k = 0; k < nc; ++k) {
      const float a = mine[k];
#pragma unroll
      for (int j = 0; j < kScanMaxQB; ++j)
        if (j < nq) acc[j] = mac<FMA>(acc[j], a, s_q[j][k]);
    }
  }
  if (row >= n_rows) return;
  const float am = amag[row];
#pragma unroll
  for (int j = 0; j < kScanMaxQB; ++j) {
    if (j < nq) {
      float v = __uint_as_float(0x7FFFFFFFu);
      if (!(am < 0.0f)) v = cosine_tail(acc[j], sqrt(static_cast<double>(am)), sb[j]);
      dist[j * ld + row] = v;
    }
  }
}

// The same scan for any width that is a multiple of 4 floats (rows 16-byte aligned), r02: rows are STAGED by the warp
// that owns them with cp.async — 32 rows x 64 floats per stage, each 128-byte line one coalesced request, XOR-swizzled
// so that lane l reads row l without bank conflicts — three stages deep across chunks AND tiles, no block-wide barrier
// after the queries are in shared memory.  (The chunk loop above loads with 4-byte requests and synchronises the block
// twice per chunk: 1.0 TB/s on the 600 k x 512 image-search matrix, §8 f4.)  One sequential fp32 accumulator per
// (row, query), products unfused unless FMA: the arithmetic of the 384 kernel.
constexpr int kWideChunk = 64;                                  // floats per row per stage (256 B)
constexpr int kWideStageBytes = 32 * kWideChunk * 4;            // 8 KB
constexpr int kWideStages = 3;
constexpr int kWideWarps = 4;
__host__ __device__ constexpr int wide_smem_bytes(int qb, int dim) {
  return kWideWarps * kWideStages * kWideStageBytes + qb * dim * 4;
}
__device__ __forceinline__ void wide_cp_async16(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
template <int QB, bool FMA>
__global__ void __launch_bounds__(kWideWarps * 32)
knn_scan_wide_kernel(const float* __restrict__ emb, const float* __restrict__ amag, int64_t n_rows, int dim,
                     const float* __restrict__ q, const double* __restrict__ sb, int nq, float* __restrict__ dist,
                     int64_t ld) {
  extern __shared__ __align__(128) unsigned char wsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wbuf = wsmem + warp * kWideStages * kWideStageBytes;
  const uint32_t wbuf_s = smem_u32(wbuf);
  float* s_q = reinterpret_cast<float*>(wsmem + kWideWarps * kWideStages * kWideStageBytes);   // [QB][dim]
  for (int i = threadIdx.x; i < QB * dim; i += blockDim.x) {
    const int j = i / dim;
    s_q[i] = j < nq ? q[i] : 0.0f;
  }
  __syncthreads();
  const int n_chunks = (dim + kWideChunk - 1) / kWideChunk;
  const int64_t n_tiles = (n_rows + 31) / 32;
  const int64_t w_global = static_cast<int64_t>(blockIdx.x) * kWideWarps + warp;
  const int64_t w_stride = static_cast<int64_t>(gridDim.x) * kWideWarps;
  const int64_t my_tiles = w_global < n_tiles ? (n_tiles - w_global + w_stride - 1) / w_stride : 0;
  const int64_t n_stages = my_tiles * n_chunks;
  // stage s = (tile s / n_chunks of this warp, chunk s % n_chunks); rows past the end re-read the last row
  auto issue = [&](int64_t s) {
    if (s < n_stages) {
      const int64_t tile = w_global + (s / n_chunks) * w_stride;
      const int c0 = static_cast<int>(s % n_chunks) * kWideChunk;
      const int nf4 = (min(kWideChunk, dim - c0)) >> 2;                  // float4s of this chunk: 1..16
      const uint32_t base = wbuf_s + static_cast<uint32_t>(s % kWideStages) * kWideStageBytes;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int r = 2 * it + (lane >> 4), f = lane & 15;
        int64_t gr = tile * 32 + r;
        if (gr >= n_rows) gr = n_rows - 1;
        if (f < nf4) wide_cp_async16(base + r * (kWideChunk * 4) + ((f ^ (r & 15)) << 4), emb + gr * dim + c0 + 4 * f);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < kWideStages - 1; ++s) issue(s);
  float acc[QB];
  for (int64_t s = 0; s < n_stages; ++s) {
    const int c = static_cast<int>(s % n_chunks);
    if (c == 0) {
#pragma unroll
      for (int j = 0; j < QB; ++j) acc[j] = 0.0f;
    }
    issue(s + kWideStages - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kWideStages - 1) : "memory");
    __syncwarp();
    const int c0 = c * kWideChunk;
    const int nf4 = (min(kWideChunk, dim - c0)) >> 2;
    const unsigned char* rb = wbuf + (s % kWideStages) * kWideStageBytes + lane * (kWideChunk * 4);
    for (int f = 0; f < nf4; ++f) {
      const float4 a = *reinterpret_cast<const float4*>(rb + ((f ^ (lane & 15)) << 4));
#pragma unroll
      for (int j = 0; j < QB; ++j) {
        const float4 qv = *reinterpret_cast<const float4*>(s_q + j * dim + c0 + 4 * f);
        acc[j] = mac<FMA>(acc[j], a.x, qv.x);
        acc[j] = mac<FMA>(acc[j], a.y, qv.y);
        acc[j] = mac<FMA>(acc[j], a.z, qv.z);
        acc[j] = mac<FMA>(acc[j], a.w, qv.w);
      }
    }
    __syncwarp();                                               // the stage is free for stage s + kWideStages
    if (c == n_chunks - 1) {
      const int64_t row = (w_global + (s / n_chunks) * w_stride) * 32 + lane;
      if (row < n_rows) {
        const float am = amag[row];
#pragma unroll
        for (int j = 0; j < QB; ++j) {
          if (j < nq) {
            float v = __uint_as_float(0x7FFFFFFFu);
            if (!(am < 0.0f)) v = cosine_tail(acc[j], sqrt(static_cast<double>(am)), sb[j]);
            dist[j * ld + row] = v;
          }
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace rse
