// bm25.cuh — K6/K7: BM25 scoring over CSR postings and top-k with the reference's
// tie order.  Replaces the Python/SQLite loop of KeywordSearch.search
// (rag_search_engine/utils/keyword_search.py:196-250).
//
// Bit-exactness: every score is the reference's own IEEE-double expression,
//   denom = tf + k1*(1.0 - b + b*(dl/avgdl))            (:241)
//   add   = idf*(tf*(k1 + 1.0)/denom)                   (:242)
//   score = scores.get(doc, 0.0) + add                  (:244)
// evaluated with explicit round-to-nearest intrinsics (never contracted), and a
// document's adds happen in query-token order because the token passes are
// separated by a CTA barrier and a document appears at most once per posting
// list — so no atomics are needed and the sum order is the reference's.
// idf (:224) is computed on the host with the same libm `log` CPython uses.
//
// B200 mapping (HBM/L2 gather bound, 8 B/posting + 8 B gathered norm):
//   * documents are cut into ranges of kBmRange = 7168; one CTA owns one
//     (query, range) pair with the range's fp64 accumulators (56 KB) and
//     first-token bytes (7 KB) in shared memory — 3 CTAs/SM;
//   * a per-term range-offset table built at load time gives each CTA its slice
//     of every posting list without searching; slices are read as coalesced
//     8-byte (doc, tf) pairs;
//   * k1*(1 - b + b*dl/avgdl) depends only on the document, so it is cached per
//     (k1, b) as normk[M] (same expression, same bits) and gathered from L2;
//   * the CTA emits its range's top-k under the key
//     (score desc, first-token asc, doc asc) = sorted(..., reverse=True) on a
//     dict in insertion order (:250); a per-query k-way merge finishes.
#pragma once
#include "common.cuh"

namespace rse {

constexpr int kBmRange = 7168;   // 7 x 1024: 63 KB of accumulators + first-token bytes → 3 CTAs per SM
constexpr int kBmThreads = 256;
constexpr int kBmMaxTokens = 255;
constexpr int kBmDocsPerThread = kBmRange / kBmThreads;   // 32
constexpr unsigned int kBmCandCap = 192;                    // survivors of the threshold pass kept in smem

struct Key128 {
  unsigned long long hi;   // ~orderable(score): ascending hi == descending score
  unsigned long long lo;   // (first_token << 32) | doc
};
__device__ __forceinline__ bool key_less(const Key128& a, const Key128& b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
__device__ __forceinline__ Key128 key_inf() { return Key128{~0ull, ~0ull}; }

// normk[d] = k1 * (1.0 - b + b * (dl[d] / avgdl))   — keyword_search.py:241
__global__ void bm25_norm_kernel(const uint32_t* __restrict__ dl, int64_t n_docs, double k1, double b,
                                 double avgdl, double* __restrict__ normk) {
  int64_t d = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (d >= n_docs) return;
  const double ratio = __ddiv_rn(static_cast<double>(dl[d]), avgdl);
  const double inner = __dadd_rn(__dsub_rn(1.0, b), __dmul_rn(b, ratio));
  normk[d] = __dmul_rn(k1, inner);
}

// Range-offset table: roff[t*(NR+1) + r] = number of postings of term t with
// doc < r*kBmRange.  One thread per (term, boundary) binary-searches the posting
// list (load time only).
__global__ void bm25_range_offsets_by_term_kernel(const int64_t* __restrict__ indptr,
                                                  const uint2* __restrict__ post, int64_t n_terms, int nr,
                                                  uint32_t* __restrict__ roff) {
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = n_terms * (nr + 1);
  if (gid >= total) return;
  const int64_t t = gid / (nr + 1);
  const int r = static_cast<int>(gid - t * (nr + 1));
  const int64_t lo0 = indptr[t], hi0 = indptr[t + 1];
  const uint32_t bound = static_cast<uint32_t>(static_cast<int64_t>(r) * kBmRange > 0xFFFFFFFFll
                                                   ? 0xFFFFFFFFu
                                                   : static_cast<uint32_t>(r) * kBmRange);
  int64_t lo = lo0, hi = hi0;   // first posting with doc >= bound
  if (r == nr) {
    lo = hi0;
  } else {
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (post[mid].x < bound) lo = mid + 1; else hi = mid;
    }
  }
  roff[gid] = static_cast<uint32_t>(lo - lo0);
}

// Block-wide arg-min of per-thread Key128 values.  Returns the winning key to all
// threads and the winning thread id through *winner.  s_k / s_w are smem scratch
// of 8 entries (one per warp).
__device__ __forceinline__ Key128 block_argmin(Key128 mine, Key128* s_k, int* s_w, int* winner) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Key128 best = mine;
  int who = threadIdx.x;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    Key128 o;
    o.hi = __shfl_xor_sync(0xFFFFFFFFu, best.hi, off);
    o.lo = __shfl_xor_sync(0xFFFFFFFFu, best.lo, off);
    int ow = __shfl_xor_sync(0xFFFFFFFFu, who, off);
    if (key_less(o, best)) { best = o; who = ow; }
  }
  __syncthreads();   // protect scratch from the previous round's readers
  if (lane == 0) { s_k[warp] = best; s_w[warp] = who; }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  Key128 r = s_k[0];
  int rw = s_w[0];
  for (int w = 1; w < nw; ++w) {
    Key128 o = s_k[w];
    if (key_less(o, r)) { r = o; rw = s_w[w]; }
  }
  *winner = rw;
  return r;
}

// K6: one CTA per (range, query).
//   cand_hi/lo: [nq][nr][k]; cand_cnt: [nq][nr]
__global__ void __launch_bounds__(kBmThreads)
bm25_score_kernel(const int64_t* __restrict__ indptr, const uint2* __restrict__ post,
                  const uint32_t* __restrict__ roff, const double* __restrict__ normk, int nr,
                  int64_t n_docs, const int32_t* __restrict__ tok_indptr, const int32_t* __restrict__ term_rows,
                  const double* __restrict__ tok_idf, int q0, int k, double k1p1,
                  unsigned long long* __restrict__ cand_hi, unsigned long long* __restrict__ cand_lo,
                  int* __restrict__ cand_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* acc = reinterpret_cast<double*>(smem_raw);                       // [kBmRange]
  unsigned char* first = reinterpret_cast<unsigned char*>(acc + kBmRange); // [kBmRange]
  // token tables (scoring phase) and selection scratch (afterwards) share one buffer
  constexpr int kTokBytes = (kBmMaxTokens + 1) * (8 + 8 + 4);
  constexpr int kSelBytes = (256 + static_cast<int>(kBmCandCap)) * static_cast<int>(sizeof(Key128));
  __shared__ __align__(16) unsigned char s_cand_mem[kTokBytes > kSelBytes ? kTokBytes : kSelBytes];
  long long* s_lo = reinterpret_cast<long long*>(s_cand_mem);
  double* s_idf = reinterpret_cast<double*>(s_cand_mem + (kBmMaxTokens + 1) * 8);
  unsigned int* s_n = reinterpret_cast<unsigned int*>(s_cand_mem + (kBmMaxTokens + 1) * 16);
  __shared__ Key128 s_k[kBmThreads / 32];
  __shared__ int s_w[kBmThreads / 32];
  __shared__ unsigned int s_total;
  __shared__ unsigned int s_ncand;
  __shared__ Key128 s_tau;

  const int r = blockIdx.x;
  const int q = q0 + blockIdx.y;
  const int t0 = tok_indptr[q], t1 = tok_indptr[q + 1];
  const int ntok = t1 - t0;
  const int64_t cbase = (static_cast<int64_t>(q) * nr + r);

  if (threadIdx.x == 0) s_total = 0u;
  __syncthreads();
  for (int t = threadIdx.x; t < ntok; t += blockDim.x) {
    const int term = term_rows[t0 + t];
    long long lo = 0;
    unsigned int n = 0;
    if (term >= 0) {
      const uint32_t* ro = roff + static_cast<int64_t>(term) * (nr + 1) + r;
      const uint32_t a = ro[0], bnd = ro[1];
      lo = indptr[term] + a;
      n = bnd - a;
    }
    s_lo[t] = lo; s_n[t] = n; s_idf[t] = tok_idf[t0 + t];
    if (n) atomicAdd(&s_total, n);
  }
  __syncthreads();
  if (s_total == 0u) {
    if (threadIdx.x == 0) cand_cnt[cbase] = 0;
    return;
  }

  {
    uint4* a4 = reinterpret_cast<uint4*>(acc);                 // 64 KB of +0.0
    for (int i = threadIdx.x; i < kBmRange / 2; i += blockDim.x) a4[i] = make_uint4(0u, 0u, 0u, 0u);
    uint4* f4 = reinterpret_cast<uint4*>(first);               // 8 KB of 0xFF
    for (int i = threadIdx.x; i < kBmRange / 16; i += blockDim.x) f4[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  }
  __syncthreads();

  const uint32_t doc_base = static_cast<uint32_t>(r) * kBmRange;
  for (int t = 0; t < ntok; ++t) {
    const unsigned int n = s_n[t];
    if (n == 0u) continue;            // uniform across the CTA
    const uint2* p = post + s_lo[t];
    const double idf = s_idf[t];
    // 4 postings per thread per trip: all posting loads first, then all norm gathers, then the
    // arithmetic — two dependent global loads per posting otherwise leave the warp on the long
    // scoreboard (r01 ncu).  A document occurs once per list, so the 4 smem updates never alias.
    for (unsigned int i0 = threadIdx.x; i0 < n; i0 += 4 * kBmThreads) {
      uint2 e[4];
      double nk[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned int i = i0 + j * kBmThreads;
        ok[j] = i < n;
        e[j] = ok[j] ? __ldg(p + i) : make_uint2(doc_base, 1u);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) nk[j] = ok[j] ? __ldg(normk + e[j].x) : 1.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!ok[j]) continue;
        const double tfd = static_cast<double>(e[j].y);
        const double denom = __dadd_rn(tfd, nk[j]);
        const double add = __dmul_rn(idf, __ddiv_rn(__dmul_rn(tfd, k1p1), denom));
        const uint32_t l = e[j].x - doc_base;
        if (first[l] == 0xFF) first[l] = static_cast<unsigned char>(t);
        acc[l] = __dadd_rn(acc[l], add);          // first touch is 0.0 + add
      }
    }
    __syncthreads();
  }

  __syncthreads();   // the token tables are dead from here on: their storage becomes selection scratch

  // ---- K7 (range part): the range's top-k under (score desc, first token asc, doc asc).
  auto make_key = [&](int l) -> Key128 {
    Key128 kk;
    kk.hi = ~f64_orderable(static_cast<uint64_t>(__double_as_longlong(acc[l])));
    kk.lo = (static_cast<unsigned long long>(first[l]) << 32) | (doc_base + static_cast<uint32_t>(l));
    return kk;
  };
  if (k <= 32) {
    // Fast path.  (1) every thread finds the best of its 32 documents; (2) each warp sorts its 32
    // thread-bests with shuffles; (3) warp 0 merges the 8 sorted lists to the k-th best overall — a
    // valid threshold because the thread-bests are distinct documents; (4) one collect pass keeps the
    // documents at least that good (≥ k of them, usually barely more); (5) rank counting orders them.
    Key128 mine = key_inf();
#pragma unroll 4
    for (int i = 0; i < kBmDocsPerThread; ++i) {
      const int l = threadIdx.x + i * kBmThreads;
      if (first[l] == 0xFF) continue;
      const Key128 kk = make_key(l);
      if (key_less(kk, mine)) mine = kk;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // warp bitonic sort, ascending (best first)
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        Key128 o;
        o.hi = __shfl_xor_sync(0xFFFFFFFFu, mine.hi, stride);
        o.lo = __shfl_xor_sync(0xFFFFFFFFu, mine.lo, stride);
        const bool up = ((lane & size) == 0);
        const bool lower = ((lane & stride) == 0);
        const bool take_min = (up == lower);
        const bool o_less = key_less(o, mine);
        if (take_min ? o_less : !o_less) mine = o;
      }
    }
    Key128* s_wb = reinterpret_cast<Key128*>(s_cand_mem);        // [8][32]
    s_wb[warp * 32 + lane] = mine;
    __syncthreads();
    if (warp == 0) {
      // 8-way merge: lane w < 8 owns list w
      int head = 0;
      Key128 cur = (lane < kBmThreads / 32) ? s_wb[lane * 32] : key_inf();
      Key128 tau = key_inf();
      for (int round = 0; round < k; ++round) {
        Key128 best = cur;
        int who = lane;
#pragma unroll
        for (int off = 4; off > 0; off >>= 1) {
          Key128 o;
          o.hi = __shfl_xor_sync(0xFFFFFFFFu, best.hi, off);
          o.lo = __shfl_xor_sync(0xFFFFFFFFu, best.lo, off);
          const int ow = __shfl_xor_sync(0xFFFFFFFFu, who, off);
          if (key_less(o, best)) { best = o; who = ow; }
        }
        best.hi = __shfl_sync(0xFFFFFFFFu, best.hi, 0);
        best.lo = __shfl_sync(0xFFFFFFFFu, best.lo, 0);
        who = __shfl_sync(0xFFFFFFFFu, who, 0);
        tau = best;
        if (best.hi == ~0ull && best.lo == ~0ull) break;         // fewer than k touched documents
        if (lane == who) {
          ++head;
          cur = (head < 32) ? s_wb[lane * 32 + head] : key_inf();
        }
      }
      if (lane == 0) { s_tau = tau; s_ncand = 0u; }
    }
    __syncthreads();
    const Key128 tau = s_tau;
    Key128* s_cand = reinterpret_cast<Key128*>(s_cand_mem) + 256;  // [kBmCandCap], after the warp lists
#pragma unroll 4
    for (int i = 0; i < kBmDocsPerThread; ++i) {
      const int l = threadIdx.x + i * kBmThreads;
      if (first[l] == 0xFF) continue;
      const Key128 kk = make_key(l);
      if (!key_less(tau, kk)) {                                  // kk <= tau
        const unsigned int slot = atomicAdd(&s_ncand, 1u);
        if (slot < kBmCandCap) s_cand[slot] = kk;
      }
    }
    __syncthreads();
    const unsigned int nc = s_ncand;
    if (nc <= kBmCandCap) {
      const int emitted = static_cast<int>(nc) < k ? static_cast<int>(nc) : k;
      if (threadIdx.x < nc) {
        const Key128 me = s_cand[threadIdx.x];
        int rank = 0;
        for (unsigned int j = 0; j < nc; ++j) rank += key_less(s_cand[j], me) ? 1 : 0;
        if (rank < k) {
          cand_hi[cbase * k + rank] = me.hi;
          cand_lo[cbase * k + rank] = me.lo;
        }
      }
      if (threadIdx.x == 0) cand_cnt[cbase] = emitted;
      return;
    }
    // survivor overflow (mass ties at the threshold): fall through to the general path
    __syncthreads();
  }

  // General path (k > 32, or overflow): k rounds of block arg-min over per-thread cached bests.
  auto scan_best = [&](const Key128& after, bool have_after) -> Key128 {
    Key128 best = key_inf();
#pragma unroll 4
    for (int i = 0; i < kBmDocsPerThread; ++i) {
      const int l = threadIdx.x + i * kBmThreads;
      if (first[l] == 0xFF) continue;
      const Key128 kk = make_key(l);
      if (have_after && !key_less(after, kk)) continue;   // already emitted (kk <= after)
      if (key_less(kk, best)) best = kk;
    }
    return best;
  };
  Key128 mine = scan_best(key_inf(), false);
  int emitted = 0;
  for (int round = 0; round < k; ++round) {
    int winner;
    const Key128 w = block_argmin(mine, s_k, s_w, &winner);
    if (w.hi == ~0ull && w.lo == ~0ull) break;            // range exhausted
    if (threadIdx.x == winner) {
      cand_hi[cbase * k + round] = w.hi;
      cand_lo[cbase * k + round] = w.lo;
      mine = scan_best(w, true);
    }
    ++emitted;
  }
  if (threadIdx.x == 0) cand_cnt[cbase] = emitted;
}

// K7 (query part): k-way merge of the per-range sorted candidate lists.
// One CTA per query; thread j owns ranges j, j+blockDim, ...
__global__ void __launch_bounds__(kBmThreads)
bm25_merge_kernel(const unsigned long long* __restrict__ cand_hi, const unsigned long long* __restrict__ cand_lo,
                  const int* __restrict__ cand_cnt, int nr, int q0, int k, double* __restrict__ out_score,
                  int* __restrict__ out_doc, int* __restrict__ out_count) {
  __shared__ Key128 s_k[kBmThreads / 32];
  __shared__ int s_w[kBmThreads / 32];
  const int q = q0 + blockIdx.x;
  // per-thread heads (a thread may own several ranges when nr > blockDim; it keeps
  // the best head among them and re-scans its ranges after winning)
  constexpr int kMaxOwn = 8;                 // nr <= 2048 ranges = 16.7 M docs
  int head[kMaxOwn];
#pragma unroll
  for (int i = 0; i < kMaxOwn; ++i) head[i] = 0;
  auto best_head = [&](int* which) -> Key128 {
    Key128 best = key_inf();
    *which = -1;
#pragma unroll
    for (int i = 0; i < kMaxOwn; ++i) {
      const int r = threadIdx.x + i * kBmThreads;
      if (r >= nr) break;
      const int64_t cb = static_cast<int64_t>(q) * nr + r;
      if (head[i] < cand_cnt[cb]) {
        Key128 kk{cand_hi[cb * k + head[i]], cand_lo[cb * k + head[i]]};
        if (key_less(kk, best)) { best = kk; *which = i; }
      }
    }
    return best;
  };
  int which;
  Key128 mine = best_head(&which);
  int emitted = 0;
  for (int round = 0; round < k; ++round) {
    int winner;
    const Key128 w = block_argmin(mine, s_k, s_w, &winner);
    if (w.hi == ~0ull && w.lo == ~0ull) break;
    if (threadIdx.x == winner) {
      const int64_t o = static_cast<int64_t>(q) * k + round;
      const uint64_t ob = ~w.hi;   // orderable(score)
      const uint64_t bits = (ob & 0x8000000000000000ull) ? (ob & 0x7FFFFFFFFFFFFFFFull) : ~ob;
      out_score[o] = __longlong_as_double(static_cast<long long>(bits));
      out_doc[o] = static_cast<int>(static_cast<uint32_t>(w.lo));
#pragma unroll
      for (int i = 0; i < kMaxOwn; ++i)
        if (i == which) head[i]++;
      mine = best_head(&which);
    }
    ++emitted;
  }
  if (threadIdx.x == 0) out_count[q] = emitted;
  for (int i = emitted + threadIdx.x; i < k; i += blockDim.x) {
    out_score[static_cast<int64_t>(q) * k + i] = 0.0;
    out_doc[static_cast<int64_t>(q) * k + i] = -1;
  }
}

}  // namespace rse
