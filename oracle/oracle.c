/*
 * oracle/oracle.c — CPU restatement of the reference's query hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rag_search_engine_b200/ may import,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or the timed CPU baseline.
 *
 * PARITY STATUS
 *   vec0 KNN (a1): **parity unpinned**.  The arithmetic lives in the
 *     third-party `sqlite-vec` extension (PyPI `sqlite-vec`, unpinned in
 *     /root/reference/requirements.txt:3, not vendored, not installable here).
 *     This file restates the published v0.1.x algorithm
 *     (`distance_cosine_float`, `min_idx`, `merge_sorted_lists`,
 *     `vec0Filter_knn_chunks_iter`) and is anchored on the reference's call
 *     sites: rag_search_engine/utils/semantic_search.py:94-101 (DDL,
 *     `float[dim] distance_metric=cosine`, default chunk_size 1024) and
 *     :254-279 (`embedding MATCH :q AND k = :k ... ORDER BY knn.distance`).
 *   BM25 (a6): pinned against the reference's own
 *     rag_search_engine/utils/keyword_search.py:180-250 imported here
 *     (oracle/make_golden.py → tests/golden/bm25_*.json).
 *
 * Compile with -O2 -ffp-contract=off (x86-64 sqlite-vec wheels are built
 * without FMA; CPython double arithmetic is never contracted).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define VEC0_CHUNK_SIZE 1024 /* sqlite-vec default chunk_size; the reference passes none (semantic_search.py:96-99) */

/* ---------------------------------------------------------------------------
 * sqlite-vec `distance_cosine_float`: three sequential fp32 accumulators,
 * double sqrt / mul / div / sub, result narrowed to f32.
 * a = stored row, b = query (vec0Filter_knn_chunks_iter passes base vector first).
 * use_fma=1 restates an aarch64 build where GCC contracts a*b+c.
 * ------------------------------------------------------------------------- */
float oracle_cosine_distance(const float *a, const float *b, int64_t n, int use_fma) {
  float dot = 0.0f, aMag = 0.0f, bMag = 0.0f;
  if (use_fma) {
    for (int64_t i = 0; i < n; i++) {
      dot = fmaf(a[i], b[i], dot);
      aMag = fmaf(a[i], a[i], aMag);
      bMag = fmaf(b[i], b[i], bMag);
    }
  } else {
    for (int64_t i = 0; i < n; i++) {
      dot += a[i] * b[i];
      aMag += a[i] * a[i];
      bMag += b[i] * b[i];
    }
  }
  return (float)(1.0 - ((double)dot / (sqrt((double)aMag) * sqrt((double)bMag))));
}

/* Σ a[i]^2 in the same sequential fp32 order (exposed for tests of the GPU's
 * precomputed row magnitudes). */
float oracle_sq_magnitude(const float *a, int64_t n, int use_fma) {
  float m = 0.0f;
  if (use_fma)
    for (int64_t i = 0; i < n; i++) m = fmaf(a[i], a[i], m);
  else
    for (int64_t i = 0; i < n; i++) m += a[i] * a[i];
  return m;
}

void oracle_all_distances(const float *emb, int64_t n_rows, int32_t dim, const float *q,
                          int use_fma, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; r++)
    out[r] = oracle_cosine_distance(emb + r * (int64_t)dim, q, dim, use_fma);
}

/* sqlite-vec `min_idx`: pick the k smallest of one chunk's distances by
 * repeated arg-min.  The inner comparison is `<=` while scanning ascending
 * slot, so among equal distances the HIGHEST slot is taken first. */
static int vec0_min_idx(const float *distances, int n, const uint8_t *candidates, int *out, int k,
                        uint8_t *taken) {
  memset(taken, 0, (size_t)n);
  for (int ik = 0; ik < k; ik++) {
    int mi = 0;
    while (mi < n && (taken[mi] || !candidates[mi])) mi++;
    if (mi >= n) return ik;
    for (int i = 0; i < n; i++) {
      if (distances[i] <= distances[mi] && !taken[i] && candidates[i]) mi = i;
    }
    out[ik] = mi;
    taken[mi] = 1;
  }
  return k;
}

/* sqlite-vec `merge_sorted_lists`: running list `a` wins ties (`a <= b`), so
 * earlier chunks beat later ones at equal distance. */
static int64_t vec0_merge(const float *a, const int64_t *a_ids, int64_t a_len, const float *b,
                          const int64_t *b_ids, const int *b_top, int64_t b_len, float *out,
                          int64_t *out_ids, int64_t out_len) {
  int64_t pa = 0, pb = 0;
  for (int64_t i = 0; i < out_len; i++) {
    if (pa >= a_len && pb >= b_len) return i;
    if (pa >= a_len) {
      out[i] = b[b_top[pb]];
      out_ids[i] = b_ids[b_top[pb]];
      pb++;
    } else if (pb >= b_len) {
      out[i] = a[pa];
      out_ids[i] = a_ids[pa];
      pa++;
    } else if (a[pa] <= b[b_top[pb]]) {
      out[i] = a[pa];
      out_ids[i] = a_ids[pa];
      pa++;
    } else {
      out[i] = b[b_top[pb]];
      out_ids[i] = b_ids[b_top[pb]];
      pb++;
    }
  }
  return out_len;
}

/*
 * Literal restatement of the vec0 KNN scan (semantic_search.py:254-261).
 *   emb      [n_rows, dim]   rows in PHYSICAL order
 *   pos      [n_rows] or NULL: physical position chunk*1024+slot of each row,
 *            strictly ascending; NULL means pos[r] = r (fresh build,
 *            semantic_search.py:164-206 inserts rowids 0..C-1 in order).
 *   out_row  index into emb (the caller maps to rowid).
 * Returns the number of results (<= k).
 */
int64_t oracle_vec0_knn(const float *emb, int64_t n_rows, int32_t dim, const int64_t *pos,
                        const float *q, int32_t k, int use_fma, float *out_dist,
                        int64_t *out_row) {
  if (k <= 0 || n_rows <= 0) return 0;
  float *cd = (float *)malloc(sizeof(float) * VEC0_CHUNK_SIZE);
  int64_t *cids = (int64_t *)malloc(sizeof(int64_t) * VEC0_CHUNK_SIZE);
  uint8_t *cand = (uint8_t *)malloc(VEC0_CHUNK_SIZE);
  uint8_t *taken = (uint8_t *)malloc(VEC0_CHUNK_SIZE);
  int *top = (int *)malloc(sizeof(int) * VEC0_CHUNK_SIZE);
  float *run_d = (float *)malloc(sizeof(float) * (size_t)k);
  int64_t *run_i = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
  float *tmp_d = (float *)malloc(sizeof(float) * (size_t)k);
  int64_t *tmp_i = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
  int64_t run_len = 0;

  int64_t r = 0;
  while (r < n_rows) {
    int64_t chunk = (pos ? pos[r] : r) / VEC0_CHUNK_SIZE;
    memset(cand, 0, VEC0_CHUNK_SIZE);
    int64_t r_end = r;
    while (r_end < n_rows && (pos ? pos[r_end] : r_end) / VEC0_CHUNK_SIZE == chunk) {
      int slot = (int)((pos ? pos[r_end] : r_end) % VEC0_CHUNK_SIZE);
      cd[slot] = oracle_cosine_distance(emb + r_end * (int64_t)dim, q, dim, use_fma);
      cids[slot] = r_end;
      cand[slot] = 1;
      r_end++;
    }
    int kk = k < VEC0_CHUNK_SIZE ? k : VEC0_CHUNK_SIZE;
    int used = vec0_min_idx(cd, VEC0_CHUNK_SIZE, cand, top, kk, taken);
    int64_t n = vec0_merge(run_d, run_i, run_len, cd, cids, top, used, tmp_d, tmp_i, k);
    memcpy(run_d, tmp_d, sizeof(float) * (size_t)n);
    memcpy(run_i, tmp_i, sizeof(int64_t) * (size_t)n);
    run_len = n;
    r = r_end;
  }
  memcpy(out_dist, run_d, sizeof(float) * (size_t)run_len);
  memcpy(out_row, run_i, sizeof(int64_t) * (size_t)run_len);
  free(cd); free(cids); free(cand); free(taken); free(top);
  free(run_d); free(run_i); free(tmp_d); free(tmp_i);
  return run_len;
}

/* Same result derived from the closed-form total order
 * (distance asc, chunk asc, slot desc) — SURVEY App. A.2.  Used by the tests
 * to prove the literal scan and the key order agree, and as the fast CPU
 * checker at sizes where O(1024*k) per chunk is too slow. */
typedef struct { float d; int64_t code; int64_t row; } okey_t;
static int okey_cmp(const void *x, const void *y) {
  const okey_t *a = (const okey_t *)x, *b = (const okey_t *)y;
  if (a->d < b->d) return -1;
  if (a->d > b->d) return 1;
  if (a->code < b->code) return -1;
  if (a->code > b->code) return 1;
  return 0;
}
int64_t oracle_vec0_knn_keyorder(const float *emb, int64_t n_rows, int32_t dim, const int64_t *pos,
                                 const float *q, int32_t k, int use_fma, float *out_dist,
                                 int64_t *out_row) {
  if (k <= 0 || n_rows <= 0) return 0;
  okey_t *keys = (okey_t *)malloc(sizeof(okey_t) * (size_t)n_rows);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rows; r++) {
    int64_t p = pos ? pos[r] : r;
    keys[r].d = oracle_cosine_distance(emb + r * (int64_t)dim, q, dim, use_fma);
    keys[r].code = p ^ (VEC0_CHUNK_SIZE - 1); /* chunk asc, slot desc */
    keys[r].row = r;
  }
  qsort(keys, (size_t)n_rows, sizeof(okey_t), okey_cmp);
  int64_t n = n_rows < k ? n_rows : k;
  for (int64_t i = 0; i < n; i++) { out_dist[i] = keys[i].d; out_row[i] = keys[i].row; }
  free(keys);
  return n;
}

/* Best-chunk-per-movie aggregation (semantic_search.py:285-317): walk the KNN
 * rows in order, keep the first row per movie (a later row replaces it only on
 * a strictly smaller distance, :301), stable sort by distance, truncate to k. */
int64_t oracle_aggregate_movies(const float *dist, const int64_t *movie, int64_t n, int32_t k,
                                int64_t *out_sel /* indices into the input */) {
  int64_t *best = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
  int64_t nb = 0;
  for (int64_t i = 0; i < n; i++) {
    int64_t j;
    for (j = 0; j < nb; j++) if (movie[best[j]] == movie[i]) break;
    if (j == nb) best[nb++] = i;
    else if (dist[i] < dist[best[j]]) best[j] = i; /* dict keeps original insertion slot */
  }
  /* stable insertion sort by distance (sorted() is stable, :314-317) */
  for (int64_t i = 1; i < nb; i++) {
    int64_t v = best[i], j = i - 1;
    while (j >= 0 && dist[best[j]] > dist[v]) { best[j + 1] = best[j]; j--; }
    best[j + 1] = v;
  }
  int64_t m = nb < k ? nb : k;
  for (int64_t i = 0; i < m; i++) out_sel[i] = best[i];
  free(best);
  return m;
}

/* ---------------------------------------------------------------------------
 * BM25 over CSR postings — restates keyword_search.py:196-250.
 *   indptr[T+1], doc[P] (dense doc index, ascending within a term — the order
 *   SQLite returns rows via the (term_id, doc_id) autoindex), tf[P], df[T]
 *   (= len(rows), :222), dl[M], N = COUNT(movies) (:196), avgdl (:197-198).
 *   Query = term rows in token order (duplicates kept, :205); -1 = unknown
 *   term (skipped, :209-210).
 * Tie order of sorted(..., reverse=True) (:250) is dict insertion order:
 * (index of first token whose postings hold the doc, doc asc).
 * Scratch: acc[M] doubles + first[M] int32, caller-provided, zero/-1 filled
 * on entry and restored on exit (touched list).
 * ------------------------------------------------------------------------- */
typedef struct { double s; int32_t first; int32_t doc; } bkey_t;
static int bkey_cmp(const void *x, const void *y) {
  const bkey_t *a = (const bkey_t *)x, *b = (const bkey_t *)y;
  if (a->s > b->s) return -1;
  if (a->s < b->s) return 1;
  if (a->first != b->first) return a->first < b->first ? -1 : 1;
  if (a->doc != b->doc) return a->doc < b->doc ? -1 : 1;
  return 0;
}

int64_t oracle_bm25_query(const int64_t *indptr, const uint32_t *doc, const uint32_t *tf,
                          const int64_t *df, const uint32_t *dl, int64_t M, int64_t N, double avgdl,
                          const int32_t *terms, int32_t n_terms, int32_t k, double k1, double b,
                          double *acc, int32_t *first, int32_t *touched, double *out_score,
                          int32_t *out_doc) {
  (void)M;
  int64_t nt = 0;
  for (int32_t t = 0; t < n_terms; t++) {
    int32_t term = terms[t];
    if (term < 0) continue;
    int64_t lo = indptr[term], hi = indptr[term + 1];
    int64_t dfi = df[term];
    if (dfi == 0) continue;
    double idf = log(((double)N - (double)dfi + 0.5) / ((double)dfi + 0.5) + 1.0);
    for (int64_t p = lo; p < hi; p++) {
      uint32_t d = doc[p];
      double tfd = (double)tf[p];
      double denom = tfd + k1 * (1.0 - b + b * ((double)dl[d] / avgdl));
      double add = idf * (tfd * (k1 + 1.0) / denom);
      if (first[d] < 0) { first[d] = t; touched[nt++] = (int32_t)d; acc[d] = 0.0 + add; }
      else acc[d] = acc[d] + add;
    }
  }
  bkey_t *keys = (bkey_t *)malloc(sizeof(bkey_t) * (size_t)(nt > 0 ? nt : 1));
  for (int64_t i = 0; i < nt; i++) {
    int32_t d = touched[i];
    keys[i].s = acc[d]; keys[i].first = first[d]; keys[i].doc = d;
    acc[d] = 0.0; first[d] = -1;
  }
  qsort(keys, (size_t)nt, sizeof(bkey_t), bkey_cmp);
  int64_t n = nt < k ? nt : k;
  for (int64_t i = 0; i < n; i++) { out_score[i] = keys[i].s; out_doc[i] = keys[i].doc; }
  free(keys);
  return n;
}

/* Batch driver (OpenMP over queries) — used by the CPU-baseline timing. */
void oracle_bm25_batch(const int64_t *indptr, const uint32_t *doc, const uint32_t *tf,
                       const int64_t *df, const uint32_t *dl, int64_t M, int64_t N, double avgdl,
                       const int32_t *tok_indptr, const int32_t *terms, int32_t nq, int32_t k,
                       double k1, double b, double *out_score, int32_t *out_doc,
                       int32_t *out_count) {
#pragma omp parallel
  {
    double *acc = (double *)calloc((size_t)M, sizeof(double));
    int32_t *first = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
    int32_t *touched = (int32_t *)malloc(sizeof(int32_t) * (size_t)M);
    for (int64_t i = 0; i < M; i++) first[i] = -1;
#pragma omp for schedule(dynamic, 1)
    for (int32_t q = 0; q < nq; q++) {
      out_count[q] = (int32_t)oracle_bm25_query(
          indptr, doc, tf, df, dl, M, N, avgdl, terms + tok_indptr[q],
          tok_indptr[q + 1] - tok_indptr[q], k, k1, b, acc, first, touched,
          out_score + (int64_t)q * k, out_doc + (int64_t)q * k);
    }
    free(acc); free(first); free(touched);
  }
}

/* Batch driver for KNN + aggregation (OpenMP over queries; each query is the
 * single-threaded literal vec0 scan, as sqlite-vec is single-threaded). */
void oracle_knn_movies_batch(const float *emb, int64_t n_rows, int32_t dim, const int64_t *pos,
                             const int64_t *movie_of_row, const float *Q, int32_t nq, int32_t k,
                             int32_t kprime, int use_fma, int literal, float *out_dist,
                             int64_t *out_row, int32_t *out_count) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int32_t qi = 0; qi < nq; qi++) {
    float *d = (float *)malloc(sizeof(float) * (size_t)kprime);
    int64_t *rw = (int64_t *)malloc(sizeof(int64_t) * (size_t)kprime);
    int64_t *mv = (int64_t *)malloc(sizeof(int64_t) * (size_t)kprime);
    int64_t *sel = (int64_t *)malloc(sizeof(int64_t) * (size_t)kprime);
    int64_t n = literal
        ? oracle_vec0_knn(emb, n_rows, dim, pos, Q + (int64_t)qi * dim, kprime, use_fma, d, rw)
        : oracle_vec0_knn_keyorder(emb, n_rows, dim, pos, Q + (int64_t)qi * dim, kprime, use_fma, d, rw);
    for (int64_t i = 0; i < n; i++) mv[i] = movie_of_row[rw[i]];
    int64_t m = oracle_aggregate_movies(d, mv, n, k, sel);
    for (int64_t i = 0; i < m; i++) {
      out_dist[(int64_t)qi * k + i] = d[sel[i]];
      out_row[(int64_t)qi * k + i] = rw[sel[i]];
    }
    out_count[qi] = (int32_t)m;
    free(d); free(rw); free(mv); free(sel);
  }
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Thread count of the following calls (OMP_NUM_THREADS is only read when the OpenMP runtime starts). */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
