/*
 * rse.h — C-ABI of librse.so, the B200-native retrieval core that replaces the
 * query hot path of JWSch4fer/rag-search-engine.
 *
 * Plain C: opaque handle, raw pointers and sizes, int status codes, no
 * exceptions and no torch/numpy types across the boundary.  The reference is
 * pure Python, so the reference-side binding is a ctypes stub;
 * each entry point names the reference interface it replaces (file:line are
 * relative to /root/reference/rag_search_engine/).
 *
 * Conventions
 *   - every function returns RSE_OK (0) or a negative RSE_ERR_* code;
 *     rse_last_error(h) gives the message (h may be NULL for rse_create).
 *   - "host" pointers are caller-owned CPU buffers; "dev" pointers are
 *     caller-owned device buffers on the handle's device (the *_dev entry
 *     points are what the torch.distributed host plumbing uses; they enqueue
 *     on the handle's stream and do not synchronise).
 *   - one handle = one device + one CUDA stream; calls on a handle are
 *     synchronous unless stated and not re-entrant (the reference objects are
 *     single-threaded too: utils/basesearch_db.py:40).
 *   - there is NO CPU fallback: without a CUDA device rse_create fails.
 */
#ifndef RSE_H_
#define RSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSE_ABI_VERSION 2

#define RSE_OK 0
#define RSE_ERR_INVALID -1     /* bad argument */
#define RSE_ERR_CUDA -2        /* CUDA runtime error (message has the call) */
#define RSE_ERR_STATE -3       /* e.g. query before load */
#define RSE_ERR_UNSUPPORTED -4 /* outside documented limits */
#define RSE_ERR_NOMEM -5

/* Limits (documented in DESIGN.md §limits) */
#define RSE_VEC0_BLOCK 1024     /* sqlite-vec default chunk_size (semantic_search.py:96-99 passes none) */
#define RSE_MAX_KPRIME 4096     /* sqlite-vec caps k at 4096 */
#define RSE_MAX_BM25_K 256
#define RSE_MAX_FUSE_LIMIT 128
#define RSE_MAX_QUERY_TOKENS 255
#define RSE_RRF_NOT_FOUND 99999 /* hybrid_search.py:247 */
#define RSE_MAX_STASH 16        /* staged batches that can be parked in HBM (rse_hybrid_stash) */

/* tie_mode for the fusion entry points */
#define RSE_TIE_REFERENCE 0 /* CPython set iteration order of the id union (hybrid_search.py:148,248) */
#define RSE_TIE_BY_ID 1     /* deterministic by ascending id (north_star wording) */

typedef struct rse_index rse_index;

int rse_abi_version(void);

/* Create a handle bound to CUDA device `device`.  Fails (RSE_ERR_CUDA) when no
 * device is present.  Replaces the object construction at
 * utils/hybrid_search.py:41-54 (one KeywordSearch + one SemanticSearch). */
int rse_create(int32_t device, rse_index **out);
void rse_destroy(rse_index *h);
const char *rse_last_error(const rse_index *h);

/* Use an externally owned CUDA stream (cudaStream_t as void*), e.g. torch's current
 * stream.  NULL is the legacy default stream (what torch uses by default), NOT "unset";
 * rse_use_own_stream goes back to the handle's private non-blocking stream. */
int rse_set_stream(rse_index *h, void *cuda_stream);
int rse_use_own_stream(rse_index *h);
/* Block until everything enqueued on the handle's stream has finished. */
int rse_synchronize(rse_index *h);

/* ------------------------------------------------------------------ a1: vec0 store
 * Replaces the vec0 virtual table `chunk_embeddings` (semantic_search.py:94-101)
 * as the thing the KNN scans.  Rows are given in vec0 PHYSICAL layout: row i of
 * `emb` is (block pos_base/1024 + i/1024, slot i%1024); `valid[i]==0` marks an
 * empty slot (NULL: all valid).  `rowid` (NULL: rowid = pos_base + i) is
 * chunks.id; `movie_idx` (NULL: aggregation unavailable) is the dense index of
 * chunks.movie_id in the caller's sorted id universe, or -1 when the JOIN at
 * semantic_search.py:262-276 would drop the row.  pos_base must be a multiple of
 * 1024 (row shards are aligned to vec0 blocks).  use_fma selects the aarch64
 * (contracted) arithmetic of sqlite-vec instead of the x86-64 one.
 * The host variant copies; the dev variant BORROWS emb_dev (caller keeps it
 * alive) and copies the small side arrays. */
int rse_load_embeddings(rse_index *h, const float *emb_host, int64_t n_rows, int32_t dim,
                        const uint8_t *valid_host, const int64_t *rowid_host,
                        const int32_t *movie_idx_host, int64_t pos_base);
int rse_attach_embeddings_dev(rse_index *h, const float *emb_dev, int64_t n_rows, int32_t dim,
                              const uint8_t *valid_dev, const int64_t *rowid_dev,
                              const int32_t *movie_idx_dev, int64_t pos_base);
int rse_set_fma(rse_index *h, int32_t use_fma);
/* K4, the tcgen05 path for query batches (probe -> filter -> EXACT re-score, results identical to the
 * streaming scan): 0 = auto (batches of >= RSE_TC_MIN_BATCH queries on >= 256 k rows, dim 384), 1 = never,
 * 2 = whenever the shape allows it.  The probe/filter GEMM runs over an fp16 normalised shadow of the corpus
 * (knn_tc3.cuh; built once at the first batch that needs it, +768 B per row of device memory).  Measured on
 * S-600k (bench.py knn_small_batches, r02): the K4 chain costs 0.75-0.82 ms per call for 1..64 queries, the exact
 * scan 1.21 / 1.41 / 1.65 / 1.72 / 2.32 / 3.98 ms for 1 / 2 / 3 / 4 / 8 / 16 — K4 wins from the first query on.  A
 * single query takes the exact streaming scan (north_star: batch-1 = the HBM-bound kernel; no shadow needed) until
 * a batch has built the shadow, and K4 from then on; mode 2 sends it through K4 from the start. */
#define RSE_TC_MIN_BATCH 2
int rse_set_tc_mode(rse_index *h, int32_t mode);

/* Which BM25 kernels serve rse_bm25 / rse_hybrid for k <= 32 (results are bit-identical in every mode):
 * 0 = fixed-point streaming kernel + exact re-score of the finalists (default), 1 = streaming kernel that
 * keeps the reference's summation order with one barrier per (range, token) slice, 2 = general kernel only
 * (one CTA per range and query; what k > 32, > 16 tokens and flagged queries always use). */
int rse_set_bm25_mode(rse_index *h, int32_t mode);

/* KNN: `embedding MATCH :q AND k = :k ... ORDER BY knn.distance`
 * (semantic_search.py:254-279).  Q is [nq, dim] fp32.  Outputs are [nq, kprime]
 * in vec0 emit order (distance asc, block asc, slot desc); out_count[q] <= kprime
 * rows are valid.  out_pos / out_rowid / out_movie_idx may be NULL. */
int rse_knn(rse_index *h, const float *q_host, int32_t nq, int32_t kprime, float *out_dist,
            int64_t *out_pos, int64_t *out_rowid, int32_t *out_movie_idx, int32_t *out_count);

/* KNN + best-chunk-per-movie aggregation = SemanticSearch.query_top_k numerics
 * (semantic_search.py:250-317): top-kprime chunks, first row per movie, first k.
 * Outputs [nq, k]. */
int rse_knn_movies(rse_index *h, const float *q_host, int32_t nq, int32_t k, int32_t kprime,
                   float *out_dist, int64_t *out_chunk_rowid, int32_t *out_movie_idx,
                   int32_t *out_count);

/* Device-pointer building blocks for the row-sharded multi-GPU path.
 * Local top-kprime as packed candidates: cand_dev is [nq, kprime, 3] int64 =
 * {key, rowid, movie_idx}; key = (orderable(distance) << 32) | (global_pos ^ 1023),
 * so ascending key IS the vec0 emit order; unused tail entries have key = -1
 * (all ones).  q_dev is [nq, dim] fp32 on the device.  Asynchronous on the handle's stream when
 * the exact scan serves the batch; the tensor-core path (rse_set_tc_mode) synchronises once to
 * read its per-query overflow flags. */
int rse_knn_local_dev(rse_index *h, const float *q_dev, int32_t nq, int32_t kprime,
                      int64_t *cand_dev);
/* Keep a row-sharded step free of host round trips: with deferral enabled rse_knn_local_dev never
 * synchronises — a query the tensor-core path could not finish (survivor overflow: a mass tie at the K'-th
 * distance) keeps key = -1 candidates and is only COUNTED: rse_knn_flags_dev enqueues `*flagged_dev +=
 * (flagged queries of the last rse_knn_local_dev)` on the handle's stream.  The caller lets that counter travel
 * with the step's results and repeats a step whose global count is non-zero with deferral off. */
int rse_set_defer_flags(rse_index *h, int32_t enabled);
int rse_knn_flags_dev(rse_index *h, int32_t *flagged_dev);
/* Merge n_lists candidate lists per query (gathered_dev is [n_lists, nq, kprime, 3],
 * the layout all_gather produces), keep the first kprime under the key order, then
 * aggregate per movie like rse_knn_movies.  Outputs are DEVICE buffers [nq, k]. */
int rse_knn_merge_movies_dev(rse_index *h, const int64_t *gathered_dev, int32_t n_lists,
                             int32_t nq, int32_t k, int32_t kprime, float *out_dist_dev,
                             int64_t *out_chunk_rowid_dev, int32_t *out_movie_idx_dev,
                             int32_t *out_count_dev);

/* ------------------------------------------------------------------ a6: BM25 index
 * Replaces the SQLite tables terms/postings/doclen as the thing
 * KeywordSearch.search (keyword_search.py:180-250) scores against.
 *   indptr[T+1], doc_idx[P] (dense doc index, ascending within a term — the
 *   order the (term_id, doc_id) autoindex yields, :214-218), tf[P] =
 *   len(positions) (:227-228), df[T] = len(rows) (:222; may exceed the CSR
 *   length when postings without a doclen row were dropped, :235-236),
 *   dl[M] = doclen.length, n_movies = COUNT(movies) (:196), avgdl = AVG(length)
 *   (:197-198). */
int rse_load_bm25(rse_index *h, const int64_t *indptr, const uint32_t *doc_idx, const uint32_t *tf,
                  const int64_t *df, int64_t n_terms, int64_t n_postings, const uint32_t *dl,
                  int64_t n_docs, int64_t n_movies, double avgdl);

/* BM25 top-k for a batch of tokenised queries.  tok_indptr[nq+1] delimits each
 * query's tokens in `term_rows` (CSR term row per token in query order,
 * duplicates kept, -1 = unknown term, skipped: keyword_search.py:205-210).
 * Outputs [nq, k]: scores (fp64, bit-exact reference arithmetic :224,:241-244),
 * dense doc indices, and per-query counts.  Order = sorted(..., reverse=True)
 * with dict-insertion tie order (:250). */
int rse_bm25(rse_index *h, const int32_t *tok_indptr, const int32_t *term_rows, int32_t nq,
             int32_t k, double k1, double b, double *out_score, int32_t *out_doc_idx,
             int32_t *out_count);

/* ------------------------------------------------------------------ a8-a10: fusion
 * ids are the caller's document ids (movies.id).  Inputs [nq, limit] with
 * per-query counts; BM25 lists are score-descending, semantic lists
 * distance-ascending (what the retrievers return).  Distances are doubles: the
 * reference does float(hit["distance"]) on whatever the plugged-in retriever
 * returns (hybrid_search.py:134); vec0 distances are f32 values widened.  Outputs [nq, limit].
 * weighted: hybrid_search.py:117-180 (min_max_norm utils.py:182-191,
 *           alpha*b + (1-alpha)*s :163).
 * rrf:      hybrid_search.py:217-272,379 (0-based ranks, NOT_FOUND=99999 :247,
 *           1/(k+r_b) + 1/(k+r_s) :255); out ranks are -1 for None. */
int rse_fuse_weighted(rse_index *h, int32_t nq, int32_t limit, double alpha, int32_t tie_mode,
                      const int64_t *bm25_id, const double *bm25_score, const int32_t *bm25_count,
                      const int64_t *sem_id, const double *sem_dist, const int32_t *sem_count,
                      int64_t *out_id, double *out_bm25, double *out_sem, double *out_score,
                      int32_t *out_count);
int rse_fuse_rrf(rse_index *h, int32_t nq, int32_t limit, double k, int32_t tie_mode,
                 const int64_t *bm25_id, const double *bm25_score, const int32_t *bm25_count,
                 const int64_t *sem_id, const double *sem_dist, const int32_t *sem_count,
                 int64_t *out_id, double *out_score, int32_t *out_bm25_rank, int32_t *out_sem_rank,
                 int32_t *out_count);

/* ------------------------------------------------------------------ hybrid, end to end
 * HybridSearch.rrf_search / weighted_search numerics for a batch (hybrid_search.py:91-180,
 * 183-272): BM25(k=limit) + KNN(k=limit, kprime=max(limit*knn_multiplier, limit),
 * semantic_search.py:251) + fusion, all on the device, one host round trip.
 * doc_ids[n_docs] / movie_ids[n_movies_uni] map dense indices to movies.id (set with
 * rse_set_id_tables).  mode: 0 = rrf (param = k), 1 = weighted (param = alpha).
 * Outputs [nq, limit]; out_a/out_b = (bm25_rank, sem_rank) as doubles for rrf
 * (-1 = None) or (bm25_norm, sem_norm) for weighted. */
int rse_set_id_tables(rse_index *h, const int64_t *doc_ids, int64_t n_docs, const int64_t *movie_ids,
                      int64_t n_movie_ids);
int rse_hybrid(rse_index *h, int32_t mode, double param, int32_t tie_mode, int32_t limit,
               int32_t knn_multiplier, int32_t nq, const float *q_host, const int32_t *tok_indptr,
               const int32_t *term_rows, double k1, double b, int64_t *out_id, double *out_score,
               double *out_a, double *out_b, int32_t *out_count);
/* The same call split in three, for callers that keep a query batch resident in HBM
 * (rse_hybrid == stage + run + fetch):
 *   stage: upload query vectors + tokens (+ host-computed idf, keyword_search.py:224);
 *   run:   BM25 + KNN + aggregation + fusion on the staged batch; asynchronous on the
 *          handle's stream, results stay on the device; may be repeated;
 *   fetch: copy the [nq, limit] results to the host and synchronise. */
int rse_hybrid_stage(rse_index *h, int32_t nq, const float *q_host, const int32_t *tok_indptr,
                     const int32_t *term_rows);
int rse_hybrid_run(rse_index *h, int32_t mode, double param, int32_t tie_mode, int32_t limit,
                   int32_t knn_multiplier, double k1, double b);
/* Row-sharded variant of `run`: the KNN stage is replaced by the merge of
 * n_lists gathered candidate lists ([n_lists, nq, kprime, 3] from rse_knn_local_dev on every
 * shard + all_gather); the staged batch is this rank's query slice.  Outputs are DEVICE
 * buffers [nq, limit] (they feed the final all_gather).  Asynchronous on the handle's stream. */
int rse_hybrid_run_merged_dev(rse_index *h, int32_t mode, double param, int32_t tie_mode, int32_t limit,
                              int32_t knn_multiplier, double k1, double b, const int64_t *gathered_dev,
                              int32_t n_lists, int64_t *out_id_dev, double *out_score_dev,
                              double *out_a_dev, double *out_b_dev, int32_t *out_count_dev);
int rse_hybrid_fetch(rse_index *h, int32_t limit, int64_t *out_id, double *out_score, double *out_a,
                     double *out_b, int32_t *out_count);
/* EXCHANGE the staged batch with stash slot `slot` (0 .. RSE_MAX_STASH-1); pointer swaps on the host, no device
 * work.  stage(A); stash(0); stage(B); stash(1) parks two batches in HBM; a later stash(0) makes A the staged
 * batch again (and parks whatever was staged).  For callers that rotate over several RESIDENT batches. */
int rse_hybrid_stash(rse_index *h, int32_t slot);
/* Pipelined form of rse_hybrid for a serving loop: at most TWO batches in flight per handle.
 *   submit : stage (host buffers -> pinned -> device, asynchronous) + run + an asynchronous copy of
 *            the results into a pinned slot owned by the handle; returns a ticket.  It returns as
 *            soon as the batch is ENQUEUED — it never waits for the device — so the caller submits
 *            batch i+1 before collecting batch i and the device does not idle between batches.
 *            Caller buffers may be reused as soon as submit returns.
 *   collect: wait for the ticket's results and copy them out; tickets are collected in submission
 *            order.  out_nq / out_limit (optional) report the batch's shape.  The tensor-core
 *            path's per-query overflow flags (a mass tie at the K'-th distance; rse_hybrid waits
 *            for them in mid-batch) are checked HERE: a batch with a flagged query is re-run
 *            through rse_hybrid from the handle's copies of its inputs before collect returns.
 * Results are those of rse_hybrid on the same inputs.  A third submit before a collect returns
 * RSE_ERR_STATE.  Not to be interleaved with stage/run/fetch between a submit and its collect. */
int rse_hybrid_submit(rse_index *h, int32_t mode, double param, int32_t tie_mode, int32_t limit,
                      int32_t knn_multiplier, int32_t nq, const float *q_host,
                      const int32_t *tok_indptr, const int32_t *term_rows, double k1, double b,
                      int64_t *ticket);
int rse_hybrid_collect(rse_index *h, int64_t ticket, int32_t *out_nq, int32_t *out_limit,
                       int64_t *out_id, double *out_score, double *out_a, double *out_b,
                       int32_t *out_count);
/* Wait for and DISCARD every ticket still in flight (a consumer that stops early, an error between a
 * submit and its collect): afterwards the handle accepts submits again.  out_dropped (optional) = how many. */
int rse_hybrid_drain(rse_index *h, int32_t *out_dropped);

/* ------------------------------------------------------------------ §8(f2)/(f3): the text encoders on the device
 * The reference encodes every query text with SentenceTransformer("all-MiniLM-L6-v2") on the CPU
 * (utils/semantic_search.py:45, :211-222: model.encode -> float32[1, 384]) and reranks the fused hits with
 * CrossEncoder("cross-encoder/ms-marco-TinyBERT-L2-v2").predict(pairs) (utils/hybrid_search.py:279-312).  Both are
 * HuggingFace BertModels; here one implementation runs either next to the index (csrc/encoder.cuh):
 *   head 0 : BertModel -> mean pooling over the tokens -> L2 normalise  (the sentence-transformers pipeline);
 *            output [n_seq, hidden]
 *   head 1 : BertModel -> pooler (tanh(dense([CLS]))) -> classifier (num_labels = 1); output [n_seq] logits
 * Weights are handed over tensor by tensor under their HuggingFace state_dict names (any prefix such as "bert." or
 * "0.auto_model." is ignored); tokenisation stays on the host (out of scope): the caller passes the
 * token ids of the batch PACKED — ids[T] (+ optional token_type_ids[T], NULL = all 0) and cu_seqlens[n_seq + 1]
 * (sequence s owns tokens cu[s] .. cu[s+1]; what attention_mask == 1 selects).  fp32 throughout.
 * rse_encode_dev leaves the result on the device (out_dev on the handle's device, asynchronous on its stream), so
 * rse_hybrid_stage_dev can take the query vectors from there: text in, no host hop for the vectors. */
#define RSE_MAX_ENCODERS 2
typedef struct rse_encoder_config {
  int32_t vocab_size, hidden, layers, heads, intermediate, max_positions, type_vocab;
  float ln_eps;
  int32_t head; /* 0 = mean pool + L2 normalise, 1 = BERT pooler + 1-logit classifier */
} rse_encoder_config;
int rse_encoder_create(rse_index *h, int32_t slot, const rse_encoder_config *cfg);
int rse_encoder_set_tensor(rse_index *h, int32_t slot, const char *name, const float *data, int64_t n_elem);
int rse_encoder_finalize(rse_index *h, int32_t slot);
/* Which GEMM the encoder's linear layers run on (results agree to ~1e-6 relative; both meet the 1e-5 bar):
 * 0 = tcgen05 kind::tf32 tensor cores with a 3-pass hi/lo split of both operands (fp32-class accuracy; default
 * when every GEMM dimension is a multiple of 128), 1 = fp32 SIMT (fmaf). */
int rse_encoder_set_mode(rse_index *h, int32_t slot, int32_t mode);
int rse_encode(rse_index *h, int32_t slot, const int32_t *ids, const int32_t *type_ids, const int32_t *cu_seqlens,
               int32_t n_seq, float *out_host);
int rse_encode_dev(rse_index *h, int32_t slot, const int32_t *ids, const int32_t *type_ids,
                   const int32_t *cu_seqlens, int32_t n_seq, float *out_dev);
/* rse_hybrid_stage with the query vectors already ON THE DEVICE ([nq, dim] fp32, e.g. rse_encode_dev's output);
 * tokens still come from the host.  Followed by rse_hybrid_run / rse_hybrid_fetch as usual. */
int rse_hybrid_stage_dev(rse_index *h, int32_t nq, const float *q_dev, const int32_t *tok_indptr,
                         const int32_t *term_rows);

/* ------------------------------------------------------------------ multi-GPU: row shards over one box
 * One process per GPU, one handle per process; the HANDLE owns the NCCL communicator (NCCL is bound at run time:
 * libnccl.so.2, or $RSE_NCCL_LIB).  The chunk-embedding matrix is cut by row into contiguous shards aligned to
 * vec0 blocks (rse_load_embeddings / rse_attach_embeddings_dev with pos_base = the shard's first global position);
 * the BM25 index is replicated; rank r owns the contiguous query slice [r*nq/N ...) of every batch (the split of
 * sharded.query_slices).  The reference has no counterpart (single process, semantic_search.py:254-261 scans one
 * table); results equal those of one handle holding the whole corpus.
 *   rse_comm_unique_id : rank 0 creates the 128-byte id and hands it to the other ranks by any means
 *                        (torch.distributed broadcast, a file, an environment variable);
 *   rse_comm_init      : every rank, collectively (ncclCommInitRank on the handle's device). */
#define RSE_COMM_ID_BYTES 128
int rse_comm_unique_id(uint8_t *out_id);
int rse_comm_init(rse_index *h, const uint8_t *id_bytes, int32_t n_ranks, int32_t rank);
int rse_comm_destroy(rse_index *h);
int rse_comm_info(rse_index *h, int32_t *n_ranks, int32_t *rank, int32_t *nccl_version);
/* The exchange step alone: cand_dev [nq, kprime, 3] (this shard's local top-K' of ALL queries, rse_knn_local_dev)
 * -> mine_dev [n_ranks, ns, kprime, 3] (every shard's candidates for THIS rank's ns queries): an all-to-all by
 * query slice (grouped ncclSend / ncclRecv) on the handle's stream. */
int rse_comm_exchange_candidates_dev(rse_index *h, const int64_t *cand_dev, int32_t nq, int32_t kprime,
                                     int64_t *mine_dev);
int rse_comm_allgather_dev(rse_index *h, const void *send_dev, void *recv_dev, int64_t bytes_per_rank);
/* configs[4]: sharded KNN + per-movie aggregation in one call — local top-K' of all nq_all queries on this
 * shard, candidate exchange, merge + aggregation of this rank's query slice.  Outputs [ns, k] DEVICE buffers.
 * flagged_dev (optional, int32 on the device): never wait for the host; *flagged_dev += the queries this shard
 * could not finish (see rse_set_defer_flags).  NULL: blocking and always exact.  Collective: every rank calls. */
int rse_knn_sharded_dev(rse_index *h, const float *q_all_dev, int32_t nq_all, int32_t k, int32_t kprime,
                        float *out_dist_dev, int64_t *out_chunk_rowid_dev, int32_t *out_movie_idx_dev,
                        int32_t *out_count_dev, int32_t *flagged_dev);
/* The hybrid step, row-sharded: the staged batch (rse_hybrid_stage) is THIS rank's query slice (its tokens; the
 * staged vectors are not used), q_all_dev holds the vectors of the whole batch.  KNN local top-K' of all queries
 * (the slice's BM25 runs underneath the filter pass) -> exchange -> merge + aggregation + fusion of the slice.
 * Outputs [ns, limit] DEVICE buffers; flagged_dev as above.  Collective. */
int rse_hybrid_sharded_run_dev(rse_index *h, int32_t mode, double param, int32_t tie_mode, int32_t limit,
                               int32_t knn_multiplier, double k1, double b, const float *q_all_dev, int32_t nq_all,
                               int64_t *out_id_dev, double *out_score_dev, double *out_a_dev, double *out_b_dev,
                               int32_t *out_count_dev, int32_t *flagged_dev);

/* ------------------------------------------------------------------ introspection
 * Counters since the last rse_stats_reset: kernels launched by this library,
 * device ms of the last call per stage (CUDA events on the handle's stream). */
typedef struct rse_stats {
  int64_t kernel_launches;
  int64_t knn_scan_launches;
  double scan_ms_total;        /* Σ device time of the K1 scan launches (event pairs, resolved in rse_get_stats) */
  int64_t scan_launches_timed; /* how many scan launches that sum covers */
  double last_knn_total_ms;    /* last rse_knn / rse_knn_movies call, H2D → D2H */
  double last_bm25_ms;
  double last_fuse_ms;
  int64_t emb_rows;
  int32_t emb_dim;
  int64_t bm25_postings;
  int64_t bm25_docs;
  int64_t tc_filter_launches;   /* K4 tensor-core filter passes (each serves ≤ 256 queries) */
  int64_t tc_queries;           /* queries answered through K4 */
  int64_t tc_fallback_queries;  /* of those, re-run through the exact scan (survivor overflow) */
  int64_t tc_second_chance_queries; /* of tc_queries, answered by the second filter pass with a tightened bound */
  int64_t bm25_queries;           /* queries scored by the BM25 kernels */
  int64_t bm25_fallback_queries;  /* of those, handed to the general kernel by the streaming path (device-side flag) */
  int64_t bm25_finalists;         /* documents re-scored exactly by the fixed-point path's finish kernel */
  int64_t bm25_candidates;        /* candidates the fixed-point kernel handed to the finish kernel */
  int64_t h2d_bytes;              /* bytes moved host -> device by the query entry points (not the index loads) */
  int64_t d2h_bytes;              /* bytes moved device -> host by the query entry points */
} rse_stats;
int rse_get_stats(rse_index *h, rse_stats *out);
/* Survivor counts of the most recent K4 filter pass (one per query of its 256-query block; a count above the
 * survivor cap of 8192 means the query overflowed and was re-run).  Synchronises.  n <= 256. */
int rse_tc_last_survivors(rse_index *h, int32_t *out_counts, int32_t n);
/* The order statistic j of the K4 probe's row sample (DESIGN.md §5 "The probe's sample"): the smallest j with
 * P(Binomial(kprime, 1 / stride) >= j) <= 1e-7, capped at kprime.  Pure host arithmetic — exported so that the claim
 * can be checked against an independent binomial tail (tests/test_host_cpu.py). */
int rse_tc_probe_rank(int32_t kprime, int32_t stride);
int rse_stats_reset(rse_index *h);
/* Enable CUDA-event timing: an event pair around every scan launch (no extra syncs). */
int rse_set_timing(rse_index *h, int32_t enabled);

#ifdef __cplusplus
}
#endif
#endif /* RSE_H_ */
