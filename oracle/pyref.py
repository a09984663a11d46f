"""Pure-Python literal restatements of the reference's in-repo arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to
/root/reference/rag_search_engine/.

These follow the reference statement by statement (dict insertion order, the
stable ``sorted``/``list.sort``, the CPython ``set`` union) so that they can be
checked here against the reference's own code (oracle/make_golden.py) and then
travel to the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple


# ----------------------------------------------------------------------------- BM25
def bm25_search(postings: Dict[str, List[Tuple[int, int]]], doclen: Dict[int, int], n_movies: int,
                query_tokens: Sequence[str], k: int = 10, k1: float = 1.5,
                b: float = 0.75) -> List[Tuple[int, float]]:
    """utils/keyword_search.py:180-250.

    ``postings[term]`` = [(doc_id, tf)] in ascending doc_id (the order the
    (term_id, doc_id) autoindex returns rows, :214-218); ``doclen`` = doclen
    table; ``n_movies`` = COUNT(movies) (:196).  Returns [(doc_id, score)].
    """
    if not query_tokens:                                     # :190-191
        return []
    if not doclen:
        return []
    avgdl = sum(doclen.values()) / len(doclen)               # AVG(length) :197-198
    if avgdl is None or avgdl == 0:                          # :199-200
        return []
    N = n_movies
    scores: Dict[int, float] = {}
    for tok in query_tokens:                                 # :205
        rows = postings.get(tok)
        if not rows:                                         # :209-210, :219-220
            continue
        df = len(rows)                                       # :222
        idf = math.log((N - df + 0.5) / (df + 0.5) + 1.0)    # :224
        for doc_id, tf in rows:                              # :226-228
            dl = doclen.get(doc_id)
            if dl is None:                                   # :235-236
                continue
            denom = tf + k1 * (1.0 - b + b * (dl / avgdl))   # :241
            score_add = idf * (tf * (k1 + 1.0) / denom)      # :242
            scores[doc_id] = scores.get(doc_id, 0.0) + score_add  # :244
    if not scores:                                           # :246-247
        return []
    return sorted(scores.items(), key=lambda x: x[1], reverse=True)[:k]  # :250


# ----------------------------------------------------------------------------- fusion helpers
def min_max_norm(nums: List[float]) -> List[float]:
    """utils/utils.py:182-191."""
    min_score = min(nums)
    max_score = max(nums)
    if min_score == max_score:
        return [1.0] * len(nums)
    return [(x - min_score) / (max_score - min_score) for x in nums]


def rrf_score(rank: int, k=60) -> float:
    """utils/utils.py:205-206."""
    return 1 / (k + rank)


def weighted_fuse(bm25_hits: Sequence[Tuple[int, float]], sem_hits: Sequence[Tuple[int, float]],
                  alpha: float, limit: int) -> List[dict]:
    """utils/hybrid_search.py:117-180 on (id, bm25 score) / (movie_id, distance) lists.

    Returns [{id, bm25, semantic, score}] (titles are resolved by the caller).
    """
    bm25_scores = [s for _, s in bm25_hits]
    bm25_norm = min_max_norm(bm25_scores) if bm25_scores else []          # :118-119
    bm25_by_id: Dict[int, float] = {}
    for (doc_id, _), norm in zip(bm25_hits, bm25_norm):                   # :122-129
        bm25_by_id[int(doc_id)] = float(norm)
    sem_sims = [1.0 - float(d) for _, d in sem_hits]                      # :134
    sem_norm = min_max_norm(sem_sims) if sem_sims else []                 # :135
    sem_by_id: Dict[int, float] = {}
    for (doc_id, _), norm in zip(sem_hits, sem_norm):                     # :138-145
        sem_by_id[int(doc_id)] = float(norm)
    all_ids = set(bm25_by_id.keys()) | set(sem_by_id.keys())              # :148
    results = []
    for doc_id in all_ids:                                                # :151
        b = bm25_by_id.get(doc_id, 0.0)                                   # :160
        s = sem_by_id.get(doc_id, 0.0)                                    # :161
        results.append({"id": doc_id, "bm25": b, "semantic": s,
                        "score": alpha * b + (1.0 - alpha) * s})          # :163
    results.sort(key=lambda r: r["score"], reverse=True)                  # :177
    return results[:limit]                                                # :180


def rrf_fuse(bm25_hits: Sequence[Tuple[int, float]], sem_hits: Sequence[Tuple[int, float]], k=60,
             limit: int = 10) -> List[dict]:
    """utils/hybrid_search.py:217-272,379 on (id, score) / (movie_id, distance) lists.

    Returns [{id, score, bm25_rank, sem_rank}].
    """
    bm25_sorted = sorted(bm25_hits, key=lambda h: h[1], reverse=True)     # :220-224
    bm25_rank: Dict[int, int] = {}
    for rank_idx, (doc_id, _) in enumerate(bm25_sorted):                  # :227-230
        bm25_rank[int(doc_id)] = rank_idx
    sem_sorted = sorted(sem_hits, key=lambda h: h[1])                     # :235-238
    sem_rank: Dict[int, int] = {}
    for rank_idx, (doc_id, _) in enumerate(sem_sorted):                   # :241-244
        sem_rank[int(doc_id)] = rank_idx
    NOT_FOUND = 99999                                                     # :247
    all_ids = set(bm25_rank.keys()) | set(sem_rank.keys())                # :248
    results = []
    for doc_id in all_ids:
        r_b = bm25_rank.get(doc_id, NOT_FOUND)
        r_s = sem_rank.get(doc_id, NOT_FOUND)
        results.append({"id": doc_id, "score": rrf_score(r_b, k) + rrf_score(r_s, k),   # :255
                        "bm25_rank": bm25_rank.get(doc_id), "sem_rank": sem_rank.get(doc_id)})
    results.sort(key=lambda r: r["score"], reverse=True)                  # :272
    return results[:limit]                                                # :379


def aggregate_movies(rows: Sequence[Tuple[int, float, int]], k: int) -> List[Tuple[int, float, int]]:
    """utils/semantic_search.py:285-317 on [(rowid, distance, movie_id)] in KNN emit order."""
    best: Dict[int, Tuple[int, float, int]] = {}
    for rowid, distance, movie_id in rows:
        prev = best.get(movie_id)
        if prev is None or distance < prev[1]:                            # :301
            best[movie_id] = (rowid, distance, movie_id)
    return sorted(best.values(), key=lambda r: r[1])[:k]                  # :314-317


def set_union_order(a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Iteration order of ``set(a) | set(b)`` built as hybrid_search.py:148,248 build it."""
    da = {int(x): None for x in a}
    db = {int(x): None for x in b}
    return list(set(da.keys()) | set(db.keys()))


# ----------------------------------------------------------------------------- vec0 KNN (second restatement)
def vec0_knn(emb, q, k: int, valid=None, chunk_size: int = 1024) -> List[Tuple[int, float]]:
    """``embedding MATCH :q AND k = :k`` on a vec0 table (semantic_search.py:254-261), restated a SECOND time,
    independently of oracle.c and in the slowest possible way, so that the committed fixture
    (tests/golden/knn_kat.json) pins the C restatement against an accidental change: numpy float32 SCALARS for the
    three sequential accumulators of sqlite-vec's ``distance_cosine_float`` (v0.1.x; a = stored row, b = query),
    Python floats (doubles) for the tail ``1 - dot / (sqrt(aMag) * sqrt(bMag))`` narrowed to f32; per 1024-row
    block ``min_idx`` (repeated arg-min with ``<=`` while the slot ascends: among equal distances the HIGHEST slot
    goes first) over the valid slots; ``merge_sorted_lists`` where the running list wins ties (earlier blocks
    first).  ``emb`` is in physical order (row r = block r // 1024, slot r % 1024); ``valid`` marks live slots.
    Returns [(row, distance)] in emit order.  Small inputs only (Python loops).  "Parity unpinned" like oracle.c:
    sqlite-vec itself is not installable here (tests/test_sqlite_vec_crosscheck.py pins both the day it is)."""
    import numpy as np
    f32 = np.float32
    n, dim = len(emb), len(q)
    run: List[Tuple[int, float]] = []
    for b0 in range(0, n, chunk_size):
        slots = list(range(b0, min(n, b0 + chunk_size)))
        dist: Dict[int, float] = {}
        for r in slots:
            if valid is not None and not valid[r]:
                continue
            dot = amag = bmag = f32(0.0)
            a = emb[r]
            for i in range(dim):
                dot = f32(dot + f32(f32(a[i]) * f32(q[i])))
                amag = f32(amag + f32(f32(a[i]) * f32(a[i])))
                bmag = f32(bmag + f32(f32(q[i]) * f32(q[i])))
            dist[r] = float(f32(1.0 - float(dot) / (math.sqrt(float(amag)) * math.sqrt(float(bmag)))))
        taken: set = set()
        top: List[int] = []
        for _ in range(min(k, chunk_size)):
            cand = [r for r in slots if r in dist and r not in taken]
            if not cand:
                break
            mi = cand[0]
            for r in slots:                                   # ascending slot, `<=`: the last equal one wins
                if r in dist and r not in taken and dist[r] <= dist[mi]:
                    mi = r
            top.append(mi)
            taken.add(mi)
        merged: List[Tuple[int, float]] = []
        pa = pb = 0
        while len(merged) < k and (pa < len(run) or pb < len(top)):
            if pa >= len(run):
                merged.append((top[pb], dist[top[pb]])); pb += 1
            elif pb >= len(top):
                merged.append(run[pa]); pa += 1
            elif run[pa][1] <= dist[top[pb]]:                 # the running list wins ties
                merged.append(run[pa]); pa += 1
            else:
                merged.append((top[pb], dist[top[pb]])); pb += 1
        run = merged
    return run
