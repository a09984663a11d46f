"""Generate tests/golden/*.json from the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden

What it pins:
  * bm25_ref.json    — reference ``KeywordSearch.search`` (keyword_search.py:180-267) on a
                       synthetic corpus indexed by the reference's own ``_rebuild_index``
                       (:102-177), ``preprocess`` frozen to a whitespace tokenizer.
  * fusion_ref.json  — reference ``HybridSearch.weighted_search`` / ``rrf_search``
                       (hybrid_search.py:91-180,183-272,379) driven through the reference's own
                       plugin seam (dummy retrievers monkeypatched on the module, exactly as
                       tests/test_hybrid_search.py:73-76 does), including structural ties.
  * knn_kat.json     — vec0 KNN known answers (duplicates across block boundaries, mass ties, holes)
                       from the pure-Python second restatement ``pyref.vec0_knn`` (sqlite-vec is not
                       installable here: this fixture pins oracle.c against pyref, not against sqlite-vec).
Floats are stored with float.hex() so the fixtures are bit-exact.
"""
from __future__ import annotations

import json
import random
import tempfile
from pathlib import Path

from oracle import ref_import

GOLDEN = Path(__file__).resolve().parents[1] / "tests" / "golden"


# ----------------------------------------------------------------------------- corpora
def make_corpus(seed: int, n_docs: int, vocab: int, dup_frac: float = 0.08):
    rnd = random.Random(seed)
    words = [f"w{i}" for i in range(vocab)]
    weights = [1.0 / (i + 1) ** 1.07 for i in range(vocab)]
    ids = sorted(rnd.sample(range(1, n_docs * 9), n_docs))
    docs = []
    for i, did in enumerate(ids):
        tl = rnd.randint(1, 4)
        bl = max(0, int(rnd.gauss(30, 14)))
        title = " ".join(rnd.choices(words, weights, k=tl))
        body = " ".join(rnd.choices(words, weights, k=bl))
        docs.append({"id": did, "title": title, "description": body})
    # exact duplicates (same title+description under another id) → exact BM25 score ties
    for _ in range(int(n_docs * dup_frac)):
        a, b = rnd.sample(range(n_docs), 2)
        docs[b]["title"] = docs[a]["title"]
        docs[b]["description"] = docs[a]["description"]
    return docs, words, weights


def make_queries(seed: int, words, weights, n: int):
    rnd = random.Random(seed)
    qs = []
    for i in range(n):
        nt = rnd.randint(1, 6)
        toks = rnd.choices(words, weights, k=nt)
        r = rnd.random()
        if r < 0.15:
            toks.insert(rnd.randrange(len(toks) + 1), "zzz_oov")       # unknown term
        elif r < 0.30:
            toks.append(toks[0])                                       # duplicate token
        elif r < 0.35:
            toks = [words[-1 - rnd.randrange(20)]]                     # rare single term
        qs.append(" ".join(toks))
    qs += ["", "   ", "zzz_oov qqq_oov", "[TITLE_END]", words[0], f"{words[0]} {words[0]} {words[0]}"]
    return qs


def bm25_golden():
    ref_kw, _, _ = ref_import.load()
    out = {"cases": []}
    for seed, n_docs, vocab, nq, k in [(11, 240, 150, 40, 10), (12, 60, 12, 25, 5), (13, 400, 600, 30, 25)]:
        docs, words, weights = make_corpus(seed, n_docs, vocab)
        queries = make_queries(seed + 100, words, weights, nq)
        with tempfile.TemporaryDirectory() as td:
            p = Path(td) / "movies.json"
            p.write_text(json.dumps({"movies": docs}), encoding="utf-8")
            ks = ref_kw.KeywordSearch.build_from_docs(docs_path=p, db_path=Path(td) / "kw.db", force=True)
            try:
                results = []
                for q in queries:
                    for (kk, k1, b) in [(k, 1.5, 0.75), (3, 1.2, 0.5)]:
                        res = ks.search(q, k=kk, k1=k1, b=b)
                        results.append({"query": q, "k": kk, "k1": k1, "b": b,
                                        "hits": [[r["id"], float(r["score"]).hex()] for r in res]})
                cur = ks.conn.cursor()
                cur.execute("SELECT COUNT(*) FROM terms"); (nterms,) = cur.fetchone()
                cur.execute("SELECT COUNT(*) FROM postings"); (npost,) = cur.fetchone()
                cur.execute("SELECT AVG(length) FROM doclen"); (avgdl,) = cur.fetchone()
            finally:
                ks.close()
        out["cases"].append({"seed": seed, "docs": docs, "results": results,
                             "n_terms": nterms, "n_postings": npost, "avgdl": float(avgdl).hex()})
    (GOLDEN / "bm25_ref.json").write_text(json.dumps(out, separators=(",", ":")))
    print("bm25_ref.json:", sum(len(c["results"]) for c in out["cases"]), "searches")


# ----------------------------------------------------------------------------- fusion
def fusion_golden():
    _, ref_hs, _ = ref_import.load()
    rnd = random.Random(7)

    class KW:
        hits = []
        def __init__(self, *a, **k): pass
        def search(self, query, k=10, k1=1.5, b=0.75): return [dict(h) for h in KW.hits[:k]]
        def close(self): pass

    class SEM:
        hits = []
        def __init__(self, *a, **k): pass
        def query_top_k(self, query_text, k=10, knn_multiplier=10): return [dict(h) for h in SEM.hits[:k]]
        def close(self): pass

    ref_hs.KeywordSearch = KW
    ref_hs.SemanticSearch = SEM
    cases = []

    def run(bm, sem, limit, alpha, k):
        KW.hits = [{"id": i, "title": f"t{i}", "description": f"d{i}", "score": s} for i, s in bm]
        SEM.hits = [{"chunk_id": 0, "distance": d, "chunk": "", "movie_id": i, "title": f"t{i}",
                     "description": f"d{i}"} for i, d in sem]
        with tempfile.TemporaryDirectory() as td:
            hs = ref_hs.HybridSearch(docs_path=None, db_path=Path(td) / "h.db")
            w = hs.weighted_search("q", alpha=alpha, limit=limit)
            r = hs.rrf_search("q", k=k, limit=limit)
        cases.append({
            "bm25": [[i, float(s).hex()] for i, s in bm[:limit]],
            "sem": [[i, float(d).hex()] for i, d in sem[:limit]],
            "limit": limit, "alpha": float(alpha).hex(), "k": k,
            "weighted": [[x["id"], float(x["bm25"]).hex(), float(x["semantic"]).hex(),
                          float(x["score"]).hex()] for x in w],
            "rrf": [[x["id"], float(x["score"]).hex(), x["bm25_rank"], x["sem_rank"]] for x in r],
        })

    import numpy as np
    # SURVEY App. B known-answer case
    run([(912345, 9.5), (17, 7.25), (400001, 7.0), (8, 3.5), (33, 1.125)],
        [(5, 0.20440000295639038), (700000, 0.23229999840259552), (17, 0.2764), (31, 0.5), (8, 0.75)],
        5, 0.5, 60)
    # reference unit-test shapes (tests/test_hybrid_search.py)
    run([(1, 3.0), (3, 2.0), (2, 1.0)], [(2, 0.2), (4, 0.4)], 10, 0.7, 60)
    run([(1, 3.0), (3, 2.0), (2, 1.0)], [], 10, 0.8, 60)
    run([], [(2, 0.2), (4, 0.4)], 3, 0.5, 60)
    run([], [], 5, 0.5, 60)
    run([(7, 2.5)], [(7, 0.3)], 5, 0.5, 60)            # single element → min_max all-equal → 1.0
    run([(7, 2.5), (9, 2.5)], [(11, 0.3), (12, 0.3)], 5, 0.25, 60)   # all-equal lists
    for case in range(160):
        limit = rnd.choice([1, 2, 3, 5, 10, 10, 10, 20, 40, 64])
        id_space = rnd.choice([max(30, 4 * limit), 200, 10**6, 10**9, 2**40])
        nb = rnd.randint(0, limit)
        ns = rnd.randint(0, limit)
        overlap = rnd.random()
        bids = rnd.sample(range(1, id_space), nb)
        sids = []
        pool = list(bids)
        while len(sids) < ns:
            if pool and rnd.random() < overlap:
                sids.append(pool.pop(rnd.randrange(len(pool))))
            else:
                c = rnd.randrange(1, id_space)
                if c not in sids and c not in bids:
                    sids.append(c)
        bs = sorted((rnd.choice([rnd.uniform(0.5, 20.0), float(rnd.randint(1, 4))]) for _ in range(nb)),
                    reverse=True)
        ds = sorted(float(np.float32(rnd.choice([rnd.uniform(0.05, 1.2), rnd.randint(1, 4) / 8.0])))
                    for _ in range(ns))
        alpha = rnd.choice([0.0, 0.2, 0.5, 0.5, 0.7, 1.0, rnd.random()])
        k = rnd.choice([60, 60, 60.0, 1, 10, 0.5])
        run(list(zip(bids, bs)), list(zip(sids, ds)), limit, alpha, k)
    (GOLDEN / "fusion_ref.json").write_text(json.dumps({"cases": cases}, separators=(",", ":")))
    print("fusion_ref.json:", len(cases), "cases")


# ----------------------------------------------------------------------------- vec0 KNN known answers
def knn_kat_inputs(case: dict):
    """(emb f32[n, dim], valid bool[n], queries f32[nq, dim]) of a fixture case, rebuilt from its seed — the fixture
    stores the seed and the ANSWERS, not the matrix."""
    import numpy as np
    rng = np.random.default_rng(case["seed"])
    n, dim = case["n"], case["dim"]
    emb = rng.standard_normal((n, dim)).astype(np.float32)
    emb *= rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)           # norms are recomputed per pair
    emb = (np.round(emb * 4) / 4).astype(np.float32) if case["coarse"] else emb   # coarse grid: exact distance ties
    for a, b in case["dups"]:
        emb[b] = emb[a]
    valid = np.ones(n, bool)
    valid[case["holes"]] = False
    Q = rng.standard_normal((case["nq"], dim)).astype(np.float32)
    for j, r in enumerate(case["query_rows"]):
        Q[j] = emb[r] * np.float32(1.5)
    return emb, valid, Q


def knn_golden():
    """tests/golden/knn_kat.json — SURVEY §8c (3): vec0 KNN known answers with duplicate rows on both sides of
    1024-row block boundaries, mass ties (coarse grid), deleted slots and k above a block's size, computed by the
    pure-Python restatement (pyref.vec0_knn); tests/test_oracle_golden.py checks oracle.c — the checker of every GPU
    parity test — against them."""
    from oracle import pyref
    cases = []
    specs = [dict(seed=21, n=2100, dim=8, coarse=False, nq=5, query_rows=[1023, 5], holes=[3, 1024, 2099],
                  dups=[[1023, 1024], [1023, 2047], [5, 900], [5, 2050], [2047, 2048]], ks=[1, 10, 100]),
             dict(seed=22, n=1500, dim=4, coarse=True, nq=4, query_rows=[7], holes=[0, 7, 1499],
                  dups=[[7, 8], [7, 1030], [100, 1023]], ks=[3, 64, 1100])]
    for sp in specs:
        emb, valid, Q = knn_kat_inputs(sp)
        answers = []
        for k in sp["ks"]:
            for qi in range(len(Q)):
                res = pyref.vec0_knn(emb, Q[qi], k, valid=valid)
                answers.append({"k": k, "q": qi, "rows": [r for r, _ in res],
                                "dist": [float(d).hex() for _, d in res]})
        cases.append({**{kk: v for kk, v in sp.items() if kk != "ks"}, "answers": answers})
    (GOLDEN / "knn_kat.json").write_text(json.dumps({"cases": cases}, separators=(",", ":")))
    print("knn_kat.json:", sum(len(c["answers"]) for c in cases), "answers")


if __name__ == "__main__":
    import sys
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if "--knn-only" not in sys.argv:
        bm25_golden()
        fusion_golden()
    knn_golden()
