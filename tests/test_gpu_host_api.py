"""GPU: the drop-in classes (same names / signatures / dicts as the reference) against the oracle
pipeline on a reference-format SQLite file, plus the reference's own unit-test scenarios
(rag_search_engine/tests/test_hybrid_search.py, test_keyword_search.py) replayed on the mirrors."""
import hashlib
import json
import sqlite3

import numpy as np
import pytest

import oracle
from oracle import pyref
from helpers.corpus import build_postings

pytestmark = pytest.mark.gpu
DIM = 384


class HashEncoder:
    """Frozen stand-in for SentenceTransformer('all-MiniLM-L6-v2') (semantic_search.py:45): a
    deterministic unit vector per text, so identical titles give byte-identical embeddings."""

    def get_sentence_embedding_dimension(self):
        return DIM

    def encode(self, texts, show_progress_bar=False):
        out = np.empty((len(texts), DIM), np.float32)
        for i, t in enumerate(texts):
            seed = int.from_bytes(hashlib.sha256(t.encode()).digest()[:8], "little")
            v = np.random.default_rng(seed).standard_normal(DIM).astype(np.float32)
            out[i] = v / np.linalg.norm(v)
        return out


def make_docs(n=400, seed=21):
    import random
    rnd = random.Random(seed)
    words = [f"w{i}" for i in range(80)]
    weights = [1.0 / (i + 1) for i in range(80)]
    docs = []
    ids = sorted(rnd.sample(range(5, 9 * n), n))
    for did in ids:
        title = " ".join(rnd.choices(words, weights, k=rnd.randint(1, 3)))
        sents = [" ".join(rnd.choices(words, weights, k=rnd.randint(3, 9))) + rnd.choice(".!?") for _ in range(rnd.randint(1, 9))]
        docs.append({"id": did, "title": title.title(), "description": " ".join(sents)})
    for _ in range(n // 10):                      # duplicate titles → exact KNN ties across movies
        a, b = rnd.sample(range(n), 2)
        docs[b]["title"] = docs[a]["title"]
    return docs, words


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rag_search_engine_b200 import store
    from rag_search_engine_b200.textutil import whitespace_tokenizer
    docs, words = make_docs()
    d = tmp_path_factory.mktemp("db")
    enc = HashEncoder()
    db = store.write_reference_db(d / "movies.db", docs, whitespace_tokenizer, embed=enc.encode)
    (d / "movies.json").write_text(json.dumps({"movies": docs}))
    return {"db": db, "docs": docs, "words": words, "enc": enc, "dir": d}


def oracle_semantic(built, text, k, mult=10):
    """semantic_search.py:250-338 restated with the oracle: vec0 KNN → aggregation → dicts."""
    from rag_search_engine_b200.textutil import chunk_text
    conn = sqlite3.connect(built["db"])
    chunks = conn.execute("SELECT id, movie_id, chunk_index, max_chunk_size, overlap FROM chunks ORDER BY id").fetchall()
    movies = {r[0]: (r[1], r[2]) for r in conn.execute("SELECT id, title, description FROM movies")}
    texts = [chunk_text(*movies[c[1]], c[2], c[3], c[4]) for c in chunks]
    emb = built["enc"].encode(texts)
    q = built["enc"].encode([text])[0]
    d, rows = oracle.vec0_knn(emb, q, max(k * mult, k))
    hits = pyref.aggregate_movies([(int(r), float(dd), chunks[r][1]) for dd, r in zip(d, rows)], k)
    conn.close()
    return [{"chunk_id": r, "distance": dd, "chunk": texts[r], "movie_id": m, "title": movies[m][0],
             "description": movies[m][1]} for r, dd, m in hits]


def test_keyword_search_matches_reference_semantics(built):
    from rag_search_engine_b200 import KeywordSearch
    from rag_search_engine_b200.textutil import whitespace_tokenizer
    ks = KeywordSearch.open_existing(db_path=built["db"], tokenizer=whitespace_tokenizer)
    try:
        postings, doclen = build_postings(built["docs"])
        meta = {d["id"]: d for d in built["docs"]}
        for q in ["w0", "w1 w5 w9", "w3 w3", "zzz w2", "w70 w71", "W0 w0 W0"]:
            got = ks.search(q, k=7)
            want = pyref.bm25_search(postings, doclen, len(doclen), q.lower().split(), k=7)
            assert [(r["id"], r["score"]) for r in got] == want
            for r in got:
                assert list(r) == ["title", "description", "id", "score"]            # dict key order (:264)
                assert r["title"] == meta[r["id"]]["title"] and r["description"] == meta[r["id"]]["description"]
        assert ks.search("", k=5) == [] and ks.search("qqq_unknown", k=5) == []       # :190-191, :246-247
        assert ks.count_movies() == len(built["docs"])
        ks.verify_db()
    finally:
        ks.close()


def test_reference_keyword_unit_test_scenario(tmp_path):
    """rag_search_engine/tests/test_keyword_search.py:32-61 replayed on the mirror (build_from_docs)."""
    from rag_search_engine_b200 import KeywordSearch
    from rag_search_engine_b200.textutil import whitespace_tokenizer
    movies = {"movies": [
        {"id": 1, "title": "The Matrix", "description": "A computer hacker learns about the true nature of reality."},
        {"id": 2, "title": "Inception", "description": "A thief enters dreams to steal secrets."},
        {"id": 3, "title": "Toy Story", "description": "Toys come to life when humans are not around."}]}
    p = tmp_path / "movies.json"
    p.write_text(json.dumps(movies))
    ks = KeywordSearch.build_from_docs(docs_path=p, db_path=tmp_path / "kw.db", force=True, tokenizer=whitespace_tokenizer)
    try:
        results = ks.search("Matrix", k=5)
        assert results and results[0]["title"] == "The Matrix" and "score" in results[0]
        scores = [r["score"] for r in results]
        assert scores == sorted(scores, reverse=True)
        assert ks.search("", k=5) == []
    finally:
        ks.close()


def test_semantic_search_matches_oracle(built):
    from rag_search_engine_b200 import SemanticSearch
    ss = SemanticSearch.open_existing(db_path=built["db"], encoder=built["enc"])
    try:
        docs = built["docs"]
        for text, k in [(docs[3]["title"], 5), ("w1 w2 w3 something else", 5), (docs[10]["description"].split(". ")[0], 10),
                        (docs[40]["title"], 3)]:
            got = ss.query_top_k(text, k=k)
            want = oracle_semantic(built, text, k)
            assert [list(g) for g in got] == [["chunk_id", "distance", "chunk", "movie_id", "title", "description"]] * len(got)
            assert got == want
        with pytest.raises(ValueError, match="cannot embed empty text"):              # :218-219
            ss.query_top_k("   ")
        vec = ss.generate_embedding("hello world")
        assert vec.shape == (1, DIM) and vec.dtype == np.float32
        ss.verify_db()
    finally:
        ss.close()


def test_hybrid_search_matches_reference_pipeline(built):
    from rag_search_engine_b200 import HybridSearch
    from rag_search_engine_b200.textutil import whitespace_tokenizer
    hs = HybridSearch(docs_path=None, db_path=built["db"], tokenizer=whitespace_tokenizer, encoder=built["enc"])
    try:
        postings, doclen = build_postings(built["docs"])
        meta = {d["id"]: d for d in built["docs"]}
        for q, limit in [("w0 w1", 5), (built["docs"][7]["title"], 10), ("w2 w2 w9 zzz", 3), ("w60", 10)]:
            bm = pyref.bm25_search(postings, doclen, len(doclen), q.lower().split(), k=limit)
            sem = [(h["movie_id"], h["distance"]) for h in oracle_semantic(built, q, limit)]
            got = hs.rrf_search(q, k=60, limit=limit)
            want = pyref.rrf_fuse(bm, sem, 60, limit)
            assert [(g["id"], g["score"], g["bm25_rank"], g["sem_rank"]) for g in got] == \
                   [(w["id"], w["score"], w["bm25_rank"], w["sem_rank"]) for w in want]
            assert all(list(g) == ["id", "title", "description", "score", "bm25_rank", "sem_rank"] for g in got)
            assert all(g["title"] == meta[g["id"]]["title"] for g in got)
            gw = hs.weighted_search(q, alpha=0.5, limit=limit)
            ww = pyref.weighted_fuse(bm, sem, 0.5, limit)
            assert [(g["id"], g["bm25"], g["semantic"], g["score"]) for g in gw] == \
                   [(w["id"], w["bm25"], w["semantic"], w["score"]) for w in ww]
            # the one-call on-device batch path gives the same fused lists
            qv = built["enc"].encode([q])
            gb = hs.rrf_search_batch([q.lower().split()], qv, k=60, limit=limit)[0]
            assert [(g["id"], g["score"], g["bm25_rank"], g["sem_rank"]) for g in gb] == \
                   [(w["id"], w["score"], w["bm25_rank"], w["sem_rank"]) for w in want]
            gwb = hs.weighted_search_batch([q.lower().split()], qv, 0.5, limit=limit)[0]
            assert [(g["id"], g["bm25"], g["semantic"], g["score"]) for g in gwb] == \
                   [(w["id"], w["bm25"], w["semantic"], w["score"]) for w in ww]
        assert hs.rrf_search("w0", limit=5, rerank_method="nonsense") == hs.rrf_search("w0", limit=5)   # :370-373
        # the serving loop (two batches in flight) yields, in order, what the blocking batch call returns
        qs = ["w0 w1", built["docs"][7]["title"], "w2 w2 w9 zzz", "w60", "w3 w4 w5", "w1"]
        batches = [([q.lower().split() for q in qs[i:i + n]], built["enc"].encode(qs[i:i + n]))
                   for i, n in ((0, 2), (2, 1), (3, 0), (3, 3))]
        streamed = list(hs.rrf_search_stream(iter(batches), k=60, limit=5))
        assert len(streamed) == len(batches) and streamed[2] == []
        for (toks, qv), got in zip(batches, streamed):
            if toks:
                assert got == hs.rrf_search_batch(toks, qv, k=60, limit=5)
    finally:
        hs.close()


# ---- the reference's own fusion tests through the same monkeypatch seam (tests/test_hybrid_search.py)
class DummyKeywordSearch:
    def __init__(self, *_, **__): pass
    def search(self, query, k=10, k1=1.5, b=0.75):
        return [{"id": 1, "title": "Doc1", "description": "d1", "score": 3.0},
                {"id": 3, "title": "Doc3", "description": "d3", "score": 2.0},
                {"id": 2, "title": "Doc2", "description": "d2", "score": 1.0}]
    def close(self): pass


class DummySemanticSearchOnly2:
    def __init__(self, *_, **__): pass
    def query_top_k(self, query_text, k=10, knn_multiplier=10):
        return [{"chunk_id": 0, "distance": 0.2, "chunk": "d2 chunk", "movie_id": 2, "title": "Doc2", "description": "d2"},
                {"chunk_id": 1, "distance": 0.4, "chunk": "d4 chunk", "movie_id": 4, "title": "Doc4", "description": "d4"}]
    def close(self): pass


class DummyEmptySemantic:
    def __init__(self, *_, **__): pass
    def query_top_k(self, query_text, k=10, knn_multiplier=10): return []
    def close(self): pass


class DummyEmptyKeyword:
    def __init__(self, *_, **__): pass
    def search(self, query, k=10, k1=1.5, b=0.75): return []
    def close(self): pass


def _make_hybrid(monkeypatch, kw_cls, sem_cls, tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from rag_search_engine_b200 import hybrid_search as hs_mod
    monkeypatch.setattr(hs_mod, "KeywordSearch", kw_cls)
    monkeypatch.setattr(hs_mod, "SemanticSearch", sem_cls)
    return hs_mod.HybridSearch(docs_path=None, db_path=tmp_path / "hybrid.db")


def test_weighted_search_combines_scores(monkeypatch, tmp_path):
    hs = _make_hybrid(monkeypatch, DummyKeywordSearch, DummySemanticSearchOnly2, tmp_path)
    try:
        results = hs.weighted_search("query", alpha=0.7, limit=10)
        assert results
        for r in results:
            assert 0.0 <= r["bm25"] <= 1.0 and 0.0 <= r["semantic"] <= 1.0
            assert pytest.approx(0.7 * r["bm25"] + 0.3 * r["semantic"]) == r["score"]
        want = pyref.weighted_fuse([(1, 3.0), (3, 2.0), (2, 1.0)], [(2, 0.2), (4, 0.4)], 0.7, 10)
        assert [(r["id"], r["bm25"], r["semantic"], r["score"]) for r in results] == \
               [(w["id"], w["bm25"], w["semantic"], w["score"]) for w in want]
        assert [r["title"] for r in results] == [f"Doc{r['id']}" for r in results]
    finally:
        hs.close()


def test_weighted_search_degrades_to_bm25_when_no_semantic(monkeypatch, tmp_path):
    hs = _make_hybrid(monkeypatch, DummyKeywordSearch, DummyEmptySemantic, tmp_path)
    try:
        assert [r["id"] for r in hs.weighted_search("query", alpha=0.8, limit=10)] == [1, 3, 2]
    finally:
        hs.close()


def test_rrf_search_degrades_to_bm25_when_only_keyword(monkeypatch, tmp_path):
    hs = _make_hybrid(monkeypatch, DummyKeywordSearch, DummyEmptySemantic, tmp_path)
    try:
        assert [r["id"] for r in hs.rrf_search("query", limit=3)] == [1, 3, 2]
    finally:
        hs.close()


def test_rrf_search_degrades_to_semantic_when_only_semantic(monkeypatch, tmp_path):
    hs = _make_hybrid(monkeypatch, DummyEmptyKeyword, DummySemanticSearchOnly2, tmp_path)
    try:
        assert [r["id"] for r in hs.rrf_search("query", limit=3)][:2] == [2, 4]
        assert hs.rrf_search("query", limit=3)[0]["bm25_rank"] is None
    finally:
        hs.close()


def test_both_empty(monkeypatch, tmp_path):
    hs = _make_hybrid(monkeypatch, DummyEmptyKeyword, DummyEmptySemantic, tmp_path)
    try:
        assert hs.weighted_search("q", alpha=0.5, limit=5) == [] and hs.rrf_search("q", limit=5) == []
    finally:
        hs.close()
