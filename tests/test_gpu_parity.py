"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes → librse.so),
against the oracle on the same seeded inputs and against the committed golden fixtures.

Bars (BASELINE.json north_star): ids / ranks / order bit-exact; BM25 and RRF scores bit-exact
(fp64, same association); cosine distances bit-exact (the scan reproduces the reference's
sequential fp32 sum), which is stricter than the stated 1e-5 relative tolerance.
"""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from oracle import pyref
from helpers.corpus import build_postings, to_csr

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def inject_ties(rng, emb, n_dup):
    n = len(emb)
    src = rng.integers(0, n, n_dup)
    dst = rng.integers(0, n, n_dup)
    emb[dst] = emb[src]
    # block-boundary duplicates
    for s, d in [(7, 1023), (7, 1024), (7, 2047), (9, 2048)]:
        if d < n:
            emb[d] = emb[s]
    return emb


# ----------------------------------------------------------------------------- KNN
@pytest.mark.parametrize("n,dim,kprime,nq", [(5000, 384, 100, 5), (40000, 384, 100, 9), (3000, 384, 1, 3),
                                              (1500, 384, 1500, 2), (4100, 384, 4096, 1), (2500, 48, 50, 4),
                                              (37, 4, 10, 3), (1025, 384, 100, 8), (31, 384, 5, 1),
                                              (3000, 512, 20, 7), (2000, 768, 10, 16), (1200, 100, 15, 3)])   # any-width scan
def test_knn_matches_oracle_bit_exact(fresh_index, n, dim, kprime, nq):
    rng = np.random.default_rng(n + dim + kprime)
    emb = inject_ties(rng, unit_rows(rng, n, dim), n // 20)
    Q = unit_rows(rng, nq, dim)
    Q[0] = emb[7]                                     # exact hit + duplicates → distance ties
    if nq > 1:
        Q[1] = emb[min(9, n - 1)] + 0.05 * rng.standard_normal(dim).astype(np.float32)
    fresh_index.load_embeddings(emb)
    dist, pos, rowid, movie, cnt = fresh_index.knn(Q, kprime)
    for qi in range(nq):
        od, orow = oracle.vec0_knn(emb, Q[qi], kprime, literal=(n <= 8000))
        assert cnt[qi] == len(orow) == min(kprime, n)
        assert pos[qi, :cnt[qi]].tolist() == orow.tolist()
        assert dist[qi, :cnt[qi]].view(np.uint32).tolist() == od.view(np.uint32).tolist()
        assert rowid[qi, :cnt[qi]].tolist() == orow.tolist()


def test_knn_all_distances_bit_exact_via_full_k(fresh_index):
    """k' = n: every row's f32 distance comes back and must equal the oracle's bit for bit."""
    rng = np.random.default_rng(3)
    n = 4000
    emb = unit_rows(rng, n, 384) * rng.uniform(0.5, 2.0, (n, 1)).astype(np.float32)   # non-unit norms
    q = rng.standard_normal(384).astype(np.float32)
    fresh_index.load_embeddings(emb)
    dist, pos, _, _, cnt = fresh_index.knn(q[None], n)
    assert cnt[0] == n
    all_d = oracle.all_distances(emb, q)
    got = np.empty(n, np.float32)
    got[pos[0]] = dist[0]
    assert got.view(np.uint32).tolist() == all_d.view(np.uint32).tolist()


def test_knn_holes_rowids_posbase_and_fma(fresh_index):
    rng = np.random.default_rng(11)
    n = 5000
    emb = inject_ties(rng, unit_rows(rng, n, 384), 200)
    valid = (rng.random(n) > 0.1).astype(np.uint8)
    rowid = rng.permutation(n).astype(np.int64) + 1000
    base = 7 * 1024
    fresh_index.load_embeddings(emb, valid=valid, rowid=rowid, pos_base=base)
    Q = unit_rows(rng, 4, 384)
    Q[0] = emb[7]
    dist, pos, rid, _, cnt = fresh_index.knn(Q, 64)
    keep = np.nonzero(valid)[0]
    for qi in range(4):
        od, orow = oracle.vec0_knn(emb[keep], Q[qi], 64, pos=(keep + base).astype(np.int64))
        assert pos[qi, :cnt[qi]].tolist() == (keep[orow] + base).tolist()
        assert rid[qi, :cnt[qi]].tolist() == rowid[keep[orow]].tolist()
        assert dist[qi, :cnt[qi]].view(np.uint32).tolist() == od.view(np.uint32).tolist()
    # fewer valid rows than k'
    valid2 = np.zeros(n, np.uint8); valid2[[5, 1030, 4000]] = 1
    fresh_index.load_embeddings(emb, valid=valid2)
    dist, pos, _, _, cnt = fresh_index.knn(Q[:1], 10)
    assert cnt[0] == 3 and sorted(pos[0, :3].tolist()) == [5, 1030, 4000]


def test_knn_fma_mode_matches_fma_oracle():
    from rag_search_engine_b200 import _lib
    idx = _lib.Index(0)
    try:
        rng = np.random.default_rng(12)
        emb = unit_rows(rng, 3000, 384)
        Q = unit_rows(rng, 3, 384)
        idx.set_fma(True)
        idx.load_embeddings(emb)
        dist, pos, _, _, cnt = idx.knn(Q, 20)
        for qi in range(3):
            od, orow = oracle.vec0_knn(emb, Q[qi], 20, use_fma=True)
            assert pos[qi].tolist() == orow.tolist()
            assert dist[qi].view(np.uint32).tolist() == od.view(np.uint32).tolist()
    finally:
        idx.close()


def test_knn_movies_matches_oracle(fresh_index):
    rng = np.random.default_rng(21)
    n_movies = 900
    per = rng.integers(2, 12, n_movies)
    movie_of = np.repeat(np.arange(n_movies, dtype=np.int32), per)
    n = len(movie_of)
    emb = unit_rows(rng, n, 384)
    # make each movie's chunks similar so one movie fills several of the top-k' slots
    centers = unit_rows(rng, n_movies, 384)
    emb = centers[movie_of] + 0.3 * emb
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    emb = inject_ties(rng, emb.astype(np.float32), 100)
    Q = (centers[rng.integers(0, n_movies, 12)] + 0.2 * unit_rows(rng, 12, 384)).astype(np.float32)
    fresh_index.load_embeddings(emb, movie_idx=movie_of)
    for k, kp in [(5, 50), (10, 100), (10, 10), (3, 7)]:
        dist, rowid, movie, cnt = fresh_index.knn_movies(Q, k, kp)
        od, orow, oc = oracle.knn_movies_batch(emb, Q, movie_of, k, kp)
        assert cnt.tolist() == oc.tolist()
        for qi in range(len(Q)):
            c = cnt[qi]
            assert rowid[qi, :c].tolist() == orow[qi, :c].tolist()
            assert movie[qi, :c].tolist() == movie_of[orow[qi, :c]].tolist()
            assert dist[qi, :c].view(np.uint32).tolist() == od[qi, :c].view(np.uint32).tolist()
    # fewer than k distinct movies inside the top-k' → fewer than k results
    dist, rowid, movie, cnt = fresh_index.knn_movies(centers[:4].astype(np.float32), 10, 12)
    od, orow, oc = oracle.knn_movies_batch(emb, centers[:4].astype(np.float32), movie_of, 10, 12)
    assert cnt.tolist() == oc.tolist() and (cnt < 10).any()


def test_knn_size_independent_properties_at_scale(fresh_index):
    """600k-shaped shard slice (≈1M rows): sortedness under the emit order, self-hit at
    distance ≈ 0, idempotence, and agreement with the key-order oracle on a sample."""
    import torch
    from rag_search_engine_b200 import synth
    se = synth.synth_embeddings(120_000, seed=1234, device="cuda")
    C = se.emb.shape[0]
    fresh_index.attach_embeddings_dev(se.emb.data_ptr(), C, 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(),
                                      keepalive=se)
    Q = synth.synth_query_vectors(se.emb, 16, seed=99).cpu().numpy()
    rows = np.array([0, 1023, 1024, C - 1])
    Q[:4] = se.emb[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    d1, p1, _, m1, c1 = fresh_index.knn(Q, 100)
    d2, p2, _, _, _ = fresh_index.knn(Q, 100)
    assert (p1 == p2).all() and (d1.view(np.uint32) == d2.view(np.uint32)).all()          # idempotent
    for qi in range(16):
        code = p1[qi] ^ 1023
        order = sorted(range(100), key=lambda i: (d1[qi, i], code[i]))
        assert order == list(range(100))                                                   # emit order
    for i, r in enumerate(rows):
        assert r in p1[i, :8] and d1[i, 0] <= 1e-6                                         # self hit
    emb_host = se.emb.cpu().numpy()
    for qi in (0, 5, 11):
        od, orow = oracle.vec0_knn(emb_host, Q[qi], 100, literal=False)
        assert p1[qi].tolist() == orow.tolist()
        assert d1[qi].view(np.uint32).tolist() == od.view(np.uint32).tolist()


# ----------------------------------------------------------------------------- BM25
def test_bm25_matches_reference_golden(fresh_index):
    for case in json.loads((GOLDEN / "bm25_ref.json").read_text())["cases"]:
        postings, doclen = build_postings(case["docs"])
        csr = to_csr(postings, doclen)
        fresh_index.load_bm25(csr["indptr"], csr["doc"], csr["tf"], csr["df"], csr["dl"], len(csr["dl"]), csr["avgdl"])
        by = {}
        for r in case["results"]:
            by.setdefault((r["k"], r["k1"], r["b"]), []).append(r)
        for (k, k1, b), rs in by.items():
            tok_indptr, terms = [0], []
            for r in rs:
                terms += [csr["row"].get(t, -1) for t in r["query"].lower().split()]
                tok_indptr.append(len(terms))
            sc, dc, cnt = fresh_index.bm25(np.array(tok_indptr, np.int32), np.array(terms, np.int32), k, k1, b)
            for i, r in enumerate(rs):
                got = [[int(csr["doc_ids"][dc[i, j]]), float(sc[i, j]).hex()] for j in range(cnt[i])]
                assert got == r["hits"], (case["seed"], r["query"])


def test_bm25_large_synthetic_matches_c_oracle(fresh_index):
    from rag_search_engine_b200 import synth
    bm = synth.synth_bm25(30_000, 5_000, seed=5, mean_len=40, sd_len=15)       # spans 4 doc ranges
    tok_indptr, terms = synth.synth_token_queries(bm, 300, seed=9)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    for k, k1, b in [(10, 1.5, 0.75), (1, 1.5, 0.75), (64, 2.0, 0.3)]:
        sc, dc, cnt = fresh_index.bm25(tok_indptr, terms, k, k1, b)
        osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl,
                                           tok_indptr, terms, k, k1, b)
        assert cnt.tolist() == ocnt.tolist()
        assert (dc == odc).all()
        assert (sc.view(np.uint64) == osc.view(np.uint64)).all()


@pytest.mark.parametrize("mode", [0, 1])
def test_bm25_streaming_kernel_fallbacks_and_ties(fresh_index, mode):
    """The streaming BM25 kernel hands a query to the general kernel when it has more than 16 tokens or when
    a mass tie overflows its candidate list; small tie groups are resolved inside it (first query token whose
    postings hold the document, then doc).  Everything must equal the C oracle bit for bit."""
    rng = np.random.default_rng(11)
    n_docs, vocab = 40_000, 400
    # documents: 12 000 byte-identical ones (mass tie), groups of 3 identical ones (small ties), the rest random
    docs = []
    for d in range(n_docs):
        if d % 3 == 0 and d < 36_000:
            toks = [1, 2, 3, 3]
        elif d % 3 == 1 and d < 36_000:
            base = (d // 30) % 50
            toks = [10 + base, 11 + base, 200 + (d // 30) % 7]
        else:
            toks = rng.integers(0, vocab, rng.integers(3, 12)).tolist()
        docs.append(toks)
    indptr = [0]; doc_idx = []; tf = []
    by_term = [dict() for _ in range(vocab)]
    for d, toks in enumerate(docs):
        for t in toks:
            by_term[t][d] = by_term[t].get(d, 0) + 1
    for t in range(vocab):
        for d in sorted(by_term[t]):
            doc_idx.append(d); tf.append(by_term[t][d])
        indptr.append(len(doc_idx))
    indptr = np.array(indptr, np.int64); doc_idx = np.array(doc_idx, np.uint32); tf = np.array(tf, np.uint32)
    df = np.diff(indptr).astype(np.int64)
    dl = np.array([len(t) + 1 for t in docs], np.uint32)
    avgdl = float(dl.sum()) / n_docs
    queries = [[1, 3], [3, 2, 1, 3], [10, 11, 200], [12, 13, 201, 12], list(range(5, 25)), [7], [399, 1, 50],
               [203, 30, 31], list(range(100, 117)), [2]]
    queries += [rng.integers(0, vocab, rng.integers(1, 7)).tolist() for _ in range(40)]
    tok_indptr = np.cumsum([0] + [len(q) for q in queries]).astype(np.int32)
    terms = np.array([t for q in queries for t in q], np.int32)
    fresh_index.set_bm25_mode(mode)
    fresh_index.load_bm25(indptr, doc_idx, tf, df, dl, n_docs, avgdl)
    for k in (10, 3, 32):
        sc, dc, cnt = fresh_index.bm25(tok_indptr, terms, k)
        osc, odc, ocnt = oracle.bm25_batch(indptr, doc_idx, tf, df, dl, n_docs, avgdl, tok_indptr, terms, k)
        assert cnt.tolist() == ocnt.tolist()
        assert (dc == odc).all()
        assert (sc.view(np.uint64) == osc.view(np.uint64)).all()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bm25_modes_agree_with_oracle(fresh_index, mode):
    """rse_set_bm25_mode: fixed-point streaming (0), exact-order streaming (1) and the general kernel (2) must all
    return the oracle's documents, order and scores bit for bit."""
    from rag_search_engine_b200 import synth
    bm = synth.synth_bm25(60_000, 8_000, seed=21, mean_len=50, sd_len=20)       # 9 doc ranges
    tok_indptr, terms = synth.synth_token_queries(bm, 400, seed=22)
    fresh_index.set_bm25_mode(mode)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    # k1 + 1 an exact power of two (1.0, 3.0, 0.0) is the edge of the 8-byte posting's weight scale; b = 0 makes
    # every normk equal; k1 = 0 gives w == k1 + 1 for every posting; b > 1 leaves the fixed-point path
    for k, k1, b in [(10, 1.5, 0.75), (32, 1.2, 0.5), (5, 3.0, 1.0), (10, 1.0, 0.75), (10, 0.0, 0.75), (7, 2.0, 0.0),
                     (10, 1.5, 1.25), (10, 100.0, 0.3)]:
        sc, dc, cnt = fresh_index.bm25(tok_indptr, terms, k, k1, b)
        osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl,
                                           tok_indptr, terms, k, k1, b)
        assert cnt.tolist() == ocnt.tolist()
        assert (dc == odc).all()
        assert (sc.view(np.uint64) == osc.view(np.uint64)).all()


def test_bm25_edge_cases(fresh_index):
    postings, doclen = build_postings([{"id": 5, "title": "a b", "description": "c a"},
                                       {"id": 9, "title": "a", "description": ""}])
    csr = to_csr(postings, doclen)
    fresh_index.load_bm25(csr["indptr"], csr["doc"], csr["tf"], csr["df"], csr["dl"], 2, csr["avgdl"])
    a = csr["row"]["a"]
    tok_indptr = np.array([0, 0, 1, 3, 4], np.int32)             # empty query, [a], [oov, a], [oov]
    terms = np.array([a, -1, a, -1], np.int32)
    sc, dc, cnt = fresh_index.bm25(tok_indptr, terms, 10)
    assert cnt.tolist() == [0, 2, 2, 0]
    want = pyref.bm25_search(postings, doclen, 2, ["a"], 10)
    assert [(int(csr["doc_ids"][dc[1, j]]), sc[1, j]) for j in range(2)] == want
    assert [(int(csr["doc_ids"][dc[2, j]]), sc[2, j]) for j in range(2)] == want
    with pytest.raises(Exception):
        fresh_index.bm25(np.array([0, 300], np.int32), np.zeros(300, np.int32), 10)   # > 255 tokens
    with pytest.raises(Exception):
        fresh_index.bm25(tok_indptr, terms, 1000)                                      # k above the limit


# ----------------------------------------------------------------------------- fusion
def _run_fusion_cases(idx, cases):
    by_limit = {}
    for c in cases:
        by_limit.setdefault((c["limit"], c["alpha"], str(c["k"])), []).append(c)
    for (limit, alpha_hex, _), cs in by_limit.items():
        nq = len(cs)
        bid = np.full((nq, limit), -1, np.int64); bsc = np.zeros((nq, limit)); bc = np.zeros(nq, np.int32)
        sid = np.full((nq, limit), -1, np.int64); sds = np.zeros((nq, limit), np.float64); scn = np.zeros(nq, np.int32)
        for i, c in enumerate(cs):
            for j, (d, s) in enumerate(c["bm25"]):
                bid[i, j] = d; bsc[i, j] = float.fromhex(s)
            for j, (d, s) in enumerate(c["sem"]):
                sid[i, j] = d; sds[i, j] = float.fromhex(s)
            bc[i] = len(c["bm25"]); scn[i] = len(c["sem"])
        alpha = float.fromhex(alpha_hex)
        oid, ob, osem, osc, oc = idx.fuse_weighted(limit, alpha, bid, bsc, bc, sid, sds, scn)
        rid, rsc, rb, rs, rc = idx.fuse_rrf(limit, cs[0]["k"], bid, bsc, bc, sid, sds, scn)
        for i, c in enumerate(cs):
            got = [[int(oid[i, j]), float(ob[i, j]).hex(), float(osem[i, j]).hex(), float(osc[i, j]).hex()]
                   for j in range(oc[i])]
            assert got == c["weighted"]
            got = [[int(rid[i, j]), float(rsc[i, j]).hex(), None if rb[i, j] < 0 else int(rb[i, j]),
                    None if rs[i, j] < 0 else int(rs[i, j])] for j in range(rc[i])]
            assert got == c["rrf"]


def test_fusion_matches_reference_golden(fresh_index):
    cases = json.loads((GOLDEN / "fusion_ref.json").read_text())["cases"]
    _run_fusion_cases(fresh_index, cases)


def test_fusion_random_against_pyref_incl_max_limit(fresh_index):
    import random
    rnd = random.Random(3)
    cases = []
    for _ in range(120):
        limit = rnd.choice([1, 7, 10, 33, 100, 128])
        space = rnd.choice([4 * limit + 5, 10**6, 2**45])
        nb, ns = rnd.randint(0, limit), rnd.randint(0, limit)
        bids = rnd.sample(range(1, space), nb)
        sids = rnd.sample(range(1, space), ns)
        for j in range(min(nb, ns)):
            if rnd.random() < 0.4 and bids[j] not in sids:
                sids[j] = bids[j]
        bs = sorted((float(rnd.randint(1, 5)) if rnd.random() < 0.5 else rnd.uniform(0, 9) for _ in range(nb)), reverse=True)
        ds = sorted(float(np.float32(rnd.randint(1, 5) / 8 if rnd.random() < 0.5 else rnd.uniform(0, 1.5))) for _ in range(ns))
        alpha, k = rnd.choice([0.0, 0.5, 1.0, 0.3]), rnd.choice([60, 1, 7.5])
        bm, sem = list(zip(bids, bs)), list(zip(sids, ds))
        w = pyref.weighted_fuse(bm, sem, alpha, limit)
        r = pyref.rrf_fuse(bm, sem, k, limit)
        cases.append({"bm25": [[i, s.hex()] for i, s in bm], "sem": [[i, d.hex()] for i, d in sem], "limit": limit,
                      "alpha": float(alpha).hex(), "k": k,
                      "weighted": [[x["id"], x["bm25"].hex(), x["semantic"].hex(), x["score"].hex()] for x in w],
                      "rrf": [[x["id"], float(x["score"]).hex(), x["bm25_rank"], x["sem_rank"]] for x in r]})
    _run_fusion_cases(fresh_index, cases)


def test_fusion_tie_by_id_mode(fresh_index):
    from rag_search_engine_b200 import _lib
    bid = np.array([[912345, 17]], np.int64); bsc = np.array([[9.5, 7.25]]); bc = np.array([2], np.int32)
    sid = np.array([[5, 700000]], np.int64); sds = np.array([[0.2, 0.3]], np.float64); scn = np.array([2], np.int32)
    rid, rsc, rb, rs, rc = fresh_index.fuse_rrf(2, 60, bid, bsc, bc, sid, sds, scn, tie_mode=_lib.TIE_BY_ID)
    assert rid[0].tolist() == [5, 912345]            # rank-0 tie resolved by ascending id
    rid, *_ = fresh_index.fuse_rrf(2, 60, bid, bsc, bc, sid, sds, scn, tie_mode=_lib.TIE_REFERENCE)
    assert rid[0].tolist() == [x["id"] for x in pyref.rrf_fuse([(912345, 9.5), (17, 7.25)], [(5, 0.2), (700000, 0.3)], 60, 2)]


# ----------------------------------------------------------------------------- hybrid end to end
def test_hybrid_end_to_end_matches_oracle_pipeline(fresh_index):
    from rag_search_engine_b200 import synth
    n_movies = 6000
    se = synth.synth_embeddings(n_movies, seed=4, device="cpu")
    emb = se.emb.numpy(); movie_of = se.movie_of_chunk.numpy()
    bm = synth.synth_bm25(n_movies, 3000, seed=4, mean_len=30, sd_len=10)
    nq = 40
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=8)
    Q = synth.synth_query_vectors(se.emb, nq, seed=8).numpy()
    ids = se.movie_ids
    fresh_index.load_embeddings(emb, movie_idx=movie_of)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    fresh_index.set_id_tables(ids, ids)
    limit = 10
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tok_indptr,
                                       terms, limit)
    kd, krow, kc = oracle.knn_movies_batch(emb, Q, movie_of, limit, limit * 10)
    oid, sc, a, b, cnt = fresh_index.hybrid(0, 60.0, limit, Q, tok_indptr, terms)
    wid, wsc, wa, wb, wcnt = fresh_index.hybrid(1, 0.5, limit, Q, tok_indptr, terms)
    for qi in range(nq):
        bmh = [(int(ids[odc[qi, j]]), float(osc[qi, j])) for j in range(ocnt[qi])]
        semh = [(int(ids[movie_of[krow[qi, j]]]), float(kd[qi, j])) for j in range(kc[qi])]
        want = pyref.rrf_fuse(bmh, semh, 60.0, limit)
        got = [(int(oid[qi, j]), float(sc[qi, j]), None if a[qi, j] < 0 else int(a[qi, j]),
                None if b[qi, j] < 0 else int(b[qi, j])) for j in range(cnt[qi])]
        assert got == [(x["id"], x["score"], x["bm25_rank"], x["sem_rank"]) for x in want]
        wwant = pyref.weighted_fuse(bmh, semh, 0.5, limit)
        wgot = [(int(wid[qi, j]), float(wa[qi, j]), float(wb[qi, j]), float(wsc[qi, j])) for j in range(wcnt[qi])]
        assert wgot == [(x["id"], x["bm25"], x["semantic"], x["score"]) for x in wwant]


# ----------------------------------------------------------------------------- multi-shard merge (1 GPU)
def test_sharded_merge_equals_single_index():
    """Row shards aligned to vec0 blocks + candidate merge == the unsharded KNN,
    emulated on one GPU with one handle per shard."""
    import torch
    from rag_search_engine_b200 import _lib
    rng = np.random.default_rng(31)
    n, n_sh, kp, k = 9000, 3, 100, 10
    movie_of = np.repeat(np.arange(n // 5, dtype=np.int32), 5)[:n]
    emb = inject_ties(rng, unit_rows(rng, n, 384), 300)
    Q = unit_rows(rng, 6, 384); Q[0] = emb[7]
    bounds = [0, 3 * 1024, 6 * 1024, n]
    qd = torch.as_tensor(Q, device="cuda")
    gathered = torch.empty((n_sh, len(Q), kp, 3), dtype=torch.int64, device="cuda")
    handles = []
    for s in range(n_sh):
        h = _lib.Index(0)
        h.load_embeddings(emb[bounds[s]:bounds[s + 1]], movie_idx=movie_of[bounds[s]:bounds[s + 1]], pos_base=bounds[s])
        h.knn_local_dev(qd.data_ptr(), len(Q), kp, gathered[s].data_ptr())
        h.synchronize()
        handles.append(h)
    od = torch.empty((len(Q), k), dtype=torch.float32, device="cuda")
    orow = torch.empty((len(Q), k), dtype=torch.int64, device="cuda")
    om = torch.empty((len(Q), k), dtype=torch.int32, device="cuda")
    oc = torch.empty((len(Q),), dtype=torch.int32, device="cuda")
    handles[0].knn_merge_movies_dev(gathered.data_ptr(), n_sh, len(Q), k, kp, od.data_ptr(), orow.data_ptr(),
                                    om.data_ptr(), oc.data_ptr())
    handles[0].synchronize()
    wd, wrow, wc = oracle.knn_movies_batch(emb, Q, movie_of, k, kp)
    assert oc.cpu().numpy().tolist() == wc.tolist()
    assert (orow.cpu().numpy() == wrow).all()
    assert (od.cpu().numpy().view(np.uint32) == wd.view(np.uint32)).all()
    for h in handles:
        h.close()


# ----------------------------------------------------------------------------- K4 tensor-core path
@pytest.mark.parametrize("n,nq,kprime", [(40_000, 70, 100), (9_000, 256, 50), (33_000, 300, 100), (5_000, 5, 10),
                                         (60_000, 64, 400)])      # K' > 256: the dense probe instead of the per-thread top-8
def test_tensor_core_path_equals_exact_scan_and_oracle(n, nq, kprime):
    """K4 (tcgen05 probe/filter + exact re-score) must return exactly what the exact scan returns:
    same rows, same order, bit-identical distances — including duplicate rows (ties) and non-unit norms.
    Mode 1 = exact scan only, 2 = the tensor-core path forced on."""
    from rag_search_engine_b200 import _lib
    rng = np.random.default_rng(n + nq)
    centers = unit_rows(rng, 50, 384)
    emb = centers[rng.integers(0, 50, n)] + 0.5 * unit_rows(rng, n, 384)          # clustered → dense near-neighbour bands
    emb *= rng.uniform(0.7, 1.4, (n, 1)).astype(np.float32)                        # non-unit norms
    emb = inject_ties(rng, emb.astype(np.float32), n // 25)
    Q = (centers[rng.integers(0, 50, nq)] + 0.3 * unit_rows(rng, nq, 384)).astype(np.float32)
    Q[0] = emb[7]
    movie_of = (np.arange(n) // 6).astype(np.int32)
    res = {}
    for mode in (1, 2):
        idx = _lib.Index(0)
        try:
            idx.set_tc_mode(mode)
            idx.load_embeddings(emb, movie_idx=movie_of)
            res[mode] = idx.knn(Q, kprime) + idx.knn_movies(Q, 10, kprime)
            st = idx.stats()
            if mode != 1:
                assert st.tc_queries == 2 * nq and st.tc_filter_launches >= 2, "tensor-core path did not run"
                assert st.tc_fallback_queries <= nq // 4
            else:
                assert st.tc_queries == 0
        finally:
            idx.close()
    for mode in (2,):
        for a, b in zip(res[1], res[mode]):
            assert a.dtype == b.dtype and (a.view(np.uint8) == b.view(np.uint8)).all(), f"tc_mode {mode}"
    dist, pos = res[2][0], res[2][1]
    for qi in (0, nq // 2, nq - 1):
        od, orow = oracle.vec0_knn(emb, Q[qi], kprime, literal=False)
        assert pos[qi].tolist() == orow.tolist()
        assert dist[qi].view(np.uint32).tolist() == od.view(np.uint32).tolist()


def test_tensor_core_path_overflow_falls_back_to_exact():
    """Thousands of identical rows tie at the threshold → the survivor list overflows → the query is
    re-run through the exact scan and still matches."""
    from rag_search_engine_b200 import _lib
    rng = np.random.default_rng(77)
    n = 30_000
    emb = unit_rows(rng, n, 384)
    emb[5000:17000] = emb[5000]                                   # 12 000 copies of one vector
    Q = unit_rows(rng, 64, 384)
    Q[3] = emb[5000] + 0.01 * unit_rows(rng, 1, 384)[0]
    out = {}
    for mode in (1, 2):
        idx = _lib.Index(0)
        try:
            idx.set_tc_mode(mode)
            idx.load_embeddings(emb)
            out[mode] = idx.knn(Q, 100)
            if mode != 1:
                assert idx.stats().tc_fallback_queries >= 1
        finally:
            idx.close()
    for mode in (2,):
        for a, b in zip(out[1], out[mode]):
            assert (a.view(np.uint8) == b.view(np.uint8)).all()
    assert (out[2][1][3] >= 5000).all() and (out[2][1][3] < 17000).all()


@pytest.mark.parametrize("zero_rows", [False, True])
def test_tensor_core_path_degenerate_rows_and_queries(zero_rows):
    """Zero-norm query, empty vec0 slots past the end of a tile, tiny / huge norms and a wide dynamic range
    inside rows: the fp16-shadow filter must never lose a row of the exact top-K' (the zero-norm query is
    handed to the exact scan).  With zero-norm ROWS in the corpus the exact distance is NaN, which the
    filter bound cannot cover: the index must stay on the exact scan altogether."""
    from rag_search_engine_b200 import _lib
    rng = np.random.default_rng(5)
    n = 20_000
    emb = unit_rows(rng, n, 384)
    emb[::7] *= np.float32(1e-3)                                  # tiny norms
    emb[1::11] *= np.float32(3e4)                                 # huge norms (fp16 could not hold these unscaled)
    emb[2::13, :200] *= np.float32(1e-6)                          # components far below the fp16 normal range
    if zero_rows:
        emb[500:520] = 0.0
    Q = unit_rows(rng, 64, 384)
    Q[1] = 0.0                                                    # zero-norm query
    Q[2] = emb[1] * np.float32(1e-4)
    Q[3] = emb[2] + np.float32(1e-3) * unit_rows(rng, 1, 384)[0]
    out = {}
    for mode in (1, 2):
        idx = _lib.Index(0)
        try:
            idx.set_tc_mode(mode)
            idx.load_embeddings(emb)
            out[mode] = idx.knn(Q, 100)
            if mode == 2:
                st = idx.stats()
                if zero_rows:
                    assert st.tc_queries == 0 and st.tc_filter_launches == 0
                else:
                    assert st.tc_queries == 64 and 1 <= st.tc_fallback_queries <= 8
        finally:
            idx.close()
    for a, b in zip(out[1], out[2]):
        assert (a.view(np.uint8) == b.view(np.uint8)).all()


def test_hybrid_large_batch_takes_tensor_core_path_and_matches_oracle(fresh_index):
    """≥ 262 144 rows and ≥ 48 queries → the hybrid call routes its KNN through K4 automatically; the
    fused lists must still equal the oracle pipeline bit for bit."""
    from rag_search_engine_b200 import synth
    n_movies = 36_000
    se = synth.synth_embeddings(n_movies, seed=6, device="cuda")
    assert se.emb.shape[0] >= 262_144
    bm = synth.synth_bm25(n_movies, 20_000, seed=6, mean_len=40, sd_len=12, device="cuda")
    nq, limit = 64, 10
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=12)
    Q = synth.synth_query_vectors(se.emb, nq, seed=12).cpu().numpy()
    ids = se.movie_ids
    fresh_index.attach_embeddings_dev(se.emb.data_ptr(), se.emb.shape[0], 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(),
                                      keepalive=se)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    fresh_index.set_id_tables(ids, ids)
    oid, sc, a, b, cnt = fresh_index.hybrid(0, 60.0, limit, Q, tok_indptr, terms)
    st = fresh_index.stats()
    assert st.tc_queries == nq and st.tc_filter_launches == 1
    emb = se.emb.cpu().numpy(); movie_of = se.movie_of_chunk.cpu().numpy()
    sel = [0, 7, 31, 63]
    kd, krow, kc = oracle.knn_movies_batch(emb, Q[sel], movie_of, limit, limit * 10, literal=False)
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tok_indptr,
                                       terms, limit)
    for i, qi in enumerate(sel):
        bmh = [(int(ids[odc[qi, j]]), float(osc[qi, j])) for j in range(ocnt[qi])]
        semh = [(int(ids[movie_of[krow[i, j]]]), float(kd[i, j])) for j in range(kc[i])]
        want = pyref.rrf_fuse(bmh, semh, 60.0, limit)
        got = [(int(oid[qi, j]), float(sc[qi, j])) for j in range(cnt[qi])]
        assert got == [(x["id"], x["score"]) for x in want]


# ----------------------------------------------------------------------------- full-size properties (S-600k shape)
def test_full_size_tensor_core_path_equals_exact_scan_and_is_idempotent():
    """At the benchmark's own size (4.8 M chunks x 384): the batched tensor-core path must return exactly what the
    exact streaming scan returns for the same 256 queries (rows, order, distance bits), twice in a row (idempotence),
    and every list must be sorted by the vec0 emit order (distance asc, block asc, slot desc)."""
    import torch
    from rag_search_engine_b200 import _lib, synth
    se = synth.synth_embeddings(600_000, seed=1234, device="cuda")
    n = int(se.emb.shape[0])
    Q = synth.synth_query_vectors(se.emb, 256, seed=99).cpu().numpy()
    out = {}
    for mode in (1, 0):
        idx = _lib.Index(0)
        try:
            idx.set_tc_mode(mode)
            idx.attach_embeddings_dev(se.emb.data_ptr(), n, 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(), keepalive=se)
            out[mode] = idx.knn(Q, 100)
            if mode == 0:
                again = idx.knn(Q, 100)
                for a, b in zip(out[0], again):
                    assert (a.view(np.uint8) == b.view(np.uint8)).all(), "second run differs from the first"
                st = idx.stats()
                assert st.tc_queries == 512 and st.tc_fallback_queries == 0
        finally:
            idx.close()
    for a, b in zip(out[1], out[0]):
        assert a.dtype == b.dtype and (a.view(np.uint8) == b.view(np.uint8)).all()
    dist, pos = out[0][0], out[0][1]
    assert (out[0][-1] == 100).all()                                  # every query found its full K'
    d = dist                                                          # f32 distances (may be a few ulps below 0)
    blk, slot = pos // 1024, pos % 1024
    ok = (d[:, 1:] > d[:, :-1]) | ((d[:, 1:] == d[:, :-1]) & ((blk[:, 1:] > blk[:, :-1]) |
                                                              ((blk[:, 1:] == blk[:, :-1]) & (slot[:, 1:] < slot[:, :-1]))))
    assert ok.all(), "a result list is not in vec0 emit order"
    del se
    torch.cuda.empty_cache()


def test_full_size_bm25_modes_agree():
    """At the benchmark's own size (600 k documents, 54 M postings, 256 queries touching ~78 M postings): the
    fixed-point streaming path, the exact-order streaming path and the general kernel return identical documents,
    order and score bits; scores are non-increasing and every hit is a real document."""
    import torch
    from rag_search_engine_b200 import _lib, synth
    bm = synth.synth_bm25(600_000, 1_000_000, seed=1234, device="cuda")
    tok_indptr, terms = synth.synth_token_queries(bm, 256, seed=99)
    out = {}
    for mode in (2, 1, 0):
        idx = _lib.Index(0)
        try:
            idx.set_bm25_mode(mode)
            idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
            out[mode] = idx.bm25(tok_indptr, terms, 10)
        finally:
            idx.close()
    for mode in (1, 0):
        for a, b in zip(out[2], out[mode]):
            assert (a.view(np.uint8) == b.view(np.uint8)).all(), f"bm25 mode {mode} differs from the general kernel"
    sc, dc, cnt = out[0]
    for q in range(256):
        c = int(cnt[q])
        assert (np.diff(sc[q, :c]) <= 0).all()
        assert ((dc[q, :c] >= 0) & (dc[q, :c] < 600_000)).all()
    torch.cuda.empty_cache()


def test_hybrid_pipelined_submit_collect_equals_the_blocking_call(fresh_index):
    """rse_hybrid_submit / rse_hybrid_collect (two batches in flight): every batch returns what rse_hybrid returns on
    the same inputs, whatever is in flight behind it; the third submit and an out-of-order collect are refused."""
    from rag_search_engine_b200 import synth
    n_movies = 36_000
    se = synth.synth_embeddings(n_movies, seed=31, device="cuda")
    bm = synth.synth_bm25(n_movies, 20_000, seed=32, mean_len=40, sd_len=12, device="cuda")
    fresh_index.attach_embeddings_dev(se.emb.data_ptr(), se.emb.shape[0], 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(),
                                      keepalive=se)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    fresh_index.set_id_tables(se.movie_ids, se.movie_ids)
    batches = []
    for i, nq in enumerate([96, 300, 17, 256, 64]):            # sizes on both sides of the tensor-core threshold
        Q = synth.synth_query_vectors(se.emb, nq, seed=40 + i).cpu().numpy()
        tp, tr = synth.synth_token_queries(bm, nq, seed=50 + i)
        batches.append((Q, tp, tr, 10 if i % 2 == 0 else 5, i % 2))
    want = [fresh_index.hybrid(mode, 60.0 if mode == 0 else 0.5, limit, Q, tp, tr) for Q, tp, tr, limit, mode in batches]
    got, tickets = [], []
    for Q, tp, tr, limit, mode in batches:
        Qc, tpc, trc = Q.copy(), tp.copy(), tr.copy()
        tickets.append(fresh_index.hybrid_submit(mode, 60.0 if mode == 0 else 0.5, limit, Qc, tpc, trc))
        Qc[:] = 0; tpc[:] = 0; trc[:] = 0                        # caller buffers are free once submit returns
        if len(tickets) == 2:
            with pytest.raises(Exception):
                fresh_index.hybrid_submit(mode, 60.0, limit, Q, tp, tr)       # a third batch in flight
            with pytest.raises(Exception):
                fresh_index.hybrid_collect(tickets[1])                          # out of order
            got.append(fresh_index.hybrid_collect(tickets.pop(0)))
    while tickets:
        got.append(fresh_index.hybrid_collect(tickets.pop(0)))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (g[0] == w[0]).all() and (g[4] == w[4]).all()
        for a, b in zip(g[1:4], w[1:4]):
            assert (a.view(np.uint64) == b.view(np.uint64)).all()
    # the blocking call still works afterwards
    Q, tp, tr, limit, mode = batches[0]
    again = fresh_index.hybrid(mode, 60.0, limit, Q, tp, tr)
    assert (again[0] == want[0][0]).all()


def test_hybrid_pipelined_overflowed_batch_is_rerun_at_collect():
    """submit does not wait for the tensor-core path's overflow flags; collect finds a flagged query (12 000 identical
    rows tie at the threshold) and re-runs that batch through the blocking call — same bits as the exact path, and
    the batch in flight behind it is not disturbed."""
    from rag_search_engine_b200 import _lib, synth
    rng = np.random.default_rng(78)
    n = 30_000
    emb = unit_rows(rng, n, 384)
    emb[5000:17000] = emb[5000]
    movie_of = (np.arange(n) // 3).astype(np.int32)
    ids = (np.arange(n // 3 + 1, dtype=np.int64) * 7 + 3)
    bm = synth.synth_bm25(len(ids), 4_000, seed=3, mean_len=30, sd_len=10)
    Qa = unit_rows(rng, 64, 384)
    Qa[3] = emb[5000] + 0.01 * unit_rows(rng, 1, 384)[0]         # overflows
    Qb = unit_rows(rng, 64, 384)                                   # does not
    tpa, tra = synth.synth_token_queries(bm, 64, seed=4)
    tpb, trb = synth.synth_token_queries(bm, 64, seed=5)
    out = {}
    for mode in (1, 2):
        idx = _lib.Index(0)
        try:
            idx.set_tc_mode(mode)
            idx.load_embeddings(emb, movie_idx=movie_of)
            idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
            idx.set_id_tables(ids, ids)
            if mode == 1:
                out[1] = (idx.hybrid(0, 60.0, 10, Qa, tpa, tra), idx.hybrid(0, 60.0, 10, Qb, tpb, trb))
            else:
                ta = idx.hybrid_submit(0, 60.0, 10, Qa, tpa, tra)
                tb = idx.hybrid_submit(0, 60.0, 10, Qb, tpb, trb)
                assert idx.stats().tc_fallback_queries == 0          # nothing has been checked yet
                ra = idx.hybrid_collect(ta)
                assert idx.stats().tc_fallback_queries >= 1
                rb = idx.hybrid_collect(tb)
                out[2] = (ra, rb)
        finally:
            idx.close()
    for want, got in zip(out[1], out[2]):
        for a, b in zip(want, got):
            assert (a.view(np.uint8) == b.view(np.uint8)).all()



def test_hybrid_repeated_large_batches_are_stable(fresh_index):
    """Back-to-back hybrid batches of 1024 queries (four tensor-core blocks, BM25 on the second stream underneath the
    filter pass) must return the same bits every time — a shared-memory race in the BM25 kernel once made this fail
    intermittently with an illegal access."""
    from rag_search_engine_b200 import synth
    n_movies = 36_000
    se = synth.synth_embeddings(n_movies, seed=6, device="cuda")
    assert se.emb.shape[0] >= 262_144
    bm = synth.synth_bm25(n_movies, 20_000, seed=6, mean_len=40, sd_len=12, device="cuda")
    nq, limit = 1024, 10
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=13)
    Q = synth.synth_query_vectors(se.emb, nq, seed=13).cpu().numpy()
    ids = se.movie_ids
    fresh_index.attach_embeddings_dev(se.emb.data_ptr(), se.emb.shape[0], 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(),
                                      keepalive=se)
    fresh_index.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    fresh_index.set_id_tables(ids, ids)
    ref = fresh_index.hybrid(0, 60.0, limit, Q, tok_indptr, terms)
    assert fresh_index.stats().tc_queries == nq
    for it in range(25):
        got = fresh_index.hybrid(0, 60.0, limit, Q, tok_indptr, terms)
        for a, b in zip(ref, got):
            assert (a.view(np.uint8) == b.view(np.uint8)).all(), f"iteration {it} differs"
    fresh_index.hybrid_stage(Q, tok_indptr, terms)
    for it in range(25):
        fresh_index.hybrid_run(0, 60.0, limit)
    got = fresh_index.hybrid_fetch(limit)
    for a, b in zip(ref, got):
        assert (a.view(np.uint8) == b.view(np.uint8)).all()
