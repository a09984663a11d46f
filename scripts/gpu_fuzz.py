"""Randomised parity fuzzer for the CUDA path (run on the GPU box: `python scripts/gpu_fuzz.py --cases 150`).

Every case draws a configuration at random — corpus size and shape, row-norm spread, cluster structure, duplicate
density, holes, rowid permutation, position base, K', batch size, BM25 parameters, fusion mode — runs it through
the C-ABI (ctypes -> librse.so) and compares:

  knn     the tensor-core path (K4) byte for byte against the exact streaming scan (K1+K2), and both against the
          oracle's vec0 restatement on a few queries of the case;
  bm25    the three BM25 modes against oracle.bm25_batch (bit-equal fp64 scores, documents, order);
  hybrid  rse_hybrid (rrf / weighted) against oracle BM25 + oracle KNN-per-movie + the Python restatement of the
          reference's fusion.

Test infrastructure: it imports oracle/ as the checker, like tests/.  tests/test_gpu_fuzz.py runs a small fixed
slice of the same seeds in the `-m gpu` suite; a longer run's log is kept under profiles/.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
if str(ROOT / "tests") not in sys.path:
    sys.path.insert(0, str(ROOT / "tests"))


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-30)
    return x


def make_corpus(rng, n, dim):
    """Rows with a randomly chosen structure; returns (emb fp32[n, dim], description)."""
    kind = rng.choice(["iso", "clustered", "tight", "scaled", "lowrank"])
    if kind == "iso":
        emb = _unit(rng, n, dim)
    elif kind in ("clustered", "tight"):
        n_c = int(rng.integers(2, 80))
        spread = float(rng.choice([0.02, 0.1, 0.3, 0.6])) if kind == "tight" else float(rng.uniform(0.3, 1.0))
        centers = _unit(rng, n_c, dim)
        emb = centers[rng.integers(0, n_c, n)] + spread * _unit(rng, n, dim)
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
        kind = f"{kind}({n_c},{spread:.2f})"
    elif kind == "scaled":
        emb = _unit(rng, n, dim) * rng.uniform(1e-3, 1e3, (n, 1)).astype(np.float32)
    else:
        r = int(rng.integers(1, 6))
        emb = (rng.standard_normal((n, r)) @ rng.standard_normal((r, dim))).astype(np.float32)
        emb += 0.05 * rng.standard_normal((n, dim)).astype(np.float32)
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    n_dup = int(n * rng.choice([0.0, 0.02, 0.2]))
    if n_dup:
        emb[rng.integers(0, n, n_dup)] = emb[rng.integers(0, n, n_dup)]
    for s, d in [(3, 1023), (3, 1024), (5, 2047), (5, 2048)]:       # duplicates across vec0 block boundaries
        if d < n:
            emb[d] = emb[s]
    return emb, kind


def make_queries(rng, emb, nq):
    n, dim = emb.shape
    how = rng.choice(["random", "near", "mixed"])
    Q = _unit(rng, nq, dim)
    if how != "random":
        pick = rng.integers(0, n, nq)
        near = emb[pick] + float(rng.choice([0.0, 0.01, 0.2])) * np.abs(emb[pick]).mean() * rng.standard_normal((nq, dim)).astype(np.float32)
        m = np.ones(nq, bool) if how == "near" else rng.random(nq) < 0.5
        Q[m] = near[m]
    return np.ascontiguousarray(Q, dtype=np.float32)


def fuzz_knn(seed, lib, oracle):
    rng = np.random.default_rng(1000 + seed)
    dim = 384 if rng.random() < 0.85 else int(rng.choice([4, 48, 100, 256, 512, 768]))
    n = int(rng.choice([rng.integers(1, 300), rng.integers(300, 5000), rng.integers(5000, 60000), rng.integers(60000, 250000)],
                       p=[0.1, 0.3, 0.45, 0.15]))
    nq = int(rng.choice([1, 2, 3, rng.integers(4, 64), rng.integers(64, 257), rng.integers(257, 700)]))
    kprime = int(rng.choice([1, 10, 50, 100, 100, 100, rng.integers(1, 400), rng.integers(400, 4097)]))
    if n * nq > 40_000_000:
        nq = max(1, 40_000_000 // n)
    emb, kind = make_corpus(rng, n, dim)
    Q = make_queries(rng, emb, nq)
    holes = rng.random() < 0.3
    valid = (rng.random(n) > 0.15).astype(np.uint8) if holes else None
    if valid is not None and valid.sum() == 0:
        valid[0] = 1
    rowid = (rng.permutation(n).astype(np.int64) + int(rng.integers(1, 10**9))) if rng.random() < 0.3 else None
    base = int(rng.integers(0, 50)) * 1024 if rng.random() < 0.3 else 0
    movie_of = np.sort(rng.integers(0, max(1, n // int(rng.integers(1, 12))), n)).astype(np.int32)
    fma = bool(rng.random() < 0.15)
    k = int(rng.integers(1, min(kprime, 128) + 1))
    desc = f"knn seed={seed} n={n} dim={dim} nq={nq} k'={kprime} k={k} {kind} holes={holes} rowid={rowid is not None} base={base} fma={fma}"
    # the probe's row sample (knn_tc3.cuh): as shipped it needs >= 512 k rows — lower the bar on some cases, and on
    # some force the probe's order statistic so low that the verified bound fails and the second chance must repair it
    knobs = {}
    pick = rng.random()
    if pick < 0.5:
        knobs = {"RSE_TC_SAMPLE_MIN_TILES": "1", "RSE_TC_SAMPLE_STRIDE": str(int(rng.choice([4, 8, 16, 32])))}
        if pick < 0.2:
            knobs["RSE_TC_PROBE_RANK"] = str(int(rng.choice([1, 2, 5])))
    desc += f" knobs={knobs}" if knobs else ""
    res = {}
    used_tc = False
    for mode in (1, 2):
        old = {name: os.environ.get(name) for name in knobs}
        if mode == 2:
            os.environ.update(knobs)
        try:
            idx = lib.Index(0)
        finally:
            for name, val in old.items():
                if val is None:
                    os.environ.pop(name, None)
                else:
                    os.environ[name] = val
        try:
            idx.set_tc_mode(mode)
            idx.set_fma(fma)
            idx.load_embeddings(emb, valid=valid, rowid=rowid, movie_idx=movie_of, pos_base=base)
            res[mode] = idx.knn(Q, kprime) + idx.knn_movies(Q, k, kprime)
            if mode == 2:
                used_tc = idx.stats().tc_queries > 0
        finally:
            idx.close()
    for i, (a, b) in enumerate(zip(res[1], res[2])):
        if not (a.shape == b.shape and (a.view(np.uint8) == b.view(np.uint8)).all()):
            return False, desc + f" :: K4 != exact scan in output {i}", used_tc
    dist, pos, rid, _, cnt = res[1][:5]
    keep = np.nonzero(valid)[0] if valid is not None else np.arange(n)
    for qi in sorted({0, nq // 2, nq - 1}):
        od, orow = oracle.vec0_knn(emb[keep], Q[qi], kprime, pos=(keep + base).astype(np.int64), use_fma=fma,
                                   literal=(len(keep) <= 4000))
        c = int(cnt[qi])
        ok = (c == len(orow) and pos[qi, :c].tolist() == (keep[orow] + base).tolist()
              and dist[qi, :c].view(np.uint32).tolist() == od.view(np.uint32).tolist())
        if ok and rowid is not None:
            ok = rid[qi, :c].tolist() == rowid[keep[orow]].tolist()
        if not ok:
            return False, desc + f" :: exact scan != oracle at query {qi}", used_tc
    return True, desc, used_tc


def fuzz_bm25(seed, lib, oracle):
    from rag_search_engine_b200 import synth
    rng = np.random.default_rng(2000 + seed)
    n_docs = int(rng.choice([rng.integers(1, 200), rng.integers(200, 20000), rng.integers(20000, 150000)]))
    vocab = int(rng.choice([rng.integers(1, 50), rng.integers(50, 5000), rng.integers(5000, 40000)]))
    bm = synth.synth_bm25(n_docs, vocab, seed=seed, mean_len=float(rng.uniform(3, 80)), sd_len=float(rng.uniform(0, 30)))
    nq = int(rng.integers(1, 400))
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=seed + 1, min_tok=int(rng.integers(0, 3)),
                                                  max_tok=int(rng.integers(3, 24)))
    terms = terms.copy()
    if len(terms):
        terms[rng.random(len(terms)) < 0.05] = -1                    # unknown terms are skipped
    k = int(rng.choice([1, 3, 10, 10, 32, 64, 128]))
    k1 = float(rng.choice([1.5, 1.5, 1.2, 2.0, 0.0, 1.0, 3.0, rng.uniform(0, 10)]))
    b = float(rng.choice([0.75, 0.75, 0.0, 1.0, rng.uniform(0, 1), 1.25]))
    desc = f"bm25 seed={seed} docs={n_docs} vocab={vocab} nq={nq} k={k} k1={k1:.3f} b={b:.3f}"
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tok_indptr, terms,
                                       k, k1, b)
    for mode in (0, 1, 2):
        idx = lib.Index(0)
        try:
            idx.set_bm25_mode(mode)
            idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
            sc, dc, cnt = idx.bm25(tok_indptr, terms, k, k1, b)
        finally:
            idx.close()
        if not (cnt.tolist() == ocnt.tolist() and (dc == odc).all() and (sc.view(np.uint64) == osc.view(np.uint64)).all()):
            return False, desc + f" :: mode {mode} != oracle", False
    return True, desc, False


def fuzz_hybrid(seed, lib, oracle):
    from oracle import pyref
    from rag_search_engine_b200 import synth
    rng = np.random.default_rng(3000 + seed)
    n_movies = int(rng.integers(50, 9000))
    dist_kind = str(rng.choice(["isotropic", "clustered"]))
    se = synth.synth_embeddings(n_movies, seed=seed, device="cpu", distribution=dist_kind)
    emb = se.emb.numpy(); movie_of = se.movie_of_chunk.numpy(); ids = se.movie_ids
    bm = synth.synth_bm25(n_movies, int(rng.integers(20, 4000)), seed=seed, mean_len=float(rng.uniform(5, 40)), sd_len=8.0)
    nq = int(rng.choice([1, 2, 7, 40, 130, 300]))
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=seed + 2)
    Q = synth.synth_query_vectors(se.emb, nq, seed=seed + 3).numpy()
    limit = int(rng.choice([1, 5, 10, 10, 25]))
    mult = int(rng.choice([1, 10, 10, 20]))
    mode = int(rng.integers(0, 2))
    param = float(rng.choice([60.0, 1.0, 7.5])) if mode == 0 else float(rng.choice([0.5, 0.0, 1.0, 0.3]))
    desc = f"hybrid seed={seed} movies={n_movies} {dist_kind} nq={nq} limit={limit} mult={mult} mode={'rrf' if mode == 0 else 'weighted'} param={param}"
    # half of the cases force the tensor-core path (the automatic choice needs 262 k rows): BM25 then runs on the second
    # stream underneath the filter, in its 896-thread shape
    force_tc = bool(rng.random() < 0.5)
    desc += f" tc={'forced' if force_tc else 'auto'}"
    knobs = {"RSE_TC_SAMPLE_MIN_TILES": "1", "RSE_TC_SAMPLE_STRIDE": "8"} if force_tc and rng.random() < 0.5 else {}
    old_env = {name: os.environ.get(name) for name in knobs}
    os.environ.update(knobs)
    try:
        idx = lib.Index(0)
    finally:
        for name, val in old_env.items():
            if val is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = val
    try:
        if force_tc:
            idx.set_tc_mode(2)
        idx.load_embeddings(emb, movie_idx=movie_of)
        idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
        idx.set_id_tables(ids, ids)
        oid, sc, a, b, cnt = idx.hybrid(mode, param, limit, Q, tok_indptr, terms, knn_multiplier=mult)
        used_tc = idx.stats().tc_queries > 0
    finally:
        idx.close()
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tok_indptr, terms, limit)
    kd, krow, kc = oracle.knn_movies_batch(emb, Q, movie_of, limit, max(limit * mult, limit))
    for qi in range(nq):
        bmh = [(int(ids[odc[qi, j]]), float(osc[qi, j])) for j in range(ocnt[qi])]
        semh = [(int(ids[movie_of[krow[qi, j]]]), float(kd[qi, j])) for j in range(kc[qi])]
        if mode == 0:
            want = [(x["id"], x["score"], x["bm25_rank"], x["sem_rank"]) for x in pyref.rrf_fuse(bmh, semh, param, limit)]
            got = [(int(oid[qi, j]), float(sc[qi, j]), None if a[qi, j] < 0 else int(a[qi, j]),
                    None if b[qi, j] < 0 else int(b[qi, j])) for j in range(cnt[qi])]
        else:
            want = [(x["id"], x["bm25"], x["semantic"], x["score"]) for x in pyref.weighted_fuse(bmh, semh, param, limit)]
            got = [(int(oid[qi, j]), float(a[qi, j]), float(b[qi, j]), float(sc[qi, j])) for j in range(cnt[qi])]
        if got != want:
            return False, desc + f" :: query {qi}", used_tc
    return True, desc, used_tc


KINDS = {"knn": fuzz_knn, "bm25": fuzz_bm25, "hybrid": fuzz_hybrid}


def run(seeds, kinds=("knn", "bm25", "hybrid"), verbose=False, out=sys.stdout):
    import oracle
    from rag_search_engine_b200 import _lib
    failures, n_tc, n = [], 0, 0
    for seed in seeds:
        for kind in kinds:
            t0 = time.perf_counter()
            ok, desc, used_tc = KINDS[kind](seed, _lib, oracle)
            n += 1
            n_tc += bool(used_tc)
            if verbose or not ok:
                print(("ok   " if ok else "FAIL ") + desc + f"  [{time.perf_counter() - t0:.1f}s{' tc' if used_tc else ''}]", file=out, flush=True)
            if not ok:
                failures.append(desc)
    print(f"{n} cases, {len(failures)} failures, {n_tc} cases went through the tensor-core path", file=out, flush=True)
    return failures


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=50)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--kinds", default="knn,bm25,hybrid")
    args = ap.parse_args()
    bad = run(range(args.first, args.first + args.cases), tuple(args.kinds.split(",")), verbose=True)
    sys.exit(1 if bad else 0)
