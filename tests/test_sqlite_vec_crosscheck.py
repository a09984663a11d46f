"""Pin the vec0 KNN oracle (and the exporter) against the REAL sqlite-vec whenever it is importable.

The reference's KNN arithmetic lives in sqlite-vec (rag_search_engine/utils/semantic_search.py:68-72 load,
:94-101 DDL, :254-261 ``embedding MATCH :q AND k = :k``), an un-vendored, unpinned dependency that is not
installable in the build container, so ``oracle/oracle.c`` restates its published algorithm ("parity
unpinned").  These tests auto-enable the first time ``import sqlite_vec`` works — same pattern
as the reference's own ``tests/test_semantic_search.py:8-9`` — and then check, on the reference's own DDL:

  * ``MATCH … k=`` rows, distances (bit for bit) and TIE ORDER == ``oracle.vec0_knn(literal=True)``, with
    duplicate vectors placed on both sides of 1024-row block boundaries;
  * ``store.export_embeddings`` on the REAL shadow tables returns exactly what was inserted, in vec0's
    physical order (incl. after deletes);
  * (``-m gpu``) ``rse_knn`` on the exported arrays == the real extension.

``bench.py`` prints ``sqlite_vec: present|absent`` in its JSON line so a record states which situation it ran in.
"""
import sqlite3

import numpy as np
import pytest

sqlite_vec = pytest.importorskip("sqlite_vec", reason="sqlite-vec not installed: the vec0 KNN oracle stays unpinned")

import oracle  # noqa: E402
from rag_search_engine_b200 import store  # noqa: E402

DIM = 384


def _unit(rng, n, d=DIM):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True).astype(np.float32)


def _corpus(seed=3, n=3 * 1024 + 137):
    rng = np.random.default_rng(seed)
    emb = _unit(rng, n)
    emb *= rng.uniform(0.8, 1.25, (n, 1)).astype(np.float32)          # the reference recomputes both norms per pair
    # exact duplicates straddling block boundaries: (slot 1023 of block 0, slot 0 of block 1), inside one block,
    # and across blocks 1 and 2 — the cases that decide (distance asc, block asc, slot desc)
    for a, b in ((1023, 1024), (1024, 1030), (2047, 2048), (5, 900), (5, 2500), (2048 + 77, 3 * 1024 + 100)):
        emb[b] = emb[a]
    return emb


def _real_db(tmp_path, emb, delete=()):
    conn = sqlite3.connect(tmp_path / "real.db")
    conn.enable_load_extension(True)
    sqlite_vec.load(conn)                                              # semantic_search.py:68-72
    conn.enable_load_extension(False)
    store._init_schema(conn)
    # the reference's DDL, verbatim (semantic_search.py:94-101)
    conn.execute(f"CREATE VIRTUAL TABLE IF NOT EXISTS chunk_embeddings USING vec0("
                 f"embedding float[{emb.shape[1]}] distance_metric=cosine)")
    conn.executemany("INSERT INTO chunk_embeddings(rowid, embedding) VALUES (?, ?)",      # :203-206
                     [(int(i), emb[i]) for i in range(emb.shape[0])])
    for r in delete:
        conn.execute("DELETE FROM chunk_embeddings WHERE rowid = ?", (int(r),))
    n_movies = emb.shape[0] // 5 + 1
    conn.executemany("INSERT INTO movies(id, title, description) VALUES (?, ?, ?)",
                     [(10 + 3 * m, f"t{m}", "d.") for m in range(n_movies)])
    conn.executemany("INSERT INTO chunks(id, movie_id, chunk_index, max_chunk_size, overlap) VALUES (?,?,?,?,?)",
                     [(i, 10 + 3 * (i // 5), i % 5, 3, 1) for i in range(emb.shape[0])])
    conn.commit()
    return conn


def _match(conn, q, k):
    return conn.execute("SELECT rowid, distance FROM chunk_embeddings WHERE embedding MATCH :q AND k = :k "
                        "ORDER BY distance", {"q": np.ascontiguousarray(q, np.float32), "k": int(k)}).fetchall()


def _queries(emb, rng, nq=24):
    Q = _unit(rng, nq)
    Q[0] = emb[1023]                      # hits the block-boundary duplicate pair at distance ~0
    Q[1] = emb[5]                         # the triple (5, 900, 2500)
    Q[2] = emb[2047] * 3.0                # non-unit query norm
    Q[3] = emb[2048 + 77] + 1e-3 * _unit(rng, 1)[0]
    return Q


def test_match_equals_oracle_rows_distances_and_tie_order(tmp_path):
    emb = _corpus()
    conn = _real_db(tmp_path, emb)
    rng = np.random.default_rng(11)
    for k in (1, 10, 100, 1030):
        for q in _queries(emb, rng):
            got = _match(conn, q, k)
            od, orow = oracle.vec0_knn(emb, q, k, literal=True)
            assert [r for r, _ in got] == orow.tolist(), f"k={k}: row / tie order differs from sqlite-vec"
            gd = np.array([d for _, d in got], np.float64).astype(np.float32)
            assert gd.view(np.uint32).tolist() == od.view(np.uint32).tolist(), f"k={k}: distances differ"
    conn.close()


def test_exporter_reads_the_real_shadow_tables(tmp_path):
    emb = _corpus(seed=4)
    deleted = (0, 1023, 1024, 2000, emb.shape[0] - 1)
    conn = _real_db(tmp_path, emb, delete=deleted)
    arr = store.export_embeddings(conn)
    assert arr.dim == DIM
    live = np.ones(len(arr.rowid), bool) if arr.valid is None else arr.valid.astype(bool)
    want = np.array([i for i in range(emb.shape[0]) if i not in deleted], np.int64)
    assert arr.rowid[live].tolist() == want.tolist()
    assert (arr.emb[live].view(np.uint32) == emb[want].view(np.uint32)).all()
    assert (arr.movie_idx[live] == want // 5).all()
    # the oracle fed with the EXPORTED arrays (physical positions, validity) == the real extension
    rng = np.random.default_rng(12)
    pos = np.nonzero(live)[0].astype(np.int64)
    for q in _queries(emb, rng, 8):
        got = _match(conn, q, 50)
        od, orow = oracle.vec0_knn(np.ascontiguousarray(arr.emb[live]), q, 50, pos=pos, literal=True)
        assert [r for r, _ in got] == arr.rowid[live][orow].tolist()
    conn.close()


@pytest.mark.gpu
def test_rse_knn_equals_real_sqlite_vec(tmp_path, fresh_index):
    emb = _corpus(seed=5, n=9 * 1024 + 11)
    conn = _real_db(tmp_path, emb, delete=(7, 1024))
    arr = store.export_embeddings(conn)
    fresh_index.load_embeddings(arr.emb, valid=arr.valid, rowid=arr.rowid, movie_idx=arr.movie_idx)
    rng = np.random.default_rng(13)
    Q = _queries(emb, rng, 32)
    for mode in (1, 2):                                     # exact scan, tensor-core path forced
        fresh_index.set_tc_mode(mode)
        dist, _pos, rowid, _movie, cnt = fresh_index.knn(Q, 100)
        for qi in range(Q.shape[0]):
            got = _match(conn, Q[qi], 100)
            assert rowid[qi, :cnt[qi]].tolist() == [r for r, _ in got]
            gd = np.array([d for _, d in got], np.float64).astype(np.float32)
            assert dist[qi, :cnt[qi]].view(np.uint32).tolist() == gd.view(np.uint32).tolist()
    conn.close()
