"""oracle/ — CPU restatement of the reference's query hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and there only as the checker or the timed CPU baseline.  Nothing
under ``rag_search_engine_b200/`` imports it (``tests/test_host_cpu.py::test_product_never_touches_the_oracle``
enforces that).

Parity status:
  * BM25 + weighted/RRF fusion: pinned against the reference's own Python code
    imported from /root/reference (``oracle/make_golden.py`` →
    ``tests/golden/*.json``) and the reference's own unit tests.
  * vec0 KNN: **parity unpinned** — sqlite-vec is an un-vendored, unpinned
    third-party dependency (requirements.txt:3) that is not installable here;
    ``oracle.c`` restates its published algorithm.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    """Compile oracle.c → liboracle.so (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = _HERE / "oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC",
               "-shared", "-o", str(_SO), str(src), "-lm"]
        subprocess.run(cmd, check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        c_f = ctypes.POINTER(ctypes.c_float)
        c_i64 = ctypes.POINTER(ctypes.c_int64)
        c_i32 = ctypes.POINTER(ctypes.c_int32)
        c_u32 = ctypes.POINTER(ctypes.c_uint32)
        c_d = ctypes.POINTER(ctypes.c_double)
        L.oracle_cosine_distance.restype = ctypes.c_float
        L.oracle_cosine_distance.argtypes = [c_f, c_f, ctypes.c_int64, ctypes.c_int]
        L.oracle_sq_magnitude.restype = ctypes.c_float
        L.oracle_sq_magnitude.argtypes = [c_f, ctypes.c_int64, ctypes.c_int]
        L.oracle_all_distances.restype = None
        L.oracle_all_distances.argtypes = [c_f, ctypes.c_int64, ctypes.c_int32, c_f, ctypes.c_int, c_f]
        for name in ("oracle_vec0_knn", "oracle_vec0_knn_keyorder"):
            fn = getattr(L, name)
            fn.restype = ctypes.c_int64
            fn.argtypes = [c_f, ctypes.c_int64, ctypes.c_int32, c_i64, c_f, ctypes.c_int32,
                           ctypes.c_int, c_f, c_i64]
        L.oracle_aggregate_movies.restype = ctypes.c_int64
        L.oracle_aggregate_movies.argtypes = [c_f, c_i64, ctypes.c_int64, ctypes.c_int32, c_i64]
        L.oracle_bm25_batch.restype = None
        L.oracle_bm25_batch.argtypes = [c_i64, c_u32, c_u32, c_i64, c_u32, ctypes.c_int64,
                                        ctypes.c_int64, ctypes.c_double, c_i32, c_i32,
                                        ctypes.c_int32, ctypes.c_int32, ctypes.c_double,
                                        ctypes.c_double, c_d, c_i32, c_i32]
        L.oracle_knn_movies_batch.restype = None
        L.oracle_knn_movies_batch.argtypes = [c_f, ctypes.c_int64, ctypes.c_int32, c_i64, c_i64, c_f,
                                              ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                              ctypes.c_int, ctypes.c_int, c_f, c_i64, c_i32]
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_threads.restype = None
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(int(n))      # for a runtime that has not started yet
    lib().oracle_set_threads(int(n))                  # and for one that has


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


# ----------------------------------------------------------------------------- KNN
def cosine_distance(a: np.ndarray, b: np.ndarray, use_fma: bool = False) -> float:
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().oracle_cosine_distance(_p(a, ctypes.c_float), _p(b, ctypes.c_float),
                                              a.size, int(use_fma)))


def all_distances(emb: np.ndarray, q: np.ndarray, use_fma: bool = False) -> np.ndarray:
    emb = np.ascontiguousarray(emb, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    out = np.empty(emb.shape[0], np.float32)
    lib().oracle_all_distances(_p(emb, ctypes.c_float), emb.shape[0], emb.shape[1],
                               _p(q, ctypes.c_float), int(use_fma), _p(out, ctypes.c_float))
    return out


def sq_magnitudes(emb: np.ndarray, use_fma: bool = False) -> np.ndarray:
    emb = np.ascontiguousarray(emb, np.float32)
    L = lib()
    return np.array([L.oracle_sq_magnitude(_p(emb[i], ctypes.c_float), emb.shape[1], int(use_fma))
                     for i in range(emb.shape[0])], np.float32)


def vec0_knn(emb: np.ndarray, q: np.ndarray, k: int, pos: np.ndarray | None = None,
             use_fma: bool = False, literal: bool = True):
    """vec0 ``embedding MATCH :q AND k = :k`` (semantic_search.py:254-261).

    Returns (dist f32[n], row i64[n]) in emit order, n <= k; ``row`` indexes ``emb``.
    """
    emb = np.ascontiguousarray(emb, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    k = int(k)
    out_d = np.empty(max(k, 1), np.float32)
    out_r = np.empty(max(k, 1), np.int64)
    posp = None
    if pos is not None:
        pos = np.ascontiguousarray(pos, np.int64)
        posp = _p(pos, ctypes.c_int64)
    fn = lib().oracle_vec0_knn if literal else lib().oracle_vec0_knn_keyorder
    n = fn(_p(emb, ctypes.c_float), emb.shape[0], emb.shape[1], posp, _p(q, ctypes.c_float), k,
           int(use_fma), _p(out_d, ctypes.c_float), _p(out_r, ctypes.c_int64))
    return out_d[:n].copy(), out_r[:n].copy()


def aggregate_movies(dist: np.ndarray, movie: np.ndarray, k: int) -> np.ndarray:
    """semantic_search.py:285-317 — returns indices (into the KNN rows) of the ≤k movie hits."""
    dist = np.ascontiguousarray(dist, np.float32)
    movie = np.ascontiguousarray(movie, np.int64)
    sel = np.empty(max(len(dist), 1), np.int64)
    n = lib().oracle_aggregate_movies(_p(dist, ctypes.c_float), _p(movie, ctypes.c_int64), len(dist),
                                      int(k), _p(sel, ctypes.c_int64))
    return sel[:n].copy()


def knn_movies_batch(emb, Q, movie_of_row, k, kprime, pos=None, use_fma=False, literal=True):
    emb = np.ascontiguousarray(emb, np.float32)
    Q = np.ascontiguousarray(Q, np.float32)
    mv = np.ascontiguousarray(movie_of_row, np.int64)
    nq = Q.shape[0]
    od = np.zeros((nq, k), np.float32)
    orow = np.full((nq, k), -1, np.int64)
    oc = np.zeros(nq, np.int32)
    posp = None
    if pos is not None:
        pos = np.ascontiguousarray(pos, np.int64)
        posp = _p(pos, ctypes.c_int64)
    lib().oracle_knn_movies_batch(_p(emb, ctypes.c_float), emb.shape[0], emb.shape[1], posp,
                                  _p(mv, ctypes.c_int64), _p(Q, ctypes.c_float), nq, int(k),
                                  int(kprime), int(use_fma), int(literal), _p(od, ctypes.c_float),
                                  _p(orow, ctypes.c_int64), _p(oc, ctypes.c_int32))
    return od, orow, oc


# ----------------------------------------------------------------------------- BM25 (C, CSR)
def bm25_batch(indptr, doc, tf, df, dl, N, avgdl, tok_indptr, terms, k, k1=1.5, b=0.75):
    """keyword_search.py:196-250 over CSR postings; returns (score f64[nq,k], doc i32[nq,k], count)."""
    indptr = np.ascontiguousarray(indptr, np.int64)
    doc = np.ascontiguousarray(doc, np.uint32)
    tf = np.ascontiguousarray(tf, np.uint32)
    df = np.ascontiguousarray(df, np.int64)
    dl = np.ascontiguousarray(dl, np.uint32)
    tok_indptr = np.ascontiguousarray(tok_indptr, np.int32)
    terms = np.ascontiguousarray(terms, np.int32)
    nq = len(tok_indptr) - 1
    os_ = np.zeros((nq, k), np.float64)
    od = np.full((nq, k), -1, np.int32)
    oc = np.zeros(nq, np.int32)
    lib().oracle_bm25_batch(_p(indptr, ctypes.c_int64), _p(doc, ctypes.c_uint32), _p(tf, ctypes.c_uint32),
                            _p(df, ctypes.c_int64), _p(dl, ctypes.c_uint32), len(dl), int(N),
                            float(avgdl), _p(tok_indptr, ctypes.c_int32), _p(terms, ctypes.c_int32),
                            nq, int(k), float(k1), float(b), _p(os_, ctypes.c_double),
                            _p(od, ctypes.c_int32), _p(oc, ctypes.c_int32))
    return os_, od, oc
