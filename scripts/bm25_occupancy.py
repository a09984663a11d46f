"""Experiment: bm25_fx_kernel stand-alone at reduced occupancy (RSE_BM25_PAD pads the CTA's shared memory)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rag_search_engine_b200 import _lib, synth  # noqa: E402

bm = synth.synth_bm25(600_000, 1_000_000, seed=1234)
tok_indptr, terms = synth.synth_token_queries(bm, 256, seed=99)
for pad in (0, 40_000, 80_000, 170_000):
    os.environ["RSE_BM25_PAD"] = str(pad)
    idx = _lib.Index(0)
    idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    for _ in range(3):
        idx.bm25(tok_indptr, terms, 10)
    import time
    ms = []
    for _ in range(7):
        t0 = time.perf_counter()
        idx.bm25(tok_indptr, terms, 10)
        ms.append((time.perf_counter() - t0) * 1e3)
    print(f"pad {pad:7d} B -> CTAs/SM {min(4, 227_000 // (39_000 + pad))}: bm25 call {np.median(ms):.3f} ms (wall, host buffers in and out)")
    idx.close()
