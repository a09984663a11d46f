"""How long does rse_load_embeddings take for an S-600k-sized matrix from a memory-mapped file (page cache warm),
as a function of the staging threads (RSE_UPLOAD_THREADS)?"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rag_search_engine_b200 import _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_800_000
f = Path(tempfile.gettempdir()) / f"rse_cold_{os.getpid()}.npy"
rng = np.random.default_rng(1)
a = np.lib.format.open_memmap(f, mode="w+", dtype=np.float32, shape=(n, 384))
for i in range(0, n, 200_000):
    a[i:i + 200_000] = rng.standard_normal((min(200_000, n - i), 384), dtype=np.float32)
a.flush(); del a
for threads in (1, 2, 4, 8, 16, None):
    if threads is None:
        os.environ.pop("RSE_UPLOAD_THREADS", None)
    else:
        os.environ["RSE_UPLOAD_THREADS"] = str(threads)
    best = 1e9
    for _ in range(2):
        m = np.load(f, mmap_mode="r")
        idx = _lib.Index(0)
        t0 = time.perf_counter()
        idx.load_embeddings(m)
        idx.synchronize()
        best = min(best, time.perf_counter() - t0)
        idx.close(); del m
    print(f"threads {threads}: {best:.3f} s = {n * 1536 / best / 1e9:.1f} GB/s (load_embeddings incl. the row-norm kernel)")
f.unlink()
