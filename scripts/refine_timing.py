"""Debug helper (needs a librse.so built with -DRSE_REFINE_TIMING): per-phase timeline of knn_refine_kernel on the
S-600k bench batch.  Prints mean / p50 / max per phase over the 256 CTAs of the last launch."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rag_search_engine_b200 import _lib, synth  # noqa: E402

n_movies = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
dist = sys.argv[2] if len(sys.argv) > 2 else "isotropic"
se = synth.synth_embeddings(n_movies, seed=1234, device="cuda", distribution=dist)
C = se.emb.shape[0]
idx = _lib.Index(0)
idx.attach_embeddings_dev(se.emb.data_ptr(), C, 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(), keepalive=se)
Q = synth.synth_query_vectors(se.emb, 256, seed=99).cpu().numpy()
for _ in range(3):
    idx.knn_movies(Q, 10, 100)
lib = _lib.load_library()
buf = np.zeros(8 * 4096, np.uint64)
rc = lib.rse_debug_refine_ns(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
assert rc == 0
t = buf.reshape(4096, 8)[:256].astype(np.int64)
t0 = t[:, 0].min()
names = ["load pairs+q", "radix select", "band collect", "exact re-score", "sort", "emit"]
print(f"rows={C} dist={dist}; CTA start skew: p50 {np.median(t[:,0]-t0)/1e3:.1f} us, max {(t[:,0]-t0).max()/1e3:.1f} us; "
      f"CTA end: p50 {np.median(t[:,6]-t0)/1e3:.1f} us, max {(t[:,6]-t0).max()/1e3:.1f} us; m (rows re-scored): p50 {np.median(t[:,7])}, max {t[:,7].max()}")
for k, nm in enumerate(names):
    d = (t[:, k + 1] - t[:, k]) / 1e3
    print(f"  {nm:16s} mean {d.mean():6.1f}  p50 {np.median(d):6.1f}  max {d.max():6.1f} us")
slow = np.argsort(t[:, 6])[-3:]
for c in slow:
    print("  slowest CTA", int(c), [(int(t[c, k + 1] - t[c, k]) / 1e3) for k in range(6)], "m", int(t[c, 7]))
idx.close()
