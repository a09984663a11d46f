"""Debug helper: which queries differ between the exact scan (mode 1) and the K4 modes on degenerate inputs."""
import sys
sys.path.insert(0, '.')
import numpy as np
from rag_search_engine_b200 import _lib
rng = np.random.default_rng(5)
def unit_rows(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32); return x / np.linalg.norm(x, axis=1, keepdims=True)
n = 20_000
emb = unit_rows(rng, n, 384)
emb[::7] *= np.float32(1e-3)
emb[1::11] *= np.float32(3e4)
emb[2::13, :200] *= np.float32(1e-6)
emb[500:520] = 0.0
Q = unit_rows(rng, 64, 384)
Q[1] = 0.0
Q[2] = emb[1] * np.float32(1e-4)
Q[3] = emb[2] + np.float32(1e-3) * unit_rows(rng, 1, 384)[0]
out = {}
for mode in (1, 2, 6):
    idx = _lib.Index(0); idx.set_tc_mode(mode); idx.load_embeddings(emb)
    out[mode] = idx.knn(Q, 100); st = idx.stats()
    print("mode", mode, "tc", st.tc_queries, "fallback", st.tc_fallback_queries); idx.close()
for mode in (2, 6):
    for ai, (a, b) in enumerate(zip(out[1], out[mode])):
        bad = np.nonzero((a.reshape(64, -1).view(np.uint8) != b.reshape(64, -1).view(np.uint8)).any(axis=1))[0]
        print("mode", mode, "array", ai, a.dtype, a.shape, "differing queries", bad[:20])
        for q in bad[:3]:
            print("   q", q, "exact", a[q][:6], "tc", b[q][:6])
