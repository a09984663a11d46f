"""Equality of the cta_group::2 TMEM-resident filter path (tc_mode 4) with the exact scan (mode 1)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from rag_search_engine_b200 import _lib
rng = np.random.default_rng(1)
def unit(n, d):
    x = rng.standard_normal((n, d)).astype(np.float32); return x / np.linalg.norm(x, axis=1, keepdims=True)
for n, nq, kp in [(9000, 256, 50), (40000, 70, 100), (33000, 300, 100)]:
    centers = unit(50, 384)
    emb = (centers[rng.integers(0, 50, n)] + 0.5 * unit(n, 384)).astype(np.float32)
    emb *= rng.uniform(0.7, 1.4, (n, 1)).astype(np.float32)
    emb[100:140] = emb[100]
    Q = (centers[rng.integers(0, 50, nq)] + 0.3 * unit(nq, 384)).astype(np.float32); Q[0] = emb[7]
    res = {}
    for mode in (1, 4):
        idx = _lib.Index(0); idx.set_tc_mode(mode); idx.load_embeddings(emb)
        t0 = time.time(); res[mode] = idx.knn(Q, kp); dt = time.time() - t0
        st = idx.stats(); print("mode", mode, "n", n, "nq", nq, "tc_queries", st.tc_queries, "fallback", st.tc_fallback_queries, f"{dt*1e3:.1f} ms", flush=True)
        idx.close()
    same = all((a.view(np.uint8) == b.view(np.uint8)).all() for a, b in zip(res[1], res[4]))
    print("  identical:", same, flush=True)
    if not same:
        d1, p1 = res[1][0], res[1][1]; d4, p4 = res[4][0], res[4][1]
        bad = np.nonzero((p1 != p4).any(axis=1))[0]
        print("  differing queries:", bad[:10], "first diff", p1[bad[0]][:10], p4[bad[0]][:10])
