import numpy as np, sys
sys.path.insert(0, '.')
import oracle
from rag_search_engine_b200 import _lib
rng = np.random.default_rng(11)
n = 5000
emb = rng.standard_normal((n, 384)).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
emb[1023] = emb[7]
valid = (rng.random(n) > 0.1).astype(np.uint8); valid[[7, 1023]] = 1
rowid = rng.permutation(n).astype(np.int64) + 1000
Q = emb[[7, 100, 200, 300]].copy()
keep = np.nonzero(valid)[0]
for name, kw, base, use_valid in [("plain", {}, 0, False), ("base", {"pos_base": 7168}, 7168, False),
                                  ("rowid", {"rowid": rowid}, 0, False), ("valid", {"valid": valid}, 0, True),
                                  ("all", {"valid": valid, "rowid": rowid, "pos_base": 7168}, 7168, True)]:
    idx = _lib.Index(0)
    idx.load_embeddings(emb, **kw)
    dist, pos, rid, _, cnt = idx.knn(Q, 8)
    if use_valid:
        od, orow = oracle.vec0_knn(emb[keep], Q[0], 8, pos=(keep + base).astype(np.int64)); want = (keep[orow] + base)
    else:
        od, orow = oracle.vec0_knn(emb, Q[0], 8); want = orow + base
    print(name, "ok" if pos[0].tolist() == want.tolist() else "BAD", pos[0].tolist(), dist[0].tolist(), want.tolist(), od.tolist())
    idx.close()
