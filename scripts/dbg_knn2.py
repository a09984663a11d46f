import numpy as np, sys
sys.path.insert(0, '.')
from rag_search_engine_b200 import _lib
rng = np.random.default_rng(11)
n = 2100
emb = rng.standard_normal((n, 384)).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
valid = np.ones(n, np.uint8); valid[[3, 40, 1000, 2050]] = 0
for nq in (1, 2, 4, 8):
    idx = _lib.Index(0)
    idx.load_embeddings(emb, valid=valid)
    Q = emb[[7, 100, 200, 300, 301, 302, 303, 304]][:nq].copy()
    dist, pos, rid, _, cnt = idx.knn(Q, n)
    print("nq", nq, "cnt", cnt.tolist(), "invalid in result:", [int((pos[q, :cnt[q]] == r).sum()) for q in range(nq) for r in (3, 40, 1000, 2050)][:8],
          "first", pos[0, :4].tolist(), dist[0, :4].view(np.uint32).tolist())
    idx.close()
