import numpy as np, sys
sys.path.insert(0, '.')
from rag_search_engine_b200 import _lib
rng = np.random.default_rng(11)
for dim in (48, 384):
  for n in (2100, 100):
    emb = rng.standard_normal((n, dim)).astype(np.float32); emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    valid = np.ones(n, np.uint8); valid[[3, 40]] = 0
    idx = _lib.Index(0)
    idx.load_embeddings(emb, valid=valid)
    dist, pos, rid, _, cnt = idx.knn(emb[7:8].copy(), 10)
    print("dim", dim, "n", n, "k=10 ->", pos[0, :6].tolist(), dist[0, :6].view(np.uint32).tolist())
    dist, pos, rid, _, cnt = idx.knn(emb[7:8].copy(), n)
    print("   k=n cnt", cnt.tolist(), pos[0, :4].tolist(), dist[0, :4].view(np.uint32).tolist(), "tail", pos[0, -3:].tolist())
    idx.close()
