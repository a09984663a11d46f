"""Scan-kernel sweep: device time of knn_scan384<QB> per pass for QB = 1..16 on an S-600k-sized shard."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from rag_search_engine_b200 import _lib, synth
n_movies = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
se = synth.synth_embeddings(n_movies, seed=1234, device="cuda")
C = se.emb.shape[0]
idx = _lib.Index(0)
idx.attach_embeddings_dev(se.emb.data_ptr(), C, 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(), keepalive=se)
Q = synth.synth_query_vectors(se.emb, 64, seed=5).cpu().numpy()
bytes_pass = C * (1536 + 4)
for fma in (False,):
    for nq in (1, 2, 4, 8, 16):
        for _ in range(3):
            idx.knn_movies(Q[:nq], 10, 100)
        idx.set_timing(True); idx.stats_reset()
        t0 = time.perf_counter()
        reps = 20
        for r in range(reps):
            idx.knn_movies(Q[r % 4 * nq % 32:][:nq], 10, 100)
        wall = (time.perf_counter() - t0) / reps
        st = idx.stats(); idx.set_timing(False)
        ms = st.scan_ms_total / st.scan_launches_timed
        print(f"QB={nq:2d} scan {ms:7.3f} ms  {bytes_pass/ms/1e6:7.0f} GB/s  per-query {ms/nq:6.3f} ms | whole call {wall*1e3:7.3f} ms "
              f"({wall*1e3/nq:6.3f}/query, {st.kernel_launches/reps:.0f} launches)")
idx.close()
