"""Reproduce the intermittent batch-1024 failure: many back-to-back hybrid calls."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from rag_search_engine_b200 import _lib, synth
nm = int(sys.argv[1]) if len(sys.argv) > 1 else 600_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 40
se = synth.synth_embeddings(nm, seed=1234, device="cuda")
bm = synth.synth_bm25(nm, 1_000_000 if nm >= 600_000 else 100_000, seed=1234, device="cuda")
tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=99)
Q = synth.synth_query_vectors(se.emb, nq, seed=99).cpu().numpy()
idx = _lib.Index(0)
idx.set_stream(torch.cuda.current_stream().cuda_stream)
idx.attach_embeddings_dev(se.emb.data_ptr(), se.emb.shape[0], 384, movie_idx_ptr=se.movie_of_chunk.data_ptr(), keepalive=se)
idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
idx.set_id_tables(se.movie_ids, se.movie_ids)
ref = idx.hybrid(0, 60.0, 10, Q, tok_indptr, terms)
idx.hybrid_stage(Q, tok_indptr, terms)
for i in range(reps):
    idx.hybrid_run(0, 60.0, 10)
torch.cuda.synchronize()
print("resident runs ok", flush=True)
for i in range(reps):
    r = idx.hybrid(0, 60.0, 10, Q, tok_indptr, terms)
    same = all((a == b).all() for a, b in zip(r, ref))
    if not same:
        print("iteration", i, "DIFFERS", flush=True)
print("host-buffer runs ok", flush=True)
idx.close()
