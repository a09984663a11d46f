"""Two librse handles on ONE GPU, sharing the corpus in HBM, each running its own stream of hybrid batches from
its own host thread: does the tail of one batch (refine, BM25 finish, aggregation, fusion — ~0.2 ms in which the
filter's SMs idle) hide behind the other handle's probe / filter?  Prints one JSON line:
single-handle ms per batch, dual-handle ms per batch (both handles together), and whether the dual run's results
equal the single run's.

    gpurun -- 'python scripts/dual_handle.py --steps 100 > gpurun_out/dual.json'
"""
from __future__ import annotations

import argparse
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import bench
    from rag_search_engine_b200 import _lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--movies", type=int, default=600_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--handles", type=int, default=2)
    a = ap.parse_args()
    a.tc_mode = a.bm25_mode = -1
    torch.cuda.set_device(0)
    se, bm, tok_indptr, terms, Q, info = bench.build_workload(a, "cuda:0")
    C, dim = info["chunks"], info["dim"]
    NB = bench.NB
    Qn, batches = bench.pinned_batches(torch, Q, tok_indptr, terms, a.batch, 0, a.batch)
    streams = [torch.cuda.Stream() for _ in range(a.handles)]
    handles = [bench.make_handle(a, 0, s, se, bm, 0, C, dim) for s in streams]
    for h in handles:
        for b, (Qb, tp, tr) in enumerate(batches):
            h.hybrid_stage(Qb, tp, tr)
            h.hybrid_stash(b)

    def loop(h, first, n, stride):
        for i in range(n):
            b = (first + i * stride) % NB
            h.hybrid_stash(b)
            h.hybrid_run(0, 60.0, a.limit)
            h.hybrid_stash(b)

    def fetch_all(h):
        out = []
        for b in range(NB):
            h.hybrid_stash(b)
            h.hybrid_run(0, 60.0, a.limit)
            out.append([np.array(x) for x in h.hybrid_fetch(a.limit)])
            h.hybrid_stash(b)
        return out

    # single handle
    loop(handles[0], 0, 8, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(handles[0], 0, a.steps, 1)
    torch.cuda.synchronize()
    single_ms = 1e3 * (time.perf_counter() - t0) / a.steps
    want = fetch_all(handles[0])

    # all handles, one host thread each; handle j takes batches j, j + H, j + 2H, ...
    H = a.handles
    for j, h in enumerate(handles):
        loop(h, j, 8, H)
    torch.cuda.synchronize()
    per = a.steps // H
    threads = [threading.Thread(target=loop, args=(h, j, per, H)) for j, h in enumerate(handles)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    torch.cuda.synchronize()
    dual_ms = 1e3 * (time.perf_counter() - t0) / (per * H)
    same = True
    for h in handles[1:]:
        got = fetch_all(h)
        same = same and all((x.view(np.uint8) == y.view(np.uint8)).all() for g, w in zip(got, want) for x, y in zip(g, w))

    # one host thread alternating between the handles (what a single-threaded server would do)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(per):
        for j, h in enumerate(handles):
            b = (j + i * H) % NB
            h.hybrid_stash(b); h.hybrid_run(0, 60.0, a.limit); h.hybrid_stash(b)
    torch.cuda.synchronize()
    alt_ms = 1e3 * (time.perf_counter() - t0) / (per * H)
    print(json.dumps({"handles": H, "queries_per_batch": a.batch, "steps": a.steps,
                      "single_handle_ms_per_batch": single_ms, "single_qps": a.batch / single_ms * 1e3,
                      "threads_ms_per_batch": dual_ms, "threads_qps": a.batch / dual_ms * 1e3,
                      "one_thread_alternating_ms_per_batch": alt_ms, "alternating_qps": a.batch / alt_ms * 1e3,
                      "same_results": bool(same), "chunks": C}), flush=True)


if __name__ == "__main__":
    main()
