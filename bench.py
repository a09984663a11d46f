#!/usr/bin/env python
"""bench.py — hybrid queries/sec (RRF, top-10) on the movies_600k-shaped synthetic corpus.

    python bench.py --gpus N --steps K --warmup W            # this repo (librse on the B200)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A "step" = one batch of hybrid queries through the whole hot path
(BM25 top-10 + exact vec0 KNN top-100 + best-chunk-per-movie + RRF fusion).
  value : whole-job queries/s with the query batch already resident in HBM
          (rse_hybrid_stage once, then K × rse_hybrid_run), CUDA events, max over ranks.
  e2e   : the same metric through the host-buffer C-ABI calls: every step uploads its query
          vectors/tokens (host -> pinned -> device) and reads its results back inside the timed
          region.  Headline = the serving loop (rse_hybrid_submit / rse_hybrid_collect, two batches
          in flight); e2e.blocking_call = rse_hybrid one batch at a time.
N > 1   : weak scaling by default — the corpus replicated, the query batch split by rank, no
          data-path collective (--parallelism rowshard: corpus row-sharded, one all_gather of the
          local top-K' candidates per step, BM25/fusion split by query slice).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

ROW_BYTES = 384 * 4 + 4          # algorithmic bytes per chunk row per pass: 1536 B vector + 4 B |a|^2
ROW_BYTES_SHADOW = 384 * 2       # the fp16-shadow filter pass (knn_tc3) streams 768 B per row and nothing else


def tc_kind(args):
    """The K4 kernel (r01's two TF32 mappings were removed in r02: 2.6-3.5x slower on every measured shape)."""
    return "f16-shadow"


def scan_kernel_desc(args, tc_used, qb):
    if not tc_used:
        return (f"knn_scan384_kernel<QB={qb}> (one pass over the shard serves {qb} queries; FFMA2-pipe-bound above QB=4, "
                f"HBM-bound at QB=1: see knn_batch1)"), ROW_BYTES
    return ("knn_tc3_kernel<filter> (tcgen05 kind::f16 256x256x16 cta_group::2 over the fp16 normalised shadow, "
            "queries resident in shared memory, TMA 4-stage ring; one pass serves 256 queries; survivors re-scored "
            "exactly in fp32; in the hybrid step one bm25_fx_kernel CTA per SM runs underneath it on a second "
            "stream, which costs the pass ~8 %)"), ROW_BYTES_SHADOW


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--movies", type=int, default=600_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=256, help="hybrid queries per step")
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--mode", default="rrf", choices=["rrf", "weighted"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="queries in the cpu_baseline sample (0 = 2 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-knn1", action="store_true", help="skip the batch-1 KNN micro-measurement")
    ap.add_argument("--workload", default="hybrid600k", choices=["hybrid600k", "bm25_10k", "knn100m"],
                    help="hybrid600k = configs[3] (the metric's configuration, default); bm25_10k = configs[1]; "
                         "knn100m = configs[4] (12.5 M-row shard per GPU, top-100, NCCL candidate merge)")
    ap.add_argument("--shard-rows", type=int, default=12_500_000)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = --batch queries per GPU per step (global batch = batch x N; every rank scans its "
                         "row shard for all of them, so per-GPU work is constant); strong = the same --batch split over N")
    ap.add_argument("--parallelism", default="auto", choices=["auto", "replicate", "rowshard"],
                    help="N > 1: replicate = every GPU holds the corpus and serves its own slice of the query batch (hybrid "
                         "queries are independent units: no data-path collective); rowshard = the corpus is cut by row, "
                         "local top-K' + one all_gather + merge (what a corpus that does not fit one GPU needs: "
                         "--workload knn100m always uses it).  auto = replicate when corpus + shadow fit in 1/3 of HBM; "
                         "the JSON line then carries a shorter rowshard measurement of the same batch as well")
    ap.add_argument("--bm25-mode", type=int, default=-1, help="rse_set_bm25_mode override (see include/rse.h); -1 = library default")
    ap.add_argument("--tc-mode", type=int, default=-1, help="rse_set_tc_mode override (see include/rse.h); -1 = library default")
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  The timed region of the default run is tens
    of milliseconds, far below nvidia-smi's loop granularity, so NVML is polled from a thread every ~1 ms
    (nvidia-smi -lms is the fallback when the NVML binding is missing)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.sm, self.reasons = [], set()
        self.sm_max = None
        self.path = Path(os.environ.get("TMPDIR", "/tmp")) / f"rse_clocks_{os.getpid()}.csv"

    def _poll(self, nv, handle):
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        def once():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(handle))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
        period = float(os.environ.get("RSE_CLOCK_POLL_MS", "1")) * 1e-3
        while not self.stop_flag:
            once()
            time.sleep(period)

    def start(self):
        try:
            import pynvml as nv
            import threading
            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: address the GPU by the PCI bus id torch reports
            handle = None
            try:
                import torch
                bus = torch.cuda.get_device_properties(self.gpu).pci_bus_id
                dom = torch.cuda.get_device_properties(self.gpu).pci_domain_id
                dev = torch.cuda.get_device_properties(self.gpu).pci_device_id
                handle = nv.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
            except Exception:
                handle = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.sm:
                out.update(sm_mhz=statistics.median(self.sm), sm_max_mhz=self.sm_max, reasons=sorted(self.reasons),
                           samples=len(self.sm), source="nvml polling thread (one query takes ~25 ms)")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in self.path.read_text().splitlines():
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       source="nvidia-smi -lms 20")
        try:
            self.path.unlink()
        except Exception:
            pass
        return out


# ----------------------------------------------------------------------------- workload
def build_workload(args, device: str):
    """Seeded S-600k corpus + query batch (identical on every rank and in both arms)."""
    import torch
    from rag_search_engine_b200 import synth
    t0 = time.time()
    se = synth.synth_embeddings(args.movies, seed=1234, device=device)
    bm = synth.synth_bm25(args.movies, args.vocab, seed=1234, device=device)
    tok_indptr, terms = synth.synth_token_queries(bm, args.batch, seed=99)
    Q = synth.synth_query_vectors(se.emb, args.batch, seed=99)
    if device != "cpu":
        torch.cuda.synchronize()
    info = {"chunks": int(se.emb.shape[0]), "dim": int(se.emb.shape[1]), "movies": args.movies,
            "postings": int(len(bm.doc_idx)), "terms": int(len(bm.df)), "build_s": round(time.time() - t0, 1)}
    return se, bm, tok_indptr, terms, Q, info


def postings_touched(bm, tok_indptr, terms) -> int:
    t = terms[terms >= 0]
    return int(bm.df[t].sum())


def config_dict(args, info, world, par="replicate"):
    return {"workload": "configs[3]: movies_600k-shaped synthetic (S-600k), hybrid rrf_search k=60 limit=10 "
                        "(BM25 top-10 + exact vec0 KNN top-100 + per-movie best chunk + RRF), Gemini disabled",
            "mode": args.mode, "limit": args.limit, "knn_kprime": max(args.limit * 10, args.limit),
            "queries_per_step": args.batch, "queries_per_gpu_per_step": args.batch // max(1, world) if getattr(args, "scaling", "weak") == "weak" else None,
            "movies": info["movies"], "chunks": info["chunks"], "dim": info["dim"],
            "bm25_postings": info["postings"], "bm25_terms": info["terms"],
            "l2": "inputs larger than L2: every step streams the 3.7 GB fp16 shadow of the corpus (one pass per 256 "
                  "queries) and ~1.3 GB of postings against a 126 MB L2; no flush",
            "parallelism": ("1 GPU" if world == 1 else
                            (f"row-shard x{world} + all_gather(top-K') + query-slice BM25/fusion" if par == "rowshard" else
                             f"corpus replicated x{world}, query batch split by rank, no data-path collective"))}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_hybrid_sample(se_emb_host, movie_of, bm, ids, Q, tok_indptr, terms, limit, mode, sample, threads):
    """The reference's CPU path restated (oracle/): literal vec0 scan + aggregation, BM25, fusion."""
    import oracle
    from oracle import pyref
    oracle.set_threads(threads)
    nq = sample
    tp = tok_indptr[: nq + 1].astype(np.int32)
    tr = terms[: tp[-1]]
    t0 = time.perf_counter()
    kd, krow, kc = oracle.knn_movies_batch(se_emb_host, Q[:nq], movie_of, limit, max(limit * 10, limit), literal=True)
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tp, tr, limit)
    out = []
    for qi in range(nq):
        bmh = [(int(ids[odc[qi, j]]), float(osc[qi, j])) for j in range(ocnt[qi])]
        semh = [(int(ids[movie_of[krow[qi, j]]]), float(kd[qi, j])) for j in range(kc[qi])]
        out.append(pyref.rrf_fuse(bmh, semh, 60, limit) if mode == "rrf" else pyref.weighted_fuse(bmh, semh, 0.5, limit))
    dt = time.perf_counter() - t0
    return nq / dt, dt, out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the box's host cores.
    The reference is Python + un-vendored sqlite-vec and cannot travel to the GPU box, so this
    times the oracle port (oracle/oracle.c: literal vec0 scan; BM25 over CSR; fusion in Python)."""
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"   # data generation only
    se, bm, tok_indptr, terms, Q, info = build_workload(args, dev)
    emb_host = se.emb.cpu().numpy()
    movie_of = se.movie_of_chunk.cpu().numpy().astype(np.int64)
    Qh = Q.cpu().numpy()
    del se.emb
    cores = os.cpu_count() or 1
    sample = min(args.batch, args.cpu_sample or cores)
    import oracle
    oracle.build()
    times = []
    budget_s = 150.0          # the whole arm must end within a few minutes whatever --steps says
    t_start = time.perf_counter()
    for step in range(args.warmup + args.steps):
        if step > args.warmup + 1 and time.perf_counter() - t_start > budget_s:
            break
        off = (step * sample) % max(1, args.batch - sample + 1)
        tp = (tok_indptr[off: off + sample + 1] - tok_indptr[off]).astype(np.int32)
        tr = terms[tok_indptr[off]: tok_indptr[off + sample]]
        qps, dt, _ = cpu_hybrid_sample(emb_host, movie_of, bm, se.movie_ids, Qh[off: off + sample], tp, tr, args.limit,
                                       args.mode, sample, cores)
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": "hybrid queries/sec", "value": value, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "replicas only (CPU path, rank 0)", "vs_baseline": None,
            "dtype": "f32 sequential scan + f64 tail/BM25/fusion (the reference's arithmetic)",
            "data": "synthetic", "config": config_dict(args, info, 1),
            "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} hybrid queries per step over the full S-600k corpus "
                                       f"(OpenMP over queries, each query a single-threaded literal vec0 scan)"},
            "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def measure_rowshard(args, rank, world, local_rank, se, bm, tok_indptr, terms, Q, Qn, info, mode, param, limit, steps):
    """The same global batch through the ROW-SHARDED path (north_star: local top-K' per shard + one NCCL all_gather
    + merge; sharded.py): device-timed steps with the batch resident, max over ranks, checked against one handle."""
    import torch
    import torch.distributed as dist
    from rag_search_engine_b200 import _lib, sharded
    device = torch.device("cuda", local_rank)
    C, nq = info["chunks"], Qn.shape[0]
    bounds = sharded.shard_bounds(C, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    idx = _lib.Index(local_rank)
    if args.tc_mode >= 0:
        idx.set_tc_mode(args.tc_mode)
    stream = torch.cuda.current_stream(device)
    idx.set_stream(stream.cuda_stream)
    shard = se.emb[lo:hi]
    mo = se.movie_of_chunk[lo:hi].contiguous()
    idx.attach_embeddings_dev(shard.data_ptr(), hi - lo, info["dim"], movie_idx_ptr=mo.data_ptr(), pos_base=lo,
                              keepalive=(se, shard, mo))
    idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    idx.set_id_tables(se.movie_ids, se.movie_ids)
    backend = sharded.LibrseShardBackend(idx, Qn, tok_indptr, terms, device)
    sh = sharded.ShardedHybrid(backend, nq)
    Qd = Q.to(device).contiguous()
    for _ in range(3):
        sh.step(Qd, mode, param, limit)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        r = sh.step(Qd, mode, param, limit)
    e1.record(stream)
    torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ok = None
    if rank == 0:
        chk = _lib.Index(local_rank)
        chk.set_stream(stream.cuda_stream)
        mo_full = se.movie_of_chunk.contiguous()
        chk.attach_embeddings_dev(se.emb.data_ptr(), C, info["dim"], movie_idx_ptr=mo_full.data_ptr(), keepalive=(se, mo_full))
        chk.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
        chk.set_id_tables(se.movie_ids, se.movie_ids)
        oid, osc, oa, ob, oc = chk.hybrid(mode, param, limit, Qn, tok_indptr, terms)
        ok = bool((r.ids.cpu().numpy() == oid).all() and (r.score.cpu().numpy() == osc).all() and
                  (r.count.cpu().numpy() == oc).all())
        chk.close()
    idx.close()
    dist.barrier()
    return {"value": nq * steps / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms / steps, "steps": steps,
            "queries_per_step": nq, "parallelism": f"row-shard x{world} + all_gather(top-K') + query-slice BM25/fusion",
            "sharded_matches_single_gpu": ok}


# ----------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from rag_search_engine_b200 import _lib, sharded

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    per_gpu_batch = args.batch
    if world > 1 and args.scaling == "weak":
        args.batch = args.batch * world          # global batch; every rank scans its shard for all of it
    se, bm, tok_indptr, terms, Q, info = build_workload(args, f"cuda:{local_rank}")
    C = info["chunks"]
    par = args.parallelism
    if par == "auto":
        fits = C * info["dim"] * 6 < torch.cuda.get_device_properties(device).total_memory / 3
        par = "replicate" if fits else "rowshard"
    rowshard = world > 1 and par == "rowshard"
    bounds = sharded.shard_bounds(C, world) if rowshard else [0] + [C] * world
    lo, hi = (bounds[rank], bounds[rank + 1]) if rowshard else (0, C)
    idx = _lib.Index(local_rank)
    if args.tc_mode >= 0:
        idx.set_tc_mode(args.tc_mode)
    if args.bm25_mode >= 0:
        idx.set_bm25_mode(args.bm25_mode)
    stream = torch.cuda.current_stream(device)
    idx.set_stream(stream.cuda_stream)
    shard = se.emb[lo:hi]
    mo = se.movie_of_chunk[lo:hi].contiguous()
    idx.attach_embeddings_dev(shard.data_ptr(), hi - lo, info["dim"], movie_idx_ptr=mo.data_ptr(), pos_base=lo,
                              keepalive=(se, shard, mo))
    idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    idx.set_id_tables(se.movie_ids, se.movie_ids)
    mode = 0 if args.mode == "rrf" else 1
    param = 60.0 if mode == 0 else 0.5
    limit, nq = args.limit, args.batch
    Qh = torch.empty((nq, info["dim"]), dtype=torch.float32, pin_memory=True)
    Qh.copy_(Q.cpu())
    Qn = Qh.numpy()
    # replicate: this rank's slice of the global batch (the whole batch when N = 1)
    qs = sharded.query_slices(nq, world)
    qlo, qhi = (qs[rank], qs[rank + 1]) if (world > 1 and not rowshard) else (0, nq)
    Qn_loc = Qn[qlo:qhi]
    tp_loc = (tok_indptr[qlo:qhi + 1] - tok_indptr[qlo]).astype(np.int32)
    tr_loc = terms[tok_indptr[qlo]:tok_indptr[qhi]]
    nq_loc = qhi - qlo

    def barrier():
        if world > 1:
            dist.barrier()

    if rowshard:
        backend = sharded.LibrseShardBackend(idx, Qn, tok_indptr, terms, device)
        sh = sharded.ShardedHybrid(backend, nq)
        Qd = Q.to(device).contiguous()
        step_fn = lambda: sh.step(Qd, mode, param, limit)        # noqa: E731
    else:
        idx.hybrid_stage(Qn_loc, tp_loc, tr_loc)
        step_fn = lambda: idx.hybrid_run(mode, param, limit)     # noqa: E731

    for _ in range(max(args.warmup, 3)):
        step_fn()
    torch.cuda.synchronize(); barrier()

    # ---- timed region: K steps, inputs resident in HBM
    idx.set_timing(True)
    idx.stats_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_fn()                             # (no inline NVML sampling here: one query costs the host ~25 ms and the
    e1.record(stream)                         #  device would idle; the sampler thread polls concurrently instead)
    torch.cuda.synchronize(); barrier()
    clocks = sampler.stop()
    sh_last = step_fn() if rowshard else None
    torch.cuda.synchronize(); barrier()
    ms = e0.elapsed_time(e1)
    st = idx.stats()
    idx.set_timing(False)
    launches = int(st.kernel_launches)
    scan_ms = st.scan_ms_total / max(1, st.scan_launches_timed)
    scan_share = st.scan_ms_total / ms if ms > 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = nq * args.steps / (ms / 1e3)

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region
    e2e = None
    res = None
    if not args.no_e2e and not rowshard:
        # every rank: its slice of the batch through the host-buffer call (H2D of queries + tokens, D2H of results)
        def wall_max(w):
            if world > 1:
                t = torch.tensor([w], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t.item())
            return w
        # (1) the blocking call, one batch at a time: the device idles while the host stages and collects
        for _ in range(2):
            idx.hybrid(mode, param, limit, Qn_loc, tp_loc, tr_loc)
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            res = idx.hybrid(mode, param, limit, Qn_loc, tp_loc, tr_loc)
        e1.record(stream)
        torch.cuda.synchronize()
        wall_blocking = wall_max(time.perf_counter() - t0)
        dev_blocking = e0.elapsed_time(e1) / args.steps
        # (2) the serving loop: rse_hybrid_submit / rse_hybrid_collect, two batches in flight — every step still
        # uploads its own inputs from host buffers and reads its own results back inside the timed region
        for _ in range(2):                                       # warm both ticket slots (their buffers are allocated lazily)
            t_prev = idx.hybrid_submit(mode, param, limit, Qn_loc, tp_loc, tr_loc)
            t_next = idx.hybrid_submit(mode, param, limit, Qn_loc, tp_loc, tr_loc)
            idx.hybrid_collect(t_prev); idx.hybrid_collect(t_next)
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        t_prev = idx.hybrid_submit(mode, param, limit, Qn_loc, tp_loc, tr_loc)
        for _ in range(args.steps - 1):
            t_next = idx.hybrid_submit(mode, param, limit, Qn_loc, tp_loc, tr_loc)
            res_p = idx.hybrid_collect(t_prev)
            t_prev = t_next
        res_p = idx.hybrid_collect(t_prev)
        torch.cuda.synchronize()
        wall = wall_max(time.perf_counter() - t0)
        pipelined_ok = all((a.view(np.uint8) == b.view(np.uint8)).all() for a, b in zip(res, res_p))
        ntok = int(tok_indptr[-1])
        h2d = Qn.nbytes + (nq + world) * 4 + ntok * (4 + 8)
        d2h = nq * limit * (8 + 8 + 8 + 8) + nq * 4
        e2e = {"value": nq * args.steps / wall, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "wall_ms_per_step": 1e3 * wall / args.steps,
               "api": "rse_hybrid_submit / rse_hybrid_collect, host buffers, two batches in flight",
               "same_results_as_blocking_call": bool(pipelined_ok),
               "blocking_call": {"value": nq * args.steps / wall_blocking, "unit": "queries/s",
                                 "api": "rse_hybrid (one batch at a time)", "device_ms_per_step": dev_blocking,
                                 "wall_ms_per_step": 1e3 * wall_blocking / args.steps}}
    elif rowshard and not args.no_e2e:
        # sharded e2e: stage (H2D of this rank's slice + the batch's query vectors) + step + D2H of the fused batch
        torch.cuda.synchronize(); barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            backend.stage_slice(sh.lo, sh.hi)
            Qd2 = Qh.to(device, non_blocking=True)
            r = sh.step(Qd2, mode, param, limit)
            _ = r.ids.cpu(), r.score.cpu(), r.count.cpu()
        torch.cuda.synchronize(); barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([wall], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
        e2e = {"value": nq * args.steps / wall, "unit": "queries/s", "h2d_bytes_per_step": int(Qn.nbytes),
               "d2h_bytes_per_step": int(nq * limit * 16 + nq * 4), "wall_ms_per_step": 1e3 * wall / args.steps}

    # ---- batch-1 KNN (north_star: batch-1 kNN as a fraction of HBM peak): scan<QB=1> alone + whole call
    knn1 = None
    knn1k = None
    if world == 1 and not args.no_knn1:
        kp = max(limit * 10, limit)
        for _ in range(3):
            idx.knn_movies(Qn[:1], limit, kp)
        idx.set_timing(True); idx.stats_reset()
        reps = 20
        t0 = time.perf_counter()
        for i in range(reps):
            idx.knn_movies(Qn[i % nq: i % nq + 1], limit, kp)
        wall1 = (time.perf_counter() - t0) / reps
        s1 = idx.stats(); idx.set_timing(False)
        sm1 = s1.scan_ms_total / max(1, s1.scan_launches_timed)
        knn1 = {"scan_ms": sm1, "call_ms_host_buffers": 1e3 * wall1, "launches_per_query": s1.kernel_launches / reps}
        # configs[2]: batch-1024 semantic search (KNN top-K' + per-movie best chunk) through the host-buffer call
        from rag_search_engine_b200 import synth as _synth
        Q1k = _synth.synth_query_vectors(se.emb, 1024, seed=7).cpu().numpy()
        for _ in range(2):
            idx.knn_movies(Q1k, limit, kp)
        idx.set_timing(True); idx.stats_reset()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            idx.knn_movies(Q1k, limit, kp)
        wall1k = (time.perf_counter() - t0) / reps
        s1k = idx.stats(); idx.set_timing(False)
        f_ms = s1k.scan_ms_total / max(1, s1k.scan_launches_timed)
        knn1k = {"call_ms_host_buffers": 1e3 * wall1k, "queries_per_s": 1024 / wall1k, "filter_pass_ms": f_ms,
                 "filter_tflops": 2.0 * 384 * (hi - lo) * 256 / (f_ms * 1e-3) / 1e12 if f_ms > 0 else None,
                 "tc_fallback_queries": int(s1k.tc_fallback_queries)}

    sharded_ok = None
    replicas_ok = None
    rowshard_extra = None
    mismatch = None
    corpora_identical = None
    if world > 1:
        # every rank generates the corpus on its own GPU from the same seeds: verify they really are the same bits
        def csum(t):
            t = t.contiguous().view(torch.int32) if t.dtype == torch.float32 else t.to(torch.int64)
            return int(t.to(torch.int64).sum().item())
        sig = [csum(se.emb), csum(se.movie_of_chunk), csum(torch.as_tensor(np.asarray(bm.doc_idx).astype(np.int64))),
               csum(torch.as_tensor(np.asarray(bm.tf).astype(np.int64))), int(np.asarray(tok_indptr).sum()),
               int(np.asarray(terms).astype(np.int64).sum()), csum(torch.as_tensor(Qn).view(torch.int32))]
        sigs = [None] * world
        dist.all_gather_object(sigs, sig)
        corpora_identical = all(x == sigs[0] for x in sigs)
    if world > 1:
        # rank 0 generated the whole corpus: check the N-GPU result against a single-handle run of the whole batch
        if rowshard:
            got = (sh_last.ids.cpu().numpy(), sh_last.score.cpu().numpy(), sh_last.count.cpu().numpy())
        else:
            mine = res if res is not None else idx.hybrid(mode, param, limit, Qn_loc, tp_loc, tr_loc)
            parts = [None] * world
            dist.all_gather_object(parts, (mine[0], mine[1], mine[4]))
            got = tuple(np.concatenate([p[i] for p in parts]) for i in range(3))
        if rank == 0:
            if rowshard:
                chk = _lib.Index(local_rank)
                chk.set_stream(stream.cuda_stream)
                mo_full = se.movie_of_chunk.contiguous()
                chk.attach_embeddings_dev(se.emb.data_ptr(), C, info["dim"], movie_idx_ptr=mo_full.data_ptr(),
                                          keepalive=(se, mo_full))
                chk.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
                chk.set_id_tables(se.movie_ids, se.movie_ids)
            else:
                chk = idx
            oid, osc, oa, ob, oc = chk.hybrid(mode, param, limit, Qn, tok_indptr, terms)
            ok = bool((got[0] == oid).all() and (got[1] == osc).all() and (got[2] == oc).all())
            if not ok:
                badq = np.nonzero((got[0] != oid).any(axis=1) | (got[1] != osc).any(axis=1) | (got[2] != oc))[0]
                mismatch = {"queries": int(len(badq)), "first": [int(x) for x in badq[:8]],
                            "ids_differ": int((got[0] != oid).any(axis=1).sum()),
                            "scores_differ": int((got[1] != osc).any(axis=1).sum()),
                            "example": {"q": int(badq[0]), "got_id": [int(x) for x in got[0][badq[0]]],
                                        "want_id": [int(x) for x in oid[badq[0]]],
                                        "got_score": [float(x) for x in got[1][badq[0]]],
                                        "want_score": [float(x) for x in osc[badq[0]]]}}
            if rowshard:
                sharded_ok = ok
                chk.close()
            else:
                replicas_ok = ok
        if not rowshard and args.parallelism == "auto":
            rowshard_extra = measure_rowshard(args, rank, world, local_rank, se, bm, tok_indptr, terms, Q, Qn, info, mode,
                                              param, limit, min(args.steps, 30))
    if world > 1:
        dist.barrier()
    if rank != 0:
        return
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rows_local = hi - lo
    qb = 16 if nq >= 9 else (8 if nq >= 5 else (4 if nq >= 3 else nq))
    tc_used = int(st.tc_filter_launches) > 0
    kname, row_bytes = scan_kernel_desc(args, tc_used, qb)
    alg_bytes = rows_local * row_bytes
    # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture (profiles/): only
    # quoted when this run has the shard size that was profiled
    traffic = None
    tj = ROOT / "profiles" / "r01_dominant_kernel_traffic.json"
    if tj.exists():
        t = json.loads(tj.read_text())
        if t.get("rows") == rows_local and t.get("tc_kind") == (tc_kind(args) if tc_used else "scan"):
            traffic = t.get("dram_bytes_read", 0) + t.get("dram_bytes_write", 0)
    hbm_achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else None
    psrc = "MEASURED_PEAKS.json" if peaks else "fallback"
    if tc_used:
        # The batched filter pass is bound by the tensor pipe (+ TMEM reads), not by the stream: with the MMAs
        # removed the same kernel streams the shadow at 7.4 TB/s (0.50 ms), with the TMEM loads removed it takes
        # 0.65 ms = 1.45 PFLOP/s.  Algorithmic work: 2*384 flop per (row, query) pair.
        q_per_pass = min(nq if rowshard else nq_loc, 256)
        flops = 2.0 * 384 * rows_local * q_per_pass
        tpeak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        tach = flops / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None
        roofline = {"bound": "tensor", "kernel": kname, "achieved": tach, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": (tach / tpeak) if tach else None,
                    "peak_source": f"{psrc} bf16_tflops_sustained (kernel timed inside the step; f16 and bf16 share the "
                                   f"tensor rate; burst figure {peaks.get('bf16_tflops')})",
                    "traffic": traffic, "algorithmic_flops_per_launch": flops, "queries_per_pass": q_per_pass,
                    "avg_launch_ms": scan_ms,
                    "hbm": {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s",
                            "frac": (hbm_achieved / peak) if hbm_achieved else None,
                            "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_row": row_bytes},
                    "scan_launches": int(st.scan_launches_timed), "scan_share_of_step": scan_share,
                    "tc_queries": int(st.tc_queries), "tc_fallback_queries": int(st.tc_fallback_queries)}
    else:
        roofline = {"bound": "hbm", "kernel": kname,
                    "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": (hbm_achieved / peak) if hbm_achieved else None,
                    "peak_source": f"{psrc} hbm_gbs", "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
                    "avg_launch_ms": scan_ms, "scan_launches": int(st.scan_launches_timed), "scan_share_of_step": scan_share,
                    "tc_queries": int(st.tc_queries), "tc_fallback_queries": int(st.tc_fallback_queries)}

    if knn1:
        g1 = rows_local * ROW_BYTES / (knn1["scan_ms"] * 1e-3) / 1e9
        knn1.update(achieved_gbs=g1, frac_of_measured_peak=g1 / peak, queries_per_s=1e3 / knn1["call_ms_host_buffers"])
    line = {"metric": "hybrid queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": (args.scaling if world > 1 else "weak"), "vs_baseline": None,
            "dtype": "f16 tensor-core filter + exact f32 re-score (f64 tail), f64 BM25/fusion", "data": "synthetic",
            "config": config_dict(args, info, world, par), "roofline": roofline, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "sharded_matches_single_gpu": sharded_ok, "replicas_match_single_gpu": replicas_ok, "mismatch": mismatch,
            "corpora_identical_across_ranks": corpora_identical,
            "rowshard": rowshard_extra, "knn_batch1": knn1, "knn_batch1024": knn1k, "postings_touched_per_step": postings_touched(bm, tok_indptr, terms)}

    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(nq, args.cpu_sample or 2 * cores)
        emb_host = se.emb.cpu().numpy()
        movie_of = se.movie_of_chunk.cpu().numpy().astype(np.int64)
        qps, dt, cpu_out = cpu_hybrid_sample(emb_host, movie_of, bm, se.movie_ids, Qn, tok_indptr, terms, limit, args.mode,
                                             sample, cores)
        # the sample doubles as a full-size parity check of the GPU results
        oid, osc, oa, ob, oc = res if e2e and world == 1 else idx.hybrid(mode, param, limit, Qn, tok_indptr, terms)
        ok = all([int(x) for x in oid[q, :oc[q]]] == [r["id"] for r in cpu_out[q]] and
                 [float(x) for x in osc[q, :oc[q]]] == [r["score"] for r in cpu_out[q]] for q in range(sample))
        # the reference itself is single-threaded: the same port on ONE core, two queries
        qps1, dt1, _ = cpu_hybrid_sample(emb_host, movie_of, bm, se.movie_ids, Qn, tok_indptr, terms, limit, args.mode, 2, 1)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "one_core": {"value": qps1, "unit": "queries/s", "sample": f"2 queries, {dt1:.1f} s"},
                                "sample": f"{sample} hybrid queries of the same batch over the full corpus, "
                                          f"{dt:.1f} s (OpenMP over queries; literal vec0 scan per query)",
                                "gpu_matches_cpu_on_sample": bool(ok)}
    print(json.dumps(line), flush=True)
    idx.close()


def run_bm25_10k(args, local_rank):
    """configs[1]: key_search BM25 top-10, 10k-query batch over the GPU CSR postings (host-buffer call)."""
    import torch
    from rag_search_engine_b200 import _lib, synth
    torch.cuda.set_device(local_rank)
    bm = synth.synth_bm25(args.movies, args.vocab, seed=1234, device=f"cuda:{local_rank}")
    nq = 10_000
    tok_indptr, terms = synth.synth_token_queries(bm, nq, seed=99)
    idx = _lib.Index(local_rank)
    idx.load_bm25(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl)
    for _ in range(max(args.warmup, 3)):
        sc, dc, cnt = idx.bm25(tok_indptr, terms, 10)
    idx.stats_reset()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        sc, dc, cnt = idx.bm25(tok_indptr, terms, 10)
    wall = (time.perf_counter() - t0) / args.steps
    launches = int(idx.stats().kernel_launches)
    touched = postings_touched(bm, tok_indptr, terms)
    import oracle
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    ns = 512
    tp = tok_indptr[: ns + 1].astype(np.int32)
    t0 = time.perf_counter()
    osc, odc, ocnt = oracle.bm25_batch(bm.indptr, bm.doc_idx, bm.tf, bm.df, bm.dl, bm.n_movies, bm.avgdl, tp, terms[: tp[-1]], 10)
    cpu_dt = time.perf_counter() - t0
    ok = bool((odc == dc[:ns]).all() and (osc.view(np.uint64) == sc[:ns].view(np.uint64)).all() and (ocnt == cnt[:ns]).all())
    peak = 6546.6
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
    ach = touched * 12 / wall / 1e9
    print(json.dumps({"metric": "BM25 top-10 queries/sec (10k-query batch)", "value": nq / wall, "unit": "queries/s",
                      "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * wall,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "configs[1]: movies_600k-shaped BM25 index, 10k-query batch, top-10",
                                 "docs": args.movies, "postings": int(len(bm.doc_idx)), "postings_touched": touched},
                      "roofline": {"bound": "hbm", "kernel": "bm25_score_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                                   "frac": ach / peak, "traffic": None,
                                   "note": "12 B per touched posting over the whole host-buffer call (H2D/D2H included)"},
                      "e2e": {"value": nq / wall, "unit": "queries/s", "h2d_bytes_per_step": int(terms.nbytes * 3 + tok_indptr.nbytes),
                              "d2h_bytes_per_step": int(sc.nbytes + dc.nbytes + cnt.nbytes)},
                      "gpu_launches": launches,
                      "cpu_baseline": {"value": ns / cpu_dt, "unit": "queries/s", "cores": cores, "kind": "port",
                                       "sample": f"first {ns} queries of the batch (oracle.c BM25 over CSR, OpenMP)",
                                       "gpu_matches_cpu_on_sample": ok}}), flush=True)
    idx.close()


def run_knn100m(args, rank, world, local_rank):
    """configs[4]: synthetic 100M x 384 corpus as 12.5 M-row shards (one per GPU), top-100 with the NCCL
    candidate merge.  Weak scaling: every rank always scans a full shard."""
    import torch
    import torch.distributed as dist
    from rag_search_engine_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rows = (args.shard_rows // 1024) * 1024
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    emb = torch.empty((rows, 384), dtype=torch.float32, device=dev)
    for s0 in range(0, rows, 1 << 20):
        e0_ = min(rows, s0 + (1 << 20))
        x = torch.randn((e0_ - s0, 384), generator=g, device=dev)
        emb[s0:e0_] = x / x.norm(dim=1, keepdim=True)
    base = rank * rows
    movie = (torch.arange(rows, device=dev, dtype=torch.int64) + base).to(torch.int32)
    idx = _lib.Index(local_rank)
    if args.tc_mode >= 0:
        idx.set_tc_mode(args.tc_mode)
    stream = torch.cuda.current_stream(dev)
    idx.set_stream(stream.cuda_stream)
    idx.attach_embeddings_dev(emb.data_ptr(), rows, 384, movie_idx_ptr=movie.data_ptr(), pos_base=base, keepalive=(emb, movie))
    nq, kp = args.batch, 100
    gq = torch.Generator(device=dev).manual_seed(99)
    Q = torch.randn((nq, 384), generator=gq, device=dev)
    Q /= Q.norm(dim=1, keepdim=True)
    if rank == 0:
        Q[: min(8, nq)] = emb[torch.arange(min(8, nq), device=dev) * 1000 + 5]       # known self-hits
    if world > 1:
        dist.broadcast(Q, 0)
    cand = torch.empty((nq, kp, 3), dtype=torch.int64, device=dev)
    flat = torch.empty((world * nq, kp, 3), dtype=torch.int64, device=dev)
    od = torch.empty((nq, kp), dtype=torch.float32, device=dev); orow = torch.empty((nq, kp), dtype=torch.int64, device=dev)
    om = torch.empty((nq, kp), dtype=torch.int32, device=dev); oc = torch.empty((nq,), dtype=torch.int32, device=dev)

    def step():
        idx.knn_local_dev(Q.data_ptr(), nq, kp, cand.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(flat, cand)
            idx.knn_merge_movies_dev(flat.data_ptr(), world, nq, kp, kp, od.data_ptr(), orow.data_ptr(), om.data_ptr(), oc.data_ptr())
        else:
            idx.knn_merge_movies_dev(cand.data_ptr(), 1, nq, kp, kp, od.data_ptr(), orow.data_ptr(), om.data_ptr(), oc.data_ptr())

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    idx.set_timing(True); idx.stats_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = idx.stats()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return
    ok = bool((od[: min(8, nq), 0] <= 1e-6).all().item() and (oc == kp).all().item())
    peak = 6546.6
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
    scan_ms = st.scan_ms_total / max(1, st.scan_launches_timed)
    kname, row_bytes = scan_kernel_desc(args, int(st.tc_filter_launches) > 0, 16)
    ach = rows * row_bytes / (scan_ms * 1e-3) / 1e9
    print(json.dumps({"metric": "kNN top-100 queries/sec x corpus chunks (row-sharded, NCCL candidate merge)",
                      "value": nq * args.steps / (ms / 1e3) * world * rows, "unit": "query-chunks/s",
                      "queries_per_s": nq * args.steps / (ms / 1e3), "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": f"{tc_kind(args)} filter + f32 exact re-score", "data": "synthetic",
                      "config": {"workload": "configs[4]: synthetic 100M x 384 fp32 (12.5 M-row shard per GPU), top-100",
                                 "rows_per_gpu": rows, "total_rows": rows * world, "queries_per_step": nq},
                      "roofline": {"bound": "hbm", "kernel": kname, "algorithmic_bytes_per_launch": rows * row_bytes,
                                   "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                                   "avg_launch_ms": scan_ms},
                      "self_hits_found_and_full_k": ok, "gpu_launches": int(st.kernel_launches)}), flush=True)
    idx.close()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "bm25_10k":
            if rank == 0:
                run_bm25_10k(args, local_rank)
        elif args.workload == "knn100m":
            run_knn100m(args, rank, world, local_rank)
        else:
            run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
