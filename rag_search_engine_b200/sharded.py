"""Row-sharded hybrid search over the GPUs of one box.

One process per GPU (``torch.distributed``; NCCL over NVLink on the B200 box, gloo in the
CPU tests).  The chunk-embedding matrix is cut by row into contiguous shards aligned to
1024-row vec0 blocks, so the emit-order key (distance, block asc, slot desc) is computable
from the global row alone.  Per step:

  1. every rank scans ITS shard for ALL queries of the batch → local top-K' candidates
     packed as int64 triples {key, rowid, movie_idx} (``rse_knn_local_dev``);
  2. ONE exchange of candidates — the path's only real exchange step: every rank needs the lists of ITS query
     slice only, so this is an all-to-all by query slice (nq/N·K'·24 B to each peer; NCCL), not an all-gather
     of everything to everyone (r01: 8x the bytes and a strided copy afterwards); gloo, which has no all-to-all,
     falls back to all_gather + slice in the CPU tests;
  3. each rank owns a contiguous slice of the query batch: it merges the exchanged lists of
     its slice under the same key, aggregates per movie AFTER the merge (the reference
     aggregates over the global top-K', semantic_search.py:285-317), runs BM25 for the
     slice on its replica of the postings and fuses (``rse_hybrid_run_merged_dev``);
  4. the fused [slice, limit] results are all-gathered so every rank holds the batch.

The compute is behind a tiny backend protocol so the plumbing (shard bounds, query
slices, gather layout, result assembly) is testable on CPU with gloo; the product
backend is librse on the rank's GPU — there is no CPU compute path in this package.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Protocol, Tuple

import torch
import torch.distributed as dist

VEC0_BLOCK = 1024


def shard_bounds(n_rows: int, world: int) -> list[int]:
    """Contiguous row ranges aligned to vec0 blocks: rank r owns bounds[r] .. bounds[r+1].
    Every interior bound is a multiple of 1024 (the library requires an aligned pos_base);
    the trailing partial block belongs to the last non-empty range."""
    n_blocks = (n_rows + VEC0_BLOCK - 1) // VEC0_BLOCK
    per, rem = divmod(n_blocks, world)
    floor_rows = (n_rows // VEC0_BLOCK) * VEC0_BLOCK
    bounds = [0]
    blocks = 0
    for r in range(world):
        blocks += per + (1 if r < rem else 0)
        bounds.append(min(blocks * VEC0_BLOCK, floor_rows))
    bounds[-1] = n_rows
    return bounds


def query_slices(nq: int, world: int) -> list[int]:
    per, rem = divmod(nq, world)
    out = [0]
    for r in range(world):
        out.append(out[-1] + per + (1 if r < rem else 0))
    return out


class CandidateExchange:
    """[nq, K', 3] local candidates of ALL queries  ->  [world, ns, K', 3]: every shard's candidates for THIS
    rank's query slice.  NCCL: one all_to_all_single with per-slice split sizes.  Backends without all-to-all
    (gloo): all_gather + slice."""

    def __init__(self, nq: int, world: int, rank: int, device, group=None):
        self.nq, self.world, self.rank, self.device, self.group = nq, world, rank, device, group
        self.slices = query_slices(nq, world)
        self.lo, self.hi = self.slices[rank], self.slices[rank + 1]
        self.max_slice = max(self.slices[r + 1] - self.slices[r] for r in range(world))
        self.kind = "none (1 rank)"
        if world > 1:
            backend = dist.get_backend(group)
            self.kind = "all_to_all (query slices)" if backend == "nccl" else f"all_gather + slice ({backend} has no all-to-all)"
        self._buf = None
        self._flat = None

    def exchange(self, cand: torch.Tensor) -> torch.Tensor:
        nq, kp, three = cand.shape
        ns = self.hi - self.lo
        if self.world == 1:
            return cand.unsqueeze(0)
        if self._buf is None or self._buf.shape[2] != kp:
            self._buf = torch.empty((self.world, max(ns, 1), kp, three), dtype=cand.dtype, device=cand.device)
            self._flat = None
        out = self._buf
        if self.kind.startswith("all_to_all"):
            per = kp * three
            in_splits = [(self.slices[r + 1] - self.slices[r]) * per for r in range(self.world)]
            out_splits = [ns * per] * self.world
            dist.all_to_all_single(out.view(-1)[: self.world * ns * per], cand.view(-1), out_splits, in_splits,
                                   group=self.group)
            return out[:, :ns] if ns > 0 else out[:, :0]
        if self._flat is None:
            self._flat = torch.empty((self.world * nq, kp, three), dtype=cand.dtype, device=cand.device)
        dist.all_gather_into_tensor(self._flat, cand, group=self.group)
        gathered = self._flat.view((self.world,) + tuple(cand.shape))
        if ns > 0:
            out[:, :ns].copy_(gathered[:, self.lo:self.hi])
        return out[:, :ns]


class ShardBackend(Protocol):
    device: torch.device

    def knn_local(self, q_all: torch.Tensor, kprime: int, flag: torch.Tensor | None = None) -> torch.Tensor:
        """[nq, dim] → packed candidates [nq, kprime, 3] int64 on ``device``.  ``flag`` (int32[1], device): do not
        wait for the host; add the number of queries that could not be finished to it.  None: blocking, exact."""

    def stage_slice(self, lo: int, hi: int) -> None:
        """Make queries lo..hi (tokens) the staged batch."""

    def fuse_merged(self, gathered_slice: torch.Tensor, mode: int, param: float, limit: int,
                    knn_multiplier: int, out=None) -> Tuple[torch.Tensor, ...]:
        """[world, nslice, kprime, 3] → (id i64, score f64, a f64, b f64 [nslice, limit], count i32 [nslice])."""


@dataclass
class HybridBatchResult:
    ids: torch.Tensor      # [nq, limit] int64
    score: torch.Tensor    # [nq, limit] float64
    a: torch.Tensor        # rrf: bm25_rank (-1 = None) / weighted: bm25_norm
    b: torch.Tensor        # rrf: sem_rank / weighted: sem_norm
    count: torch.Tensor    # [nq] int32
    flagged: torch.Tensor | None = None   # device scalar: queries (over all ranks) the step could not finish; 0 = exact


class ShardedHybrid:
    """The multi-GPU hybrid step.  ``group`` is a torch.distributed process group (None = default)."""

    def __init__(self, backend: ShardBackend, nq: int, group=None):
        self.backend = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nq = nq
        self.ex = CandidateExchange(nq, self.world, self.rank, backend.device, group)
        self.exchange_kind = self.ex.kind
        self.slices = self.ex.slices
        self.lo, self.hi = self.ex.lo, self.ex.hi
        self.max_slice = self.ex.max_slice
        self.flagged_steps = 0
        backend.stage_slice(self.lo, self.hi)

    def _buffers(self, limit: int, dev):
        if getattr(self, "_buf_key", None) != limit:
            pad = self.max_slice
            # one packed result block per rank: planes id(bits) / score / a / b / count, each [pad, limit] f64, plus
            # one extra row whose first element carries this rank's count of flagged (unfinished) queries
            self._pack = torch.zeros((5, pad + 1, limit), dtype=torch.float64, device=dev)
            self._allp = torch.empty((self.world * 5, pad + 1, limit), dtype=torch.float64, device=dev)
            self._cnt = torch.zeros((pad,), dtype=torch.int32, device=dev)
            self._flag = torch.zeros((1,), dtype=torch.int32, device=dev)
            self._buf_key = limit
        return self._pack, self._allp, self._cnt, self._flag

    def step(self, q_all: torch.Tensor, mode: int, param: float, limit: int, knn_multiplier: int = 10,
             check: bool = True) -> HybridBatchResult:
        """One hybrid step for the whole batch; every rank returns the whole batch's results.  The step itself never
        waits for the host: queries the tensor-core path could not finish are only counted on the device and the
        count travels with the results.  ``check=True`` (default) reads that count (one synchronisation) and
        repeats a flagged step through the blocking path, so the results are always exact; a caller that
        pipelines steps passes ``check=False`` and looks at ``result.flagged`` when it consumes the results."""
        res = self._step(q_all, mode, param, limit, knn_multiplier, defer=True)
        if check and res.flagged is not None and int(res.flagged.item()) > 0:
            self.flagged_steps += 1
            res = self._step(q_all, mode, param, limit, knn_multiplier, defer=False)
        return res

    def _step(self, q_all, mode, param, limit, knn_multiplier, defer: bool) -> HybridBatchResult:
        kprime = max(limit * knn_multiplier, limit)
        dev = self.backend.device
        pack, allp, cnt, flag = self._buffers(limit, dev)
        flag.zero_()
        cand = self.backend.knn_local(q_all, kprime, flag if defer else None)           # [nq, kp, 3]
        ns = self.hi - self.lo
        if self.world == 1:
            i, s, a, b, c = self.backend.fuse_merged(cand.unsqueeze(0), mode, param, limit, knn_multiplier)
            return HybridBatchResult(i, s, a, b, c, flag.to(torch.int64).sum() if defer else None)
        mine = self.ex.exchange(cand)                                                   # the exchange step
        if ns > 0:
            self.backend.fuse_merged(mine, mode, param, limit, knn_multiplier,
                                     out=(pack[0, :ns].view(torch.int64), pack[1, :ns], pack[2, :ns],
                                          pack[3, :ns], cnt[:ns]))
            pack[4, :ns, 0] = cnt[:ns].to(torch.float64)
        pack[4, self.max_slice, 0] = flag[0].to(torch.float64)
        # assemble the batch on every rank (small: nq·limit·40 B): ONE all_gather of the packed block
        dist.all_gather_into_tensor(allp, pack, group=self.group)
        allv = allp.view(self.world, 5, self.max_slice + 1, limit)
        flagged = allv[:, 4, self.max_slice, 0].sum().to(torch.int64) if defer else None
        body = allv[:, :, : self.max_slice]
        if self.nq % self.world == 0:
            full = body.permute(1, 0, 2, 3).reshape(5, self.nq, limit)
        else:
            full = torch.cat([body[r, :, : self.slices[r + 1] - self.slices[r]] for r in range(self.world)], 1)
        return HybridBatchResult(full[0].contiguous().view(torch.int64), full[1], full[2], full[3],
                                 full[4, :, 0].to(torch.int32), flagged)


class LibrseShardBackend:
    """Product backend: this rank's librse handle on its GPU, sharing torch's current stream."""

    def __init__(self, index, q_host, tok_indptr, term_rows, device: torch.device, k1: float = 1.5, b: float = 0.75,
                 tie_mode: int = 0):
        import numpy as np
        self.index = index
        self.device = device
        self.q_host = np.ascontiguousarray(q_host, np.float32)
        self.tok_indptr = np.ascontiguousarray(tok_indptr, np.int32)
        self.term_rows = np.ascontiguousarray(term_rows, np.int32)
        self.k1, self.b, self.tie_mode = k1, b, tie_mode
        index.set_stream(torch.cuda.current_stream(device).cuda_stream)

    def knn_local(self, q_all: torch.Tensor, kprime: int, flag: torch.Tensor | None = None) -> torch.Tensor:
        nq = q_all.shape[0]
        key = (nq, kprime)
        if getattr(self, "_cand_key", None) != key:
            self._cand = torch.empty((nq, kprime, 3), dtype=torch.int64, device=self.device)
            self._cand_key = key
        self.index.set_defer_flags(flag is not None)
        self.index.knn_local_dev(q_all.data_ptr(), nq, kprime, self._cand.data_ptr())
        if flag is not None:
            self.index.knn_flags_dev(flag.data_ptr())
        return self._cand

    def stage_slice(self, lo: int, hi: int) -> None:
        import numpy as np
        if hi <= lo:
            return
        t0, t1 = int(self.tok_indptr[lo]), int(self.tok_indptr[hi])
        ptr = (self.tok_indptr[lo:hi + 1] - t0).astype(np.int32)
        self.index.hybrid_stage(self.q_host[lo:hi], ptr, self.term_rows[t0:t1])

    def fuse_merged(self, gathered_slice, mode, param, limit, knn_multiplier, out=None):
        ns = gathered_slice.shape[1]
        dev = self.device
        if out is None:
            out = (torch.empty((ns, limit), dtype=torch.int64, device=dev),
                   torch.empty((ns, limit), dtype=torch.float64, device=dev),
                   torch.empty((ns, limit), dtype=torch.float64, device=dev),
                   torch.empty((ns, limit), dtype=torch.float64, device=dev),
                   torch.empty((ns,), dtype=torch.int32, device=dev))
        oid, osc, oa, ob, oc = out
        self.index.hybrid_run_merged_dev(mode, param, limit, gathered_slice.data_ptr(), gathered_slice.shape[0],
                                         oid.data_ptr(), osc.data_ptr(), oa.data_ptr(), ob.data_ptr(), oc.data_ptr(),
                                         knn_multiplier=knn_multiplier, k1=self.k1, b=self.b, tie_mode=self.tie_mode)
        return oid, osc, oa, ob, oc


class NcclShardedHybrid:
    """The same step with the collective owned by the LIBRARY (include/rse.h "multi-GPU"): the handle holds the NCCL
    communicator, ``rse_hybrid_sharded_run_dev`` does local top-K' -> all-to-all -> merge -> fusion in one call and
    ``rse_comm_allgather_dev`` assembles the batch on every rank — a binder that only has the C ABI gets the same
    multi-GPU path.  Python only distributes the 128-byte communicator id (``torch.distributed`` here; any
    transport works).  Same result object and the same repeat-when-flagged rule as ``ShardedHybrid``."""

    def __init__(self, index, q_host, tok_indptr, term_rows, device: torch.device, nq: int, k1: float = 1.5,
                 b: float = 0.75, tie_mode: int = 0, group=None):
        import numpy as np
        self.index, self.device, self.nq = index, device, nq
        self.k1, self.b, self.tie_mode = k1, b, tie_mode
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.slices = query_slices(nq, self.world)
        self.lo, self.hi = self.slices[self.rank], self.slices[self.rank + 1]
        self.max_slice = max(self.slices[r + 1] - self.slices[r] for r in range(self.world))
        self.q_host = np.ascontiguousarray(q_host, np.float32)
        self.tok_indptr = np.ascontiguousarray(tok_indptr, np.int32)
        self.term_rows = np.ascontiguousarray(term_rows, np.int32)
        self.flagged_steps = 0
        index.set_stream(torch.cuda.current_stream(device).cuda_stream)
        ids = [index.comm_unique_id() if self.rank == 0 else None]
        if self.world > 1:
            dist.broadcast_object_list(ids, src=0, group=group)
        index.comm_init(ids[0], self.world, self.rank)
        n_ranks, _rank, version = index.comm_info()
        self.exchange_kind = f"library-owned NCCL {version}: grouped ncclSend/ncclRecv all-to-all by query slice ({n_ranks} ranks)"
        self.stage_slice()

    def stage_slice(self):
        import numpy as np
        if self.hi <= self.lo:
            return
        t0, t1 = int(self.tok_indptr[self.lo]), int(self.tok_indptr[self.hi])
        ptr = (self.tok_indptr[self.lo:self.hi + 1] - t0).astype(np.int32)
        self.index.hybrid_stage(self.q_host[self.lo:self.hi], ptr, self.term_rows[t0:t1])

    def _buffers(self, limit: int):
        if getattr(self, "_buf_key", None) != limit:
            pad, dev = self.max_slice, self.device
            self._pack = torch.zeros((5, pad + 1, limit), dtype=torch.float64, device=dev)
            self._allp = torch.empty((self.world * 5, pad + 1, limit), dtype=torch.float64, device=dev)
            self._cnt = torch.zeros((pad,), dtype=torch.int32, device=dev)
            self._flag = torch.zeros((1,), dtype=torch.int32, device=dev)
            self._buf_key = limit
        return self._pack, self._allp, self._cnt, self._flag

    def step(self, q_all: torch.Tensor, mode: int, param: float, limit: int, knn_multiplier: int = 10,
             check: bool = True) -> HybridBatchResult:
        res = self._step(q_all, mode, param, limit, knn_multiplier, defer=True)
        if check and int(res.flagged.item()) > 0:
            self.flagged_steps += 1
            res = self._step(q_all, mode, param, limit, knn_multiplier, defer=False)
        return res

    def _step(self, q_all, mode, param, limit, knn_multiplier, defer: bool) -> HybridBatchResult:
        pack, allp, cnt, flag = self._buffers(limit)
        ns = self.hi - self.lo
        flag.zero_()
        self.index.hybrid_sharded_run_dev(mode, param, limit, q_all.data_ptr(), q_all.shape[0],
                                          pack[0].data_ptr(), pack[1].data_ptr(), pack[2].data_ptr(), pack[3].data_ptr(),
                                          cnt.data_ptr(), flag.data_ptr() if defer else 0,
                                          knn_multiplier=knn_multiplier, k1=self.k1, b=self.b, tie_mode=self.tie_mode)
        if ns > 0:
            pack[4, :ns, 0] = cnt[:ns].to(torch.float64)
        pack[4, self.max_slice, 0] = flag[0].to(torch.float64)
        if self.world > 1:
            self.index.comm_allgather_dev(pack.data_ptr(), allp.data_ptr(), pack.numel() * 8)
        else:
            allp.copy_(pack)
        allv = allp.view(self.world, 5, self.max_slice + 1, limit)
        flagged = allv[:, 4, self.max_slice, 0].sum().to(torch.int64)
        body = allv[:, :, : self.max_slice]
        if self.nq % self.world == 0:
            full = body.permute(1, 0, 2, 3).reshape(5, self.nq, limit)
        else:
            full = torch.cat([body[r, :, : self.slices[r + 1] - self.slices[r]] for r in range(self.world)], 1)
        return HybridBatchResult(full[0].contiguous().view(torch.int64), full[1], full[2], full[3],
                                 full[4, :, 0].to(torch.int32), flagged)
