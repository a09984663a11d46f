"""Row-sharded hybrid search over the GPUs of one box.

One process per GPU (``torch.distributed``; NCCL over NVLink on the B200 box, gloo in the
CPU tests).  The chunk-embedding matrix is cut by row into contiguous shards aligned to
1024-row vec0 blocks, so the emit-order key (distance, block asc, slot desc) is computable
from the global row alone.  Per step:

  1. every rank scans ITS shard for ALL queries of the batch → local top-K' candidates
     packed as int64 triples {key, rowid, movie_idx} (``rse_knn_local_dev``);
  2. ONE ``all_gather`` of nq·K'·24 B per rank — the path's only real exchange step;
  3. each rank owns a contiguous slice of the query batch: it merges the gathered lists of
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


class ShardBackend(Protocol):
    device: torch.device

    def knn_local(self, q_all: torch.Tensor, kprime: int) -> torch.Tensor:
        """[nq, dim] → packed candidates [nq, kprime, 3] int64 on ``device``."""

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


class ShardedHybrid:
    """The multi-GPU hybrid step.  ``group`` is a torch.distributed process group (None = default)."""

    def __init__(self, backend: ShardBackend, nq: int, group=None):
        self.backend = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nq = nq
        self.slices = query_slices(nq, self.world)
        self.lo, self.hi = self.slices[self.rank], self.slices[self.rank + 1]
        self.max_slice = max(self.slices[r + 1] - self.slices[r] for r in range(self.world))
        backend.stage_slice(self.lo, self.hi)

    def _buffers(self, nq_all: int, kprime: int, limit: int, dev):
        key = (nq_all, kprime, limit)
        if getattr(self, "_buf_key", None) != key:
            pad = self.max_slice
            self._flat = torch.empty((self.world * nq_all, kprime, 3), dtype=torch.int64, device=dev)
            self._mine = torch.empty((self.world, max(self.hi - self.lo, 1), kprime, 3), dtype=torch.int64, device=dev)
            # one packed result block per rank: planes id(bits) / score / a / b / count, each [pad, limit] f64
            self._pack = torch.zeros((5, pad, limit), dtype=torch.float64, device=dev)
            self._allp = torch.empty((self.world * 5, pad, limit), dtype=torch.float64, device=dev)
            self._cnt = torch.zeros((pad,), dtype=torch.int32, device=dev)
            self._buf_key = key
        return self._flat, self._mine, self._pack, self._allp, self._cnt

    def step(self, q_all: torch.Tensor, mode: int, param: float, limit: int, knn_multiplier: int = 10) -> HybridBatchResult:
        kprime = max(limit * knn_multiplier, limit)
        dev = self.backend.device
        cand = self.backend.knn_local(q_all, kprime)                                   # [nq, kp, 3]
        ns = self.hi - self.lo
        if self.world == 1:
            i, s, a, b, c = self.backend.fuse_merged(cand.unsqueeze(0), mode, param, limit, knn_multiplier)
            return HybridBatchResult(i, s, a, b, c)
        flat, mine, pack, allp, cnt = self._buffers(cand.shape[0], kprime, limit, dev)
        dist.all_gather_into_tensor(flat, cand, group=self.group)                      # the exchange step
        gathered = flat.view((self.world,) + tuple(cand.shape))
        if ns > 0:
            mine.copy_(gathered[:, self.lo:self.hi])                                   # this rank's query slice, contiguous
            i, s, a, b, c = self.backend.fuse_merged(mine, mode, param, limit, knn_multiplier,
                                                     out=(pack[0, :ns].view(torch.int64), pack[1, :ns], pack[2, :ns],
                                                          pack[3, :ns], cnt[:ns]))
            pack[4, :ns, 0] = cnt[:ns].to(torch.float64)
        # assemble the batch on every rank (small: nq·limit·40 B): ONE all_gather of the packed block
        dist.all_gather_into_tensor(allp, pack, group=self.group)
        allv = allp.view(self.world, 5, self.max_slice, limit)
        if self.nq % self.world == 0:
            full = allv.permute(1, 0, 2, 3).reshape(5, self.nq, limit)
        else:
            full = torch.cat([allv[r, :, : self.slices[r + 1] - self.slices[r]] for r in range(self.world)], 1)
        return HybridBatchResult(full[0].contiguous().view(torch.int64), full[1], full[2], full[3],
                                 full[4, :, 0].to(torch.int32))


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

    def knn_local(self, q_all: torch.Tensor, kprime: int) -> torch.Tensor:
        nq = q_all.shape[0]
        cand = torch.empty((nq, kprime, 3), dtype=torch.int64, device=self.device)
        self.index.knn_local_dev(q_all.data_ptr(), nq, kprime, cand.data_ptr())
        return cand

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
