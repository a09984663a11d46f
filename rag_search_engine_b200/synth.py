"""Seeded synthetic corpora with the shape of the reference's data.

``movies_600k.json`` is absent from the reference checkout (.MISSING_LARGE_BLOBS), so
the benchmark and the large parity tests run on generated data of the named shape:

  * chunk embeddings: contiguous chunks per movie, title chunk first, movies ascending
    (semantic_search.py:164-188); L2-normalised N(0,1) rows with exact duplicates
    ("same title" → byte-identical embedding → exact distance ties) and 1-ulp
    near-duplicates;
  * BM25 postings: Zipf term draw, doc length ~ max(5, N(120, 40)) plus the
    ``[TITLE_END]`` sentinel that every document holds exactly once
    (keyword_search.py:24,132-133) — one term with df = N, tf = 1, counted in dl;
  * queries: document-frequency-weighted token draws with OOV and duplicate tokens;
    query vectors half perturbed corpus rows, half random unit vectors.

Generation uses torch so the big corpora can be produced directly in HBM; nothing here
is on the timed path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


# ----------------------------------------------------------------------------- embeddings
@dataclass
class SynthEmbeddings:
    emb: torch.Tensor            # [C, dim] fp32 (device of generation)
    movie_of_chunk: torch.Tensor  # [C] int32 dense movie index (non-decreasing)
    movie_ids: np.ndarray        # [M] int64 sparse increasing ids
    n_movies: int


def synth_movie_ids(n_movies: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    gaps = rng.integers(1, 4, size=n_movies, dtype=np.int64)
    return np.cumsum(gaps) + 10          # sparse increasing ints, not 0..M-1


def synth_embeddings(n_movies: int, seed: int = 1234, dim: int = 384, mean_desc_chunks: float = 7.0,
                     dup_frac: float = 0.02, neardup_frac: float = 0.005, device: str = "cpu",
                     chunk_rows: int = 1 << 20, distribution: str = "isotropic", n_centres: int = 20_000,
                     spread: float = 0.6) -> SynthEmbeddings:
    """``distribution="isotropic"``: L2-normalised N(0, I) rows — the K'-th neighbour of a query sits ~3.8 sigma
    out in a thin tail (the friendliest case for a threshold filter).  ``"clustered"``: every movie draws one of
    ``n_centres`` unit centres (its "genre") and each of its chunks is normalise(centre + spread * unit noise) —
    sentence-embedding corpora are clustered like this, so the neighbourhood of a query is DENSE: the top-K' lie
    inside a cluster of rows_per_centre rows whose cosines to the query differ by a few 1e-3."""
    if distribution not in ("isotropic", "clustered"):
        raise ValueError("distribution must be 'isotropic' or 'clustered'")
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_desc = torch.poisson(torch.full((n_movies,), float(mean_desc_chunks)), generator=g).clamp_(min=1)
    per_movie = (1 + n_desc).to(torch.int64)
    C = int(per_movie.sum())
    movie_of_chunk = torch.repeat_interleave(torch.arange(n_movies, dtype=torch.int32), per_movie)
    dev = torch.device(device)
    emb = torch.empty((C, dim), dtype=torch.float32, device=dev)
    gd = torch.Generator(device=dev).manual_seed(seed + 1)
    centres = centre_of_chunk = None
    if distribution == "clustered":
        centres = torch.randn((n_centres, dim), generator=gd, device=dev, dtype=torch.float32)
        centres /= centres.norm(dim=1, keepdim=True)
        centre_of_movie = torch.randint(0, n_centres, (n_movies,), generator=g)
        centre_of_chunk = torch.repeat_interleave(centre_of_movie, per_movie).to(dev)
    for s in range(0, C, chunk_rows):
        e = min(C, s + chunk_rows)
        x = torch.randn((e - s, dim), generator=gd, device=dev, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        if centres is not None:
            x = centres[centre_of_chunk[s:e]] + spread * x
            x /= x.norm(dim=1, keepdim=True)
        emb[s:e] = x
    # exact duplicates and 1-ulp near-duplicates
    n_dup = int(C * dup_frac)
    n_near = int(C * neardup_frac)
    if C > 1 and (n_dup or n_near):
        idx = torch.randint(0, C, (2, n_dup + n_near), generator=g)
        col_all = torch.randint(0, dim, (n_near,), generator=g) if n_near else None

        def first_occurrence(d: torch.Tensor) -> torch.Tensor:
            """Positions of the first occurrence of every destination row.  An indexed assignment with repeated
            destinations is nondeterministic on CUDA (whichever thread writes last wins), which made the corpus
            differ between runs and between the GPUs of a multi-GPU bench in ~1000 rows."""
            _, first = np.unique(d.numpy(), return_index=True)
            return torch.from_numpy(np.sort(first))

        keep = first_occurrence(idx[1, :n_dup])
        src = idx[0, :n_dup][keep].to(dev)
        dst = idx[1, :n_dup][keep].to(dev)
        emb[dst] = emb[src]
        if n_near:
            keep = first_occurrence(idx[1, n_dup:])
            src = idx[0, n_dup:][keep].to(dev)
            dst = idx[1, n_dup:][keep].to(dev)
            col = col_all[keep].to(dev)
            rows = emb[src].clone()
            ar = torch.arange(rows.shape[0], device=dev)
            v = rows[ar, col]
            rows[ar, col] = torch.nextafter(v, torch.full_like(v, 2.0))
            emb[dst] = rows
    return SynthEmbeddings(emb=emb, movie_of_chunk=movie_of_chunk.to(dev), movie_ids=synth_movie_ids(n_movies, seed),
                           n_movies=n_movies)


def synth_query_vectors(emb: torch.Tensor, nq: int, seed: int = 99, sigma: float = 0.05) -> torch.Tensor:
    dev = emb.device
    g = torch.Generator(device="cpu").manual_seed(seed)
    C, dim = emb.shape
    n_pert = nq // 2
    rows = torch.randint(0, C, (n_pert,), generator=g).to(dev)
    gd = torch.Generator(device=dev).manual_seed(seed + 7)
    q1 = emb[rows] + sigma * torch.randn((n_pert, dim), generator=gd, device=dev)
    q2 = torch.randn((nq - n_pert, dim), generator=gd, device=dev)
    q = torch.cat([q1, q2], 0)
    q /= q.norm(dim=1, keepdim=True)
    perm = torch.randperm(nq, generator=g).to(dev)
    return q[perm].contiguous().to(torch.float32)


# ----------------------------------------------------------------------------- BM25
@dataclass
class SynthBm25:
    indptr: np.ndarray    # [T+1] int64
    doc_idx: np.ndarray   # [P] uint32, ascending within a term
    tf: np.ndarray        # [P] uint32
    df: np.ndarray        # [T] int64
    dl: np.ndarray        # [M] uint32 (sentinel counted)
    n_movies: int
    avgdl: float
    sentinel_term: int    # CSR row of "[TITLE_END]"


def synth_bm25(n_docs: int, vocab: int, seed: int = 1234, mean_len: float = 120.0, sd_len: float = 40.0,
               zipf_s: float = 1.07, device: str = "cpu") -> SynthBm25:
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    lens = torch.normal(mean_len, sd_len, (n_docs,), generator=g, device=dev).round_().clamp_(min=5).to(torch.int64)
    total = int(lens.sum())
    doc_of_tok = torch.repeat_interleave(torch.arange(n_docs, device=dev, dtype=torch.int64), lens)
    w = 1.0 / torch.arange(1, vocab + 1, device=dev, dtype=torch.float64) ** zipf_s
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    u = torch.rand((total,), generator=g, device=dev, dtype=torch.float64)
    term = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
    key = term * n_docs + doc_of_tok
    del term, doc_of_tok, u
    key, _ = torch.sort(key)
    uniq, counts = torch.unique_consecutive(key, return_counts=True)
    del key
    t = uniq // n_docs
    d = uniq - t * n_docs
    # sentinel "[TITLE_END]": last term row, every doc once
    t = torch.cat([t, torch.full((n_docs,), vocab, device=dev, dtype=torch.int64)])
    d = torch.cat([d, torch.arange(n_docs, device=dev, dtype=torch.int64)])
    counts = torch.cat([counts, torch.ones(n_docs, device=dev, dtype=torch.int64)])
    per_term = torch.bincount(t, minlength=vocab + 1)
    indptr = torch.zeros(vocab + 2, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(per_term, 0)
    dl = (lens + 1).to(torch.int64)
    avgdl = float(dl.sum().item()) / float(n_docs)          # == SQLite AVG(length)
    return SynthBm25(indptr=indptr.cpu().numpy(), doc_idx=d.to(torch.int64).cpu().numpy().astype(np.uint32),
                     tf=counts.cpu().numpy().astype(np.uint32), df=per_term.cpu().numpy().astype(np.int64),
                     dl=dl.cpu().numpy().astype(np.uint32), n_movies=n_docs, avgdl=avgdl, sentinel_term=vocab)


def synth_token_queries(bm: SynthBm25, nq: int, seed: int = 99, min_tok: int = 2, max_tok: int = 6,
                        oov_frac: float = 0.05, dup_frac: float = 0.03):
    """Returns (tok_indptr int32[nq+1], term_rows int32[ntok]); -1 marks an OOV token."""
    rng = np.random.default_rng(seed)
    df = bm.df.astype(np.float64).copy()
    df[bm.sentinel_term] = 0.0              # "[TITLE_END]" cannot come out of preprocess (utils.py:116)
    p = df / df.sum()
    cdf = np.cumsum(p)
    ntoks = rng.integers(min_tok, max_tok + 1, size=nq)
    tok_indptr = np.zeros(nq + 1, np.int32)
    tok_indptr[1:] = np.cumsum(ntoks)
    total = int(tok_indptr[-1])
    terms = np.searchsorted(cdf, rng.random(total)).clip(max=len(df) - 1).astype(np.int32)
    r = rng.random(total)
    terms[r < oov_frac] = -1
    dup = np.nonzero((r >= oov_frac) & (r < oov_frac + dup_frac))[0]
    for i in dup:                            # duplicate the query's first token
        q = np.searchsorted(tok_indptr, i, side="right") - 1
        terms[i] = terms[tok_indptr[q]]
    return tok_indptr, terms
