"""KeywordSearch — drop-in for rag_search_engine.utils.keyword_search.KeywordSearch whose
``search`` runs on the B200 (librse BM25 kernels) instead of the Python/SQLite loop.

Same constructor, methods, return dicts and empty/error behaviour as the reference
(keyword_search.py:26-40, :180-267, :270-307); SQLite remains the storage layer: the index is
exported once at open time (store.export_bm25) and titles/descriptions are still fetched from
the ``movies`` table for the ≤k final hits (:253-265).
"""
from __future__ import annotations

import sqlite3
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import runtime, store
from .textutil import Tokenizer, default_tokenizer

try:  # reference default (config.py:15) when the reference package is importable
    from rag_search_engine.config import DEFAULT_DB_PATH  # type: ignore
except Exception:  # pragma: no cover
    DEFAULT_DB_PATH = Path("rag_search_engine") / "cache" / "movies.db"


class KeywordSearch:
    TITLE_END_TOKEN = store.TITLE_END_TOKEN                 # keyword_search.py:24

    def __init__(self, docs_path: Path | str | None, db_path: Path | str | None = None, force: bool = False, *,
                 tokenizer: Optional[Tokenizer] = None, device: int = 0) -> None:
        self.db_path = Path(db_path) if db_path else Path(DEFAULT_DB_PATH)
        self.db_path.parent.mkdir(parents=True, exist_ok=True)
        self.docs_path: Optional[Path] = Path(docs_path) if docs_path else None
        self.force = force
        self.device = device
        self._tokenizer = tokenizer
        if self.docs_path is not None:
            self._build(force)
        self.conn = sqlite3.connect(self.db_path)            # basesearch_db.py:40-42
        self.conn.execute("PRAGMA journal_mode=WAL")
        self.conn.execute("PRAGMA synchronous=NORMAL")
        store._init_schema(self.conn)
        self._index = runtime.acquire(self.db_path, device)
        self._closed = False
        self._load()

    # ------------------------------------------------------------------ build (stays on SQLite)
    def _build(self, force: bool) -> None:
        """Index build is NOT accelerated: delegate to the reference's own build when
        the reference package is importable, else write the same tables with the configured tokenizer."""
        try:
            from rag_search_engine.utils.keyword_search import KeywordSearch as RefKS  # type: ignore
            ref = RefKS(docs_path=self.docs_path, db_path=self.db_path, force=force)
            ref.close()
            return
        except ImportError:
            pass
        import json
        data = json.loads(Path(self.docs_path).read_text(encoding="utf-8"))["movies"]
        conn = sqlite3.connect(self.db_path)
        try:
            store._init_schema(conn)
            (n_dl,) = conn.execute("SELECT COUNT(*) FROM doclen").fetchone()
            if force or n_dl != len(data):                   # keyword_search.py:87-100
                cur = conn.cursor()
                cur.execute("DELETE FROM movies")
                cur.executemany("INSERT INTO movies(id, title, description) VALUES (?, ?, ?)",
                                [(int(d["id"]), d["title"], d["description"]) for d in data])
                conn.commit()
                store.write_keyword_index(conn, data, self._tok())
        finally:
            conn.close()

    def _tok(self) -> Tokenizer:
        if self._tokenizer is None:
            self._tokenizer = default_tokenizer()
        return self._tokenizer

    def _load(self) -> None:
        reg = runtime.parts(self.db_path, self.device)
        full = store._db_fingerprint(self.conn, self.db_path)
        fp = store.part_fingerprint(full, "bm25")
        if "bm25" not in reg or reg.get("bm25_fp") != fp:
            # first open of this database in the process — or the keyword tables changed under a live handle (a
            # rebuild in place): export again and replace the device copy; every object on the handle reads its
            # arrays through the registry (``_arr``), so they all move to the new index together
            arr = store.load_or_export(self.conn, self.db_path, "bm25", fingerprint=full)
            self._index.load_bm25(arr.indptr, arr.doc_idx, arr.tf, arr.df, arr.dl, arr.n_movies, arr.avgdl)
            reg["bm25"], reg["bm25_fp"] = arr, fp
            reg.pop("ids", None)                              # the fused path's id tables follow the arrays

    @property
    def _arr(self) -> store.Bm25Arrays:
        own = self.__dict__.get("_arr_own")                  # set directly by from_loaded / tests
        return own if own is not None else runtime.parts(self.db_path, self.device)["bm25"]

    @_arr.setter
    def _arr(self, value) -> None:
        self.__dict__["_arr_own"] = value

    # ------------------------------------------------------------------ helpers mirrored from BaseSearchDB
    def count_movies(self) -> int:                            # basesearch_db.py:95-99
        (count,) = self.conn.execute("SELECT COUNT(*) FROM movies").fetchone()
        return count

    def close(self) -> None:                                  # basesearch_db.py:101-102
        if not self._closed:
            self._closed = True
            self.conn.close()
            runtime.release(self.db_path, self.device)

    @classmethod
    def build_from_docs(cls, docs_path, db_path=None, force: bool = False, **kw) -> "KeywordSearch":
        return cls(docs_path=docs_path, db_path=db_path, force=force, **kw)

    @classmethod
    def open_existing(cls, db_path=None, **kw) -> "KeywordSearch":
        return cls(docs_path=None, db_path=db_path, **kw)

    # ------------------------------------------------------------------ BM25 search
    def _term_rows(self, token_lists):
        """Query tokens → (tok_indptr int32[nq+1], term_rows int32[ntok]) for rse_bm25 / rse_hybrid: the CSR row of
        every token in query order, duplicates kept, -1 for a term the index does not hold
        (keyword_search.py:205-210).  Accepts a sequence of token lists, or a pre-flattened
        ``(tok_indptr, tokens)`` pair where ``tokens`` is a numpy array of str — that form is resolved with one
        ``searchsorted`` over the sorted vocabulary instead of a dict lookup per token."""
        if isinstance(token_lists, tuple) and len(token_lists) == 2 and isinstance(token_lists[1], np.ndarray):
            tok_indptr, toks = token_lists
            tok_indptr = np.ascontiguousarray(tok_indptr, np.int32)
            if toks.size == 0:
                return tok_indptr, np.zeros(1, np.int32)
            vocab, rows_of = self._sorted_vocab()
            pos = np.searchsorted(vocab, toks)
            pos[pos >= len(vocab)] = 0
            hit = vocab[pos] == toks
            return tok_indptr, np.where(hit, rows_of[pos], -1).astype(np.int32)
        get = self._arr.term_row.get
        lens = np.fromiter((len(t) for t in token_lists), np.int64, count=len(token_lists))
        tok_indptr = np.zeros(len(token_lists) + 1, np.int32)
        np.cumsum(lens, out=tok_indptr[1:])
        n = int(tok_indptr[-1])
        if n == 0:
            return tok_indptr, np.zeros(1, np.int32)
        rows = np.fromiter((get(t, -1) for toks in token_lists for t in toks), np.int32, count=n)
        return tok_indptr, rows

    def _sorted_vocab(self):
        term_row = self._arr.term_row
        sv = getattr(self, "_vocab_sorted", None)
        if sv is None or sv[0] is not term_row:              # (rebuilt when the arrays were replaced)
            terms = np.array(list(term_row.keys()), dtype=np.str_)
            rows = np.fromiter(term_row.values(), np.int32, count=len(terms))
            order = np.argsort(terms, kind="stable")
            sv = self._vocab_sorted = (term_row, terms[order], rows[order])
        return sv[1], sv[2]

    def search_tokens(self, token_lists: Sequence[Sequence[str]], k: int = 10, k1: float = 1.5, b: float = 0.75):
        """Batch entry point on pre-tokenised queries → [(doc_id int64[n], score float64[n])] per query."""
        tok_indptr, rows = self._term_rows(token_lists)
        score, doc, cnt = self._index.bm25(tok_indptr, rows, k, k1, b)
        ids = self._arr.doc_ids
        return [(ids[doc[q, :cnt[q]]], score[q, :cnt[q]]) for q in range(len(token_lists))]

    def search(self, query: str, k: int = 10, k1: float = 1.5, b: float = 0.75) -> List[Dict[str, Any]]:
        """keyword_search.py:180-267 — same results, same dict key order (title, description, id, score)."""
        query_tokens = self._tok()([query])[0]
        if not query_tokens:                                  # :190-191
            return []
        doc_ids, scores = self.search_tokens([query_tokens], k=k, k1=k1, b=b)[0]
        if len(doc_ids) == 0:                                 # :199-200, :246-247
            return []
        ids = [int(d) for d in doc_ids]
        cur = self.conn.cursor()
        placeholders = ",".join("?" for _ in ids)
        cur.execute(f"SELECT id, title, description FROM movies WHERE id IN ({placeholders})", ids)
        meta = {row[0]: (row[1], row[2]) for row in cur.fetchall()}
        results = []
        for doc_id, score in zip(ids, scores):
            title, desc = meta.get(doc_id, ("<missing>", ""))  # :262-263
            results.append({"title": title, "description": desc, "id": doc_id, "score": float(score)})
        return results

    # ------------------------------------------------------------------ verify (keyword_search.py:270-307)
    def verify_db(self) -> None:
        cur = self.conn.cursor()
        counts = {}
        for t in ("movies", "terms", "postings", "doclen"):
            (counts[t],) = cur.execute(f"SELECT COUNT(*) FROM {t}").fetchone()
        (avgdl,) = cur.execute("SELECT AVG(length) FROM doclen").fetchone()
        avgdl = avgdl or 0.0
        print(f"Keyword index DB path:   {self.db_path}")
        print(f"Movies table count:      {counts['movies']}")
        print(f"Terms table count:       {counts['terms']}")
        print(f"Postings table count:    {counts['postings']}")
        print(f"Doclen table count:      {counts['doclen']}")
        print(f"Average doc length:      {avgdl:.2f}")
        if counts["movies"] != counts["doclen"]:
            print("WARNING: movies.count != doclen.count (index may be out of sync)")
        if counts["postings"] == 0:
            print("WARNING: postings table is empty")
        if counts["terms"] == 0:
            print("WARNING: terms table is empty")
