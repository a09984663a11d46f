"""SemanticSearch — drop-in for rag_search_engine.utils.semantic_search.SemanticSearch whose
``query_top_k`` runs the vec0 KNN + best-chunk-per-movie aggregation on the B200.

Same constructor, method names, return dicts and error behaviour as the reference
(semantic_search.py:36-65, :211-340, :375-395).  The chunk embeddings are read once from the
vec0 shadow tables of the reference-built SQLite file (store.export_embeddings) and kept in HBM;
chunk text / titles for the ≤k hits still come from SQLite (:262-279, :321-328).
"""
from __future__ import annotations

import sqlite3
from pathlib import Path
from typing import Any, Dict, List, Optional

import numpy as np

from . import runtime, store
from .keyword_search import DEFAULT_DB_PATH
from .textutil import chunk_text


def _default_encoder():
    try:
        from sentence_transformers import SentenceTransformer  # type: ignore
    except Exception:
        return None
    return SentenceTransformer("all-MiniLM-L6-v2")              # semantic_search.py:45


class SemanticSearch:
    def __init__(self, docs_path: Path | str | None = None, db_path: Path | str | None = None,
                 max_chunk_size: int = 3, overlap: int = 1, force: bool = False, *, encoder=None,
                 device: int = 0, fallback_build: bool = False) -> None:
        self.model = encoder if encoder is not None else _default_encoder()
        self.max_chunk_size = max_chunk_size
        self.overlap = overlap
        self.db_path = Path(db_path) if db_path else Path(DEFAULT_DB_PATH)
        self.db_path.parent.mkdir(parents=True, exist_ok=True)
        self.docs_path = Path(docs_path) if docs_path else None
        self.device = device
        self.fallback_build = fallback_build
        if self.docs_path is not None:
            self._build(force)
        self.conn = sqlite3.connect(self.db_path)
        self.conn.execute("PRAGMA journal_mode=WAL")
        self.conn.execute("PRAGMA synchronous=NORMAL")
        store._init_schema(self.conn)                         # a fresh database opens empty, as the reference's does
        self._index = runtime.acquire(self.db_path, device)
        self._closed = False
        self._load()
        self.embedding_dim = (self.model.get_sentence_embedding_dimension() if self.model is not None
                              else self._arr.dim)

    # ------------------------------------------------------------------ build (stays on SQLite)
    def _build(self, force: bool) -> None:
        """Chunking + embedding + vec0 insert are NOT accelerated: use the reference's
        build when it is importable, else freeze embeddings with the given encoder into the same
        tables (store.write_reference_db)."""
        try:
            from rag_search_engine.utils.semantic_search import SemanticSearch as RefSS  # type: ignore
            ref = RefSS(docs_path=self.docs_path, db_path=self.db_path, max_chunk_size=self.max_chunk_size,
                        overlap=self.overlap, force=force)
            ref.close()
            return
        except ImportError:
            pass
        if not self.fallback_build:
            raise RuntimeError("the index build stays on the reference (SQLite + sqlite-vec, north_star) and the "
                               "reference package is not importable here; pass fallback_build=True to freeze the "
                               "embeddings into a GPU-ONLY database (vec0 shadow tables without the virtual table: "
                               "sqlite-vec / the reference cannot open it)")
        if self.model is None:
            raise RuntimeError("SemanticSearch build needs an encoder (sentence-transformers is not installed): "
                               "pass encoder=<object with .encode(texts)>")
        import json
        data = json.loads(Path(self.docs_path).read_text(encoding="utf-8"))["movies"]
        conn = sqlite3.connect(self.db_path)
        try:
            store._init_schema(conn)
            cur = conn.cursor()
            (n_movies,) = cur.execute("SELECT COUNT(*) FROM movies").fetchone()
            if n_movies != len(data):
                cur.execute("DELETE FROM movies")
                cur.executemany("INSERT INTO movies(id, title, description) VALUES (?, ?, ?)",
                                [(int(d["id"]), d["title"], d["description"]) for d in data])
                conn.commit()
            (n_chunks,) = cur.execute("SELECT COUNT(*) FROM chunks").fetchone()
            params = cur.execute("SELECT DISTINCT max_chunk_size, overlap FROM chunks").fetchall()
            in_sync = (n_chunks > 0 and len(params) == 1 and params[0] == (self.max_chunk_size, self.overlap)
                       and store._table_exists(conn, f"{store.VEC_TABLE}_chunks"))          # :145-154
            if force or not in_sync:
                movies = cur.execute("SELECT id, title, description FROM movies ORDER BY id").fetchall()
                rows, texts = store.plan_chunks(movies, self.max_chunk_size, self.overlap)
                cur.execute("DELETE FROM chunks")
                cur.executemany("INSERT INTO chunks(id, movie_id, chunk_index, max_chunk_size, overlap) "
                                "VALUES (?,?,?,?,?)", rows)
                emb = np.asarray(self.model.encode(texts), np.float32)
                store.write_vec0_shadow(conn, np.array([r[0] for r in rows], np.int64), emb)
                conn.commit()
        finally:
            conn.close()

    def _load(self) -> None:
        reg = runtime.parts(self.db_path, self.device)
        full = store._db_fingerprint(self.conn, self.db_path)
        fp = store.part_fingerprint(full, "emb")
        if "emb" not in reg or reg.get("emb_fp") != fp:          # first open, or re-embedded under a live handle
            arr = store.load_or_export(self.conn, self.db_path, "emb", fingerprint=full)
            if arr.emb.shape[0] > 0:
                self._index.load_embeddings(arr.emb, valid=arr.valid, rowid=arr.rowid, movie_idx=arr.movie_idx)
            else:
                # no chunk embeddings (keyword-only / fresh database): query_top_k returns [] like the reference;
                # the handle still needs a dimension for the batch entry points' reshape
                self._index.dim = int(arr.dim) or (int(self.model.get_sentence_embedding_dimension())
                                                   if self.model is not None else 384)
                self._index.n_rows = 0
            reg["emb"], reg["emb_fp"] = arr, fp
            reg.pop("ids", None)

    @property
    def _arr(self) -> store.EmbArrays:
        own = self.__dict__.get("_arr_own")                      # set directly by from_loaded / tests
        return own if own is not None else runtime.parts(self.db_path, self.device)["emb"]

    @_arr.setter
    def _arr(self, value) -> None:
        self.__dict__["_arr_own"] = value

    # ------------------------------------------------------------------ BaseSearchDB helpers
    def count_movies(self) -> int:
        (count,) = self.conn.execute("SELECT COUNT(*) FROM movies").fetchone()
        return count

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self.conn.close()
            runtime.release(self.db_path, self.device)

    @classmethod
    def build_from_docs(cls, docs_path, db_path=None, force: bool = False, **kw) -> "SemanticSearch":
        return cls(docs_path=docs_path, db_path=db_path, force=force, **kw)

    @classmethod
    def open_existing(cls, db_path=None, **kw) -> "SemanticSearch":
        return cls(docs_path=None, db_path=db_path, **kw)

    # ------------------------------------------------------------------ embeddings (semantic_search.py:211-222)
    def generate_embedding(self, text):
        texts = [text] if isinstance(text, str) else text
        if any(len(t.strip()) == 0 for t in texts):
            raise ValueError("cannot embed empty text")
        if self.model is None:
            raise RuntimeError("no query encoder (sentence-transformers is not installed): pass encoder=... "
                               "or call query_top_k_vec with frozen query vectors")
        return np.asarray(self.model.encode(texts), dtype=np.float32)

    # ------------------------------------------------------------------ query
    def query_top_k_vec(self, query_vecs, k: int = 5, knn_multiplier: int = 10) -> List[List[Dict[str, Any]]]:
        """Batch entry point on frozen query vectors [B, dim] → one reference-shaped result list per query."""
        Q = np.atleast_2d(np.ascontiguousarray(query_vecs, np.float32))
        if self._arr.emb.shape[0] == 0:                         # no chunk embeddings: [] per query, like the reference
            return [[] for _ in range(Q.shape[0])]
        Q = Q.reshape(-1, self._arr.dim)
        internal_k = max(k * knn_multiplier, k)                 # semantic_search.py:251
        dist, rowid, movie, cnt = self._index.knn_movies(Q, k, internal_k)
        # the two JOINs of semantic_search.py:262-279 for the <= k final hits of every query: ONE lookup per table
        # for the whole batch (a point query per hit costs more than the device call from ~100 hits on)
        movie_ids = self._arr.movie_ids
        live = np.arange(rowid.shape[1])[None, :] < np.asarray(cnt)[:, None]
        chunk_meta = self._fetch("SELECT id, chunk_index, max_chunk_size, overlap FROM chunks WHERE id IN ({})",
                                 np.unique(rowid[live]))
        movie_meta = self._fetch("SELECT id, title, description FROM movies WHERE id IN ({})",
                                 np.unique(movie_ids[movie[live]]))
        out = []
        for q in range(Q.shape[0]):
            hits = []
            for j in range(cnt[q]):
                cid = int(rowid[q, j])
                movie_id = int(movie_ids[movie[q, j]])
                ci, mcs, ov = chunk_meta[cid]
                title, desc = movie_meta[movie_id]
                hits.append({"chunk_id": cid, "distance": float(dist[q, j]),
                             "chunk": chunk_text(title, desc, int(ci), int(mcs), int(ov)),
                             "movie_id": movie_id, "title": title, "description": desc})
            out.append(hits)
        return out

    def _fetch(self, sql: str, ids) -> Dict[int, tuple]:
        """{id: rest of the row} for ``ids`` (int64 array), in slices below SQLite's bound-variable limit."""
        got: Dict[int, tuple] = {}
        cur = self.conn.cursor()
        ids = [int(i) for i in ids]
        for lo in range(0, len(ids), 900):
            part = ids[lo:lo + 900]
            for row in cur.execute(sql.format(",".join("?" * len(part))), part):
                got[row[0]] = row[1:]
        return got

    def query_top_k(self, query_text: str, k: int = 5, knn_multiplier: int = 10) -> List[Dict[str, Any]]:
        """semantic_search.py:225-340."""
        query_vec = self.generate_embedding(query_text)[0]
        return self.query_top_k_vec(query_vec[None, :], k=k, knn_multiplier=knn_multiplier)[0]

    def _reconstruct_chunk_text(self, title, description, chunk_index, max_chunk_size, overlap) -> str:
        return chunk_text(title, description, chunk_index, max_chunk_size, overlap)

    # ------------------------------------------------------------------ verify (semantic_search.py:368-395)
    def verify_model(self) -> None:
        print(f"Model loaded: {self.model}")
        print(f"Max sequence length: {getattr(self.model, 'max_seq_length', None)}")
        print(f"Embedding dim: {self.embedding_dim}")

    def verify_db(self) -> None:
        cur = self.conn.cursor()
        (movie_count,) = cur.execute("SELECT COUNT(*) FROM movies").fetchone()
        (chunk_count,) = cur.execute("SELECT COUNT(*) FROM chunks").fetchone()
        vec_count = int(self._arr.valid.sum()) if self._arr.valid is not None else int(self._arr.emb.shape[0])
        print(f"Vector DB path: {self.db_path}")
        print("sqlite-vec version: (not loaded: vec0 shadow tables read directly; KNN runs in librse on the GPU)")
        print(f"Movies count:            {movie_count}")
        print(f"Chunks table count:      {chunk_count}")
        print(f"Embeddings (vec0) count: {vec_count}")
        print(f"Embedding dim:           {self.embedding_dim}")
