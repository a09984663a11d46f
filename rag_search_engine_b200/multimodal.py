"""MultimodalSearch — drop-in for rag_search_engine.llm.multimodal.MultimodalSearch whose
image -> text search ranks on the B200.

The reference encodes every movie text with CLIP once, then per query image computes the cosine similarity to
ALL text embeddings in numpy and sorts them (llm/multimodal.py:86-95: normalise, ``text_normed @ image_vec``,
``np.argsort(...)[::-1][:top_k]``).  Here the text embeddings live in HBM and the ranking is the same cosine
top-k kernel the semantic path uses (``rse_knn``: exact fp32 cosine distance per row + radix top-k), so a search
is one scan instead of an O(N log N) host sort.  The CLIP model itself (image / text encoder) is NOT rewritten: it
is the pluggable ``model`` (``SentenceTransformer(model_name)`` when sentence-transformers is installed).

Same constructor arguments, method names, result dict keys and error behaviour as the reference.
``similarity`` = 1 - cosine distance; it agrees with the reference's fp32 BLAS dot product to ~1e-7 (different
summation order), so two documents closer than that may swap places — numpy's argsort gives no tie guarantee
either.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np

from . import _lib


class MultimodalSearch:
    def __init__(self, documents: Optional[List[Dict[str, object]]] = None, model_name: str = "clip-ViT-B-32", *,
                 model=None, device: int = 0) -> None:
        if model is None:
            from sentence_transformers import SentenceTransformer  # type: ignore  (llm/multimodal.py:28)
            model = SentenceTransformer(model_name)
        self.model = model
        self.documents: List[Dict[str, object]] = documents or []
        self.texts: List[str] = []
        self.text_embeddings: Optional[np.ndarray] = None
        self._index: Optional[_lib.Index] = None
        if self.documents:
            for doc in self.documents:                                           # llm/multimodal.py:35-38
                title = str(doc.get("title", "")).strip()
                desc = str(doc.get("description", doc.get("document", ""))).strip()
                self.texts.append(f"{title}: {desc}")
            self.text_embeddings = np.asarray(self.model.encode(self.texts, convert_to_numpy=True,
                                                                show_progress_bar=True), np.float32)
            self._index = _lib.Index(device)
            self._index.set_tc_mode(1)          # any embedding width (CLIP: 512): the exact streaming scan
            self._index.load_embeddings(self.text_embeddings)

    def embed_image(self, image_path: str) -> np.ndarray:                        # llm/multimodal.py:46-58
        from PIL import Image  # type: ignore
        image = Image.open(image_path).convert("RGB")
        return self.model.encode([image], convert_to_numpy=True)[0]

    def search_with_vector(self, image_emb, top_k: int = 5) -> List[Dict[str, object]]:
        """The ranking step alone, for a precomputed (e.g. frozen) image embedding."""
        if not self.documents or self.text_embeddings is None:
            raise ValueError("MultimodalSearch was initialized without documents; cannot run search_with_image.")
        top_k = min(int(top_k), len(self.documents))                             # llm/multimodal.py:93
        if top_k <= 0:
            return []
        q = np.ascontiguousarray(image_emb, np.float32).reshape(1, -1)
        if top_k > _lib.RSE_MAX_KPRIME:
            raise ValueError(f"top_k above {_lib.RSE_MAX_KPRIME} is not supported")
        dist, pos, _rowid, _movie, cnt = self._index.knn(q, top_k)
        results: List[Dict[str, object]] = []
        for j in range(int(cnt[0])):
            idx = int(pos[0, j])
            doc = self.documents[idx]
            results.append({"id": doc.get("id", idx), "title": doc.get("title", ""),
                            "description": doc.get("description", doc.get("document", "")),
                            "similarity": float(1.0 - np.float64(dist[0, j]))})
        return results

    def search_with_image(self, image_path: str, top_k: int = 5) -> List[Dict[str, object]]:
        if not self.documents or self.text_embeddings is None:
            raise ValueError("MultimodalSearch was initialized without documents; cannot run search_with_image.")
        return self.search_with_vector(self.embed_image(image_path), top_k=top_k)

    def close(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None
