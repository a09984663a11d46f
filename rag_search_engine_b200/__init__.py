"""rag_search_engine_b200 — B200-native retrieval core for the query path of
JWSch4fer/rag-search-engine (BM25 + vec0 KNN + per-movie aggregation + weighted/RRF fusion).

Layout
  csrc/            hand-written sm_100a kernels + the C-ABI (librse.so, include/rse.h)
  _lib.py          ctypes binding of the C-ABI (fails loudly without the library / a GPU)
  keyword_search.py, semantic_search.py, hybrid_search.py
                   drop-in mirrors of the reference classes (same names, signatures, dicts)
  store.py         load-time exporter from the reference's SQLite file (+ writer of that format)
  encoder.py       GpuSentenceEncoder / GpuCrossEncoder: the reference's MiniLM query encoder and TinyBERT
                   cross-encoder on the device (csrc/encoder.cuh)
  multimodal.py    MultimodalSearch mirror: CLIP image -> text ranking on the device
  sharded.py       row-sharded multi-GPU drivers (library-owned NCCL communicator; torch.distributed for tests)
  synth.py         seeded synthetic corpora of the movies_600k shape
  cli.py           the reference CLI with these classes injected

Importing the package needs neither a GPU nor the built library; using it does.
"""
from .keyword_search import KeywordSearch          # noqa: F401
from .semantic_search import SemanticSearch        # noqa: F401
from .hybrid_search import HybridSearch            # noqa: F401

__all__ = ["KeywordSearch", "SemanticSearch", "HybridSearch"]
__version__ = "0.1.0"
