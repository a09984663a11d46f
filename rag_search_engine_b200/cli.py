"""The reference CLI (`rag-search …`) with the GPU classes injected.

The reference's handlers look the three classes up on the `cli` module at call time
(rag_search_engine/cli/cli.py:5-7, :262-271, :320-445) and its own tests swap them the same way
(tests/test_cli.py:64-65, :92, :117, :141), so the drop-in is two `setattr`s: nothing of the CLI is
re-implemented here (`image_search` gets the MultimodalSearch mirror the same way).  Requires the reference package
to be importable.

    python -m rag_search_engine_b200.cli key_search "toy" --limit 5
"""
from __future__ import annotations

import sys


def inject():
    try:
        import rag_search_engine.cli.cli as ref_cli                     # type: ignore
        import rag_search_engine.utils.hybrid_search as ref_hybrid      # type: ignore
    except Exception as e:  # pragma: no cover - depends on the environment
        raise RuntimeError("the reference package `rag_search_engine` is not importable; "
                           "use the classes of rag_search_engine_b200 directly") from e
    from . import HybridSearch, KeywordSearch, SemanticSearch
    ref_cli.KeywordSearch = KeywordSearch
    ref_cli.SemanticSearch = SemanticSearch
    ref_cli.HybridSearch = HybridSearch
    # code that constructs the reference's own HybridSearch still gets GPU retrievers
    ref_hybrid.KeywordSearch = KeywordSearch
    ref_hybrid.SemanticSearch = SemanticSearch
    # `rag-search image_search` builds MultimodalSearch by module-global name inside image_search_command
    # (llm/multimodal.py:127-150): the mirror keeps the CLIP model and ranks on the GPU
    try:
        import rag_search_engine.llm.multimodal as ref_mm               # type: ignore
        from .multimodal import MultimodalSearch
        ref_mm.MultimodalSearch = MultimodalSearch
    except Exception:                                                   # CLIP / PIL not installed: leave the command as it is
        pass
    return ref_cli


def main(argv=None) -> int:
    ref_cli = inject()
    if argv is not None:
        sys.argv = [sys.argv[0], *argv]
    ref_cli.main()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
