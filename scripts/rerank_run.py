"""bench.py's rerank measurement alone (target of ncu launch lists for the cross-encoder)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from rag_search_engine_b200 import _lib  # noqa: E402

idx = _lib.Index(0)
print(json.dumps(bench.measure_rerank(idx, 10)))
idx.close()
