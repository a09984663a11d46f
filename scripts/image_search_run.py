"""bench.py's image_search measurement alone (600 k x 512 text embeddings, top-5 per query, host buffers)."""
import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402

ns = argparse.Namespace(movies=600_000)
print(json.dumps(bench.measure_image_search(ns, 0)))
