"""A fixed slice of the randomised parity fuzzer (scripts/gpu_fuzz.py) inside the `-m gpu` suite: random corpus
shapes, norms, duplicates, holes, rowids, K', batch sizes, BM25 parameters and fusion modes — K4 against the exact
scan byte for byte, both against the oracle, BM25 in its three modes against the oracle, the hybrid call against
oracle BM25 + oracle KNN + the restated reference fusion.  The long run's log is profiles/r02_fuzz.txt."""
import importlib.util
import io
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def _fuzz():
    spec = importlib.util.spec_from_file_location("gpu_fuzz", Path(__file__).resolve().parents[1] / "scripts" / "gpu_fuzz.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind,seeds", [("knn", range(0, 10)), ("bm25", range(0, 6)), ("hybrid", range(0, 6))])
def test_fuzz_slice(fresh_index, kind, seeds):
    buf = io.StringIO()
    failures = _fuzz().run(seeds, (kind,), verbose=True, out=buf)
    assert not failures, buf.getvalue()
