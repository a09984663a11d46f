"""Run the MiniLM-architecture encoder a few times (target of the ncu capture of enc_gemm_tc_kernel)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rag_search_engine_b200 import _lib                                                    # noqa: E402
from rag_search_engine_b200.encoder import MINILM_L6_CONFIG, GpuSentenceEncoder, random_state_dict  # noqa: E402

idx = _lib.Index(0)
enc = GpuSentenceEncoder(idx, random_state_dict(MINILM_L6_CONFIG, seed=5), MINILM_L6_CONFIG)
if len(sys.argv) > 1:
    idx.encoder_set_mode(enc.slot, int(sys.argv[1]))
rng = np.random.default_rng(1)
ids = [[101] + rng.integers(1000, 30000, int(L) - 2).tolist() + [102] for L in rng.integers(4, 17, 256)]
import time
for _ in range(3):
    out = enc.encode_ids(ids)
t0 = time.perf_counter()
for _ in range(20):
    out = enc.encode_ids(ids)
print(out.shape, float(np.linalg.norm(out[0])), f"{(time.perf_counter() - t0) / 20 * 1e3:.3f} ms per encode call (host buffers, 256 queries)")
idx.close()
