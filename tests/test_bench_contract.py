"""The driver-facing contract of ``bench.py --impl reference`` (the CPU arm runs without a GPU): ONE JSON line
with the keys the driver parses, on a corpus small enough for the CPU suite.  The b200 arm's line is checked on
the GPU box by the driver itself; the keys both arms share are listed once here."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

SHARED_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
               "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def test_reference_arm_prints_one_contract_line():
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--movies", "3000",
                           "--vocab", "5000", "--steps", "2", "--warmup", "1"],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in SHARED_KEYS:
        assert k in d, k
    assert d["metric"] == "hybrid queries/sec" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("configs[3]") and d["config"]["movies"] == 3000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value = queries of the timed steps / their time
    assert abs(d["value"] * d["ms_per_step"] / 1e3 - int(cb["sample"].split()[0])) < 1e-6 * max(1, d["value"])
