#!/bin/bash
# One gpurun trip: GPU tests, then the default bench line.  Every step under its own timeout; logs in gpurun_out/.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")})
    print("e2e", d["e2e"]["value"] if d.get("e2e") else None)
    print("roofline", {k: d["roofline"].get(k) for k in ("achieved", "peak", "frac", "avg_launch_ms", "tc_fallback_queries")})
    print("weighted", {k: (d.get("weighted") or {}).get(k) for k in ("value", "ms_per_step")})
    for c in d.get("clustered") or []:
        print("clustered", c)
    print("python", d.get("e2e_python"))
    print("text_in", d.get("text_in"))
    print("image_search", d.get("image_search"))
    print("rerank", (d.get("text_in") or {}).get("rerank"))
    print("small", d.get("knn_small_batches"))
    print("knn100m", d.get("knn100m"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench line:", e)
PY
