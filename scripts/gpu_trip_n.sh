#!/bin/bash
# Multi-GPU trip: the default bench line at N GPUs (torchrun), logs in gpurun_out/.
set -u
N=${1:-2}
TAG=${2:-r02n$N}
shift 2 || true
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_$TAG.json").read().strip().splitlines() if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "replicas_match_single_gpu", "corpora_identical_across_ranks")})
    print("e2e", d["e2e"]["value"] if d.get("e2e") else None)
    print("rowshard", d.get("rowshard"))
    print("knn100m", d.get("knn100m"))
except Exception as e:
    print("no bench line:", e)
PY
