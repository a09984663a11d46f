#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r02e.log
for D in 3 5 8; do
  RSE_TC_SURVIVOR_DIV=$D timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn100m --no-knn1 > gpurun_out/bench_div$D.json 2> gpurun_out/bench_div$D.err
  echo "div=$D rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_div$D.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("div $D: ms/step", round(d["ms_per_step"], 4), "filter ms", round(r["avg_launch_ms"], 4), "surv", r["survivors_per_query"]["p50"], r["survivors_per_query"]["p99"])
for c in d.get("clustered") or []:
    print("   ", c["corpus"], round(c["ms_per_step"], 4), "2nd", c["tc_second_chance_queries"], "fb", c["tc_fallback_queries"], "surv", c["survivors_per_query"]["p50"], c["survivors_per_query"]["p99"])
print("    text_in", d["text_in"]["value"], d["text_in"]["encoder_ms_per_batch"])
PY
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02.json 2> gpurun_out/bench_ref_r02.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_ref_r02.json
