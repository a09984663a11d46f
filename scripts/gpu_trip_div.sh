#!/bin/bash
set -u
mkdir -p gpurun_out

for D in ${DIVS:-5 8 12}; do
  RSE_TC_SURVIVOR_DIV=$D timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn100m --no-e2e > gpurun_out/bench_div$D.json 2> gpurun_out/bench_div$D.err
  echo "div=$D rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_div$D.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("div $D: ms/step", round(d["ms_per_step"], 4), "filter ms", round(r["avg_launch_ms"], 4), "surv", r["survivors_per_query"]["p50"], r["survivors_per_query"]["p99"], r["tc_equals_exact_scan"])
for c in d.get("clustered") or []:
    print("   ", c["corpus"], round(c["ms_per_step"], 4), "2nd", c["tc_second_chance_queries"], "fb", c["tc_fallback_queries"], "surv", c["survivors_per_query"]["p50"], c["survivors_per_query"]["p99"], c["tc_equals_exact_scan"]["identical_rows_order_and_distances"])
print("    k1024", d["knn_batch1024"]["call_ms_host_buffers"])
PY
done
