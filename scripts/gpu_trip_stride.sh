#!/bin/bash
# sweep of the probe's sample stride (RSE_TC_SAMPLE_STRIDE) on the default bench workload + clustered corpora
set -u
mkdir -p gpurun_out
for S in ${STRIDES:-32 16 64 32}; do
  RSE_TC_SAMPLE_STRIDE=$S timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn100m --no-e2e > gpurun_out/bench_stride$S.json 2> gpurun_out/bench_stride$S.err
  echo "stride=$S rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_stride$S.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("stride $S: ms/step", round(d["ms_per_step"], 4), "filter ms", round(r["avg_launch_ms"], 4), "surv", r["survivors_per_query"]["p50"], r["survivors_per_query"]["p99"], "2nd", r["tc_second_chance_queries"], "fb", r["tc_fallback_queries"], r["tc_equals_exact_scan"])
for c in d.get("clustered") or []:
    print("   ", c["corpus"], round(c["ms_per_step"], 4), "2nd", c["tc_second_chance_queries"], "fb", c["tc_fallback_queries"], "surv", c["survivors_per_query"]["p50"], c["survivors_per_query"]["p99"], c["tc_equals_exact_scan"]["identical_rows_order_and_distances"])
print("    k1024", d["knn_batch1024"]["call_ms_host_buffers"], "weighted", d["weighted"]["ms_per_step"])
PY
done
