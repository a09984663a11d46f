#!/bin/bash
# A/B of an environment knob on the same box: scripts/gpu_ab_env.sh "NAME=VALUE" [rounds]
set -u
KNOB=$1; ROUNDS=${2:-2}
for r in $(seq $ROUNDS); do
  for v in base knob; do
    if [ $v = knob ]; then export $KNOB; else unset ${KNOB%%=*}; fi
    RSE_TIMELINE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-knn1 2>/tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v', 'ms/step', round(d['ms_per_step'],4), 'filter', round(r['avg_launch_ms'],4), 'e2e', round(d['e2e']['value']), 'same', d['e2e'].get('same_results_as_blocking_call'))"
    grep "rse timeline" /tmp/ab.err | tail -2
  done
done
