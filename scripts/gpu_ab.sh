#!/bin/bash
# A/B of two builds of librse.so on the same box: scripts/gpu_ab.sh <other.so> [rounds]; prints ms/step and the filter pass
set -u
OTHER=$1; ROUNDS=${2:-2}
cp rag_search_engine_b200/csrc/librse.so /tmp/base.so
for r in $(seq $ROUNDS); do
  for v in base other; do
    if [ $v = other ]; then cp $OTHER rag_search_engine_b200/csrc/librse.so; else cp /tmp/base.so rag_search_engine_b200/csrc/librse.so; fi
    RSE_TIMELINE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --no-knn1 2>/tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v', 'ms/step', round(d['ms_per_step'],4), 'filter', round(r['avg_launch_ms'],4), 'e2e', round(d['e2e']['value']))"
    grep "rse timeline" /tmp/ab.err | tail -2
  done
done
cp /tmp/base.so rag_search_engine_b200/csrc/librse.so
