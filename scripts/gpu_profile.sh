#!/bin/bash
# ncu evidence for profiles/: launch list of the default step + full captures of the filter, BM25 and encoder GEMM kernels.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-knn1 --no-e2e"
timeout 300 python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py $ARGS > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tc3_kernel -s 10 -c 1 -f -o gpurun_out/filter_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_f_$TAG.log 2>&1; echo "filter capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bm25_fx_kernel -s 3 -c 1 -f -o gpurun_out/bm25fx_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_b_$TAG.log 2>&1; echo "bm25 capture rc=$?"
timeout 300 python scripts/enc_run.py > gpurun_out/enc_plain_$TAG.log 2>&1 || { echo "encoder run failed"; tail -5 gpurun_out/enc_plain_$TAG.log; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enc_gemm_tc_kernel -s 48 -c 4 -f -o gpurun_out/encgemm_$TAG \
  python scripts/enc_run.py > gpurun_out/ncu_e_$TAG.log 2>&1; echo "encoder gemm capture rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_enc_$TAG.csv \
  python scripts/enc_run.py > gpurun_out/ncu_el_$TAG.log 2>&1; echo "encoder launch list rc=$?"
ls -la gpurun_out/*$TAG* | head -20
