#!/bin/bash
# One gpurun trip for the randomised parity fuzzer: log in gpurun_out/fuzz_<tag>.txt
set -u
TAG=${1:-r02}
CASES=${2:-120}
mkdir -p gpurun_out
timeout 1500 python scripts/gpu_fuzz.py --cases $CASES > gpurun_out/fuzz_$TAG.txt 2>&1
echo "fuzz rc=$?"
grep -c "^ok" gpurun_out/fuzz_$TAG.txt
grep "^FAIL\|cases," gpurun_out/fuzz_$TAG.txt | head -40
tail -3 gpurun_out/fuzz_$TAG.txt
