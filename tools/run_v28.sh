#!/bin/bash
set -x
TAG=${1:-v28}
mkdir -p gpurun_out
CDETR_GEMM_EPI_DIRECT=1 timeout 200 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_direct_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_direct_$TAG.log; grep -a "FAIL\|rror" gpurun_out/gemm_probe_direct_$TAG.log | head; grep -a "^time" gpurun_out/gemm_probe_direct_$TAG.log | head -9
timeout 200 python tests/gpu_gemm_probe.py 2>&1 | grep -a "^time" | head -9
for cfg in "CDETR_GEMM_EPI_DIRECT=1" "X=1"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_c3_${TAG}_${cfg%%=*}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'])"
done
