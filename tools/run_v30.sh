#!/bin/bash
set -x
TAG=${1:-v30}
mkdir -p gpurun_out
timeout 200 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_$TAG.log; grep -a "FAIL\|rror" gpurun_out/gemm_probe_$TAG.log | head; grep -a "^time" gpurun_out/gemm_probe_$TAG.log | head -15
for cfg in "X=1" "CDETR_GEMM_EPI_PAIRS=0"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_c3_${TAG}_${cfg%%=*}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'])"
done
