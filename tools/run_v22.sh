#!/bin/bash
set -x
TAG=${1:-v22}
mkdir -p gpurun_out
for cfg in "X=1" "CDETR_PDL_LIGHT=0"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_c3_${TAG}_${cfg%%=*}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['config']['launch'])"
  tail -2 gpurun_out/bench_$TAG.err
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/gpu_ops_probe.py > gpurun_out/ops_probe_$TAG.log 2>&1; grep -c "^OK" gpurun_out/ops_probe_$TAG.log; grep -a "FAIL\|rror" gpurun_out/ops_probe_$TAG.log | head
