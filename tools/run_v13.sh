#!/bin/bash
# PDL + single-CTA deep ring: floor probe, correctness probe, bench A/B (PDL on/off, resident-B on).
set -x
TAG=${1:-v13}
mkdir -p gpurun_out
timeout 200 python tools/gemm_floor.py > gpurun_out/gemm_floor_$TAG.log 2>&1; grep -a "us/launch\|launch 3\|rror" gpurun_out/gemm_floor_$TAG.log
timeout 300 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_$TAG.log; grep -a "FAIL\|rror" gpurun_out/gemm_probe_$TAG.log | head -20
for cfg in "X=1" "CDETR_PDL=0" "CDETR_GEMM_RESIDENT=1"; do
  echo "== $cfg"
  env $cfg timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu 2> gpurun_out/bench_$TAG.err | tee gpurun_out/bench_c3_${TAG}_${cfg%%=*}.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step'], d['config']['launch'])"
  tail -2 gpurun_out/bench_$TAG.err
done
