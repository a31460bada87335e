#!/bin/bash
# quick check after a kernel change: op probe (every kernel vs torch / oracle) + one short bench line
set -x
TAG=${1:-x}
mkdir -p gpurun_out
timeout 200 python tests/gpu_ops_probe.py > gpurun_out/ops_probe_$TAG.log 2>&1; echo "rc=$?"; grep -c "^OK" gpurun_out/ops_probe_$TAG.log; grep -a "FAIL\|rror" gpurun_out/ops_probe_$TAG.log | head
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
python -c "import sys,json; d=json.loads(open('gpurun_out/bench_c3_$TAG.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['gemm_ms_per_step']); print(d['roofline_attention']['per_kernel'])"; tail -2 gpurun_out/bench_c3_$TAG.err
