#!/bin/bash
# TMA-epilogue validation: GEMM probe (correctness + timings), A/B bench against the generic epilogue, GPU suite.
set -x
TAG=${1:-v12}
mkdir -p gpurun_out
timeout 300 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_$TAG.log; grep -a "FAIL\|rror" gpurun_out/gemm_probe_$TAG.log | head -20; grep -a "^time" gpurun_out/gemm_probe_$TAG.log
CDETR_GEMM_TMA_EPI=0 timeout 300 python tests/gpu_gemm_probe.py 2>&1 | grep -a "^time" | head -9
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
head -c 300 gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
