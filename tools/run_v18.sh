#!/bin/bash
# CTA-pair (cta_group::2) GEMM bring-up: probe with pairs forced on (short timeout: a hang must not eat the box).
set -x
TAG=${1:-v18}
mkdir -p gpurun_out
CDETR_GEMM_PAIR=1 timeout 150 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_pair_$TAG.log 2>&1; echo "rc=$?"; grep -c OK gpurun_out/gemm_probe_pair_$TAG.log; grep -a "FAIL\|rror" gpurun_out/gemm_probe_pair_$TAG.log | head -20; grep -a "^time" gpurun_out/gemm_probe_pair_$TAG.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
