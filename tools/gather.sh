#!/bin/bash
# One GPU-box pass: GEMM/conv probe, per-shape GEMM table, ncu launch list of one C3 step, ncu --set full of the
# attention kernels (1-layer model: every kernel type once, real shapes).  Outputs (small!) under gpurun_out/.
set -x
TAG=${1:-v6}
mkdir -p gpurun_out
timeout 300 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_$TAG.log; grep -a "FAIL" gpurun_out/gemm_probe_$TAG.log | head -20
timeout 300 python tools/gemm_breakdown.py c3 > gpurun_out/gemm_breakdown_c3_$TAG.txt 2> gpurun_out/gemm_breakdown_c3_$TAG.err
head -3 gpurun_out/gemm_breakdown_c3_$TAG.txt; tail -3 gpurun_out/gemm_breakdown_c3_$TAG.err
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
head -c 400 gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_c3_$TAG.csv python tools/profile_step.py c3 > gpurun_out/ncu_launches_$TAG.log 2>&1
PROFILE_LAYERS=1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'rcda|mha' \
  -f -o gpurun_out/attn_full_$TAG python tools/profile_step.py c3 > gpurun_out/ncu_attn_$TAG.log 2>&1
du -sh gpurun_out
