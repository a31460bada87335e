#!/bin/bash
# Session-3 validation pass of HEAD: GPU parity suite, GEMM probe, bench (resident-B on/off), ncu launch list of one
# C3 step, ncu --set full of representative GEMMs + attention kernels.
set -x
TAG=${1:-v11}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_$TAG.log 2>&1; grep -c OK gpurun_out/gemm_probe_$TAG.log; grep -a "FAIL" gpurun_out/gemm_probe_$TAG.log | head -20
timeout 400 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
head -c 600 gpurun_out/bench_c3_$TAG.json; tail -3 gpurun_out/bench_c3_$TAG.err
CDETR_GEMM_RESIDENT=0 timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_${TAG}_nores.json 2> gpurun_out/bench_c3_${TAG}_nores.err
head -c 300 gpurun_out/bench_c3_${TAG}_nores.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
head -c 300 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_c3_$TAG.csv python tools/profile_step.py c3 > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_c3_$TAG.csv > gpurun_out/launches_c3_${TAG}_summary.txt; head -30 gpurun_out/launches_c3_${TAG}_summary.txt
timeout 400 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/gemm_full_$TAG python tools/ncu_gemm.py > gpurun_out/ncu_gemm_$TAG.log 2>&1
PROFILE_LAYERS=1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'rcda|mha' \
  -f -o gpurun_out/attn_full_$TAG python tools/profile_step.py c3 > gpurun_out/ncu_attn_$TAG.log 2>&1
du -sh gpurun_out
