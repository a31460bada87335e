#!/bin/bash
# End-of-round evidence pass: GPU suite, default bench + reference arm, ncu launch list, DRAM traffic of every GEMM of
# one step, ncu --set full of representative GEMMs (incl. the CTA-pair kernel) and of the attention cores.
set -x
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err; head -c 400 gpurun_out/bench_c3_$TAG.json; tail -2 gpurun_out/bench_c3_$TAG.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; head -c 300 gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_c3_$TAG.csv python tools/profile_step.py c3 > gpurun_out/ncu_launches_$TAG.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_c3_$TAG.csv > gpurun_out/launches_c3_${TAG}_summary.txt; head -24 gpurun_out/launches_c3_${TAG}_summary.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_split --csv \
  --log-file gpurun_out/gemm_traffic_c3_$TAG.csv python tools/profile_step.py c3 > gpurun_out/ncu_traffic_$TAG.log 2>&1
python tools/gemm_traffic.py gpurun_out/gemm_traffic_c3_$TAG.csv > gpurun_out/gemm_traffic_c3_${TAG}.json; cat gpurun_out/gemm_traffic_c3_${TAG}.json
timeout 400 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/gemm_full_$TAG python tools/ncu_gemm.py > gpurun_out/ncu_gemm_$TAG.log 2>&1
python tools/ncu_extract.py gpurun_out/gemm_full_$TAG.ncu-rep > gpurun_out/ncu_gemm_full_${TAG}_summary.txt
PROFILE_LAYERS=1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'rcda|mha' \
  -f -o gpurun_out/attn_full_$TAG python tools/profile_step.py c3 > gpurun_out/ncu_attn_$TAG.log 2>&1
python tools/ncu_extract.py gpurun_out/attn_full_$TAG.ncu-rep > gpurun_out/ncu_attn_full_${TAG}_summary.txt
du -sh gpurun_out
