#!/bin/bash
# End-of-round evidence pass (one B200): GPU suite, default bench + reference arm, C4 bench, ncu launch list, DRAM traffic
# of every GEMM of one step, ncu --set full of the attention cores and the stem, critical-path attribution.
set -x
TAG=${1:-r02_final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py 2> gpurun_out/${TAG}_bench_c3.err | grep "^{" > gpurun_out/${TAG}_bench_c3.json; head -c 300 gpurun_out/${TAG}_bench_c3.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err | grep "^{" > gpurun_out/${TAG}_bench_ref.json; head -c 300 gpurun_out/${TAG}_bench_ref.json; echo
timeout 600 python bench.py --workload c4 --skip-cpu 2> gpurun_out/${TAG}_bench_c4.err | grep "^{" > gpurun_out/${TAG}_bench_c4.json; head -c 300 gpurun_out/${TAG}_bench_c4.json; echo
timeout 300 python tools/criticality.py c3 2>/dev/null | grep "ms/step" > gpurun_out/${TAG}_criticality_c3.txt; cat gpurun_out/${TAG}_criticality_c3.txt
timeout 300 python tools/step_timeline.py c3 2>/dev/null > gpurun_out/${TAG}_step_timeline_c3.txt; head -12 gpurun_out/${TAG}_step_timeline_c3.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_launches_c3.csv python tools/profile_step.py c3 > gpurun_out/${TAG}_ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_c3.csv > gpurun_out/${TAG}_launches_c3_summary.txt; head -30 gpurun_out/${TAG}_launches_c3_summary.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:gemm_split --csv \
  --log-file gpurun_out/${TAG}_gemm_traffic_c3.csv python tools/profile_step.py c3 > gpurun_out/${TAG}_ncu_traffic.log 2>&1
python tools/gemm_traffic.py gpurun_out/${TAG}_gemm_traffic_c3.csv > gpurun_out/${TAG}_gemm_traffic_c3.json; cat gpurun_out/${TAG}_gemm_traffic_c3.json
PROFILE_LAYERS=1 timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'rcda|mha|stem_conv' \
  -f -o gpurun_out/${TAG}_attn_stem python tools/profile_step.py c3 > gpurun_out/${TAG}_ncu_attn.log 2>&1
python tools/ncu_summary.py gpurun_out/${TAG}_attn_stem.ncu-rep > gpurun_out/${TAG}_ncu_attn_stem_summary.txt
du -sh gpurun_out
