set -x
timeout 200 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_v9.log 2>&1; grep -a "FAIL" gpurun_out/gemm_probe_v9.log | head; tail -8 gpurun_out/gemm_probe_v9.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_v9.json 2> gpurun_out/bench_c3_v9.err; head -c 450 gpurun_out/bench_c3_v9.json; tail -3 gpurun_out/bench_c3_v9.err
timeout 300 python tools/gemm_breakdown.py c3 > gpurun_out/gemm_breakdown_c3_v9.txt 2> gpurun_out/gemm_breakdown_c3_v9.err; head -2 gpurun_out/gemm_breakdown_c3_v9.txt
timeout 300 ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/gemm_full_v9 python tools/ncu_gemm.py > gpurun_out/ncu_gemm_v9.log 2>&1; tail -2 gpurun_out/ncu_gemm_v9.log
