set -x
timeout 200 python tests/gpu_gemm_probe.py > gpurun_out/gemm_probe_v10.log 2>&1; grep -a "FAIL" gpurun_out/gemm_probe_v10.log | head; grep -a "resident\|TN M=4096\|TN M=5000\|TN M=40000\|TN M=3000\|TN M=65536 N=512" gpurun_out/gemm_probe_v10.log | head
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_v10.json 2> gpurun_out/bench_c3_v10.err; head -c 450 gpurun_out/bench_c3_v10.json; tail -3 gpurun_out/bench_c3_v10.err
SWEEP_FULL=1 timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v10.txt 2>&1; grep -a "^M=" gpurun_out/gemm_sweep_v10.txt
