set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v8.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_v8.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_v8.json 2> gpurun_out/bench_c3_v8.err; head -c 450 gpurun_out/bench_c3_v8.json; tail -3 gpurun_out/bench_c3_v8.err
timeout 600 python tools/gemm_sweep.py > gpurun_out/gemm_sweep_v8.txt 2>&1; tail -25 gpurun_out/gemm_sweep_v8.txt
