#!/bin/bash
# Closing numbers after the last changes: GPU suite, default bench (+ reference arm), C4 bench.
TAG=${1:-r02_close}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py 2> gpurun_out/${TAG}_bench_c3.err | grep "^{" > gpurun_out/${TAG}_bench_c3.json; head -c 200 gpurun_out/${TAG}_bench_c3.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep "^{" > gpurun_out/${TAG}_bench_ref.json; head -c 120 gpurun_out/${TAG}_bench_ref.json; echo
timeout 600 python bench.py --workload c4 --skip-cpu 2>/dev/null | grep "^{" > gpurun_out/${TAG}_bench_c4.json; head -c 200 gpurun_out/${TAG}_bench_c4.json; echo
timeout 300 python tools/criticality.py c3 2>/dev/null | grep "ms/step" > gpurun_out/${TAG}_criticality_c3.txt; cat gpurun_out/${TAG}_criticality_c3.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
