#!/bin/bash
# Robustness pass: other workloads (C2 stage 1, C4 800x800/Q=500 on the CUDA-core RCDA + legacy MHA paths), 2-GPU bench.
set -x
TAG=${1:-v21}
mkdir -p gpurun_out
for wl in c2 c4; do
  timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${wl}_$TAG.json').read()); print('$wl', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['launch'], d['loss'])"; tail -2 gpurun_out/bench_${wl}_$TAG.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --skip-cpu > gpurun_out/bench_c3_2gpu_$TAG.json 2> gpurun_out/bench_c3_2gpu_$TAG.err
tail -1 gpurun_out/bench_c3_2gpu_$TAG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('2gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'])"; tail -3 gpurun_out/bench_c3_2gpu_$TAG.err
