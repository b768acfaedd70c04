#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 2500 gpurun_out/bench_n2.json; tail -15 gpurun_out/bench_n2.err | cut -c1-300
