#!/bin/bash
# BASELINE.json configs[2] / configs[3] style lines.  usage: bash tools/gpu_configs.sh <N> <tag> <bench args...>
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
SECONDS=0
timeout ${LIMIT:-800} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" \
  > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "$TAG rc=$? (${SECONDS}s)"; grep -v "^\s" gpurun_out/bench_$TAG.err | grep -v "^W1\|OMP_NUM\|^\*\*\*\|NCCL version" | tail -8
python tools/show_bench.py gpurun_out/bench_$TAG.json | head -12
df -h /tmp | tail -1; free -g | sed -n 2p
