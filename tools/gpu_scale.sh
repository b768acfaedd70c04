#!/bin/bash
# Multi-GPU bench lines on one box (gpurun --gpus N): bash tools/gpu_scale.sh "2 4" [extra bench.py args]
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
NS=${1:-2}; shift
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c; nproc; df -h /tmp | tail -1
for N in $NS; do
  SECONDS=0
  timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N "$@" \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
  echo "N=$N rc=$? (${SECONDS}s)"; grep -v "^\s" gpurun_out/bench_n$N.err | grep -v "^W1\|OMP_NUM\|^\*\*\*" | tail -12
  python tools/show_bench.py gpurun_out/bench_n$N.json
done
