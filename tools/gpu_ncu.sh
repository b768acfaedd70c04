#!/bin/bash
# ncu --set full capture (with source) of the heavy kernels on the bench-size workload; $1 = output tag, $2 = kernel regex
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
TAG=${1:-r1g}
RE=${2:-'k_align_lean|k_sketch_tiled|k_replay$'}
ncu --set full --clock-control none --import-source on -k regex:"$RE" -c ${3:-4} -o gpurun_out/prof_$TAG -f \
    python tools/probe.py 50e6 30 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out | head
