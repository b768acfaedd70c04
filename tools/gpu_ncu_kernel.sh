#!/bin/bash
# One `ncu --set full` capture of ONE kernel (first matching launch) of the bench workload.
#   bash tools/gpu_ncu_kernel.sh <kernel regex> <tag> [env assignments for bench.py, e.g. PGB_ALIGN_VARIANT=20]
K=$1; TAG=$2; shift 2
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -c ${NCU_COUNT:-1} -f -o gpurun_out/prof_$TAG \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log; ls -la gpurun_out/prof_$TAG.ncu-rep
