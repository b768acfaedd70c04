#!/bin/bash
# Multi-GPU verification on one box (gpurun --gpus 4): the tests that need more than one GPU, then the weak-scaling lines.
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "${1:-nccl or two_rank or bad_strips}" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
bash tools/gpu_scale.sh "${2:-2 4}"
