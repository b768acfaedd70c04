#!/bin/bash
# first GPU session: parity tests, then a perf probe (outputs under gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
PGB_VERBOSE=1 timeout 600 python tools/probe.py 5e6 30 2 > gpurun_out/probe_5mb.log 2>&1
tail -30 gpurun_out/probe_5mb.log
