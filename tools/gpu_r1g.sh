#!/bin/bash
# tests + e2e probe (1 GPU part)
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python tools/probe_e2e.py 50e6 96 32 256 2>&1 | grep -v '"iter": 0' | tee gpurun_out/probe_e2e.log
