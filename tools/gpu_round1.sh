#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
PGB_VERBOSE=1 timeout 900 python tools/probe.py 50e6 30 3 > gpurun_out/probe_50mb.log 2>&1
grep -v "replay pass" gpurun_out/probe_50mb.log | tail -1 | cut -c1-2200
