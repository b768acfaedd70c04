#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 900 python tools/probe.py 50e6 30 2 > gpurun_out/probe_50mb.log 2>&1
tail -2 gpurun_out/probe_50mb.log | cut -c1-1800
