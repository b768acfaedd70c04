#!/bin/bash
# One GPU box visit: parity tests, smoke(), then the bench line.  Usage (from the repo root, on the GPU box via gpurun):
#   bash tools/gpu_check.sh [pytest -k expression]
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -8
nproc; free -g | head -2; df -h /tmp | tail -1
SECONDS=0
if [ -n "$1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest rc=$? (${SECONDS}s)"; tail -15 gpurun_out/pytest_gpu.log
SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "smoke rc=$? (${SECONDS}s)"
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
echo "bench rc=$? (${SECONDS}s)"; tail -3 gpurun_out/bench_ours.err
python tools/show_bench.py gpurun_out/bench_ours.json
