#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "two_bit or cli_chain or mkseqdb or config1" > gpurun_out/pytest_2bit.log 2>&1; echo "2-bit parity rc=$?"; tail -12 gpurun_out/pytest_2bit.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_2bit.json 2> gpurun_out/bench_2bit.err
echo "== bench rc=$?"; tail -2 gpurun_out/bench_2bit.err; python tools/show_bench.py gpurun_out/bench_2bit.json > gpurun_out/show_2bit.txt; head -4 gpurun_out/show_2bit.txt
bash tools/gpu_ncu_kernel.sh k_sketch_strip r2_strip
