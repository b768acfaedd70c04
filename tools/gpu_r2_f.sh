#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "index or sweep or adversarial or single_chunk or multi_chunk or config5 or abi" > gpurun_out/pytest_f.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_f.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench_f.json | sed -n 1,4p
PGB_ALIGN_SORT=target timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_sorttarget.json 2> gpurun_out/bench_sorttarget.err; echo "sort=target rc=$?"; python tools/show_bench.py gpurun_out/bench_sorttarget.json | sed -n 4p
PGB_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "replay pass" | tail -14
timeout 1200 python tools/cli_e2e.py --genome-mb 500 > gpurun_out/cli_e2e_500.json 2> gpurun_out/cli_e2e_500.err; echo "cli_e2e 500 rc=$?"; cat gpurun_out/cli_e2e_500.json; tail -2 gpurun_out/cli_e2e_500.err
