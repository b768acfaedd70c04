#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? (${SECONDS}s)"; tail -10 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench_ours.json | head -6
for v in 12 13 14; do
  PGB_ALIGN_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  echo "== align variant $v rc=$?"; python tools/show_bench.py gpurun_out/bench_v$v.json | sed -n 4p
done
PGB_REDUCE=warp timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_reduce_warp.json 2> gpurun_out/bench_reduce_warp.err
echo "== reduce warp rc=$?"; python tools/show_bench.py gpurun_out/bench_reduce_warp.json | sed -n 3p
timeout 900 python tools/cli_e2e.py > gpurun_out/cli_e2e.json 2> gpurun_out/cli_e2e.err; echo "cli_e2e rc=$?"; cat gpurun_out/cli_e2e.json; tail -3 gpurun_out/cli_e2e.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r2.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_r2.csv
