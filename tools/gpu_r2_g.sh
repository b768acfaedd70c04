#!/bin/bash
# round 2, run g: warp-walked replay buckets (k_replay_warp) and the (length, target) alignment order: parity, then A/B
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "replay or single_chunk or multi_chunk or adversarial or noisy or config5 or overflow or abi or engine_api" > gpurun_out/pytest_g.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_g.log
timeout 600 python bench.py > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/bench_g.json | sed -n 1,4p; grep -o '"parity": {[^}]*}' gpurun_out/bench_g.json
for rw in 64 6 8 16 24; do
  PGB_REPLAY_WARP_MIN=$rw timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_rw$rw.json 2> gpurun_out/bench_rw$rw.err; echo "rw_min=$rw rc=$?"; python tools/show_bench.py gpurun_out/bench_rw$rw.json | sed -n '1p;4p'
done
for so in 0 2 3 5 6 9; do
  PGB_ALIGN_SORT=$so timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_so$so.json 2> gpurun_out/bench_so$so.err; echo "align_sort=$so rc=$?"; python tools/show_bench.py gpurun_out/bench_so$so.json | sed -n '1p;4p'
done
PGB_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "replay pass" | tail -14
