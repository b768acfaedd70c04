#!/bin/bash
# tests + bench (both arms) + ncu launch list + one full capture of the heavy kernels.  Outputs under gpurun_out/
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 3200 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sketch_tiled|k_align|k_replay' -c 5 -o gpurun_out/prof_r1b \
    python tools/probe.py 10e6 30 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -20
