#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 3300 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# first (big) k_align launch, the tiled sketch kernel and the first replay pass, full sections
ncu --set full --clock-control none --import-source on -k regex:'k_align$|k_sketch_tiled|k_replay$' -c 4 -o gpurun_out/prof_r1c \
    python tools/probe.py 20e6 30 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -20
