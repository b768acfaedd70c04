#!/bin/bash
# bench (both arms) + ncu launch list + one full capture of the heavy kernels.  Outputs under gpurun_out/
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 1500 gpurun_out/bench_ref.json
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
# launch list of one step (cold-cache, serialised): compare shares
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# full capture of the heavy kernels (few launches each)
ncu --set full --clock-control none --import-source on -k regex:'k_sketch_exact|k_align|k_replay' -c 6 -o gpurun_out/prof_r1 \
    python tools/probe.py 10e6 30 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
