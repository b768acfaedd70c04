#!/bin/bash
# round bench: both arms, ncu launch list of one bench step, ncu --set full of the heavy kernels (bench-size workload)
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -c 3500 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_align_lean|k_sketch_tiled|k_replay$|k_replay_block|k_align_warp|k_pack_reads' -c 6 -o gpurun_out/prof_${1:-r1f} \
    python tools/probe.py 50e6 30 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | head -30
