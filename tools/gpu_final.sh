#!/bin/bash
# round-end single-GPU evidence: parity tests, bench line (with CPU baseline), ncu launch list of one bench step, ncu --set full of the sketch kernel
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
TAG=${1:-r1g}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log
cat gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_$TAG.json 2> gpurun_out/bench_ours_$TAG.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ours_$TAG.json'))
print('dev ms', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e value', round(d['e2e']['value']), 'launches', d['gpu_launches'])
print({k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
print(d['roofline']); print(d.get('cpu_baseline')); print(d['clocks'])"
tail -3 gpurun_out/bench_ours_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sketch_tiled|k_tile_gather|k_pack_reads' -c 3 -o gpurun_out/prof_${TAG}_sketch -f \
    python tools/probe.py 50e6 30 1 > gpurun_out/ncu_full_${TAG}_sketch.log 2>&1
ls -la gpurun_out | tail -8
