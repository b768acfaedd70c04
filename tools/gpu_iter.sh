#!/bin/bash
# one development iteration on the GPU: parity tests, a short bench, optionally an ncu capture ($1 = tag, $2 = kernel regex)
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('dev ms', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'e2e value', round(d['e2e']['value']))
print({k:round(v,2) for k,v in d['stage_ms_per_step'].items()})
print({k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
tail -3 gpurun_out/bench_ours.err
if [ -n "$2" ]; then
  ncu --set full --clock-control none --import-source on -k regex:"$2" -c ${3:-4} -o gpurun_out/prof_$1 -f python tools/probe.py 50e6 30 1 > gpurun_out/ncu_full_$1.log 2>&1
  tail -2 gpurun_out/ncu_full_$1.log
fi
