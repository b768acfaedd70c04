#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
export PGB_WORK=/tmp/pgb_bench
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('dev ms', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'])
print(d['stage_ms_per_step'])"
tail -3 gpurun_out/bench_ours.err
PGB_LOAD_CHUNK_MB=32 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunk32: e2e ms', d['e2e']['ms_per_step'])"
PGB_LOAD_CHUNK_MB=256 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunk256: e2e ms', d['e2e']['ms_per_step'])"
