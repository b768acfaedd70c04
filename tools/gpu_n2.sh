#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
N=${1:-2}
for mode in routed gathered; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --exchange $mode \
    > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  tail -3 gpurun_out/bench_n${N}_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_$mode.json').read().strip().splitlines()[-1])
print('$mode', 'N', d['n_gpus'], 'dev ms', round(d['ms_per_step'],2), 'wall', round(d['config']['wall_ms_per_step_device_resident'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'ovl', d['config']['overlaps_per_step'])
print(d['stage_ms_per_step'])
PY
done
