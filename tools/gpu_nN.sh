#!/bin/bash
# multi-GPU bench line: $1 = N, $2 = exchange mode(s)
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
N=${1:-2}
for mode in ${2:-routed}; do
  SECONDS=0; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 --exchange $mode \
    > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  true
  echo "elapsed ${SECONDS}s"; grep -v "^\s" gpurun_out/bench_n${N}_$mode.err | tail -5
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('$mode', 'N', d['n_gpus'], 'dev ms', round(d['ms_per_step'],2), 'wall', round(d['config']['wall_ms_per_step_device_resident'],2), 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'ovl', d['config']['overlaps_per_step'])
    print({k:round(v,1) for k,v in d['stage_ms_per_step'].items()})
except Exception as e:
    print('no json', e)
PY
done
