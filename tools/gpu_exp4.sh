#!/bin/bash
export PGB_WORK=/tmp/pgb_bench
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_replay','ms_k_replay','ms_emit','ms_align','n_replay_passes','overlaps','wall_overlap_s')})"
}
run PGB_TAIL_RUN=40000
run PGB_TAIL_RUN=70000
run PGB_TAIL_RUN=120000
run PGB_TAIL_RUN=200000
run PGB_TAIL_RUN=120000 PGB_REPLAY_BIG_TAIL=16
run PGB_TAIL_RUN=120000 PGB_REPLAY_BIG_TAIL=24
run PGB_TAIL_RUN=120000 PGB_REPLAY_BIG_TAIL=12
run PGB_TAIL_RUN=400000 PGB_REPLAY_BIG_TAIL=24
run PGB_DRY_PASSES=1
run PGB_DRY_PASSES=3
