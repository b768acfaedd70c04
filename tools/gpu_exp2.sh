#!/bin/bash
export PGB_WORK=/tmp/pgb_bench
mkdir -p gpurun_out
for big in 48 32 24 16 12 8 4; do
  echo "== PGB_REPLAY_BIG=$big"
  PGB_REPLAY_BIG=$big python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_replay','ms_k_replay','ms_emit','ms_align','n_replay_passes','overlaps','wall_overlap_s')})"
done
for seg in 512 1024; do
  echo "== PGB_EXACT_SEG=$seg"
  PGB_EXACT_SEG=$seg python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_sketch','ms_k_sketch_tiled','ms_k_sketch_count','ms_k_sketch_write','overlaps')})"
done
