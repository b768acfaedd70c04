#!/bin/bash
export PGB_WORK=/tmp/pgb_bench
mkdir -p gpurun_out
for v in 7 8 9 10; do
  echo "== PGB_ALIGN_VARIANT=$v"
  PGB_ALIGN_VARIANT=$v python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_align','ms_k_align','ms_replay','overlaps','wall_overlap_s')})"
done
