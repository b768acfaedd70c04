#!/bin/bash
export PGB_WORK=/tmp/pgb_bench
mkdir -p gpurun_out
for seg in 256 64 32; do
  echo "== PGB_EXACT_SEG=$seg"
  PGB_EXACT_SEG=$seg python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_sketch','ms_k_sketch_tiled','ms_k_sketch_count','ms_k_sketch_write','n_sketch_fallback_reads','overlaps')})"
done
echo "== verbose replay passes"
PGB_VERBOSE=1 python tools/probe.py 50e6 30 2 2>&1 | grep "replay pass\|outer khash" | tail -40
