#!/bin/bash
# next GPU session: A/B of the unmeasured k_align experiments (7 = production, 12 = + L2 prefetch, 13 = persistent lanes,
# 14 = persistent lanes + L2 prefetch); parity of 13/14 first (CLI vs reference on a small set), then the 50 Mb workload
export PGB_WORK=/tmp/pgb_bench
mkdir -p gpurun_out
for v in 12 13 14; do
  PGB_ALIGN_VARIANT=$v PGB_ALIGN_WARP_MAX=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_chunk or noisy or adversarial" 2>&1 | tail -2
done
for v in 7 12 13 14; do
  echo "== PGB_ALIGN_VARIANT=$v"
  PGB_ALIGN_VARIANT=$v python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_align','ms_k_align','ms_replay','overlaps','wall_overlap_s')})"
done
# k_reduce_warp: parity (index sweep incl. r = 36, 2, 1-level) then timing
PGB_REDUCE=warp timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "index_parameter_sweep or single_chunk or multi_chunk or adversarial" 2>&1 | tail -2
for m in thread warp; do
  echo "== PGB_REDUCE=$m"
  PGB_REDUCE=$m python tools/probe.py 50e6 30 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('ms_reduce','ms_sketch','overlaps','wall_index_s')})"
done
