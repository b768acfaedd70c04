#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" PGB_VERBOSE=1 timeout 600 python tools/probe.py 50e6 30 3 > gpurun_out/probe_$name.log 2>&1
  echo "== $name"; grep -v "replay pass\|outer khash" gpurun_out/probe_$name.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print({k:d[k] for k in ('wall_index_s','wall_overlap_s','overlaps','ms_pack','ms_sketch','ms_k_sketch_tiled','ms_reduce','ms_pairs','ms_buckets','ms_host_order','ms_replay','ms_align','ms_emit','ms_k_align','ms_k_replay','n_alignments','n_replay_passes','kernel_launches')})"
}
run default A=1
run g9 PGB200_LIB=build/libpgb200_g9.so
run g17 PGB200_LIB=build/libpgb200_g17.so
ncu --set full --clock-control none --import-source on -k regex:'k_align_lean|k_sketch_tiled' -c 2 -o gpurun_out/prof_r1e \
    python tools/probe.py 20e6 30 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
