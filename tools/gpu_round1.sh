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
print({k:d[k] for k in ('wall_index_s','wall_overlap_s','overlaps','ms_pack','ms_sketch','ms_replay','ms_align','ms_emit','ms_k_align','ms_k_replay','n_alignments','n_replay_passes','kernel_launches')})"
}
run default A=1
run tr_all8 PGB_TAIL_RUN=4000000000
run tr_all16 PGB_TAIL_RUN=4000000000 PGB_REPLAY_BIG_TAIL=16
run tr_all24 PGB_TAIL_RUN=4000000000 PGB_REPLAY_BIG_TAIL=24
run big24 PGB_REPLAY_BIG=24
run big48 PGB_REPLAY_BIG=48
