#!/bin/bash
# round 2, run i: batched shmr_aln, replay thresholds of the incremental passes, allocator counters in the e2e line
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -s -k "abi or utils or replay_kernel" > gpurun_out/pytest_i.log 2>&1; echo "abi rc=$?"; grep -E "shmr_aln:|ovlp_match|passed|failed|Error" gpurun_out/pytest_i.log | tail -8
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag rc=$?"; python tools/show_bench.py gpurun_out/bench_$tag.json | sed -n '1p;4p'
  grep -o '"steps_ms_rank0": [^}]*}' gpurun_out/bench_$tag.json
}
run i_default PGB_X=0
run i_inc4g4 PGB_REPLAY_WARP_MIN_INC=4 PGB_REPLAY_GROUP=4
run i_inc6 PGB_REPLAY_WARP_MIN_INC=6
run i_inc8g4 PGB_REPLAY_WARP_MIN_INC=8 PGB_REPLAY_GROUP=4
run i_inc12 PGB_REPLAY_WARP_MIN_INC=12
run i_tail4 PGB_REPLAY_BIG_TAIL=48
PGB_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "replay pass" | tail -13
