#!/bin/bash
# round 2, run h: lane-group replay (k_replay_group<G>): parity of every form, then A/B over group width and thresholds
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "replay or single_chunk or multi_chunk or adversarial or noisy or overflow" > gpurun_out/pytest_h.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/pytest_h.log
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "$tag rc=$?"; python tools/show_bench.py gpurun_out/bench_$tag.json | sed -n '1p;4p'
}
run h_default PGB_X=0
run h_off PGB_REPLAY_WARP_MIN=64
run h_g4 PGB_REPLAY_GROUP=4
run h_g16 PGB_REPLAY_GROUP=16
run h_min8 PGB_REPLAY_WARP_MIN=8
run h_min24 PGB_REPLAY_WARP_MIN=24
run h_inc8 PGB_REPLAY_WARP_MIN=64 PGB_REPLAY_WARP_MIN_INC=8
run h_inc16 PGB_REPLAY_WARP_MIN=64 PGB_REPLAY_WARP_MIN_INC=16
run h_g4min8 PGB_REPLAY_GROUP=4 PGB_REPLAY_WARP_MIN=8
PGB_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "replay pass" | tail -13
