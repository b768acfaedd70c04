#!/bin/bash
# A/B of the bulk alignment kernels on one GPU: parity under every variant, then bench lines.
#   VARIANTS="20 21" FULLTESTS=1 bash tools/gpu_align_ab.sh
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "thread_per_alignment" > gpurun_out/pytest_align.log 2>&1; echo "variants rc=$?"; tail -5 gpurun_out/pytest_align.log
if [ -n "$FULLTESTS" ]; then
for v in ${VARIANTS:-20 21}; do
  PGB_ALIGN_VARIANT=$v PGB_ALIGN_WARP_MAX=0 timeout 900 python -m pytest tests -m gpu -x -q -k "adversarial or single_chunk or config5 or config2_50Mb or overflow" > gpurun_out/pytest_align_v$v.log 2>&1
  echo "variant $v on the parity tests rc=$?"; tail -4 gpurun_out/pytest_align_v$v.log
done
fi
for v in ${BASE:-7} ${VARIANTS:-20 21}; do
  PGB_ALIGN_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  echo "== bench variant $v rc=$?"; tail -2 gpurun_out/bench_v$v.err; python tools/show_bench.py gpurun_out/bench_v$v.json > gpurun_out/show_v$v.txt; head -3 gpurun_out/show_v$v.txt
done
