#!/bin/bash
# End-of-round verification on one GPU: the whole -m gpu suite, smoke(), the bench line with the CPU baseline / parity check,
# then the ncu evidence of the same bench command: launch list + one `--set full` capture each of the three heavy kernels.
bash tools/gpu_check.sh
export PGB_WORK=/tmp/pgb_bench
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_r2.log 2>&1; echo "launch list rc=$?"
bash tools/gpu_ncu_kernel.sh k_align_lean r2_align
bash tools/gpu_ncu_kernel.sh k_sketch_strip r2_sketch
NCU_COUNT=3 bash tools/gpu_ncu_kernel.sh "k_replay" r2_replay
