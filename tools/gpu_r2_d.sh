#!/bin/bash
mkdir -p gpurun_out
export PGB_WORK=/tmp/pgb_bench
timeout 900 python -m pytest tests -m gpu -x -q -k "index or sweep or adversarial or single_chunk or multi_chunk or config5 or mkseqdb or abi or two_bit or config2_50Mb" > gpurun_out/pytest_strip.log 2>&1; echo "strip parity rc=$?"; tail -12 gpurun_out/pytest_strip.log
for sk in strip tiled; do
  PGB_SKETCH=$sk timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_sk_$sk.json 2> gpurun_out/bench_sk_$sk.err
  echo "== bench sketch=$sk rc=$?"; tail -2 gpurun_out/bench_sk_$sk.err; python tools/show_bench.py gpurun_out/bench_sk_$sk.json > gpurun_out/show_sk_$sk.txt; head -4 gpurun_out/show_sk_$sk.txt; grep -o '"n_sketch_fallback_reads[^,]*' gpurun_out/bench_sk_$sk.json
done
