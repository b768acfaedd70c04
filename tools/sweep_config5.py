#!/usr/bin/env python
"""BASELINE.json configs[4]: throughput of the (k, w, aln_bw) sweep on the 50 Mb synthetic set, one B200.
Runs bench.py once per parameter set (k x w at aln_bw=100, then aln_bw 50 / 200 at the default k, w) and writes the JSON lines
plus a small table.  usage: python tools/sweep_config5.py [out.json]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "config5_sweep.json")
combos = [(k, w, 100) for k in (14, 16, 18) for w in (60, 80, 120)] + [(16, 80, 50), (16, 80, 200)]
rows = []
for k, w, bw in combos:
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "2", "--no-cpu-baseline", "--k", str(k), "--w", str(w), "--aln-bw", str(bw)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode or not line:
        rows.append({"k": k, "w": w, "aln_bw": bw, "error": (r.stderr or "")[-400:]})
        print(k, w, bw, "FAILED", file=sys.stderr)
        continue
    j = json.loads(line[-1])
    rows.append({"k": k, "w": w, "aln_bw": bw, "overlaps_per_s": j["value"], "ms_per_step": j["ms_per_step"], "e2e_overlaps_per_s": j["e2e"]["value"],
                 "read_bases_per_s": j.get("read_bases_per_s"), "overlaps_per_step": j["counts_per_step"].get("n_overlaps") if "counts_per_step" in j else None,
                 "stage_ms": j.get("stage_ms_per_step"), "kernel_ms": j["roofline"]["kernel_ms_per_step"], "counts": j.get("counts_per_step"),
                 "sketch_kernel": "k_sketch_strip<u64>" if k > 16 else "k_sketch_strip<u32>"})
    print(f"k={k} w={w} bw={bw}: {j['value'] / 1e6:.2f} M ovl/s dev ({j['ms_per_step']:.1f} ms), e2e {j['e2e']['value'] / 1e6:.2f} M ovl/s, "
          f"kernels {{{', '.join(f'{a}: {b:.1f}' for a, b in j['roofline']['kernel_ms_per_step'].items())}}}")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump({"workload": "synthetic 50 Mb genome, 30x 15 kb reads @99.5%, T=1, one B200; bench.py --steps 2 --warmup 2", "rows": rows}, open(out_path, "w"), indent=1)
