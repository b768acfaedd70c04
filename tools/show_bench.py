#!/usr/bin/env python
"""Print the essentials of a bench.py JSON line (last line of the given file)."""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:  # noqa: BLE001
    print("no json:", e)
    sys.exit(0)
if "error" in d:
    print("ERROR:", d["error"])
    print(d.get("traceback_tail", ""))
    sys.exit(0)
r = d.get("roofline", {})
print(f"N={d['n_gpus']} value={d['value']:.4g} {d['unit']}  dev {d['ms_per_step']:.2f} ms  e2e {d['e2e']['value']:.4g} ({d['e2e'].get('ms_per_step', 0):.2f} ms)  "
      f"bases/s {d.get('read_bases_per_s', 0):.4g}  ovl/step {d['config'].get('overlaps_per_step')}  launches {d.get('gpu_launches')}")
if "e2e_seqdb" in d:
    print(f"e2e from .seqdb: {d['e2e_seqdb']['value']:.4g} ({d['e2e_seqdb']['ms_per_step']:.2f} ms, h2d {d['e2e_seqdb']['h2d_bytes_per_step']/1e9:.2f} GB)  e2e h2d {d['e2e']['h2d_bytes_per_step']/1e9:.2f} GB")
print("stages:", {k: round(v, 2) for k, v in d.get("stage_ms_per_step", {}).items() if v})
print("kernels:", {k: round(v, 2) for k, v in r.get("kernel_ms_per_step", {}).items()}, "frac", round(r.get("frac", 0), 4), r.get("kernel"))
if "rooflines" in d:
    print("rooflines:", {k: round(v["frac"], 4) for k, v in d["rooflines"].items()})
print("counts:", d.get("counts_per_step"))
print("clocks:", d.get("clocks"))
for k in ("parity", "cpu_baseline"):
    if k in d:
        print(k + ":", d[k])
