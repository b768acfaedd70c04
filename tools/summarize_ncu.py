"""Summarise gpurun_out ncu artefacts into profiles/ (markdown + csv).  usage: summarize_ncu.py <tag> [launches.csv] [prof.ncu-rep]"""
import csv, subprocess, sys, os, collections, io
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else None
out = [f"# ncu summary — {tag}\n"]
if launches and os.path.exists(launches):
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"): v /= 1e6
        elif r[ui] in ("us", "usecond"): v /= 1e3
        elif r[ui] in ("s", "second"): v *= 1e3
        name = r[ki].split("(")[0].replace("void ", "").replace("pgb::", "")
        name = name[:70]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    total = sum(v[1] for v in tot.values())
    out.append(f"## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`), {sum(v[0] for v in tot.values())} launches, {total:.2f} ms of kernel time\n")
    out.append("cold-cache, serialised per-launch times: compare SHARES, not absolutes\n")
    out.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100*v[1]/total:.1f}% |")
    out.append("")
if rep and os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed.sum"]
    idx = [hdr.index(w) for w in want if w in hdr]
    out.append(f"## `ncu --set full --clock-control none` ({os.path.basename(rep)})\n")
    out.append("| kernel | " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |")
    out.append("|---|" + "---:|" * len(idx))
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        out.append(f"| `{r[ki].split('(')[0].replace('void ','')[:40]}` | " + " | ".join(r[i] for i in idx) + " |")
    out.append("")
os.makedirs("profiles", exist_ok=True)
open(f"profiles/{tag}_ncu.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
