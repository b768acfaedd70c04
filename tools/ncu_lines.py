"""Join an ncu SASS-level source page with nvdisasm line info: per source line, instructions executed and stall samples.

  ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME > sass.csv
  cuobjdump -xelf all libpgb200.so ; nvdisasm -g -c pgb200.sm_100a.cubin > dis.txt
  python tools/ncu_lines.py sass.csv dis.txt MANGLED_SUBSTRING [top_n] [launch_index]
"""
import csv, re, sys
from collections import defaultdict

sass_csv, dis_txt, name = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0

# ---- nvdisasm: instruction index -> (file, line) inside the function
lines = open(dis_txt, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and name in l)
loc = []
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.lstrip().startswith(".section"):
        if loc:
            break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        loc.append(cur)

# ---- ncu csv: possibly several launches; take launch `which`
rows = list(csv.reader(open(sass_csv)))
blocks, cur_b = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur_b = {"name": r[1], "hdr": None, "ins": []}
        blocks.append(cur_b)
    elif r and r[0] == "Address":
        cur_b["hdr"] = r
    elif cur_b is not None and cur_b["hdr"] is not None and r:
        cur_b["ins"].append(r)
b = blocks[which]
h = b["hdr"]
ci = {n: h.index(n) for n in (x for x in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared Excessive",
                               "stall_long_sb", "stall_short_sb", "stall_barrier", "stall_mio", "stall_math", "stall_wait", "stall_not_selected", "stall_lg",
                               "stall_branch_resolving", "stall_no_inst") if x in h)}
print(f"kernel: {b['name'][:90]}  sass instructions: ncu {len(b['ins'])} / nvdisasm {len(loc)}")
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
for i, r in enumerate(b["ins"]):
    key = loc[i] if i < len(loc) else ("?", 0)
    for n, c in ci.items():
        if n == "Source":
            continue
        v = float(r[c] or 0)
        agg[key][n] += v
        tot[n] += v
print("totals:", {k: int(v) for k, v in tot.items()})
srcs = {}
def src(f, n):
    if f not in srcs:
        try:
            srcs[f] = open(f"peregrine_b200/csrc/{f}").read().splitlines()
        except OSError:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:100] if 0 < n <= len(srcs[f]) else ""
print(f"{'file:line':28s} {'inst%':>6s} {'smp%':>6s} {'long':>5s} {'short':>5s} {'barr':>5s} {'mio':>5s} {'math':>5s} {'wait':>5s} {'nsel':>5s} {'bank+':>8s}  source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    s = max(a["# Samples"], 1)
    print(f"{key[0] + ':' + str(key[1]):28s} {100 * a['Instructions Executed'] / tot['Instructions Executed']:6.2f} {100 * a['# Samples'] / tot['# Samples']:6.2f} "
          f"{100 * a['stall_long_sb'] / s:5.0f} {100 * a['stall_short_sb'] / s:5.0f} {100 * a['stall_barrier'] / s:5.0f} {100 * a['stall_mio'] / s:5.0f} "
          f"{100 * a['stall_math'] / s:5.0f} {100 * a['stall_wait'] / s:5.0f} {100 * a['stall_not_selected'] / s:5.0f} {int(a.get('L1 Wavefronts Shared Excessive', 0)):8d}  {src(*key)}")
