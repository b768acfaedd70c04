#!/usr/bin/env python
"""Time the DROP-IN tools as pg_run.py / test/ecoli_K12/run_test.sh run them: one process per chunk, files in, files out.

    python tools/cli_e2e.py [--genome-mb 50] [--t-idx 12] [--t-ovlp 8] [--jobs 4] [--skip-ref]

Both arms run tools/run_chain.sh (shmr_mkseqdb -> shmr_index x T_idx -> shmr_overlap x T_ovlp -> cat | shmr_dedup) on the same
FASTA: once with this repository's bin/ (every process shares GPU 0), once with the unmodified reference in oracle/_ref
(`xargs -P jobs`, jobs = host cores for the reference arm).  Prints one JSON object: wall seconds per arm and per stage,
records, and whether preads.ovl is byte-identical.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_arm(bindir, lst, wd, ti, to, jobs):
    env = dict(os.environ, BIN=bindir)
    t = {}
    os.makedirs(os.path.join(wd, "index"), exist_ok=True)
    os.makedirs(os.path.join(wd, "ovlp"), exist_ok=True)
    os.makedirs(os.path.join(wd, "asm"), exist_ok=True)
    sh = lambda cmd: subprocess.run(["bash", "-c", cmd], check=True, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    t0 = time.perf_counter()
    sh(f'"$BIN/shmr_mkseqdb" -p "{wd}/index/seq_dataset" -d "{lst}"')
    t["mkseqdb_s"] = time.perf_counter() - t0
    t1 = time.perf_counter()
    sh(f'seq 1 {ti} | xargs -P {jobs} -I{{}} "$BIN/shmr_index" -p "{wd}/index/seq_dataset" -r 6 -t {ti} -c {{}} -m 0 -o "{wd}/index/shmr"')
    t["index_s"] = time.perf_counter() - t1
    t2 = time.perf_counter()
    sh(f'seq -f "%02g" 1 {to} | xargs -P {jobs} -I{{}} "$BIN/shmr_overlap" -p "{wd}/index/seq_dataset" -l "{wd}/index/shmr-L2" -t {to} -c {{}} -o "{wd}/ovlp/ovlp.{{}}"')
    t["overlap_s"] = time.perf_counter() - t2
    t3 = time.perf_counter()
    sh(f'cat "{wd}"/ovlp/ovlp.* | "$BIN/shmr_dedup" > "{wd}/asm/preads.ovl"')
    t["dedup_s"] = time.perf_counter() - t3
    t["total_s"] = time.perf_counter() - t0
    t["index_overlap_s"] = t["index_s"] + t["overlap_s"]
    t["records"] = sum(os.path.getsize(os.path.join(wd, "ovlp", f)) for f in os.listdir(os.path.join(wd, "ovlp"))) // 64
    return t


def main():
    import datasets as D

    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mb", type=float, default=50.0)
    ap.add_argument("--cov", type=float, default=30.0)
    ap.add_argument("--t-idx", type=int, default=12)
    ap.add_argument("--t-ovlp", type=int, default=8)
    ap.add_argument("--jobs", type=int, default=4, help="concurrent processes of OUR arm (they share one GPU)")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--work", default=os.environ.get("PGB_WORK", "/tmp/pgb_bench"))
    a = ap.parse_args()
    g = int(a.genome_mb * 1e6)
    d = os.path.join(a.work, f"cli_g{g}")
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "seq.fa")
    if not os.path.exists(fa):
        D.run([D.ensure_simreads(), "-g", str(g), "-c", str(a.cov), "-e", "0.005", "-S", "42", "-f", "-p", os.path.join(d, "seq")])
    lst = os.path.join(d, "in.lst")
    with open(lst, "w") as f:
        f.write(fa + "\n")
    cores = os.cpu_count() or 1
    out = {"workload": f"synthetic {a.genome_mb:g} Mb genome, {a.cov:g}x 15 kb reads @99.5%, T_idx={a.t_idx}, T_ovlp={a.t_ovlp}, files on local disk",
           "ours": run_arm(os.path.join(ROOT, "bin"), lst, os.path.join(d, "our"), a.t_idx, a.t_ovlp, a.jobs)}
    out["ours"]["jobs"] = a.jobs
    if not a.skip_ref:
        out["reference"] = run_arm(os.path.join(ROOT, "oracle", "_ref"), lst, os.path.join(d, "ref"), a.t_idx, a.t_ovlp, cores)
        out["reference"]["jobs"] = cores
        out["host_cores"] = cores
        same = open(os.path.join(d, "our", "asm", "preads.ovl"), "rb").read() == open(os.path.join(d, "ref", "asm", "preads.ovl"), "rb").read()
        out["preads_ovl_identical"] = same
        out["speedup_index_overlap"] = out["reference"]["index_overlap_s"] / out["ours"]["index_overlap_s"]
        out["speedup_total"] = out["reference"]["total_s"] / out["ours"]["total_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
