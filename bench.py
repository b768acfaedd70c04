#!/usr/bin/env python
"""bench.py — SHIMMER index + read-to-read overlap throughput on B200 (BASELINE.json metric: overlaps/s and read-bases/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (pack -> mm_sketch -> mm_reduce x2 -> count -> build_map -> bucket order ->
replay/ovlp_match fix-point -> ovlp records) over one synthetic read set.  Workload at N=1 = BASELINE.json configs[1]:
synthetic 50 Mb genome, 30x 15 kb reads at 99.5 % accuracy, k=16 w=80 r=6 l=2, single chunk (T=1).  At N>1 (torchrun)
the genome grows with N (weak scaling: 50 Mb per GPU) and hash chunk c of T=N is owned by rank c-1.

  value  = overlaps/s with the .seqdb image already resident in HBM (device-resident step, CUDA-event timed)
  e2e    = the same through the public C ABI with HOST buffers: pinned .seqdb -> H2D -> ... -> ovlp records -> D2H
  --impl reference : the unmodified reference tools (oracle/_ref, built from /root/reference/src) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

PARAMS = dict(w=80, k=16, r=6, levels=2, bestn=4, mc_lower=2, mc_upper=240, bw=100, ovlp_upper=120)
REF = os.path.join(ROOT, "oracle", "_ref")


def work_dir():
    d = os.environ.get("PGB_WORK", "/tmp/pgb_bench")
    os.makedirs(d, exist_ok=True)
    return d


def make_dataset(name, genome, cov, seed=42, err=0.005, mod=1, res=0):
    import datasets as D

    return D.make_sim(work_dir(), name, genome=genome, cov=cov, err=err, seed=seed, mod=mod, res=res)


class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms through NVML while the timed region runs (a timed region of a few
    steps lasts ~0.2 s, too short for `nvidia-smi -lms`); falls back to one `nvidia-smi` query loop if pynvml is unusable."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.sm, self.mx, self.reasons, self.stop_flag, self.t, self.p, self.rows = device, [], 0, set(), False, None, None, []
        self.nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices; NVML wants the physical one
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device
            if vis:
                toks = [t.strip() for t in vis.split(",") if t.strip()]
                if device < len(toks) and toks[device].isdigit():
                    idx = int(toks[device])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.h))
                for bit, name in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def start(self):
        self.stop_flag = False
        if self.nvml:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def stop(self):
        self.stop_flag = True
        if self.nvml:
            if self.t:
                self.t.join(timeout=1)
        elif self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()
            for r in self.rows:
                try:
                    self.sm.append(float(r[1]))
                    self.mx = max(self.mx, float(r[2]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
                except (ValueError, IndexError):
                    pass
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons), "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ reference / cpu legs
def ref_one(prefix, outdir):
    """index + overlap of one read set with the unmodified reference tools, T=1 (one core)."""
    import datasets as D

    t = time.perf_counter()
    rp = D.ref_index(REF, prefix, outdir, T=1, extra=["-m", "0", "-k", str(PARAMS["k"]), "-w", str(PARAMS["w"])])
    ro = D.ref_overlap(REF, prefix, rp, 2, outdir, T=1, extra=["-w", str(PARAMS["bw"])])
    return os.path.getsize(ro[0]) // 64, time.perf_counter() - t


def cpu_baseline_single_core(genome=15_000_000, cov=30, device=0):  # ~12 s of single-core reference work
    """The unmodified reference on one core over a bounded sample of the bench workload (same read model, 15 Mb genome).  Its
    records are then the checker for OUR engine on the same set: `parity` says whether the two streams are identical
    (same records, same order); the bench fails if they are not."""
    from peregrine_b200 import Engine, formats as F

    p = make_dataset(f"cpu1_g{genome}", genome, cov, seed=1234)
    rid, ln, off = F.read_idx(p + ".idx")
    outdir = os.path.join(work_dir(), f"cpu1_g{genome}", "ref")
    n, dt = ref_one(p, outdir)
    want = F.normalise_ovlp(F.read_ovlp(os.path.join(outdir, "ovlp.01")))
    eng = Engine(device)
    eng.load_reads(np.fromfile(p + ".seqdb", dtype=np.uint8), rid, ln, off)
    P = PARAMS
    eng.index(P["w"], P["k"], P["r"], P["levels"], 0)
    eng.set_shimmers_from_index(2)
    got = eng.overlap(1, 1, P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"])
    eng.close()
    same = len(got) == len(want) and got.tobytes() == want.tobytes()
    parity = {"records": int(len(want)), "records_ours": int(len(got)), "identical": bool(same),
              "checked": f"ovlp_t stream of this engine vs oracle/_ref shmr_overlap on the {genome/1e6:g} Mb sample, field-normalised, in order"}
    return {"value": n / dt, "unit": "overlaps/s", "cores": 1, "kind": "reference",
            "read_bases_per_s": float(ln.sum()) / dt,
            "sample": f"oracle/_ref shmr_index+shmr_overlap (unmodified reference), {genome/1e6:g} Mb genome {cov:g}x, T=1, one core, {dt:.1f} s"}, parity


def run_reference_arm(args):
    from concurrent.futures import ThreadPoolExecutor
    from peregrine_b200 import formats as F

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    g = int(args.ref_genome_mb * 1e6)
    sets = [make_dataset(f"ref_g{g}_s{i}", g, args.cov, seed=1000 + i) for i in range(cores)]
    bases = sum(int(F.read_idx(p + ".idx")[1].sum()) for p in sets)

    def step():
        t = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            res = list(ex.map(lambda i: ref_one(sets[i], os.path.join(os.path.dirname(sets[i]), "ref")), range(cores)))
        return sum(r[0] for r in res), time.perf_counter() - t

    for _ in range(args.warmup):
        step()
    tot_n, tot_t = 0, 0.0
    for _ in range(args.steps):
        n, dt = step()
        tot_n += n
        tot_t += dt
    v = tot_n / tot_t
    sample = (f"{cores} independent {args.ref_genome_mb:g} Mb genomes at {args.cov:g}x (same read model), one reference process per core, "
              f"T=1 each; {tot_n // args.steps} overlaps per step")
    print(json.dumps({
        "impl": "reference", "metric": "overlaps/s (index+overlap)", "value": v, "unit": "overlaps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "read_bases_per_s": bases * args.steps / tot_t,
        "config": {"workload": f"synthetic 30x 15 kb reads @99.5%, k={PARAMS['k']} w={PARAMS['w']} r=6 l=2, T=1 (bounded sample of configs[1])", "sample": sample},
        "cpu_baseline": {"value": v, "unit": "overlaps/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "overlaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import gc

    import torch

    gc.disable()  # a generation-2 collection inside a 60 ms step is a 30 ms outlier (seen as 1 step in 12); nothing here needs the collector

    from peregrine_b200 import Engine, formats as F

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        # NCCL prints its version banner on STDOUT when the first communicator is created; stdout carries the one JSON
        # line, so native writes to fd 1 go to stderr until the result is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    genome = int(args.genome_mb * 1e6) * world
    if world > 1:
        import faulthandler

        faulthandler.dump_traceback_later(240, repeat=True, file=sys.stderr)  # a stuck rank shows where it is stuck
        killer = threading.Timer(1500, lambda: os._exit(3))  # ... and does not hold 8 GPUs forever
        killer.daemon = True
        killer.start()
    # Every rank generates ITS share of the read set (rid % N == (rank+1) % N, src/shmr_index.c:157) straight from the
    # counter-based simulator: read i depends on (seed, i) only, so the shares are exactly the selection of the full set,
    # and no rank ever writes or scans the whole 1.5 GB x N image.
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))  # simreads (torchrun pins it to 1)
    prefix = make_dataset(f"g{genome}" if world == 1 else f"g{genome}_c{rank + 1}of{world}", genome, args.cov, mod=world, res=(rank + 1) % world)
    rid, ln, off = F.read_idx(prefix + ".idx")
    nbytes = os.path.getsize(prefix + ".seqdb")
    if nbytes != int(ln.sum()):
        raise RuntimeError(f"{prefix}.seqdb has {nbytes} bytes, the .idx table says {int(ln.sum())} (disk full while generating?)")
    pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    seqdb = pinned.numpy()
    with open(prefix + ".seqdb", "rb") as f:
        f.readinto(memoryview(seqdb))
    if world > 1:  # a rank's share is only ever read here: do not leave N x 1.5 GB per scaling point in the work directory
        import shutil

        shutil.rmtree(os.path.dirname(prefix), ignore_errors=True)
    bases = int(ln.sum())
    if world > 1:
        t = torch.tensor([bases], dtype=torch.int64, device=f"cuda:{local}")
        dist.all_reduce(t)
        bases = int(t.item())
    P = PARAMS
    T = world
    c = rank + 1
    sampler = ClockSampler(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    if world == 1:
        eng = Engine(local)
        engines = [eng]

        def compute(copy):
            eng.index(P["w"], P["k"], P["r"], P["levels"], 0)
            eng.set_shimmers_from_index(2)
            return eng.overlap(1, 1, P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"], copy=copy)

        def e2e_step():  # host buffers in (2-bit image as this library's shmr_mkseqdb writes it), host records out
            eng.load_reads_2bit(words2b, nrec2b, rid, ln, 1, 1, defer=True)  # H2D overlapped with sketching inside index()
            return len(compute(copy="view"))  # records land in the engine's page-locked host buffer

        def e2e_step_seqdb():  # the same from the reference's own 1-byte/base .seqdb image
            eng.load_reads(seqdb, rid, ln, off, 1, 1, keep_raw=False, defer=True)  # H2D overlapped with pack + sketch inside index()
            return len(compute(copy="view"))

        def dev_prepare():
            eng.load_reads(seqdb, rid, ln, off, 1, 1, keep_raw=True)

        def dev_step():  # .seqdb image resident in HBM, records stay in HBM
            eng.repack()
            return compute(copy=False)
    else:
        from peregrine_b200 import multigpu as M

        idx_eng, ovl_eng = Engine(local), Engine(local)
        engines = [idx_eng, ovl_eng]
        job = M.ShardedJob(idx_eng, ovl_eng, rank, world, torch.device("cuda", local))

        def exchange():
            if args.exchange == "routed":  # SHIMMER-pair records travel: one NCCL all-to-all (north_star / SURVEY 8e)
                job.index_and_route(P["w"], P["k"], P["r"], P["mc_lower"], P["mc_upper"])
            else:  # every rank gathers all SHIMMER lists and rescans them
                job.index_and_exchange(P["w"], P["k"], P["r"])

        def e2e_step():  # this rank's share of the 2-bit image from pinned host memory, its chunk's records back to the host
            idx_eng.load_reads_2bit(words2b, nrec2b, rid, ln, 1, 1, defer=True)
            exchange()
            return len(job.overlap(P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"], copy="view"))

        def e2e_step_seqdb():
            idx_eng.load_reads(seqdb, rid, ln, off, 1, 1, keep_raw=False, defer=True)
            exchange()
            return len(job.overlap(P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"], copy="view"))

        def dev_prepare():
            idx_eng.load_reads(seqdb, rid, ln, off, 1, 1, keep_raw=True)

        def dev_step():
            idx_eng.repack()
            exchange()
            return job.overlap(P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"], copy=False)

    def all_stats():
        out = {}
        for e in engines:
            for k_, v in e.stats().items():
                out[k_] = out.get(k_, 0) + v
        return out

    # the 2-bit image of this rank's reads, in page-locked memory: what this library's shmr_mkseqdb leaves next to the .seqdb
    # (<prefix>.seq2b / .seq2n); produced once, outside every timed region, like the .seqdb itself
    nw_total = int(((ln.astype(np.int64) + 31) // 32).sum())
    pinned_w = torch.empty(nw_total, dtype=torch.int64, pin_memory=True)
    words2b, nrec2b = engines[0].pack_2bit(seqdb, off, ln, words_out=pinned_w.numpy().view(np.uint64))

    e2e_allocs = []  # cudaMalloc calls / fix-point restarts inside the timed steps of the two loops (0 / 0 in a steady-state job)
    e2e_steps = []  # rank 0's per-step host times of the two end-to-end loops (a stall in one step shows here)

    def time_e2e(step):  # (this also warms the allocator pool)
        for _ in range(max(args.warmup, 1)):
            n = step()
        for e in engines:
            e.stats_reset()
        barrier()
        t0 = time.perf_counter()
        per_step = []
        for _ in range(args.steps):
            t1 = time.perf_counter()
            n = step()
            per_step.append(round((time.perf_counter() - t1) * 1e3, 2))
        barrier()
        sec = max_over_ranks((time.perf_counter() - t0) / args.steps)
        st_ = all_stats()
        e2e_steps.append(per_step)
        e2e_allocs.append({"device_mallocs": int(st_["n_device_mallocs"]), "replay_restarts": int(st_["n_replay_restarts"])})
        return n, sec, int(sum_over_ranks(st_["h2d_bytes"])), int(sum_over_ranks(st_["d2h_bytes"]))  # bytes of the whole job

    # ---- end-to-end: from the reference's 1-byte/base image, then from the 2-bit image (the headline `e2e`)
    n_mine_db, e2e_db_s, h2d_db, d2h_db = time_e2e(e2e_step_seqdb)
    n_mine, e2e_s, h2d_all, d2h_all = time_e2e(e2e_step)
    assert n_mine == n_mine_db
    n_ovl = int(sum_over_ranks(n_mine))

    # ---- device-resident; CUDA events on the library's stream (+ wall clock, which includes the NCCL exchange at N > 1)
    dev_prepare()
    for _ in range(args.warmup):
        dev_step()
    sampler.start()
    for e in engines:
        e.stats_reset()
    barrier()
    engines[-1].event_record(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_dev = dev_step()
    engines[-1].event_record(1)
    barrier()
    dev_wall = max_over_ranks((time.perf_counter() - t0) / args.steps)
    dev_ms = max_over_ranks(engines[-1].event_elapsed_ms(0, 1) / args.steps)
    clocks = sampler.stop()
    st = all_stats()
    for e in engines:
        e.close()
    assert int(sum_over_ranks(n_dev)) == n_ovl
    bases_total = bases
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (HBM-bound integer work; algorithmic bytes per SURVEY 8d / DESIGN.md)
    peak, peak_src = measured_peaks()
    K = args.steps
    kern = {
        "k_sketch_tiled": (st["ms_k_sketch_tiled"], st["n_k_sketch_tiled"], (st["bases_sketched"] / 4.0 + 16.0 * st["n_l0"]) / max(st["n_k_sketch_tiled"], 1)),  # (rank 0's own reads)
        "k_align": (st["ms_k_align"], st["n_k_align"], (st["n_align_bases"] / 4.0 + 64.0 * st["n_alignments"]) / max(st["n_k_align"], 1)),
        # n_candidates = candidate pairs of the eligible buckets, counted once per step: per launch = total over the steps / launches
        "k_replay": (st["ms_k_replay"], st["n_k_replay"], (16.0 * st["n_candidates"] + 9.0 * st["n_pair_records"]) / max(st["n_k_replay"], 1)),
    }
    top = max(kern, key=lambda k_: kern[k_][0])
    ms_tot, n_l, bytes_per_launch = kern[top]
    avg_ms = ms_tot / max(n_l, 1)
    achieved = bytes_per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    traffic = None
    try:  # DRAM bytes of that kernel from the committed ncu capture, per launch like `achieved`
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if top in tj and n_l:
            traffic = tj[top]["dram_bytes_per_step"] / (n_l / K)
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_ms, "launches_per_step": n_l / K,
                "kernel_ms_per_step": {**{k_: v[0] / K for k_, v in kern.items()},
                                       # the exact automaton on the reads the fast sketch kernel hands back (automaton pass + placement)
                                       "k_sketch_exact": (st["ms_k_sketch_count"] + st["ms_k_sketch_write"]) / K}}
    # the same figure for each of the three heavy kernels (north_star asks for mm_sketch and the chaining kernels, not only the top
    # one); `limiter` = what the committed ncu captures (profiles/) show each kernel is actually bound by
    limiter = {"k_sketch_tiled": "integer ALU pipe (ncu: 65 % of the ALU pipe's peak, issue active 49 %; 117 thread-instructions per base; DRAM reads 0.255 B/base = the packed words once)",
               "k_align": "dependent-load latency + divergence (ncu: issue active 54 %, 14.7 of 32 lanes live, 36 % occupancy, DRAM 6 % of peak)",
               "k_replay": "latency of dependent random probes into a 1.5 GB table (ncu: 6 of 32 lanes live, DRAM 0.45 TB/s = 7 % of peak in full passes)"}
    rooflines = {}
    for k_, (ms_k, n_k, b_k) in kern.items():
        a_ms = ms_k / max(n_k, 1)
        ach = b_k / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
        rooflines[k_] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_launch": b_k,
                         "avg_launch_ms": a_ms, "launches_per_step": n_k / K, "limiter": limiter[k_]}
    cpu, parity = cpu_baseline_single_core(device=local) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else (None, None)
    out = {
        "metric": "overlaps/s (index+overlap)", "value": n_ovl / (dev_ms * 1e-3), "unit": "overlaps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "read_bases_per_s": bases / (dev_ms * 1e-3),
        "config": {"workload": f"synthetic {args.genome_mb * world:g} Mb genome, {args.cov:g}x 15 kb reads @99.5%, k={P['k']} w={P['w']} r=6 l=2, T={T}",
                   "reads_per_rank": int(len(rid)), "bases": bases, "overlaps_per_step": int(n_ovl), "l2_flush": "inputs (1.5 GB image) exceed the 126 MB L2",
                   "wall_ms_per_step_device_resident": dev_wall * 1e3,
                   "sharding": ("reads by rid % N for the index, SHIMMER-hash chunk c of T=N for the overlap; " +
                                ("NCCL all-to-all of SHIMMER-pair records to the owning chunk + all-gather of packed reads and partial count tables" if args.exchange == "routed"
                                 else "NCCL all-gather of packed reads + L2 lists")) if world > 1 else "single GPU"},
        "e2e": {"value": n_ovl / e2e_s, "unit": "overlaps/s", "h2d_bytes_per_step": h2d_all // K, "d2h_bytes_per_step": d2h_all // K,
                "ms_per_step": e2e_s * 1e3, "read_bases_per_s": bases / e2e_s,
                "input": "pinned host 2-bit image (<prefix>.seq2b as written by this library's shmr_mkseqdb) -> ovlp_t records in pinned host memory",
                "steps_ms_rank0": e2e_steps[1], **e2e_allocs[1]},
        "e2e_seqdb": {"value": n_ovl / e2e_db_s, "unit": "overlaps/s", "h2d_bytes_per_step": h2d_db // K, "d2h_bytes_per_step": d2h_db // K,
                      "ms_per_step": e2e_db_s * 1e3, "input": "pinned host 1-byte/base .seqdb image (the reference's own format)",
                      "steps_ms_rank0": e2e_steps[0], **e2e_allocs[0]},
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": clocks,
        "roofline": roofline,
        "rooflines": rooflines,
        "stage_ms_per_step": {k_: st[k_] / K for k_ in st if k_.startswith("ms_") and not k_.startswith("ms_k_")},
        "counts_per_step": {k_: st[k_] // K for k_ in ("n_l0", "n_l1", "n_l2", "n_pair_records", "n_buckets", "n_eligible_buckets", "n_candidates",
                                                         "n_alignments", "n_replay_passes", "n_sketch_fallback_reads")},
    }
    if cpu:
        out["cpu_baseline"] = cpu
        out["parity"] = parity
    if world > 1:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["identical"]:
        print("bench.py: PARITY FAILURE: this engine's records differ from the reference's on the cpu_baseline sample", file=sys.stderr)
        sys.exit(4)


# ------------------------------------------------------------------------------------------------ our arm, T chunks on N GPUs
def run_chunked(args):
    """BASELINE.json configs[2]/[3]-style jobs: ONE genome of --genome-mb (total), T = --chunks index chunks and T hash chunks
    (README.md:155-165 runs human data as 24 / 24 chunks) on N GPUs, T a multiple of N.  Rank r owns index chunks and hash
    chunks c = r + 1 + j N.  A step = every rank sketches its index chunks -> the packed reads and the per-chunk SHIMMER lists are
    all-gathered in chunk order (what the reference's processes find on the shared file system, src/shmr_overlap.c:359-382) ->
    every rank runs build_map + process_overlaps for its hash chunks.  Records per chunk are the reference's `shmr_overlap -t T -c c`
    (tests/test_gpu_parity.py::test_sharded_exchange_matches_reference checks this path chunk by chunk)."""
    import shutil

    import torch
    import torch.distributed as dist

    from peregrine_b200 import Engine, formats as F, multigpu as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    T = args.chunks
    if T % world:
        raise SystemExit("--chunks must be a multiple of the number of GPUs")
    per = T // world
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    genome = int(args.genome_mb * 1e6)
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    mine = [rank + 1 + j * world for j in range(per)]
    shares = []
    for ci in mine:  # one share at a time: generate, read into page-locked memory, delete
        prefix = make_dataset(f"g{genome}_c{ci}of{T}", genome, args.cov, mod=T, res=ci % T)
        rid, ln, off = F.read_idx(prefix + ".idx")
        nbytes = os.path.getsize(prefix + ".seqdb")
        if nbytes != int(ln.sum()):
            raise RuntimeError(f"{prefix}.seqdb is truncated (disk full?)")
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        with open(prefix + ".seqdb", "rb") as f:
            f.readinto(memoryview(pinned.numpy()))
        shutil.rmtree(os.path.dirname(prefix), ignore_errors=True)
        shares.append((pinned, rid, ln, off))
    bases_mine = sum(int(s_[2].sum()) for s_ in shares)
    P = PARAMS
    idx_eng, ovl_eng = Engine(local), Engine(local)
    sampler = ClockSampler(local)

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    free_min = [torch.cuda.mem_get_info(dev)[0]]

    def step():
        parts, lists = [], []
        for pinned, rid, ln, off in shares:
            idx_eng.load_reads(pinned.numpy(), rid, ln, off, 1, 1, keep_raw=False, defer=True)
            idx_eng.index(P["w"], P["k"], P["r"], P["levels"], 0)
            parts.append(M.export_reads(idx_eng, dev))
            lists.append(M.export_level(idx_eng, 2, dev))
        M._handoff(dev)
        reads, l2_all = M.gather_chunks(parts, lists)  # straight into buffers in chunk order 1..T
        del parts, lists
        M._handoff(dev)
        M.import_reads(ovl_eng, reads)
        M.import_shimmers(ovl_eng, l2_all)
        del reads, l2_all
        torch.cuda.empty_cache()  # the gathered copies go back to the driver before the overlap stage sizes its tables
        free_min[0] = min(free_min[0], torch.cuda.mem_get_info(dev)[0])
        n = 0
        for c_ in mine:
            n += len(ovl_eng.overlap(T, c_, P["bestn"], P["mc_lower"], P["mc_upper"], P["bw"], P["ovlp_upper"], copy="view"))
            free_min[0] = min(free_min[0], torch.cuda.mem_get_info(dev)[0])
        return n

    for _ in range(max(args.warmup, 1)):
        n_mine = step()
    for e in (idx_eng, ovl_eng):
        e.stats_reset()
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_mine = step()
    barrier()
    sec = allmax((time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop()
    st = {}
    for e in (idx_eng, ovl_eng):
        for k_, v in e.stats().items():
            st[k_] = st.get(k_, 0) + v
    n_ovl = int(allsum(n_mine))
    bases = int(allsum(bases_mine))
    n_aln = int(allsum(st["n_alignments"])) // args.steps
    launches = int(st["kernel_launches"])
    h2d, d2h = int(allsum(st["h2d_bytes"])) // args.steps, int(allsum(st["d2h_bytes"])) // args.steps
    idx_eng.close()
    ovl_eng.close()
    if rank == 0:
        K = args.steps
        out = {
            "metric": "overlaps/s (index+overlap)", "value": n_ovl / sec, "unit": "overlaps/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "read_bases_per_s": bases / sec,
            "config": {"workload": f"synthetic {args.genome_mb:g} Mb genome, {args.cov:g}x 15 kb reads @99.5%, k={P['k']} w={P['w']} r=6 l=2, T={T} index chunks and {T} "
                                   f"hash chunks on {world} GPU(s), {per} of each per GPU",
                       "bases": bases, "overlaps_per_step": n_ovl, "alignments_per_step": n_aln,
                       "timing": "host clock around barrier + synchronize (the step spans two engines and the NCCL streams); inputs are host buffers: "
                                 "the same number is the end-to-end figure",
                       "l2_flush": "inputs exceed the 126 MB L2", "hbm_free_min_gb_rank0": round(free_min[0] / 1e9, 1),
                       "sharding": "index chunks by rid % T, hash chunks by (hash % T); all-gather of packed reads + per-chunk SHIMMER lists in chunk order"},
            "e2e": {"value": n_ovl / sec, "unit": "overlaps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": sec * 1e3,
                    "read_bases_per_s": bases / sec},
            "gpu_launches": launches, "clocks": clocks,
        }
        out["stage_ms_per_step"] = {k_: st[k_] / K for k_ in st if k_.startswith("ms_") and not k_.startswith("ms_k_")}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours_guarded(args):
    """A rank that dies must say why: torchrun's summary only has the exit code (SCALE_r01: "rank 7 exitcode 1, error_file
    <N/A>").  The traceback and the library's last error go to stderr, a one-line JSON error record goes to stdout (rank 0)
    or stderr (other ranks), and torch's elastic error file is written by `record`."""
    import traceback

    try:
        from torch.distributed.elastic.multiprocessing.errors import record
    except Exception:  # pragma: no cover
        def record(f):
            return f

    @record
    def go():
        try:
            if args.chunks:
                run_chunked(args)
            else:
                run_ours(args)
        except BaseException as e:  # noqa: BLE001 - report, then re-raise for torchrun
            if isinstance(e, SystemExit):
                raise
            rank = os.environ.get("RANK", "0")
            tb = traceback.format_exc()
            sys.stderr.write(f"\n[bench.py rank {rank}] FAILED: {type(e).__name__}: {e}\n{tb}\n")
            sys.stderr.flush()
            line = json.dumps({"error": f"{type(e).__name__}: {e}", "rank": int(rank), "n_gpus": args.gpus, "traceback_tail": tb[-1500:]})
            print(line, file=sys.stdout if rank == "0" else sys.stderr, flush=True)
            raise

    go()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=50.0, help="genome size per GPU (BASELINE.json configs[1]: 50 Mb)")
    ap.add_argument("--cov", type=float, default=30.0)
    ap.add_argument("--chunks", type=int, default=0, help="T index chunks and T hash chunks of ONE genome of --genome-mb (total) on the N GPUs "
                                                          "(strong-scaling / configs[2..3] mode; T must be a multiple of N)")
    ap.add_argument("--ref-genome-mb", type=float, default=4.0, help="reference arm: genome size of each per-core sample")
    ap.add_argument("--k", type=int, default=16, help="k-mer size (BASELINE.json configs[4] sweeps 14/16/18; k > 16 takes the 64-bit sketch kernel)")
    ap.add_argument("--w", type=int, default=80, help="minimizer window (configs[4]: 60/80/120)")
    ap.add_argument("--aln-bw", type=int, default=100, help="ovlp_match band tolerance, shmr_overlap -w (configs[4]: 50/100/200)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="routed", choices=["routed", "gathered"], help="multi-GPU exchange step (N > 1)")
    args = ap.parse_args()
    PARAMS.update(k=args.k, w=args.w, bw=args.aln_bw)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours_guarded(args)


if __name__ == "__main__":
    main()
