"""Stage benchmarks of the tools next to the index/overlap path: shmr_mkseqdb's encode_biseq, shmr_dedup, shmr_map.
Development tool (not bench.py's contract): prints one JSON line per stage with the GPU stage time (CUDA events inside the
library), the algorithmic-bytes roofline fraction, and the unmodified reference binary timed on one host core on the same input.

    python tools/bench_stages.py [genome_bp=20e6] [cov=30]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import datasets as D  # noqa: E402
from peregrine_b200 import Engine, formats as F  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
genome = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30
wd = os.environ.get("PGB_WORK", "/tmp/pgb_stages")
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError, ValueError):
    pass


def timed(cmd, **kw):
    t = time.perf_counter()
    r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, **kw)
    return r.stdout, time.perf_counter() - t


p = D.make_sim(wd, f"g{genome}", genome=genome, cov=cov)
rid, ln, off = F.read_idx(p + ".idx")
seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
eng = Engine(0)

# ---- overlap records of the set (input of dedup), produced by the engine itself
eng.load_reads(seqdb, rid, ln, off)
eng.index(80, 16, 6, 2)
eng.set_shimmers_from_index(2)
ov = eng.overlap(1, 1)
stream = np.concatenate([ov, ov[::-1]])  # every pair twice, second time in reverse stream order
for _ in range(2):
    eng.stats_reset()
    text = eng.dedup(stream)
    st = eng.stats()
ref_text, ref_s = timed([os.path.join(REF, "shmr_dedup")], input=stream.tobytes())
assert text == ref_text
alg = 64.0 * len(stream) + len(text)
print(json.dumps({"stage": "shmr_dedup", "records_in": int(len(stream)), "lines": int(st["n_dedup_kept"]), "gpu_ms": st["ms_dedup"],
                  "records_per_s": len(stream) / (st["ms_dedup"] * 1e-3), "algorithmic_bytes": alg,
                  "roofline_frac": alg / (st["ms_dedup"] * 1e-3) / 1e9 / peak, "reference_1core_s": ref_s,
                  "reference_records_per_s": len(stream) / ref_s}))

# ---- encode_biseq: ASCII of the same reads (decoded from the low nibbles) -> .seqdb bytes
lut = np.zeros(256, dtype=np.uint8)
lut[[1, 2, 4, 8]] = np.frombuffer(b"ACGT", dtype=np.uint8)
lut[0] = ord("N")
ascii_ = lut[seqdb & 0x0F]
for _ in range(2):
    eng.stats_reset()
    enc = eng.encode_biseq(ascii_, off, ln)
    st = eng.stats()
assert np.array_equal(enc, seqdb)
fa = os.path.join(wd, "stage.fa")
if not os.path.exists(fa):
    with open(fa, "wb") as f:
        for i in range(len(rid)):
            f.write(b">r%d\n" % i + ascii_[int(off[i]): int(off[i]) + int(ln[i])].tobytes() + b"\n")
with open(os.path.join(wd, "stage.lst"), "w") as f:
    f.write(fa + "\n")
_, ref_s = timed([os.path.join(REF, "shmr_mkseqdb"), "-d", os.path.join(wd, "stage.lst"), "-p", os.path.join(wd, "stage_ref")])
_, our_s = timed([os.path.join(ROOT, "bin", "shmr_mkseqdb"), "-d", os.path.join(wd, "stage.lst"), "-p", os.path.join(wd, "stage_our")])
alg = 2.0 * seqdb.size
print(json.dumps({"stage": "shmr_mkseqdb", "bases": int(seqdb.size), "k_encode_biseq_ms": st["ms_k_encode"], "encode_call_ms_incl_copies": st["ms_encode"],
                  "kernel_GBps": alg / (st["ms_k_encode"] * 1e-3) / 1e9, "roofline_frac": alg / (st["ms_k_encode"] * 1e-3) / 1e9 / peak,
                  "tool_wall_s": our_s, "reference_tool_wall_s_1core": ref_s}))

# ---- shmr_map: the longest reads as "contigs" against the index of all reads
eng.stats_reset()
l2 = eng.level(2)
order = np.argsort(-ln.astype(np.int64))[:200]
ctg_rids = set(int(rid[i]) for i in order)
ref_mm = l2[np.isin((l2["y"] >> np.uint64(32)).astype(np.int64), list(ctg_rids))]
mcount = eng.level_counts(2) if hasattr(eng, "level_counts") else None
eng.index(80, 16, 6, 2, 4)
mc = eng.level_counts(2)
eng.set_shimmers(l2, mc)
for _ in range(2):
    eng.stats_reset()
    hits = eng.map(ref_mm)
    st = eng.stats()
print(json.dumps({"stage": "shmr_map", "contig_shimmers": int(len(ref_mm)), "read_shimmers": int(len(l2)), "hits": int(st["n_map_hits"]),
                  "gpu_ms_after_pair_records": st["ms_map"], "pair_records_ms": st["ms_pairs"], "text_bytes": len(hits)}))
eng.close()
