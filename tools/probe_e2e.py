"""Development probe: where does the end-to-end step (pinned host .seqdb -> records on the host) spend its wall time?"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import datasets as D
from peregrine_b200 import Engine, formats as F

genome = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
wd = os.environ.get("PGB_WORK", "/tmp/pgb_bench")
p = D.make_sim(wd, f"g{genome}", genome=genome, cov=30)
rid, ln, off = F.read_idx(p + ".idx")
nbytes = os.path.getsize(p + ".seqdb")
pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
seqdb = pinned.numpy()
with open(p + ".seqdb", "rb") as f:
    f.readinto(memoryview(seqdb))
# raw pinned H2D rate of this box
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); dev.copy_(pinned, non_blocking=True); torch.cuda.synchronize()
    print(f"H2D {nbytes/1e9:.2f} GB pinned: {1e3*(time.perf_counter()-t):.1f} ms = {nbytes/1e9/(time.perf_counter()-t):.1f} GB/s", flush=True)
del dev
for chunk in sys.argv[2:] or ["96"]:
    os.environ["PGB_LOAD_CHUNK_MB"] = chunk
    eng = Engine(0)
    for it in range(5):
        eng.stats_reset()
        t0 = time.perf_counter(); eng.load_reads(seqdb, rid, ln, off, 1, 1, keep_raw=False, defer=(chunk != "0")); t1 = time.perf_counter()
        eng.index(80, 16, 6, 2, 0); t2 = time.perf_counter()
        eng.set_shimmers_from_index(2); t3 = time.perf_counter()
        n = eng.overlap(1, 1, copy=False); t4 = time.perf_counter()
        recs = eng.overlap_records(view=True); t5 = time.perf_counter()
        st = eng.stats()
        print(json.dumps({"chunk_mb": chunk, "iter": it, "load": round(1e3*(t1-t0), 1), "index": round(1e3*(t2-t1), 1), "set": round(1e3*(t3-t2), 1),
                          "overlap": round(1e3*(t4-t3), 1), "d2h": round(1e3*(t5-t4), 1), "total": round(1e3*(t5-t0), 1),
                          "ms_sketch": round(st["ms_sketch"], 1), "ms_k_sketch_tiled": round(st["ms_k_sketch_tiled"], 1),
                          "ms_host_order": round(st["ms_host_order"], 1), "ms_replay": round(st["ms_replay"], 1), "ms_align": round(st["ms_align"], 1)}), flush=True)
    eng.close()
