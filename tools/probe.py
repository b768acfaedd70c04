"""Development probe: run the engine once or a few times on a synthetic set and print stage timings."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import datasets as D
from peregrine_b200 import Engine, formats as F

genome = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wd = os.environ.get("PGB_WORK", "/tmp/pgb_probe")
t0 = time.time()
p = D.make_sim(wd, f"g{genome}", genome=genome, cov=cov)
rid, ln, off = F.read_idx(p + ".idx")
seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
print(f"dataset: {len(rid)} reads {seqdb.size} bases ({time.time()-t0:.1f}s)", flush=True)
eng = Engine(0)
for it in range(reps):
    eng.stats_reset()
    t = time.time(); eng.load_reads(seqdb, rid, ln, off); t_load = time.time() - t
    t = time.time(); eng.index(80, 16, 6, 2); t_idx = time.time() - t
    t = time.time(); eng.set_shimmers_from_index(2); n = eng.overlap(1, 1, copy=False); t_ov = time.time() - t
    st = eng.stats()
    print(json.dumps({"iter": it, "wall_load_s": round(t_load, 4), "wall_index_s": round(t_idx, 4), "wall_overlap_s": round(t_ov, 4),
                      "overlaps": n, **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}}), flush=True)
eng.close()
