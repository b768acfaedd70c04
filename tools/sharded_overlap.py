#!/usr/bin/env python
"""One process per GPU (torchrun): index + overlap of ONE read set sharded over the ranks, T = world size, rank r owns index
chunk r+1 (rid % T) and hash chunk r+1 — the job `shmr_index -t T -c c` x T followed by `shmr_overlap -t T -c c` x T does on a
shared file system, with the exchange done by peregrine_b200.multigpu.ShardedJob over NCCL.  Writes <out>/ovlp.CC, the same
raw ovlp_t stream shmr_overlap writes for chunk CC (used by tests/test_zz_gpu_scale.py to diff a real multi-process run against
the reference).

    python -m torch.distributed.run --nproc-per-node N tools/sharded_overlap.py --prefix <seqdb prefix> --out <dir> [--exchange routed|gathered]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--prefix", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--exchange", default="routed", choices=["routed", "gathered"])
    ap.add_argument("--backend", default="nccl")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    from peregrine_b200 import Engine, formats as F, multigpu as M

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    os.dup2(2, 1)  # NCCL's banner goes to stderr
    torch.cuda.set_device(local)
    dist.init_process_group(a.backend, device_id=torch.device("cuda", local))
    rid, ln, off = F.read_idx(a.prefix + ".idx")
    seqdb = np.fromfile(a.prefix + ".seqdb", dtype=np.uint8)
    idx_eng, ovl_eng = Engine(local), Engine(local)
    idx_eng.load_reads(seqdb, rid, ln, off, world, rank + 1)
    job = M.ShardedJob(idx_eng, ovl_eng, rank, world, torch.device("cuda", local))
    if a.exchange == "routed":
        job.index_and_route(80, 16, 6, 2, 240)
    else:
        job.index_and_exchange(80, 16, 6)
    ov = job.overlap(4, 2, 240, 100, 120, copy=True)
    os.makedirs(a.out, exist_ok=True)
    ov.tofile(os.path.join(a.out, f"ovlp.{rank + 1:02d}"))
    idx_eng.close()
    ovl_eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
