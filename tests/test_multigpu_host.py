"""CPU tests of the multi-GPU host logic: world_size-2 gloo run of the exchange (variable-size all-gather, concatenation
order, word-offset rebasing).  The GPU-side equivalent (three ranks simulated on one device, compared with the reference's
T=3 chunk files) is tests/test_gpu_parity.py::test_sharded_exchange_matches_reference."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from peregrine_b200 import multigpu as M


def _part(rank):
    g = torch.Generator().manual_seed(100 + rank)
    n_rows = 3 + rank
    lens = torch.randint(1, 200, (n_rows,), generator=g, dtype=torch.int32)
    words_per = (lens.to(torch.int64) + 31) // 32
    woff = torch.cumsum(torch.cat([torch.tensor([2]), words_per[:-1]]), 0)
    n_words = int(words_per.sum()) + 4
    return {
        "words": torch.randint(0, 2**62, (n_words,), generator=g, dtype=torch.int64),
        "nmask": torch.randint(0, 2**31 - 1, (n_words,), generator=g, dtype=torch.int32),
        "row_rid": (torch.arange(n_rows, dtype=torch.int32) * 2 + rank),
        "row_len": lens,
        "row_woff": woff.to(torch.int64),
        "row_hasn": torch.zeros(n_rows, dtype=torch.int32),
    }


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = _part(rank)
    reads = M.exchange_reads(part)
    l2 = torch.arange(2 * (5 + 3 * rank), dtype=torch.int64).reshape(-1, 2) + 1000 * rank
    l2_all = M.exchange_shimmers(l2)
    # routed exchange: rank r sends (r + 1) * (d + 2) rows to rank d, rows tagged (source, destination, running number)
    sizes = [(rank + 1) * (d + 2) for d in range(world)]
    send = torch.tensor([[rank, d, i, 0, 0] for d in range(world) for i in range(sizes[d])], dtype=torch.int64).reshape(-1, 5)
    recv, in_sizes = M.all_to_all_var(send, sizes)
    before = M.first_found_before(rank == 1, torch.device("cpu"))
    # cross-rank record stream for shmr_dedup: rank r holds 2 + r records, tagged (rank, i)
    recs = torch.tensor([[rank, i, 0, 0, 0, 0, 0, 0] for i in range(2 + rank)], dtype=torch.int64).reshape(-1, 8)
    stream = M.gather_records(recs)
    assert stream[:, :2].tolist() == [[r, i] for r in range(world) for i in range(2 + r)]  # chunk order, stream order inside
    # several chunks per rank (T = 2 x world): chunk c = rank + 1 + j * world; the gather must come back in chunk order 1..T
    cparts = [_part(10 * j + rank) for j in range(2)]
    clists = [torch.full((3 + j + rank, 2), 100 * j + rank, dtype=torch.int64) for j in range(2)]
    creads, cl2 = M.gather_chunks(cparts, clists)
    want_c = M.concat_reads([_part(10 * j + r) for j in range(2) for r in range(world)])
    for k in M.READ_KEYS:
        assert torch.equal(creads[k], want_c[k]), k
    assert cl2[:, 0].tolist() == [100 * j + r for j in range(2) for r in range(world) for _ in range(3 + j + r)]
    q.put((rank, {k: v.numpy() for k, v in reads.items()}, l2_all.numpy(), recv.numpy(), in_sizes, before))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_with_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [_part(0), _part(1)]
    want = M.concat_reads(parts)
    for rank, _, _, recv, in_sizes, before in got:
        assert in_sizes == [(src + 1) * (rank + 2) for src in range(2)]
        want_rows = [[src, rank, i, 0, 0] for src in range(2) for i in range((src + 1) * (rank + 2))]  # source-rank order, scan order inside
        assert recv.tolist() == want_rows
        assert before is False  # only rank 1 holds a "first" element, and no rank lies after it
    for rank, reads, l2_all, *_ in got:
        for k in M.READ_KEYS:
            assert np.array_equal(reads[k], want[k].numpy()), (rank, k)
        # every row's words are where row_woff says they are, in the concatenated array
        base = 0
        row = 0
        for p in parts:
            for i in range(len(p["row_rid"])):
                nw = (int(p["row_len"][i]) + 31) // 32
                a = reads["words"][reads["row_woff"][row]: reads["row_woff"][row] + nw]
                b = p["words"].numpy()[int(p["row_woff"][i]): int(p["row_woff"][i]) + nw]
                assert np.array_equal(a, b)
                row += 1
            base += len(p["words"])
        assert l2_all.shape == (8 + 5, 2) and l2_all[0, 0] == 0 and l2_all[5, 0] == 1000  # chunk order = rank order
