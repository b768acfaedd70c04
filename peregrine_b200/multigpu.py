"""One process per GPU: shard the SHIMMER index by read id and the overlap by SHIMMER-hash chunk, exactly as the reference
shards its processes (rid % T for shmr_index, src/shmr_index.c:157; (hash % T) for shmr_overlap, src/shmr_utils.c:337,362),
with T = world size and rank r owning chunk r + 1.

The reference's "exchange" is the shared file system: every shmr_overlap process reads ALL L2 chunk files and mmaps the whole
.seqdb (src/shmr_overlap.c:359-382,200).  Here it is one collective step over NVLink (torch.distributed / NCCL), in two forms:

  routed (ShardedJob.index_and_route, the default of bench.py; BASELINE.json north_star / SURVEY 8e):
  * all-gather + sum of the per-rank partial multiplicity tables (aggregate_mm_count over the -MC- files, shmr_utils.c:162-176),
  * every rank runs build_map over the shimmers of ITS reads and emits the SHIMMER-pair records of every hash chunk; ONE
    all-to-all sends each 40-byte record to the rank that owns its chunk; records received from ranks 0..N-1 in rank order
    are exactly the insertion order of the reference's scan over the concatenated chunk files,
  * all-gather of the 2-bit packed reads + read table (so that the owner can align any pair of reads).

  gathered (ShardedJob.index_and_exchange): all-gather of the packed reads and of the per-chunk SHIMMER lists in chunk order;
  every rank then scans the whole list and keeps the records of its chunk (closest to what the reference processes do).

After the exchange no rank needs another rank again (rid_pairs is per chunk in the reference as well).  The functions below
are device-agnostic torch code so that the layout logic is covered by gloo/CPU tests; the engine calls move bytes between
libpgb200's buffers and torch tensors.
"""
from __future__ import annotations

import torch

READ_KEYS = ("words", "nmask", "row_rid", "row_len", "row_woff", "row_hasn")


def _handoff(device=None):
    """The collectives run on torch's streams, libpgb200's kernels on the library's own (non-blocking) stream: before a tensor
    produced here is handed to the library by pointer, everything torch has queued must have finished.  (Round 1 missed this
    between the all-gather of the count tables and pgb_counts_set_device: at N = 2 the race was won, at N >= 4 the last rank
    read its table before the gather had written it - "mer missing from count table", SCALE_r01.)"""
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def all_gather_var(t: torch.Tensor, group=None, sizes=None, out=None):
    """all_gather of tensors whose first dimension differs per rank -> ONE tensor, the ranks' blocks concatenated in rank order,
    plus the per-rank row counts.  NCCL: the blocks are received straight into slices of the result (ProcessGroupNCCL handles
    uneven outputs as a group of broadcasts), nothing is padded or copied again.  gloo (CPU tests): pad to the longest."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    if sizes is None:
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        all_n = torch.empty(world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(all_n, n, group=group)
        sizes = [int(x) for x in all_n.tolist()]
    tail = tuple(t.shape[1:])
    if out is None:
        out = torch.empty((sum(sizes),) + tail, dtype=t.dtype, device=t.device)
    assert out.shape[0] == sum(sizes)  # (a caller may hand in a slice of a larger preallocated buffer)
    t = t.contiguous()
    if dist.get_backend(group) == "nccl":
        dist.all_gather(list(out.split(sizes)), t, group=group)
    else:
        m = max(sizes)
        pad = torch.zeros((m,) + tail, dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        o = 0
        for p_, s_ in zip(parts, sizes):
            out[o: o + s_] = p_[:s_]
            o += s_
    return out, sizes


def concat_reads(parts):
    """parts[r] = dict of rank r's buffers (READ_KEYS).  Concatenate in rank order; row_woff is rebased onto the
    concatenated word array (each rank's block keeps its own guard words, which is harmless)."""
    out = {}
    base = 0
    woffs = []
    for p in parts:
        woffs.append(p["row_woff"] + base)
        base += int(p["words"].shape[0])
    out["row_woff"] = torch.cat(woffs)
    for k in ("words", "nmask", "row_rid", "row_len", "row_hasn"):
        out[k] = torch.cat([p[k] for p in parts])
    return out


def exchange_reads(part, group=None):
    """all-gather of every rank's packed reads + read table; row_woff is rebased onto the concatenated word array."""
    words, wsizes = all_gather_var(part["words"], group)
    out = {"words": words}
    out["nmask"], _ = all_gather_var(part["nmask"], group, sizes=wsizes)
    out["row_rid"], rsizes = all_gather_var(part["row_rid"], group)
    for k in ("row_len", "row_hasn", "row_woff"):
        out[k], _ = all_gather_var(part[k], group, sizes=rsizes)
    base, o = 0, 0
    for ws, rs in zip(wsizes, rsizes):  # each rank's block keeps its own guard words, which is harmless
        out["row_woff"][o: o + rs] += base
        base += ws
        o += rs
    return out


def gather_chunks(parts, lists, group=None):
    """Several chunks per rank (T = per x N chunks, chunk c = r + 1 + j N held by rank r as its j-th): all-gather every rank's
    packed reads and SHIMMER lists straight into buffers laid out in chunk order 1..T - block j (chunks 1 + jN .. N + jN, rank
    order) follows block j - 1 -, without intermediate copies.  parts[j] / lists[j]: this rank's j-th chunk (export_reads /
    export_level).  Returns (reads dict with row_woff rebased onto the concatenated word array, l2_all)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = len(parts)
    dev = parts[0]["words"].device
    mine = torch.tensor([[p_["words"].shape[0], p_["row_rid"].shape[0], l_.shape[0]] for p_, l_ in zip(parts, lists)], dtype=torch.int64, device=dev)
    if world > 1:
        flat = torch.empty(world * per * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat, mine.reshape(-1).contiguous(), group=group)
        allm = flat.reshape(world, per, 3)
    else:
        allm = mine[None]
    allm = allm.cpu().numpy()  # [rank][j][kind]
    n_w = [[int(allm[r][j][0]) for r in range(world)] for j in range(per)]
    n_r = [[int(allm[r][j][1]) for r in range(world)] for j in range(per)]
    n_l = [[int(allm[r][j][2]) for r in range(world)] for j in range(per)]
    tot_w, tot_r, tot_l = sum(map(sum, n_w)), sum(map(sum, n_r)), sum(map(sum, n_l))
    out = {"words": torch.empty(tot_w, dtype=torch.int64, device=dev), "nmask": torch.empty(tot_w, dtype=torch.int32, device=dev),
           "row_rid": torch.empty(tot_r, dtype=torch.int32, device=dev), "row_len": torch.empty(tot_r, dtype=torch.int32, device=dev),
           "row_hasn": torch.empty(tot_r, dtype=torch.int32, device=dev), "row_woff": torch.empty(tot_r, dtype=torch.int64, device=dev)}
    l2_all = torch.empty((tot_l, 2), dtype=torch.int64, device=dev)
    ow = orow = ol = 0
    for j in range(per):
        sw, sr, sl = sum(n_w[j]), sum(n_r[j]), sum(n_l[j])
        for key, o_, s_, sizes in (("words", ow, sw, n_w[j]), ("nmask", ow, sw, n_w[j]), ("row_rid", orow, sr, n_r[j]), ("row_len", orow, sr, n_r[j]),
                                   ("row_hasn", orow, sr, n_r[j]), ("row_woff", orow, sr, n_r[j])):
            if world > 1:
                all_gather_var(parts[j][key], group, sizes=sizes, out=out[key][o_: o_ + s_])
            else:
                out[key][o_: o_ + s_] = parts[j][key]
        if world > 1:
            all_gather_var(lists[j], group, sizes=n_l[j], out=l2_all[ol: ol + sl])
        else:
            l2_all[ol: ol + sl] = lists[j]
        base, o2 = ow, orow
        for r in range(world):  # each chunk keeps its own guard words; rebase its rows onto the concatenated word array
            out["row_woff"][o2: o2 + n_r[j][r]] += base
            base += n_w[j][r]
            o2 += n_r[j][r]
        ow += sw
        orow += sr
        ol += sl
    return out, l2_all


def exchange_shimmers(l2: torch.Tensor, group=None):
    """l2: (n, 2) int64 view of this rank's mm128_t list.  Result: all chunks concatenated in chunk (= rank) order."""
    return all_gather_var(l2, group)[0]


def all_to_all_var(send: torch.Tensor, split_sizes, group=None):
    """send: (n, ...) rows grouped by destination rank, split_sizes[d] rows for rank d.  One all-to-all of the counts, one of
    the rows.  Returns (received rows, concatenated in source-rank order; list of per-source row counts)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    assert len(split_sizes) == world and sum(split_sizes) == send.shape[0]
    n_out = torch.tensor(list(split_sizes), dtype=torch.int64, device=send.device)
    n_in = torch.empty_like(n_out)
    dist.all_to_all_single(n_in, n_out, group=group)
    in_sizes = [int(x) for x in n_in.tolist()]
    recv = torch.empty((sum(in_sizes),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=in_sizes, input_split_sizes=list(split_sizes), group=group)
    return recv, in_sizes


def first_found_before(has_first: bool, device, group=None) -> bool:
    """build_map treats the FIRST kept element of the concatenated list specially (src/shmr_utils.c:311-321); with the list
    split by rank, a rank needs to know whether a lower rank already holds that element."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    f = torch.tensor([1 if has_first else 0], dtype=torch.int64, device=device)
    fs = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(fs, f, group=group)
    return bool(fs[:rank].any().item()) if rank else False


def gather_records(mine: torch.Tensor, group=None):
    """`cat ovlp-01.dat ovlp-02.dat ...` across ranks: every rank's ovlp_t records (an (n, 8) int64 view of the 64-byte records)
    concatenated in chunk (= rank) order, which is the stream order shmr_dedup sees (py/scripts/pg_run.py:352)."""
    return all_gather_var(mine, group)[0]


# ------------------------------------------------------------------------------------------------ engine <-> torch
def export_reads(eng, device):
    """Copy the engine's packed reads + read table into fresh torch tensors on `device` (a CUDA device)."""
    _handoff(device)  # (a recycled block of torch's allocator may still be in use by work queued on torch's stream)
    E = eng
    spec = (("words", E.BUF_WORDS, torch.int64), ("nmask", E.BUF_NMASK, torch.int32), ("row_rid", E.BUF_ROW_RID, torch.int32),
            ("row_len", E.BUF_ROW_LEN, torch.int32), ("row_woff", E.BUF_ROW_WOFF, torch.int64), ("row_hasn", E.BUF_ROW_HASN, torch.int32))
    out = {}
    for name, which, dt in spec:
        t = torch.empty(E.buffer_elems(which), dtype=dt, device=device)
        E.buffer_copy_out(which, t.data_ptr())
        out[name] = t
    return out


def export_level(eng, level, device):
    _handoff(device)
    n = eng.buffer_elems(eng.BUF_LEVEL0 + level)
    t = torch.empty((n, 2), dtype=torch.int64, device=device)
    eng.buffer_copy_out(eng.BUF_LEVEL0 + level, t.data_ptr())
    return t


def export_counts(eng, device):
    """This rank's partial multiplicity table as an (n, 2) int64 tensor of (mer, count) rows (mm_count_t, 16 bytes)."""
    _handoff(device)
    n = eng.counts_dump()
    t = torch.empty((n, 2), dtype=torch.int64, device=device)
    eng.buffer_copy_out(eng.BUF_COUNTS, t.data_ptr())
    return t


def export_route(eng, device):
    """The routed SHIMMER-pair records of pgb_route_build: (n, 5) int64 rows {x0, x1, y0, y1, direction}, grouped by owner."""
    _handoff(device)
    n = eng.buffer_elems(eng.BUF_ROUTE)
    t = torch.empty((n, 5), dtype=torch.int64, device=device)
    eng.buffer_copy_out(eng.BUF_ROUTE, t.data_ptr())
    return t


def import_reads(eng, reads):
    """Hand torch tensors (e.g. the result of concat_reads) to the engine by pointer: torch's queued work is finished first."""
    r = {k: v.contiguous() for k, v in reads.items()}
    _handoff(r["words"].device)
    eng.load_packed_device(r["words"].data_ptr(), r["nmask"].data_ptr(), r["words"].shape[0], r["row_rid"].data_ptr(), r["row_len"].data_ptr(),
                           r["row_woff"].data_ptr(), r["row_hasn"].data_ptr(), r["row_rid"].shape[0])


def import_shimmers(eng, l2_all):
    l2_all = l2_all.contiguous()
    _handoff(l2_all.device)
    eng.set_shimmers_device(l2_all.data_ptr(), l2_all.shape[0])


class ShardedJob:
    """index (own reads) -> exchange -> overlap (own hash chunk) for one rank.

    idx_eng holds the rank's own reads (index stage); ovl_eng receives the gathered read set (overlap stage).  They may be
    the same Engine when the raw image does not have to stay resident between steps."""

    def __init__(self, idx_eng, ovl_eng, rank, world, device, group=None):
        self.idx_eng, self.ovl_eng, self.rank, self.world, self.device, self.group = idx_eng, ovl_eng, rank, world, device, group
        self.routed = None  # set by index_and_route: the records this rank owns

    def index_and_exchange(self, w, k, r):
        self.routed = None
        self.idx_eng.index(w, k, r, 2, 0)
        part = export_reads(self.idx_eng, self.device)
        l2 = export_level(self.idx_eng, 2, self.device)
        _handoff(self.device)
        reads = exchange_reads(part, self.group)
        l2_all = exchange_shimmers(l2, self.group)
        _handoff(self.device)
        import_reads(self.ovl_eng, reads)
        import_shimmers(self.ovl_eng, l2_all)
        return int(reads["words"].shape[0]) * 12 + int(l2_all.shape[0]) * 16  # bytes this rank ends up holding from the exchange

    def overlap(self, bestn=4, mc_lower=2, mc_upper=240, bw=100, ovlp_upper=120, copy=True):
        if self.routed is not None:
            return self.ovl_eng.overlap_routed(self.routed.data_ptr(), int(self.routed.shape[0]), bestn, bw, ovlp_upper, copy=copy, total_chunk=self.world)
        return self.ovl_eng.overlap(self.world, self.rank + 1, bestn, mc_lower, mc_upper, bw, ovlp_upper, copy=copy)

    def dedup_all(self, dedup_eng=None):
        """`cat ovlp-*.dat | shmr_dedup` for the whole job: the ranks' record streams are gathered in chunk order and rank 0
        runs the dedup kernels over the concatenation (first-seen per read pair ACROSS chunks).  Returns the preads.ovl text
        on rank 0, None elsewhere.  Call after overlap(copy=False)."""
        E = self.ovl_eng
        n = E.L.pgb_overlap_size(E.h)
        mine = torch.empty((n, 8), dtype=torch.int64, device=self.device)
        if n:
            E.buffer_copy_out(E.BUF_OVLP, mine.data_ptr())
        _handoff(self.device)
        stream = gather_records(mine, self.group).contiguous()
        _handoff(self.device)
        if self.rank != 0:
            return None
        return (dedup_eng or E).dedup(device_ptr=stream.data_ptr(), n=int(stream.shape[0]))

    def index_and_route(self, w, k, r, mc_lower=2, mc_upper=240):
        """north_star's exchange: the SHIMMER-pair records travel.  Every rank runs build_map over its own reads' shimmers
        with GLOBAL multiplicities (all-gather + sum of the partial count tables, as aggregate_mm_count does over the -MC-
        files) and sends each record to the rank that owns its hash chunk with ONE all-to-all; the packed reads are
        all-gathered so that the owner can align any pair.  Returns the bytes this rank received."""
        E = self.idx_eng
        E.index(w, k, r, 2, 0)
        E.set_shimmers_from_index(2)
        part = export_reads(E, self.device)
        counts = export_counts(E, self.device)
        _handoff(self.device)
        all_counts, _ = all_gather_var(counts, self.group)
        _handoff(self.device)  # the library reads all_counts next, on its own stream
        E.counts_set_device(all_counts.data_ptr(), int(all_counts.shape[0]))
        has_first = E.route_scan(mc_lower, mc_upper)
        before = first_found_before(has_first, self.device, self.group)
        per_chunk = E.route_build(self.world, mc_lower, mc_upper, before)
        send = export_route(E, self.device)
        _handoff(self.device)
        self.routed, _ = all_to_all_var(send, per_chunk, self.group)  # chunk c is owned by rank c-1
        reads = exchange_reads(part, self.group)
        _handoff(self.device)
        import_reads(self.ovl_eng, reads)
        return int(reads["words"].shape[0]) * 12 + int(self.routed.shape[0]) * 40 + int(all_counts.shape[0]) * 16
