"""GPU parity tests: the CUDA path (through the drop-in CLI tools and the C ABI) against the UNMODIFIED reference
compiled into oracle/_ref, on the same inputs.  Bar: byte-identical L0/L1/L2 files, set-identical -MC- files, and
ovlp files identical field-for-field AND in order (padding bytes masked, SURVEY A-1)."""
import os
import subprocess

import numpy as np
import pytest

import datasets as D
from peregrine_b200 import formats as F

pytestmark = pytest.mark.gpu

ROOT = D.ROOT
BIN = os.path.join(ROOT, "bin")


def ours_index(prefix, outdir, T=1, extra=()):
    os.makedirs(outdir, exist_ok=True)
    for c in range(1, T + 1):
        D.run([os.path.join(BIN, "shmr_index"), "-p", prefix, "-t", str(T), "-c", str(c), "-o", os.path.join(outdir, "shmr"), *extra])
    return os.path.join(outdir, "shmr")


def ours_overlap(prefix, idx_prefix, level, outdir, T=1, extra=()):
    os.makedirs(outdir, exist_ok=True)
    outs = []
    for c in range(1, T + 1):
        o = os.path.join(outdir, f"ovlp.{c:02d}")
        D.run([os.path.join(BIN, "shmr_overlap"), "-p", prefix, "-l", f"{idx_prefix}-L{level}", "-t", str(T), "-c", str(c), "-o", o, *extra])
        outs.append(o)
    return outs


def assert_same_bytes(a, b):
    with open(a, "rb") as fa, open(b, "rb") as fb:
        da, db = fa.read(), fb.read()
    assert len(da) == len(db), f"{a} ({len(da)} B) vs {b} ({len(db)} B)"
    assert da == db, f"{a} differs from {b}"


def assert_same_mc(a, b):
    ma, mb = F.mc_as_sorted_pairs(F.read_mc(a)), F.mc_as_sorted_pairs(F.read_mc(b))
    assert ma.shape == mb.shape and np.array_equal(ma, mb), f"{a} vs {b}: count tables differ as sets"


def assert_same_ovlp(a, b):
    ra, rb = F.normalise_ovlp(F.read_ovlp(a)), F.normalise_ovlp(F.read_ovlp(b))
    assert len(ra) == len(rb), f"{a}: {len(ra)} records vs reference {len(rb)}"
    if len(ra):
        neq = np.nonzero(ra.view(np.uint8).reshape(-1, 64) != rb.view(np.uint8).reshape(-1, 64))[0]
        assert neq.size == 0, f"{a}: first differing record {neq[0]}: {ra[neq[0]]} vs {rb[neq[0]]}"


def compare_index(ref_p, our_p, T, levels=("L0", "L2")):
    for c in range(1, T + 1):
        for lv in levels:
            sfx = f"{c:02d}-of-{T:02d}.dat"
            assert_same_bytes(f"{our_p}-{lv}-{sfx}", f"{ref_p}-{lv}-{sfx}")
            assert_same_mc(f"{our_p}-{lv}-MC-{sfx}", f"{ref_p}-{lv}-MC-{sfx}")


@pytest.fixture(scope="module")
def sim1(workdir):
    return D.make_sim(workdir, "sim1", genome=1_000_000, cov=20)


def test_index_and_overlap_single_chunk(sim1, workdir, ref_dir):
    """BASELINE.json configs[0] shape: T_idx = T_ovlp = 1, k=16 w=80 r=6 l=2."""
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    op = ours_index(sim1, os.path.join(workdir, "sim1/our1"), T=1, extra=["-m", "1"])
    compare_index(rp, op, 1)
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref1"), T=1)
    oo = ours_overlap(sim1, op, 2, os.path.join(workdir, "sim1/our1"), T=1)
    assert_same_ovlp(oo[0], ro[0])
    assert os.path.getsize(ro[0]) > 64 * 1000


def test_multi_chunk(sim1, workdir, ref_dir):
    """T_idx=3, T_ovlp=2 (exercises rid % T selection, file concatenation order and c % T == 0)."""
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref3"), T=3, extra=["-m", "0"])
    op = ours_index(sim1, os.path.join(workdir, "sim1/our3"), T=3, extra=["-m", "0"])
    compare_index(rp, op, 3, levels=("L2",))
    assert not os.path.exists(f"{op}-L0-01-of-03.dat")
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref3"), T=2)
    # our overlap reads OUR index files
    oo = ours_overlap(sim1, op, 2, os.path.join(workdir, "sim1/our3"), T=2)
    for a, b in zip(oo, ro):
        assert_same_ovlp(a, b)


@pytest.mark.parametrize("k,w,r,l", [(14, 60, 6, 2), (18, 120, 3, 2), (16, 80, 36, 1), (12, 24, 2, 2), (28, 255, 6, 2)])
def test_index_parameter_sweep(sim1, workdir, ref_dir, k, w, r, l):
    tag = f"k{k}w{w}r{r}l{l}"
    ex = ["-k", str(k), "-w", str(w), "-r", str(r), "-l", str(l), "-m", "1"]
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, f"sim1/ref_{tag}"), T=1, extra=ex)
    op = ours_index(sim1, os.path.join(workdir, f"sim1/our_{tag}"), T=1, extra=ex)
    compare_index(rp, op, 1, levels=("L0", f"L{l}"))


@pytest.mark.parametrize("extra", [["-w", "50"], ["-w", "200"], ["-b", "2", "-n", "40"], ["-m", "3", "-M", "60"], ["-b", "8", "-n", "500", "-M", "1000"],
                                   ["-b", "256"]])  # (bestn is a uint8_t: 256 wraps to 0 and no candidate is visited, src/shmr_overlap.c:245)
def test_overlap_parameter_sweep(sim1, workdir, ref_dir, extra):
    tag = "_".join(x.strip("-") for x in extra)
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, f"sim1/refo_{tag}"), T=1, extra=extra)
    oo = ours_overlap(sim1, rp, 2, os.path.join(workdir, f"sim1/ouro_{tag}"), T=1, extra=extra)
    assert_same_ovlp(oo[0], ro[0])


def test_adversarial_reads(workdir, ref_dir):
    """N runs, tandem repeats / poly-A (hash ties), palindromic k-mers, reads shorter than k or than one window."""
    p = D.make_from_fasta(workdir, "adv", D.adversarial_records(), ref_dir)
    for ex in (["-m", "1"], ["-m", "1", "-k", "12", "-w", "24", "-r", "2"], ["-m", "1", "-k", "18", "-w", "120", "-r", "3"]):
        tag = "".join(ex).replace("-", "")
        rp = D.ref_index(ref_dir, p, os.path.join(workdir, f"adv/ref_{tag}"), T=2, extra=ex)
        op = ours_index(p, os.path.join(workdir, f"adv/our_{tag}"), T=2, extra=ex)
        compare_index(rp, op, 2)
        ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, f"adv/ref_{tag}"), T=1)
        oo = ours_overlap(p, op, 2, os.path.join(workdir, f"adv/our_{tag}"), T=1)
        assert_same_ovlp(oo[0], ro[0])


def test_noisy_reads_one_percent(workdir, ref_dir):
    """The E. coli test's error rate (1 %, test/ecoli_K12/simulate_reads.py:13): wider live band, more rejected alignments."""
    p = D.make_sim(workdir, "sim_e1", genome=400_000, cov=25, err=0.01, seed=7)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "sim_e1/ref"), T=1, extra=["-m", "0"])
    op = ours_index(p, os.path.join(workdir, "sim_e1/our"), T=1, extra=["-m", "0"])
    compare_index(rp, op, 1, levels=("L2",))
    for ex in ([], ["-w", "50"]):
        ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "sim_e1/ref" + "".join(ex)), T=1, extra=ex)
        oo = ours_overlap(p, op, 2, os.path.join(workdir, "sim_e1/our" + "".join(ex)), T=1, extra=ex)
        assert_same_ovlp(oo[0], ro[0])


def test_engine_api_matches_cli(sim1, workdir, ref_dir):
    """Stage-level C ABI (device hand-off of L2, no files) gives the same records as the file-based tools."""
    from peregrine_b200 import Engine

    rid, ln, off = F.read_idx(sim1 + ".idx")
    seqdb = np.fromfile(sim1 + ".seqdb", dtype=np.uint8)
    eng = Engine(0)
    eng.load_reads(seqdb, rid, ln, off)
    eng.index(80, 16, 6, 2, with_counts=0b100)
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    assert np.array_equal(eng.level(2), F.read_mmlist(rp + "-L2-01-of-01.dat"))
    assert np.array_equal(F.mc_as_sorted_pairs(eng.level_counts(2)), F.mc_as_sorted_pairs(F.read_mc(rp + "-L2-MC-01-of-01.dat")))
    eng.set_shimmers_from_index(2)
    ov = eng.overlap(1, 1)
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref1"), T=1)
    ref = F.normalise_ovlp(F.read_ovlp(ro[0]))
    assert len(ov) == len(ref) and ov.tobytes() == ref.tobytes()
    st = eng.stats()
    assert st["kernel_launches"] > 10 and st["n_alignments"] >= len(ref)
    eng.close()


def test_deferred_load_overlaps_copy_with_sketch(sim1, workdir, ref_dir, monkeypatch):
    """PGB_LOAD_DEFER: index() copies, packs and sketches chunk by chunk (tiny chunks here, so that several are in flight);
    a consumer other than index() completes the copy first.  Same L0/L2 and same records as the plain load."""
    from peregrine_b200 import Engine

    monkeypatch.setenv("PGB_LOAD_CHUNK_MB", "1")
    rid, ln, off = F.read_idx(sim1 + ".idx")
    seqdb = np.fromfile(sim1 + ".seqdb", dtype=np.uint8)
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref1"), T=1)
    ref = F.normalise_ovlp(F.read_ovlp(ro[0]))
    eng = Engine(0)
    eng.load_reads(seqdb, rid, ln, off, defer=True)
    eng.index(80, 16, 6, 2)
    assert np.array_equal(eng.level(0), F.read_mmlist(rp + "-L0-01-of-01.dat"))
    assert np.array_equal(eng.level(2), F.read_mmlist(rp + "-L2-01-of-01.dat"))
    eng.set_shimmers_from_index(2)
    ov = eng.overlap(1, 1, copy="view")
    assert len(ov) == len(ref) and ov.tobytes() == ref.tobytes()
    # deferred load consumed by overlap() directly (shimmers from the reference's files)
    eng.load_reads(seqdb, rid, ln, off, defer=True)
    eng.set_shimmers(F.read_mmlist(rp + "-L2-01-of-01.dat"), F.read_mc(rp + "-L2-MC-01-of-01.dat"))
    ov = eng.overlap(1, 1)
    assert len(ov) == len(ref) and ov.tobytes() == ref.tobytes()
    eng.close()


def test_empty_and_tiny_inputs(workdir, ref_dir):
    """Chunks with no reads, reads with no minimizer, and an index with no eligible bucket."""
    recs = [("t/000000/0_5", "ACGTA"), ("t/000001/0_40", "ACGTTGCAAGGCTTAACCGGTTAACCGGATATCGCGATAT" )]
    p = D.make_from_fasta(workdir, "tiny", recs, ref_dir)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "tiny/ref"), T=3, extra=["-m", "1"])
    op = ours_index(p, os.path.join(workdir, "tiny/our"), T=3, extra=["-m", "1"])
    compare_index(rp, op, 3)
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "tiny/ref"), T=1)
    oo = ours_overlap(p, op, 2, os.path.join(workdir, "tiny/our"), T=1)
    assert_same_ovlp(oo[0], ro[0])


def test_exact_automaton_path_on_all_reads(sim1, workdir, ref_dir, monkeypatch):
    """PGB_SKETCH=exact forces every read through k_sketch_exact (normally only the reads the tiled kernel flags)."""
    monkeypatch.setenv("PGB_SKETCH", "exact")
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    op = ours_index(sim1, os.path.join(workdir, "sim1/our_exact"), T=1, extra=["-m", "1"])
    compare_index(rp, op, 1)


def test_tiled_sketch_reports_fallbacks(workdir, ref_dir):
    """The tiled kernel hands flagged reads (N, ties, palindromes, short) to the exact automaton and says how many."""
    from peregrine_b200 import Engine

    p = D.make_from_fasta(workdir, "adv", D.adversarial_records(), ref_dir)
    rid, ln, off = F.read_idx(p + ".idx")
    seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
    eng = Engine(0)
    eng.load_reads(seqdb, rid, ln, off)
    eng.index(80, 16, 6, 2)
    st = eng.stats()
    assert st["n_k_sketch_tiled"] == 1 and 0 < st["n_sketch_fallback_reads"] < len(rid)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "adv/ref_T1"), T=1, extra=["-m", "1"])
    assert np.array_equal(eng.level(0), F.read_mmlist(rp + "-L0-01-of-01.dat"))
    eng.close()


@pytest.mark.parametrize("big,warp_min,group", [("3", "64", "8"), ("1000000", "64", "8"), ("64", "2", "8"), ("64", "2", "32"), ("64", "3", "4"), ("48", "16", "16")])
def test_replay_kernel_forms(sim1, workdir, ref_dir, monkeypatch, big, warp_min, group):
    """The three forms of the bucket scan: PGB_REPLAY_BIG=3 sends every bucket through the block-cooperative kernel
    (k_replay_block), a huge value with PGB_REPLAY_WARP_MIN=64 through the one-thread-per-bucket kernel (k_replay), and
    PGB_REPLAY_BIG=64 with PGB_REPLAY_WARP_MIN=2 through the lane-group kernel (k_replay_group<G>, G = PGB_REPLAY_GROUP lanes
    per bucket; buckets in which a read occurs twice fall back to one lane).  The default mixes them by bucket size.  Every
    mix must reproduce the reference stream, on clean reads, on 1 % error reads, with a small bestn (the row cut-off inside
    a group's G candidates) and on the adversarial set (tandem repeats: reads with several records per bucket)."""
    monkeypatch.setenv("PGB_REPLAY_BIG", big)
    monkeypatch.setenv("PGB_REPLAY_BIG_TAIL", big)
    monkeypatch.setenv("PGB_REPLAY_WARP_MIN", warp_min)
    monkeypatch.setenv("PGB_REPLAY_GROUP", group)
    tag = f"rb{big}_{warp_min}_{group}"
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    for extra in ([], ["-b", "2", "-n", "40"], ["-b", "1"]):
        et = "".join(extra).replace("-", "")
        ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, f"sim1/refo_{et}"), T=1, extra=extra)
        oo = ours_overlap(sim1, rp, 2, os.path.join(workdir, f"sim1/our_{tag}{et}"), T=1, extra=extra)
        assert_same_ovlp(oo[0], ro[0])
    p = D.make_sim(workdir, "sim_e1", genome=400_000, cov=25, err=0.01, seed=7)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "sim_e1/ref"), T=1, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "sim_e1/ref"), T=1)
    oo = ours_overlap(p, rp, 2, os.path.join(workdir, f"sim_e1/our_{tag}"), T=1)
    assert_same_ovlp(oo[0], ro[0])
    p = D.make_from_fasta(workdir, "adv", D.adversarial_records(), ref_dir)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "adv/ref_m1"), T=1, extra=["-m", "1"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "adv/ref_m1"), T=1, extra=["-M", "1000"])
    oo = ours_overlap(p, rp, 2, os.path.join(workdir, f"adv/our_{tag}"), T=1, extra=["-M", "1000"])
    assert_same_ovlp(oo[0], ro[0])


def test_warp_per_alignment_kernel_on_every_alignment(sim1, workdir, ref_dir, monkeypatch):
    """PGB_ALIGN_WARP_MAX=huge routes every alignment batch (not only the small tail batches) through k_align_warp."""
    monkeypatch.setenv("PGB_ALIGN_WARP_MAX", "4000000000")
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref1"), T=1, extra=["-m", "1"])
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref1"), T=1)
    oo = ours_overlap(sim1, rp, 2, os.path.join(workdir, "sim1/our_aw"), T=1)
    assert_same_ovlp(oo[0], ro[0])


@pytest.mark.parametrize("variant", ["7", "0", "20", "21", "22", "23", "30", "31"])
def test_thread_per_alignment_kernel_on_every_alignment(workdir, ref_dir, monkeypatch, variant):
    """PGB_ALIGN_WARP_MAX=0 routes every alignment batch through the bulk kernel: k_align_lean (variant 7: band-row prefetch +
    register-cached band trim; variant 0: the plain form) or k_align_quad (20: 4 lanes per alignment, 21: 8 lanes; operand
    windows staged by cp.async.bulk), on clean and on 3 %-error reads (wide bands, deep trims, multi-chunk rows) and with a
    narrow band limit (-w 30: the band-limit exit of DWmatch.c:120-122)."""
    monkeypatch.setenv("PGB_ALIGN_WARP_MAX", "0")
    monkeypatch.setenv("PGB_ALIGN_VARIANT", variant)
    for name, kw in (("lean_a", dict(genome=400_000, cov=20)), ("lean_b", dict(genome=150_000, cov=15, err=0.03))):
        p = D.make_sim(workdir, name, **kw)
        rp = D.ref_index(ref_dir, p, os.path.join(workdir, name, "ref"), T=1, extra=["-m", "0"])
        for extra in ([], ["-w", "30"]):
            ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, name, "ref" + "".join(extra)), T=1, extra=extra)
            oo = ours_overlap(p, rp, 2, os.path.join(workdir, name, f"our{variant}" + "".join(extra)), T=1, extra=extra)
            assert_same_ovlp(oo[0], ro[0])


def test_sharded_exchange_matches_reference(sim1, workdir, ref_dir):
    """Multi-GPU data path on one device: three 'ranks' (engines) each sketch the reads of their index chunk, the packed
    reads and SHIMMER lists are concatenated exactly as the NCCL all-gather of peregrine_b200.multigpu does, and every rank
    then produces its hash chunk; records must equal the reference's shmr_overlap -t 3 -c {1,2,3} on 3 index chunk files."""
    import torch
    from peregrine_b200 import Engine, multigpu as M

    T = 3
    dev = torch.device("cuda", 0)
    rid, ln, off = F.read_idx(sim1 + ".idx")
    seqdb = np.fromfile(sim1 + ".seqdb", dtype=np.uint8)
    engs = [Engine(0) for _ in range(T)]
    parts, l2s = [], []
    for r in range(T):
        engs[r].load_reads(seqdb, rid, ln, off, T, r + 1)
        engs[r].index(80, 16, 6, 2)
        parts.append(M.export_reads(engs[r], dev))
        l2s.append(M.export_level(engs[r], 2, dev))
    reads = M.concat_reads(parts)
    l2_all = torch.cat(l2s)
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref3"), T=T, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref3o"), T=T)
    for r in range(T):
        M.import_reads(engs[r], reads)
        M.import_shimmers(engs[r], l2_all)
        ov = engs[r].overlap(T, r + 1)
        want = F.normalise_ovlp(F.read_ovlp(ro[r]))
        assert len(ov) == len(want) and ov.tobytes() == want.tobytes(), f"chunk {r + 1}"
    for e in engs:
        e.close()


def test_routed_exchange_matches_reference(sim1, workdir, ref_dir):
    """north_star's exchange on one device: three 'ranks' each scan only their own reads' shimmers with the summed
    multiplicity tables, emit the pair records of every hash chunk grouped by owner (pgb_route_build), the groups are moved
    as the all-to-all of peregrine_b200.multigpu does (source ranks concatenated in rank order), and every owner runs
    pgb_overlap_routed; records must equal the reference's shmr_overlap -t 3 -c {1,2,3} over 3 index chunk files."""
    import torch
    from peregrine_b200 import Engine, multigpu as M

    T = 3
    dev = torch.device("cuda", 0)
    rid, ln, off = F.read_idx(sim1 + ".idx")
    seqdb = np.fromfile(sim1 + ".seqdb", dtype=np.uint8)
    engs = [Engine(0) for _ in range(T)]
    parts, counts = [], []
    for r in range(T):
        engs[r].load_reads(seqdb, rid, ln, off, T, r + 1)
        engs[r].index(80, 16, 6, 2)
        engs[r].set_shimmers_from_index(2)
        parts.append(M.export_reads(engs[r], dev))
        counts.append(M.export_counts(engs[r], dev))
    all_counts = torch.cat(counts).contiguous()
    torch.cuda.synchronize()  # torch's stream -> the library's stream (the engine is handed raw pointers below)
    has_first, sends, splits = [], [], []
    for r in range(T):
        engs[r].counts_set_device(all_counts.data_ptr(), int(all_counts.shape[0]))
        has_first.append(engs[r].route_scan(2, 240))
    for r in range(T):
        per_chunk = engs[r].route_build(T, 2, 240, any(has_first[:r]))
        send = M.export_route(engs[r], dev)
        assert sum(per_chunk) == send.shape[0]
        sends.append(send)
        splits.append(np.concatenate([[0], np.cumsum(per_chunk)]))
    reads = M.concat_reads(parts)
    rp = D.ref_index(ref_dir, sim1, os.path.join(workdir, "sim1/ref3"), T=T, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, sim1, rp, 2, os.path.join(workdir, "sim1/ref3o"), T=T)
    ovl = Engine(0)
    M.import_reads(ovl, reads)
    for d in range(T):  # owner of chunk d + 1
        recv = torch.cat([sends[src][int(splits[src][d]): int(splits[src][d + 1])] for src in range(T)]).contiguous()
        torch.cuda.synchronize()  # (without it the library may read recv before torch.cat has written it: seen once as missing records)
        ov = ovl.overlap_routed(recv.data_ptr(), int(recv.shape[0]), total_chunk=T)
        want = F.normalise_ovlp(F.read_ovlp(ro[d]))
        assert len(ov) == len(want) and ov.tobytes() == want.tobytes(), f"chunk {d + 1}"
    ovl.close()
    for e in engs:
        e.close()
