"""GPU parity tests added after the round's last full GPU run (they sort last so that a surprise in one of them cannot hide
the established tests behind `pytest -x`): the full config-5 parameter grid and the overflow / restart path of the replay
tables.  Same helpers and checker (the unmodified reference in oracle/_ref) as tests/test_gpu_parity.py."""
import os

import pytest

import datasets as D
from test_gpu_parity import assert_same_ovlp, compare_index, ours_index, ours_overlap

pytestmark = pytest.mark.gpu


def test_config5_grid(workdir, ref_dir):
    """BASELINE.json configs[4]: every (k, w) of {14,16,18} x {60,80,120} through index AND overlap with aln_bw of 50/100/200,
    on a small noisy set (1 % error: wider bands than the clean sets)."""
    p = D.make_sim(workdir, "c5", genome=250_000, cov=15, err=0.01)
    bws = ("50", "100", "200")
    n = 0
    for k in (14, 16, 18):
        for w in (60, 80, 120):
            tag = f"k{k}w{w}"
            ex = ["-k", str(k), "-w", str(w), "-m", "0"]
            rp = D.ref_index(ref_dir, p, os.path.join(workdir, f"c5/ref_{tag}"), T=1, extra=ex)
            op = ours_index(p, os.path.join(workdir, f"c5/our_{tag}"), T=1, extra=ex)
            compare_index(rp, op, 1, levels=("L2",))
            bw = bws[n % 3]
            n += 1
            ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, f"c5/refo_{tag}"), T=1, extra=["-w", bw])
            oo = ours_overlap(p, op, 2, os.path.join(workdir, f"c5/ouro_{tag}"), T=1, extra=["-w", bw])
            assert_same_ovlp(oo[0], ro[0])


@pytest.mark.parametrize("scale", ["0.02", "0.3"])
def test_replay_tables_overflow_and_restart(workdir, ref_dir, monkeypatch, scale):
    """PGB_TABLE_SCALE shrinks the initial pair table / alignment cache so that they fill up: probes are bounded
    (PGB_MAX_PROBE), the pass is abandoned, the host doubles the table and restarts the fix-point, and the records are
    still the reference's.  (At T >= 4 chunks of a big job the default sizing overflows for real: rid_pairs is per chunk,
    so a read pair is aligned in several chunks.)"""
    monkeypatch.setenv("PGB_TABLE_SCALE", scale)
    monkeypatch.setenv("PGB_VERBOSE", "1")
    p = D.make_sim(workdir, "ovf", genome=300_000, cov=20)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "ovf/ref"), T=1, extra=["-m", "0"])
    for T in (1, 4):
        ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, f"ovf/ref{T}"), T=T)
        oo = ours_overlap(p, rp, 2, os.path.join(workdir, f"ovf/our{T}_{scale}"), T=T)
        for a, b in zip(oo, ro):
            assert_same_ovlp(a, b)


def test_cli_chain_like_run_test_sh(workdir, ref_dir):
    """tools/run_chain.sh = the reference's own shell recipe (test/ecoli_K12/run_test.sh:16-28): shmr_mkseqdb -> shmr_index x 3
    -> shmr_overlap x 2 (two processes sharing the GPU) -> cat | shmr_dedup, once with bin/ and once with the unmodified
    reference; preads.ovl and every intermediate file must be identical."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = D.make_sim(workdir, "chain", genome=500_000, cov=15)
    fa = os.path.join(workdir, "chain", "reads.fa")
    if not os.path.exists(fa):  # FASTA of the simulated set (decode the low nibbles of the .seqdb image)
        import numpy as np
        from peregrine_b200 import formats as F

        rid, ln, off = F.read_idx(p + ".idx")
        db = np.fromfile(p + ".seqdb", dtype=np.uint8)
        lut = np.zeros(16, dtype=np.uint8)
        lut[[1, 2, 4, 8]] = np.frombuffer(b"ACGT", dtype=np.uint8)
        lut[0] = ord("N")
        with open(fa, "wb") as f:
            for i in range(len(rid)):
                f.write(b">r/%06d/0_%d\n" % (i, ln[i]) + lut[db[int(off[i]): int(off[i]) + int(ln[i])] & 0x0F].tobytes() + b"\n")
    lst = os.path.join(workdir, "chain", "in.lst")
    with open(lst, "w") as f:
        f.write(fa + "\n")
    outs = {}
    for tag, bindir in (("ref", ref_dir), ("our", os.path.join(root, "bin"))):
        wd = os.path.join(workdir, "chain", tag)
        subprocess.run(["bash", os.path.join(root, "tools", "run_chain.sh"), lst, wd, "3", "2", "2"], env=dict(os.environ, BIN=bindir), check=True,
                       stdout=subprocess.DEVNULL)
        outs[tag] = wd
    for rel in ["asm/preads.ovl", "index/seq_dataset.idx", "index/seq_dataset.seqdb"] + [f"index/shmr-L2-{c:02d}-of-03.dat" for c in (1, 2, 3)]:
        a, b = (open(os.path.join(outs[t], rel), "rb").read() for t in ("our", "ref"))
        assert a == b and len(a) > 0, rel
    for c in ("01", "02"):
        assert_same_ovlp(os.path.join(outs["our"], "ovlp", "ovlp." + c), os.path.join(outs["ref"], "ovlp", "ovlp." + c))
    assert open(os.path.join(outs["our"], "asm/preads.ovl"), "rb").read().count(b"\n") > 500


def test_two_bit_handoff(workdir, ref_dir):
    """pgb_pack_2bit + pgb_load_reads_2bit (the .seq2b / .seq2n hand-off of this library's shmr_mkseqdb) against the 1-byte/base
    load: same packed image => same L2 and the same overlap records, on reads with N runs, for a chunked selection (gathered
    copy), the whole set (one copy) and the deferred form that overlaps the copy with sketching."""
    import numpy as np
    from peregrine_b200 import Engine, formats as F

    p = D.make_from_fasta(workdir, "adv2b", D.adversarial_records(seed=11), ref_dir)
    rid, ln, off = F.read_idx(p + ".idx")
    seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
    a, b = Engine(0), Engine(0)
    words, nrec = a.pack_2bit(seqdb, off, ln)
    assert nrec.size > 0  # the adversarial set has reads with N
    for T, chunk, defer in ((1, 1, False), (1, 1, True), (3, 2, False), (3, 3, False)):
        a.load_reads(seqdb, rid, ln, off, T, chunk)
        b.load_reads_2bit(words, nrec, rid, ln, T, chunk, defer=defer)
        a.index(80, 16, 6, 2)
        b.index(80, 16, 6, 2)
        for lv in (0, 2):
            assert np.array_equal(a.level(lv), b.level(lv)), (T, chunk, defer, lv)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "adv2b/ref"), T=1, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "adv2b/ref"), T=1)
    b.load_reads_2bit(words, nrec, rid, ln)
    b.index(80, 16, 6, 2)
    b.set_shimmers_from_index(2)
    ov = b.overlap(1, 1)
    want = F.normalise_ovlp(F.read_ovlp(ro[0]))
    assert len(ov) == len(want) and ov.tobytes() == want.tobytes()
    a.close()
    b.close()


def test_sketch_bad_strips_are_spliced(workdir, ref_dir, monkeypatch):
    """The strip sketch kernel hands single 512-position strips (not whole reads) to the exact automaton and splices the two
    record streams by position.  Reads with planted trouble at chosen places: tandem duplications (equal k-mer hashes inside one
    window: ties) in the middle of a read, in two adjacent strips, across a strip boundary, in the last partial strip and beyond
    strip 63 (> 32 kb: whole-read fallback); palindromic k-mers (no window slot) before the first full window, two within one
    window, and one next to a duplication.  L0 must be byte-identical to the reference for three (k, w) and the counters must show
    that reads were redone in part."""
    import numpy as np
    from peregrine_b200 import Engine, formats as F

    recs = D.bad_strip_records()
    p = D.make_from_fasta(workdir, "badstrips", recs, ref_dir)
    rid, ln, off = F.read_idx(p + ".idx")
    seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
    eng = Engine(0)
    eng.load_reads(seqdb, rid, ln, off)
    for k, w in ((16, 80), (14, 40), (18, 120)):
        rp = D.ref_index(ref_dir, p, os.path.join(workdir, f"badstrips/ref_k{k}w{w}"), T=1, extra=["-m", "1", "-k", str(k), "-w", str(w)])
        eng.stats_reset()
        eng.index(w, k, 6, 2)
        st = eng.stats()
        assert 0 < st["n_sketch_fallback_reads"] < len(rid), (k, w, st["n_sketch_fallback_reads"])
        for lv in (0, 2):  # (-m 1 writes L0 next to the final level)
            assert np.array_equal(eng.level(lv), F.read_mmlist(rp + f"-L{lv}-01-of-01.dat")), (k, w, lv)
    eng.close()
