"""CPU tests (no GPU): the plain-C oracle restatement against (a) the committed golden vectors generated from the
unmodified reference and (b) the reference itself (oracle/_ref) on fresh seeded inputs; plus host-side plumbing."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import datasets as D
import goldenset as G
import oracle as O
from peregrine_b200 import formats as F

ROOT = D.ROOT


@pytest.fixture(scope="module")
def gold():
    return G.load()


def test_golden_seqdb_image(gold):
    g, reads = gold
    seqdb, rid, ln, off = G.seqdb_arrays(reads)
    assert G.sha(seqdb.tobytes()) == g["seqdb_sha256"]


def test_oracle_sketch_golden(gold):
    g, reads = gold
    for case in g["mm_sketch"]:
        out = [O.orc_sketch(O.ascii_to_nib(s), case["w"], case["k"], i) for i, s in enumerate(reads)]
        a = np.concatenate(out)
        assert len(a) == case["n"] and G.sha(a.tobytes()) == case["sha256"], case


def test_oracle_reduce_golden(gold):
    g, reads = gold
    l0 = np.concatenate([O.orc_sketch(O.ascii_to_nib(s), 80, 16, i) for i, s in enumerate(reads)])
    for case in g["mm_reduce"]:
        a = O.orc_reduce(l0, case["r"])
        b = O.orc_reduce(a, case["r"])
        assert (len(a), len(b)) == (case["n1"], case["n2"])
        assert G.sha(a.tobytes()) == case["sha256_1"] and G.sha(b.tobytes()) == case["sha256_2"]


def test_oracle_ovlp_match_golden(gold):
    g, reads = gold
    enc = [O.encode_biseq(s) for s in reads]
    for case in g["ovlp_match"]:
        got = O.orc_ovlp_match(enc[case["i"]][case["start"]:], case["s0"], enc[case["j"]], case["s1"], case["bw"])
        assert list(map(int, got)) == case["match"], case


def test_oracle_index_and_overlap_golden(gold):
    g, reads = gold
    seqdb, rid, ln, off = G.seqdb_arrays(reads)
    p = g["params"]
    l2_all, mc_all = [], []
    for c in (1, 2):
        lv = O.orc_index_chunk(seqdb, rid, ln, off, p["T_idx"], c, p["w"], p["k"], p["r"], 2)
        for l in (0, 2):
            hdr = np.uint64(len(lv[l])).tobytes()
            assert G.sha(hdr + lv[l].tobytes()) == g["files"][f"shmr-L{l}-{c:02d}-of-02.dat"]
            assert G.sha(F.mc_as_sorted_pairs(O.orc_count(lv[l])).tobytes()) == g["files"][f"shmr-L{l}-MC-{c:02d}-of-02.dat"]
        l2_all.append(lv[2])
        mc_all.append(O.orc_count(lv[2]))
    mm, mc = np.concatenate(l2_all), np.concatenate(mc_all)
    ov, _ = O.orc_overlap_chunk(seqdb, rid, ln, off, mm, mc, T=1, c=1)
    want = np.fromfile(os.path.join(G.HERE, "ovlp_T1.bin"), dtype=F.OVLP)
    assert len(ov) == g["ovlp_T1"]["records"] and ov.tobytes() == want.tobytes()
    for c in (1, 2):
        ov, _ = O.orc_overlap_chunk(seqdb, rid, ln, off, mm, mc, T=2, c=c)
        assert len(ov) == g[f"ovlp_T2_c{c}"]["records"] and G.sha(ov.tobytes()) == g[f"ovlp_T2_c{c}"]["sha256"]


# ------------------------------------------------------------------------------------------------ against the live reference
def _need_ref():
    L = O.reflib()
    if L is None:
        pytest.skip("oracle/_ref not present")
    return L


def test_oracle_vs_reference_calls():
    L = _need_ref()
    rnd = np.random.default_rng(5)
    B = np.array(list("ACGT"))
    for trial in range(60):
        n = int(rnd.integers(1, 3000))
        s = "".join(rnd.choice(B, n))
        if trial % 5 == 0:
            unit = "".join(rnd.choice(B, int(rnd.integers(1, 40))))
            s = (s[: n // 3] + unit * 30 + s[n // 3:])[: max(n, 50)]
        if trial % 7 == 0 and len(s) > 60:
            s = s[:30] + "N" * int(rnd.integers(1, 20)) + s[40:]
        w, k = [(80, 16), (24, 12), (120, 18), (60, 28), (255, 13)][trial % 5]
        a = O.abi_sketch(L, s, w, k, trial)
        b = O.orc_sketch(O.ascii_to_nib(s), w, k, trial)
        assert np.array_equal(a, b), (trial, w, k, len(a), len(b))
        for r in (2, 6):
            assert np.array_equal(O.abi_reduce(L, a, r), O.orc_reduce(b, r))
    # alignments: noisy copies (true overlaps), unrelated pairs, every strand combination
    for trial in range(60):
        n = int(rnd.integers(600, 4000))
        s = "".join(rnd.choice(B, n))
        t = list(s[int(rnd.integers(0, 200)):])
        for _ in range(int(len(t) * 0.01 * (trial % 4))):
            p = int(rnd.integers(0, len(t)))
            t[p] = str(rnd.choice(B)) if rnd.random() < 0.5 else ""
        t = "".join(t) if trial % 6 else "".join(rnd.choice(B, n))
        q, tt = O.encode_biseq(s), O.encode_biseq(t)
        for qs in (0, 1):
            for ts in (0, 1):
                bw = (20, 50, 100, 200)[trial % 4]
                assert np.array_equal(O.abi_ovlp_match(L, q, qs, tt, ts, bw), O.orc_ovlp_match(q, qs, tt, ts, bw))


def test_oracle_vs_reference_tools(workdir, ref_dir):
    """Chunk drivers of the oracle against the reference binaries on a simulated set (multi-chunk, 1 % error)."""
    p = D.make_sim(workdir, "orc_sim", genome=300_000, cov=20, err=0.01, seed=3)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "orc_sim/ref"), T=3, extra=["-m", "1"])
    rid, ln, off = F.read_idx(p + ".idx")
    seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
    l2, mc = [], []
    for c in (1, 2, 3):
        lv = O.orc_index_chunk(seqdb, rid, ln, off, 3, c, 80, 16, 6, 2)
        assert np.array_equal(lv[0], F.read_mmlist(f"{rp}-L0-{c:02d}-of-03.dat"))
        assert np.array_equal(lv[2], F.read_mmlist(f"{rp}-L2-{c:02d}-of-03.dat"))
        assert np.array_equal(F.mc_as_sorted_pairs(O.orc_count(lv[2])), F.mc_as_sorted_pairs(F.read_mc(f"{rp}-L2-MC-{c:02d}-of-03.dat")))
        l2.append(lv[2]); mc.append(O.orc_count(lv[2]))
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "orc_sim/ref"), T=2)
    for c in (1, 2):
        ov, _ = O.orc_overlap_chunk(seqdb, rid, ln, off, np.concatenate(l2), np.concatenate(mc), T=2, c=c)
        want = F.normalise_ovlp(F.read_ovlp(ro[c - 1]))
        assert len(ov) == len(want) and ov.tobytes() == want.tobytes()


# ------------------------------------------------------------------------------------------------ host plumbing
def test_library_exports_every_declared_symbol():
    """libpgb200.so loads without a GPU and exports every function include/pgb200.h declares."""
    from peregrine_b200 import lib_path

    hdr = open(os.path.join(ROOT, "include", "pgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([a-z_0-9]+)\s*\([^;{}]*\)\s*;", hdr))
    assert {"pgb_overlap", "pgb_index", "mm_sketch", "ovlp_match", "pgb_shmr_overlap_main"} <= names
    L = C.CDLL(lib_path())
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_no_gpu_means_loud_failure():
    """Without a CUDA device the product refuses to run instead of falling back to the CPU."""
    from peregrine_b200 import Engine, NoDeviceError, load_library

    L = load_library()
    if L.pgb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(NoDeviceError):
        Engine(0)
    r = subprocess.run([os.path.join(ROOT, "bin", "shmr_index"), "-p", "/nonexistent/x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode != 0 and b"no CUDA device" in r.stderr


def test_formats_roundtrip(tmp_path):
    a = np.zeros(5, dtype=F.MM128)
    a["x"] = np.arange(5) << 8 | 16
    a["y"] = np.arange(5) << 32 | 7
    F.write_mmlist(str(tmp_path / "a.dat"), a)
    assert np.array_equal(F.read_mmlist(str(tmp_path / "a.dat")), a)
    m = np.zeros(3, dtype=F.MMCOUNT)
    m["mer"] = [5, 3, 9]
    m["count"] = [1, 2, 3]
    F.write_mc(str(tmp_path / "m.dat"), m)
    assert F.mc_as_sorted_pairs(F.read_mc(str(tmp_path / "m.dat"))).tolist() == [[3, 2], [5, 1], [9, 3]]


def test_kernel_logic_on_host(workdir, ref_dir):
    """tests/hostsim compiles the product's __host__ __device__ per-item functions for the CPU and compares them with the
    reference: sketch automaton + reduce, packed-word ovlp_match, and the fix-point replay of the greedy bucket scan."""
    subprocess.check_call(["make", "-C", ROOT, "hostsim"], stdout=subprocess.DEVNULL)
    hs = os.path.join(ROOT, "build", "hostsim")
    env = dict(os.environ, PGB_REF_LIB=os.path.join(ref_dir, "libshimmer_ref.so"))
    p = D.make_from_fasta(workdir, "adv_h", D.adversarial_records(seed=11), ref_dir)
    for w, k, r in ((80, 16, 6), (24, 12, 2), (120, 18, 3)):
        subprocess.check_call([hs, "sketch", p, str(w), str(k), str(r)], env=env, stdout=subprocess.DEVNULL)
    # the strip sketch kernel's per-strip fallback: a host model of the kernel's window decisions + the exact automaton on the bad
    # strips, spliced by position, must give the reference's minimizers (planted ties and palindromic k-mers; even and odd k)
    pb = D.make_from_fasta(workdir, "badstrips_h", D.bad_strip_records(seed=43, n_reads=150), ref_dir)
    for w, k in ((80, 16), (40, 14), (120, 18), (33, 15), (64, 12)):
        out = subprocess.run([hs, "sketch", pb, str(w), str(k), "6"], env=env, stdout=subprocess.PIPE, check=True).stdout.decode()
        line = [x for x in out.splitlines() if x.startswith("strip model")][0]
        f = dict(kv.split("=") for kv in line.replace("of them in part", "partial").replace("automaton pieces", "pieces").split(":")[1].split())
        assert int(f["partial"]) > 10 and int(f["pieces"]) > 10 and int(f["mismatching"]) == 0, line
    # ... and the check can fail: without the positions before a bad strip the splice is wrong
    r = subprocess.run([hs, "sketch", pb, "80", "16", "6"], env=dict(env, PGB_SPLICE_REACH="0"), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
    assert r.returncode == 3 and b"mismatching=0\n" not in [x for x in r.stdout.splitlines(True) if x.startswith(b"strip model")][0]
    # the lane-group bucket walk of k_replay_group: its cut-off logic against the sequential replay_bucket on random buckets
    subprocess.check_call([hs, "grouprow", "20000"], stdout=subprocess.DEVNULL)
    subprocess.check_call([hs, "match", p, "1500", "50"], env=env, stdout=subprocess.DEVNULL)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "adv_h/ref"), T=2, extra=["-m", "0", "-k", "18", "-w", "120", "-r", "3"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "adv_h/ref"), T=1)
    subprocess.check_call([hs, "overlap", p, rp + "-L2", "1", "1", ro[0]], env=env, stdout=subprocess.DEVNULL)
    env["PGB_SIM_STRICT"] = "1"
    subprocess.check_call([hs, "overlap", p, rp + "-L2", "1", "1", ro[0], "jacobi", "4"], env=env, stdout=subprocess.DEVNULL)
