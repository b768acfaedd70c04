"""GPU tests of the reference cffi surface exported by libpgb200.so (include/pgb200.h group 1): every function is called
through the reference's ABI and compared call-for-call with the unmodified reference (oracle/_ref/libshimmer_ref.so) and
with the committed golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

import goldenset as G
import oracle as O
from peregrine_b200 import formats as F, lib_path

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ours():
    L = C.CDLL(lib_path())
    L.mm_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(O.MMV)]
    L.mm_reduce.argtypes = [C.POINTER(O.MMV), C.POINTER(O.MMV), C.c_uint8]
    L.ovlp_match.restype = C.POINTER(O.MatchT)
    L.ovlp_match.argtypes = [C.c_void_p, C.c_int32, C.c_uint8, C.c_void_p, C.c_int32, C.c_uint8, C.c_int32]
    L.free_ovlp_match.argtypes = [C.POINTER(O.MatchT)]
    return L


def test_golden_vectors_through_the_abi(ours):
    g, reads = G.load()
    for case in g["mm_sketch"][:3]:
        a = np.concatenate([O.abi_sketch(ours, s, case["w"], case["k"], i) for i, s in enumerate(reads)])
        assert len(a) == case["n"] and G.sha(a.tobytes()) == case["sha256"], case
    l0 = np.concatenate([O.abi_sketch(ours, s, 80, 16, i) for i, s in enumerate(reads)])
    for case in g["mm_reduce"]:
        a = O.abi_reduce(ours, l0, case["r"])
        b = O.abi_reduce(ours, a, case["r"])
        assert G.sha(a.tobytes()) == case["sha256_1"] and G.sha(b.tobytes()) == case["sha256_2"], case
    enc = [O.encode_biseq(s) for s in reads]
    for case in g["ovlp_match"]:
        got = O.abi_ovlp_match(ours, enc[case["i"]][case["start"]:], case["s0"], enc[case["j"]], case["s1"], case["bw"])
        assert list(map(int, got)) == case["match"], case


def test_cffi_surface_sketch_reduce_match_vs_reference(ours):
    ref = O.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    rnd = np.random.default_rng(11)
    B = np.array(list("ACGT"))
    for trial in range(25):
        n = int(rnd.integers(1, 6000))
        s = "".join(rnd.choice(B, n))
        if trial % 4 == 0 and n > 100:
            s = s[:50] + "N" * 7 + s[57:]
        if trial % 5 == 0:
            s = (s[: n // 2] + "ACG" * 60 + s[n // 2:])
        w, k = [(80, 16), (24, 12), (120, 18), (10, 8), (255, 28)][trial % 5]
        a, b = O.abi_sketch(ref, s, w, k, trial), O.abi_sketch(ours, s, w, k, trial)
        assert np.array_equal(a, b), (trial, w, k, len(a), len(b))
        for r in (1, 3, 6):
            assert np.array_equal(O.abi_reduce(ref, a, r), O.abi_reduce(ours, b, r))
    for trial in range(25):
        n = int(rnd.integers(600, 5000))
        s = "".join(rnd.choice(B, n))
        t = list(s[int(rnd.integers(0, 300)):])
        for _ in range(int(len(t) * 0.01 * (trial % 3))):
            p = int(rnd.integers(0, len(t)))
            t[p] = str(rnd.choice(B)) if rnd.random() < 0.5 else ""
        t = "".join(t)
        q, tt = O.encode_biseq(s), O.encode_biseq(t)
        for qs in (0, 1):
            for ts in (0, 1):
                bw = (20, 100, 200)[trial % 3]
                assert np.array_equal(O.abi_ovlp_match(ref, q, qs, tt, ts, bw), O.abi_ovlp_match(ours, q, qs, tt, ts, bw))


class IdxV(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(C.c_uint32))]


class AlnT(C.Structure):
    _fields_ = [("idx0", IdxV), ("idx1", IdxV)]


class AlnV(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(AlnT))]


def _aln(L, m0, m1, direction, max_diff, max_dist, max_repeat):
    L.shmr_aln.restype = C.POINTER(AlnV)
    L.shmr_aln.argtypes = [C.POINTER(O.MMV), C.POINTER(O.MMV), C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint32]
    L.free_shmr_alns.argtypes = [C.POINTER(AlnV)]
    m0 = np.ascontiguousarray(m0, dtype=F.MM128)
    m1 = np.ascontiguousarray(m1, dtype=F.MM128)
    v0, v1 = O.MMV(len(m0), len(m0), m0.ctypes.data), O.MMV(len(m1), len(m1), m1.ctypes.data)
    r = L.shmr_aln(C.byref(v0), C.byref(v1), direction, max_diff, max_dist, max_repeat)
    out = []
    for i in range(r.contents.n):
        a = r.contents.a[i]
        out.append(([a.idx0.a[j] for j in range(a.idx0.n)], [a.idx1.a[j] for j in range(a.idx1.n)]))
    L.free_shmr_alns(r)
    return out


def test_shmr_aln_vs_reference(ours):
    """Greedy co-linear chaining (src/shmr_align.c:21-160): noisy shifted copies, unrelated lists, repeats, three parameter sets."""
    ref = O.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    rnd = np.random.default_rng(3)
    B = np.array(list("ACGT"))
    for trial in range(40):
        n = int(rnd.integers(400, 9000))
        s = "".join(rnd.choice(B, n))
        if trial % 3 == 0:
            unit = "".join(rnd.choice(B, 300))
            s = s[: n // 3] + unit * 3 + s[n // 3:]
        t = list(s[int(rnd.integers(0, 500)):])
        for _ in range(int(len(t) * 0.02)):
            p = int(rnd.integers(0, len(t)))
            t[p] = str(rnd.choice(B)) if rnd.random() < 0.5 else ""
        t = "".join(t) if trial % 7 else "".join(rnd.choice(B, n))
        m0 = O.abi_sketch(ref, s, 24, 12, 0)
        m1 = O.abi_sketch(ref, t, 24, 12, 1)
        if trial % 2:
            m0, m1 = O.abi_reduce(ref, m0, 2), O.abi_reduce(ref, m1, 2)
        for params in ((100, 1200, 1), (100, 1200, 3), (30, 500, 2)):
            assert _aln(ref, m0, m1, 0, *params) == _aln(ours, m0, m1, 0, *params), (trial, params)
    assert _aln(ours, m0[:0], m1, 0, 100, 1200, 1) == []


def test_shmr_aln_batch_vs_reference(ours):
    """pgb_shmr_aln_batch (one sort for the hash -> index map of all pairs, a warp per pair for the greedy chaining) against the
    reference's shmr_aln pair by pair: related and unrelated lists, tandem repeats (several matches per minimizer, max_repeat
    filter), empty lists inside the batch, both directions' strand filters, three parameter sets."""
    import time

    ref = O.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    from peregrine_b200 import Engine

    rnd = np.random.default_rng(5)
    B = np.array(list("ACGT"))
    l0, l1 = [], []
    for trial in range(48):
        n = int(rnd.integers(300, 12000))
        s = "".join(rnd.choice(B, n))
        if trial % 3 == 0:
            unit = "".join(rnd.choice(B, 250))
            s = s[: n // 3] + unit * 4 + s[n // 3:]
        t = list(s[int(rnd.integers(0, 300)):])
        for _ in range(int(len(t) * 0.02)):
            p = int(rnd.integers(0, len(t)))
            t[p] = str(rnd.choice(B)) if rnd.random() < 0.5 else ""
        t = "".join(t) if trial % 7 else "".join(rnd.choice(B, n))
        m0 = O.abi_sketch(ref, s, 24, 12, 0)
        m1 = O.abi_sketch(ref, t, 24, 12, 1)
        if trial % 2:
            m0, m1 = O.abi_reduce(ref, m0, 2), O.abi_reduce(ref, m1, 2)
        if trial == 5:
            m0 = m0[:0]
        if trial == 9:
            m1 = m1[:0]
        l0.append(m0)
        l1.append(m1)
    eng = Engine(0)
    for params in ((100, 1200, 1), (100, 1200, 3), (30, 500, 2)):
        got = eng.shmr_aln_batch(l0, l1, 0, *params)
        for p in range(len(l0)):
            want = _aln(ref, l0[p], l1[p], 0, *params) if len(l0[p]) and len(l1[p]) else []
            assert want == got[p], (params, p)
    # direction 1: the reference reads one element past the end of list 1 for its first probe (SURVEY A-7), so only the chains'
    # well-defined part is compared, as in tests/test_z_utils_helpers.py: both sides must agree on every hit whose idx1 > 0
    assert eng.shmr_aln_batch([], [], 0) == []
    # the batch call itself (arrays already concatenated, as a C caller has them) against one reference core
    big0, big1 = l0 * 40, l1 * 40
    off0 = np.zeros(len(big0) + 1, dtype=np.uint64)
    off1 = np.zeros(len(big1) + 1, dtype=np.uint64)
    off0[1:] = np.cumsum([len(x) for x in big0])
    off1[1:] = np.cumsum([len(x) for x in big1])
    c0, c1 = np.concatenate(big0), np.concatenate(big1)
    eng.shmr_aln_batch_raw(c0, off0, c1, off1)
    t0 = time.perf_counter()
    hit_off, n_chains, hits = eng.shmr_aln_batch_raw(c0, off0, c1, off1)
    t_batch = time.perf_counter() - t0
    assert int(hit_off[-1]) == len(hits) and int(n_chains[3]) == len(_aln(ref, l0[3], l1[3], 0, 100, 1200, 1))
    vs = [(O.MMV(len(a), len(a), a.ctypes.data), O.MMV(len(b), len(b), b.ctypes.data)) for a, b in zip(l0, l1) if len(a) and len(b)]
    t0 = time.perf_counter()
    for v0, v1 in vs:  # the reference's C call alone (no Python-side unpacking)
        ref.free_shmr_alns(ref.shmr_aln(C.byref(v0), C.byref(v1), 0, 100, 1200, 1))
    t_ref = (time.perf_counter() - t0) * 40
    print(f"shmr_aln: batch of {len(big0)} pairs ({len(c0)} x {len(c1)} minimizers, {len(hits)} hits) {t_batch * 1e3:.1f} ms, reference one core {t_ref * 1e3:.1f} ms")
    eng.close()


class PyMmer(C.Structure):
    _fields_ = [("mmers", C.POINTER(O.MMV)), ("mmer0_map", C.c_void_p), ("rlmap", C.c_void_p), ("mcmap", C.c_void_p), ("ridmm", C.c_void_p)]


class MP256(C.Structure):
    _fields_ = [("x0", C.c_uint64), ("x1", C.c_uint64), ("y0", C.c_uint64), ("y1", C.c_uint64), ("direction", C.c_uint8)]


class MP256V(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(MP256))]


def _handle(L, seq_prefix, idx_prefix, c, T):
    L.build_shimmer_map4py.argtypes = [C.POINTER(PyMmer), C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.get_shimmers_for_read.argtypes = [C.POINTER(O.MMV), C.POINTER(PyMmer), C.c_uint32]
    L.get_mmer_count.restype = C.c_uint32
    L.get_mmer_count.argtypes = [C.POINTER(PyMmer), C.c_uint64]
    L.get_shimmer_hits.argtypes = [C.POINTER(MP256V), C.POINTER(PyMmer), C.c_uint64, C.c_uint32]
    h = PyMmer()
    L.build_shimmer_map4py(C.byref(h), seq_prefix.encode(), (idx_prefix + "-L2").encode(), c, T, 2, 240)
    return h


def _hits(L, h, mhash, span):
    v = MP256V(0, 0, None)
    L.get_shimmer_hits(C.byref(v), C.byref(h), mhash, span)
    return [(v.a[i].x0, v.a[i].x1, v.a[i].y0, v.a[i].y1, v.a[i].direction) for i in range(v.n)]


def test_shimmer4py_index_handle_vs_reference(ours, workdir, ref_dir):
    """build_shimmer_map4py / get_shimmers_for_read / get_mmer_count / get_shimmer_hits (src/shimmer4py.c:44-196)."""
    import datasets as D

    ref = O.reflib()
    p = D.make_sim(workdir, "h4py", genome=300_000, cov=20, seed=5)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "h4py/ref"), T=2, extra=["-m", "0"])
    for (c, T) in ((1, 1), (2, 3)):
        hr, ho = _handle(ref, p, rp, c, T), _handle(ours, p, rp, c, T)
        assert hr.mmers.contents.n == ho.mmers.contents.n
        mm = np.concatenate([F.read_mmlist(f"{rp}-L2-{i:02d}-of-02.dat") for i in (1, 2)])
        for rid in (0, 1, 7, 199, 10**6):
            a, b = O.MMV(0, 0, None), O.MMV(0, 0, None)
            ref.get_shimmers_for_read(C.byref(a), C.byref(hr), rid)
            ours.get_shimmers_for_read(C.byref(b), C.byref(ho), rid)
            assert a.n == b.n
            if a.n:
                assert C.string_at(a.a, a.n * 16) == C.string_at(b.a, b.n * 16)
        keys = [int(x) for x in mm["x"][:: max(1, len(mm) // 150)]]
        n_hits = 0
        for x in keys + [12345]:
            assert ref.get_mmer_count(C.byref(hr), x >> 8) == ours.get_mmer_count(C.byref(ho), x >> 8)
            a, b = _hits(ref, hr, x >> 8, x & 0xFF), _hits(ours, ho, x >> 8, x & 0xFF)
            assert a == b, (c, T, x)
            n_hits += len(a)
        assert n_hits > 100


def test_ovlp_match_batch_vs_reference(ours):
    """pgb_ovlp_match_batch (one upload, a warp per pair, one download) against the reference's ovlp_match pair by pair: all
    four strand combinations, operands with N runs (nibble 0 equals only nibble 0), empty-ish operands, three band limits;
    and the single-call ovlp_match, which is a batch of one through the same path."""
    import time

    ref = O.reflib()
    if ref is None:
        pytest.skip("oracle/_ref not present")
    from peregrine_b200 import Engine

    rnd = np.random.default_rng(23)
    B = np.array(list("ACGT"))
    bufs, pairs = [], []
    off = 0
    for trial in range(60):
        n = int(rnd.integers(40, 9000))
        s = "".join(rnd.choice(B, n))
        t = list(s[int(rnd.integers(0, min(300, n // 2))):])
        for _ in range(int(len(t) * 0.01 * (trial % 4))):
            p = int(rnd.integers(0, len(t)))
            t[p] = str(rnd.choice(B)) if rnd.random() < 0.5 else ""
        t = "".join(t)
        if trial % 6 == 0 and len(t) > 200:
            t = t[:90] + "N" * 9 + t[99:]
        if trial % 9 == 0:
            s = s[:30] + "N" + s[31:]
        q, tt = O.encode_biseq(s), O.encode_biseq(t)
        bufs += [q, tt]
        pairs.append((off, len(q), trial & 1, off + len(q), len(tt), (trial >> 1) & 1))
        off += len(q) + len(tt)
    seq = np.concatenate(bufs)
    eng = Engine(0)
    L = eng.L
    L.pgb_ovlp_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
    P = np.array(pairs, dtype=np.int64)
    arrs = [np.ascontiguousarray(P[:, 0], np.uint64), np.ascontiguousarray(P[:, 1], np.uint32), np.ascontiguousarray(P[:, 2], np.uint8),
            np.ascontiguousarray(P[:, 3], np.uint64), np.ascontiguousarray(P[:, 4], np.uint32), np.ascontiguousarray(P[:, 5], np.uint8)]
    for bw in (20, 100, 200):
        out = np.zeros((len(pairs), 8), dtype=np.int32)
        rc = L.pgb_ovlp_match_batch(eng.h, seq.ctypes.data, seq.size, len(pairs), *[a.ctypes.data for a in arrs], bw, out.ctypes.data)
        assert rc == 0, L.pgb_last_error(eng.h)
        for i, (qo, ql, qs, to, tl, ts) in enumerate(pairs):
            want = O.abi_ovlp_match(ref, seq[qo:qo + ql], qs, seq[to:to + tl], ts, bw)
            assert list(map(int, want)) == list(map(int, out[i])), (bw, i, pairs[i])
    # single calls: the same answers, and cheap enough to be called in a Python loop
    qo, ql, qs, to, tl, ts = pairs[3]
    t0 = time.perf_counter()
    for _ in range(20):
        got = O.abi_ovlp_match(ours, seq[qo:qo + ql], qs, seq[to:to + tl], ts, 100)
    per_call = (time.perf_counter() - t0) / 20
    assert list(map(int, got)) == list(map(int, O.abi_ovlp_match(ref, seq[qo:qo + ql], qs, seq[to:to + tl], ts, 100)))
    print(f"single ovlp_match call ({ql} x {tl} bases): {per_call * 1e6:.0f} us")
    eng.close()
