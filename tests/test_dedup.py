"""shmr_dedup: raw ovlp_t stream -> preads.ovl text, byte-for-byte against the unmodified reference binary
(oracle/_ref/shmr_dedup, src/shmr_dedup.c).

CPU: the product's formatting / pair-key functions compiled for the host (tests/hostsim `dedup`).
GPU: libpgb200's pgb_dedup* entry points and the bin/shmr_dedup drop-in."""
import os
import subprocess

import numpy as np
import pytest

import datasets as D
from peregrine_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ref_dedup(ref_dir, stream: bytes) -> bytes:
    return subprocess.run([os.path.join(ref_dir, "shmr_dedup")], input=stream, stdout=subprocess.PIPE, check=True).stdout


def adversarial_stream(n=6000, seed=5):
    """Records that exercise every branch and wrap of the reference's arithmetic: both strands on both sides, coordinates
    beyond the read ends (unsigned clamps), negative intermediate values printed with %d, read ids above 2^31 (%09d of a
    negative int), m_size 0 (nan / inf), err_est values on x.x5 ties, repeated pairs in both orders."""
    rng = np.random.default_rng(seed)
    a = np.zeros(n, dtype=F.OVLP)
    rid0 = rng.integers(0, 300, n).astype(np.uint64)
    rid1 = rng.integers(0, 300, n).astype(np.uint64)
    big = rng.random(n) < 0.02
    rid0[big] = rng.integers(2**31, 2**32 - 2, big.sum()).astype(np.uint64)
    pos0 = rng.integers(0, 30000, n).astype(np.uint64)
    pos1 = rng.integers(0, 30000, n).astype(np.uint64)
    a["y0"] = (rid0 << np.uint64(32)) | (pos0 << np.uint64(1)) | rng.integers(0, 2, n).astype(np.uint64)
    a["y1"] = (rid1 << np.uint64(32)) | (pos1 << np.uint64(1)) | rng.integers(0, 2, n).astype(np.uint64)
    a["rl0"] = rng.integers(1, 40000, n)
    a["rl1"] = rng.integers(1, 40000, n)
    a["strand0"] = rng.integers(0, 2, n)
    a["strand1"] = rng.integers(0, 2, n)
    a["ovlp_type"] = rng.integers(0, 3, n)
    m = a["match"]
    m["q_bgn"] = rng.integers(0, 60, n)
    m["t_bgn"] = rng.integers(0, 60, n)
    m["q_end"] = rng.integers(0, 45000, n)
    m["t_end"] = rng.integers(0, 45000, n)
    m["dist"] = rng.integers(0, 4000, n)
    m["m_size"] = rng.integers(1, 45000, n)
    # ties of "%0.1f": 100 - 100 * d / m with m = 400 * j, d = j * (2 i + 1) -> x.25 / x.75 exactly; m = 2000 j -> x.x5
    t = rng.random(n) < 0.2
    j = rng.integers(1, 20, n)
    i = rng.integers(0, 150, n)
    m["m_size"][t] = (2000 * j)[t]
    m["dist"][t] = (j * (2 * i + 1))[t]
    z = rng.random(n) < 0.01
    m["m_size"][z] = 0
    m["dist"][z & (rng.random(n) < 0.5)] = 0
    neg = rng.random(n) < 0.01
    m["m_size"][neg] = -rng.integers(1, 1000, neg.sum())
    return a


def real_stream(workdir, ref_dir):
    p = D.make_sim(workdir, "dd", genome=400_000, cov=20)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "dd/ref"), T=2, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, "dd/ref"), T=3)
    return b"".join(open(f, "rb").read() for f in ro)  # `cat ovlp-*.dat`


def test_dedup_logic_on_host(workdir, ref_dir):
    subprocess.check_call(["make", "-C", ROOT, "hostsim"], stdout=subprocess.DEVNULL)
    hs = os.path.join(ROOT, "build", "hostsim")
    for stream in (real_stream(workdir, ref_dir), adversarial_stream().tobytes(), adversarial_stream(1, seed=9).tobytes()):
        want = ref_dedup(ref_dir, stream)
        got = subprocess.run([hs, "dedup"], input=stream, stdout=subprocess.PIPE, check=True).stdout
        assert len(want) > 0 and got == want


@pytest.mark.gpu
def test_dedup_gpu_matches_reference(workdir, ref_dir):
    from peregrine_b200 import Engine

    eng = Engine(0)
    for stream in (real_stream(workdir, ref_dir), adversarial_stream().tobytes(), adversarial_stream(1, seed=9).tobytes(),
                   adversarial_stream(200_000, seed=3).tobytes()):
        want = ref_dedup(ref_dir, stream)
        got = eng.dedup(np.frombuffer(stream, dtype=F.OVLP))
        assert got == want
        assert eng.stats()["n_dedup_kept"] > 0
    assert eng.dedup(np.zeros(0, dtype=F.OVLP)) == b""  # empty stream: nothing (the reference prints an uninitialised record)
    eng.close()


@pytest.mark.gpu
def test_dedup_cli_and_in_place(workdir, ref_dir):
    """bin/shmr_dedup as `cat ovlp-*.dat | shmr_dedup`, and the device-resident form over the engine's own overlap output."""
    from peregrine_b200 import Engine

    stream = real_stream(workdir, ref_dir)
    want = ref_dedup(ref_dir, stream)
    got = subprocess.run([os.path.join(ROOT, "bin", "shmr_dedup")], input=stream, stdout=subprocess.PIPE, check=True).stdout
    assert got == want
    p = os.path.join(workdir, "dd", "seq")
    rid, ln, off = F.read_idx(p + ".idx")
    seqdb = np.fromfile(p + ".seqdb", dtype=np.uint8)
    eng = Engine(0)
    eng.load_reads(seqdb, rid, ln, off)
    eng.index(80, 16, 6, 2)
    eng.set_shimmers_from_index(2)
    ov = eng.overlap(1, 1)
    assert eng.dedup() == ref_dedup(ref_dir, ov.tobytes())
    eng.close()


@pytest.mark.gpu
def test_dedup_in_bounded_batches(workdir, ref_dir):
    """bin/shmr_dedup reads stdin in batches (PGB_DEDUP_BATCH records) and keeps the pair table on the device between them: batch
    sizes from 1 record up (the table grows and is rehashed several times on the way), a stream with a truncated trailing
    record, and the stream API of the library; the text must be the reference's for the whole stream."""
    import ctypes as C

    from peregrine_b200 import Engine

    tool = os.path.join(ROOT, "bin", "shmr_dedup")
    big = adversarial_stream(200_000, seed=3).tobytes()
    for stream, batches in ((real_stream(workdir, ref_dir), ("7", "1000", "65536")), (big, ("3001", "50000")), (adversarial_stream(40, seed=2).tobytes(), ("1",))):
        want = ref_dedup(ref_dir, stream)
        for b in batches:
            got = subprocess.run([tool], input=stream, stdout=subprocess.PIPE, check=True, env=dict(os.environ, PGB_DEDUP_BATCH=b)).stdout
            assert got == want, (len(stream), b)
    cut = big[: 64 * 1000 + 17]  # 1000 records and a torn one
    got = subprocess.run([tool], input=cut, stdout=subprocess.PIPE, check=True, env=dict(os.environ, PGB_DEDUP_BATCH="333")).stdout
    assert got == ref_dedup(ref_dir, cut[: 64 * 1000])
    eng = Engine(0)
    L = eng.L
    for f in (L.pgb_dedup_stream_begin, L.pgb_dedup_stream_end):
        f.argtypes = [C.c_void_p]
    L.pgb_dedup_stream_push.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    recs = np.frombuffer(big, dtype=F.OVLP)
    assert L.pgb_dedup_stream_begin(eng.h) == 0
    out = []
    for lo in range(0, len(recs), 77_777):
        part = np.ascontiguousarray(recs[lo: lo + 77_777])
        assert L.pgb_dedup_stream_push(eng.h, part.ctypes.data, len(part)) == 0, L.pgb_last_error(eng.h)
        buf = np.empty(L.pgb_dedup_text_bytes(eng.h), dtype=np.uint8)
        if buf.size:
            assert L.pgb_dedup_text_copy(eng.h, buf.ctypes.data) == 0
        out.append(buf.tobytes())
    assert L.pgb_dedup_stream_end(eng.h) == 0
    assert b"".join(out) == ref_dedup(ref_dir, big)
    eng.close()
