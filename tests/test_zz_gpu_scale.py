"""GPU parity at the sizes and shapes the bench runs at (VERDICT r1 "parity gaps"): BASELINE.json configs[0] as specified
(the reference's own simulator, (T_idx, T_ovlp) = (12, 8) and (1, 1)), the routed multi-GPU path at T = 8 on a 15 Mb set with
default table sizes, configs[1] itself (50 Mb, T = 1) and a real 2-rank NCCL run against `shmr_overlap -t 2`.  Checker: the
unmodified reference in oracle/_ref.  They sort last: the slow ones must not hide the established tests behind `pytest -x`."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import datasets as D
import ecoli_standin as ES
from peregrine_b200 import formats as F
from test_gpu_parity import assert_same_bytes, assert_same_ovlp

pytestmark = pytest.mark.gpu
ROOT = D.ROOT


def test_config1_ecoli_standin(workdir, ref_dir):
    """configs[0]: stand-in genome (4,639,675 bp) through the reference's simulator (1 % error, 8 x 623 reads), then the shell
    recipe of test/ecoli_K12/run_test.sh:18-28 — shmr_mkseqdb, shmr_index -t 12 (:21), shmr_overlap -t 8 (:25), cat | shmr_dedup —
    once with bin/ and once with the reference; and the single-chunk case the config names (T_idx = T_ovlp = 1)."""
    lst, paths = ES.make_set(workdir)
    pin = json.load(open(os.path.join(ROOT, "tests", "golden", "ecoli_standin.json")))
    assert ES.sha256_files(paths) == pin["sha256_reference_script"], "the restated simulator no longer matches simulate_reads.py"
    outs = {}
    for ti, to in ((12, 8), (1, 1)):
        for tag, bindir in (("ref", ref_dir), ("our", os.path.join(ROOT, "bin"))):
            wd = os.path.join(workdir, "ecoli_standin", f"{tag}_{ti}_{to}")
            subprocess.run(["bash", os.path.join(ROOT, "tools", "run_chain.sh"), lst, wd, str(ti), str(to), "4"], env=dict(os.environ, BIN=bindir),
                           check=True, stdout=subprocess.DEVNULL)
            outs[tag] = wd
        for rel in ["index/seq_dataset.idx", "index/seq_dataset.seqdb", "asm/preads.ovl"] + [f"index/shmr-L2-{c:02d}-of-{ti:02d}.dat" for c in range(1, ti + 1)]:
            assert_same_bytes(os.path.join(outs["our"], rel), os.path.join(outs["ref"], rel))
        n = 0
        for c in range(1, to + 1):
            assert_same_ovlp(os.path.join(outs["our"], "ovlp", f"ovlp.{c:02d}"), os.path.join(outs["ref"], "ovlp", f"ovlp.{c:02d}"))
            n += os.path.getsize(os.path.join(outs["ref"], "ovlp", f"ovlp.{c:02d}")) // 64
        assert n > 40_000, n  # 53,360 records at T=1, 249,251 at T=8 on the survey's stand-in


@pytest.fixture(scope="module")
def sim15(workdir):
    return D.make_sim(workdir, "sim15", genome=15_000_000, cov=30, seed=1234)


@pytest.fixture(scope="module")
def ref15_t8(sim15, workdir, ref_dir):
    """the reference on the 15 Mb set as 8 index chunks and 8 overlap chunks (processes in parallel, like pg_run.py runs them)"""
    from concurrent.futures import ThreadPoolExecutor

    T = 8
    out = os.path.join(workdir, "sim15", "ref8")
    os.makedirs(out, exist_ok=True)
    with ThreadPoolExecutor(T) as ex:
        list(ex.map(lambda c: D.run([os.path.join(ref_dir, "shmr_index"), "-p", sim15, "-t", str(T), "-c", str(c), "-o", os.path.join(out, "shmr"), "-m", "0"]),
                    range(1, T + 1)))
        list(ex.map(lambda c: D.run([os.path.join(ref_dir, "shmr_overlap"), "-p", sim15, "-l", os.path.join(out, "shmr-L2"), "-t", str(T), "-c", str(c),
                                     "-o", os.path.join(out, f"ovlp.{c:02d}")]), range(1, T + 1)))
    return [os.path.join(out, f"ovlp.{c:02d}") for c in range(1, T + 1)]


def test_routed_T8_15Mb_default_tables(sim15, ref15_t8):
    """The multi-GPU data path of bench.py --gpus 8 on ONE device, at a size where the per-chunk replay tables matter (rid_pairs
    is per chunk: a read pair is aligned in up to T chunks): eight 'ranks' index their reads, sum the partial count tables,
    route the SHIMMER-pair records to the owning chunk (same concatenation as the all-to-all), and every owner runs
    pgb_overlap_routed with DEFAULT table sizing.  Records per chunk must equal shmr_overlap -t 8 -c c."""
    import torch
    from peregrine_b200 import Engine, multigpu as M

    T = 8
    dev = torch.device("cuda", 0)
    rid, ln, off = F.read_idx(sim15 + ".idx")
    seqdb = np.fromfile(sim15 + ".seqdb", dtype=np.uint8)
    parts, counts, l2 = [], [], []
    eng = Engine(0)
    for r in range(T):  # one engine, one rank at a time: its exports are copied out before the next load
        eng.load_reads(seqdb, rid, ln, off, T, r + 1)
        eng.index(80, 16, 6, 2)
        eng.set_shimmers_from_index(2)
        parts.append(M.export_reads(eng, dev))
        counts.append(M.export_counts(eng, dev))
        l2.append(M.export_level(eng, 2, dev))
    all_counts = torch.cat(counts).contiguous()
    torch.cuda.synchronize()  # torch's stream -> the library's stream (the engine is handed raw pointers below)
    has_first, sends, splits = [], [], []
    for r in range(T):
        M.import_reads(eng, parts[r])
        eng.set_shimmers_device(l2[r].data_ptr(), int(l2[r].shape[0]))
        eng.counts_set_device(all_counts.data_ptr(), int(all_counts.shape[0]))
        has_first.append(eng.route_scan(2, 240))
        per_chunk = eng.route_build(T, 2, 240, any(has_first[:r]))
        send = M.export_route(eng, dev)
        assert sum(per_chunk) == send.shape[0]
        sends.append(send)
        splits.append(np.concatenate([[0], np.cumsum(per_chunk)]))
    M.import_reads(eng, M.concat_reads(parts))
    total = 0
    for d in range(T):  # owner of chunk d + 1
        recv = torch.cat([sends[src][int(splits[src][d]): int(splits[src][d + 1])] for src in range(T)]).contiguous()
        torch.cuda.synchronize()  # (without it the library may read recv before torch.cat has written it: seen once as missing records)
        ov = eng.overlap_routed(recv.data_ptr(), int(recv.shape[0]), total_chunk=T)
        want = F.normalise_ovlp(F.read_ovlp(ref15_t8[d]))
        assert len(ov) == len(want) and ov.tobytes() == want.tobytes(), f"chunk {d + 1}: {len(ov)} records vs reference {len(want)}"
        total += len(ov)
    st = eng.stats()
    eng.close()
    assert total > 1_000_000, total  # x4.5 of the T=1 count
    print(f"routed T=8 on 15 Mb: {total} records identical; {st['n_replay_passes']} replay passes over 8 chunks")


def test_config2_50Mb_single_chunk(workdir, ref_dir):
    """configs[1] at full size: 50 Mb genome, 30x, T = 1 — the exact workload bench.py times (outer khash of ~195 k keys with
    several rehashes, inner groups beyond the 48-key device replay, 1.4 M alignments).  The reference needs ~70 core-seconds."""
    from peregrine_b200 import Engine

    p = D.make_sim(workdir, "g50", genome=50_000_000, cov=30)
    out = os.path.join(workdir, "g50", "ref")
    rp = D.ref_index(ref_dir, p, out, T=1, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, p, rp, 2, out, T=1)
    rid, ln, off = F.read_idx(p + ".idx")
    eng = Engine(0)
    eng.load_reads(np.fromfile(p + ".seqdb", dtype=np.uint8), rid, ln, off)
    eng.index(80, 16, 6, 2)
    assert np.array_equal(eng.level(2), F.read_mmlist(rp + "-L2-01-of-01.dat")), "L2 differs from the reference"
    eng.set_shimmers_from_index(2)
    ov = eng.overlap(1, 1)
    eng.close()
    want = F.normalise_ovlp(F.read_ovlp(ro[0]))
    assert len(ov) == len(want) and ov.tobytes() == want.tobytes(), f"{len(ov)} records vs reference {len(want)}"
    assert len(ov) > 1_200_000


def test_two_rank_nccl_matches_reference(sim15, workdir, ref_dir):
    """A real 2-process NCCL run of peregrine_b200.multigpu.ShardedJob (the code path of bench.py --gpus 2, routed exchange):
    rank r's records must equal the reference's shmr_overlap -t 2 -c r+1 over two index chunk files."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    T = 2
    out = os.path.join(workdir, "sim15", "ref2")
    rp = D.ref_index(ref_dir, sim15, out, T=T, extra=["-m", "0"])
    ro = D.ref_overlap(ref_dir, sim15, rp, 2, out, T=T)
    our = os.path.join(workdir, "sim15", "our2")
    os.makedirs(our, exist_ok=True)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(T), "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "sharded_overlap.py"), "--prefix", sim15, "--out", our]
    subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
    for c in range(1, T + 1):
        assert_same_ovlp(os.path.join(our, f"ovlp.{c:02d}"), ro[c - 1])


def test_decode_biseq_abi(ref_dir):
    """decode_biseq through the C ABI (src/shmr_utils.c:56-62): both strands, every nibble value incl. the invalid ones,
    against the reference's own function in libshimmer_ref.so."""
    import ctypes as C

    ours = C.CDLL(os.path.join(ROOT, "peregrine_b200", "libpgb200.so"))
    ref = C.CDLL(os.path.join(ref_dir, "libshimmer_ref.so"))
    rng = np.random.default_rng(5)
    fwd, rev = np.array([1, 2, 4, 8], dtype=np.uint8), np.array([8, 4, 2, 1], dtype=np.uint8)
    for n in (0, 1, 7, 1000, 40001):
        b = rng.integers(0, 4, n)
        src = (fwd[b] | (rev[b[::-1]] << 4)).astype(np.uint8) if n else np.zeros(0, np.uint8)
        cases = [src]
        if n:
            junk = src.copy()
            junk[rng.integers(0, n, max(1, n // 10))] = rng.integers(0, 256, max(1, n // 10)).astype(np.uint8)  # any byte value
            cases.append(junk)
        for s in cases:
            for strand in (0, 1):
                a, r = C.create_string_buffer(n + 1), C.create_string_buffer(n + 1)
                for lib, dst in ((ours, a), (ref, r)):
                    lib.decode_biseq.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint8]
                    lib.decode_biseq.restype = None
                    lib.decode_biseq(s.ctypes.data_as(C.c_void_p), dst, n, strand)
                assert a.raw[:n] == r.raw[:n], (n, strand)
    # the round trip of a real read: encode (reference layout) -> decode strand 1 = reverse complement
    seq = b"ACGTTGCANNACGT"
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    lut = {65: 1, 67: 2, 71: 4, 84: 8, 78: 0}
    lutr = {65: 8, 67: 4, 71: 2, 84: 1, 78: 0}
    enc = np.array([lut[seq[p]] | (lutr[seq[len(seq) - 1 - p]] << 4) for p in range(len(seq))], dtype=np.uint8)
    out = C.create_string_buffer(len(seq) + 1)
    ours.decode_biseq(enc.ctypes.data_as(C.c_void_p), out, len(seq), 0)
    assert out.raw[:len(seq)] == seq
    ours.decode_biseq(enc.ctypes.data_as(C.c_void_p), out, len(seq), 1)
    assert out.raw[:len(seq)] == seq.translate(comp)[::-1]
