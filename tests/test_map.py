"""shmr_map: contig shimmers against the reads' SHIMMER-pair index, byte-for-byte against the unmodified
reference binary (oracle/_ref/shmr_map, src/shmr_map.c) on the same files."""
import os
import subprocess

import numpy as np
import pytest

import datasets as D
from peregrine_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
COMP = str.maketrans("ACGT", "TGCA")


def make_case(workdir, ref_dir, name="map", genome_len=300_000, cov=20, seed=23):
    """reads (both strands, 0.5 % substitutions) of a random genome + 'contigs' = pieces of that genome, one of them
    reverse-complemented and one carrying a few edits, all pushed through the reference's own mkseqdb / index tools."""
    rng = np.random.default_rng(seed)
    g = "".join(rng.choice(list("ACGT"), genome_len))
    reads = []
    for i in range(int(genome_len * cov / 12000)):
        ln = int(rng.integers(8000, 16000))
        st = int(rng.integers(0, genome_len - ln))
        s = np.array(list(g[st: st + ln]))
        mut = rng.random(ln) < 0.005
        s[mut] = rng.choice(list("ACGT"), int(mut.sum()))
        s = "".join(s)
        if rng.random() < 0.5:
            s = s.translate(COMP)[::-1]
        reads.append((f"r{i}", s))
    reads_p = D.make_from_fasta(workdir, name + "_reads", reads, ref_dir)
    c2 = np.array(list(g[100_000:200_000]))
    mut = rng.random(len(c2)) < 0.002
    c2[mut] = rng.choice(list("ACGT"), int(mut.sum()))
    ctgs = [("ctg0", g[:100_000]), ("ctg1", "".join(c2)), ("ctg2", g[200_000:].translate(COMP)[::-1]), ("tiny", g[5000:5300])]
    ctg_p = D.make_from_fasta(workdir, name + "_ctg", ctgs, ref_dir)
    return reads_p, ctg_p


def run_map(tool, ctg_p, ctg_idx, reads_p, reads_idx, extra=()):
    r = subprocess.run([tool, "-r", ctg_p, "-m", ctg_idx + "-L2", "-p", reads_p, "-l", reads_idx + "-L2", *extra], stdout=subprocess.PIPE, stderr=subprocess.PIPE, check=True)
    return r.stdout, r.stderr


@pytest.fixture(scope="module")
def case(workdir, ref_dir):
    reads_p, ctg_p = make_case(workdir, ref_dir)
    reads_idx = D.ref_index(ref_dir, reads_p, os.path.join(workdir, "map_reads/idx"), T=3, extra=["-m", "0"])
    ctg_idx = D.ref_index(ref_dir, ctg_p, os.path.join(workdir, "map_ctg/idx"), T=1, extra=["-m", "0"])
    return reads_p, ctg_p, reads_idx, ctg_idx


@pytest.mark.parametrize("extra", [[], ["-M", "60", "-n", "2"], ["-t", "2", "-c", "1"], ["-t", "2", "-c", "2"]])
def test_map_cli_matches_reference(case, ref_dir, extra):
    reads_p, ctg_p, reads_idx, ctg_idx = case
    want, want_err = run_map(os.path.join(ref_dir, "shmr_map"), ctg_p, ctg_idx, reads_p, reads_idx, extra)
    got, got_err = run_map(os.path.join(ROOT, "bin", "shmr_map"), ctg_p, ctg_idx, reads_p, reads_idx, extra)
    assert want.count(b"\n") > 1000
    assert got == want
    assert got_err == want_err  # same progress messages on stderr


def test_map_engine_api(case, ref_dir):
    from peregrine_b200 import Engine

    reads_p, ctg_p, reads_idx, ctg_idx = case
    want, _ = run_map(os.path.join(ref_dir, "shmr_map"), ctg_p, ctg_idx, reads_p, reads_idx)
    rid, ln, _ = F.read_idx(reads_p + ".idx")
    mm = np.concatenate([F.read_mmlist(f"{reads_idx}-L2-{c:02d}-of-03.dat") for c in (1, 2, 3)])
    mc = np.concatenate([F.read_mc(f"{reads_idx}-L2-MC-{c:02d}-of-03.dat") for c in (1, 2, 3)])
    ref_mm = F.read_mmlist(f"{ctg_idx}-L2-01-of-01.dat")
    eng = Engine(0)
    eng.set_read_lengths(rid, ln)
    eng.set_shimmers(mm, mc)
    assert eng.map(ref_mm) == want
    assert eng.stats()["n_map_hits"] == want.count(b"\n")
    assert eng.map(ref_mm[:0]) == b""
    eng.close()
