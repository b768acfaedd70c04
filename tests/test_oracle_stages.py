"""Pins oracle/stages_oracle.c (the plain-C restatement of shmr_mkseqdb / shmr_dedup / shmr_map) to the unmodified reference
binaries (oracle/_ref) and to the committed golden vectors (tests/golden/golden_stages.json).  CPU only."""
import hashlib
import json
import os
import subprocess

import numpy as np

import datasets as D
import oracle as O
from peregrine_b200 import formats as F
from test_dedup import adversarial_stream, ref_dedup
from test_mkseqdb import ref_mkseqdb, tricky_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
G = json.load(open(os.path.join(GOLD, "golden_stages.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_oracle_mkseqdb_vs_reference_and_golden(tmp_path, ref_dir):
    lst = tricky_inputs(str(tmp_path))
    want_idx, want_db = ref_mkseqdb(ref_dir, lst, str(tmp_path / "ref"))
    idx, db = O.orc_mkseqdb([l.strip() for l in open(lst)])
    assert idx == want_idx and db == want_db
    for tag in ("tricky", "reads"):
        g = G["mkseqdb_" + tag]
        idx, db = O.orc_mkseqdb([os.path.join(GOLD, n) for n in g["files"]])
        assert sha(idx) == g["idx_sha256"] and sha(db) == g["seqdb_sha256"]


def test_oracle_dedup_vs_reference_and_golden(workdir, ref_dir):
    for stream in (adversarial_stream(), adversarial_stream(1, seed=9), adversarial_stream(50_000, seed=4)):
        assert O.orc_dedup(stream) == ref_dedup(ref_dir, stream.tobytes())
    gold_in = np.fromfile(os.path.join(GOLD, "dedup_in.bin"), dtype=F.OVLP)
    assert O.orc_dedup(gold_in) == open(os.path.join(GOLD, "dedup_expected.txt"), "rb").read()


def test_oracle_map_vs_reference(workdir, ref_dir):
    """reads of a random genome against pieces of that genome as contigs (the case of tests/test_map.py), T = 1 and 2."""
    from test_map import make_case, run_map

    reads_p, ctg_p = make_case(workdir, ref_dir, name="omap", genome_len=300_000, cov=12)
    reads_idx = D.ref_index(ref_dir, reads_p, os.path.join(workdir, "omap_reads/idx"), T=2, extra=["-m", "0"])
    ctg_idx = D.ref_index(ref_dir, ctg_p, os.path.join(workdir, "omap_ctg/idx"), T=1, extra=["-m", "0"])
    rid, ln, _ = F.read_idx(reads_p + ".idx")
    mm = np.concatenate([F.read_mmlist(f"{reads_idx}-L2-{c:02d}-of-02.dat") for c in (1, 2)])
    mc = np.concatenate([F.read_mc(f"{reads_idx}-L2-MC-{c:02d}-of-02.dat") for c in (1, 2)])
    ref_mm = F.read_mmlist(f"{ctg_idx}-L2-01-of-01.dat")
    for extra, kw in (([], {}), (["-M", "40", "-n", "2"], dict(lower=2, upper=40)), (["-t", "2", "-c", "2"], dict(T=2, c=2))):
        want, _ = run_map(os.path.join(ref_dir, "shmr_map"), ctg_p, ctg_idx, reads_p, reads_idx, extra)
        assert want.count(b"\n") > 500
        assert O.orc_map(ref_mm, mm, mc, rid, ln, **kw) == want
