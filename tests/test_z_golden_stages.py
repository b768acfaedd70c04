"""Committed golden vectors of shmr_mkseqdb / shmr_dedup / shmr_map (tests/golden/make_golden_stages.py generated them from
the unmodified reference): these pin the stages on machines where /root/reference and oracle/_ref do not exist.

CPU: the host-compiled per-item functions (tests/hostsim) against the stored answers.
GPU: the tool chain bin/shmr_mkseqdb -> bin/shmr_index -> bin/shmr_map and the dedup entry points against them."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from peregrine_b200 import formats as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
G = json.load(open(os.path.join(GOLD, "golden_stages.json")))
BIN = os.path.join(ROOT, "bin")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def write_list(path, names):
    with open(path, "w") as f:
        for n in names:
            f.write(os.path.join(GOLD, n) + "\n")
    return path


@pytest.fixture(scope="module")
def hostsim():
    subprocess.check_call(["make", "-C", ROOT, "hostsim"], stdout=subprocess.DEVNULL)
    return os.path.join(ROOT, "build", "hostsim")


def test_golden_dedup_on_host(hostsim):
    stream = open(os.path.join(GOLD, "dedup_in.bin"), "rb").read()
    got = subprocess.run([hostsim, "dedup"], input=stream, stdout=subprocess.PIPE, check=True).stdout
    assert got == open(os.path.join(GOLD, "dedup_expected.txt"), "rb").read()
    assert got.count(b"\n") == G["dedup"]["lines"] and sha(got) == G["dedup"]["sha256"]


def test_golden_fasta_scanner_on_host(hostsim, tmp_path):
    for tag in ("tricky", "reads"):
        g = G["mkseqdb_" + tag]
        lst = write_list(str(tmp_path / (tag + ".lst")), g["files"])
        got = subprocess.run([hostsim, "fastaidx", lst], stdout=subprocess.PIPE, check=True).stdout
        assert sha(got) == g["idx_sha256"] and got.count(b"\n") == g["n_records"]
        if g["idx_text"]:
            assert got.decode() == g["idx_text"]


@pytest.mark.gpu
def test_golden_mkseqdb_gpu(tmp_path):
    for tag in ("tricky", "reads"):
        g = G["mkseqdb_" + tag]
        lst = write_list(str(tmp_path / (tag + ".lst")), g["files"])
        subprocess.run([os.path.join(BIN, "shmr_mkseqdb"), "-d", lst, "-p", str(tmp_path / tag)], check=True, stdout=subprocess.DEVNULL)
        assert sha(open(tmp_path / (tag + ".idx"), "rb").read()) == g["idx_sha256"]
        db = open(tmp_path / (tag + ".seqdb"), "rb").read()
        assert len(db) == g["bases"] and sha(db) == g["seqdb_sha256"]


@pytest.mark.gpu
def test_golden_dedup_gpu():
    from peregrine_b200 import Engine

    stream = np.fromfile(os.path.join(GOLD, "dedup_in.bin"), dtype=F.OVLP)
    want = open(os.path.join(GOLD, "dedup_expected.txt"), "rb").read()
    eng = Engine(0)
    assert eng.dedup(stream) == want
    eng.close()
    got = subprocess.run([os.path.join(BIN, "shmr_dedup")], input=stream.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    assert got == want


@pytest.mark.gpu
def test_golden_tool_chain_mkseqdb_index_map(tmp_path):
    """FASTA -> bin/shmr_mkseqdb -> bin/shmr_index -> bin/shmr_map, every step ours, against the reference's stored output."""
    g = G["map"]
    recs, name = [], None
    for line in open(os.path.join(GOLD, "reads.fa")):
        if line.startswith(">"):
            name = line[1:].strip()
        else:
            recs.append((name, line.strip()))
    with open(tmp_path / "ctg.fa", "w") as f:
        for n, s in recs[: g["n_contigs"]]:
            f.write(f">{n}\n{s}\n")
    (tmp_path / "ctg.lst").write_text(str(tmp_path / "ctg.fa") + "\n")
    write_list(str(tmp_path / "reads.lst"), ["reads.fa"])
    run = lambda cmd: subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE).stdout  # noqa: E731
    run([os.path.join(BIN, "shmr_mkseqdb"), "-d", str(tmp_path / "reads.lst"), "-p", str(tmp_path / "reads")])
    run([os.path.join(BIN, "shmr_mkseqdb"), "-d", str(tmp_path / "ctg.lst"), "-p", str(tmp_path / "ctg")])
    for prefix, out in (("reads", "ridx"), ("ctg", "cidx")):
        run([os.path.join(BIN, "shmr_index"), "-p", str(tmp_path / prefix), "-t", "1", "-c", "1", "-o", str(tmp_path / out), *g["index_args"]])
    got = run([os.path.join(BIN, "shmr_map"), "-r", str(tmp_path / "ctg"), "-m", str(tmp_path / "cidx-L2"), "-p", str(tmp_path / "reads"),
               "-l", str(tmp_path / "ridx-L2")])
    assert got == open(os.path.join(GOLD, "map_expected.txt"), "rb").read()
    assert got.count(b"\n") == g["lines"] and sha(got) == g["sha256"]
