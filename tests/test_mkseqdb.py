"""shmr_mkseqdb: FASTA/FASTQ(.gz) -> .idx + .seqdb, byte-for-byte against the unmodified reference binary
(oracle/_ref/shmr_mkseqdb, src/shmr_mkseqdb.c + kseq.h).

CPU: the product's record scanner (peregrine_b200/csrc/fasta_reader.hpp, compiled into tests/hostsim) must produce the
     reference's .idx on inputs that exercise kseq's grammar.
GPU: bin/shmr_mkseqdb (scanner + k_encode_biseq) must produce identical .idx and .seqdb files; pgb_encode_biseq against
     the reference's encoding of every byte value."""
import gzip
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tricky_inputs(d):
    """Returns the list-file path.  Every file stresses one corner of kseq_read (src/kseq.h:185-224)."""
    rng = np.random.default_rng(17)

    def dna(n, alphabet="ACGT"):
        return "".join(rng.choice(list(alphabet), n))

    files = {}
    # multi-line FASTA: blank lines, CRLF, lower case, N / IUPAC, tabs and '>' inside the comment, junk before the first header
    s1, s2, s3 = dna(1000), dna(777, "ACGTacgtNnRYKM"), dna(130)
    files["a.fa"] = ("junk line before any header\n>r1 comment with > and @ inside\n" + "\n".join(s1[i:i + 60] for i in range(0, 1000, 60)) + "\n\n\n"
                     + ">r2\tTabbed comment\r\n" + "\r\n".join(s2[i:i + 70] for i in range(0, 777, 70)) + "\r\n"
                     + ">empty\n>r3\n" + s3)  # no trailing newline
    # FASTQ, single- and multi-line, quality lines that start with '@' and '>'
    q1 = dna(200)
    qual1 = "@" + "I" * 99 + "\n>" + "#" * 99
    files["b.fq"] = ("@q1 first\n" + q1[:100] + "\n" + q1[100:] + "\n+q1\n" + qual1 + "\n"
                     + "@q2\n" + dna(50) + "\n+\n" + "@" * 50 + "\n"
                     + ">mixed_fasta_after_fastq\n" + dna(90) + "\n")
    # gzip-compressed FASTA with long single-line reads
    files["c.fa.gz"] = "".join(f">g{i}/x/0_{n}\n{dna(n)}\n" for i, n in enumerate((15000, 1, 2, 33, 20011)))
    # truncated FASTQ: the second record's quality is short -> the reference stops reading this file there
    files["d.fq"] = "@ok\nACGTACGT\n+\nIIIIIIII\n@bad\nACGTACGTAC\n+\nIII\n@never\nAC\n+\nII\n"
    files["e.fa"] = ">after_truncated\n" + dna(300) + "\n>last lower\n" + dna(64, "acgtn") + "\n"
    lst = os.path.join(d, "in.lst")
    with open(lst, "w") as f:
        for name, text in files.items():
            p = os.path.join(d, name)
            if name.endswith(".gz"):
                with gzip.open(p, "wt") as g:
                    g.write(text)
            else:
                with open(p, "w", newline="") as g:
                    g.write(text)
            f.write(p + "\n")
    return lst


def ref_mkseqdb(ref_dir, lst, prefix):
    subprocess.check_call([os.path.join(ref_dir, "shmr_mkseqdb"), "-d", lst, "-p", prefix], stdout=subprocess.DEVNULL)
    return open(prefix + ".idx", "rb").read(), open(prefix + ".seqdb", "rb").read()


def test_fasta_scanner_on_host(tmp_path, ref_dir):
    subprocess.check_call(["make", "-C", ROOT, "hostsim"], stdout=subprocess.DEVNULL)
    lst = tricky_inputs(str(tmp_path))
    want_idx, want_db = ref_mkseqdb(ref_dir, lst, str(tmp_path / "ref"))
    got = subprocess.run([os.path.join(ROOT, "build", "hostsim"), "fastaidx", lst], stdout=subprocess.PIPE, check=True).stdout
    assert want_idx.count(b"\n") >= 12 and got == want_idx
    # the block-wise reader shmr_mkseqdb uses (GzRecordStream): every block size, down to blocks far smaller than a record, must
    # give the same records (a record that touches the end of the buffered text is scanned again once more text arrived)
    for block in (1, 3, 7, 64, 333, 1 << 20):
        got = subprocess.run([os.path.join(ROOT, "build", "hostsim"), "fastaidx", lst, str(block)], stdout=subprocess.PIPE, check=True).stdout
        assert got == want_idx, block


@pytest.mark.gpu
def test_mkseqdb_cli_matches_reference(tmp_path, ref_dir):
    lst = tricky_inputs(str(tmp_path))
    want_idx, want_db = ref_mkseqdb(ref_dir, lst, str(tmp_path / "ref"))
    out = subprocess.run([os.path.join(ROOT, "bin", "shmr_mkseqdb"), "-d", lst, "-p", str(tmp_path / "our")], stdout=subprocess.PIPE, check=True).stdout
    ref_out = subprocess.run([os.path.join(ref_dir, "shmr_mkseqdb"), "-d", lst, "-p", str(tmp_path / "ref2")], stdout=subprocess.PIPE, check=True).stdout
    assert out.replace(b"our", b"XXX") == ref_out.replace(b"ref2", b"XXX")  # same messages
    assert open(tmp_path / "our.idx", "rb").read() == want_idx
    assert open(tmp_path / "our.seqdb", "rb").read() == want_db
    # missing list file: message + exit 1, like the reference
    r = subprocess.run([os.path.join(ROOT, "bin", "shmr_mkseqdb"), "-d", str(tmp_path / "nope.lst"), "-p", str(tmp_path / "x")], capture_output=True)
    assert r.returncode == 1 and b"open error" in r.stderr


@pytest.mark.gpu
def test_encode_biseq_every_byte_value(tmp_path, ref_dir):
    """All 256 byte values on both strands, ragged lengths incl. 0 and 1 and a read spanning several kernel tiles, against
    the reference tool run on the same sequences (bytes that kseq cannot deliver are checked against its tables' rule)."""
    from peregrine_b200 import Engine

    rng = np.random.default_rng(3)
    lens = np.array([0, 1, 2, 31, 32, 33, 8191, 8192, 8193, 50000, 7], dtype=np.uint32)
    off = np.concatenate([[0], np.cumsum(lens.astype(np.uint64))[:-1]]).astype(np.uint64)
    total = int(lens.sum())
    printable = np.frombuffer(b"ACGTacgtNnRYKMSWBDHVU*-.", dtype=np.uint8)
    a = printable[rng.integers(0, len(printable), total)]
    eng = Engine(0)
    got = eng.encode_biseq(a, off, lens)
    fa = tmp_path / "e.fa"
    with open(fa, "wb") as f:
        for i, (o, l) in enumerate(zip(off, lens)):
            f.write(b">r%d\n" % i + a[int(o): int(o) + int(l)].tobytes() + b"\n")
    (tmp_path / "e.lst").write_text(str(fa) + "\n")
    _, want = ref_mkseqdb(ref_dir, str(tmp_path / "e.lst"), str(tmp_path / "e"))
    assert got.tobytes() == want
    # every byte value: only A C G T a c g t carry a code
    allb = np.arange(256, dtype=np.uint8)
    enc = eng.encode_biseq(allb, np.array([0], dtype=np.uint64), np.array([256], dtype=np.uint32))
    f = np.zeros(256, dtype=np.uint8)
    r = np.zeros(256, dtype=np.uint8)
    for ch, fv, rv in ((b"A", 1, 8), (b"C", 2, 4), (b"G", 4, 2), (b"T", 8, 1)):
        for c in (ch, ch.lower()):
            f[c[0]] = fv
            r[c[0]] = rv
    assert np.array_equal(enc, f[allb] | (r[allb[::-1]] << 4))
    assert eng.stats()["n_k_encode"] == 2
    eng.close()
