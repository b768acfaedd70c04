"""Golden vectors of the stages next to the index/overlap path: shmr_mkseqdb, shmr_dedup, shmr_map.
Generated from the UNMODIFIED reference (oracle/_ref).  Run where /root/reference is mounted:

    python tests/golden/make_golden_stages.py

Outputs (committed):
  tricky.fa / tricky.fq        inputs that exercise kseq's grammar (multi-line, CRLF, blank lines, '>' '@' in odd places, FASTQ)
  dedup_in.bin                 an ovlp_t stream: the golden overlap stream twice (second copy with the read roles swapped) plus
                               synthetic records that hit every branch / wrap of shmr_dedup's arithmetic
  dedup_expected.txt           what the reference's shmr_dedup prints for it
  map_expected.txt             shmr_map of the first 6 reads of reads.fa ("contigs") against the index of all reads
  golden_stages.json           sha256 of the reference's .idx / .seqdb for (reads.fa) and (tricky.fa, tricky.fq), sizes, parameters
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from peregrine_b200 import formats as F  # noqa: E402
from test_dedup import adversarial_stream  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
N_CTG = 6


def sha(b):
    return hashlib.sha256(b).hexdigest()


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def tricky_files():
    rng = np.random.default_rng(5)
    dna = lambda n, a="ACGT": "".join(rng.choice(list(a), n))  # noqa: E731
    s1, s2 = dna(333), dna(250, "ACGTacgtNnRY")
    fa = ("leading junk\n>t1 a comment > with @ signs\n" + "\n".join(s1[i:i + 50] for i in range(0, 333, 50)) + "\n\n"
          + ">t2\ttab comment\r\n" + "\r\n".join(s2[i:i + 61] for i in range(0, 250, 61)) + "\r\n>empty\n>t3\n" + dna(77))
    q = dna(120)
    fq = ("@q1 x\n" + q[:60] + "\n" + q[60:] + "\n+\n@" + "I" * 59 + "\n>" + "#" * 59 + "\n@q2\n" + dna(40) + "\n+q2\n" + "@" * 40 + "\n"
          + ">fa_after_fq\n" + dna(55) + "\n@bad\nACGTACGT\n+\nIII\n@never\nAC\n+\nII\n")
    return fa, fq


def main():
    wd = tempfile.mkdtemp()
    gold = {}
    # ---- mkseqdb
    fa, fq = tricky_files()
    for name, text in (("tricky.fa", fa), ("tricky.fq", fq)):
        with open(os.path.join(HERE, name), "w", newline="") as f:
            f.write(text)
    for tag, files in (("reads", ["reads.fa"]), ("tricky", ["tricky.fa", "tricky.fq"])):
        lst = os.path.join(wd, tag + ".lst")
        with open(lst, "w") as f:
            for n in files:
                f.write(os.path.join(HERE, n) + "\n")
        run([os.path.join(REF, "shmr_mkseqdb"), "-d", lst, "-p", os.path.join(wd, tag)])
        idx = open(os.path.join(wd, tag + ".idx"), "rb").read()
        db = open(os.path.join(wd, tag + ".seqdb"), "rb").read()
        gold["mkseqdb_" + tag] = {"files": files, "idx_sha256": sha(idx), "seqdb_sha256": sha(db), "n_records": idx.count(b"\n"), "bases": len(db),
                                  "idx_text": idx.decode() if tag == "tricky" else None}
    # ---- dedup
    ov = np.fromfile(os.path.join(HERE, "ovlp_T1.bin"), dtype=F.OVLP)
    sw = ov.copy()
    sw["y0"], sw["y1"] = ov["y1"], ov["y0"]
    stream = np.concatenate([ov, sw, adversarial_stream(300, seed=11)])
    stream.tofile(os.path.join(HERE, "dedup_in.bin"))
    txt = run([os.path.join(REF, "shmr_dedup")], input=stream.tobytes()).stdout
    open(os.path.join(HERE, "dedup_expected.txt"), "wb").write(txt)
    gold["dedup"] = {"records_in": int(len(stream)), "lines": txt.count(b"\n"), "sha256": sha(txt)}
    # ---- map: first N_CTG reads as contigs
    recs, name = [], None
    for l in open(os.path.join(HERE, "reads.fa")):
        if l.startswith(">"):
            name = l[1:].strip()
        else:
            recs.append((name, l.strip()))
    with open(os.path.join(wd, "ctg.fa"), "w") as f:
        for n, s in recs[:N_CTG]:
            f.write(f">{n}\n{s}\n")
    with open(os.path.join(wd, "ctg.lst"), "w") as f:
        f.write(os.path.join(wd, "ctg.fa") + "\n")
    run([os.path.join(REF, "shmr_mkseqdb"), "-d", os.path.join(wd, "ctg.lst"), "-p", os.path.join(wd, "ctg")])
    for prefix, out in (("reads", "ridx"), ("ctg", "cidx")):
        run([os.path.join(REF, "shmr_index"), "-p", os.path.join(wd, prefix), "-t", "1", "-c", "1", "-o", os.path.join(wd, out), "-m", "0", "-r", "3"])
    m = run([os.path.join(REF, "shmr_map"), "-r", os.path.join(wd, "ctg"), "-m", os.path.join(wd, "cidx-L2"), "-p", os.path.join(wd, "reads"), "-l", os.path.join(wd, "ridx-L2")]).stdout
    open(os.path.join(HERE, "map_expected.txt"), "wb").write(m)
    gold["map"] = {"n_contigs": N_CTG, "index_args": ["-m", "0", "-r", "3"], "lines": m.count(b"\n"), "sha256": sha(m)}
    with open(os.path.join(HERE, "golden_stages.json"), "w") as f:
        json.dump(gold, f, indent=1)
    shutil.rmtree(wd)
    print({k: {kk: vv for kk, vv in v.items() if kk != "idx_text"} for k, v in gold.items()})


if __name__ == "__main__":
    main()
