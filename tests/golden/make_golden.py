"""Generate the golden vectors in this directory from the UNMODIFIED reference (oracle/_ref, built from
/root/reference/src by oracle/Makefile).  Run where /root/reference is mounted:

    python tests/golden/make_golden.py

Outputs (committed):
  reads.fa        deterministic read set (own LCG; repeats, N runs, palindromes, short reads, noisy overlapping copies)
  golden.json     sha256 of every reference output file for that set, the per-call vectors for mm_sketch / mm_reduce /
                  ovlp_match, and the parameters used
  ovlp_T1.bin     the reference's full ovlp_t stream (T=1), padding bytes zeroed
The reference repository itself ships no golden vectors; these pin our oracle and the CUDA path to the
reference's behaviour on machines where /root/reference does not exist.
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

import oracle as O  # noqa: E402
from peregrine_b200 import formats as F  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


class LCG:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return self.s >> 33

    def below(self, n):
        return self.next() % n


def make_reads():
    g = LCG(20261017)
    B = "ACGT"
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rs = lambda n: "".join(B[g.below(4)] for _ in range(n))  # noqa: E731
    rc = lambda s: "".join(comp[c] for c in reversed(s))  # noqa: E731
    genome = rs(24000)
    pal = rs(8)
    genome = (genome[:3000] + "CA" * 150 + genome[3300:7000] + "A" * 200 + genome[7200:11000] + pal + rc(pal) + genome[11016:15000]
              + rs(41) * 8 + genome[15328:])
    recs = []
    for i in range(70):
        ln = 1500 + g.below(2500)
        s0 = g.below(len(genome) - ln)
        s = list(genome[s0:s0 + ln])
        out = []
        for c in s:
            if g.below(1000) < 8:
                m = g.below(9)
                c = (B[m] if m < 4 else "" if m == 4 else c + B[m - 5])
            out.append(c)
        s = "".join(out)
        if g.below(2):
            s = rc(s)
        if g.below(8) == 0:
            p = g.below(len(s) - 40)
            s = s[:p] + "N" * (1 + g.below(12)) + s[p + 12:]
        recs.append(s)
    recs += ["A", "ACGT", rs(15), rs(16), rs(17), rs(95), rs(96), "N" * 30, "A" * 300, "AC" * 200, rs(20) + "N" + rs(20), (pal + rc(pal)) * 12]
    return recs


def sha(b):
    return hashlib.sha256(b).hexdigest()


def main():
    reads = make_reads()
    with open(os.path.join(HERE, "reads.fa"), "w") as f:
        for i, s in enumerate(reads):
            f.write(f">g/{i:06d}/0_{len(s)}\n{s}\n")
    wd = tempfile.mkdtemp()
    run = lambda cmd: subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)  # noqa: E731
    with open(os.path.join(wd, "fa.lst"), "w") as f:
        f.write(os.path.join(HERE, "reads.fa") + "\n")
    run([os.path.join(REF, "shmr_mkseqdb"), "-d", os.path.join(wd, "fa.lst"), "-p", os.path.join(wd, "seq")])
    gold = {"params": {"w": 80, "k": 16, "r": 6, "l": 2, "T_idx": 2}, "files": {}, "seqdb_sha256": sha(open(os.path.join(wd, "seq.seqdb"), "rb").read())}
    for c in (1, 2):
        run([os.path.join(REF, "shmr_index"), "-p", os.path.join(wd, "seq"), "-t", "2", "-c", str(c), "-o", os.path.join(wd, "shmr"), "-m", "1"])
    for c in (1, 2):
        for lv in ("L0", "L2"):
            fn = f"shmr-{lv}-{c:02d}-of-02.dat"
            gold["files"][fn] = sha(open(os.path.join(wd, fn), "rb").read())
            fn = f"shmr-{lv}-MC-{c:02d}-of-02.dat"
            gold["files"][fn] = sha(F.mc_as_sorted_pairs(F.read_mc(os.path.join(wd, fn))).tobytes())
    run([os.path.join(REF, "shmr_overlap"), "-p", os.path.join(wd, "seq"), "-l", os.path.join(wd, "shmr-L2"), "-t", "1", "-c", "1", "-o", os.path.join(wd, "ovlp.T1")])
    ov = F.normalise_ovlp(F.read_ovlp(os.path.join(wd, "ovlp.T1")))
    ov.tofile(os.path.join(HERE, "ovlp_T1.bin"))
    gold["ovlp_T1"] = {"records": int(len(ov)), "sha256": sha(ov.tobytes())}
    for c in (1, 2):
        run([os.path.join(REF, "shmr_overlap"), "-p", os.path.join(wd, "seq"), "-l", os.path.join(wd, "shmr-L2"), "-t", "2", "-c", str(c), "-o", os.path.join(wd, f"ovlp.T2.{c}")])
        o2 = F.normalise_ovlp(F.read_ovlp(os.path.join(wd, f"ovlp.T2.{c}")))
        gold[f"ovlp_T2_c{c}"] = {"records": int(len(o2)), "sha256": sha(o2.tobytes())}
    # per-call vectors through the reference's cffi ABI
    L = O.reflib()
    sk = []
    for (w, k) in ((80, 16), (24, 12), (60, 14), (120, 18), (255, 28), (30, 17)):
        h = hashlib.sha256()
        n = 0
        for i, s in enumerate(reads):
            a = O.abi_sketch(L, s, w, k, i)
            h.update(a.tobytes())
            n += len(a)
        sk.append({"w": w, "k": k, "n": n, "sha256": h.hexdigest()})
    gold["mm_sketch"] = sk
    l0 = np.concatenate([O.abi_sketch(L, s, 80, 16, i) for i, s in enumerate(reads)])
    rd = []
    for r in (1, 2, 3, 6, 24, 36):
        a = O.abi_reduce(L, l0, r)
        b = O.abi_reduce(L, a, r)
        rd.append({"r": r, "n1": int(len(a)), "n2": int(len(b)), "sha256_1": sha(a.tobytes()), "sha256_2": sha(b.tobytes())})
    gold["mm_reduce"] = rd
    # ovlp_match: true candidate pairs (from the ovlp stream) and unrelated pairs, all strand combinations, three band widths
    enc = [O.encode_biseq(s) for s in reads]
    g = LCG(99)
    om = []
    for t in range(96):
        if t < 48 and len(ov):
            r = ov[g.below(len(ov))]
            i, j = int(r["y0"] >> 32), int(r["y1"] >> 32)
            p0, p1 = ((int(r["y0"]) & 0xFFFFFFFF) >> 1) + 1, ((int(r["y1"]) & 0xFFFFFFFF) >> 1) + 1
            st, s0, s1 = p0 - p1, int(r["strand0"]), int(r["strand1"])
        else:
            i, j = g.below(70), g.below(70)
            st, s0, s1 = g.below(len(reads[i]) // 2), g.below(2), g.below(2)
        bw = (50, 100, 200)[t % 3]
        res = O.abi_ovlp_match(L, enc[i][st:], s0, enc[j], s1, bw)
        om.append({"i": i, "start": st, "s0": s0, "j": j, "s1": s1, "bw": bw, "match": [int(x) for x in res]})
    gold["ovlp_match"] = om
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    shutil.rmtree(wd)
    print("golden vectors written:", {k: (v if not isinstance(v, (list, dict)) else "...") for k, v in gold.items()})


if __name__ == "__main__":
    main()
