"""Synthetic read sets + reference runs shared by the tests (cached per pytest session)."""
import os
import random
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIMREADS = os.path.join(ROOT, "build", "simreads")


def run(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, **kw)


def ensure_simreads():
    if not os.path.exists(SIMREADS):
        os.makedirs(os.path.dirname(SIMREADS), exist_ok=True)
        run(["gcc", "-O3", "-fopenmp", "-o", SIMREADS, os.path.join(ROOT, "tools", "simreads.c"), "-lm"])
    return SIMREADS


def make_sim(workdir, name, genome, cov=20, err=0.005, seed=42, mean=15000, sd=1500, mod=1, res=0):
    """mod/res: only the reads with rid % mod == res (one rank's share of a sharded job; rids stay global)."""
    d = os.path.join(workdir, name)
    prefix = os.path.join(d, "seq")
    if not os.path.exists(prefix + ".seqdb") or not os.path.exists(prefix + ".done"):
        os.makedirs(d, exist_ok=True)
        run([ensure_simreads(), "-g", str(genome), "-c", str(cov), "-e", str(err), "-S", str(seed), "-l", str(mean), "-s", str(sd),
             "-m", str(mod), "-r", str(res), "-p", prefix])
        open(prefix + ".done", "w").close()
    return prefix


def make_from_fasta(workdir, name, records, ref_dir):
    """records: list of (name, sequence).  Goes through the reference's own shmr_mkseqdb."""
    d = os.path.join(workdir, name)
    prefix = os.path.join(d, "seq")
    if not os.path.exists(prefix + ".seqdb"):
        os.makedirs(d, exist_ok=True)
        fa = os.path.join(d, "reads.fa")
        with open(fa, "w") as f:
            for n, s in records:
                f.write(f">{n}\n{s}\n")
        with open(os.path.join(d, "fa.lst"), "w") as f:
            f.write(fa + "\n")
        run([os.path.join(ref_dir, "shmr_mkseqdb"), "-d", os.path.join(d, "fa.lst"), "-p", prefix])
    return prefix


def adversarial_records(seed=7):
    """Reads that exercise the corner cases of SURVEY App. A-5/A-6: N runs, tandem repeats (hash ties), poly-A,
    palindromic k-mers, reads shorter than k / shorter than one window, lower case, plus overlapping noisy copies."""
    rnd = random.Random(seed)
    B = "ACGT"
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}

    def rs(n):
        return "".join(rnd.choice(B) for _ in range(n))

    def rc(s):
        return "".join(comp[c] for c in reversed(s))

    def noisy(s, p=0.01):
        out = []
        for c in s:
            if rnd.random() < p:
                c = rnd.choice(["A", "C", "G", "T", "", c + "A", c + "C", c + "G", c + "T"])
            out.append(c)
        return "".join(out)

    genome = rs(60000)
    # plant structure into the genome so that overlapping reads share it
    genome = genome[:5000] + "CA" * 300 + genome[5600:12000] + "A" * 400 + genome[12400:20000] + ("ACGTTGCA" * 50) + genome[20400:]
    pal = rs(8)
    genome = genome[:30000] + pal + rc(pal) + genome[30016:40000] + (rs(37) * 12) + genome[40444:]
    recs = []
    i = 0
    for _ in range(140):
        ln = int(rnd.gauss(6000, 1500))
        ln = max(700, ln)
        s0 = rnd.randint(0, len(genome) - ln)
        s = noisy(genome[s0:s0 + ln])
        if rnd.random() < 0.5:
            s = rc(s)
        if rnd.random() < 0.15:  # sprinkle N runs
            p = rnd.randint(0, len(s) - 40)
            s = s[:p] + "N" * rnd.randint(1, 30) + s[p + 30:]
        if rnd.random() < 0.1:
            s = s.lower()
        recs.append((f"adv/{i:06d}/0_{len(s)}", s))
        i += 1
    for s in ["A", "ACGT", rs(15), rs(16), rs(17), rs(94), rs(95), rs(96), rs(97), "N" * 50, "A" * 500, "AC" * 300,
              rs(20) + "N" + rs(20), "N" + rs(300), rs(300) + "N", (pal + rc(pal)) * 20, rs(200) + "NNNN" + rs(200)]:
        recs.append((f"adv/{i:06d}/0_{len(s)}", s))
        i += 1
    return recs


def ref_index(ref_dir, prefix, outdir, T=1, extra=()):
    os.makedirs(outdir, exist_ok=True)
    for c in range(1, T + 1):
        run([os.path.join(ref_dir, "shmr_index"), "-p", prefix, "-t", str(T), "-c", str(c), "-o", os.path.join(outdir, "shmr"), *extra])
    return os.path.join(outdir, "shmr")


def ref_overlap(ref_dir, prefix, idx_prefix, level, outdir, T=1, extra=()):
    os.makedirs(outdir, exist_ok=True)
    outs = []
    for c in range(1, T + 1):
        o = os.path.join(outdir, f"ovlp.{c:02d}")
        run([os.path.join(ref_dir, "shmr_overlap"), "-p", prefix, "-l", f"{idx_prefix}-L{level}", "-t", str(T), "-c", str(c), "-o", o, *extra])
        outs.append(o)
    return outs


def bad_strip_records(seed=41, n_reads=60):
    """Random reads with planted trouble for the strip sketch kernel's per-strip fallback (sketch_strip.cuh): tandem duplications
    (equal k-mer hashes inside one window: ties) in the middle of a read, in two adjacent strips, across a strip boundary, in the
    last partial strip and beyond strip 63 (> 32 kb: whole-read fallback); palindromic k-mers (no window slot) before the first
    full window, two within one window, and one next to a duplication."""
    rnd = random.Random(seed)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rs = lambda n: "".join(rnd.choice("ACGT") for _ in range(n))
    rc = lambda s: "".join(comp[c] for c in reversed(s))

    def plant_dup(s, at, unit):  # s[at:at+unit] twice in a row
        return s[:at + unit] + s[at:at + unit] + s[at + unit:]

    def plant_pal(s, at, half):  # a reverse-complement palindrome of 2*half bases at `at`
        h = rs(half)
        return s[:at] + h + rc(h) + s[at + 2 * half:]

    recs = []
    for i in range(n_reads):
        n = rnd.choice([700, 3000, 9000, 15000, 20000])
        s = rs(n)
        kind = i % 10
        if kind == 0 and n > 2500:
            s = plant_dup(s, 1700, 30)
        elif kind == 1 and n > 2500:
            s = plant_dup(plant_dup(s, 1000, 25), 1500, 40)          # adjacent strips 1/2 and 2/3
        elif kind == 2 and n > 2500:
            s = plant_dup(s, 2030, 28)                                # across the boundary at 2048
        elif kind == 3:
            s = plant_dup(s, len(s) - 120, 35)                        # last (partial) strip
        elif kind == 4:
            s = plant_pal(s, 20, rnd.choice([6, 7, 8, 9]))            # before the first full window
        elif kind == 5 and n > 2500:
            h = rnd.choice([6, 7, 8, 9])
            s = plant_pal(plant_pal(s, 1200, h), 1240, h)             # two in one window
        elif kind == 6 and n > 2500:
            s = plant_dup(plant_pal(s, 2500, 8), 2530, 30)
        elif kind == 7 and n > 2500:
            s = plant_pal(plant_pal(s, 600, 7), 2100, 9)              # k = 14 / k = 18 palindromes
        elif kind == 8:
            for at in range(100, len(s) - 200, 900):
                s = plant_dup(s, at, rnd.choice([17, 24, 33, 50]))
        recs.append((f"r/{i}/0_{len(s)}", s))
    long_read = rs(40000)
    recs.append(("r/long/0_1", plant_dup(plant_dup(long_read, 5000, 30), 35000, 30)))  # bad windows before and beyond strip 63
    return recs
