"""Pins tests/ecoli_standin.py: runs the reference's OWN simulator (test/ecoli_K12/simulate_reads.py, unmodified, from
/root/reference) on the stand-in genome and records the SHA-256 of its eight FASTA files next to the SHA-256 of what
ecoli_standin.simulate writes for the same genome.  Run in the build container (needs /root/reference):

    python tests/golden/make_ecoli_standin.py
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ecoli_standin as E  # noqa: E402

REF_SCRIPT = "/root/reference/test/ecoli_K12/simulate_reads.py"

with tempfile.TemporaryDirectory() as wd:
    seq = E.make_genome(os.path.join(wd, "K12MG1655.fa"))
    os.makedirs(os.path.join(wd, "reads"))
    subprocess.run([sys.executable, REF_SCRIPT], cwd=wd, check=True)  # reads ./K12MG1655.fa, writes reads/reads_{j}.fa
    ref_paths = [os.path.join(wd, "reads", f"reads_{j}.fa") for j in range(E.N_FILES)]
    ref_sha = E.sha256_files(ref_paths)
    ours = E.simulate(seq, os.path.join(wd, "ours"))
    our_sha = E.sha256_files(ours)
    n_reads = sum(1 for p in ref_paths for line in open(p) if line.startswith(">"))
    n_bases = sum(len(line) - 1 for p in ref_paths for line in open(p) if not line.startswith(">"))
out = {"genome_bp": E.GENOME_BP, "reads": n_reads, "bases": n_bases, "sha256_reference_script": ref_sha, "sha256_restatement": our_sha,
       "identical": ref_sha == our_sha}
json.dump(out, open(os.path.join(HERE, "ecoli_standin.json"), "w"), indent=1)
print(out)
assert ref_sha == our_sha
