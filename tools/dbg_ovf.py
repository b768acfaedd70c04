import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import datasets as D
wd = tempfile.mkdtemp()
ref = os.path.join(ROOT, "oracle", "_ref")
p = D.make_sim(wd, "ovf", genome=300_000, cov=20)
rp = D.ref_index(ref, p, os.path.join(wd, "ovf/ref"), T=1, extra=["-m", "0"])
env = dict(os.environ, PGB_TABLE_SCALE=sys.argv[1] if len(sys.argv) > 1 else "0.02", PGB_VERBOSE="1")
r = subprocess.run([os.path.join(ROOT, "bin", "shmr_overlap"), "-p", p, "-l", rp + "-L2", "-t", "4", "-c", "1", "-o", os.path.join(wd, "o.dat")], env=env, capture_output=True, text=True)
print("rc", r.returncode)
lines = r.stderr.splitlines()
print("\n".join(l[:230] for l in lines if "replay pass" not in l))
print("\n".join(l[:200] for l in lines if "replay pass" in l)[-1500:])
