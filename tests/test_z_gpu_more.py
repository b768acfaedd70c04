"""GPU parity tests added after the round's last full GPU run (they sort last so that a surprise in one of them cannot hide
the established tests behind `pytest -x`): the full config-5 parameter grid and the overflow / restart path of the replay
tables.  Same helpers and checker (the unmodified reference in oracle/_ref) as tests/test_gpu_parity.py."""
import os

import pytest

import datasets as D
from test_gpu_parity import assert_same_ovlp, compare_index, ours_index, ours_overlap

pytestmark = pytest.mark.gpu


def test_config5_grid(workdir, ref_dir):
    """BASELINE.json configs[4]: every (k, w) of {14,16,18} x {60,80,120} through index AND overlap with aln_bw of 50/100/200,
    on a small noisy set (1 % error: wider bands than the clean sets)."""
    p = D.make_sim(workdir, "c5", genome=250_000, cov=15, err=0.01)
    bws = ("50", "100", "200")
    n = 0
    for k in (14, 16, 18):
        for w in (60, 80, 120):
            tag = f"k{k}w{w}"
            ex = ["-k", str(k), "-w", str(w), "-m", "0"]
            rp = D.ref_index(ref_dir, p, os.path.join(workdir, f"c5/ref_{tag}"), T=1, extra=ex)
            op = ours_index(p, os.path.join(workdir, f"c5/our_{tag}"), T=1, extra=ex)
            compare_index(rp, op, 1, levels=("L2",))
            bw = bws[n % 3]
            n += 1
            ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, f"c5/refo_{tag}"), T=1, extra=["-w", bw])
            oo = ours_overlap(p, op, 2, os.path.join(workdir, f"c5/ouro_{tag}"), T=1, extra=["-w", bw])
            assert_same_ovlp(oo[0], ro[0])


@pytest.mark.parametrize("scale", ["0.02", "0.3"])
def test_replay_tables_overflow_and_restart(workdir, ref_dir, monkeypatch, scale):
    """PGB_TABLE_SCALE shrinks the initial pair table / alignment cache so that they fill up: probes are bounded
    (PGB_MAX_PROBE), the pass is abandoned, the host doubles the table and restarts the fix-point, and the records are
    still the reference's.  (At T >= 4 chunks of a big job the default sizing overflows for real: rid_pairs is per chunk,
    so a read pair is aligned in several chunks.)"""
    monkeypatch.setenv("PGB_TABLE_SCALE", scale)
    monkeypatch.setenv("PGB_VERBOSE", "1")
    p = D.make_sim(workdir, "ovf", genome=300_000, cov=20)
    rp = D.ref_index(ref_dir, p, os.path.join(workdir, "ovf/ref"), T=1, extra=["-m", "0"])
    for T in (1, 4):
        ro = D.ref_overlap(ref_dir, p, rp, 2, os.path.join(workdir, f"ovf/ref{T}"), T=T)
        oo = ours_overlap(p, rp, 2, os.path.join(workdir, f"ovf/our{T}_{scale}"), T=T)
        for a, b in zip(oo, ro):
            assert_same_ovlp(a, b)
