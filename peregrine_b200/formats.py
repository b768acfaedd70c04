"""numpy views of the reference's on-disk formats (all little-endian native struct dumps).

=====================  =========================================  ==========================================
file                   producer in the reference                  layout
=====================  =========================================  ==========================================
<p>.idx                src/shmr_mkseqdb.c:112                     text "%09d %s %u %lu" rid name len offset
<p>.seqdb              src/shmr_mkseqdb.c:114                     1 byte/base, lo nibble fwd, hi nibble rc
*-L{0,1,2}-cc-of-TT    write_mmlist  src/shmr_utils.c:98-108      size_t n + n x {u64 x, u64 y}
*-MC-cc-of-TT          write_mm_count src/shmr_utils.c:178-188    size_t n + n x {u64 mer, u32 count, 4 pad}
ovlp.cc                fwrite(ovlp_t) src/shmr_overlap.c:173      headerless stream of 64-byte ovlp_t
=====================  =========================================  ==========================================
"""
from __future__ import annotations

import numpy as np

MM128 = np.dtype([("x", "<u8"), ("y", "<u8")])
MMCOUNT = np.dtype([("mer", "<u8"), ("count", "<u4"), ("pad", "<u4")])
MATCH = np.dtype([(n, "<i4") for n in ("m_size", "dist", "q_bgn", "q_end", "t_bgn", "t_end", "t_m_end", "q_m_end")])
OVLP = np.dtype(
    [("y0", "<u8"), ("y1", "<u8"), ("rl0", "<u4"), ("rl1", "<u4"), ("strand0", "u1"), ("strand1", "u1"),
     ("ovlp_type", "u1"), ("pad0", "u1"), ("match", MATCH), ("pad1", "<u4")]
)
assert MM128.itemsize == 16 and MMCOUNT.itemsize == 16 and MATCH.itemsize == 32 and OVLP.itemsize == 64


def read_idx(path):
    """-> (rid u32[n], len u32[n], offset u64[n]) in file order."""
    rid, ln, off = [], [], []
    with open(path) as f:
        for line in f:
            t = line.split()
            if len(t) < 4:
                continue
            rid.append(int(t[0]))
            ln.append(int(t[2]))
            off.append(int(t[3]))
    return np.asarray(rid, dtype=np.uint32), np.asarray(ln, dtype=np.uint32), np.asarray(off, dtype=np.uint64)


def _read_counted(path, dtype):
    with open(path, "rb") as f:
        n = int(np.frombuffer(f.read(8), dtype="<u8")[0])
        a = np.fromfile(f, dtype=dtype, count=n)
    if len(a) != n:
        raise IOError(f"{path}: header says {n} records, file holds {len(a)}")
    return a


def read_mmlist(path):
    return _read_counted(path, MM128)


def read_mc(path):
    return _read_counted(path, MMCOUNT)


def write_mmlist(path, a):
    a = np.ascontiguousarray(a, dtype=MM128)
    with open(path, "wb") as f:
        f.write(np.uint64(len(a)).tobytes())
        a.tofile(f)


def write_mc(path, a):
    a = np.ascontiguousarray(a, dtype=MMCOUNT)
    with open(path, "wb") as f:
        f.write(np.uint64(len(a)).tobytes())
        a.tofile(f)


def read_ovlp(path):
    return np.fromfile(path, dtype=OVLP)


def normalise_ovlp(a):
    """Zero the bytes the reference leaves uninitialised (struct padding at byte 27 and 60..63, SURVEY A-1)."""
    a = a.copy()
    a["pad0"] = 0
    a["pad1"] = 0
    return a


def mc_as_sorted_pairs(a):
    """(mer,count) set in canonical order — the reference writes khash slot order, only the set matters downstream."""
    o = np.argsort(a["mer"], kind="stable")
    return np.stack([a["mer"][o], a["count"][o].astype(np.uint64)], axis=1)
