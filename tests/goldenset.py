"""Load tests/golden (committed golden vectors + their input reads) without needing /root/reference."""
import hashlib
import json
import os

import numpy as np

import oracle as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(b):
    return hashlib.sha256(b).hexdigest()


def load():
    gold = json.load(open(os.path.join(HERE, "golden.json")))
    reads = [l.strip() for l in open(os.path.join(HERE, "reads.fa")) if not l.startswith(">")]
    return gold, reads


def seqdb_arrays(reads):
    """What shmr_mkseqdb writes for reads.fa (src/shmr_mkseqdb.c:99-121): image + (rid, len, offset)."""
    enc = [O.encode_biseq(s) for s in reads]
    ln = np.array([len(e) for e in enc], dtype=np.uint32)
    off = np.concatenate([[0], np.cumsum(ln[:-1], dtype=np.uint64)]).astype(np.uint64)
    rid = np.arange(len(reads), dtype=np.uint32)
    return np.concatenate(enc), rid, ln, off


def write_seqdb(prefix, reads):
    seqdb, rid, ln, off = seqdb_arrays(reads)
    seqdb.tofile(prefix + ".seqdb")
    with open(prefix + ".idx", "w") as f:
        for i in range(len(reads)):
            f.write("%09d g/%06d/0_%d %u %lu\n" % (i, i, ln[i], ln[i], off[i]))
    return seqdb, rid, ln, off
