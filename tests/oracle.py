"""ctypes wrapper of oracle/_build/liboracle.so (the plain-C restatement) and of oracle/_ref/libshimmer_ref.so (the
unmodified reference).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from peregrine_b200 import formats as F  # noqa: E402


class MMV(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.c_void_p)]


def _mmv_to_np(v, free):
    out = np.empty(v.n, dtype=F.MM128)
    if v.n:
        C.memmove(out.ctypes.data, v.a, v.n * 16)
    if v.a:
        free(v.a)
    return out


_orc = None
FASTA_CB = C.CFUNCTYPE(None, C.POINTER(C.c_char), C.c_size_t, C.POINTER(C.c_char), C.c_size_t, C.c_void_p)


def oracle():
    global _orc
    if _orc is None:
        p = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
        if not os.path.exists(p):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"], stdout=subprocess.DEVNULL)
        L = C.CDLL(p)
        L.orc_sketch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(MMV)]
        L.orc_reduce.argtypes = [C.POINTER(MMV), C.POINTER(MMV), C.c_int]
        L.orc_ovlp_match.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_index_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.POINTER(MMV * 3)]
        L.orc_count.restype = C.c_void_p
        L.orc_count.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_overlap_chunk.restype = C.c_void_p
        L.orc_overlap_chunk.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                        C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                                        C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        L.orc_free.argtypes = [C.c_void_p]
        # stages next to the path (oracle/stages_oracle.c)
        L.orc_encode_biseq.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_fasta_records.restype = C.c_size_t
        L.orc_fasta_records.argtypes = [C.c_char_p, C.c_size_t, FASTA_CB, C.c_void_p]
        L.orc_dedup.restype = C.c_void_p
        L.orc_dedup.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_map.restype = C.c_void_p
        L.orc_map.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_uint32,
                              C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
        _orc = L
    return _orc


def ascii_to_nib(seq: str) -> np.ndarray:
    lut = np.zeros(256, dtype=np.uint8)
    for ch, v in (("A", 1), ("C", 2), ("G", 4), ("T", 8), ("a", 1), ("c", 2), ("g", 4), ("t", 8)):
        lut[ord(ch)] = v
    return lut[np.frombuffer(seq.encode(), dtype=np.uint8)]


def encode_biseq(seq: str) -> np.ndarray:
    """.seqdb bytes of one read (src/shmr_utils.c:44-51)."""
    f = ascii_to_nib(seq)
    rmap = np.zeros(16, dtype=np.uint8)
    rmap[1], rmap[2], rmap[4], rmap[8] = 8, 4, 2, 1
    return (f | (rmap[f[::-1]] << 4)).astype(np.uint8)


def orc_sketch(nib, w, k, rid):
    L = oracle()
    nib = np.ascontiguousarray(nib, dtype=np.uint8)
    v = MMV(0, 0, None)
    L.orc_sketch(nib.ctypes.data, len(nib), w, k, rid, C.byref(v))
    return _mmv_to_np(v, L.orc_free)


def orc_reduce(mm, rs):
    L = oracle()
    mm = np.ascontiguousarray(mm, dtype=F.MM128)
    vin = MMV(len(mm), len(mm), mm.ctypes.data)
    v = MMV(0, 0, None)
    L.orc_reduce(C.byref(vin), C.byref(v), rs)
    return _mmv_to_np(v, L.orc_free)


def orc_ovlp_match(q, qs, t, ts, bw):
    L = oracle()
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    out = np.zeros(8, dtype=np.int32)
    L.orc_ovlp_match(q.ctypes.data, len(q), qs, t.ctypes.data, len(t), ts, bw, out.ctypes.data)
    return out


def orc_index_chunk(seqdb, rid, ln, off, T, c, w, k, r, levels):
    L = oracle()
    out = (MMV * 3)()
    L.orc_index_chunk(seqdb.ctypes.data, rid.ctypes.data, ln.ctypes.data, off.ctypes.data, len(rid), T, c, w, k, r, levels, C.byref(out))
    return [_mmv_to_np(out[i], L.orc_free) for i in range(3)]


def orc_count(mm):
    L = oracle()
    mm = np.ascontiguousarray(mm, dtype=F.MM128)
    n = C.c_size_t(0)
    p = L.orc_count(mm.ctypes.data, len(mm), C.byref(n))
    out = np.empty(n.value, dtype=F.MMCOUNT)
    if n.value:
        C.memmove(out.ctypes.data, p, n.value * 16)
    L.orc_free(p)
    return out


def orc_overlap_chunk(seqdb, rid, ln, off, mm, mc, T=1, c=1, bestn=4, lo=2, hi=240, bw=100, upper=120):
    L = oracle()
    mm = np.ascontiguousarray(mm, dtype=F.MM128)
    mc = np.ascontiguousarray(mc, dtype=F.MMCOUNT)
    n = C.c_size_t(0)
    na = C.c_uint64(0)
    p = L.orc_overlap_chunk(seqdb.ctypes.data, rid.ctypes.data, ln.ctypes.data, off.ctypes.data, len(rid), mm.ctypes.data, len(mm),
                            mc.ctypes.data, len(mc), T, c, bestn, lo, hi, bw, upper, C.byref(n), C.byref(na))
    out = np.empty(n.value, dtype=F.OVLP)
    if n.value:
        C.memmove(out.ctypes.data, p, n.value * 64)
    L.orc_free(p)
    return out, na.value


# ------------------------------------------------------------------------------------------------ the real reference
_ref = None


class MatchT(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("m_size", "dist", "q_bgn", "q_end", "t_bgn", "t_end", "t_m_end", "q_m_end")]


def reflib():
    """libshimmer_ref.so = the reference's own cffi source list (py/peregrine/build_shimmer4py.py:86-96) as a shared library."""
    global _ref
    if _ref is None:
        p = os.path.join(ROOT, "oracle", "_ref", "libshimmer_ref.so")
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.mm_sketch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(MMV)]
        L.mm_reduce.argtypes = [C.POINTER(MMV), C.POINTER(MMV), C.c_uint8]
        L.ovlp_match.restype = C.POINTER(MatchT)
        L.ovlp_match.argtypes = [C.c_void_p, C.c_int32, C.c_uint8, C.c_void_p, C.c_int32, C.c_uint8, C.c_int32]
        L.free_ovlp_match.argtypes = [C.POINTER(MatchT)]
        _ref = L
    return _ref


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


def abi_sketch(L, seq: str, w, k, rid):
    """Call <lib>.mm_sketch through the reference ABI (works for libshimmer_ref.so and libpgb200.so alike)."""
    v = MMV(0, 0, None)
    L.mm_sketch(None, seq.encode(), len(seq), w, k, rid, 0, C.byref(v))
    return _mmv_to_np(v, _libc.free)


def abi_reduce(L, mm, rs):
    mm = np.ascontiguousarray(mm, dtype=F.MM128)
    vin = MMV(len(mm), len(mm), mm.ctypes.data)
    v = MMV(0, 0, None)
    L.mm_reduce(C.byref(vin), C.byref(v), rs)
    return _mmv_to_np(v, _libc.free)


def abi_ovlp_match(L, q, qs, t, ts, bw):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    m = L.ovlp_match(q.ctypes.data, len(q), qs, t.ctypes.data, len(t), ts, bw)
    out = np.array([getattr(m.contents, n) for n, _ in MatchT._fields_], dtype=np.int32)
    L.free_ovlp_match(m)
    return out


# ------------------------------------------------------------------------------------------------ stages next to the path
def orc_mkseqdb(paths):
    """What shmr_mkseqdb writes for the listed files (plain or gzip): (.idx text, .seqdb bytes), via the oracle's kseq
    restatement and encode_biseq."""
    import gzip

    L = oracle()
    idx, db = [], []
    state = {"rid": 0, "off": 0}

    def rec(name, nl, seq, sl, _):
        nm = C.string_at(name, nl)
        sq = C.string_at(seq, sl)
        out = np.empty(sl, dtype=np.uint8)
        L.orc_encode_biseq(sq, sl, out.ctypes.data_as(C.c_void_p))
        idx.append(b"%09d %s %u %lu\n".replace(b"%lu", b"%d") % (state["rid"], nm, sl, state["off"]))
        db.append(out.tobytes())
        state["rid"] += 1
        state["off"] += sl

    cb = FASTA_CB(rec)
    for p in paths:
        raw = open(p, "rb").read()
        if raw[:2] == b"\x1f\x8b":
            raw = gzip.decompress(raw)
        L.orc_fasta_records(raw, len(raw), cb, None)
    return b"".join(idx), b"".join(db)


def orc_dedup(stream: np.ndarray) -> bytes:
    L = oracle()
    a = np.ascontiguousarray(stream, dtype=F.OVLP)
    n = C.c_size_t()
    p = L.orc_dedup(a.ctypes.data_as(C.c_void_p), len(a), C.byref(n))
    out = C.string_at(p, n.value)
    L.orc_free(p)
    return out


def orc_map(ref_mm, mm, mc, rid, ln, T=1, c=1, lower=1, upper=240) -> bytes:
    L = oracle()
    ref_mm = np.ascontiguousarray(ref_mm, dtype=F.MM128)
    mm = np.ascontiguousarray(mm, dtype=F.MM128)
    mc = np.ascontiguousarray(mc, dtype=F.MMCOUNT)
    by_rid = np.zeros(int(rid.max()) + 1, dtype=np.uint32)
    by_rid[rid] = ln
    n = C.c_size_t()
    p = L.orc_map(ref_mm.ctypes.data_as(C.c_void_p), len(ref_mm), mm.ctypes.data_as(C.c_void_p), len(mm), mc.ctypes.data_as(C.c_void_p), len(mc),
                  by_rid.ctypes.data_as(C.c_void_p), T, c, lower, upper, C.byref(n))
    out = C.string_at(p, n.value)
    L.orc_free(p)
    return out
