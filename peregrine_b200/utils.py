"""SHIMMER helpers of the reference's ``peregrine.utils`` (py/peregrine/utils.py:10-73) on top of libpgb200.so.

Same names, arguments and return values as the reference functions that sit on the index/overlap path — ``rc``,
``mmer2tuple``, ``get_shimmers_from_seq``, ``get_shimmer_alns`` — so notebooks and scripts written against
``peregrine.utils`` keep working with ``import peregrine_b200.utils as utils``.  The consensus helpers of that module
(``get_tag_from_seqs``, ``get_cns_from_reads``) sit on ``_falcon4py`` and are out of scope.

The helpers only talk to a cffi ``(ffi, lib)`` pair that exposes the reference cdef, so the same code can be pointed at the
reference's own library (``ShimmerTools.for_library(path)``): the tests run both and compare.
"""
import numpy as np

_COMPLEMENT = bytes.maketrans(b"ACGT", b"TGCA")


def rc(seq: bytes) -> bytes:
    """Reverse complement of an upper-case ACGT byte string (py/peregrine/utils.py:13-14)."""
    return seq.translate(_COMPLEMENT)[::-1]


def mmer2tuple(mmer):
    """(minimizer hash, span, read id, end position, strand) of an mm128_t (py/peregrine/utils.py:17-25)."""
    x, y = int(mmer.x), int(mmer.y)
    return (x >> 8, x & 0xFF, y >> 32, ((y & 0xFFFFFFFF) >> 1) + 1, y & 0x1)


class ShimmerTools:
    def __init__(self, ffi, lib):
        self.ffi, self.lib = ffi, lib

    @classmethod
    def for_library(cls, path):
        """Bind the same cdef to another shared library that exports the reference ABI (e.g. oracle/_ref/libshimmer_ref.so)."""
        from .shimmer4py import ffi

        return cls(ffi, ffi.dlopen(path))

    def get_shimmers_from_seq(self, seq, rid=0, levels=2, reduction_factor=3, k=16, w=80):
        """mm128_v* of level `levels` (0 = plain (w,k) minimizers) for one sequence; the caller owns it
        (``lib.free(v.a)``, ``ffi.release(v)``), as in py/peregrine/utils.py:28-49."""
        assert 0 <= levels <= 2
        ffi, lib = self.ffi, self.lib
        cur = ffi.new("mm128_v *")
        lib.mm_sketch(ffi.NULL, seq, len(seq), w, k, rid, 0, cur)
        for _ in range(levels):
            nxt = ffi.new("mm128_v *")
            lib.mm_reduce(cur, nxt, reduction_factor)
            if cur.a != ffi.NULL:
                lib.free(cur.a)  # (the reference releases only the cffi struct and leaks the array)
            ffi.release(cur)
            cur = nxt
        return cur

    def get_shimmer_alns(self, shimmers0, shimmers1, direction=0, max_diff=100, max_dist=1200, max_repeat=1):
        """Chains of shared minimizers: list of (chain, max, mean, min) with chain = [(mmer2tuple(m0), mmer2tuple(m1)), ...].
        The three statistics are those of the LAST pair's offset, as in the reference (py/peregrine/utils.py:69-71 reduces the
        scalar `d`, not the `offsets` array); the per-pair offsets are end0 - end1 (same direction) or end0 + end1."""
        lib = self.lib
        aln = lib.shmr_aln(shimmers0, shimmers1, direction, max_diff, max_dist, max_repeat)
        chains = []
        for i in range(aln.n):
            a = aln.a[i]
            chain, d = [], 0
            for j in range(a.idx0.n):
                m0 = mmer2tuple(shimmers0.a[a.idx0.a[j]])
                m1 = mmer2tuple(shimmers1.a[a.idx1.a[j]])
                chain.append((m0, m1))
                d = m0[3] - m1[3] if direction == 0 else m0[3] + m1[3]
            chains.append((chain, np.max(d), np.mean(d), np.min(d)))
        lib.free_shmr_alns(aln)
        return chains


_default = None


def _tools():
    global _default
    if _default is None:
        from .shimmer4py import ffi, lib

        _default = ShimmerTools(ffi, lib)
    return _default


def get_shimmers_from_seq(seq, rid=0, levels=2, reduction_factor=3, k=16, w=80):
    return _tools().get_shimmers_from_seq(seq, rid, levels, reduction_factor, k, w)


def get_shimmer_alns(shimmers0, shimmers1, direction=0, max_diff=100, max_dist=1200, max_repeat=1):
    return _tools().get_shimmer_alns(shimmers0, shimmers1, direction, max_diff, max_dist, max_repeat)
