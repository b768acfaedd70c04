"""ctypes binding of libpgb200.so's stage-level API (include/pgb200.h, group 3).

``Engine`` mirrors the two reference tools as method calls over host (numpy) buffers:

    eng = Engine(device=0)
    eng.load_reads(seqdb_u8, rid, length, offset)          # .seqdb image + .idx table  (src/shmr_index.c:133-163)
    eng.index(w=80, k=16, r=6, levels=2)                   # mm_sketch + mm_reduce x2   (src/shmr_index.c:155-216)
    l2 = eng.level(2); mc = eng.level_counts(2)
    eng.set_shimmers(l2_all_chunks, mc_all_chunks)         # what shmr_overlap globs    (src/shmr_overlap.c:359-382)
    ov = eng.overlap(total_chunk=1, mychunk=1)             # build_map + process_overlaps (src/shmr_overlap.c:394-397)

Nothing here computes: every call lands in a CUDA kernel.  Without the shared library or without a CUDA device the
constructor raises (there is deliberately no fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import formats

_HERE = os.path.dirname(os.path.abspath(__file__))


class NoDeviceError(RuntimeError):
    pass


def lib_path() -> str:
    return os.environ.get("PGB200_LIB", os.path.join(_HERE, "libpgb200.so"))


class _Stats(C.Structure):
    _fields_ = (
        [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]
        + [(n, C.c_double) for n in ("ms_pack", "ms_sketch", "ms_reduce", "ms_count", "ms_pairs", "ms_buckets",
                                     "ms_host_order", "ms_replay", "ms_align", "ms_emit")]
        + [(n, C.c_uint64) for n in ("bases_packed", "bases_sketched", "n_l0", "n_l1", "n_l2", "n_pair_records",
                                     "n_buckets", "n_eligible_buckets", "n_candidates", "n_alignments",
                                     "n_align_bases", "n_replay_passes", "n_overlaps")]
        + [(n, C.c_double) for n in ("ms_k_sketch_count", "ms_k_sketch_write", "ms_k_align", "ms_k_replay")]
        + [(n, C.c_uint64) for n in ("n_k_sketch_count", "n_k_sketch_write", "n_k_align", "n_k_replay")]
        + [("ms_k_sketch_tiled", C.c_double), ("n_k_sketch_tiled", C.c_uint64), ("n_sketch_fallback_reads", C.c_uint64),
           ("n_replay_buckets", C.c_uint64), ("ms_dedup", C.c_double), ("n_dedup_in", C.c_uint64), ("n_dedup_kept", C.c_uint64),
           ("ms_encode", C.c_double), ("ms_k_encode", C.c_double), ("n_k_encode", C.c_uint64), ("bases_encoded", C.c_uint64),
           ("ms_map", C.c_double), ("n_map_hits", C.c_uint64), ("n_device_mallocs", C.c_uint64), ("n_replay_restarts", C.c_uint64)]
    )


_lib = None


def load_library():
    """dlopen libpgb200.so and declare the prototypes.  Raises FileNotFoundError if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not found: build it with `make` (or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(p)
    vp, u8p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.pgb_device_count.restype = C.c_int
    L.pgb_create.restype = vp
    L.pgb_create.argtypes = [C.c_int]
    L.pgb_destroy.argtypes = [vp]
    L.pgb_last_error.restype = C.c_char_p
    L.pgb_last_error.argtypes = [vp]
    L.pgb_load_reads.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
    L.pgb_repack.argtypes = [vp]
    L.pgb_pack_2bit.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_size_t, vp, vp, vp]
    L.pgb_load_reads_2bit.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_int]
    L.pgb_load_reads_from_files.argtypes = [vp, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int]
    L.pgb_index.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pgb_index_size.restype = C.c_size_t
    L.pgb_index_size.argtypes = [vp, C.c_int]
    L.pgb_index_copy.argtypes = [vp, C.c_int, vp]
    L.pgb_index_count_size.restype = C.c_size_t
    L.pgb_index_count_size.argtypes = [vp, C.c_int]
    L.pgb_index_count_copy.argtypes = [vp, C.c_int, vp]
    L.pgb_set_shimmers.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t]
    L.pgb_set_shimmers_from_index.argtypes = [vp, C.c_int]
    L.pgb_overlap.argtypes = [vp] + [C.c_uint32] * 7
    L.pgb_overlap_size.restype = C.c_size_t
    L.pgb_overlap_size.argtypes = [vp]
    L.pgb_overlap_copy.argtypes = [vp, vp]
    L.pgb_overlap_host.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.pgb_buffer_elems.restype = C.c_size_t
    L.pgb_buffer_elems.argtypes = [vp, C.c_int]
    L.pgb_buffer_copy_out.argtypes = [vp, C.c_int, vp]
    L.pgb_load_packed_device.argtypes = [vp, vp, vp, C.c_size_t, vp, vp, vp, vp, C.c_size_t]
    L.pgb_set_shimmers_device.argtypes = [vp, vp, C.c_size_t]
    L.pgb_counts_dump.argtypes = [vp, C.POINTER(C.c_size_t)]
    L.pgb_counts_set_device.argtypes = [vp, vp, C.c_size_t]
    L.pgb_route_scan.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    L.pgb_route_build.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u64p]
    L.pgb_overlap_routed.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.pgb_ovlp_match_batch.argtypes = [vp, vp, C.c_size_t, C.c_size_t] + [vp] * 6 + [C.c_int, vp]
    L.pgb_shmr_aln_batch.argtypes = [vp, vp, vp, vp, vp, C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp, C.POINTER(vp)]
    L.pgb_host_free.argtypes = [vp]
    L.pgb_dedup.argtypes = [vp, vp, C.c_size_t]
    L.pgb_dedup_device.argtypes = [vp, vp, C.c_size_t]
    L.pgb_dedup_overlaps.argtypes = [vp]
    L.pgb_dedup_kept.restype = C.c_size_t
    L.pgb_dedup_kept.argtypes = [vp]
    L.pgb_dedup_text_bytes.restype = C.c_size_t
    L.pgb_dedup_text_bytes.argtypes = [vp]
    L.pgb_dedup_text_copy.argtypes = [vp, vp]
    L.pgb_shmr_dedup_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.pgb_shmr_mkseqdb_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.pgb_shmr_map_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.pgb_set_read_lengths.argtypes = [vp, vp, vp, C.c_size_t]
    L.pgb_map.argtypes = [vp, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.pgb_map_hits.restype = C.c_size_t
    L.pgb_map_hits.argtypes = [vp]
    L.pgb_map_text_bytes.restype = C.c_size_t
    L.pgb_map_text_bytes.argtypes = [vp]
    L.pgb_map_text_copy.argtypes = [vp, vp]
    L.pgb_encode_biseq.argtypes = [vp, vp, C.c_size_t, vp, vp, C.c_size_t, vp]
    L.pgb_stats_reset.argtypes = [vp]
    L.pgb_stats_get.argtypes = [vp, C.POINTER(_Stats)]
    L.pgb_event_record.argtypes = [vp, C.c_int]
    L.pgb_event_elapsed_ms.restype = C.c_double
    L.pgb_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int]
    L.pgb_shmr_index_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    L.pgb_shmr_overlap_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, device: int = 0):
        self.L = load_library()
        if self.L.pgb_device_count() <= 0:
            raise NoDeviceError("libpgb200 needs a CUDA device (sm_100a); none is visible and there is no CPU path")
        self.h = self.L.pgb_create(device)
        if not self.h:
            raise NoDeviceError(f"pgb_create({device}) failed")
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.pgb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed: {self.L.pgb_last_error(self.h).decode()}")

    # ------------------------------------------------------------------ reads
    def load_reads(self, seqdb, rid, length, offset, total_chunk=1, mychunk=1, keep_raw=False, defer=False):
        """seqdb: uint8 array (numpy, may wrap pinned memory); rid/length/offset as parsed from .idx.
        defer=True (PGB_LOAD_DEFER): the bulk copy is overlapped with the next index() call; seqdb must stay alive and
        unchanged until that call returns."""
        self._seqdb_ref = seqdb if defer else None
        seqdb = np.ascontiguousarray(seqdb, dtype=np.uint8) if not isinstance(seqdb, np.ndarray) else seqdb
        rid = np.ascontiguousarray(rid, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        self._ck(self.L.pgb_load_reads(self.h, _ptr(seqdb), seqdb.size, _ptr(rid), _ptr(length), _ptr(offset), len(rid),
                                       total_chunk, mychunk, int(bool(keep_raw)) | (2 if defer else 0)), "pgb_load_reads")

    def load_reads_ptr(self, seqdb_ptr: int, seqdb_bytes: int, rid, length, offset, total_chunk=1, mychunk=1, keep_raw=False):
        """Same, from a raw host address (e.g. a pinned torch tensor's data_ptr())."""
        rid = np.ascontiguousarray(rid, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        self._ck(self.L.pgb_load_reads(self.h, C.c_void_p(seqdb_ptr), seqdb_bytes, _ptr(rid), _ptr(length), _ptr(offset),
                                       len(rid), total_chunk, mychunk, int(keep_raw)), "pgb_load_reads")

    def load_reads_from_files(self, prefix, total_chunk=1, mychunk=1, keep_raw=False):
        self._ck(self.L.pgb_load_reads_from_files(self.h, prefix.encode(), total_chunk, mychunk, int(keep_raw)),
                 "pgb_load_reads_from_files")

    def pack_2bit(self, seqdb, offset, length, words_out=None):
        """.seqdb bytes -> (words uint64[sum ceil(len/32)], N records uint32[...] in the .seq2n layout): what this library's
        shmr_mkseqdb writes as <prefix>.seq2b / .seq2n.  words_out: optional preallocated (e.g. pinned) uint64 array."""
        offset = np.ascontiguousarray(offset, dtype=np.uint64)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        nw = (length.astype(np.int64) + 31) // 32
        total = int(nw.sum())
        words = words_out if words_out is not None else np.empty(total, dtype=np.uint64)
        assert words.size == total
        nmask = np.empty(total, dtype=np.uint32)
        hasn = np.empty(len(length), dtype=np.uint8)
        self._ck(self.L.pgb_pack_2bit(self.h, _ptr(seqdb), seqdb.size, _ptr(offset), _ptr(length), len(length), _ptr(words), _ptr(nmask), _ptr(hasn)),
                 "pgb_pack_2bit")
        recs = []
        woff = np.concatenate([[0], np.cumsum(nw)])
        for i in np.nonzero(hasn)[0]:
            recs.append(np.array([i, nw[i]], dtype=np.uint32))
            recs.append(nmask[woff[i]: woff[i + 1]])
        return words, (np.concatenate(recs) if recs else np.empty(0, dtype=np.uint32))

    def load_reads_2bit(self, words, n_records, rid, length, total_chunk=1, mychunk=1, defer=False):
        """Read set from the 2-bit image of ALL reads of the table (pack_2bit / .seq2b); defer as for load_reads."""
        self._seqdb_ref = (words, n_records) if defer else None
        rid = np.ascontiguousarray(rid, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        n_records = np.ascontiguousarray(n_records, dtype=np.uint32)
        self._ck(self.L.pgb_load_reads_2bit(self.h, _ptr(words), words.size, _ptr(n_records), n_records.size, _ptr(rid), _ptr(length), len(rid),
                                            total_chunk, mychunk, 2 if defer else 0), "pgb_load_reads_2bit")

    def repack(self):
        self._ck(self.L.pgb_repack(self.h), "pgb_repack")

    # ------------------------------------------------------------------ index
    def index(self, w=80, k=16, r=6, levels=2, with_counts=0):
        self._ck(self.L.pgb_index(self.h, w, k, r, levels, with_counts), "pgb_index")

    def level(self, level):
        n = self.L.pgb_index_size(self.h, level)
        out = np.empty(n, dtype=formats.MM128)
        if n:
            self._ck(self.L.pgb_index_copy(self.h, level, _ptr(out)), "pgb_index_copy")
        return out

    def level_size(self, level):
        return self.L.pgb_index_size(self.h, level)

    def level_counts(self, level):
        n = self.L.pgb_index_count_size(self.h, level)
        out = np.empty(n, dtype=formats.MMCOUNT)
        if n:
            self._ck(self.L.pgb_index_count_copy(self.h, level, _ptr(out)), "pgb_index_count_copy")
        return out

    # ------------------------------------------------------------------ overlap
    def set_shimmers(self, mmers, counts):
        mmers = np.ascontiguousarray(mmers, dtype=formats.MM128)
        counts = np.ascontiguousarray(counts, dtype=formats.MMCOUNT)
        self._ck(self.L.pgb_set_shimmers(self.h, _ptr(mmers), len(mmers), _ptr(counts), len(counts)), "pgb_set_shimmers")

    def set_shimmers_from_index(self, level=2):
        self._ck(self.L.pgb_set_shimmers_from_index(self.h, level), "pgb_set_shimmers_from_index")

    def overlap(self, total_chunk=1, mychunk=1, bestn=4, mc_lower=2, mc_upper=240, align_bandwidth=100, ovlp_upper=120,
                copy=True):
        """copy=True: records as an independent numpy array; copy="view": zero-copy view of the page-locked staging buffer
        (valid until the next overlap()); copy=False: leave the records on the device, return their number."""
        self._ck(self.L.pgb_overlap(self.h, total_chunk, mychunk, bestn, mc_lower, mc_upper, align_bandwidth, ovlp_upper),
                 "pgb_overlap")
        if copy is False:
            return self.L.pgb_overlap_size(self.h)
        return self.overlap_records(view=(copy == "view"))

    def overlap_records(self, view=False):
        """The chunk's ovlp_t records on the host.  view=True: zero-copy numpy view of the context's page-locked staging
        buffer (valid until the next overlap() / close() of this engine); default: an independent array."""
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self.L.pgb_overlap_host(self.h, C.byref(p), C.byref(n)), "pgb_overlap_host")
        if n.value == 0:
            return np.empty(0, dtype=formats.OVLP)
        buf = (C.c_uint8 * (n.value * formats.OVLP.itemsize)).from_address(p.value)
        a = np.frombuffer(buf, dtype=formats.OVLP, count=n.value)
        return a if view else a.copy()

    # ------------------------------------------------------------------ multi-GPU plumbing (device buffers)
    BUF_WORDS, BUF_NMASK, BUF_ROW_RID, BUF_ROW_LEN, BUF_ROW_WOFF, BUF_ROW_HASN = 0, 1, 2, 3, 4, 5
    BUF_LEVEL0, BUF_LEVEL1, BUF_LEVEL2 = 8, 9, 10

    def buffer_elems(self, which):
        return self.L.pgb_buffer_elems(self.h, which)

    def buffer_copy_out(self, which, dst_device_ptr: int):
        self._ck(self.L.pgb_buffer_copy_out(self.h, which, C.c_void_p(dst_device_ptr)), "pgb_buffer_copy_out")

    def load_packed_device(self, words_ptr, nmask_ptr, n_words, row_rid_ptr, row_len_ptr, row_woff_ptr, row_hasn_ptr, n_rows):
        self._ck(self.L.pgb_load_packed_device(self.h, C.c_void_p(words_ptr), C.c_void_p(nmask_ptr), n_words, C.c_void_p(row_rid_ptr),
                                               C.c_void_p(row_len_ptr), C.c_void_p(row_woff_ptr), C.c_void_p(row_hasn_ptr), n_rows),
                 "pgb_load_packed_device")

    def set_shimmers_device(self, mm_ptr: int, n: int):
        self._ck(self.L.pgb_set_shimmers_device(self.h, C.c_void_p(mm_ptr), n), "pgb_set_shimmers_device")

    # ------------------------------------------------------------------ routed exchange (pair records travel)
    BUF_COUNTS, BUF_ROUTE, BUF_OVLP = 11, 12, 13

    def counts_dump(self) -> int:
        n = C.c_size_t()
        self._ck(self.L.pgb_counts_dump(self.h, C.byref(n)), "pgb_counts_dump")
        return n.value

    def counts_set_device(self, entries_ptr: int, n: int):
        self._ck(self.L.pgb_counts_set_device(self.h, C.c_void_p(entries_ptr), n), "pgb_counts_set_device")

    def route_scan(self, mc_lower=2, mc_upper=240) -> bool:
        f = C.c_int()
        self._ck(self.L.pgb_route_scan(self.h, mc_lower, mc_upper, C.byref(f)), "pgb_route_scan")
        return bool(f.value)

    def route_build(self, total_chunk, mc_lower=2, mc_upper=240, first_found_before=False):
        out = (C.c_uint64 * total_chunk)()
        self._ck(self.L.pgb_route_build(self.h, total_chunk, mc_lower, mc_upper, int(first_found_before), out), "pgb_route_build")
        return [int(x) for x in out]

    def overlap_routed(self, records_ptr: int, n: int, bestn=4, align_bandwidth=100, ovlp_upper=120, copy=True, total_chunk=1):
        self._ck(self.L.pgb_overlap_routed(self.h, C.c_void_p(records_ptr), n, bestn, align_bandwidth, ovlp_upper, total_chunk), "pgb_overlap_routed")
        if copy is False:
            return self.L.pgb_overlap_size(self.h)
        return self.overlap_records(view=(copy == "view"))

    # ------------------------------------------------------------------ shmr_dedup
    def ovlp_match_batch(self, seq, q_off, q_len, q_strand, t_off, t_len, t_strand, band_tolerance=100):
        """ovlp_match (src/DWmatch.c:66-204) for many operand pairs at once: `seq` holds .seqdb bytes, pair i aligns
        seq[q_off[i]:+q_len[i]] on strand q_strand[i] with seq[t_off[i]:+t_len[i]] on t_strand[i].  Returns an (n, 8) int32 array with
        the fields of ovlp_match_t per pair."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        a = [np.ascontiguousarray(x, dtype=t) for x, t in ((q_off, np.uint64), (q_len, np.uint32), (q_strand, np.uint8), (t_off, np.uint64),
                                                           (t_len, np.uint32), (t_strand, np.uint8))]
        out = np.zeros((len(a[0]), 8), dtype=np.int32)
        self._ck(self.L.pgb_ovlp_match_batch(self.h, _ptr(seq), seq.size, len(a[0]), *[_ptr(x) for x in a], int(band_tolerance), _ptr(out)), "ovlp_match_batch")
        return out

    def shmr_aln_batch_raw(self, m0, off0, m1, off1, direction=0, max_diff=100, max_dist=1200, max_repeat=1):
        """pgb_shmr_aln_batch on concatenated MM128 arrays: pair p = m0[off0[p]:off0[p+1]] against m1[off1[p]:off1[p+1]].  Returns
        (hit_off[n+1], n_chains[n], hits[total, 3] = (chain, idx0, idx1) rows in the reference's order)."""
        from . import formats as F

        n = len(off0) - 1
        m0, m1 = np.ascontiguousarray(m0, dtype=F.MM128), np.ascontiguousarray(m1, dtype=F.MM128)
        off0, off1 = np.ascontiguousarray(off0, dtype=np.uint64), np.ascontiguousarray(off1, dtype=np.uint64)
        hit_off = np.zeros(n + 1, dtype=np.uint64)
        n_chains = np.zeros(max(n, 1), dtype=np.uint32)
        hp = C.c_void_p()
        self._ck(self.L.pgb_shmr_aln_batch(self.h, _ptr(m0), _ptr(off0), _ptr(m1), _ptr(off1), n, int(direction), int(max_diff), int(max_dist),
                                           int(max_repeat), _ptr(hit_off), _ptr(n_chains), C.byref(hp)), "shmr_aln_batch")
        total = int(hit_off[n])
        hits = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint32)), shape=(max(total, 1), 3))[:total].copy()
        self.L.pgb_host_free(hp)
        return hit_off, n_chains[:n], hits

    def shmr_aln_batch(self, lists0, lists1, direction=0, max_diff=100, max_dist=1200, max_repeat=1):
        """shmr_aln (src/shmr_align.c:21-160) for many pairs of minimizer lists at once.  lists0 / lists1: sequences of MM128
        arrays (pair p = lists0[p] against lists1[p]).  Returns, per pair, the chains as the reference returns them:
        [(idx0 list, idx1 list), ...] in the order of shmr_aln_v."""
        from . import formats as F

        n = len(lists0)
        assert len(lists1) == n
        if n == 0:
            return []
        off0 = np.zeros(n + 1, dtype=np.uint64)
        off1 = np.zeros(n + 1, dtype=np.uint64)
        off0[1:] = np.cumsum([len(x) for x in lists0])
        off1[1:] = np.cumsum([len(x) for x in lists1])
        cat = lambda ls: np.concatenate([np.asarray(x, dtype=F.MM128) for x in ls])
        hit_off, n_chains, hits = self.shmr_aln_batch_raw(cat(lists0), off0, cat(lists1), off1, direction, max_diff, max_dist, max_repeat)
        out = []
        for p in range(n):
            h = hits[int(hit_off[p]): int(hit_off[p + 1])]
            chains = [([], []) for _ in range(int(n_chains[p]))]
            for ch, i0, i1 in h.tolist():
                chains[ch][0].append(i0)
                chains[ch][1].append(i1)
            out.append(chains)
        return out

    def dedup(self, records=None, device_ptr=None, n=None, text=True):
        """First record of every read pair in stream order -> preads.ovl text (bytes).  records: numpy array of ovlp_t on the
        host; device_ptr/n: a stream already in HBM; neither: the records of this engine's last overlap(), in place.
        text=False leaves the text on the device and returns (kept, bytes)."""
        if records is not None:
            records = np.ascontiguousarray(records)
            self._ck(self.L.pgb_dedup(self.h, _ptr(records), len(records)), "pgb_dedup")
        elif device_ptr is not None:
            self._ck(self.L.pgb_dedup_device(self.h, C.c_void_p(device_ptr), n), "pgb_dedup_device")
        else:
            self._ck(self.L.pgb_dedup_overlaps(self.h), "pgb_dedup_overlaps")
        nb = self.L.pgb_dedup_text_bytes(self.h)
        if not text:
            return self.L.pgb_dedup_kept(self.h), nb
        out = np.empty(nb, dtype=np.uint8)
        if nb:
            self._ck(self.L.pgb_dedup_text_copy(self.h, _ptr(out)), "pgb_dedup_text_copy")
        return out.tobytes()

    # ------------------------------------------------------------------ shmr_mkseqdb
    def encode_biseq(self, ascii_bytes, offset, length):
        """encode_biseq of a batch of reads: ascii_bytes (uint8 array), read i at offset[i] with length[i] -> .seqdb bytes."""
        a = np.ascontiguousarray(ascii_bytes, dtype=np.uint8)
        off = np.ascontiguousarray(offset, dtype=np.uint64)
        ln = np.ascontiguousarray(length, dtype=np.uint32)
        out = np.empty(a.size, dtype=np.uint8)
        self._ck(self.L.pgb_encode_biseq(self.h, _ptr(a), a.size, _ptr(off), _ptr(ln), len(ln), _ptr(out)), "pgb_encode_biseq")
        return out

    # ------------------------------------------------------------------ shmr_map
    def set_read_lengths(self, rid, length):
        rid = np.ascontiguousarray(rid, dtype=np.uint32)
        length = np.ascontiguousarray(length, dtype=np.uint32)
        self._ck(self.L.pgb_set_read_lengths(self.h, _ptr(rid), _ptr(length), len(rid)), "pgb_set_read_lengths")

    def map(self, ref_mmers, total_chunk=1, mychunk=1, mc_lower=1, mc_upper=240) -> bytes:
        """shmr_map's hit lines for the contig shimmer list ref_mmers against this engine's shimmers (set_shimmers)."""
        ref_mmers = np.ascontiguousarray(ref_mmers)
        self._ck(self.L.pgb_map(self.h, _ptr(ref_mmers), len(ref_mmers), total_chunk, mychunk, mc_lower, mc_upper), "pgb_map")
        out = np.empty(self.L.pgb_map_text_bytes(self.h), dtype=np.uint8)
        if out.size:
            self._ck(self.L.pgb_map_text_copy(self.h, _ptr(out)), "pgb_map_text_copy")
        return out.tobytes()

    # ------------------------------------------------------------------ stats
    def event_record(self, slot):
        self.L.pgb_event_record(self.h, slot)

    def event_elapsed_ms(self, a, b):
        return self.L.pgb_event_elapsed_ms(self.h, a, b)

    def stats_reset(self):
        self.L.pgb_stats_reset(self.h)

    def stats(self) -> dict:
        s = _Stats()
        self.L.pgb_stats_get(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in _Stats._fields_}
