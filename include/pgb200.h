/*
 * pgb200.h — C ABI of libpgb200.so, the B200-native SHIMMER index + read-to-read overlap engine.
 *
 * Three groups of entry points; every one is plain C (pointers + sizes, no C++/torch types):
 *
 *  (1) the reference's own cffi surface, same names, argument meaning, struct layouts and ownership rules as the
 *      cdef at py/peregrine/build_shimmer4py.py:8-84 (a library built from this header can be dlopen'ed in place
 *      of peregrine._shimmer4py's C side);
 *  (2) the two command-line tools as callable mains — same getopt strings, defaults, file names and messages as
 *      src/shmr_index.c:37-245 and src/shmr_overlap.c:233-419 — which bin/shmr_index and bin/shmr_overlap wrap;
 *  (3) a stage-level context API (load reads -> index chunk -> overlap chunk) over host buffers, used by bench.py,
 *      the tests and the multi-GPU driver; it is what (2) is written on.
 *
 * All computation runs in CUDA kernels (sm_100a).  There is no CPU fallback: without a usable CUDA device every
 * computing entry point prints a message and fails (returns non-zero / exits like the reference does on errors).
 */
#ifndef PGB200_H
#define PGB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ shared types (layouts are ABI) -------- */
typedef struct { uint64_t x, y; } mm128_t;                  /* src/shimmer.h:24-26 : x = hash<<8|span, y = rid<<32|pos<<1|strand */
typedef struct { size_t n, m; mm128_t *a; } mm128_v;        /* src/shimmer.h:27-30 : .a is libc realloc memory */
typedef struct { uint64_t mer; uint32_t count; } mm_count_t;/* src/shimmer.h:61-64 : 16 bytes, 4 trailing pad bytes (zeroed here) */
typedef int32_t seq_coor_t;                                 /* src/shimmer.h:95 */
typedef struct {                                            /* src/shimmer.h:97-102 */
  seq_coor_t m_size, dist;
  seq_coor_t q_bgn, q_end;
  seq_coor_t t_bgn, t_end;
  seq_coor_t t_m_end, q_m_end;
} ovlp_match_t;
typedef struct {                                            /* src/shimmer.h:104-110 : 64 bytes; pad bytes 27, 60..63 zeroed */
  uint64_t y0, y1;
  uint32_t rl0, rl1;
  uint8_t strand0, strand1;
  uint8_t ovlp_type;                                        /* 0 overlap, 1 contains, 2 contained (src/shmr_overlap.c:37-39) */
  ovlp_match_t match;
} ovlp_t;
typedef struct { uint64_t x0, x1, y0, y1; uint8_t direction; } mp256_t;  /* src/shimmer.h:123-126 */
typedef struct { size_t n, m; mp256_t *a; } mp256_v;
typedef struct { mm128_v *mmers; void *mmer0_map; void *rlmap; void *mcmap; void *ridmm; } py_mmer_t; /* src/shimmer.h:132-138 */
typedef uint32_t mm_idx_t;
typedef struct { size_t n, m; mm_idx_t *a; } mm_idx_v;
typedef struct { mm_idx_v idx0; mm_idx_v idx1; } shmr_aln_t;            /* src/shimmer.h:144-147 */
typedef struct { size_t n, m; shmr_aln_t *a; } shmr_aln_v;

/* ------------------------------------------------------------------ (1) reference cffi surface ------------ */
/* replaces src/shmr_utils.c:56-62 */
void decode_biseq(uint8_t *src, char *seq, size_t len, uint8_t strand);
/* replaces src/DWmatch.c:66-204; operands are .seqdb-format bytes; result is calloc'd, free with free_ovlp_match */
ovlp_match_t *ovlp_match(uint8_t *query_seq, seq_coor_t q_len, uint8_t q_strand, uint8_t *target_seq, seq_coor_t t_len,
                         uint8_t t_strand, seq_coor_t band_tolerance);
void free_ovlp_match(ovlp_match_t *match);                 /* src/DWmatch.c:206 */
/* replaces src/shmr_utils.c:110-123; returned BY VALUE, .a freed by the caller with free() */
mm128_v read_mmlist(char *fn);
/* replaces src/mm_sketch.c:70-151; APPENDS to p; is_hpc must be 0 (every reference call site passes 0) */
void mm_sketch(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p);
/* replaces src/shmr_reduce.c:53-90; appends to out */
void mm_reduce(mm128_v *in, mm128_v *out, uint8_t rs);

/* replaces src/shmr_align.c:21-160 (greedy co-linear chaining of shared minimizers); result freed with free_shmr_alns.
 * direction 1 reads one element past the end of the second list in the reference (UB); that element is skipped here. */
shmr_aln_v *shmr_aln(mm128_v *mmers0, mm128_v *mmers1, uint8_t direction, uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat);
void free_shmr_alns(shmr_aln_v *alns);                      /* src/shmr_align.c:162-169 */

/* replaces src/shimmer4py.c:44-196: an in-memory SHIMMER-pair index for Python.  The multiplicity table, the pair buckets and
 * their khash visiting order live in HBM behind the handle's void* fields; `mmers` is a host copy of the concatenated list. */
void build_shimmer_map4py(py_mmer_t *py_mmer, char *seqdb_prefix, char *shimmer_prefix, uint32_t mychunk, uint32_t total_chunk,
                          uint32_t lowerbound, uint32_t upperbound);
void get_shimmers_for_read(mm128_v *out, py_mmer_t *py_mmer, uint32_t rid);   /* out aliases py_mmer->mmers; do not free */
uint32_t get_mmer_count(py_mmer_t *py_mmer, uint64_t mhash);
void get_shimmer_hits(mp256_v *append_to, py_mmer_t *py_mmer, uint64_t mhash0, uint32_t span);

/* ------------------------------------------------------------------ (2) command-line tools ---------------- */
/* replaces main() of src/shmr_index.c:37-245 : -p seqdb_prefix -o out_prefix -t T -c c [-r 6] [-l 2] [-m 1] [-w 80] [-k 16] */
int pgb_shmr_index_main(int argc, char **argv);
/* replaces main() of src/shmr_overlap.c:233-419 : -p seqdb_prefix -l index_prefix -t T -c c -o file [-b 4] [-m 2] [-M 240] [-w 100] [-n 120] */
int pgb_shmr_overlap_main(int argc, char **argv);

/* ------------------------------------------------------------------ (3) stage-level API ------------------- */
typedef struct pgb_ctx pgb_ctx;

int pgb_device_count(void);                       /* number of usable CUDA devices (0 => nothing can run) */
pgb_ctx *pgb_create(int device);                  /* NULL on failure (message on stderr) */
void pgb_destroy(pgb_ctx *);
const char *pgb_last_error(pgb_ctx *);            /* "" when the last call succeeded */

/* Read set = .seqdb image + .idx table (src/shmr_mkseqdb.c:108-118).  Rows with rid % total_chunk == mychunk % total_chunk
 * (src/shmr_index.c:157) are copied to the device and 2-bit packed; total_chunk = 1 selects everything.
 * seqdb may be pageable or pinned host memory.  keep_raw is a bit set: PGB_LOAD_KEEP_RAW keeps the 1-byte/base image in
 * HBM so that pgb_repack() can redo the packing without another host copy (bench.py's device-resident timing);
 * PGB_LOAD_DEFER leaves the bulk of the image on the host and lets the next pgb_index() overlap its host->device copy
 * with packing and sketching, chunk by chunk (any other call that needs the reads completes the copy first) - the
 * caller must keep seqdb alive and unchanged until then; it only pays off for page-locked memory. */
#define PGB_LOAD_KEEP_RAW 1
#define PGB_LOAD_DEFER 2
int pgb_load_reads(pgb_ctx *, const uint8_t *seqdb, size_t seqdb_bytes, const uint32_t *rid, const uint32_t *len,
                   const uint64_t *offset, size_t n_reads, uint32_t total_chunk, uint32_t mychunk, int keep_raw);
int pgb_repack(pgb_ctx *);
/* The 2-bit hand-off.  pgb_pack_2bit: .seqdb bytes of a batch of reads -> their packed form on the host: per read
 * ceil(len/32) uint64 words (base p at bits 2(p&31) of word p/32, A0 C1 G2 T3, reads in the order given), the parallel N-mask
 * words (uint32 per 32 bases) and a per-read "contains N" flag.  This library's shmr_mkseqdb writes them next to the .seqdb as
 * <prefix>.seq2b (the words of all reads in .idx order) and <prefix>.seq2n (uint32 stream {read index, word count, mask words}
 * for the reads with N); shmr_index / shmr_overlap use them when present and current (PGB_NO_SEQ2B=1: ignore).
 * pgb_load_reads_2bit: the read set from that image (ALL reads of the table, table order); rows are selected like
 * pgb_load_reads does; a quarter of the bytes crosses the bus and nothing is packed on the device.  flags: PGB_LOAD_DEFER. */
int pgb_pack_2bit(pgb_ctx *, const uint8_t *seqdb, size_t seqdb_bytes, const uint64_t *offset, const uint32_t *len, size_t n_reads,
                  uint64_t *words_out, uint32_t *nmask_out, uint8_t *hasn_out);
int pgb_load_reads_2bit(pgb_ctx *, const uint64_t *words, size_t n_words_total, const uint32_t *n_records, size_t n_record_words,
                        const uint32_t *rid, const uint32_t *len, size_t n_reads, uint32_t total_chunk, uint32_t mychunk, int flags);
/* convenience: parse <prefix>.idx, mmap <prefix>.seqdb, then pgb_load_reads */
int pgb_load_reads_from_files(pgb_ctx *, const char *seqdb_prefix, uint32_t total_chunk, uint32_t mychunk, int keep_raw);

/* Sketch + hierarchical reduction of the loaded (selected) reads: L0 = mm_sketch(w,k), L1 = mm_reduce(L0,r),
 * L2 = mm_reduce(L1,r) (src/shmr_index.c:155-216).  levels = 1 or 2.  with_counts: bit i set => also build the
 * multiplicity table of level i (the -MC- files, src/shmr_index.c:25-32). */
int pgb_index(pgb_ctx *, int w, int k, int r, int levels, int with_counts);
size_t pgb_index_size(pgb_ctx *, int level);                    /* number of mm128_t at level 0/1/2 */
int pgb_index_copy(pgb_ctx *, int level, mm128_t *out);          /* device -> host */
size_t pgb_index_count_size(pgb_ctx *, int level);              /* distinct mers of that level */
int pgb_index_count_copy(pgb_ctx *, int level, mm_count_t *out); /* (mer,count), order unspecified (a set) */

/* Overlap input: the concatenated SHIMMER list of all index chunks in file order and all count entries
 * (src/shmr_overlap.c:359-382).  Host-buffer form, and a device hand-off form that re-uses the level just built by
 * pgb_index (single-chunk jobs: T_idx == 1). */
int pgb_set_shimmers(pgb_ctx *, const mm128_t *mmers, size_t n, const mm_count_t *counts, size_t n_counts);
int pgb_set_shimmers_from_index(pgb_ctx *, int level);

/* build_map + process_overlaps for hash chunk `mychunk` of `total_chunk` (src/shmr_overlap.c:394-397).  Needs ALL reads
 * loaded (pgb_load_reads with total_chunk = 1).  Records come back in the reference's output order. */
int pgb_overlap(pgb_ctx *, uint32_t total_chunk, uint32_t mychunk, uint32_t bestn, uint32_t mc_lower, uint32_t mc_upper,
                uint32_t align_bandwidth, uint32_t ovlp_upper);
size_t pgb_overlap_size(pgb_ctx *);
int pgb_overlap_copy(pgb_ctx *, ovlp_t *out);
/* Same records through a page-locked host buffer owned by the context (a device-to-host copy into pageable memory runs at
 * a fraction of the PCIe rate).  *out stays valid until the next pgb_overlap / pgb_destroy on this context. */
int pgb_overlap_host(pgb_ctx *, const ovlp_t **out, size_t *n);

/* ---- multi-GPU plumbing (one process per GPU; the collective itself is the caller's, e.g. torch.distributed / NCCL) ----
 * A rank sketches the reads of ITS index chunk, then the ranks all-gather (a) the packed reads + read table so that any rank
 * can align any pair and (b) the SHIMMER lists (chunk order = file order of src/shmr_overlap.c:359-369).  These calls move
 * the library's device buffers to / from caller-owned DEVICE buffers; every call synchronises the library's stream. */
enum pgb_buffer {
  PGB_BUF_WORDS = 0,    /* uint64[n]: 2-bit packed reads of the loaded rows, incl. 2 guard words at each end          */
  PGB_BUF_NMASK = 1,    /* uint32[n]: N mask, parallel to PGB_BUF_WORDS                                               */
  PGB_BUF_ROW_RID = 2,  /* uint32[rows]                                                                                */
  PGB_BUF_ROW_LEN = 3,  /* uint32[rows]                                                                                */
  PGB_BUF_ROW_WOFF = 4, /* uint64[rows]: word offset of each row inside PGB_BUF_WORDS                                  */
  PGB_BUF_ROW_HASN = 5, /* uint32[rows]: 1 if the read contains a non-ACGT base                                        */
  PGB_BUF_LEVEL0 = 8, PGB_BUF_LEVEL1 = 9, PGB_BUF_LEVEL2 = 10, /* mm128_t[n] of the index level                       */
  PGB_BUF_COUNTS = 11,  /* mm_count_t[n]: the multiplicity table dumped by pgb_counts_dump                             */
  PGB_BUF_ROUTE = 12,   /* mp256_t-like 40-byte records {x0,x1,y0,y1,direction(u64)} grouped by owner chunk 1..T       */
  PGB_BUF_OVLP = 13     /* ovlp_t[n]: the records of the last pgb_overlap / pgb_overlap_routed (for a cross-rank shmr_dedup) */
};
size_t pgb_buffer_elems(pgb_ctx *, int which);                       /* element count of a buffer                      */
int pgb_buffer_copy_out(pgb_ctx *, int which, void *dst_device);     /* device -> caller's device buffer               */
/* Replace the context's read set by an already packed one (device pointers, e.g. the concatenation of every rank's
 * buffers with row_woff rebased to the concatenated word array). */
int pgb_load_packed_device(pgb_ctx *, const uint64_t *words, const uint32_t *nmask, size_t n_words, const uint32_t *row_rid,
                           const uint32_t *row_len, const uint64_t *row_woff, const uint32_t *row_hasn, size_t n_rows);
/* Overlap input from a DEVICE array of mm128_t (all chunks concatenated in chunk order); the multiplicity table is rebuilt
 * from it, which equals summing the per-chunk -MC- files (src/shmr_utils.c:162-176). */
int pgb_set_shimmers_device(pgb_ctx *, const mm128_t *mmers_device, size_t n);

/* ---- routed exchange: the SHIMMER-pair records travel, not the lists (one all-to-all) --------------------------
 * Each rank keeps the shimmers of its own reads (pgb_set_shimmers_from_index).  build_map (src/shmr_utils.c:295-404) is a
 * single scan over the concatenation of all chunk lists; because a read never straddles two index chunks the scan splits
 * by rank, except for (a) the multiplicities, which are global, and (b) the asymmetric bound of the very first kept element
 * (:318 `< upper` vs :327 `<= upper`), which only the first rank that has such an element applies.
 *   pgb_counts_dump        compact the context's multiplicity table into PGB_BUF_COUNTS (this rank's partial counts)
 *   pgb_counts_set_device  replace the table by the SUM of the given entries (all ranks' partials concatenated; duplicates
 *                          add up, as aggregate_mm_count does over the -MC- files, src/shmr_utils.c:162-176)
 *   pgb_route_scan         global-count lookup of every local shimmer; *has_first = 1 if one has lower <= count < upper
 *   pgb_route_build        first_found_before = 1 if a LOWER rank reported has_first.  Emits the forward and the reverse record
 *                          of every kept pair into PGB_BUF_ROUTE, grouped by owner chunk (chunk c owns (x>>8) % T == c % T,
 *                          src/shmr_utils.c:337,362), scan order kept inside a group; n_per_chunk[c-1] = size of group c
 *   pgb_overlap_routed     process_overlaps over the records an owner received: source ranks concatenated in rank order.
 *                          Needs all reads loaded (pgb_load_packed_device); the result is read like pgb_overlap's. */
int pgb_counts_dump(pgb_ctx *, size_t *n_entries);
int pgb_counts_set_device(pgb_ctx *, const mm_count_t *entries_device, size_t n);
int pgb_route_scan(pgb_ctx *, uint32_t mc_lower, uint32_t mc_upper, int *has_first);
int pgb_route_build(pgb_ctx *, uint32_t total_chunk, uint32_t mc_lower, uint32_t mc_upper, int first_found_before, uint64_t *n_per_chunk);
int pgb_overlap_routed(pgb_ctx *, const void *records_device, size_t n, uint32_t bestn, uint32_t align_bandwidth, uint32_t ovlp_upper,
                       uint32_t total_chunk /* sizes the per-chunk replay tables: rid_pairs is per chunk, src/shmr_overlap.c:202 */);

/* ---- batched ovlp_match for the cffi callers (py/scripts/path_to_contig.py:82-105 stitches contigs with one call per
 * overlap; py/peregrine/utils.py).  Pair i aligns seq[q_off[i], +q_len[i]) read on strand q_strand[i] with seq[t_off[i], +t_len[i]) on
 * t_strand[i]; seq holds .seqdb bytes (the strand selects the nibble, src/DWmatch.c:90-91).  out[i] = what ovlp_match returns
 * for that pair.  One upload, one warp per pair, one download; the context's loaded reads are not touched. */
int pgb_ovlp_match_batch(pgb_ctx *, const uint8_t *seq, size_t seq_bytes, size_t n_pairs, const uint64_t *q_off, const uint32_t *q_len,
                         const uint8_t *q_strand, const uint64_t *t_off, const uint32_t *t_len, const uint8_t *t_strand, int band_tolerance,
                         ovlp_match_t *out);

/* ---- batched shmr_aln for the cffi callers (py/peregrine/utils.py:52-73 get_shimmer_alns; src/shmr_align.c:21-160).  Pair p chains
 * mm0[off0[p], off0[p+1]) against mm1[off1[p], off1[p+1]) with the reference's parameters.  On return the hits of pair p are
 * (*hits)[hit_off[p] .. hit_off[p+1]) in the order the reference appends them (idx0 / idx1 index the pair's own lists, chain = the
 * position of the chain in the reference's shmr_aln_v), n_chains[p] = alns->n.  *hits is malloc'd: release it with pgb_host_free.
 * One upload, one radix sort for the hash -> index map of all pairs, one warp per pair for the greedy chaining, one download. */
typedef struct { uint32_t chain, idx0, idx1; } pgb_aln_hit_t;
int pgb_shmr_aln_batch(pgb_ctx *, const mm128_t *mm0, const uint64_t *off0, const mm128_t *mm1, const uint64_t *off1, uint32_t n_pairs,
                       uint8_t direction, uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat, uint64_t *hit_off /* [n_pairs+1] */,
                       uint32_t *n_chains /* [n_pairs] or NULL */, pgb_aln_hit_t **hits);
void pgb_host_free(void *);

/* ---- shmr_dedup: raw ovlp_t stream -> preads.ovl text ------------------------------------------------------
 * replaces main() of src/shmr_dedup.c:19-101: keeps the FIRST record of every unordered read pair in stream order (the
 * concatenation of the chunk files, `cat ovlp-*.dat | shmr_dedup`, py/scripts/pg_run.py:352) and prints
 * "%09d %09d %d %0.1f %u %d %d %u %u %d %d %u %s\n" per kept record, byte-identical to the reference (incl. its unsigned
 * coordinate arithmetic and "%0.1f" rounding).  An EMPTY stream yields no output (the reference prints one line from an
 * uninitialised struct).  pgb_shmr_dedup_main: stdin -> stdout, no options. */
int pgb_shmr_dedup_main(int argc, char **argv);
int pgb_dedup(pgb_ctx *, const ovlp_t *records, size_t n);               /* host stream                                   */
int pgb_dedup_device(pgb_ctx *, const ovlp_t *records_device, size_t n); /* stream already in HBM                         */
/* the same stream in bounded batches (what bin/shmr_dedup does with stdin): the pair table persists on the device between the pushes, a
 * push leaves the lines of the records IT keeps (pgb_dedup_text_bytes / pgb_dedup_text_copy); the concatenation of the pushes' texts is
 * the text of the whole stream.  Host memory: one batch; device memory: 16 B per distinct pair at load <= 0.5 + one batch. */
int pgb_dedup_stream_begin(pgb_ctx *);
int pgb_dedup_stream_push(pgb_ctx *, const ovlp_t *records, size_t n);
int pgb_dedup_stream_end(pgb_ctx *);
int pgb_dedup_overlaps(pgb_ctx *);                                       /* the records of the last pgb_overlap, in place */
size_t pgb_dedup_kept(pgb_ctx *);
size_t pgb_dedup_text_bytes(pgb_ctx *);
int pgb_dedup_text_copy(pgb_ctx *, char *out);                           /* device -> host, pgb_dedup_text_bytes bytes     */

/* ---- shmr_mkseqdb: FASTA/FASTQ(.gz) -> <prefix>.idx + <prefix>.seqdb -------------------------------------------
 * pgb_shmr_mkseqdb_main replaces main() of src/shmr_mkseqdb.c:15-132 (-d file list, -p output prefix; same defaults and
 * messages).  pgb_encode_biseq replaces encode_biseq (src/shmr_utils.c:44-51) for a batch: read i is
 * ascii[offset[i] .. offset[i] + len[i]); its .seqdb bytes are written to seqdb_out at the same offsets (host buffers). */
int pgb_shmr_mkseqdb_main(int argc, char **argv);
int pgb_encode_biseq(pgb_ctx *, const char *ascii, size_t total_bytes, const uint64_t *offset, const uint32_t *len, size_t n_reads,
                     uint8_t *seqdb_out);

/* ---- shmr_map: contig shimmers against the reads' SHIMMER-pair index, no alignment ------------------------------
 * pgb_shmr_map_main replaces main() of src/shmr_map.c:168-380 (-r -m -p -l -M -n -t -c; hits on stdout as
 * "%u %u %u %u %u %u %d %u %u\n" = ref_id ref_bgn ref_end read_id read_bgn read_end direction mcount0 mcount1).
 * pgb_map replaces build_map + process_map (src/shmr_map.c:48-166) for the context's shimmers (pgb_set_shimmers: the reads'
 * lists + count files) and read lengths (pgb_set_read_lengths, or any pgb_load_reads*); ref_mmers = the contigs' lists
 * concatenated in file order. */
int pgb_shmr_map_main(int argc, char **argv);
int pgb_set_read_lengths(pgb_ctx *, const uint32_t *rid, const uint32_t *len, size_t n_reads);
int pgb_map(pgb_ctx *, const mm128_t *ref_mmers, size_t n_ref, uint32_t total_chunk, uint32_t mychunk, uint32_t mc_lower, uint32_t mc_upper);
size_t pgb_map_hits(pgb_ctx *);
size_t pgb_map_text_bytes(pgb_ctx *);
int pgb_map_text_copy(pgb_ctx *, char *out);

/* counters for bench.py / profiles */
typedef struct {
  uint64_t kernel_launches;      /* launches of this library's kernels since pgb_stats_reset */
  uint64_t h2d_bytes, d2h_bytes;
  double ms_pack, ms_sketch, ms_reduce, ms_count, ms_pairs, ms_buckets, ms_host_order, ms_replay, ms_align, ms_emit;
  uint64_t bases_packed, bases_sketched, n_l0, n_l1, n_l2;
  uint64_t n_pair_records, n_buckets, n_eligible_buckets, n_candidates;
  uint64_t n_alignments, n_align_bases, n_replay_passes, n_overlaps;
  /* per-kernel CUDA-event time (sum over launches) and launch counts of the three heavy kernels */
  double ms_k_sketch_count, ms_k_sketch_write, ms_k_align, ms_k_replay;
  uint64_t n_k_sketch_count, n_k_sketch_write, n_k_align, n_k_replay;
  double ms_k_sketch_tiled;
  uint64_t n_k_sketch_tiled, n_sketch_fallback_reads;
  uint64_t n_replay_buckets;     /* buckets replayed, summed over passes (incremental passes replay only dirty buckets) */
  double ms_dedup;               /* shmr_dedup stage */
  uint64_t n_dedup_in, n_dedup_kept;
  double ms_encode, ms_k_encode; /* pgb_encode_biseq: whole call incl. copies / kernel only */
  uint64_t n_k_encode, bases_encoded;
  double ms_map;                 /* pgb_map after the pair records are built */
  uint64_t n_map_hits;
  uint64_t n_device_mallocs;     /* cudaMalloc calls (block-cache misses + new scratch slabs): 0 per step in a steady-state job */
  uint64_t n_replay_restarts;    /* fix-point restarts after a replay table filled up */
} pgb_stats;
void pgb_stats_reset(pgb_ctx *);
/* CUDA events on the context's stream (the stream every kernel of this library is launched on): record slot 0..7, then
 * read the device time between two recorded slots. */
int pgb_event_record(pgb_ctx *, int slot);
double pgb_event_elapsed_ms(pgb_ctx *, int slot_a, int slot_b);
void pgb_stats_get(pgb_ctx *, pgb_stats *out);

#ifdef __cplusplus
}
#endif
#endif
