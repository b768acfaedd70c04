// kernels.cuh — sm_100a CUDA kernels of the SHIMMER index / overlap path.
//
// Kernel                      replaces (reference)                               work item
// k_pack_reads                decode_biseq use, src/shmr_index.c:159             1 thread = 32 bases -> one u64 word + N mask
// k_sketch_exact<WRITE>       mm_sketch, src/mm_sketch.c:70-151                  1 thread = 1 read (exact automaton)
// k_reduce<WRITE>             mm_reduce, src/shmr_reduce.c:53-90                 1 thread = 1 read's mmer run
// k_mc_insert / k_mc_add      mm_count / aggregate_mm_count, shmr_utils.c:131-176  1 thread = 1 mmer, open-addressing table
// k_count_lookup..k_pair_*    build_map, src/shmr_utils.c:295-404                1 thread = 1 (kept) mmer / adjacent pair
// k_bucket_insert/scatter/sort   MMER0/MMER1 khash + qsort, shmr_overlap.c:206-217  1 thread = 1 record / 1 bucket
// k_replay                    shimmer_to_overlap, src/shmr_overlap.c:52-180      1 thread = 1 (x0,x1) bucket
// k_align                     ovlp_match, src/DWmatch.c:66-204                   1 thread = 1 alignment (band state in local mem)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "shimmer_core.cuh"
#include "sketch_tile.cuh"
#include "sketch_strip.cuh"
#include "khash_small.cuh"

namespace pgb {

struct AlnReqPOD { uint32_t rid0, start0, rid1, strands, slot; };  // = AlnReq (defined with the replay tables below)

// ------------------------------------------------------------------------------------------------ hash table (u64 keys)
#define PGB_EMPTY 0xFFFFFFFFFFFFFFFFULL
#define PGB_NOSLOT 0xFFFFFFFFu

__device__ __forceinline__ uint32_t ht_mix(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}
// returns slot, or PGB_NOSLOT when the table is full (caller raises the error flag)
__device__ __forceinline__ uint32_t ht_insert(uint64_t *keys, uint32_t mask, uint64_t key) {
  uint32_t h = ht_mix(key) & mask;
  for (uint32_t probe = 0; probe <= mask; probe++) {
    uint64_t cur = keys[h];
    if (cur == key) return h;
    if (cur == PGB_EMPTY) {
      uint64_t prev = atomicCAS((unsigned long long *)&keys[h], (unsigned long long)PGB_EMPTY, (unsigned long long)key);
      if (prev == PGB_EMPTY || prev == key) return h;
    }
    h = (h + 1) & mask;
  }
  return PGB_NOSLOT;
}
__device__ __forceinline__ uint32_t ht_find(const uint64_t *keys, uint32_t mask, uint64_t key) {
  uint32_t h = ht_mix(key) & mask;
  for (uint32_t probe = 0; probe <= mask; probe++) {
    uint64_t cur = keys[h];
    if (cur == key) return h;
    if (cur == PGB_EMPTY) return PGB_NOSLOT;
    h = (h + 1) & mask;
  }
  return PGB_NOSLOT;
}

__global__ void k_fill_u64(uint64_t *p, uint64_t v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
__global__ void k_fill_u32(uint32_t *p, uint32_t v, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ pack
// raw: staged .seqdb bytes of the selected reads, read `row` at raw[row_raw_off[row] ...).  One thread builds one packed
// word (32 bases) of one read.  row_woff is ascending, so the owning row is found by binary search on the word index.
__global__ void k_pack_reads(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ row_raw_off,
                             const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                             const uint32_t *__restrict__ row_rid, uint32_t n_rows, uint64_t first_word, uint64_t n_words,
                             uint64_t *__restrict__ w, uint32_t *__restrict__ nm, uint32_t *__restrict__ hasn_by_rid, uint64_t raw_shift = 0) {
  // raw_shift: `raw` is a staging window that holds the image from byte raw_shift on (pgb_load_reads packs big inputs window by window)
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_words) return;
  uint64_t word = first_word + g;
  // last row with row_woff <= word
  uint32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (row_woff[mid] <= word) lo = mid; else hi = mid;
  }
  uint32_t row = lo;
  uint64_t p0 = (word - row_woff[row]) * 32;
  uint32_t len = row_len[row];
  if (p0 >= len) {  // padding word of an empty read / guard
    w[word] = 0;
    nm[word] = 0;
    return;
  }
  const uint8_t *s = raw + (row_raw_off[row] - raw_shift) + p0;
  uint32_t cnt = (len - p0) < 32 ? (uint32_t)(len - p0) : 32u;
  uint64_t bits = 0;
  uint32_t nmask = 0;
  if (cnt == 32) {
    // full word: 32 bytes as 8 (unaligned) u32, 4 bases at a time with byte-parallel arithmetic
    const uintptr_t a = (uintptr_t)s;
    const uint32_t *p = (const uint32_t *)(a & ~(uintptr_t)3);  // raw has 64 bytes of slack behind the last read
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t prev = p[0];
#pragma unroll
    for (uint32_t q = 0; q < 8; q++) {
      const uint32_t nx = p[q + 1];
      const uint32_t x = __funnelshift_r(prev, nx, sh) & 0x0F0F0F0Fu;  // low nibbles = forward bases
      prev = nx;
      uint32_t pc = (x & 0x05050505u) + ((x >> 1) & 0x05050505u);       // per-byte popcount of the nibble
      pc = (pc & 0x03030303u) + ((pc >> 2) & 0x03030303u);
      const uint32_t z = pc ^ 0x01010101u;                                // non-zero byte <=> not exactly one bit <=> 'N'
      const uint32_t nz = (((z + 0x7F7F7F7Fu) | z) & 0x80808080u) >> 7;   // 1 per 'N' byte
      uint32_t code = ((x >> 1) & 0x07070707u) - ((x >> 3) & 0x01010101u);  // 1 2 4 8 -> 0 1 2 3
      code &= ~(nz * 0xFFu);
      bits |= (uint64_t)((code * 0x01041040u) >> 24) << (8 * q);          // gather the four 2-bit codes
      nmask |= ((nz * 0x01020408u) >> 24) << (4 * q);                     // gather the four N bits
    }
    w[word] = bits;
    nm[word] = nmask;
    if (nmask) atomicOr(&hasn_by_rid[row_rid[row]], 1u);
    return;
  }
#pragma unroll 8
  for (uint32_t j = 0; j < 32; j++) {
    if (j < cnt) {
      uint32_t nib = s[j] & 0xF;
      // A=1 C=2 G=4 T=8 -> 0 1 2 3 ; anything else is 'N' (src/shmr_utils.c:53-54 bits_to_base)
      uint32_t code = (nib == 2) ? 1u : (nib == 4) ? 2u : (nib == 8) ? 3u : 0u;
      uint32_t isn = !(nib == 1 || nib == 2 || nib == 4 || nib == 8);
      bits |= (uint64_t)code << (2 * j);
      nmask |= isn << j;
    }
  }
  w[word] = bits;
  nm[word] = nmask;
  if (nmask) atomicOr(&hasn_by_rid[row_rid[row]], 1u);
}

// Operands of the cffi ovlp_match calls: .seqdb bytes whose strand picks the nibble (src/DWmatch.c:90-91,136-137).  One thread
// builds one packed word of one operand from the nibble its strand selects (shift 0 or 4); anything but A/C/G/T sets the N bit.
__global__ void k_pack_nibbles(const uint8_t *__restrict__ raw, const uint64_t *__restrict__ row_raw_off, const uint32_t *__restrict__ row_len,
                               const uint64_t *__restrict__ row_woff, const uint8_t *__restrict__ row_shift, uint32_t n_rows, uint64_t n_words,
                               uint64_t *__restrict__ w, uint32_t *__restrict__ nm, uint32_t *__restrict__ hasn_by_row) {
  uint64_t word = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (word >= n_words) return;
  uint32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (row_woff[mid] <= word) lo = mid; else hi = mid;
  }
  const uint32_t row = lo, len = row_len[row], sh = row_shift[row];
  const uint64_t p0 = (word - row_woff[row]) * 32;
  uint64_t bits = 0;
  uint32_t nmask = 0;
  if (word >= row_woff[row] && p0 < len) {
    const uint8_t *s = raw + row_raw_off[row] + p0;
    const uint32_t cnt = (len - p0) < 32 ? (uint32_t)(len - p0) : 32u;
    for (uint32_t j = 0; j < cnt; j++) {
      const uint32_t nib = (s[j] >> sh) & 0xF;
      const uint32_t code = (nib == 2) ? 1u : (nib == 4) ? 2u : (nib == 8) ? 3u : 0u;
      const uint32_t isn = !(nib == 1 || nib == 2 || nib == 4 || nib == 8);
      bits |= (uint64_t)code << (2 * j);
      nmask |= isn << j;
    }
    if (nmask) atomicOr(&hasn_by_row[row], 1u);
  }
  w[word] = bits;
  nm[word] = nmask;
}
__global__ void k_pair_requests(uint32_t n, AlnReqPOD *reqs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AlnReqPOD q;
  q.rid0 = 2 * i; q.start0 = 0; q.rid1 = 2 * i + 1; q.strands = 0; q.slot = i;  // the strands were applied when the nibbles were packed
  reqs[i] = q;
}

// Reverse-complement image of the packed reads (same word offsets as the forward image): base p of read r in wrc is
// 3 - (forward base len-1-p), which is the high nibble of .seqdb byte p (src/shmr_utils.c:44-51).  A reverse-strand
// alignment operand then reads forward through wrc exactly like a forward-strand operand reads w (ovlp_match_lean).
// One thread builds one word; bases past the end of the read are zero.
__global__ void k_make_rc(const uint64_t *__restrict__ w, const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                          uint32_t n_rows, uint64_t first_word, uint64_t n_words, uint64_t *__restrict__ wrc) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_words) return;
  uint64_t word = first_word + g;
  uint32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (row_woff[mid] <= word) lo = mid; else hi = mid;
  }
  const uint64_t base = row_woff[lo];
  const int64_t len = row_len[lo];
  const int64_t a = len - (int64_t)(word - base) * 32 - 32;  // forward bases [a, a+32) feed this word, reversed
  uint64_t out = 0;
  if (a > -32) {
    uint64_t v;
    if (a >= 0) {
      const uint64_t i = base + (uint64_t)(a >> 5);
      const int s2 = (int)(a & 31) * 2;
      v = w[i] >> s2;
      if (s2) v |= w[i + 1] << (64 - s2);
      out = ~rev2(v);
    } else {
      const int valid = (int)(32 + a);  // forward bases [0, valid)
      v = w[base] << (2 * (32 - valid));
      out = ~rev2(v) & ((1ULL << (2 * valid)) - 1ULL);
    }
  }
  wrc[word] = out;
}
__global__ void k_count_nonzero_u32(const uint32_t *__restrict__ a, size_t n, unsigned int *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && a[i]) atomicAdd(out, 1u);
}

// ------------------------------------------------------------------------------------------------ read-table helpers
__global__ void k_rows_hasn(const uint32_t *__restrict__ row_rid, const uint32_t *__restrict__ hasn_by_rid, uint32_t n, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = hasn_by_rid[row_rid[i]];
}
__global__ void k_rows_to_rid_tables(const uint32_t *__restrict__ row_rid, const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                                     const uint32_t *__restrict__ row_hasn, uint32_t n, uint32_t *rlen_by_rid, uint64_t *woff_by_rid,
                                     uint32_t *hasn_by_rid) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t rid = row_rid[i];
  rlen_by_rid[rid] = row_len[i];
  woff_by_rid[rid] = row_woff[i];
  hasn_by_rid[rid] = row_hasn[i];
}
__global__ void k_max_u32(const uint32_t *__restrict__ a, uint32_t n, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t v = i < n ? a[i] : 0;
  v = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) atomicMax(out, v);
}

// ------------------------------------------------------------------------------------------------ sketch (exact automaton)
// Thread per read of `list` (row indices).  WRITE=false: cnt_by_row[row] = number of minimizers.  WRITE=true: write at
// out + off_by_row[row].
template <bool WRITE>
__global__ void k_sketch_exact(const uint64_t *__restrict__ w, const uint32_t *__restrict__ nm,
                               const uint32_t *__restrict__ list, uint32_t n_list, const uint32_t *__restrict__ row_rid,
                               const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                               const uint32_t *__restrict__ hasn_by_rid, int wsz, int k, uint32_t *__restrict__ cnt_by_row,
                               const uint64_t *__restrict__ off_by_row, mm128 *__restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_list) return;
  uint32_t row = list[t];
  uint32_t rid = row_rid[row];
  int len = (int)row_len[row];
  uint64_t ring_x[256];
  uint32_t ring_p[256];
  uint32_t n = 0;
  mm128 *dst = WRITE ? out + off_by_row[row] : nullptr;
  if (len > 0) {
    sketch_exact(w, hasn_by_rid[rid] ? nm : nullptr, row_woff[row], len, wsz, k, rid, ring_x, ring_p,
                 [&](uint64_t x, uint64_t y) {
                   if (WRITE) {
                     dst[n].x = x;
                     dst[n].y = y;
                   }
                   n++;
                 });
  }
  if (!WRITE) cnt_by_row[row] = n;
}

// Segment-parallel form of the exact automaton for the reads the tiled kernel hands back: one thread replays one
// SEG-position segment of one read (warm-up before it, w slots after it; see sketch_exact_range).
// MODE 0: count per segment.  MODE 1: write at the final place (needs the counts' scan).  MODE 2: count AND stage the records
// in a per-segment buffer of `stage_cap` records (one pass of the automaton instead of two; a segment that overflows its
// buffer raises *overflow and the caller falls back to MODE 1).
// A segment is a "piece" [seg_lo, seg_hi) of a read.  seg_kind 0: redone by the automaton (this kernel).  seg_kind 1 (reads of
// which only some strips are redone, sketch_strip.cuh): the fast path's records with positions in the piece stand; this kernel
// skips such pieces, k_seg_fast_count counts them and k_seg_place copies them.
template <int MODE>
__global__ void k_sketch_exact_seg(const uint64_t *__restrict__ w, const uint32_t *__restrict__ nm, const uint32_t *__restrict__ seg_row,
                                   const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi, const uint8_t *__restrict__ seg_kind,
                                   const uint32_t *__restrict__ seg_first, uint32_t n_seg, const uint32_t *__restrict__ row_rid, const uint32_t *__restrict__ row_len,
                                   const uint64_t *__restrict__ row_woff, const uint32_t *__restrict__ hasn_by_rid, int wsz, int k,
                                   uint32_t *__restrict__ seg_cnt, const uint32_t *__restrict__ seg_pos,
                                   const uint64_t *__restrict__ off_by_row, mm128 *__restrict__ out, uint32_t stage_cap, int *overflow,
                                   uint32_t lane_stride) {
  // Only every lane_stride-th lane takes a piece.  The lanes of a warp rescan their rings at different positions (whenever a
  // lane's minimum leaves its window), and a rescan is a chain of ~2w dependent shared-memory loads that every other lane of the
  // warp waits for: with few pieces (the usual case: a handful of bad strips) one piece per warp is 10x faster than 32.
  if (threadIdx.x % lane_stride) return;
  const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / lane_stride;
  if (s >= n_seg || seg_kind[s]) return;
  const uint32_t row = seg_row[s];
  const uint32_t rid = row_rid[row];
  const int len = (int)row_len[row];
  const int lo = (int)seg_lo[s], hi = (int)seg_hi[s];
  // the w-slot rings of the CTA's threads, interleaved in shared memory (slot j of thread t at [j * blockDim.x + t]: conflict
  // free).  The automaton rescans its ring whenever the minimum leaves the window (every ~w/2 positions per thread, so in
  // nearly every step of a warp); with the rings in local memory that rescan ran at L2 latency and the kernel took 1.5 ms for
  // the 0.4 % of the reads that need it
  extern __shared__ uint64_t seg_smem[];
  const int rs = (int)(blockDim.x / lane_stride);
  uint64_t *ring_x = seg_smem + threadIdx.x / lane_stride;
  uint32_t *ring_p = (uint32_t *)(seg_smem + (size_t)wsz * rs) + threadIdx.x / lane_stride;
  uint32_t n = 0;
  mm128 *dst = MODE == 1 ? out + off_by_row[row] + (seg_pos[s] - seg_pos[seg_first[s]]) : (MODE == 2 ? out + (size_t)s * stage_cap : nullptr);
  auto em = [&](uint64_t x, uint64_t y) {
    if (MODE == 1 || (MODE == 2 && n < stage_cap)) {
      dst[n].x = x;
      dst[n].y = y;
    }
    n++;
  };
  const uint32_t *nmp = hasn_by_rid[rid] ? nm : nullptr;
  int st = lo - sketch_warmup_len(wsz, k);
  if (st < 0) st = 0;
  if (!sketch_exact_range(w, nmp, row_woff[row], len, wsz, k, rid, st, lo, hi, ring_x, ring_p, em, rs)) {
    n = 0;  // warm-up too short (palindrome-dense stretch): replay from the read start
    sketch_exact_range(w, nmp, row_woff[row], len, wsz, k, rid, 0, lo, hi, ring_x, ring_p, em, rs);
  }
  if (MODE != 1) seg_cnt[s] = n;
  if (MODE == 2 && n > stage_cap) atomicOr(overflow, 1);
}
// pieces of kind 1: rank range of the read's fast-path records (position order in its slab) with seg_lo <= position < seg_hi
__global__ void k_seg_fast_count(const uint32_t *__restrict__ seg_row, const uint32_t *__restrict__ seg_lo, const uint32_t *__restrict__ seg_hi,
                                 const uint8_t *__restrict__ seg_kind, uint32_t n_seg, const uint64_t *__restrict__ tmp_off, const mm128 *__restrict__ tmp,
                                 const uint32_t *__restrict__ fast_cnt, uint32_t *__restrict__ seg_cnt, uint32_t *__restrict__ seg_src) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_seg || !seg_kind[s]) return;
  const uint32_t row = seg_row[s], n = fast_cnt[row];
  const mm128 *a = tmp + tmp_off[row];
  auto rank = [&](uint32_t pos) -> uint32_t {  // records with position < pos
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if ((((uint32_t)a[mid].y) >> 1) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const uint32_t r0 = rank(seg_lo[s]), r1 = rank(seg_hi[s]);
  seg_src[s] = r0;
  seg_cnt[s] = r1 - r0;
}
// records of a piece -> their final place: staged automaton records (kind 0, unless they were written in place: !place_exact) or
// the fast path's records of that position range (kind 1)
__global__ void k_seg_place(const uint32_t *__restrict__ seg_row, const uint32_t *__restrict__ seg_first, const uint8_t *__restrict__ seg_kind,
                            const uint32_t *__restrict__ seg_src, uint32_t n_seg, const uint32_t *__restrict__ seg_pos,
                            const uint64_t *__restrict__ off_by_row, const mm128 *__restrict__ stage, uint32_t stage_cap,
                            const uint64_t *__restrict__ tmp_off, const mm128 *__restrict__ tmp, mm128 *__restrict__ out, int place_exact) {
  const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, l = threadIdx.x & 7;  // 8 lanes per piece
  if (s >= n_seg) return;
  const uint32_t kind = seg_kind[s];
  if (!kind && !place_exact) return;
  const uint32_t n = seg_pos[s + 1] - seg_pos[s];
  const uint32_t row = seg_row[s];
  mm128 *dst = out + off_by_row[row] + (seg_pos[s] - seg_pos[seg_first[s]]);
  const mm128 *src = kind ? tmp + tmp_off[row] + seg_src[s] : stage + (size_t)s * stage_cap;
  for (uint32_t i = l; i < n; i += 8) dst[i] = src[i];
}
__global__ void k_seg_row_counts(const uint32_t *__restrict__ list, const uint32_t *__restrict__ list_first_seg, uint32_t n_list,
                                 const uint32_t *__restrict__ seg_pos, uint32_t *cnt_by_row) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_list) return;
  cnt_by_row[list[i]] = seg_pos[list_first_seg[i + 1]] - seg_pos[list_first_seg[i]];
}

// ------------------------------------------------------------------------------------------------ sketch (tiled fast path)
// exclusive block scan of one u32 per thread (256 threads); returns the prefix, *total gets the block sum
__device__ __forceinline__ uint32_t block_exscan_256(uint32_t v, uint32_t *scratch /* >= 9 u32 */, uint32_t *total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  if (lane == 31) scratch[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t t = lane < 8 ? scratch[lane] : 0;
#pragma unroll
    for (int d = 1; d < 8; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, t, d);
      if (lane >= d) t += y;
    }
    if (lane < 8) scratch[lane] = t;  // inclusive warp totals
  }
  __syncthreads();
  uint32_t base = wid ? scratch[wid - 1] : 0;
  *total = scratch[7];
  __syncthreads();  // scratch may be reused by the caller
  return base + x - v;
}

// Per-tile descriptor, written once per pgb_index call by k_tile_desc (one thread per row): the sketch CTAs find their read
// with ONE broadcast load instead of a 17-step binary search by thread 0 behind a barrier (31 % of the kernel's warp time
// was spent waiting at that barrier, profiles/r1g_ncu.md).
struct TileDesc { uint64_t word_off; uint32_t len, rid, j, row; };
__global__ void k_tile_desc(const uint32_t *__restrict__ tile_off, uint32_t n_rows, const uint32_t *__restrict__ row_rid,
                            const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff, TileDesc *desc) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  TileDesc d;
  d.word_off = row_woff[row]; d.len = row_len[row]; d.rid = row_rid[row]; d.row = row;
  const uint32_t t0 = tile_off[row], t1 = tile_off[row + 1];
  for (uint32_t t = t0; t < t1; t++) { d.j = t - t0; desc[t] = d; }
}

// One CTA = one tile of one read.
// Output: tile_cnt[tile] records in tmp[tile * SK_CAP ...] (position order); row_flags[row] |= reason when the read must be
// redone by the exact automaton (the tile then reports 0 records).
template <class HT>
__global__ void __launch_bounds__(SK_THREADS) k_sketch_tiled(const uint64_t *__restrict__ w, const TileDesc *__restrict__ desc,
                                                             const uint32_t *__restrict__ hasn_by_rid,
                                                             int wsz, int k, uint32_t *__restrict__ tile_cnt, uint32_t *row_flags,
                                                             mm128 *__restrict__ tmp, uint32_t tile_cap, uint32_t tile_base) {
  extern __shared__ __align__(16) unsigned char sk_smem[];
  SkTile<HT> sh;
  sk_tile_layout<HT>(sh, sk_smem, wsz);
  const uint32_t tile = tile_base + blockIdx.x;
  const int tid = threadIdx.x;
  const TileDesc td = desc[tile];  // same address for the whole CTA: one broadcast transaction
  if (tid >= 32 && tid < 36) sh.ctr[tid - 32] = 0;
  __syncthreads();
  const uint32_t row = td.row;
  const int j = (int)td.j;
  SkParams p;
  p.w = w; p.word_off = td.word_off; p.len = (int)td.len; p.rid = td.rid; p.wsz = wsz; p.k = k;
  const int H = sk_halo(wsz), TILE = sk_tile_len(wsz);
  p.r0 = j * TILE - H; p.first_tile = j == 0;
  if (hasn_by_rid[p.rid] || p.len < sk_min_len(wsz, k)) {  // block-uniform
    if (tid == 0) {
      atomicOr(&row_flags[row], hasn_by_rid[p.rid] ? (uint32_t)SK_FLAG_N : (uint32_t)SK_FLAG_SHORT);
      tile_cnt[tile] = 0;
    }
    return;
  }
  uint32_t slot_mask, np, hs;
  {
    HT hv[SK_G];
    uint16_t ps[SK_G];
    sk_phase1<HT>(tid, p, H, hv, ps, &slot_mask, &np, &hs);
    if (np) atomicAdd(&sh.ctr[SK_N_PAL], np);
    if (hs) atomicAdd(&sh.ctr[SK_N_HALO], hs);
    uint32_t total = 0;
    const uint32_t slot_base = block_exscan_256((uint32_t)__popc(slot_mask), sh.scan, &total);
    if (tid == 0) sh.ctr[SK_N_SLOTS] = total;
    sk_phase2_write<HT>(sh, hv, ps, slot_mask, slot_base);
  }
  __syncthreads();
  if (sh.ctr[SK_N_PAL] > SK_PALPAD) {
    if (tid == 0) { atomicOr(&row_flags[row], (uint32_t)SK_FLAG_PAL); tile_cnt[tile] = 0; }
    return;
  }
  // one thread per block of B slots (at most SK_R / 9 + 1 < 2 * SK_THREADS blocks)
  const int n_blocks = ((int)sh.ctr[SK_N_SLOTS] + sh.B - 1) / sh.B;
  for (int b = tid; b < n_blocks; b += SK_THREADS) sk_phase3_suffix<HT>(b, sh);
  __syncthreads();
  for (int b = tid; b < n_blocks; b += SK_THREADS) sk_phase3_prefix<HT>(b, sh);
  __syncthreads();
  int s_eval, s_emit, s_first_full;
  sk_ranges(p, sh.ctr[SK_N_HALO], &s_eval, &s_emit, &s_first_full);
  uint32_t emit_mask, tie;
  const uint32_t cnt = sk_phase4<HT>(tid, sh, wsz, s_eval, s_emit, s_first_full, &emit_mask, &tie);
  if (tie) atomicOr(&sh.ctr[SK_FLAGS], (uint32_t)SK_FLAG_TIE);
  uint32_t total = 0;
  const uint32_t pre = block_exscan_256(cnt, sh.scan, &total);  // its barriers also publish the tie flag
  if (sh.ctr[SK_FLAGS]) {
    if (tid == 0) { atomicOr(&row_flags[row], sh.ctr[SK_FLAGS]); tile_cnt[tile] = 0; }
    return;
  }
  if (total > tile_cap) {
    if (tid == 0) { atomicOr(&row_flags[row], (uint32_t)SK_FLAG_OVERFLOW); tile_cnt[tile] = 0; }
    return;
  }
  if (cnt) sk_phase5_write<HT>(tid, sh, p, emit_mask, tmp + (size_t)tile * tile_cap + pre);
  if (tid == 0) tile_cnt[tile] = total;
}

// ---- strip kernel plumbing: record budget per row, flagged rows, final placement
__global__ void k_row_caps(const uint32_t *__restrict__ row_len, uint32_t n_rows, int wsz, uint32_t *caps) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < n_rows) caps[row] = (uint32_t)((6ull * row_len[row]) / (uint32_t)(wsz + 1)) + 64u;  // 3x the expected 2/(w+1) density
  if (row == n_rows) caps[row] = 0;
}
__global__ void k_row_exact_flags(const uint32_t *__restrict__ row_flags, uint32_t n_rows, uint32_t *exact_flag) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row < n_rows) exact_flag[row] = row_flags[row] != 0;
}
// one warp per row: its records from the per-row budget area to their final place
__global__ void k_row_gather(const uint32_t *__restrict__ cnt_by_row, const uint32_t *__restrict__ row_flags, uint32_t n_rows,
                             const uint64_t *__restrict__ tmp_off, const mm128 *__restrict__ tmp, const uint64_t *__restrict__ off_by_row,
                             mm128 *__restrict__ out) {
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n_rows || row_flags[row]) return;
  const uint32_t n = cnt_by_row[row];
  const mm128 *src = tmp + tmp_off[row];
  mm128 *dst = out + off_by_row[row];
  for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

// per row: minimizer count from its tiles, or mark it for the exact automaton
__global__ void k_row_counts(const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ row_flags,
                             uint32_t n_rows, uint32_t *cnt_by_row, uint32_t *exact_flag) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  if (row_flags[row]) { cnt_by_row[row] = 0; exact_flag[row] = 1; return; }
  uint32_t n = 0;
  for (uint32_t t = tile_off[row]; t < tile_off[row + 1]; t++) n += tile_cnt[t];
  cnt_by_row[row] = n;
  exact_flag[row] = 0;
}
// 64 threads per tile copy its records to their final place
__global__ void k_tile_gather(const uint32_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_cnt, const uint32_t *__restrict__ row_flags,
                              uint32_t n_rows, uint32_t n_tiles, const uint64_t *__restrict__ off_by_row, const mm128 *__restrict__ tmp,
                              uint32_t tile_cap, mm128 *__restrict__ out) {
  const uint32_t tile = blockIdx.x * (blockDim.x / 64) + threadIdx.x / 64;
  if (tile >= n_tiles) return;
  uint32_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  const uint32_t row = lo;
  if (row_flags[row]) return;
  uint64_t at = off_by_row[row];
  for (uint32_t t = tile_off[row]; t < tile; t++) at += tile_cnt[t];
  const uint32_t n = tile_cnt[tile];
  const mm128 *src = tmp + (size_t)tile * tile_cap;
  for (uint32_t i = threadIdx.x % 64; i < n; i += 64) out[at + i] = src[i];
}

// ------------------------------------------------------------------------------------------------ reduce
// Thread per read: in-run offsets [in_off[t], in_off[t+1]) of the input level.
template <bool WRITE>
__global__ void k_reduce(const mm128 *__restrict__ in, const uint64_t *__restrict__ in_off, uint32_t n_sel, uint32_t rs,
                         uint32_t *__restrict__ counts, const uint64_t *__restrict__ out_off, mm128 *__restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sel) return;
  const mm128 *a = in + in_off[t];
  uint32_t n_in = (uint32_t)(in_off[t + 1] - in_off[t]);
  uint32_t n = 0;
  mm128 *dst = WRITE ? out + out_off[t] : nullptr;
  uint64_t last_y = ~0ULL;  // a pick's y carries this read's rid, so the cross-read carry of shmr_reduce.c:83 can never match
  for (uint32_t o = rs - 1; o < n_in; o++) {
    uint32_t p = reduce_pick(a, o, rs);
    mm128 m = a[p];
    if (m.y != last_y) {
      if (WRITE) dst[n] = m;
      n++;
      last_y = m.y;
    }
  }
  if (!WRITE) counts[t] = n;
}

// Warp per read (the default; 0.93 ms per step against 2.03 ms for k_reduce on the bench workload): lane l decides the window ending at
// in-read offset base + l.  A window emits its pick iff the pick's y differs from the previous window's pick (the first
// window always emits) - the per-element formulation of shmr_reduce.c:79-88 that tests/hostsim checks against the reference
// (sim_reduce).  Thread-per-read keeps only ~5 warps per SM busy (100 k reads), each walking ~370 windows sequentially.
template <bool WRITE>
__global__ void k_reduce_warp(const mm128 *__restrict__ in, const uint64_t *__restrict__ in_off, uint32_t n_sel, uint32_t rs,
                              uint32_t *__restrict__ counts, const uint64_t *__restrict__ out_off, mm128 *__restrict__ out) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_sel) return;  // warp-uniform
  const mm128 *a = in + in_off[t];
  const uint32_t n_in = (uint32_t)(in_off[t + 1] - in_off[t]);
  mm128 *dst = WRITE ? out + out_off[t] : nullptr;
  uint32_t n = 0;
  uint64_t carry_y = ~0ULL;  // y of the previous window's pick (none before the first window; a real y carries this read's rid)
  for (uint32_t base = rs - 1; base < n_in; base += 32) {
    const uint32_t o = base + lane;
    const bool valid = o < n_in;
    mm128 m;
    m.x = 0; m.y = ~0ULL;
    if (valid) m = a[reduce_pick(a, o, rs)];
    uint64_t prev_y = __shfl_up_sync(0xffffffffu, m.y, 1);
    if (lane == 0) prev_y = carry_y;
    const bool emit = valid && m.y != prev_y;
    const uint32_t bal = __ballot_sync(0xffffffffu, emit);
    if (WRITE && emit) dst[n + __popc(bal & ((1u << lane) - 1u))] = m;
    n += __popc(bal);
    carry_y = __shfl_sync(0xffffffffu, m.y, 31);
  }
  if (!WRITE && lane == 0) counts[t] = n;
}

// ------------------------------------------------------------------------------------------------ multiplicity table
__global__ void k_mc_insert(const mm128 *__restrict__ mm, size_t n, uint64_t *keys, uint32_t *vals, uint32_t mask, int *err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = ht_insert(keys, mask, mm[i].x >> 8);
  if (s == PGB_NOSLOT) { atomicOr(err, 4); return; }
  atomicAdd(&vals[s], 1u);
}
struct mc_entry { uint64_t mer; uint32_t count; uint32_t pad; };
__global__ void k_mc_add(const mc_entry *__restrict__ mc, size_t n, uint64_t *keys, uint32_t *vals, uint32_t mask, int *err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = ht_insert(keys, mask, mc[i].mer);
  if (s == PGB_NOSLOT) { atomicOr(err, 4); return; }
  atomicAdd(&vals[s], mc[i].count);
}
__global__ void k_mc_flags(const uint64_t *__restrict__ keys, size_t cap, uint32_t *flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) flags[i] = keys[i] != PGB_EMPTY;
}
__global__ void k_mc_dump(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, const uint32_t *__restrict__ pos,
                          size_t cap, mc_entry *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || keys[i] == PGB_EMPTY) return;
  mc_entry e;
  e.mer = keys[i];
  e.count = vals[i];
  e.pad = 0;
  out[pos[i]] = e;
}

// ------------------------------------------------------------------------------------------------ build_map
__global__ void k_count_lookup(const mm128 *__restrict__ mm, size_t n, const uint64_t *__restrict__ keys,
                               const uint32_t *__restrict__ vals, uint32_t mask, uint32_t *cnt, uint32_t lower, uint32_t upper,
                               unsigned long long *first_strict, int *err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = ht_find(keys, mask, mm[i].x >> 8);
  if (s == PGB_NOSLOT) {  // the reference asserts here (src/shmr_utils.c:314)
    atomicOr(err, 8);
    cnt[i] = 0;
    return;
  }
  uint32_t c = vals[s];
  cnt[i] = c;
  // src/shmr_utils.c:318.  ~90 % of the elements qualify: read the current minimum first, or millions of atomics on ONE address
  // serialise (this kernel took 2.25 ms for 3.7 M lookups under ncu, profiles/r1g_ncu.md); the result is the same minimum
  if (c >= lower && c < upper && (unsigned long long)i < *(volatile unsigned long long *)first_strict) atomicMin(first_strict, (unsigned long long)i);
}
__global__ void k_kept_flags(const uint32_t *__restrict__ cnt, size_t n, uint32_t lower, uint32_t upper,
                             const unsigned long long *first_strict, uint32_t *flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long s = *first_strict;
  uint32_t c = cnt[i];
  flags[i] = (i == s) || (i > s && !(c < lower || c > upper));  // src/shmr_utils.c:327
}
__global__ void k_compact_idx(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos, size_t n, uint32_t *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flags[i]) out[pos[i]] = (uint32_t)i;
}
// pair t = (kept[t], kept[t+1]); n_rec[t] = number of records it contributes to chunk c of T
__global__ void k_pair_count(const mm128 *__restrict__ mm, const uint32_t *__restrict__ kept, uint32_t n_kept, uint32_t T,
                             uint32_t c, uint32_t *n_rec) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n_kept) { if (t < n_kept) n_rec[t] = 0; return; }
  mm128 m0 = mm[kept[t]], m1 = mm[kept[t + 1]];
  uint32_t r = 0;
  if ((m0.y >> 32) == (m1.y >> 32) && pair_far_enough(m0.y, m1.y)) {
    if ((m0.x >> 8) % T == c % T) r++;
    if ((m1.x >> 8) % T == c % T) r++;
  }
  n_rec[t] = r;
}
struct PairSoA {
  uint64_t *k0, *k1, *y0, *y1;
  uint32_t *seq;
  uint8_t *dir;
};
__global__ void k_pair_write(const mm128 *__restrict__ mm, const uint32_t *__restrict__ kept, uint32_t n_kept, uint32_t T,
                             uint32_t c, const uint32_t *__restrict__ rec_off, const uint32_t *__restrict__ rlen_by_rid,
                             PairSoA o) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n_kept) return;
  mm128 m0 = mm[kept[t]], m1 = mm[kept[t + 1]];
  if ((m0.y >> 32) != (m1.y >> 32) || !pair_far_enough(m0.y, m1.y)) return;
  uint32_t at = rec_off[t];
  if ((m0.x >> 8) % T == c % T) {  // src/shmr_utils.c:337-359
    o.k0[at] = m0.x; o.k1[at] = m1.x; o.y0[at] = m0.y; o.y1[at] = m1.y; o.seq[at] = 2 * t; o.dir[at] = 0;
    at++;
  }
  if ((m1.x >> 8) % T == c % T) {  // src/shmr_utils.c:362-400
    uint32_t rl = rlen_by_rid[(uint32_t)(m0.y >> 32)];
    o.k0[at] = m1.x; o.k1[at] = m0.x; o.y0[at] = rev_y(m1.y, m1.x, rl); o.y1[at] = rev_y(m0.y, m0.x, rl);
    o.seq[at] = 2 * t + 1; o.dir[at] = 1;
  }
}

// ---- routed form (multi-GPU): a rank scans only ITS OWN shimmer list and emits the records of EVERY hash chunk, each tagged
// with the chunk that owns it (src/shmr_utils.c:337,362: chunk c of T owns hashes with (x>>8) % T == c % T, so residue v
// belongs to chunk v, residue 0 to chunk T).  Wire format: 40 bytes {x0, x1, y0, y1, direction} = mp256_t (shimmer.h:123-126).
struct route_rec { uint64_t k0, k1, y0, y1, dir; };
__device__ __forceinline__ uint32_t route_dest(uint64_t x, uint32_t T) {  // 0-based owner chunk index (chunk id - 1)
  uint32_t v = (uint32_t)((x >> 8) % T);
  return v ? v - 1 : T - 1;
}
// kept flags when the global first element (src/shmr_utils.c:311-321) lies in an EARLIER rank's list: every element uses
// the non-strict bound
__global__ void k_kept_flags_nonfirst(const uint32_t *__restrict__ cnt, size_t n, uint32_t lower, uint32_t upper, uint32_t *flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c = cnt[i];
  flags[i] = !(c < lower || c > upper);
}
__global__ void k_pair_count_all(const mm128 *__restrict__ mm, const uint32_t *__restrict__ kept, uint32_t n_kept, uint32_t *n_rec) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n_kept) { if (t < n_kept) n_rec[t] = 0; return; }
  mm128 m0 = mm[kept[t]], m1 = mm[kept[t + 1]];
  n_rec[t] = ((m0.y >> 32) == (m1.y >> 32) && pair_far_enough(m0.y, m1.y)) ? 2u : 0u;
}
__global__ void k_pair_write_all(const mm128 *__restrict__ mm, const uint32_t *__restrict__ kept, uint32_t n_kept, uint32_t T,
                                 const uint32_t *__restrict__ rec_off, const uint32_t *__restrict__ rlen_by_rid, route_rec *out,
                                 uint32_t *dest, uint32_t *idx, uint32_t *per_dest) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t + 1 >= n_kept) return;
  mm128 m0 = mm[kept[t]], m1 = mm[kept[t + 1]];
  if ((m0.y >> 32) != (m1.y >> 32) || !pair_far_enough(m0.y, m1.y)) return;
  uint32_t at = rec_off[t];
  uint32_t rl = rlen_by_rid[(uint32_t)(m0.y >> 32)];
  route_rec f, r;
  f.k0 = m0.x; f.k1 = m1.x; f.y0 = m0.y; f.y1 = m1.y; f.dir = 0;
  r.k0 = m1.x; r.k1 = m0.x; r.y0 = rev_y(m1.y, m1.x, rl); r.y1 = rev_y(m0.y, m0.x, rl); r.dir = 1;
  out[at] = f; out[at + 1] = r;
  uint32_t d0 = route_dest(m0.x, T), d1 = route_dest(m1.x, T);
  dest[at] = d0; dest[at + 1] = d1;
  idx[at] = at; idx[at + 1] = at + 1;
  atomicAdd(&per_dest[d0], 1u);
  atomicAdd(&per_dest[d1], 1u);
}
__global__ void k_route_gather(const route_rec *__restrict__ in, const uint32_t *__restrict__ perm, uint32_t n, route_rec *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[perm[i]];
}
// received records (source ranks concatenated in rank order = insertion order of the owning chunk) -> the SoA build_map form
__global__ void k_route_unpack(const route_rec *__restrict__ in, uint32_t n, PairSoA o) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  route_rec r = in[i];
  o.k0[i] = r.k0; o.k1[i] = r.k1; o.y0[i] = r.y0; o.y1[i] = r.y1; o.seq[i] = i; o.dir[i] = (uint8_t)r.dir;
}

// ------------------------------------------------------------------------------------------------ bucket tables
// X table: distinct full x values -> slot (a dense id).  B table: (slot(x0)<<32 | slot(x1)) -> bucket.
__global__ void k_bucket_insert(PairSoA r, uint32_t n_rec, uint64_t *xkeys, uint32_t xmask, uint64_t *bkeys, uint32_t bmask,
                                uint32_t *bcount, uint32_t *bfirst, uint32_t *blast, uint32_t *xfirst, uint32_t *rec_bucket, int *err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  uint32_t s0 = ht_insert(xkeys, xmask, r.k0[i]);
  uint32_t s1 = ht_insert(xkeys, xmask, r.k1[i]);
  if (s0 == PGB_NOSLOT || s1 == PGB_NOSLOT) { atomicOr(err, 16); return; }
  uint32_t b = ht_insert(bkeys, bmask, ((uint64_t)s0 << 32) | s1);
  if (b == PGB_NOSLOT) { atomicOr(err, 16); return; }
  atomicAdd(&bcount[b], 1u);
  atomicMin(&bfirst[b], r.seq[i]);
  atomicMax(&blast[b], r.seq[i]);
  atomicMin(&xfirst[s0], r.seq[i]);  // first put of this x as an OUTER key (kh_put(MMER0, ..), src/shmr_utils.c:338,363)
  rec_bucket[i] = b;
}
// ---- visiting order of the buckets, computed on the GPU (DESIGN.md "bucket order")
// occupied bucket slots -> compact list + their first-insertion sequence numbers (sort keys)
__global__ void k_bucket_list(const uint64_t *__restrict__ bkeys, const uint32_t *__restrict__ bfirst, const uint32_t *__restrict__ pos,
                              size_t cap, uint32_t *slot_out, uint32_t *key_out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || bkeys[i] == PGB_EMPTY) return;
  slot_out[pos[i]] = (uint32_t)i;
  key_out[pos[i]] = bfirst[i];
}
// buckets sorted by first sequence: is this bucket the one that inserted its outer key?
__global__ void k_outer_first(const uint32_t *__restrict__ sslot, uint32_t n, const uint64_t *__restrict__ bkeys,
                              const uint32_t *__restrict__ bfirst, const uint32_t *__restrict__ xfirst, uint32_t *isfirst) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t slot = sslot[i];
  uint32_t s0 = (uint32_t)(bkeys[slot] >> 32);
  isfirst[i] = bfirst[slot] == xfirst[s0];
}
// outer id = number of distinct outer keys inserted before (first-insertion order)
__global__ void k_outer_id_first(const uint32_t *__restrict__ sslot, uint32_t n, const uint64_t *__restrict__ bkeys,
                                 const uint32_t *__restrict__ isfirst, const uint32_t *__restrict__ firstpos, uint32_t *xid) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !isfirst[i]) return;
  xid[(uint32_t)(bkeys[sslot[i]] >> 32)] = firstpos[i];
}
__global__ void k_outer_id_all(const uint32_t *__restrict__ sslot, uint32_t n, const uint64_t *__restrict__ bkeys,
                               const uint32_t *__restrict__ xid, uint32_t *oid, uint32_t *idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  oid[i] = xid[(uint32_t)(bkeys[sslot[i]] >> 32)];
  idx[i] = i;
}
// grouped order (stable sort by outer id): gather per-bucket attributes, mark group starts
struct GroupedBuckets {
  uint32_t *slot, *first, *last, *count, *oid;
  uint64_t *k1;
};
__global__ void k_group_gather(const uint32_t *__restrict__ gidx, const uint32_t *__restrict__ goid_sorted, uint32_t n,
                               const uint32_t *__restrict__ sslot, const uint64_t *__restrict__ bkeys, const uint64_t *__restrict__ xkeys,
                               const uint32_t *__restrict__ bfirst, const uint32_t *__restrict__ blast, const uint32_t *__restrict__ bcount,
                               GroupedBuckets g, uint32_t *goff, uint32_t n_outer, uint64_t *okey, unsigned int *last_seq_all) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t slot = sslot[gidx[j]];
  uint64_t key = bkeys[slot];
  uint32_t o = goid_sorted[j];
  g.slot[j] = slot;
  g.first[j] = bfirst[slot];
  g.last[j] = blast[slot];
  g.count[j] = bcount[slot];
  g.oid[j] = o;
  g.k1[j] = xkeys[(uint32_t)key];
  if (j == 0 || goid_sorted[j - 1] != o) {
    goff[o] = j;
    okey[o] = xkeys[(uint32_t)(key >> 32)];
  }
  if (j == n - 1) goff[n_outer] = n;
  const uint32_t ls = blast[slot];
  if (ls > *(volatile unsigned int *)last_seq_all) atomicMax(last_seq_all, ls);  // one address for every bucket: test before the atomic
}
// one thread per outer key: replay its inner khash, write each bucket's visiting position inside the group
__global__ void k_inner_order(const uint32_t *__restrict__ goff, uint32_t n_outer, GroupedBuckets g, uint32_t *ipos, uint32_t *big_list,
                              uint32_t *n_big) {
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_outer) return;
  uint32_t b = goff[o], n = goff[o + 1] - b;
  if (n == 1) { ipos[b] = 0; return; }
  if (n > KHS_MAX_KEYS) {  // replayed on the host
    big_list[atomicAdd(n_big, 1u)] = o;
    return;
  }
  uint64_t keys[KHS_MAX_KEYS];
  uint8_t rank[KHS_MAX_KEYS];
  uint32_t last = 0;
  for (uint32_t i = 0; i < n; i++) {
    keys[i] = g.k1[b + i];
    uint32_t l = g.last[b + i];
    if (l > last) last = l;
  }
  khs_order(keys, n, last > g.first[b + n - 1], rank);
  for (uint32_t i = 0; i < n; i++) ipos[b + i] = rank[i];
}
// [0] start of the last group (= the newest outer key), [1] the sequence number of its first record
__global__ void k_last_group_first(const uint32_t *__restrict__ goff, uint32_t n_outer, const uint32_t *__restrict__ gfirst, uint32_t *out) {
  const uint32_t b = goff[n_outer - 1];
  out[0] = b;
  out[1] = gfirst[b];
}
// group sizes in outer visiting order
__global__ void k_group_sizes_by_rank(const uint32_t *__restrict__ goff, const uint32_t *__restrict__ orank, uint32_t n_outer, uint32_t *size_by_rank) {
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_outer) return;
  size_by_rank[orank[o]] = goff[o + 1] - goff[o];
}
// place every bucket at its visiting position; eligibility per src/shmr_overlap.c:216
__global__ void k_visit_place(GroupedBuckets g, uint32_t n, const uint32_t *__restrict__ orank, const uint32_t *__restrict__ vstart,
                              const uint32_t *__restrict__ ipos, uint32_t ovlp_upper, uint32_t *vis_slot, uint32_t *vis_elig, uint32_t *vis_cnt,
                              unsigned long long *n_cand) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long cand = 0;
  if (j < n) {
    uint32_t v = vstart[orank[g.oid[j]]] + ipos[j];
    uint32_t c = g.count[j];
    bool el = !(c <= 2 || c > ovlp_upper);
    vis_slot[v] = g.slot[j];
    vis_elig[v] = el;
    vis_cnt[v] = el ? c : 0;
    if (el) cand = (unsigned long long)c * (c - 1) / 2;
  }
  // candidate-pair statistic: one atomic per warp, not one per bucket on a single address (blockDim is a multiple of 32)
  for (int d = 16; d; d >>= 1) cand += __shfl_down_sync(0xffffffffu, cand, d);
  if ((threadIdx.x & 31) == 0 && cand) atomicAdd(n_cand, cand);
}
__global__ void k_visit_rank(const uint32_t *__restrict__ vis_slot, const uint32_t *__restrict__ vis_elig, const uint32_t *__restrict__ rank_of,
                             const uint32_t *__restrict__ off_of, uint32_t n, uint32_t *slot2rank, uint32_t *rank_off) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  if (vis_elig[v]) {
    slot2rank[vis_slot[v]] = rank_of[v];
    rank_off[rank_of[v]] = off_of[v];
  }
}

// records of eligible buckets -> rank-ordered arrays (arbitrary order inside the bucket; sorted next)
__global__ void k_scatter(PairSoA r, uint32_t n_rec, const uint32_t *__restrict__ rec_bucket, const uint32_t *__restrict__ slot2rank,
                          const uint32_t *__restrict__ rank_off, uint32_t *fill, uint64_t *sy0, uint64_t *sy1, uint32_t *sseq,
                          uint8_t *sdir) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  uint32_t rank = slot2rank[rec_bucket[i]];
  if (rank == PGB_NOSLOT) return;
  uint32_t at = rank_off[rank] + atomicAdd(&fill[rank], 1u);
  sy0[at] = r.y0[i]; sy1[at] = r.y1[i]; sseq[at] = r.seq[i]; sdir[at] = r.dir[i];
}
// glibc qsort with the boolean comparator of src/shmr_overlap.c:46-50 behaves as a stable sort, descending by position
// (SURVEY a-9): total order here = (position desc, insertion sequence asc).
__global__ void k_sort_buckets(uint32_t n_ranks, const uint32_t *__restrict__ rank_off, uint64_t *sy0, uint64_t *sy1, uint32_t *sseq,
                               uint8_t *sdir) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_ranks) return;
  uint32_t b = rank_off[r], n = rank_off[r + 1] - b;
  for (uint32_t i = 1; i < n; i++) {
    uint64_t y0 = sy0[b + i], y1 = sy1[b + i];
    uint32_t sq = sseq[b + i];
    uint8_t d = sdir[b + i];
    uint32_t p = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1);
    uint32_t j = i;
    while (j > 0) {
      uint64_t yj = sy0[b + j - 1];
      uint32_t pj = (uint32_t)((yj & 0xFFFFFFFFULL) >> 1);
      bool before = (pj > p) || (pj == p && sseq[b + j - 1] < sq);  // element j-1 stays in front
      if (before) break;
      sy0[b + j] = yj; sy1[b + j] = sy1[b + j - 1]; sseq[b + j] = sseq[b + j - 1]; sdir[b + j] = sdir[b + j - 1];
      j--;
    }
    sy0[b + j] = y0; sy1[b + j] = y1; sseq[b + j] = sq; sdir[b + j] = d;
  }
}

// ------------------------------------------------------------------------------------------------ replay + align
// Tables of the fix-point.  Both are open addressing with linear probing over a capacity that is not a
// power of two (slot = mulhi(mix(key), cap)), sized close to what a chunk needs so that they stay largely L2-resident:
//  * time-stamped rid_pairs E: 16-byte entries {read-id pair, value of iteration 0, value of iteration 1}; value =
//    rank<<2 | type of the first bucket (in visiting order) that accepted the pair, PGB_VNONE = none.  One 16-byte load
//    answers "was this pair in rid_pairs when the reference reached my bucket" for both iterations.
//  * alignment cache: key rank<<32 | i<<16 | j -> match_t stored AT the key's slot (m_size == PGB_ALN_PENDING until
//    k_align* filled it in); a request carries its slot.
struct AlnReq { uint32_t rid0, start0, rid1, strands, slot; };  // strands: bit0 = strand0, bit1 = strand1
struct EEntry { uint64_t key; uint32_t v[2]; };
#define PGB_VNONE 0xFFFFFFFFu
#define PGB_ALN_PENDING ((int32_t)0x80000000)
#define PGB_MAXV 264  // band rows of ovlp_match: supports band_tolerance <= 256
struct ReplayState {
  EEntry *E; uint32_t ecap; uint32_t cur;  // v[cur] is being built in this pass, v[cur ^ 1] is the previous pass
  uint64_t *akeys; match_t *ares; uint32_t acap;
  AlnReq *reqs; uint32_t req_cap;
  uint32_t *n_req;  // device counter
  const uint32_t *rlen_by_rid;
  unsigned long long *ctr;  // [0] predicted (unknown) alignments [1] table diffs
  int *err;
};
__device__ __forceinline__ uint32_t ht_home(uint64_t key, uint32_t cap) { return __umulhi(ht_mix(key), cap); }
// Longest probe sequence of the pair table and of the alignment cache.  An insert that does not find a slot within this
// many steps reports the table as full (the host grows it and restarts the fix-point); lookups stop there too, since no
// key can live further from its home.  Without the bound a FULL table turns every operation into a scan of the whole
// table (millions of steps per thread: the N = 4 / N = 8 bench runs of round 1g never came back from their first wet pass).
#define PGB_MAX_PROBE 512u

struct DevReplayCtx {
  ReplayState s;
  uint32_t rank;
  bool request_enabled;
  ovlp_rec *out;
  __device__ uint32_t rlen(uint32_t rid) const { return s.rlen_by_rid[rid]; }
  __device__ void pair_get(uint64_t p, uint64_t *vold, uint64_t *vnew) const {
    *vold = *vnew = ~0ULL;
    uint32_t h = ht_home(p, s.ecap);
    const uint32_t lim = s.ecap < PGB_MAX_PROBE ? s.ecap : PGB_MAX_PROBE;
    for (uint32_t probe = 0; probe < lim; probe++) {
      const uint4 e = __ldcg((const uint4 *)&s.E[h]);  // L2: entries are updated with atomics by other SMs (and by this thread)
      const uint64_t k = (uint64_t)e.x | ((uint64_t)e.y << 32);
      if (k == p) {
        const uint32_t vo = s.cur ? e.z : e.w, vn = s.cur ? e.w : e.z;
        if (vo != PGB_VNONE) *vold = vo;
        if (vn != PGB_VNONE) *vnew = vn;
        return;
      }
      if (k == PGB_EMPTY) return;
      if (++h == s.ecap) h = 0;
    }
  }
  __device__ void pair_set(uint64_t p, uint64_t v) {
    uint32_t h = ht_home(p, s.ecap);
    const uint32_t lim = s.ecap < PGB_MAX_PROBE ? s.ecap : PGB_MAX_PROBE;
    for (uint32_t probe = 0; probe < lim; probe++) {
      uint64_t k = __ldcg(&s.E[h].key);
      if (k == PGB_EMPTY) {
        k = atomicCAS((unsigned long long *)&s.E[h].key, (unsigned long long)PGB_EMPTY, (unsigned long long)p);
        if (k == PGB_EMPTY) k = p;
      }
      if (k == p) {
        atomicMin(&s.E[h].v[s.cur], (uint32_t)v);
        return;
      }
      if (++h == s.ecap) h = 0;
    }
    atomicOr(s.err, 32);
  }
  __device__ uint32_t aln_slot(uint64_t key, bool insert) const {
    uint32_t h = ht_home(key, s.acap);
    const uint32_t lim = s.acap < PGB_MAX_PROBE ? s.acap : PGB_MAX_PROBE;
    for (uint32_t probe = 0; probe < lim; probe++) {
      uint64_t k = __ldcg(&s.akeys[h]);
      if (k == PGB_EMPTY) {
        if (!insert) return PGB_NOSLOT;
        k = atomicCAS((unsigned long long *)&s.akeys[h], (unsigned long long)PGB_EMPTY, (unsigned long long)key);
        if (k == PGB_EMPTY) k = key;
      }
      if (k == key) return h;
      if (++h == s.acap) h = 0;
    }
    return PGB_NOSLOT;
  }
  __device__ uint64_t aln_key(uint32_t i, uint32_t j) const { return ((uint64_t)rank << 32) | ((uint64_t)i << 16) | j; }
  __device__ bool aln_get(uint32_t i, uint32_t j, match_t *m) const {
    const uint32_t sl = aln_slot(aln_key(i, j), false);
    if (sl == PGB_NOSLOT) return false;
    const int4 a = __ldcg((const int4 *)&s.ares[sl]), b = __ldcg((const int4 *)&s.ares[sl] + 1);
    if (a.x == PGB_ALN_PENDING) return false;
    m->m_size = a.x; m->dist = a.y; m->q_bgn = a.z; m->q_end = a.w; m->t_bgn = b.x; m->t_end = b.y; m->t_m_end = b.z; m->q_m_end = b.w;
    return true;
  }
  __device__ void aln_request(uint32_t i, uint32_t j, uint32_t rid0, uint32_t start0, uint32_t s0, uint32_t rid1, uint32_t s1) {
    if (!request_enabled) return;
    const uint32_t sl = aln_slot(aln_key(i, j), true);
    if (sl == PGB_NOSLOT) { atomicOr(s.err, 64); return; }
    const uint32_t idx = atomicAdd(s.n_req, 1u);
    if (idx >= s.req_cap) { atomicOr(s.err, 64); return; }
    s.ares[sl].m_size = PGB_ALN_PENDING;
    AlnReq q;
    q.rid0 = rid0; q.start0 = start0; q.rid1 = rid1; q.strands = s0 | (s1 << 1); q.slot = sl;
    s.reqs[idx] = q;
  }
  __device__ void emit(uint32_t n, const ovlp_rec &o) { out[n] = o; }
};

// one thread = one bucket; only the n_ranks buckets named in list are replayed
__global__ void __launch_bounds__(64) k_replay(ReplayState st, uint32_t n_ranks, const uint32_t *__restrict__ list, const uint32_t *__restrict__ rank_off,
                                               const uint64_t *__restrict__ sy0, const uint8_t *__restrict__ sdir, uint8_t *contained, uint32_t bestn,
                                               int request_enabled, int do_emit, uint32_t *acc_count, const uint32_t *__restrict__ out_off,
                                               ovlp_rec *out, uint8_t *unk_flag, const uint32_t *__restrict__ cnt_dev, uint32_t warp_min, uint32_t warp_max) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (cnt_dev) n_ranks = cnt_dev[0];  // run list sized on the device (no host round trip): [0] small buckets, [1] all
  if (r >= n_ranks) return;
  if (*(volatile int *)st.err & (32 | 64)) return;  // a table filled up: this pass is void, the host restarts with larger tables
  r = list[r];
  uint32_t b = rank_off[r], n = rank_off[r + 1] - b;
  if (n >= warp_min && n < warp_max) return;  // walked by a lane group (k_replay_group)
  DevReplayCtx c;
  c.s = st;
  c.rank = r;
  c.request_enabled = request_enabled != 0;
  c.out = do_emit ? out + out_off[r] : nullptr;
  uint32_t unk = 0;
  uint32_t acc = replay_bucket(c, r, sy0 + b, sdir + b, n, contained + b, bestn, do_emit != 0, &unk);
  acc_count[r] = acc;
  unk_flag[r] = unk != 0;
  if (unk) atomicAdd(&st.ctr[0], (unsigned long long)unk);
}

// Lane-group form for MEDIUM buckets (min_n <= n < max_n).  A thread that walks such a bucket alone performs one dependent table
// probe after the other (several per row, rows n-2 .. 0): an incremental pass lasts as long as its slowest thread.  Here G lanes
// walk the rows of one bucket (32/G buckets per warp) and probe G candidates of the row IN PARALLEL; the row's sequential
// semantics (stop at bestn overlaps, a CONTAINED result ends the row, src/shmr_overlap.c:97-176) are then applied to the
// classified candidates with ballots: the candidates up to the cut-off are exactly those the reference visits, only they cause
// table updates, requests and records.  G is small on purpose: candidates probed beyond the cut-off are wasted random accesses
// (a whole warp per bucket, G = 32, doubles the time of a full pass).  Valid when no read has two records in the bucket (every
// read pair then occurs once in the bucket, so the rows do not see each other's table entries); other buckets are walked by
// the group's first lane with the generic replay_bucket.  The warp never diverges in the main loop: every ballot is full-mask
// and a group that is done idles under predicates.
#define PGB_RW_WARPS 4
#define PGB_RW_MAXN 64
template <int G>
__global__ void __launch_bounds__(PGB_RW_WARPS * 32) k_replay_group(ReplayState st, uint32_t n_list, const uint32_t *__restrict__ list,
                                                                    const uint32_t *__restrict__ rank_off, const uint64_t *__restrict__ sy0,
                                                                    const uint8_t *__restrict__ sdir, uint8_t *contained, uint32_t bestn,
                                                                    int request_enabled, int do_emit, uint32_t *acc_count,
                                                                    const uint32_t *__restrict__ out_off, ovlp_rec *out, uint8_t *unk_flag,
                                                                    const uint32_t *__restrict__ cnt_dev, uint32_t min_n, uint32_t max_n) {
  constexpr int NG = 32 / G, GPB = PGB_RW_WARPS * NG;
  constexpr uint32_t FULL = 0xffffffffu, GM = G == 32 ? FULL : ((1u << (G & 31)) - 1u);
  __shared__ uint64_t s_y0[GPB][PGB_RW_MAXN];
  __shared__ uint32_t s_rlen[GPB][PGB_RW_MAXN];
  __shared__ uint8_t s_dir[GPB][PGB_RW_MAXN], s_cont[GPB][PGB_RW_MAXN];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, g = lane / G, gl = lane % G, gsh = g * G;
  const uint32_t slot = wid * NG + g, li = blockIdx.x * GPB + slot;
  if (cnt_dev) n_list = cnt_dev[0];
  bool have = li < n_list && !(*(volatile int *)st.err & (32 | 64));  // (a table filled up: this pass is void)
  uint32_t r = 0, b = 0, n = 0;
  if (have) {
    r = list[li];
    b = rank_off[r];
    n = rank_off[r + 1] - b;
    have = n >= min_n && n < max_n;  // the others: thread-walked (k_replay) or CTA-walked (k_replay_block)
  }
  if (!__any_sync(FULL, have)) return;
  DevReplayCtx c;
  c.s = st;
  c.rank = r;
  c.request_enabled = request_enabled != 0;
  c.out = (do_emit && have) ? out + out_off[r] : nullptr;
  uint64_t *y0s = s_y0[slot];
  uint32_t *rl = s_rlen[slot];
  uint8_t *dr = s_dir[slot], *ct = s_cont[slot];
  bool dup = false;
  if (have) {
    for (uint32_t t = gl; t < n; t += G) {
      const uint64_t y = sy0[b + t];
      y0s[t] = y;
      rl[t] = st.rlen_by_rid[(uint32_t)(y >> 32)];
      dr[t] = sdir[b + t];
      ct[t] = 0;
    }
  }
  __syncwarp();
  if (have)  // a read with two records in the bucket?
    for (uint32_t t = gl; t < n; t += G) {
      const uint32_t rid = (uint32_t)(y0s[t] >> 32);
      for (uint32_t u = t + 1; u < n; u++) dup |= (uint32_t)(y0s[u] >> 32) == rid;
    }
  dup = ((__ballot_sync(FULL, dup) >> gsh) & GM) != 0;
  const uint64_t NONE = ~0ULL;
  uint32_t n_acc = 0, n_unk = 0, overlap_count = 0;
  uint32_t k0 = (have && !dup && bestn > 0) ? n - 1 : 0;  // current row is k0 - 1, its candidates start at k0; 0 = this group is done
  uint32_t base = k0;
  while (__any_sync(FULL, k0 > 0)) {
    const bool act = k0 > 0;
    const uint32_t i = act ? k0 - 1 : 0, j = base + gl;
    const bool valid = act && j < n && !ct[j];  // (rid0 != rid1: no read occurs twice in this bucket)
    bool hit = false, known = true, accepted = false;
    uint32_t type = 0, rid0 = 0, pos0 = 0, rlen0 = 0, rid1 = 0, rlen1 = 0, pos1 = 0;
    uint64_t y0 = 0, y1 = 0, ridp = 0;
    match_t m;
    if (valid) {
      y0 = y0s[i];
      rid0 = (uint32_t)(y0 >> 32); pos0 = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1) + 1; rlen0 = rl[i];
      y1 = y0s[j];
      rid1 = (uint32_t)(y1 >> 32);
      ridp = rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
      uint64_t v, vnew;
      c.pair_get(ridp, &v, &vnew);
      hit = (v != NONE) && ((uint32_t)(v >> 2) < r);
      if (!hit) {
        v = vnew;
        hit = (v != NONE) && ((uint32_t)(v >> 2) <= r);
      }
      if (hit) {
        type = (uint32_t)(v & 3);
      } else {
        pos1 = (uint32_t)((y1 & 0xFFFFFFFFULL) >> 1) + 1;
        rlen1 = rl[j];
        known = c.aln_get(i, j, &m);
        if (!known) predict_match(rlen0, rlen1, pos0 - pos1, &m);
        accepted = classify_match(m, rlen0, rlen1, rlen0 - pos0 + pos1, rlen1, &type);
      }
    }
    // the row in candidate order: overlap_count rises on a rid_pairs hit of type OVERLAP and on an accepted OVERLAP; the
    // candidate that brings it to bestn is the last one visited; an accepted CONTAINED ends the row after its candidate
    const bool inc = valid && type == OVL_OVERLAP && (hit || accepted);
    const bool ends = valid && !hit && accepted && type == OVL_CONTAINED;
    const uint32_t incm = (__ballot_sync(FULL, inc) >> gsh) & GM, endm = (__ballot_sync(FULL, ends) >> gsh) & GM;
    const uint32_t need = bestn - overlap_count;  // >= 1 for an active group
    uint32_t cut = G;                             // candidates of lanes <= cut are visited
    bool row_done = false;
    if (act && (uint32_t)__popc(incm) >= need) cut = __fns(incm, 0, (int)need);
    if (endm) {
      const uint32_t e = __ffs((int)endm) - 1;
      if (e <= cut) { cut = e; row_done = true; }
    }
    const uint32_t vis = cut >= G - 1 ? GM : ((2u << cut) - 1u);
    const bool mine = valid && ((vis >> gl) & 1u);
    const uint32_t accm = (__ballot_sync(FULL, mine && !hit && accepted) >> gsh) & GM;
    const uint32_t unkm = (__ballot_sync(FULL, mine && !hit && !known) >> gsh) & GM;
    if (mine && !hit) {
      if (!known) c.aln_request(i, j, rid0, pos0 - pos1, dr[i], rid1, dr[j]);
      if (accepted) {
        if (type == OVL_CONTAINS) ct[j] = 1;
        c.pair_set(ridp, ((uint64_t)r << 2) | type);
        if (do_emit) {
          ovlp_rec o;
          o.y0 = y0; o.y1 = y1; o.rl0 = rlen0; o.rl1 = rlen1;
          o.strand0 = dr[i]; o.strand1 = dr[j]; o.ovlp_type = (uint8_t)type; o.pad0 = 0;
          o.match = m;
          o.pad1 = 0;
          c.out[n_acc + __popc(accm & ((1u << gl) - 1u))] = o;
        }
      }
    }
    overlap_count += __popc(incm & vis);
    n_acc += __popc(accm);
    n_unk += __popc(unkm);
    const bool next_row = act && (row_done || overlap_count >= bestn || base + G >= n);
    if (act && row_done && gl == 0) ct[i] = 1;
    if (act && !next_row) base += G;
    __syncwarp();
    if (next_row) {  // the next row that is not contained (src/shmr_overlap.c:71)
      k0--;
      while (k0 > 0 && ct[k0 - 1]) k0--;
      base = k0;
      overlap_count = 0;
    }
  }
  if (have && gl == 0) {
    if (dup) {
      n_acc = replay_bucket(c, r, sy0 + b, sdir + b, n, contained + b, bestn, do_emit != 0, &n_unk);
    }
    acc_count[r] = n_acc;
    unk_flag[r] = n_unk != 0;
    if (n_unk) atomicAdd(&st.ctr[0], (unsigned long long)n_unk);
  }
}

// Block-cooperative form of the same scan for BIG buckets (one CTA of 128 threads = one bucket of up to PGB_RB_MAXN records).
// A thread walking a 120-record bucket alone performs hundreds of DEPENDENT table probes (~0.5-1 ms), which is the critical
// path of every pass, however few buckets run.  Here the probes of all n(n-1)/2 candidate pairs are issued in parallel:
//   phase 1  (all threads) every candidate (i, j): pair-table probe; if the pair is not in rid_pairs, alignment-cache probe
//            and acceptance test -> one code byte in shared memory (nothing is written to global memory);
//   phase 2  (thread 0) the reference's sequential scan (rows n-2..0, bestn, contained[], CONTAINED ends the row) over the
//            code bytes: shared-memory speed, produces the list of events (accepted / predicted candidates) in order;
//   phase 3  (all threads) the events' side effects: pair_set, alignment requests, records.
// Phase 1 sees rid_pairs as it was when the bucket started, so a read PAIR that occurs twice inside the bucket (possible
// only if some read has two records in it) would miss its own bucket's entry: such buckets, and buckets whose event list
// overflows, are scanned by thread 0 with the generic replay_bucket instead.
#define PGB_RB_THREADS 128
#define PGB_RB_MAXN 128
#define PGB_RB_MAXEV 1536
__global__ void __launch_bounds__(PGB_RB_THREADS) k_replay_block(ReplayState st, uint32_t n_ranks, const uint32_t *__restrict__ list,
                                                                 const uint32_t *__restrict__ rank_off, const uint64_t *__restrict__ sy0,
                                                                 const uint8_t *__restrict__ sdir, uint8_t *contained, uint32_t bestn,
                                                                 int request_enabled, int do_emit, uint32_t *acc_count,
                                                                 const uint32_t *__restrict__ out_off, ovlp_rec *out, uint8_t *unk_flag,
                                                                 const uint32_t *__restrict__ cnt_dev) {
  if (cnt_dev) { n_ranks = cnt_dev[1] - cnt_dev[0]; list += cnt_dev[0]; }  // the big buckets follow the small ones in the run list
  __shared__ uint8_t code[PGB_RB_MAXN * PGB_RB_MAXN];  // [i * n + j], j > i: kind | type << 2 | accepted << 4
  __shared__ uint64_t s_y0[PGB_RB_MAXN];
  __shared__ uint32_t s_rlen[PGB_RB_MAXN];
  __shared__ uint8_t s_dir[PGB_RB_MAXN], s_cont[PGB_RB_MAXN];
  __shared__ uint32_t s_ev[PGB_RB_MAXEV];  // i | j << 8 | type << 16 | accepted << 18 | known << 19
  __shared__ uint32_t s_nev, s_generic;
  if (blockIdx.x >= n_ranks) return;
  const uint32_t tid = threadIdx.x;
  __shared__ int s_abort;  // a table filled up: this pass is void (the host restarts with larger tables), finish quickly
  if (tid == 0) s_abort = *(volatile int *)st.err & (32 | 64);
  __syncthreads();
  if (s_abort) return;
  const uint32_t r = list[blockIdx.x];
  const uint32_t b = rank_off[r], n = rank_off[r + 1] - b;
  DevReplayCtx c;
  c.s = st;
  c.rank = r;
  c.request_enabled = request_enabled != 0;
  c.out = do_emit ? out + out_off[r] : nullptr;
  const uint64_t NONE = ~0ULL;
  if (tid == 0) { s_nev = 0; s_generic = n > PGB_RB_MAXN ? 1u : 0u; }
  for (uint32_t t = tid; t < n && t < PGB_RB_MAXN; t += PGB_RB_THREADS) {
    const uint64_t y = sy0[b + t];
    s_y0[t] = y;
    s_rlen[t] = st.rlen_by_rid[(uint32_t)(y >> 32)];
    s_dir[t] = sdir[b + t];
    s_cont[t] = 0;
  }
  __syncthreads();
  if (!s_generic) {
    // ---- phase 1
    for (uint32_t cell = tid; cell < n * n; cell += PGB_RB_THREADS) {
      const uint32_t i = cell / n, j = cell - i * n;
      if (j <= i) continue;
      const uint64_t y0 = s_y0[i], y1 = s_y0[j];
      const uint32_t rid0 = (uint32_t)(y0 >> 32), rid1 = (uint32_t)(y1 >> 32);
      uint32_t cd = 0;
      if (rid0 == rid1) {
        s_generic = 1;  // a read with two records in this bucket (benign race: every writer stores 1)
      } else {
        const uint64_t ridp = rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
        uint64_t v, vnew;
        c.pair_get(ridp, &v, &vnew);
        bool hit = (v != NONE) && ((uint32_t)(v >> 2) < r);
        if (!hit) {
          v = vnew;
          hit = (v != NONE) && ((uint32_t)(v >> 2) <= r);
        }
        if (hit) {
          cd = 1u | ((uint32_t)(v & 3) << 2);
        } else {
          const uint32_t pos0 = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1) + 1, pos1 = (uint32_t)((y1 & 0xFFFFFFFFULL) >> 1) + 1;
          const uint32_t rlen0 = s_rlen[i], rlen1 = s_rlen[j], start0 = pos0 - pos1;
          match_t m;
          const bool known = c.aln_get(i, j, &m);
          if (!known) predict_match(rlen0, rlen1, start0, &m);
          uint32_t type = 0;
          const bool accepted = classify_match(m, rlen0, rlen1, rlen0 - pos0 + pos1, rlen1, &type);
          cd = (known ? 2u : 3u) | (type << 2) | ((uint32_t)accepted << 4);
        }
      }
      code[cell] = (uint8_t)cd;
    }
  }
  __syncthreads();
  // ---- phase 2
  if (tid == 0) {
    uint32_t n_acc = 0, n_unk = 0, n_ev = 0;
    bool generic = s_generic != 0;
    if (!generic) {
      for (uint32_t k0 = n - 1; k0 > 0 && !generic; k0--) {
        const uint32_t i = k0 - 1;
        if (s_cont[i]) continue;
        uint32_t oc = 0;
        for (uint32_t j = i + 1; j < n && oc < bestn; j++) {
          if (s_cont[j]) continue;
          const uint32_t cd = code[i * n + j], kind = cd & 3u, type = (cd >> 2) & 3u;
          if (kind == 0) continue;
          if (kind == 1) {
            if (type == OVL_OVERLAP) oc++;
            continue;
          }
          const bool accepted = (cd >> 4) & 1u, known = kind == 2;
          if (!known) n_unk++;
          if (accepted || !known) {
            if (n_ev >= PGB_RB_MAXEV) { generic = true; break; }
            s_ev[n_ev++] = i | (j << 8) | (type << 16) | ((uint32_t)accepted << 18) | ((uint32_t)known << 19);
          }
          if (accepted) {
            if (type == OVL_CONTAINS) s_cont[j] = 1;
            else if (type == OVL_CONTAINED) s_cont[i] = 1;
            else oc++;
            n_acc++;
          }
          if (s_cont[i]) break;
        }
      }
    }
    if (generic) {  // exact sequential path (duplicate reads in the bucket, oversized bucket or event list)
      uint32_t unk = 0;
      n_acc = replay_bucket(c, r, sy0 + b, sdir + b, n, contained + b, bestn, do_emit != 0, &unk);
      n_unk = unk;
      n_ev = 0;
    }
    s_nev = n_ev;
    acc_count[r] = n_acc;
    unk_flag[r] = n_unk != 0;
    if (n_unk) atomicAdd(&st.ctr[0], (unsigned long long)n_unk);
  }
  __syncthreads();
  // ---- phase 3: side effects of the events, in parallel; the record index of an accepted event = accepted events before it
  __shared__ uint32_t s_wtot[PGB_RB_THREADS / 32];
  const uint32_t n_ev = s_nev;
  uint32_t base = 0;
  for (uint32_t e0 = 0; e0 < n_ev; e0 += PGB_RB_THREADS) {
    const uint32_t e = e0 + tid;
    uint32_t ev = 0;
    bool accepted = false;
    if (e < n_ev) { ev = s_ev[e]; accepted = (ev >> 18) & 1u; }
    const uint32_t bal = __ballot_sync(0xffffffffu, accepted);
    if ((tid & 31) == 0) s_wtot[tid >> 5] = __popc(bal);
    __syncthreads();
    uint32_t before = __popc(bal & ((1u << (tid & 31)) - 1u)), chunk_total = 0;
    for (uint32_t wv = 0; wv < PGB_RB_THREADS / 32; wv++) {
      if (wv < (tid >> 5)) before += s_wtot[wv];
      chunk_total += s_wtot[wv];
    }
    if (e < n_ev) {
      const uint32_t i = ev & 0xFF, j = (ev >> 8) & 0xFF, type = (ev >> 16) & 3u;
      const bool known = (ev >> 19) & 1u;
      const uint64_t y0 = s_y0[i], y1 = s_y0[j];
      const uint32_t rid0 = (uint32_t)(y0 >> 32), rid1 = (uint32_t)(y1 >> 32);
      const uint32_t pos0 = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1) + 1, pos1 = (uint32_t)((y1 & 0xFFFFFFFFULL) >> 1) + 1;
      if (!known) c.aln_request(i, j, rid0, pos0 - pos1, s_dir[i], rid1, s_dir[j]);
      if (accepted) {
        const uint64_t ridp = rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
        c.pair_set(ridp, ((uint64_t)r << 2) | type);
        if (do_emit) {
          ovlp_rec o;
          o.y0 = y0; o.y1 = y1; o.rl0 = s_rlen[i]; o.rl1 = s_rlen[j];
          o.strand0 = s_dir[i]; o.strand1 = s_dir[j]; o.ovlp_type = (uint8_t)type; o.pad0 = 0;
          if (!c.aln_get(i, j, &o.match)) atomicOr(st.err, 64);  // emission runs after convergence: every alignment is known
          o.pad1 = 0;
          c.out[base + before] = o;
        }
      }
    }
    __syncthreads();  // s_wtot is rewritten by the next chunk
    base += chunk_total;
  }
}

// sort key of an alignment request = predicted overlap length in 256-base units: warps then hold alignments of similar
// length and finish together
__global__ void k_align_keys(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n, const uint32_t *__restrict__ rlen_by_rid,
                             uint32_t *keys, uint32_t *idx, int mode) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AlnReq q = reqs[first + i];
  uint32_t a = rlen_by_rid[q.rid0] - q.start0, b = rlen_by_rid[q.rid1];
  uint32_t e = (a < b ? a : b) >> 8;
  const uint32_t lk = 255u - (e > 255 ? 255 : e);  // longest first: the tail of the launch is made of short alignments
  // mode m > 0 (default m = 4): alignments against the same target read side by side -- every alignment starts at base 0 of its
  // target, so the lanes of a warp walk the same cache lines of it in step (one L1 tag look-up instead of one per lane) -- with
  // the length class (2^(m-1) x 256 bases) as the major key so that warps still finish together.  mode 0: length only.
  keys[i] = mode > 0 ? ((lk >> (mode - 1)) << 23) | (q.rid1 & 0x7FFFFFu) : lk;
  idx[i] = i;
}
// 1 thread = 1 alignment, flattened state machine (ovlp_match_flat) over (read, strand) views with N masks; the result goes
// to results[request.slot].  only_n: take only the requests that involve a read with N (the rest is done by k_align_lean)
__global__ void k_align(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n, const uint32_t *__restrict__ perm,
                        const uint64_t *__restrict__ w, const uint32_t *__restrict__ nm, const uint64_t *__restrict__ woff_by_rid,
                        const uint32_t *__restrict__ rlen_by_rid, const uint32_t *__restrict__ hasn_by_rid, int bw, match_t *results,
                        int *err, unsigned long long *bases_total, int only_n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (perm) i = perm[i];
  AlnReq q = reqs[first + i];
  if (only_n && !(hasn_by_rid[q.rid0] | hasn_by_rid[q.rid1])) return;
  uint32_t rl0 = rlen_by_rid[q.rid0], rl1 = rlen_by_rid[q.rid1];
  SeqView qv = make_view(w, nm, woff_by_rid[q.rid0], rl0, q.start0, q.strands & 1, (int)hasn_by_rid[q.rid0]);
  SeqView tv = make_view(w, nm, woff_by_rid[q.rid1], rl1, 0, (q.strands >> 1) & 1, (int)hasn_by_rid[q.rid1]);
  int V[2 * PGB_MAXV];
  match_t m;
  int e = 0;
  ovlp_match_flat(qv, (int)(rl0 - q.start0), tv, (int)rl1, bw, V, PGB_MAXV, &m, &e);
  if (e) atomicOr(err, 128 | (e << 8));
  results[q.slot] = m;
  // bases compared along the final path (algorithmic-bytes accounting, SURVEY 8d: (q_end + t_end)/4 per alignment)
  atomicAdd(bases_total, (unsigned long long)(m.q_end + m.t_end));
}

// 1 thread = 1 alignment between N-free reads, forward views over the forward / reverse-complement images
// (ovlp_match_lean); requests that involve a read with N are left to k_align(only_n = 1)
#define PGB_ALIGN_THREADS 64
#ifndef PGB_ALIGN_MINBLOCKS
#define PGB_ALIGN_MINBLOCKS 16
#endif
template <bool PRE, bool TRIMREG, int MINBLOCKS>
__global__ void __launch_bounds__(PGB_ALIGN_THREADS, MINBLOCKS) k_align_lean(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n,
                                                                  const uint32_t *__restrict__ perm, const uint64_t *__restrict__ w,
                                                                  const uint64_t *__restrict__ wrc, const uint64_t *__restrict__ woff_by_rid,
                                                                  const uint32_t *__restrict__ rlen_by_rid, const uint32_t *__restrict__ hasn_by_rid,
                                                                  int bw, match_t *results, unsigned long long *bases_total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (perm) i = perm[i];
  const AlnReq q = reqs[first + i];
  if (hasn_by_rid[q.rid0] | hasn_by_rid[q.rid1]) return;
  const uint32_t rl0 = rlen_by_rid[q.rid0], rl1 = rlen_by_rid[q.rid1];
  const uint64_t *qw = ((q.strands & 1) ? wrc : w) + woff_by_rid[q.rid0];
  const uint64_t *tw = ((q.strands & 2) ? wrc : w) + woff_by_rid[q.rid1];
  int V[2 * PGB_MAXV];
  match_t m;
  ovlp_match_lean_t<PRE, TRIMREG>(qw, q.start0, (int)(rl0 - q.start0), tw, 0u, (int)rl1, bw, V, PGB_MAXV, &m);
  results[q.slot] = m;
  atomicAdd(bases_total, (unsigned long long)(m.q_end + m.t_end));
}

// Low-latency form for the small alignment batches of the fix-point's tail passes: 1 warp = 1 alignment.  Same algorithm
// (the structure of ovlp_match_core: rows d, diagonals k, snake), every lane carries the same scalar state; only the snake
// is parallel: lane l compares bases [32 l, 32 l + 32) past the snake's start, a ballot finds the first mismatch, so a
// snake of up to 1024 bases costs one round of loads instead of 32 dependent word-steps.  The band rows live in shared
// memory.  A single-thread alignment takes ~1 ms of dependent loads; a batch of a few hundred is latency-bound, and
// this form finishes it in ~0.1 ms.  Reads with N are left to k_align(only_n = 1), as in k_align_lean.
#define PGB_AW_WARPS 4
__global__ void __launch_bounds__(PGB_AW_WARPS * 32) k_align_warp(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n,
                                                                  const uint64_t *__restrict__ w, const uint64_t *__restrict__ wrc,
                                                                  const uint64_t *__restrict__ woff_by_rid, const uint32_t *__restrict__ rlen_by_rid,
                                                                  const uint32_t *__restrict__ hasn_by_rid, int bw, match_t *results,
                                                                  unsigned long long *bases_total) {
  __shared__ int sV[PGB_AW_WARPS][2 * PGB_MAXV];
  const uint32_t FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t i = blockIdx.x * PGB_AW_WARPS + wid;
  if (i >= n) return;
  const AlnReq q = reqs[first + i];
  if (hasn_by_rid[q.rid0] | hasn_by_rid[q.rid1]) return;
  const uint32_t rl0 = rlen_by_rid[q.rid0], rl1 = rlen_by_rid[q.rid1];
  const uint64_t *qw = ((q.strands & 1) ? wrc : w) + woff_by_rid[q.rid0];
  const uint64_t *tw = ((q.strands & 2) ? wrc : w) + woff_by_rid[q.rid1];
  const uint32_t qo = q.start0, to = 0;
  const int q_len = (int)(rl0 - q.start0), t_len = (int)rl1;
  match_t r;
  r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = r.t_m_end = r.q_m_end = 0;
  const int max_d = (int)(0.3 * (double)(q_len + t_len));  // DWmatch.c:96
  const int band_size = bw * 2;
  uint32_t longest_match = 0;
  bool start = false, matched = false;
  int best_m = -1, min_k = 0, max_k = 0, pbase = 0, x = 0, y = 0, d;
  int *Vp = sV[wid], *Vc = sV[wid] + PGB_MAXV;
  for (d = 0; d < max_d; d++) {
    if (max_k - min_k > band_size) break;  // DWmatch.c:120-122
    int idx = 0;
    for (int k = min_k; k <= max_k; k += 2, idx++) {
      if (d == 0) {
        x = 0;
      } else if (k == min_k) {  // DWmatch.c:125-130
        x = Vp[(k + 1 - pbase) >> 1];
      } else if (k == max_k) {
        x = Vp[(k - 1 - pbase) >> 1] + 1;
      } else {
        const int vm = Vp[(k - 1 - pbase) >> 1], vp = Vp[(k + 1 - pbase) >> 1];
        x = (vm < vp) ? vp : vm + 1;
      }
      y = x - k;
      const int x1 = x, y1 = y;
      for (;;) {  // snake, DWmatch.c:135-140, 1024 bases per round
        int rem = q_len - x;
        if (t_len - y < rem) rem = t_len - y;
        if (rem <= 0) break;
        int n_l = 32;
        if ((int)(32 * lane) < rem) {
          const uint32_t qb = qo + (uint32_t)x + 32 * lane, tb = to + (uint32_t)y + 32 * lane;
          const uint64_t *qp = qw + (qb >> 5), *tp = tw + (tb >> 5);
          const uint64_t df = window64(qp[0], qp[1], (qb & 31) * 2) ^ window64(tp[0], tp[1], (tb & 31) * 2);
          if (df) n_l = ctz64(df) >> 1;
        }
        const uint32_t bal = __ballot_sync(FULL, n_l < 32);
        int nn = 1024;
        if (bal) {
          const int L = __ffs((int)bal) - 1;
          nn = 32 * L + __shfl_sync(FULL, n_l, L);
        }
        if (nn > rem) nn = rem;
        x += nn;
        y += nn;
        if (nn < 1024) break;
      }
      if ((x - x1 > 16) && !start) { r.q_bgn = x1; r.t_bgn = y1; start = true; }                                        // DWmatch.c:142-146
      if ((uint32_t)(x - x1) > longest_match) { longest_match = (uint32_t)(x - x1); r.q_m_end = x; r.t_m_end = y; }  // :148-152
      if (lane == 0) Vc[idx] = x;
      if (x + y > best_m) best_m = x + y;
      if (x >= q_len || y >= t_len) {  // :161-164
        matched = true;
        break;
      }
    }
    if (matched) {  // :185-194
      r.q_end = x;
      r.t_end = y;
      r.dist = d;
      r.m_size = (r.q_end - r.q_bgn + r.t_end - r.t_bgn + 2 * d) / 2;
      break;
    }
    __syncwarp();
    int new_min_k = max_k, new_max_k = min_k;  // band trimming, :168-183
    idx = 0;
    for (int k2 = min_k; k2 <= max_k; k2 += 2, idx++) {
      if (2 * Vc[idx] - k2 >= best_m - bw) {
        if (k2 < new_min_k) new_min_k = k2;
        if (k2 > new_max_k) new_max_k = k2;
      }
    }
    pbase = min_k;
    max_k = new_max_k + 1;
    min_k = new_min_k - 1;
    int *tmp = Vp; Vp = Vc; Vc = tmp;
    __syncwarp();
  }
  if (!matched) { r.q_bgn = 0; r.t_bgn = 0; }  // :196-199
  if (lane == 0) {
    results[q.slot] = r;
    atomicAdd(bases_total, (unsigned long long)(r.q_end + r.t_end));
  }
}

}  // namespace pgb
#include "align_quad.cuh"
#include "align_coop.cuh"
namespace pgb {

// ---- pair-table maintenance between passes
__global__ void k_e_init(EEntry *E, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { E[i].key = PGB_EMPTY; E[i].v[0] = PGB_VNONE; E[i].v[1] = PGB_VNONE; }
}
__global__ void k_e_fill_new(EEntry *E, size_t n, uint32_t cur) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) E[i].v[cur] = PGB_VNONE;
}
// entries whose value changed between the two iterations: count, and list the read-id pairs (up to cap)
__global__ void k_e_diff_list(const EEntry *__restrict__ E, size_t n, unsigned long long *diffs, uint64_t *changed, uint32_t cap) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint4 e = *(const uint4 *)&E[i];
    if (e.z != e.w) {
      unsigned long long at = atomicAdd(diffs, 1ULL);
      if (at < cap) changed[at] = (uint64_t)e.x | ((uint64_t)e.y << 32);
    }
  }
}
// incremental pass: entries owned by clean buckets are carried over, entries owned by dirty buckets are rebuilt by their replay
__global__ void k_e_carry(EEntry *E, size_t n, uint32_t cur, const uint8_t *__restrict__ dirty) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint32_t v = E[i].v[cur ^ 1];
    E[i].v[cur] = (v != PGB_VNONE && !dirty[v >> 2]) ? v : PGB_VNONE;
  }
}

// ---- incremental replay (DESIGN.md "dirty buckets"): a bucket has to be replayed again only if it still had predicted
// alignments, or if the time-stamped pair table changed for a pair of reads that both occur in it.
__device__ __forceinline__ uint32_t bloom_bit(uint32_t rid) { return ht_mix(rid) & 255u; }
// per eligible bucket: (rid, rank) of every record + a 256-bit Bloom filter of its read ids
__global__ void k_bucket_reads(uint32_t n_ranks, const uint32_t *__restrict__ rank_off, const uint64_t *__restrict__ sy0, uint32_t *rec_rid,
                               uint32_t *rec_rank, uint64_t *bloom) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_ranks) return;
  uint64_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
  for (uint32_t t = rank_off[r]; t < rank_off[r + 1]; t++) {
    uint32_t rid = (uint32_t)(sy0[t] >> 32);
    rec_rid[t] = rid;
    rec_rank[t] = r;
    uint32_t bit = bloom_bit(rid);
    uint64_t m = 1ULL << (bit & 63);
    if ((bit >> 6) == 0) b0 |= m; else if ((bit >> 6) == 1) b1 |= m; else if ((bit >> 6) == 2) b2 |= m; else b3 |= m;
  }
  bloom[4 * (size_t)r] = b0; bloom[4 * (size_t)r + 1] = b1; bloom[4 * (size_t)r + 2] = b2; bloom[4 * (size_t)r + 3] = b3;
}
// thread per changed pair: every bucket holding read a that may also hold read b becomes dirty
__global__ void k_mark_dirty_pairs(const uint64_t *__restrict__ changed, uint32_t n_changed, const uint32_t *__restrict__ rid_sorted,
                                   const uint32_t *__restrict__ rank_sorted, uint32_t n_rec, const uint64_t *__restrict__ bloom, uint8_t *dirty) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_changed) return;
  const uint32_t a = (uint32_t)(changed[i] >> 32), b = (uint32_t)changed[i];
  uint32_t lo = 0, hi = n_rec;  // lower_bound(a)
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (rid_sorted[mid] < a) lo = mid + 1; else hi = mid;
  }
  const uint32_t bit = bloom_bit(b);
  for (uint32_t t = lo; t < n_rec && rid_sorted[t] == a; t++) {
    uint32_t r = rank_sorted[t];
    if (bloom[4 * (size_t)r + (bit >> 6)] >> (bit & 63) & 1ULL) dirty[r] = 1;
  }
}
__global__ void k_dirty_from_unknown(const uint8_t *__restrict__ unk_flag, uint32_t n_ranks, uint8_t *dirty) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_ranks) dirty[r] = unk_flag[r];
}
// Size classes of the buckets that run in a pass: flags[r] = runs && small, flags[n_ranks + r] = runs && big.  Small
// buckets are replayed one per thread (k_replay), big ones one per CTA (k_replay_block).  dirty == nullptr: all buckets run.
__global__ void k_class_flags(const uint32_t *__restrict__ rank_off, uint32_t n_ranks, const uint8_t *__restrict__ dirty, uint32_t big_n,
                              uint32_t *flags) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_ranks) return;
  const bool run = dirty ? dirty[r] != 0 : true;
  const bool big = rank_off[r + 1] - rank_off[r] >= big_n;
  flags[r] = run && !big;
  flags[n_ranks + r] = run && big;
}
// what the host needs to know after a pass, gathered into one struct (one device-to-host copy, one synchronisation per pass)
struct PassOut { unsigned long long unknown, diffs; uint32_t n_req; int err; uint32_t n_small, n_run; unsigned long long align_bases; };
__global__ void k_run_counts(const uint32_t *__restrict__ dpos, uint32_t n_ranks, uint32_t *cnt) {
  cnt[0] = dpos[n_ranks];       // small buckets of the run list
  cnt[1] = dpos[2 * n_ranks];   // all of them
}
__global__ void k_pass_out(const unsigned long long *__restrict__ ctr, const uint32_t *__restrict__ n_req, const int *__restrict__ err,
                           const uint32_t *__restrict__ cnt, const unsigned long long *__restrict__ align_bases, PassOut *out) {
  PassOut o;
  o.unknown = ctr[0]; o.diffs = ctr[1]; o.n_req = *n_req; o.err = *err; o.n_small = cnt[0]; o.n_run = cnt[1]; o.align_bases = *align_bases;
  *out = o;
}
__global__ void k_compact_classes(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos, uint32_t n_ranks, uint32_t *list) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * n_ranks) return;
  if (flags[i]) list[pos[i]] = i < n_ranks ? i : i - n_ranks;
}

// ------------------------------------------------------------------------------------------------ shimmer4py index handle
// per read id: index of its first minimizer in the concatenated list and how many it has (get_ridmm, src/shmr_utils.c:415-443)
__global__ void k_ridmm(const mm128 *__restrict__ mm, size_t n, uint32_t *first, uint32_t *count) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t rid = (uint32_t)(mm[i].y >> 32);
  atomicMin(&first[rid], (uint32_t)i);
  atomicAdd(&count[rid], 1u);
}
__global__ void k_slot_to_group(const uint32_t *__restrict__ gslot, uint32_t n, uint32_t *slot2j) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) slot2j[gslot[j]] = j;
}
__global__ void k_count_one(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t mask, uint64_t mer, uint32_t *out) {
  uint32_t s = ht_find(keys, mask, mer);
  *out = s == PGB_NOSLOT ? 0u : vals[s];
}
// outer key -> (group id or NOSLOT)
__global__ void k_outer_lookup(const uint64_t *__restrict__ xkeys, uint32_t xmask, const uint32_t *__restrict__ xfirst, const uint32_t *__restrict__ xid,
                               uint64_t key, uint32_t *out) {
  uint32_t s = ht_find(xkeys, xmask, key);
  *out = (s == PGB_NOSLOT || xfirst[s] == 0xFFFFFFFFu) ? PGB_NOSLOT : xid[s];
}

// ------------------------------------------------------------------------------------------------ shmr_aln (co-linear chaining)
// src/shmr_align.c:21-160, batched: pair p chains list0[off0[p], off0[p+1]) against list1[off1[p], off1[p+1]).
//  * the reference's MMIDX hash map (:40-57: hash -> ascending indices into list 0) is one radix sort of all list-0 elements by
//    (pair, low 40 hash bits) with the element index as the value (k_aln_keys0; the sort is stable, so equal keys stay in
//    ascending index order; elements of a pair stay inside the pair's own range [off0[p], off0[p+1]));
//  * k_aln_lookup<false/true>: one thread per list-1 element binary-searches its pair's range and counts / writes the list-0
//    indices whose FULL hash equals its own;
//  * k_aln_chain_warp: the greedy chaining is sequential in the hits (every hit extends the best existing chain or opens a new one,
//    :97-147), so one WARP walks a pair's hits in order and its lanes evaluate the existing chains in parallel (arg-min of
//    (diff, chain id) by shuffle = the reference's first chain with the smallest diff); it emits (chain id, idx0, idx1) per hit.
struct AlnHit { uint32_t chain, i0, i1; };
struct AlnChain { uint32_t last_i0, lp0; int32_t delta1; uint32_t n; };
__device__ __forceinline__ int64_t aln_abs_u32(uint32_t v) { int32_t x = (int32_t)v; return x < 0 ? -(int64_t)x : (int64_t)x; }  // abs((int)uint32)
#define PGB_ALN_HBITS 40
// pair id of every element of a concatenated list
__global__ void k_aln_pair_ids(const uint64_t *__restrict__ off, uint32_t n_pairs, uint64_t n, uint32_t *pid) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t lo = 0, hi = n_pairs;  // last p with off[p] <= i
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  pid[i] = lo;
}
__global__ void k_aln_keys0(const mm128 *__restrict__ a0, const uint32_t *__restrict__ pid0, uint64_t n0, uint64_t *keys, uint32_t *idx) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n0) return;
  keys[i] = ((uint64_t)pid0[i] << PGB_ALN_HBITS) | ((a0[i].x >> 8) & ((1ULL << PGB_ALN_HBITS) - 1));
  idx[i] = (uint32_t)i;
}
template <bool FILL>
__global__ void k_aln_lookup(const mm128 *__restrict__ a0, const mm128 *__restrict__ a1, const uint32_t *__restrict__ pid1, uint64_t n1,
                             const uint64_t *__restrict__ off0, const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sidx,
                             uint32_t *cnt, const uint64_t *__restrict__ moff, uint32_t *midx) {
  uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n1) return;
  const uint32_t p = pid1[s];
  const uint64_t h = a1[s].x >> 8, key = ((uint64_t)p << PGB_ALN_HBITS) | (h & ((1ULL << PGB_ALN_HBITS) - 1));
  uint64_t lo = off0[p], hi = off0[p + 1];
  const uint64_t end = hi, first = lo;
  while (lo < hi) {  // lower bound of key in the pair's range
    uint64_t mid = (lo + hi) >> 1;
    if (skeys[mid] < key) lo = mid + 1; else hi = mid;
  }
  uint32_t c = 0;
  uint64_t at = FILL ? moff[s] : 0;
  for (; lo < end && skeys[lo] == key; lo++) {
    const uint32_t i = sidx[lo];
    if ((a0[i].x >> 8) != h) continue;
    if (FILL) midx[at++] = (uint32_t)(i - first);  // index inside the pair's list 0
    c++;
  }
  if (!FILL) cnt[s] = c;
}
// per pair: upper bound of its hits = sum of the match counts that pass the max_repeat filter (:71-80)
__global__ void k_aln_hit_bound(const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ off1, uint32_t n_pairs, uint32_t max_repeat, uint32_t *ub) {
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  uint32_t t = 0;
  for (uint64_t s = off1[p] + lane; s < off1[p + 1]; s += 32) { const uint32_t c = cnt[s]; if (c && c <= max_repeat) t += c; }
  t = __reduce_add_sync(0xffffffffu, t);
  if (lane == 0) ub[p] = t;
}
__global__ void __launch_bounds__(128) k_aln_chain_warp(const mm128 *__restrict__ a0, const mm128 *__restrict__ a1, const uint64_t *__restrict__ off0,
                                                        const uint64_t *__restrict__ off1, uint32_t n_pairs, const uint32_t *__restrict__ cnt,
                                                        const uint64_t *__restrict__ moff, const uint32_t *__restrict__ midx, uint32_t direction,
                                                        uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat, const uint64_t *__restrict__ hoff,
                                                        AlnChain *chains, AlnHit *hits, uint32_t *n_hits_out, uint32_t *n_chains_out) {
  constexpr uint32_t FULL = 0xffffffffu;
  const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const mm128 *l0 = a0 + off0[p], *l1 = a1 + off1[p];
  const uint64_t s_base = off1[p];
  const uint32_t n1 = (uint32_t)(off1[p + 1] - s_base);
  AlnChain *ch = chains + hoff[p];
  AlnHit *out = hits + hoff[p];
  uint32_t n_hits = 0, n_chains = 0, small_aln_count = 0;
  for (uint32_t ss = 0; ss < n1; ss++) {
    uint32_t s = ss;
    if (direction == 1) {         // the reference indexes a[n - ss] here, i.e. one past the end for ss == 0 (UB, SURVEY A-7);
      if (ss == 0) continue;      // that element is skipped
      s = n1 - ss;
    }
    const uint32_t c = cnt[s_base + s];
    if (c == 0 || c > max_repeat) continue;            // :71-80
    const mm128 m1 = l1[s];
    const uint32_t pos1 = (uint32_t)((m1.y & 0xFFFFFFFFULL) >> 1);
    const uint64_t mo = moff[s_base + s];
    for (uint32_t q = 0; q < c; q++) {
      const uint32_t i = midx[mo + q];
      const mm128 m0 = l0[i];
      const uint32_t pos0 = (uint32_t)((m0.y & 0xFFFFFFFFULL) >> 1);
      if (direction == 0 && (m0.y & 1) != (m1.y & 1)) continue;   // :86-92
      if (direction == 1 && (m0.y & 1) == (m1.y & 1)) continue;
      const int64_t delta0 = direction == 1 ? aln_abs_u32(pos0 + pos1) : aln_abs_u32(pos0 - pos1);
      // :103-132 over the existing chains, 32 at a time: smallest diff below max_diff, the first chain on ties
      uint64_t bestk = ~0ULL;
      uint32_t small = 0;
      __syncwarp();  // (chain records written by lane 0 for the previous hit)
      for (uint32_t a = lane; a < ((n_chains + 31) & ~31u); a += 32) {
        uint64_t k = ~0ULL;
        bool is_small = false;
        if (a < n_chains) {
          const AlnChain cc = ch[a];
          is_small = cc.n < 3;
          if (i >= cc.last_i0) {
            const int64_t mm_dist = aln_abs_u32(pos0 - cc.lp0);
            if (mm_dist < (int64_t)max_dist) {
              int32_t dd = (int32_t)delta0 - cc.delta1;
              const uint32_t diff = (uint32_t)(dd < 0 ? -dd : dd);
              if (diff < max_diff) k = ((uint64_t)diff << 32) | a;
            }
          }
        }
        small += __popc(__ballot_sync(FULL, is_small));
        if (k < bestk) bestk = k;
      }
      small_aln_count = small;
      for (int o = 16; o; o >>= 1) { const uint64_t t = __shfl_xor_sync(FULL, bestk, o); if (t < bestk) bestk = t; }
      uint32_t best = (uint32_t)bestk;
      const bool fresh = bestk == ~0ULL;
      if (fresh) best = n_chains++;
      if (lane == 0) {
        AlnChain cc;
        cc.last_i0 = i;
        cc.lp0 = pos0;
        cc.delta1 = (int32_t)delta0;  // what a later hit computes from this chain's last pair (:118-121) is this hit's delta0
        cc.n = fresh ? 1u : ch[best].n + 1u;
        ch[best] = cc;
        AlnHit h;
        h.chain = best; h.i0 = i; h.i1 = s;
        out[n_hits] = h;
      }
      n_hits++;
    }
    if (small_aln_count > 4800) break;                 // MAX_SMALL_ALNS, :19,149
  }
  if (lane == 0) { n_hits_out[p] = n_hits; n_chains_out[p] = n_chains; }
}

}  // namespace pgb

#include "dedup.cuh"
#include "map.cuh"
