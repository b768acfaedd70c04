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

namespace pgb {

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
                             uint64_t *__restrict__ w, uint32_t *__restrict__ nm, uint32_t *__restrict__ hasn_by_rid) {
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
  const uint8_t *s = raw + row_raw_off[row] + p0;
  uint32_t cnt = (len - p0) < 32 ? (uint32_t)(len - p0) : 32u;
  uint64_t bits = 0;
  uint32_t nmask = 0;
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

// ------------------------------------------------------------------------------------------------ sketch (exact automaton)
// Thread per read.  WRITE=false: count only.  WRITE=true: write at out + out_off[sel].
template <bool WRITE>
__global__ void k_sketch_exact(const uint64_t *__restrict__ w, const uint32_t *__restrict__ nm,
                               const uint32_t *__restrict__ sel_rows, uint32_t n_sel, const uint32_t *__restrict__ row_rid,
                               const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                               const uint32_t *__restrict__ hasn_by_rid, int wsz, int k, uint32_t *__restrict__ counts,
                               const uint64_t *__restrict__ out_off, mm128 *__restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sel) return;
  uint32_t row = sel_rows[t];
  uint32_t rid = row_rid[row];
  int len = (int)row_len[row];
  uint64_t ring_x[256];
  uint32_t ring_p[256];
  uint32_t n = 0;
  mm128 *dst = WRITE ? out + out_off[t] : nullptr;
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
  if (!WRITE) counts[t] = n;
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
  if (c >= lower && c < upper) atomicMin(first_strict, (unsigned long long)i);  // src/shmr_utils.c:318
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

// ------------------------------------------------------------------------------------------------ bucket tables
// X table: distinct full x values -> slot (a dense id).  B table: (slot(x0)<<32 | slot(x1)) -> bucket.
__global__ void k_bucket_insert(PairSoA r, uint32_t n_rec, uint64_t *xkeys, uint32_t xmask, uint64_t *bkeys, uint32_t bmask,
                                uint32_t *bcount, uint32_t *bfirst, uint32_t *blast, uint32_t *rec_bucket, int *err) {
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
  rec_bucket[i] = b;
}
struct BucketInfo { uint64_t k0, k1; uint32_t first_seq, count, slot, last_seq; };
__global__ void k_bucket_dump(const uint64_t *__restrict__ xkeys, const uint64_t *__restrict__ bkeys,
                              const uint32_t *__restrict__ bcount, const uint32_t *__restrict__ bfirst,
                              const uint32_t *__restrict__ blast, const uint32_t *__restrict__ pos, size_t cap, BucketInfo *out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap || bkeys[i] == PGB_EMPTY) return;
  BucketInfo b;
  uint64_t key = bkeys[i];
  b.k0 = xkeys[(uint32_t)(key >> 32)];
  b.k1 = xkeys[(uint32_t)key];
  b.first_seq = bfirst[i];
  b.count = bcount[i];
  b.slot = (uint32_t)i;
  b.last_seq = blast[i];
  out[pos[i]] = b;
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
struct AlnReq { uint32_t rid0, start0, rid1, strands; };  // strands: bit0 = strand0, bit1 = strand1
struct ReplayState {
  // time-stamped rid_pairs
  uint64_t *ekeys; uint64_t *eold; uint64_t *enew; uint32_t emask;
  // alignment cache
  uint64_t *akeys; uint32_t *aidx; uint32_t amask;
  AlnReq *reqs; match_t *results; uint32_t req_cap;
  uint32_t *n_req;      // device counter
  uint32_t n_done;      // requests [0, n_done) have results
  const uint32_t *rlen_by_rid;
  int *err;
};
struct DevReplayCtx {
  ReplayState s;
  uint32_t rank;
  bool request_enabled;
  ovlp_rec *out;
  __device__ uint32_t rlen(uint32_t rid) const { return s.rlen_by_rid[rid]; }
  __device__ uint64_t pair_old(uint64_t p) const {
    uint32_t sl = ht_find(s.ekeys, s.emask, p);
    return sl == PGB_NOSLOT ? ~0ULL : s.eold[sl];
  }
  __device__ uint64_t pair_new(uint64_t p) const {
    uint32_t sl = ht_find(s.ekeys, s.emask, p);
    return sl == PGB_NOSLOT ? ~0ULL : *(volatile uint64_t *)&s.enew[sl];
  }
  __device__ void pair_set(uint64_t p, uint64_t v) {
    uint32_t sl = ht_insert(s.ekeys, s.emask, p);
    if (sl == PGB_NOSLOT) { atomicOr(s.err, 32); return; }
    atomicMin((unsigned long long *)&s.enew[sl], (unsigned long long)v);
  }
  __device__ bool aln_get(uint32_t i, uint32_t j, match_t *m) const {
    uint32_t sl = ht_find(s.akeys, s.amask, ((uint64_t)rank << 32) | ((uint64_t)i << 16) | j);
    if (sl == PGB_NOSLOT) return false;
    uint32_t idx = s.aidx[sl];
    if (idx >= s.n_done) return false;
    *m = s.results[idx];
    return true;
  }
  __device__ void aln_request(uint32_t i, uint32_t j, uint32_t rid0, uint32_t start0, uint32_t s0, uint32_t rid1, uint32_t s1) {
    if (!request_enabled) return;
    uint64_t key = ((uint64_t)rank << 32) | ((uint64_t)i << 16) | j;
    uint32_t sl = ht_find(s.akeys, s.amask, key);
    if (sl != PGB_NOSLOT) return;  // already requested in this pass (cannot happen: each (i,j) is visited once)
    uint32_t idx = atomicAdd(s.n_req, 1u);
    if (idx >= s.req_cap) { atomicOr(s.err, 64); return; }
    sl = ht_insert(s.akeys, s.amask, key);
    if (sl == PGB_NOSLOT) { atomicOr(s.err, 64); return; }
    s.aidx[sl] = idx;
    AlnReq q;
    q.rid0 = rid0; q.start0 = start0; q.rid1 = rid1; q.strands = s0 | (s1 << 1);
    s.reqs[idx] = q;
  }
  __device__ void emit(uint32_t n, const ovlp_rec &o) { out[n] = o; }
};

__global__ void k_replay(ReplayState st, uint32_t n_ranks, const uint32_t *__restrict__ rank_off, const uint64_t *__restrict__ sy0,
                         const uint8_t *__restrict__ sdir, uint8_t *contained, uint32_t bestn, int request_enabled, int do_emit,
                         uint32_t *acc_count, const uint32_t *__restrict__ out_off, ovlp_rec *out,
                         unsigned long long *n_unknown_total) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_ranks) return;
  uint32_t b = rank_off[r], n = rank_off[r + 1] - b;
  DevReplayCtx c;
  c.s = st;
  c.rank = r;
  c.request_enabled = request_enabled != 0;
  c.out = do_emit ? out + out_off[r] : nullptr;
  uint32_t unk = 0;
  uint32_t acc = replay_bucket(c, r, sy0 + b, sdir + b, n, contained + b, bestn, do_emit != 0, &unk);
  acc_count[r] = acc;
  if (unk) atomicAdd(n_unknown_total, (unsigned long long)unk);
}

#define PGB_MAXV 264  // supports band_tolerance <= 256
__global__ void k_align(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n, const uint64_t *__restrict__ w,
                        const uint32_t *__restrict__ nm, const uint64_t *__restrict__ woff_by_rid,
                        const uint32_t *__restrict__ rlen_by_rid, const uint32_t *__restrict__ hasn_by_rid, int bw, match_t *results,
                        int *err, unsigned long long *bases_total) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AlnReq q = reqs[first + i];
  uint32_t rl0 = rlen_by_rid[q.rid0], rl1 = rlen_by_rid[q.rid1];
  SeqView qv = make_view(w, nm, woff_by_rid[q.rid0], rl0, q.start0, q.strands & 1, (int)hasn_by_rid[q.rid0]);
  SeqView tv = make_view(w, nm, woff_by_rid[q.rid1], rl1, 0, (q.strands >> 1) & 1, (int)hasn_by_rid[q.rid1]);
  int Va[PGB_MAXV], Vb[PGB_MAXV];
  match_t m;
  int e = 0;
  ovlp_match_core(qv, (int)(rl0 - q.start0), tv, (int)rl1, bw, Va, Vb, PGB_MAXV, &m, &e);
  if (e) atomicOr(err, 128 | (e << 8));
  results[first + i] = m;
  // bases actually compared along the final path (algorithmic-bytes accounting, SURVEY 8d: (q_end + t_end)/4 per alignment)
  atomicAdd(bases_total, (unsigned long long)(m.q_end + m.t_end));
}

__global__ void k_table_diff(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, size_t n, unsigned long long *diffs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned int local = 0;
  for (; i < n; i += stride) local += a[i] != b[i];
  if (local) atomicAdd(diffs, (unsigned long long)local);
}

}  // namespace pgb
