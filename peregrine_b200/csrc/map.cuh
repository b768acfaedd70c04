// map.cuh — shmr_map's process_map (src/shmr_map.c:48-166) on the GPU: contig ("ref") shimmers against the SHIMMER-pair
// index of the reads, no alignment.
//
// The reference walks the contig list once.  mmer0 starts at the first element whose x is an OUTER key of the pair index
// (:85-91); afterwards every element whose hash is in the reads' multiplicity table with lower <= count <= upper becomes
// mmer1 and then the next mmer0 (:93-100 and every branch below ends in `mmer0 = mmer1`).  A consecutive pair on the same
// contig, at least 100 bases apart, whose (x0, x1) bucket exists, prints one line per record of the bucket in INSERTION
// order (no qsort here, :128-151).  So the walk is a compaction (the kept elements) followed by independent work per
// adjacent kept pair:
//   k_map_ref_flags   1 thread = 1 contig shimmer: count filter; atomicMin of the first outer-key element
//   k_map_kept        kept flag = (i == first) || (i > first && count ok)
//   k_map_pair_count  1 thread = 1 adjacent kept pair: bucket lookup -> number of lines
//   k_map_emit        1 thread = 1 pair: one 9-field tuple per bucket record, in insertion order
//   k_map_len / k_map_write   "%u %u %u %u %u %u %d %u %u\n" per tuple (two passes: line lengths -> scan -> text)
#pragma once
#include "dedup.cuh"

namespace pgb {

struct map_hit { uint32_t ref_id, ref_bgn, ref_end, read_id, read_bgn, read_end, dir, mcount0, mcount1; };

PGB_HD int map_format(const map_hit &h, char *buf) {  // src/shmr_map.c:149-150
  int n = 0;
  n += fmt_u64(buf + n, h.ref_id); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.ref_bgn); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.ref_end); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.read_id); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.read_bgn); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.read_end); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.dir); buf[n++] = ' ';  // uint8_t printed with %d
  n += fmt_u64(buf + n, h.mcount0); buf[n++] = ' ';
  n += fmt_u64(buf + n, h.mcount1);
  buf[n++] = '\n';
  return n;
}

#if defined(__CUDACC__)
// cnt[i] = multiplicity of the element's hash in the READS' table (0xFFFFFFFF = absent, :96-97)
__global__ void k_map_ref_flags(const mm128 *__restrict__ ref, size_t n, const uint64_t *__restrict__ mckeys, const uint32_t *__restrict__ mcvals,
                                uint32_t mcmask, const uint64_t *__restrict__ xkeys, uint32_t xmask, const uint32_t *__restrict__ xfirst,
                                uint32_t *cnt, unsigned long long *first_outer) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const mm128 m = ref[i];
  const uint32_t s = ht_find(mckeys, mcmask, m.x >> 8);
  cnt[i] = s == PGB_NOSLOT ? 0xFFFFFFFFu : mcvals[s];
  const uint32_t xs = ht_find(xkeys, xmask, m.x);
  if (xs != PGB_NOSLOT && xfirst[xs] != 0xFFFFFFFFu && (unsigned long long)i < *(volatile unsigned long long *)first_outer)
    atomicMin(first_outer, (unsigned long long)i);  // kh_get(MMER0, mmer0.x) hits, :88-89
}
__global__ void k_map_kept(const uint32_t *__restrict__ cnt, size_t n, uint32_t lower, uint32_t upper, const unsigned long long *first_outer,
                           uint32_t *flags) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long s = *first_outer;
  const uint32_t c = cnt[i];
  flags[i] = (i == s) || (i > s && c != 0xFFFFFFFFu && !(c < lower || c > upper));
}
// bucket of the adjacent kept pair t, or PGB_NOSLOT (:102-123)
__device__ __forceinline__ uint32_t map_pair_bucket(const mm128 &m0, const mm128 &m1, const uint64_t *xkeys, uint32_t xmask, const uint64_t *bkeys,
                                                    uint32_t bmask) {
  if ((m0.y >> 32) != (m1.y >> 32)) return PGB_NOSLOT;
  const uint32_t s0 = ht_find(xkeys, xmask, m0.x);
  if (s0 == PGB_NOSLOT) return PGB_NOSLOT;
  const uint32_t s1 = ht_find(xkeys, xmask, m1.x);
  if (s1 == PGB_NOSLOT) return PGB_NOSLOT;
  const uint32_t b = ht_find(bkeys, bmask, ((uint64_t)s0 << 32) | s1);
  if (b == PGB_NOSLOT) return PGB_NOSLOT;
  if (!pair_far_enough(m0.y, m1.y)) return PGB_NOSLOT;
  return b;
}
__global__ void k_map_pair_count(const mm128 *__restrict__ ref, const uint32_t *__restrict__ kept, uint32_t n_kept, const uint64_t *__restrict__ xkeys,
                                 uint32_t xmask, const uint64_t *__restrict__ bkeys, uint32_t bmask, const uint32_t *__restrict__ bcount,
                                 uint32_t *n_hits, uint32_t *pair_bucket) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_kept) return;
  uint32_t b = PGB_NOSLOT;
  if (t + 1 < n_kept) b = map_pair_bucket(ref[kept[t]], ref[kept[t + 1]], xkeys, xmask, bkeys, bmask);
  pair_bucket[t] = b;
  n_hits[t] = b == PGB_NOSLOT ? 0u : bcount[b];
}
// by_bucket: record indices stably sorted by bucket slot (insertion order inside a bucket); bstart[b] = first position of slot b
__global__ void k_map_emit(const mm128 *__restrict__ ref, const uint32_t *__restrict__ kept, uint32_t n_kept, const uint32_t *__restrict__ cnt,
                           const uint32_t *__restrict__ pair_bucket, const uint32_t *__restrict__ n_hits, const uint64_t *__restrict__ hit_off,
                           const uint32_t *__restrict__ bstart, const uint32_t *__restrict__ by_bucket, PairSoA r, map_hit *out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_kept || n_hits[t] == 0) return;
  const uint32_t i0 = kept[t], i1 = kept[t + 1];
  const mm128 m0 = ref[i0], m1 = ref[i1];
  map_hit h;
  h.ref_id = (uint32_t)(m0.y >> 32);
  h.ref_bgn = (uint32_t)((m0.y & 0xFFFFFFFFULL) >> 1);
  h.ref_end = (uint32_t)((m1.y & 0xFFFFFFFFULL) >> 1);
  h.mcount0 = cnt[i0];
  h.mcount1 = cnt[i1];
  const uint32_t b0 = bstart[pair_bucket[t]], nh = n_hits[t];
  map_hit *o = out + hit_off[t];
  for (uint32_t j = 0; j < nh; j++) {
    const uint32_t rec = by_bucket[b0 + j];
    const uint64_t y0 = r.y0[rec], y1 = r.y1[rec];
    h.read_id = (uint32_t)(y0 >> 32);
    h.read_bgn = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1);
    h.read_end = (uint32_t)((y1 & 0xFFFFFFFFULL) >> 1);
    h.dir = r.dir[rec];
    o[j] = h;
  }
}
__global__ void k_map_len(const map_hit *__restrict__ hits, size_t n, uint32_t *len) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  char buf[112];
  len[i] = (uint32_t)map_format(hits[i], buf);
}
__global__ void k_map_write(const map_hit *__restrict__ hits, size_t n, const uint64_t *__restrict__ off, char *text) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  char buf[112];
  const int l = map_format(hits[i], buf);
  char *o = text + off[i];
  for (int j = 0; j < l; j++) o[j] = buf[j];
}
__global__ void k_iota_u32(uint32_t *p, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
#endif

}  // namespace pgb
