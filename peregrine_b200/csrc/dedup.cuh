// dedup.cuh — shmr_dedup (src/shmr_dedup.c:19-101) on the GPU: raw ovlp_t stream -> preads.ovl text.
//
// The reference reads the concatenated per-chunk streams record by record, keeps the FIRST record of every unordered read
// pair (khash RPAIR, :41-44,97) and prints one text line per kept record (:46-96).  Here:
//   k_dedup_insert   1 thread = 1 record: pair key -> open-addressing table, atomicMin of the record index per pair
//   k_dedup_len      1 thread = 1 record: kept iff it holds its pair's minimum index; length of its text line
//   (exclusive scan of the lengths -> byte offset of every line; kept lines stay in stream order)
//   k_dedup_write    1 thread = 1 kept record: formats the line (printf-exact, see fmt_* below) at its offset
// All integer conversions follow the C of the reference literally (uint32_t a_bgn/a_end/b_bgn/b_end printed with %d, the
// `< 0` clamps that can never fire on unsigned values, the unsigned `>= rlen` clamps); "%0.1f" is reproduced exactly by
// integer arithmetic on the IEEE-754 bits (round-half-even on the exact binary value, as glibc does).
// The functions are PGB_HD so that tests/hostsim can run them on the CPU against the reference binary.
#pragma once
#include <string.h>
#include "shimmer_core.cuh"

namespace pgb {

// ---- printf-exact number formatting into a char buffer; every function returns the number of characters written
PGB_HD int fmt_u64(char *o, uint64_t v) {
  char t[20];
  int n = 0;
  do { t[n++] = (char)('0' + (int)(v % 10)); v /= 10; } while (v);
  for (int i = 0; i < n; i++) o[i] = t[n - 1 - i];
  return n;
}
// "%d" (zero_width = 0) or "%0Wd": the sign counts towards the width
PGB_HD int fmt_int(char *o, int32_t v, int zero_width) {
  int n = 0;
  uint64_t mag = v < 0 ? (uint64_t)(-(int64_t)v) : (uint64_t)v;
  if (v < 0) o[n++] = '-';
  char t[20];
  int d = fmt_u64(t, mag);
  for (int i = n + d; i < zero_width; i++) o[n++] = '0';
  for (int i = 0; i < d; i++) o[n++] = t[i];
  return n;
}
PGB_HD uint64_t dbl_bits(double v) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(v);
#else
  uint64_t b;
  memcpy(&b, &v, 8);
  return b;
#endif
}
// "%0.1f" of a finite double, |v| < 2^52: exact decimal rounding (nearest, ties to even) of the binary value
PGB_HD int fmt_f1(char *o, double v) {
  const uint64_t b = dbl_bits(v);
  int n = 0;
  if (b >> 63) o[n++] = '-';
  const int ex = (int)((b >> 52) & 0x7FF);
  uint64_t q = 0;  // round(|v| * 10)
  if (ex != 0) {   // zero / subnormal print as 0.0
    const uint64_t m = (b & ((1ULL << 52) - 1)) | (1ULL << 52);
    const int s = 1075 - ex;  // |v| = m * 2^-s
    const uint64_t N = m * 10;  // < 2^57
    if (s <= 0) q = N << (-s);  // not reached for the values shmr_dedup produces (|v| < 2^38)
    else if (s <= 63) {
      q = N >> s;
      const uint64_t rem = N & ((1ULL << s) - 1), half = 1ULL << (s - 1);
      if (rem > half || (rem == half && (q & 1))) q++;
    }
  }
  n += fmt_u64(o + n, q / 10);
  o[n++] = '.';
  o[n++] = (char)('0' + (int)(q % 10));
  return n;
}
PGB_HD int fmt_str(char *o, const char *s) {
  int n = 0;
  while (s[n]) { o[n] = s[n]; n++; }
  return n;
}

// One line of preads.ovl for a kept record (src/shmr_dedup.c:46-96); returns its length incl. the newline.  buf >= 192 bytes.
PGB_HD int dedup_format(const ovlp_rec &ov, char *buf) {
  const uint32_t rid0 = (uint32_t)(ov.y0 >> 32), rid1 = (uint32_t)(ov.y1 >> 32);
  const uint32_t pos0 = (uint32_t)((ov.y0 & 0xFFFFFFFFULL) >> 1) + 1, pos1 = (uint32_t)((ov.y1 & 0xFFFFFFFFULL) >> 1) + 1;
  const uint32_t rlen0 = ov.rl0, rlen1 = ov.rl1;
  const uint8_t strand0 = ov.strand0, strand1 = ov.strand1;
  int32_t q_bgn = ov.match.q_bgn, q_end = ov.match.q_end, t_bgn = ov.match.t_bgn, t_end = ov.match.t_end;
  uint32_t a_bgn, a_end, b_bgn, b_end;
  // signed arithmetic of the reference done modulo 2^32 (two's complement wrap, what gcc -O3 emits for it)
  q_bgn = (int32_t)((uint32_t)q_bgn - (uint32_t)t_bgn);  // :62
  t_bgn = 0;
  const uint32_t dp = pos0 - pos1;  // (seq_coor_t)(pos0 - pos1)
  if (strand0 == 0) {
    a_bgn = dp + (uint32_t)q_bgn;  // :65-69 (a_bgn < 0 is never true for a uint32_t)
    a_end = dp + (uint32_t)q_end;
    a_end = a_end >= rlen0 ? rlen0 : a_end;
  } else {
    a_bgn = rlen0 - dp - (uint32_t)q_end;  // :71-76
    a_end = rlen0 - dp - (uint32_t)q_bgn;
    a_end = a_end >= rlen0 ? rlen0 : a_end;
  }
  if (strand1 == 0) {
    b_bgn = (uint32_t)t_bgn;  // :78-82
    b_end = (uint32_t)t_end;
    b_end = b_end >= rlen1 ? rlen1 : b_end;
  } else {
    b_bgn = rlen1 - (uint32_t)t_end;  // :83-88
    b_end = rlen1 - (uint32_t)t_bgn;
    b_end = b_end >= rlen1 ? rlen1 : b_end;
  }
  int n = 0;
  n += fmt_int(buf + n, (int32_t)rid0, 9); buf[n++] = ' ';
  n += fmt_int(buf + n, (int32_t)rid1, 9); buf[n++] = ' ';
  n += fmt_int(buf + n, (int32_t)(0u - (uint32_t)ov.match.m_size), 0); buf[n++] = ' ';
  // err_est = 100.0 - 100.0 * (double)dist / (double)m_size  (:90-91), each operation rounded once (no contraction)
  const int32_t dist = ov.match.dist, msz = ov.match.m_size;
  if (msz == 0) {  // x / 0.0: glibc prints the x86 default NaN of 0.0/0.0 as "-nan"; 100 - (+-inf) = -+inf
    n += fmt_str(buf + n, dist == 0 ? "-nan" : (dist > 0 ? "-inf" : "inf"));
  } else {
#if defined(__CUDA_ARCH__)
    const double e = __dsub_rn(100.0, __ddiv_rn(__dmul_rn(100.0, (double)dist), (double)msz));
#else
    volatile double t1 = 100.0 * (double)dist;
    volatile double t2 = t1 / (double)msz;
    const double e = 100.0 - t2;
#endif
    n += fmt_f1(buf + n, e);
  }
  buf[n++] = ' ';
  buf[n++] = '0'; buf[n++] = ' ';  // ORIGINAL
  n += fmt_int(buf + n, (int32_t)a_bgn, 0); buf[n++] = ' ';
  n += fmt_int(buf + n, (int32_t)a_end, 0); buf[n++] = ' ';
  n += fmt_u64(buf + n, rlen0); buf[n++] = ' ';
  n += fmt_u64(buf + n, (uint32_t)(strand0 == 0 ? (int)strand1 : 1 - (int)strand1)); buf[n++] = ' ';  // int printed with %u
  n += fmt_int(buf + n, (int32_t)b_bgn, 0); buf[n++] = ' ';
  n += fmt_int(buf + n, (int32_t)b_end, 0); buf[n++] = ' ';
  n += fmt_u64(buf + n, rlen1); buf[n++] = ' ';
  n += fmt_str(buf + n, ov.ovlp_type == OVL_OVERLAP ? "overlap" : (ov.ovlp_type == OVL_CONTAINS ? "contains" : "contained"));
  buf[n++] = '\n';
  return n;
}
PGB_HD uint64_t dedup_pair_key(const ovlp_rec &ov) {  // :37-40
  const uint32_t rid0 = (uint32_t)(ov.y0 >> 32), rid1 = (uint32_t)(ov.y1 >> 32);
  return rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
}

// ---------------------------------------------------------------------------------------------- shmr_mkseqdb: encode_biseq
// .seqdb byte p of a read = fourbit_map_f[seq[p]] | fourbit_map_r[seq[len-1-p]] << 4 (src/shmr_utils.c:18-51): A/a 1, C/c 2,
// G/g 4, T/t 8 in the low nibble, the complement of the mirrored base in the high nibble, every other byte 0.
PGB_HD uint32_t biseq_f(uint32_t ch) {
  const uint32_t u = ch & 0xDFu;  // 'a'..'t' -> 'A'..'T'; no other byte folds onto A, C, G or T
  return u == 'A' ? 1u : u == 'C' ? 2u : u == 'G' ? 4u : u == 'T' ? 8u : 0u;
}
PGB_HD uint32_t biseq_r(uint32_t ch) {
  const uint32_t u = ch & 0xDFu;
  return u == 'A' ? 8u : u == 'C' ? 4u : u == 'G' ? 2u : u == 'T' ? 1u : 0u;
}
struct EncTile { uint64_t off; uint32_t len, p0; };  // bytes [p0, p0 + ENC_TILE) of the read at ascii[off .. off + len)
enum { ENC_TILE = 8192 };

#if defined(__CUDACC__)
// 1 CTA = one tile of one read; forward bytes are read ascending, mirrored bytes descending, both coalesced
__global__ void k_encode_biseq(const uint8_t *__restrict__ ascii, const EncTile *__restrict__ tiles, uint8_t *__restrict__ out) {
  const EncTile t = tiles[blockIdx.x];
  const uint8_t *s = ascii + t.off;
  uint8_t *o = out + t.off;
  const uint32_t p1 = t.p0 + ENC_TILE < t.len ? t.p0 + ENC_TILE : t.len;
  for (uint32_t p = t.p0 + threadIdx.x; p < p1; p += blockDim.x) o[p] = (uint8_t)(biseq_f(s[p]) | (biseq_r(s[t.len - 1 - p]) << 4));
}

// The pair table (keys / first) persists over the batches of one record stream: first[slot] = stream index of the first record of the
// pair, `base` = stream index of this batch's record 0.  A record is kept iff it holds its pair's minimum, i.e. no earlier batch
// and no earlier record of this batch had the pair (src/shmr_dedup.c:36-47 keeps the first record of every unordered pair).
__global__ void k_dedup_insert(const ovlp_rec *__restrict__ recs, size_t n, uint64_t *keys, uint32_t mask, unsigned long long *first,
                               uint32_t *slot_of, int *err, unsigned long long base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t key = dedup_pair_key(recs[i]);
  if (key == PGB_EMPTY) { atomicOr(err, 256); slot_of[i] = PGB_NOSLOT; return; }
  const uint32_t s = ht_insert(keys, mask, key);
  slot_of[i] = s;
  if (s == PGB_NOSLOT) { atomicOr(err, 256); return; }
  atomicMin(&first[s], base + (unsigned long long)i);
}
// the table grew: move every (pair, first) entry into the new one
__global__ void k_dedup_rehash(const uint64_t *__restrict__ old_keys, const unsigned long long *__restrict__ old_first, size_t old_cap, uint64_t *keys,
                               uint32_t mask, unsigned long long *first, int *err) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_cap) return;
  const uint64_t key = old_keys[i];
  if (key == PGB_EMPTY) return;
  const uint32_t s = ht_insert(keys, mask, key);
  if (s == PGB_NOSLOT) { atomicOr(err, 256); return; }
  first[s] = old_first[i];
}
__global__ void k_dedup_len(const ovlp_rec *__restrict__ recs, size_t n, const unsigned long long *__restrict__ first,
                            const uint32_t *__restrict__ slot_of, uint32_t *len, unsigned long long *n_kept, unsigned long long base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slot_of[i];
  uint32_t l = 0;
  if (s != PGB_NOSLOT && first[s] == base + (unsigned long long)i) {
    char buf[192];
    l = (uint32_t)dedup_format(recs[i], buf);
    const unsigned m = __activemask();  // one atomic per converged group of kept records
    if ((int)(threadIdx.x & 31) == __ffs((int)m) - 1) atomicAdd(n_kept, (unsigned long long)__popc(m));
  }
  len[i] = l;
}
__global__ void k_dedup_write(const ovlp_rec *__restrict__ recs, size_t n, const uint32_t *__restrict__ len, const uint64_t *__restrict__ off,
                              char *text) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || len[i] == 0) return;
  char buf[192];
  const int l = dedup_format(recs[i], buf);
  char *o = text + off[i];
  for (int j = 0; j < l; j++) o[j] = buf[j];
}
#endif

}  // namespace pgb
