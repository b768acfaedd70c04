// shimmer_core.cuh — exact per-item algorithms of the SHIMMER index / overlap path, written once as
// __host__ __device__ inline functions.  The CUDA kernels in this directory call them on the device; the
// host build of the same functions exists ONLY for tests/hostsim (a CPU simulator of kernel logic used while
// developing without a GPU) — libpgb200.so never executes them on the host for product results.
//
// Reference semantics followed (file:line relative to the reference tree):
//   hash64                    src/mm_sketch.c:23-32
//   (w,k)-minimizer automaton src/mm_sketch.c:70-151
//   hierarchical reduction    src/shmr_reduce.c:33-90
//   banded O(ND) overlap      src/DWmatch.c:66-204
//   greedy bucket scan        src/shmr_overlap.c:52-180
// Data layout differs from the reference on purpose (B200-first): reads are 2-bit packed (32 bases per
// 64-bit word, base p at bits 2*(p&31), A=0 C=1 G=2 T=3) with a separate 1-bit/base N mask, instead of the
// reference's 1-byte/base two-nibble .seqdb image.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PGB_HD __host__ __device__ __forceinline__
#else
#define PGB_HD inline
#endif

namespace pgb {

struct mm128 {
  uint64_t x, y;
};  // same layout as mm128_t, src/shimmer.h:24-26

struct match_t {  // = ovlp_match_t, src/shimmer.h:97-102
  int32_t m_size, dist, q_bgn, q_end, t_bgn, t_end, t_m_end, q_m_end;
};

struct ovlp_rec {  // = ovlp_t, src/shimmer.h:104-110 (64 bytes; pad bytes written as zero)
  uint64_t y0, y1;
  uint32_t rl0, rl1;
  uint8_t strand0, strand1, ovlp_type, pad0;
  match_t match;
  uint32_t pad1;
};

enum { OVL_OVERLAP = 0, OVL_CONTAINS = 1, OVL_CONTAINED = 2 };  // src/shmr_overlap.c:37-39
enum { READ_END_FUZZINESS = 48 };                               // src/shmr_overlap.c:34

// ---------------------------------------------------------------------------------------------- bit helpers
PGB_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)v) - 1;
#else
  return __builtin_ctzll(v);
#endif
}
PGB_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}
// reverse the order of the 32 two-bit groups of a word
PGB_HD uint64_t rev2(uint64_t v) {
#if defined(__CUDA_ARCH__)
  v = __brevll(v);
  return ((v >> 1) & 0x5555555555555555ULL) | ((v & 0x5555555555555555ULL) << 1);
#else
  v = ((v >> 2) & 0x3333333333333333ULL) | ((v & 0x3333333333333333ULL) << 2);
  v = ((v >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((v & 0x0F0F0F0F0F0F0F0FULL) << 4);
  return __builtin_bswap64(v);
#endif
}
PGB_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
  return __builtin_bswap32(v);
#endif
}

// ---------------------------------------------------------------------------------------------- hash64
// src/mm_sketch.c:23-32 (invertible integer mix, masked to 2k bits after every add step)
PGB_HD uint64_t hash64(uint64_t key, uint64_t mask) {
  key = (~key + (key << 21)) & mask;
  key = key ^ key >> 24;
  key = ((key + (key << 3)) + (key << 8)) & mask;
  key = key ^ key >> 14;
  key = ((key + (key << 2)) + (key << 4)) & mask;
  key = key ^ key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

// ---------------------------------------------------------------------------------------------- packed sequence view
// Logical base i of a (read, strand, start) view:
//   strand 0: forward base  (start + i)
//   strand 1: reverse-complement of the read, i.e. 3 - forward base (rlen-1-(start+i))   (the .seqdb high nibble,
//             src/shmr_utils.c:44-51)
// `a0` is the absolute base index (in the global packed array) of logical base 0; `dirn` is +1 / -1.
struct SeqView {
  const uint64_t *w;   // packed 2-bit words (global array, guard words at both ends)
  const uint32_t *nm;  // N mask, 1 bit per base, parallel to w (32 bases per u32); may be null when !has_n
  int64_t a0;
  int rev;    // 0 forward, 1 reverse-complement
  int has_n;  // consult nm
};

PGB_HD SeqView make_view(const uint64_t *w, const uint32_t *nm, uint64_t word_off, uint32_t rlen, uint32_t start,
                         int strand, int has_n) {
  SeqView v;
  v.w = w;
  v.nm = nm;
  v.rev = strand ? 1 : 0;
  v.has_n = has_n;
  int64_t base0 = (int64_t)word_off * 32;
  v.a0 = strand ? base0 + (int64_t)rlen - 1 - (int64_t)start : base0 + (int64_t)start;
  return v;
}

PGB_HD uint64_t fetch_fwd64(const uint64_t *w, int64_t a) {  // 32 bases starting at absolute base a
  int64_t i = a >> 5;
  int s = (int)(a & 31) * 2;
  uint64_t lo = w[i] >> s;
  uint64_t hi = w[i + 1];
  return s ? (lo | (hi << (64 - s))) : lo;
}
PGB_HD uint32_t fetch_fwd_n32(const uint32_t *nm, int64_t a) {  // 32 N-bits starting at absolute base a
  int64_t i = a >> 5;
  int s = (int)(a & 31);
  uint32_t lo = nm[i] >> s;
  uint32_t hi = nm[i + 1];
  return s ? (lo | (hi << (32 - s))) : lo;
}
// 32 logical bases starting at logical index x; base x+j at bits 2j
PGB_HD uint64_t fetch32(const SeqView &v, int x) {
  if (!v.rev) return fetch_fwd64(v.w, v.a0 + x);
  return ~rev2(fetch_fwd64(v.w, v.a0 - x - 31));
}
PGB_HD uint32_t fetchn32(const SeqView &v, int x) {
  if (!v.rev) return fetch_fwd_n32(v.nm, v.a0 + x);
  return brev32(fetch_fwd_n32(v.nm, v.a0 - x - 31));
}

// ---------------------------------------------------------------------------------------------- ovlp_match
// Banded greedy O(ND) forward extension from (0,0), src/DWmatch.c:66-204, restated with O(band) state:
// only the previous d's furthest-reaching x per diagonal is kept (the reference's V[2*max_d+1] is only ever
// read on diagonals written at d-1; U[k] = x+y = 2*V[k]-k is recomputed instead of stored).  Snakes compare
// 32 bases per step on the packed words (XOR + count-trailing-zeros) instead of byte-wise nibbles.
// Vp / Vc: caller-provided scratch of `cap` ints each; needs cap >= band_tolerance + 3.
// *err is OR-ed with 1 if the (never observed) empty-band state is reached, 2 if cap is too small.
PGB_HD void ovlp_match_core(const SeqView &q, int q_len, const SeqView &t, int t_len, int band_tolerance, int *Vp,
                            int *Vc, int cap, match_t *out, int *err) {
  match_t r;
  r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = r.t_m_end = r.q_m_end = 0;
  const int max_d = (int)(0.3 * (double)(q_len + t_len));  // DWmatch.c:96 (double multiply, truncation)
  const int band_size = band_tolerance * 2;
  uint32_t longest_match = 0;
  bool start = false, matched = false;
  int best_m = -1, min_k = 0, max_k = 0, pbase = 0;
  int x = 0, y = 0, d;
  const int has_n = q.has_n | t.has_n;
  for (d = 0; d < max_d; d++) {
    if (max_k - min_k > band_size) break;  // DWmatch.c:120-122
    if (min_k > max_k) {                   // empty band: the reference would read stale V entries from older d
      *err |= 1;
      break;
    }
    if (((max_k - min_k) >> 1) + 1 > cap) {
      *err |= 2;
      break;
    }
    int idx = 0;
    for (int k = min_k; k <= max_k; k += 2, idx++) {
      if (d == 0) {
        x = 0;  // V[1] of the calloc'd array
      } else {
        // DWmatch.c:125-130
        if (k == min_k) {
          x = Vp[(k + 1 - pbase) >> 1];
        } else if (k == max_k) {
          x = Vp[(k - 1 - pbase) >> 1] + 1;
        } else {
          int vm = Vp[(k - 1 - pbase) >> 1], vp = Vp[(k + 1 - pbase) >> 1];
          x = (vm < vp) ? vp : vm + 1;
        }
      }
      y = x - k;
      const int x1 = x, y1 = y;
      // snake, DWmatch.c:135-140
      for (;;) {
        int rem = q_len - x;
        if (t_len - y < rem) rem = t_len - y;
        if (rem <= 0) break;
        uint64_t df = fetch32(q, x) ^ fetch32(t, y);
        df = (df | (df >> 1)) & 0x5555555555555555ULL;
        int n = df ? (ctz64(df) >> 1) : 32;
        if (has_n) {
          uint32_t qn = q.has_n ? fetchn32(q, x) : 0u, tn = t.has_n ? fetchn32(t, y) : 0u;
          uint32_t m = qn ^ tn;  // N matches only N (nibble 0 == nibble 0 in the reference)
          // both-N positions must match whatever the 2-bit filler holds
          uint32_t both = qn & tn;
          if (both) {
            // clear diff bits under both-N positions
            uint64_t keep = 0;
            uint64_t dd = df;
            // expand `both` to even bit positions
            uint64_t b = both;
            b = (b | (b << 16)) & 0x0000FFFF0000FFFFULL;
            b = (b | (b << 8)) & 0x00FF00FF00FF00FFULL;
            b = (b | (b << 4)) & 0x0F0F0F0F0F0F0F0FULL;
            b = (b | (b << 2)) & 0x3333333333333333ULL;
            b = (b | (b << 1)) & 0x5555555555555555ULL;
            keep = ~b;
            dd &= keep;
            n = dd ? (ctz64(dd) >> 1) : 32;
          }
          if (m) {
            int nn = ctz32(m);
            if (nn < n) n = nn;
          }
        }
        if (n > rem) n = rem;
        x += n;
        y += n;
        if (n < 32) break;
      }
      if ((x - x1 > 16) && !start) {  // DWmatch.c:142-146
        r.q_bgn = x1;
        r.t_bgn = y1;
        start = true;
      }
      if ((uint32_t)(x - x1) > longest_match) {  // DWmatch.c:148-152
        longest_match = (uint32_t)(x - x1);
        r.q_m_end = x;
        r.t_m_end = y;
      }
      Vc[idx] = x;
      if (x + y > best_m) best_m = x + y;
      if (x >= q_len || y >= t_len) {  // DWmatch.c:161-164
        matched = true;
        break;
      }
    }
    if (matched) {  // DWmatch.c:185-194 (the band update in between cannot change the result)
      r.q_end = x;
      r.t_end = y;
      r.dist = d;
      r.m_size = (r.q_end - r.q_bgn + r.t_end - r.t_bgn + 2 * d) / 2;
      break;
    }
    // band trimming, DWmatch.c:168-183
    int new_min_k = max_k, new_max_k = min_k;
    idx = 0;
    for (int k2 = min_k; k2 <= max_k; k2 += 2, idx++) {
      int U = 2 * Vc[idx] - k2;
      if (U >= best_m - band_tolerance) {
        if (k2 < new_min_k) new_min_k = k2;
        if (k2 > new_max_k) new_max_k = k2;
      }
    }
    pbase = min_k;
    max_k = new_max_k + 1;
    min_k = new_min_k - 1;
    int *tmp = Vp;
    Vp = Vc;
    Vc = tmp;
  }
  if (!matched) {  // DWmatch.c:196-199
    r.q_bgn = 0;
    r.t_bgn = 0;
  }
  *out = r;
}

// ---------------------------------------------------------------------------------------------- ovlp_match, flattened
// Same algorithm and results as ovlp_match_core, restructured for SIMT: ONE loop whose every iteration performs one
// 32-base snake word-step, preceded by the cell set-up when a new diagonal starts and followed by the cell / edit-distance
// bookkeeping when the snake ended.  All lanes of a warp therefore run the same instruction stream whatever (d, k) they are
// at; only the short band-trimming loop at the end of an edit distance remains data dependent.
// V: caller scratch of 2*cap ints (two rows used alternately).
PGB_HD void ovlp_match_flat(const SeqView &q, int q_len, const SeqView &t, int t_len, int band_tolerance, int *V, int cap,
                            match_t *out, int *err) {
  match_t r;
  r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = r.t_m_end = r.q_m_end = 0;
  const int max_d = (int)(0.3 * (double)(q_len + t_len));
  const int band_size = band_tolerance * 2;
  const int has_n = q.has_n | t.has_n;
  uint32_t longest_match = 0;
  bool start = false, matched = false;
  int best_m = -1, min_k = 0, max_k = 0, pbase = 0;
  int d = 0, k = 0, idx = 0, x = 0, y = 0, x1 = 0, y1 = 0;
  int cur = 0;              // V + cur*cap is the row being written, the other row holds d-1
  bool new_cell = true;
  bool running = max_d > 0;  // d = 0: band is a single diagonal, always within band_size
  while (running) {
    int *Vc = V + cur * cap;
    const int *Vp = V + (cur ^ 1) * cap;
    if (new_cell) {  // DWmatch.c:125-131
      if (d == 0) x = 0;
      else if (k == min_k) x = Vp[(k + 1 - pbase) >> 1];
      else if (k == max_k) x = Vp[(k - 1 - pbase) >> 1] + 1;
      else {
        const int vm = Vp[(k - 1 - pbase) >> 1], vp = Vp[(k + 1 - pbase) >> 1];
        x = (vm < vp) ? vp : vm + 1;
      }
      y = x - k;
      x1 = x;
      y1 = y;
      new_cell = false;
    }
    // one snake word-step (DWmatch.c:135-140)
    int rem = q_len - x;
    if (t_len - y < rem) rem = t_len - y;
    bool more = false;
    if (rem > 0) {
      uint64_t df = fetch32(q, x) ^ fetch32(t, y);
      df = (df | (df >> 1)) & 0x5555555555555555ULL;
      if (has_n) {
        const uint32_t qn = q.has_n ? fetchn32(q, x) : 0u, tn = t.has_n ? fetchn32(t, y) : 0u;
        uint64_t both = qn & tn, one = qn ^ tn;  // N equals only N (nibble 0 == nibble 0 in the reference)
        both = (both | (both << 16)) & 0x0000FFFF0000FFFFULL; one = (one | (one << 16)) & 0x0000FFFF0000FFFFULL;
        both = (both | (both << 8)) & 0x00FF00FF00FF00FFULL;  one = (one | (one << 8)) & 0x00FF00FF00FF00FFULL;
        both = (both | (both << 4)) & 0x0F0F0F0F0F0F0F0FULL;  one = (one | (one << 4)) & 0x0F0F0F0F0F0F0F0FULL;
        both = (both | (both << 2)) & 0x3333333333333333ULL;  one = (one | (one << 2)) & 0x3333333333333333ULL;
        both = (both | (both << 1)) & 0x5555555555555555ULL;  one = (one | (one << 1)) & 0x5555555555555555ULL;
        df = (df & ~both) | one;
      }
      int n = df ? (ctz64(df) >> 1) : 32;
      if (n > rem) n = rem;
      x += n;
      y += n;
      more = (n == 32) && (rem > 32);
    }
    if (more) continue;
    // ---- the snake of cell (d, k) ended
    if ((x - x1 > 16) && !start) { r.q_bgn = x1; r.t_bgn = y1; start = true; }             // DWmatch.c:142-146
    if ((uint32_t)(x - x1) > longest_match) { longest_match = (uint32_t)(x - x1); r.q_m_end = x; r.t_m_end = y; }  // :148-152
    Vc[idx] = x;
    if (x + y > best_m) best_m = x + y;
    if (x >= q_len || y >= t_len) {  // :161-164, :185-194
      matched = true;
      r.q_end = x;
      r.t_end = y;
      r.dist = d;
      r.m_size = (r.q_end - r.q_bgn + r.t_end - r.t_bgn + 2 * d) / 2;
      break;
    }
    new_cell = true;
    if (k + 2 <= max_k) {
      k += 2;
      idx++;
      continue;
    }
    // ---- end of edit distance d: trim the band (DWmatch.c:168-183), advance d
    int new_min_k = max_k, new_max_k = min_k, i2 = 0;
    for (int k2 = min_k; k2 <= max_k; k2 += 2, i2++) {
      if (2 * Vc[i2] - k2 >= best_m - band_tolerance) {
        if (k2 < new_min_k) new_min_k = k2;
        if (k2 > new_max_k) new_max_k = k2;
      }
    }
    pbase = min_k;
    max_k = new_max_k + 1;
    min_k = new_min_k - 1;
    cur ^= 1;
    d++;
    k = min_k;
    idx = 0;
    if (d >= max_d) break;
    if (max_k - min_k > band_size) break;  // DWmatch.c:120-122
    if (min_k > max_k) { *err |= 1; break; }
    if (((max_k - min_k) >> 1) + 1 > cap) { *err |= 2; break; }
  }
  if (!matched) { r.q_bgn = 0; r.t_bgn = 0; }
  *out = r;
}

// ---------------------------------------------------------------------------------------------- ovlp_match, lean form
// The production form of the same algorithm (identical results; checked call-for-call against the reference in
// tests/hostsim and on the GPU).  Differences from ovlp_match_flat, all of them about cost per lane-iteration:
//  * operands are FORWARD views only: a reverse-strand operand reads the materialised reverse-complement image of the
//    packed reads (k_make_rc), so the snake needs no bit reversal (BREV/FLO run on the quarter-rate XU pipe) and the
//    four strand combinations share one instruction stream; reads that contain N take ovlp_match_flat instead;
//  * the high word of each operand is kept across word-steps of a snake (2 loads per continued step instead of 4), the
//    shift amounts are per-snake constants, and the mismatch position (count-trailing-zeros) is computed once per
//    snake, not once per step;
//  * the previous row of the band is streamed: V[k-1] of a cell is V[k+1] of the cell before it (1 load + 1 store per
//    cell), and the band trim scans inwards from both ends and stops at the diagonal that holds best_m, which is valid by
//    definition (0 loads in the common 3-diagonal band), instead of re-reading the whole row;
//  * cell set-up is folded into the tail of the previous cell, so an iteration is "word-step, then (if the snake ended)
//    one tail block" - one divergent region instead of three.
PGB_HD uint32_t funnel_r32(uint32_t lo, uint32_t hi, uint32_t sh) {  // low 32 bits of (hi:lo) >> sh, sh in [0,31]
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}
// 32 bases (64 bits) starting s2 bits (even, 0..62) into the 128-bit value B:A
PGB_HD uint64_t window64(uint64_t A, uint64_t B, uint32_t s2) {
  const uint32_t a0 = (uint32_t)A, a1 = (uint32_t)(A >> 32), b0 = (uint32_t)B, b1 = (uint32_t)(B >> 32);
  const bool up = s2 >= 32;
  const uint32_t w0 = up ? a1 : a0, w1 = up ? b0 : a1, w2 = up ? b1 : b0, sh = s2 & 31;
  return (uint64_t)funnel_r32(w0, w1, sh) | ((uint64_t)funnel_r32(w1, w2, sh) << 32);
}
// qw / tw: first packed word of the read in the image of the operand's strand; qo / to: base offset of logical base 0.
// V: caller scratch of 2*cap ints, cap >= band_tolerance + 2.
// PRE: load V[d-1][k+3] one cell ahead; TRIMREG: serve the first two trim steps per side from registers
// (Measured and dropped in round 2, profiles/r2_align.md: an L2 prefetch of the next operand line, 20.6 vs 20.2 ms, and a
// persistent form in which a lane takes the next alignment from a queue when its own ends, 22.6 ms.)
template <bool PRE, bool TRIMREG>
PGB_HD void ovlp_match_lean_t(const uint64_t *qw, uint32_t qo, int q_len, const uint64_t *tw, uint32_t to, int t_len,
                              int band_tolerance, int *V, int cap, match_t *out) {
  match_t r;
  r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = r.t_m_end = r.q_m_end = 0;
  const int max_d = (int)(0.3 * (double)(q_len + t_len));  // DWmatch.c:96
  const int band_size = band_tolerance * 2;
  uint32_t longest_match = 0;
  bool start = false, matched = false;
  int best_m = -1, k_best = 0, u_first = 0;
  int min_k = 0, max_k = 0, d = 0, k = 0, idx = 0, x = 0, x1 = 0;
  int *Vc = V, *Vp = V + cap;  // row being written / row of d-1
  int vcarry = 0;  // V[d-1][k+1] read by the previous cell of this row = V[d-1][k-1] of the current cell
  int vpre = 0;    // V[d-1][k+3], loaded one cell ahead: the band-row load of a cell is off its critical path
  int poff = 0;    // index of V[d-1][min_k+1] in the previous row's storage
  int u1 = 0, u2 = 0;             // x+y of the row's 2nd and 3rd cell   } the band trim normally inspects one or two cells per
  int ul0 = 0, ul1 = 0, ul2 = 0;  // x+y of the last three cells so far  } side: served from registers, not from the row array
  const uint64_t *qp = qw + (qo >> 5), *tp = tw + (to >> 5);
  uint32_t qs = (qo & 31) * 2, ts = (to & 31) * 2;
  uint64_t qA = 0, qB = 0, tA = 0, tB = 0;
  bool fresh = true;
  bool running = max_d > 0;
  // ONE loop, no continue / break: every iteration is a word-step followed by either "stay in the snake" or the tail of
  // the cell, and the lanes of a warp reconverge at the end of every iteration (a `continue` makes the compiler build an
  // inner snake loop at whose exit all lanes wait for the longest snake of the warp).
  while (running) {
#include "ovlp_match_lean_body.inc"
  }
  if (!matched) {  // DWmatch.c:196-199
    r.q_bgn = 0;
    r.t_bgn = 0;
  }
  *out = r;
}
PGB_HD void ovlp_match_lean(const uint64_t *qw, uint32_t qo, int q_len, const uint64_t *tw, uint32_t to, int t_len,
                            int band_tolerance, int *V, int cap, match_t *out) {
  ovlp_match_lean_t<true, true>(qw, qo, q_len, tw, to, t_len, band_tolerance, V, cap, out);
}

// ---------------------------------------------------------------------------------------------- mm_sketch (exact automaton)
// One read, forward strand, sequential — a literal restatement of src/mm_sketch.c:84-150 for is_hpc == 0 over the
// packed representation.  Used (a) for reads the tiled fast kernel flags (N, hash ties, palindrome-dense halos) and
// (b) by the single-read C-ABI mm_sketch().  ring_x/ring_p: caller scratch of w entries each.
// emit(x, y) is called in output order.
template <class Emit>
PGB_HD void sketch_exact(const uint64_t *w, const uint32_t *nm, uint64_t word_off, int len, int wsz, int k, uint32_t rid,
                         uint64_t *ring_x, uint32_t *ring_p, Emit &&emit) {
  const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
  const uint64_t XMAX = ~0ULL;
  const uint32_t PMAX = ~0u;
  uint64_t kmer0 = 0, kmer1 = 0;
  int l = 0, buf_pos = 0, min_pos = 0;
  uint64_t min_x = XMAX;
  uint32_t min_p = PMAX;
  for (int j = 0; j < wsz; j++) {
    ring_x[j] = XMAX;
    ring_p[j] = PMAX;
  }
  const uint64_t ridhi = (uint64_t)rid << 32;
  uint64_t cw = 0;
  uint32_t cn = 0;
  for (int i = 0; i < len; ++i) {
    if ((i & 31) == 0) {
      cw = w[word_off + (i >> 5)];
      cn = nm ? nm[word_off + (i >> 5)] : 0u;
    }
    int c = (int)((cw >> (2 * (i & 31))) & 3);
    int isn = (int)((cn >> (i & 31)) & 1);
    uint64_t info_x = XMAX;
    uint32_t info_p = PMAX;
    if (!isn) {
      kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
      kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
      if (kmer0 == kmer1) continue;  // mm_sketch.c:104-105, before ++l / ring write / buf_pos++
      int z = kmer0 < kmer1 ? 0 : 1;
      ++l;
      if (l >= k) {  // kmer_span == k < 256 always when is_hpc == 0
        info_x = hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)k;
        info_p = (uint32_t)i << 1 | (uint32_t)z;
      }
    } else {
      l = 0;
    }
    ring_x[buf_pos] = info_x;
    ring_p[buf_pos] = info_p;
    if (l == wsz + k - 1 && min_x != XMAX) {  // mm_sketch.c:116-125
      for (int j = buf_pos + 1; j < wsz; ++j)
        if (min_x == ring_x[j] && ring_p[j] != min_p) emit(ring_x[j], ridhi | ring_p[j]);
      for (int j = 0; j < buf_pos; ++j)
        if (min_x == ring_x[j] && ring_p[j] != min_p) emit(ring_x[j], ridhi | ring_p[j]);
    }
    if (info_x <= min_x) {  // mm_sketch.c:126-128
      if (l >= wsz + k && min_x != XMAX) emit(min_x, ridhi | min_p);
      min_x = info_x;
      min_p = info_p;
      min_pos = buf_pos;
    } else if (buf_pos == min_pos) {  // mm_sketch.c:129-147
      if (l >= wsz + k - 1 && min_x != XMAX) emit(min_x, ridhi | min_p);
      min_x = XMAX;
      for (int j = buf_pos + 1; j < wsz; ++j)
        if (min_x >= ring_x[j]) min_x = ring_x[j], min_p = ring_p[j], min_pos = j;
      for (int j = 0; j <= buf_pos; ++j)
        if (min_x >= ring_x[j]) min_x = ring_x[j], min_p = ring_p[j], min_pos = j;
      if (l >= wsz + k - 1 && min_x != XMAX) {
        for (int j = buf_pos + 1; j < wsz; ++j)
          if (min_x == ring_x[j] && min_p != ring_p[j]) emit(ring_x[j], ridhi | ring_p[j]);
        for (int j = 0; j <= buf_pos; ++j)
          if (min_x == ring_x[j] && min_p != ring_p[j]) emit(ring_x[j], ridhi | ring_p[j]);
      }
    }
    if (++buf_pos == wsz) buf_pos = 0;
  }
  if (min_x != XMAX) emit(min_x, ridhi | min_p);  // mm_sketch.c:150
}

// ---------------------------------------------------------------------------------------------- mm_sketch (exact, one segment)
// The automaton's state after a base depends only on the last w ring slots and on l (capped at w+k), so a read can be cut
// into segments that are replayed independently: start `start_pos` early enough that by position seg_lo (a) at least w
// slots were written (ring identical to the full run) and (b) l is either >= w+k in both runs or was reset by an N inside
// the warm-up (identical in both runs).  Records are attributed by POSITION: only minimizers with seg_lo <= pos < seg_hi
// are reported; the replay continues until w slots past seg_hi (an entry leaves the ring, and can no longer be emitted,
// w slots after it entered) or to the end of the read (final flush, mm_sketch.c:150).
// Returns false when start_pos > 0 and the warm-up condition was not met (caller retries with start_pos = 0).
template <class Emit>
PGB_HD bool sketch_exact_range(const uint64_t *w, const uint32_t *nm, uint64_t word_off, int len, int wsz, int k, uint32_t rid,
                               int start_pos, int seg_lo, int seg_hi, uint64_t *ring_x, uint32_t *ring_p, Emit &&emit, int rs = 1) {
  // (slot j of the ring lives at index j * rs: rs = 1 for a private ring, rs = CTA size for rings interleaved in shared memory)
  const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
  const uint64_t XMAX = ~0ULL;
  const uint32_t PMAX = ~0u;
  uint64_t kmer0 = 0, kmer1 = 0;
  int l = 0, buf_pos = 0, min_pos = 0;
  uint64_t min_x = XMAX;
  uint32_t min_p = PMAX;
  for (int j = 0; j < wsz; j++) {
    ring_x[(j) * rs] = XMAX;
    ring_p[(j) * rs] = PMAX;
  }
  const uint64_t ridhi = (uint64_t)rid << 32;
  const uint32_t plo = (uint32_t)seg_lo << 1, phi = (uint32_t)seg_hi << 1;
#define PGB_EMIT_IF(x_, p_) do { uint32_t pp_ = (p_); if (pp_ >= plo && pp_ < phi) emit((x_), ridhi | pp_); } while (0)
  uint64_t cw = 0;
  uint32_t cn = 0;
  int slots_before = 0, slots_after = 0;
  bool saw_n = false, warm_checked = start_pos == 0;
  int i = start_pos;
  if (i < len) {
    cw = w[word_off + (i >> 5)];
    cn = nm ? nm[word_off + (i >> 5)] : 0u;
  }
  for (; i < len; ++i) {
    if ((i & 31) == 0) {
      cw = w[word_off + (i >> 5)];
      cn = nm ? nm[word_off + (i >> 5)] : 0u;
    }
    if (!warm_checked && i >= seg_lo) {
      if (!(slots_before >= wsz && (saw_n || l >= wsz + k))) { return false; }
      warm_checked = true;
    }
    if (slots_after >= wsz) break;  // every entry with pos < seg_hi has left the ring
    int c = (int)((cw >> (2 * (i & 31))) & 3);
    int isn = (int)((cn >> (i & 31)) & 1);
    uint64_t info_x = XMAX;
    uint32_t info_p = PMAX;
    if (!isn) {
      kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
      kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
      if (kmer0 == kmer1) continue;
      int z = kmer0 < kmer1 ? 0 : 1;
      ++l;
      if (l >= k) {
        info_x = hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)k;
        info_p = (uint32_t)i << 1 | (uint32_t)z;
      }
    } else {
      l = 0;
      saw_n = true;
    }
    if (i < seg_lo) slots_before++;
    if (i >= seg_hi) slots_after++;
    ring_x[(buf_pos) * rs] = info_x;
    ring_p[(buf_pos) * rs] = info_p;
    if (l == wsz + k - 1 && min_x != XMAX) {
      for (int j = buf_pos + 1; j < wsz; ++j)
        if (min_x == ring_x[(j) * rs] && ring_p[(j) * rs] != min_p) PGB_EMIT_IF(ring_x[(j) * rs], ring_p[(j) * rs]);
      for (int j = 0; j < buf_pos; ++j)
        if (min_x == ring_x[(j) * rs] && ring_p[(j) * rs] != min_p) PGB_EMIT_IF(ring_x[(j) * rs], ring_p[(j) * rs]);
    }
    if (info_x <= min_x) {
      if (l >= wsz + k && min_x != XMAX) PGB_EMIT_IF(min_x, min_p);
      min_x = info_x;
      min_p = info_p;
      min_pos = buf_pos;
    } else if (buf_pos == min_pos) {
      if (l >= wsz + k - 1 && min_x != XMAX) PGB_EMIT_IF(min_x, min_p);
      min_x = XMAX;
      for (int j = buf_pos + 1; j < wsz; ++j)
        if (min_x >= ring_x[(j) * rs]) min_x = ring_x[(j) * rs], min_p = ring_p[(j) * rs], min_pos = j;
      for (int j = 0; j <= buf_pos; ++j)
        if (min_x >= ring_x[(j) * rs]) min_x = ring_x[(j) * rs], min_p = ring_p[(j) * rs], min_pos = j;
      if (l >= wsz + k - 1 && min_x != XMAX) {
        for (int j = buf_pos + 1; j < wsz; ++j)
          if (min_x == ring_x[(j) * rs] && min_p != ring_p[(j) * rs]) PGB_EMIT_IF(ring_x[(j) * rs], ring_p[(j) * rs]);
        for (int j = 0; j <= buf_pos; ++j)
          if (min_x == ring_x[(j) * rs] && min_p != ring_p[(j) * rs]) PGB_EMIT_IF(ring_x[(j) * rs], ring_p[(j) * rs]);
      }
    }
    if (++buf_pos == wsz) buf_pos = 0;
  }
  if (!warm_checked) {  // the read ended inside the warm-up (seg_lo >= len cannot happen for a valid segment)
    if (!(slots_before >= wsz && (saw_n || l >= wsz + k))) return false;
  }
  if (i >= len && min_x != XMAX) PGB_EMIT_IF(min_x, min_p);  // final flush only when the read's end was reached
#undef PGB_EMIT_IF
  return true;
}
PGB_HD int sketch_warmup_len(int wsz, int k) { return 2 * (wsz + k) + 64; }

// ---------------------------------------------------------------------------------------------- mm_reduce
// Window pick for the element at in-read offset o (o >= rs-1), src/shmr_reduce.c:33-50,79-88: the ring slot of
// the element with offset t is t % rs; slots are scanned 0..rs-1 with strict '<' on x>>8, so ties go to the lowest
// slot.  `a` points at the read's first mmer.  Returns the in-read offset of the pick.
PGB_HD uint32_t reduce_pick(const mm128 *a, uint32_t o, uint32_t rs) {
  uint32_t best_t = 0;
  uint64_t best_h = 0;
  for (uint32_t s = 0; s < rs; s++) {
    uint32_t t = o - ((o + rs - s) % rs);  // the largest offset <= o congruent to s
    uint64_t h = a[t].x >> 8;
    if (s == 0 || h < best_h) {
      best_h = h;
      best_t = t;
    }
  }
  return best_t;
}

// ---------------------------------------------------------------------------------------------- build_map helpers
// rev(y, x) of src/shmr_utils.c:376-395
PGB_HD uint64_t rev_y(uint64_t y, uint64_t x, uint32_t rlen) {
  uint32_t span = (uint32_t)(x & 0xFF);
  uint32_t pos = (uint32_t)((y & 0xFFFFFFFFULL) >> 1) + 1;
  uint32_t rpos = rlen - pos + span - 1;
  return ((y & 0xFFFFFFFF00000001ULL) | (uint64_t)(uint32_t)(rpos << 1)) ^ 0x1ULL;
}
// "too close" test of src/shmr_utils.c:332 (28-bit masked positions, u64 subtraction)
PGB_HD bool pair_far_enough(uint64_t y0, uint64_t y1) {
  return (((y1 >> 1) & 0xFFFFFFFULL) - ((y0 >> 1) & 0xFFFFFFFULL)) >= 100ULL;
}

// ---------------------------------------------------------------------------------------------- greedy bucket scan
// One (x0,x1) bucket, records already in the reference's post-qsort order (stable, descending position).
// Restates src/shmr_overlap.c:52-180 with the process-wide rid_pairs table replaced by a *time-stamped* table:
// an entry is (rank of the bucket that first accepted the pair)<<2 | type, so that each bucket can be replayed
// independently against "what was in rid_pairs when the reference reached this bucket" (DESIGN.md §replay).
//
// Ctx must provide:
//   uint32_t rlen(uint32_t rid)
//   void     pair_get(uint64_t ridp, uint64_t *vold, uint64_t *vnew)   value from the previous iteration and value being
//                                               built in this iteration (NONE = ~0), one table probe for both
//   void     pair_set(uint64_t ridp, uint64_t v)   atomic-min into the table being built
//   bool     aln_get(uint32_t i, uint32_t j, match_t *m)     true if the alignment of records (i,j) is known
//   void     aln_request(uint32_t i, uint32_t j, uint32_t rid0, uint32_t start0, uint32_t strand0,
//                        uint32_t rid1, uint32_t strand1)
//   void     emit(const ovlp_rec &)             called only when `do_emit`
// Returns the number of accepted records; *n_unknown counts alignments that were predicted, not known.
PGB_HD void predict_match(uint32_t rlen0, uint32_t rlen1, uint32_t start0, match_t *m) {
  // an overlap that extends from (0,0) of the shifted frame to the end of the shorter operand
  uint32_t slen0 = rlen0 - start0;
  uint32_t e = slen0 < rlen1 ? slen0 : rlen1;
  m->m_size = (int32_t)e;
  m->dist = 0;
  m->q_bgn = 0;
  m->t_bgn = 0;
  m->q_end = (int32_t)e;
  m->t_end = (int32_t)e;
  m->q_m_end = (int32_t)e;
  m->t_m_end = (int32_t)e;
}

// Acceptance test and overlap type of src/shmr_overlap.c:134-160 for the alignment m of records (i, j): slen0 / slen1 are
// the operand lengths handed to ovlp_match.  Returns false when the alignment is rejected.
PGB_HD bool classify_match(const match_t &m, uint32_t rlen0, uint32_t rlen1, uint32_t slen0, uint32_t slen1, uint32_t *type) {
  const int64_t q_bgn = m.q_bgn, q_end = m.q_end, t_bgn = m.t_bgn, t_end = m.t_end;
  int64_t dq = (int64_t)slen0 - q_end, dt = (int64_t)slen1 - t_end;  // abs() on int in the reference, operands fit easily
  if (dq < 0) dq = -dq;
  if (dt < 0) dt = -dt;
  if (!(q_bgn < READ_END_FUZZINESS && t_bgn < READ_END_FUZZINESS && (dq < READ_END_FUZZINESS || dt < READ_END_FUZZINESS) && q_end > 500 &&
        t_end > 500))
    return false;
  int64_t c0 = (int64_t)rlen0 - (q_end - q_bgn), c1 = (int64_t)rlen1 - (t_end - t_bgn);
  if (c0 < 0) c0 = -c0;
  if (c1 < 0) c1 = -c1;
  if (c0 < READ_END_FUZZINESS * 2 || c1 < READ_END_FUZZINESS * 2) *type = rlen0 >= rlen1 ? OVL_CONTAINS : OVL_CONTAINED;  // :142-154
  else *type = OVL_OVERLAP;
  return true;
}

template <class Ctx>
PGB_HD uint32_t replay_bucket(Ctx &c, uint32_t rank, const uint64_t *y0s, const uint8_t *dirs, uint32_t n,
                              uint8_t *contained, uint32_t bestn, bool do_emit, uint32_t *n_unknown) {
  const uint64_t NONE = ~0ULL;
  uint32_t n_acc = 0, n_unk = 0;
  for (uint32_t i = 0; i < n; i++) contained[i] = 0;
  for (uint32_t k0 = n - 1; k0 > 0; k0--) {
    const uint32_t i = k0 - 1;
    if (contained[i]) continue;
    const uint64_t y0 = y0s[i];
    const uint32_t rid0 = (uint32_t)(y0 >> 32);
    const uint32_t pos0 = (uint32_t)((y0 & 0xFFFFFFFFULL) >> 1) + 1;
    const uint32_t rlen0 = c.rlen(rid0);
    const uint32_t strand0 = dirs[i];
    uint32_t overlap_count = 0;
    for (uint32_t j = i + 1; j < n && overlap_count < bestn; j++) {
      if (contained[j]) continue;
      const uint64_t y1 = y0s[j];
      const uint32_t rid1 = (uint32_t)(y1 >> 32);
      if (rid0 == rid1) continue;
      const uint64_t ridp = rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
      {
        uint64_t v, vnew;
        c.pair_get(ridp, &v, &vnew);
        bool hit = (v != NONE) && ((uint32_t)(v >> 2) < rank);
        if (!hit) {
          v = vnew;
          hit = (v != NONE) && ((uint32_t)(v >> 2) <= rank);
        }
        if (hit) {
          if ((v & 3) == OVL_OVERLAP) overlap_count += 1;
          continue;
        }
      }
      const uint32_t pos1 = (uint32_t)((y1 & 0xFFFFFFFFULL) >> 1) + 1;
      const uint32_t rlen1 = c.rlen(rid1);
      const uint32_t strand1 = dirs[j];
      const uint32_t start0 = pos0 - pos1;
      const uint32_t slen0 = rlen0 - pos0 + pos1;
      const uint32_t slen1 = rlen1;
      match_t m;
      if (!c.aln_get(i, j, &m)) {
        n_unk++;
        c.aln_request(i, j, rid0, start0, strand0, rid1, strand1);
        predict_match(rlen0, rlen1, start0, &m);
      }
      uint32_t type;
      if (classify_match(m, rlen0, rlen1, slen0, slen1, &type)) {  // src/shmr_overlap.c:134-160
        if (type == OVL_CONTAINS) contained[j] = 1;
        else if (type == OVL_CONTAINED) contained[i] = 1;
        else overlap_count++;
        c.pair_set(ridp, ((uint64_t)rank << 2) | type);
        if (do_emit) {
          ovlp_rec o;
          o.y0 = y0;
          o.y1 = y1;
          o.rl0 = rlen0;
          o.rl1 = rlen1;
          o.strand0 = (uint8_t)strand0;
          o.strand1 = (uint8_t)strand1;
          o.ovlp_type = (uint8_t)type;
          o.pad0 = 0;
          o.match = m;
          o.pad1 = 0;
          c.emit(n_acc, o);
        }
        n_acc++;
      }
      if (contained[i]) break;  // :176
    }
  }
  *n_unknown = n_unk;
  return n_acc;
}

}  // namespace pgb
