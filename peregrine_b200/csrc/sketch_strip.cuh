// sketch_strip.cuh — (w,k)-minimizers of a read by ONE WARP that walks the read in strips of 512 positions
// (the fast path of mm_sketch, src/mm_sketch.c:70-151; replaces the block-tiled kernel of round 1).
//
// Why a strip per warp.  The round-1 kernel tiled reads over CTAs: 3 % of the positions were hashed twice (halo), the van
// Herk block scans kept a third of the CTA busy between barriers, and 152 thread-instructions were spent per base
// (profiles/r1g_ncu.md).  Here lane l of the warp owns positions [512 c + 16 l, 512 c + 16 l + 16) of strip c:
//   1. it extracts its first k-mer pair from the packed words, rolls 15 more bases through it and hashes the canonical
//      k-mer of every position (32-bit arithmetic when k <= 16: hash64 masks to 2k bits after every step);
//   2. suffix minima of its 16 hashes (value, in-row offset of the rightmost occurrence, "occurs twice" bit) go to a ring of rows
//      in shared memory (64 rows of 16 = the last two strips), one row per lane per strip;
//   3. a window [e-w+1, e] ending in the lane's segment is  suffix(row of e-w+1)  +  the whole rows in between  +  prefix of
//      the own segment: the minimum over the whole rows is computed ONCE per lane and strip (two variants, because the
//      16 window starts of a lane straddle at most two rows), the prefix is a running minimum in registers, so a window
//      costs one shared-memory load and two combines; no barrier, no halo: the ring carries the history into the next strip;
//   4. a position is emitted when it becomes the window's (rightmost) arg-min; emitted records are staged per lane and
//      leave in position order after one warp scan per strip.
// Exactness: on tie-free, N-free data "rightmost arg-min of every full window, reported when it changes" IS the reference
// output.  Everything else is detected and handed to the exact automaton (k_sketch_exact_seg):
//  * whole reads: reads with N, reads shorter than one window, a lane with more than SS_STAGE records in a strip, a read that
//    overflows its record budget, more than SS_MAXPAL palindromic k-mers, a bad window beyond strip 63;
//  * single strips ("bad" strips, bit c of the read's mask): a strip in which some evaluated window's minimum occurs twice, or
//    holds two palindromic k-mers, or a palindromic k-mer precedes the first full window (strip 0).  The automaton's state is a
//    function of the last w window slots, so a record at position p is decided by the windows that contain p alone: every
//    position inside a bad window ("tainted": at most w + SS_MAXPAL positions back from the window's end) is redone by the
//    automaton, the fast path's records at tainted positions are dropped, all its other records stand (the walk continues
//    through a bad strip: the minimum carried into the next strip is that of a window, whatever happened before it).
// Palindromic k-mers (src/mm_sketch.c:104-105: they occupy NO window slot) are rare (4^-k/2 per position) but hit 20 % of
// 15 kb reads at k = 16: a strip within w positions of one takes a slower, general window evaluation in which the windows
// that contain the palindrome reach one position further back.
#pragma once
#include "shimmer_core.cuh"
#include "sketch_tile.cuh"

namespace pgb {

enum { SS_SPL = 16, SS_STRIP = 512, SS_ROWS = 64, SS_ROWPAD = 17, SS_WARPS = 4, SS_STAGE = 8, SS_MAXPAL = 8 };

template <class HT>
struct SsWarpSmem {
  HT sv[SS_ROWS * SS_ROWPAD];        // suffix minimum of the row from this offset on; row = (position >> 4) & 63
  uint8_t sa[SS_ROWS * SS_ROWPAD];   // ... its offset in the row (rightmost on ties)
  uint16_t st[SS_ROWS];              // ... bit o: that minimum occurs twice in [o, 16)
  HT stv[32 * SS_STAGE];             // staged records of the strip: hash
  uint32_t stp[32 * SS_STAGE];       // ... position
  int pal[SS_MAXPAL];                // positions of the most recent palindromic k-mers (ring)
  unsigned long long bad;            // strips (bit c = positions [512 c, 512 c + 512)) with a window the fast path cannot decide; lane 0 writes
};

// a running minimum: value, "occurs twice" flag, position
template <class HT>
struct SsMin {
  HT v;
  bool t;
  int p;
};
// r lies to the RIGHT of l: ties go to r (the reference keeps the newest of equal k-mers, mm_sketch.c:126,135-138)
template <class HT>
__device__ __forceinline__ SsMin<HT> ss_combine(const SsMin<HT> &l, const SsMin<HT> &r) {
  SsMin<HT> o;
  const bool lt = r.v < l.v, eq = r.v == l.v;
  o.v = lt ? r.v : l.v;
  o.t = eq || (lt ? r.t : l.t);
  o.p = (lt || eq) ? r.p : l.p;
  return o;
}

// 2k-bit window of the packed sequence -> the reference's k-mer pair.  V holds 32 bases, earliest in the low bits (base b at
// bits 2b); R is V with its 32 two-bit groups reversed.  The k bases that END at base index (k-1+i) of V are, earliest base in
// the low bits, W = (V >> 2i) & mask: the reverse-complement k-mer of mm_sketch.c:101 is ~W, the forward k-mer (earliest base in
// the HIGH bits, :100) is the same group range of R.
// (k <= 16: everything a lane needs, k-1+16 <= 31 bases, sits in one 64-bit value; k > 16 takes the 128-bit form below.)
template <class HT>
__device__ __forceinline__ void ss_kmers(uint64_t V, uint64_t R, int i, int k, HT mask, HT *kmer0, HT *kmer1) {
  *kmer1 = (HT)(~(V >> (2 * i))) & mask;
  *kmer0 = (HT)(R >> (64 - 2 * i - 2 * k)) & mask;
}
// 64 bits starting s bits (0..127) into hi:lo
__device__ __forceinline__ uint64_t ss_field128(uint64_t lo, uint64_t hi, int s) {
  return s == 0 ? lo : (s < 64 ? (lo >> s) | (hi << (64 - s)) : hi >> (s - 64));
}
// the same over 64 bases: Vb:Va hold bases 0..63 (earliest in the low bits of Va); the 128-bit value with all 64 two-bit groups
// reversed has rev2(Vb) in its low and rev2(Va) in its high half
__device__ __forceinline__ void ss_kmers128(uint64_t Va, uint64_t Vb, uint64_t RVa, uint64_t RVb, int i, int k, uint64_t mask, uint64_t *kmer0,
                                            uint64_t *kmer1) {
  *kmer1 = ~ss_field128(Va, Vb, 2 * i) & mask;
  *kmer0 = ss_field128(RVb, RVa, 128 - 2 * i - 2 * k) & mask;
}

// per-read state of the walking warp
template <class HT>
struct SsState {
  const uint64_t *w;
  int64_t base0;
  int len, wsz, k, e_ff, s_eval;
  HT mask;
  uint64_t mask64;
  uint32_t n_out, flags;
  HT carry_v;  // minimum of the last window of the previous strip
  bool have_carry;
  int n_pal, last_pal;
  uint64_t ridhi, cap;
  mm128 *out;
};

// One strip.  PLAIN: every position has a complete k-mer, every window that ends in the strip is a full window after the first
// one and no palindromic k-mer is within reach - no per-position validity test is compiled in.
template <class HT, bool PLAIN>
__device__ __forceinline__ void ss_strip(SsState<HT> &S, SsWarpSmem<HT> &sh, const int cp, const int lane, const HT (&h)[SS_SPL], const uint32_t vm,
                                         const bool slow) {
  constexpr uint32_t FULL = 0xffffffffu;
  const HT MAXV = (HT) ~(HT)0;
  auto row_ix = [](int r) -> int { return (r & (SS_ROWS - 1)) * SS_ROWPAD; };
  const int pos0 = cp + SS_SPL * lane, wsz = S.wsz, len = S.len, e_ff = S.e_ff, s_eval = S.s_eval;
  const int r_e = pos0 >> 4;
  bool tie = false;  // this lane saw a window of the strip that the fast path cannot decide (tie, or palindromes beyond its reach)
  // ---------------- suffix minima of the lane's segment -> its row of the ring
  {
    const int rb = row_ix(r_e);
    HT run = h[SS_SPL - 1];
    bool rt = false;
    int ra = SS_SPL - 1;
    uint32_t tm = 0;
    sh.sv[rb + SS_SPL - 1] = run;
    sh.sa[rb + SS_SPL - 1] = (uint8_t)ra;
#pragma unroll
    for (int j = SS_SPL - 2; j >= 0; j--) {
      const bool lt = h[j] < run, eq = h[j] == run;
      rt = eq || (!lt && rt);
      ra = lt ? j : ra;
      run = lt ? h[j] : run;
      tm |= (uint32_t)rt << j;
      sh.sv[rb + j] = run;
      sh.sa[rb + j] = (uint8_t)ra;
    }
    sh.st[r_e & (SS_ROWS - 1)] = (uint16_t)tm;
  }
  __syncwarp();
  // ---------------- the windows that end in the lane's segment
  uint32_t n_st = 0;        // staged records of this lane
  HT first_v = MAXV;        // minimum of the lane's first evaluated window (its emission is decided after the shuffle)
  HT first_h = MAXV;        // hash of the first evaluated position
  HT last_v = MAXV;         // minimum of its last evaluated window
  bool first_cond = false;  // stage entry 0 is that first window's record, valid only if it differs from the left neighbour's minimum
  bool have_prev = false;
  if (PLAIN || (pos0 < len && pos0 + SS_SPL - 1 >= s_eval)) {
    const int lo0 = pos0 - wsz + 1, r_lo0 = lo0 >> 4;
    // minimum over the whole rows between the window's first row and the lane's own row: with and without row r_lo0 + 1
    SsMin<HT> m_short, m_long;
    m_short.v = MAXV; m_short.t = false; m_short.p = 0;
    auto row_total = [&](int r) -> SsMin<HT> {
      SsMin<HT> t;
      const int ix = row_ix(r);
      t.v = sh.sv[ix];
      t.t = (sh.st[r & (SS_ROWS - 1)] & 1u) != 0;
      t.p = r * SS_SPL + (int)sh.sa[ix];
      return t;
    };
    if (!slow) {
      for (int r = (r_lo0 + 2 > 0 ? r_lo0 + 2 : 0); r < r_e; r++) m_short = ss_combine(m_short, row_total(r));
      m_long = m_short;
      if (r_lo0 + 1 >= 0 && r_lo0 + 1 < r_e) m_long = ss_combine(row_total(r_lo0 + 1), m_short);
    }
    SsMin<HT> pre;
    pre.v = MAXV; pre.t = false; pre.p = 0;
    HT prev = MAXV;
#pragma unroll
    for (int j = 0; j < SS_SPL; j++) {
      const int e = pos0 + j;
      {
        SsMin<HT> c;
        c.v = h[j]; c.t = false; c.p = e;
        pre = j == 0 ? c : ss_combine(pre, c);
      }
      const bool ev = PLAIN || (e < len && e >= s_eval && ((vm >> j) & 1u));
      if (ev) {
        SsMin<HT> win;
        if (PLAIN || !slow) {
          const int lo = e - wsz + 1, ix = row_ix(lo >> 4) + (lo & 15);
          SsMin<HT> s_;
          s_.v = sh.sv[ix];
          s_.t = ((sh.st[(lo >> 4) & (SS_ROWS - 1)] >> (lo & 15)) & 1u) != 0;
          s_.p = -1;  // its position is looked up only if this part holds the window's minimum
          win = ss_combine(ss_combine(s_, (lo >> 4) == r_lo0 ? m_long : m_short), pre);
          if (win.p < 0) win.p = (lo & ~15) + (int)sh.sa[ix];
        } else {
          // general form: the window holds w SLOTS; a palindromic k-mer inside it is no slot, so the window reaches one
          // position further back (two of them: exact automaton)
          int cpal = 0;
          bool edge_pal = false;
          const int np = S.n_pal < SS_MAXPAL ? S.n_pal : SS_MAXPAL;
          for (int t = 0; t < np; t++) {
            const int q = sh.pal[t];
            cpal += (q >= e - wsz + 1 && q <= e);
            edge_pal |= (q == e - wsz);
          }
          if (cpal > 1 || (cpal == 1 && edge_pal)) tie = true;
          const int lo = e - wsz + 1 - (cpal ? 1 : 0);
          const int r_lo = lo >> 4, ix = row_ix(r_lo) + (lo & 15);
          win.v = sh.sv[ix];
          win.t = ((sh.st[r_lo & (SS_ROWS - 1)] >> (lo & 15)) & 1u) != 0;
          win.p = (lo & ~15) + (int)sh.sa[ix];
          for (int r = r_lo + 1; r < r_e; r++) win = ss_combine(win, row_total(r));
          win = ss_combine(win, pre);
        }
        tie |= win.t && (PLAIN || win.v != MAXV);
        // a new element equal to the previous window's minimum: a new minimizer the value alone cannot show (conservative)
        if (have_prev) tie |= (h[j] == prev) && (PLAIN || prev != MAXV);
        if (PLAIN || e >= e_ff) {
          const bool is_first = !have_prev;
          if ((!PLAIN && e == e_ff) || is_first || win.v != prev) {
            if (n_st < SS_STAGE) {
              sh.stv[lane * SS_STAGE + n_st] = win.v;
              sh.stp[lane * SS_STAGE + n_st] = (uint32_t)win.p;
            }
            if (is_first && (PLAIN || e != e_ff)) first_cond = true;
            n_st++;
          }
        }
        if (!have_prev) { first_v = win.v; first_h = h[j]; have_prev = true; }
        prev = win.v;
        last_v = win.v;
      }
    }
  }
  // ---------------- resolve the lanes' first windows against their left neighbours, then write in position order
  {
    const uint32_t have = __ballot_sync(FULL, have_prev);
    const uint32_t below = have & ((1u << lane) - 1u);
    const int src = below ? 31 - __clz((int)below) : 0;
    HT left = (HT)__shfl_sync(FULL, last_v, src);
    bool have_left = below != 0;
    if (!below) { left = S.carry_v; have_left = S.have_carry; }
    uint32_t skip = 0;
    if (have_prev && have_left) {
      if (first_cond && first_v == left) skip = 1;   // same minimum as the window before it: not a new minimizer
      tie |= (first_h == left) && left != MAXV;      // (the lane's first element against the previous window's minimum, as inside the lane)
    }
    const uint32_t top = have ? 31 - __clz((int)have) : 0;
    const HT new_carry = (HT)__shfl_sync(FULL, last_v, top);
    if (have) { S.carry_v = new_carry; S.have_carry = true; }
    if (__any_sync(FULL, n_st > SS_STAGE)) S.flags |= SK_FLAG_OVERFLOW;
    if (__any_sync(FULL, tie)) {  // redo this strip (and the w + SS_MAXPAL positions before it) with the exact automaton
      if ((cp >> 9) >= 64) S.flags |= SK_FLAG_TIE;
      else if (lane == 0) sh.bad |= 1ull << (cp >> 9);
    }
    S.flags = __reduce_or_sync(FULL, S.flags);
    const uint32_t cnt = n_st - skip;
    uint32_t inc = cnt;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const uint32_t o = __shfl_up_sync(FULL, inc, dlt);
      if (lane >= dlt) inc += o;
    }
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    if ((uint64_t)S.n_out + total > S.cap) S.flags |= SK_FLAG_OVERFLOW;
    if (!S.flags) {
      uint32_t at = S.n_out + inc - cnt;
      for (uint32_t i = skip; i < n_st; i++) {
        const HT v = sh.stv[lane * SS_STAGE + i];
        const int p = (int)sh.stp[lane * SS_STAGE + i];
        // strand of the canonical k-mer at p (mm_sketch.c:106)
        const uint64_t Vp_ = fetch_fwd64(S.w, S.base0 + p - S.k + 1);
        const HT kmer1 = (HT)(~Vp_ & S.mask64), kmer0 = (HT)((rev2(Vp_) >> (64 - 2 * S.k)) & S.mask64);
        mm128 m;
        m.x = (uint64_t)v << 8 | (uint64_t)S.k;
        m.y = S.ridhi | (uint64_t)((uint32_t)p << 1 | (kmer0 < kmer1 ? 0u : 1u));
        S.out[at++] = m;
      }
      S.n_out += total;
    }
  }
  __syncwarp();  // the stage and the ring rows are rewritten by the next strip
}

// One warp = one read.  Output: fast_cnt[row] records at tmp + tmp_off[row] (position order); row_flags[row] != 0 when the read
// (SK_FLAG_PARTIAL: only the strips in row_bad[row]) must be redone by the exact automaton; cnt_by_row[row] is then 0.
template <class HT>
__global__ void __launch_bounds__(SS_WARPS * 32, sizeof(HT) == 8 ? 4 : 6) k_sketch_strip(const uint64_t *__restrict__ w, const uint32_t *__restrict__ row_rid,
                                                                   const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                                                                   const uint32_t *__restrict__ hasn_by_rid, uint32_t row_first, uint32_t n_rows, int wsz, int k,
                                                                   const uint64_t *__restrict__ tmp_off, mm128 *__restrict__ tmp,
                                                                   uint32_t *__restrict__ cnt_by_row, uint32_t *__restrict__ row_flags,
                                                                   uint64_t *__restrict__ row_bad, uint32_t *__restrict__ fast_cnt) {
  extern __shared__ __align__(16) unsigned char ss_smem[];
  constexpr uint32_t FULL = 0xffffffffu;
  const HT MAXV = (HT) ~(HT)0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t row = row_first + blockIdx.x * SS_WARPS + wid;
  if (row >= row_first + n_rows) return;  // warp-uniform
  SsWarpSmem<HT> &sh = reinterpret_cast<SsWarpSmem<HT> *>(ss_smem)[wid];
  const uint32_t rid = row_rid[row];
  SsState<HT> S;
  S.w = w;
  S.len = (int)row_len[row];
  if (hasn_by_rid[rid] || S.len < sk_min_len(wsz, k)) {
    if (lane == 0) {
      row_flags[row] = hasn_by_rid[rid] ? (uint32_t)SK_FLAG_N : (uint32_t)SK_FLAG_SHORT;
      cnt_by_row[row] = 0; row_bad[row] = 0; fast_cnt[row] = 0;
    }
    return;
  }
  S.base0 = (int64_t)row_woff[row] * 32;
  S.wsz = wsz; S.k = k;
  S.mask64 = ((uint64_t)1 << 2 * k) - 1;
  S.mask = (HT)S.mask64;
  S.e_ff = wsz + k - 2;   // the first full window ends here (l == w+k-1, mm_sketch.c:116) when no palindrome precedes it
  S.s_eval = S.e_ff - 1;  // the window before it is evaluated for its tie check only (first-window special case)
  const uint64_t out0 = tmp_off[row];
  S.cap = tmp_off[row + 1] - out0;
  S.out = tmp + out0;
  S.ridhi = (uint64_t)rid << 32;
  S.n_out = 0; S.flags = 0;
  if (lane == 0) sh.bad = 0;
  S.carry_v = MAXV; S.have_carry = false;
  S.n_pal = 0; S.last_pal = -0x40000000;
  const int len = S.len;

  for (int cp = 0; cp < len && !S.flags; cp += SS_STRIP) {
    const int pos0 = cp + SS_SPL * lane;
    const bool interior = cp > S.e_ff + 1 && cp + SS_STRIP <= len;
    // ---------------- hashes of the lane's 16 positions
    HT h[SS_SPL];
    uint32_t palm = 0, vm = 0xFFFFu;  // vm: positions that are window slots with a complete k-mer (bit i)
    const uint64_t V = fetch_fwd64(w, S.base0 + pos0 - k + 1);  // bases pos0-k+1 .. pos0+32-k (guard words in front of the first read)
    const uint64_t R = rev2(V);
    uint64_t V2 = 0, R2 = 0;  // k > 16: the next 32 bases as well
    if (sizeof(HT) == 8) { V2 = fetch_fwd64(w, S.base0 + pos0 - k + 33); R2 = rev2(V2); }
#pragma unroll
    for (int i = 0; i < SS_SPL; i++) {
      HT kmer0, kmer1;
      if (sizeof(HT) == 8) {
        uint64_t a, b;
        ss_kmers128(V, V2, R, R2, i, k, S.mask64, &a, &b);
        kmer0 = (HT)a; kmer1 = (HT)b;
      } else {
        ss_kmers<HT>(V, R, i, k, S.mask, &kmer0, &kmer1);
      }
      if (kmer0 == kmer1) palm |= 1u << i;
      h[i] = sk_hash<HT>(kmer0 < kmer1 ? kmer0 : kmer1, S.mask);
    }
    if (!interior) {  // warp-uniform: the read's first and last strips
#pragma unroll
      for (int i = 0; i < SS_SPL; i++) {
        // an existing position with an incomplete k-mer is a sentinel SLOT (l < k, mm_sketch.c:108-111); one past the end is nothing:
        // both carry MAXV here, and no window that is evaluated reaches either kind (windows start at k-2 and end before len)
        if (!(pos0 + i < len && pos0 + i >= k - 1)) { vm &= ~(1u << i); palm &= ~(1u << i); }
      }
    }
    // ---------------- palindromic k-mers of the strip (warp-uniform bookkeeping)
    const uint32_t pal_lanes = __ballot_sync(FULL, palm != 0);
    if (pal_lanes) {
      for (uint32_t m = pal_lanes; m; m &= m - 1) {  // lanes in order, positions in order
        const int src = __ffs((int)m) - 1;
        uint32_t pm = __shfl_sync(FULL, palm, src);
        for (; pm; pm &= pm - 1) {
          const int q = cp + SS_SPL * src + (__ffs((int)pm) - 1);
          if (q <= S.e_ff + 1 && lane == 0) sh.bad |= 1ull;   // before the first full window (it moves): strip 0 is redone
          if (S.n_pal >= SS_MAXPAL) S.flags |= SK_FLAG_PAL;   // too many for the ring: whole read
          if (lane == 0) sh.pal[S.n_pal & (SS_MAXPAL - 1)] = q;
          S.n_pal++;
          S.last_pal = q;
        }
      }
      __syncwarp();
    }
    const bool slow = S.last_pal >= cp - wsz - 1;  // some window of this strip may contain a palindromic k-mer
    if (interior && !slow) {
      ss_strip<HT, true>(S, sh, cp, lane, h, vm, false);
    } else {
      vm &= ~palm;
#pragma unroll
      for (int i = 0; i < SS_SPL; i++)
        if (!((vm >> i) & 1u)) h[i] = MAXV;
      ss_strip<HT, false>(S, sh, cp, lane, h, vm, slow);
    }
  }
  if (lane == 0) {
    const unsigned long long bad = sh.bad;
    const bool partial = !S.flags && bad;
    row_flags[row] = partial ? (uint32_t)SK_FLAG_PARTIAL : S.flags;
    row_bad[row] = partial ? bad : 0ull;
    fast_cnt[row] = S.flags ? 0u : S.n_out;            // records in the read's slab (a partial read: incl. those at tainted positions)
    cnt_by_row[row] = (S.flags || bad) ? 0u : S.n_out;  // final count, known now only for clean reads
  }
}

template <class HT>
inline size_t ss_smem_bytes() { return sizeof(SsWarpSmem<HT>) * SS_WARPS; }

}  // namespace pgb
