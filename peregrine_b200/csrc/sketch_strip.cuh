// sketch_strip.cuh — (w,k)-minimizers of a read by ONE WARP that walks the read in strips of 512 positions
// (the fast path of mm_sketch, src/mm_sketch.c:70-151; replaces the block-tiled kernel of round 1).
//
// Why a strip per warp.  The round-1 kernel tiled reads over CTAs: 3 % of the positions were hashed twice (halo), the van
// Herk block scans kept a third of the CTA busy between barriers, and 152 thread-instructions were spent per base
// (profiles/r1g_ncu.md).  Here lane l of the warp owns positions [512 c + 16 l, 512 c + 16 l + 16) of strip c:
//   1. it extracts its first k-mer pair from the packed words, rolls 15 more bases through it and hashes the canonical
//      k-mer of every position (32-bit arithmetic when k <= 16: hash64 masks to 2k bits after every step);
//   2. suffix minima of its 16 hashes (rightmost on ties, "occurs twice" bit) go to a ring of rows in shared memory
//      (64 rows of 16 = the last two strips), one row per lane per strip;
//   3. a window [e-w+1, e] ending in the lane's segment is  suffix(row of e-w+1)  +  the whole rows in between  +  prefix of
//      the own segment: the minimum over the whole rows is computed ONCE per lane and strip (two variants, because the
//      16 window starts of a lane straddle at most two rows), the prefix is a running minimum in registers, so a window
//      costs one shared-memory load and two combines; no barrier, no halo: the ring carries the history into the next strip;
//   4. a position is emitted when it becomes the window's (rightmost) arg-min; emitted records are staged per lane and
//      leave in position order after one warp scan per strip.
// Exactness: on tie-free, N-free data "rightmost arg-min of every full window, reported when it changes" IS the reference
// output.  Everything else is detected and the whole read is handed to the exact automaton
// (k_sketch_exact_seg): reads with N, reads shorter than one window, any evaluated window whose minimum occurs twice,
// a palindromic k-mer before the first full window, two palindromic k-mers within one window, a lane with more than
// SS_STAGE records in a strip, a read that overflows its record budget.
// Palindromic k-mers (src/mm_sketch.c:104-105: they occupy NO window slot) are rare (4^-k/2 per position) but hit 20 % of
// 15 kb reads at k = 16: a strip within w positions of one takes a slower, general window evaluation in which the windows
// that contain the palindrome reach one position further back.
#pragma once
#include "shimmer_core.cuh"
#include "sketch_tile.cuh"

namespace pgb {

enum { SS_SPL = 16, SS_STRIP = 512, SS_ROWS = 64, SS_ROWPAD = 17, SS_WARPS = 4, SS_STAGE = 6, SS_MAXPAL = 8 };
#define SS_TIE 0x80000000u

template <class HT>
struct SsWarpSmem {
  HT sv[SS_ROWS * SS_ROWPAD];        // suffix minimum of the row from this offset on
  uint32_t sa[SS_ROWS * SS_ROWPAD];  // its arg: (position << 1 | strand) | SS_TIE
  HT stv[32 * SS_STAGE];             // staged records of the strip: value
  uint32_t sta[32 * SS_STAGE];       // ... arg
  int pal[SS_MAXPAL];                // positions of the most recent palindromic k-mers (ring)
};

template <class HT>
struct SsMin {
  HT v;
  uint32_t a;
};
// r lies to the RIGHT of l: ties go to r (the reference keeps the newest of equal k-mers, mm_sketch.c:126,135-138)
template <class HT>
__device__ __forceinline__ SsMin<HT> ss_combine(const SsMin<HT> &l, const SsMin<HT> &r) {
  SsMin<HT> o;
  const bool lt = r.v < l.v, eq = r.v == l.v;
  o.v = (lt || eq) ? r.v : l.v;
  o.a = lt ? r.a : (eq ? (r.a | SS_TIE) : l.a);
  return o;
}

// One warp = one read.  Output: cnt_by_row[row] records at tmp + tmp_off[row] (position order); row_flags[row] != 0 when the
// read must be redone by the exact automaton (its count is then 0).
template <class HT>
__global__ void __launch_bounds__(SS_WARPS * 32) k_sketch_strip(const uint64_t *__restrict__ w, const uint32_t *__restrict__ row_rid,
                                                                const uint32_t *__restrict__ row_len, const uint64_t *__restrict__ row_woff,
                                                                const uint32_t *__restrict__ hasn_by_rid, uint32_t row_first, uint32_t n_rows, int wsz, int k,
                                                                const uint64_t *__restrict__ tmp_off, mm128 *__restrict__ tmp,
                                                                uint32_t *__restrict__ cnt_by_row, uint32_t *__restrict__ row_flags) {
  extern __shared__ __align__(16) unsigned char ss_smem[];
  constexpr uint32_t FULL = 0xffffffffu;
  const HT MAXV = (HT) ~(HT)0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t row = row_first + blockIdx.x * SS_WARPS + wid;
  if (row >= row_first + n_rows) return;  // warp-uniform
  SsWarpSmem<HT> &sh = reinterpret_cast<SsWarpSmem<HT> *>(ss_smem)[wid];
  const uint32_t rid = row_rid[row];
  const int len = (int)row_len[row];
  if (hasn_by_rid[rid] || len < sk_min_len(wsz, k)) {
    if (lane == 0) { row_flags[row] = hasn_by_rid[rid] ? (uint32_t)SK_FLAG_N : (uint32_t)SK_FLAG_SHORT; cnt_by_row[row] = 0; }
    return;
  }
  const int64_t base0 = (int64_t)row_woff[row] * 32;
  const uint64_t mask64 = ((uint64_t)1 << 2 * k) - 1;
  const HT mask = (HT)mask64;
  const int shift1 = 2 * (k - 1);
  const int e_ff = wsz + k - 2;  // the first full window ends here (l == w+k-1, mm_sketch.c:116) when no palindrome precedes it
  const int s_eval = e_ff - 1;   // the window before it is evaluated for its tie bit only (first-window special case)
  const uint64_t out0 = tmp_off[row], cap = tmp_off[row + 1] - out0;
  mm128 *out = tmp + out0;
  const uint64_t ridhi = (uint64_t)rid << 32;
  uint32_t n_out = 0, flags = 0;
  uint32_t carry_a = 0xFFFFFFFFu;  // arg-min of the last window of the previous strip
  int n_pal = 0, last_pal = -0x40000000;

  for (int cp = 0; cp < len && !flags; cp += SS_STRIP) {
    const int pos0 = cp + SS_SPL * lane;
    // ---------------- 1. hashes of the lane's 16 positions
    HT h[SS_SPL];
    uint32_t zm = 0, palm = 0;
    {
      HT kmer0 = 0, kmer1 = 0;
      if (pos0 >= k - 1 && pos0 + SS_SPL <= len) {  // every position has a complete k-mer (all lanes but a few at the read's ends)
        const uint64_t v = fetch_fwd64(w, base0 + pos0 - k + 1) & mask64;  // bases pos0-k+1 .. pos0, earliest in the low bits
        kmer1 = (HT)((~v) & mask64);
        kmer0 = (HT)(rev2(v) >> (64 - 2 * k));
        const uint64_t bases = fetch_fwd64(w, base0 + pos0 + 1);
#pragma unroll
        for (int i = 0; i < SS_SPL; i++) {
          if (i > 0) {
            const HT c = (HT)((bases >> (2 * (i - 1))) & 3);
            kmer0 = (HT)((HT)(kmer0 << 2) | c) & mask;
            kmer1 = (HT)(kmer1 >> 2) | (HT)((HT)(3 ^ c) << shift1);
          }
          const bool z = !(kmer0 < kmer1);
          const bool p = kmer0 == kmer1;
          palm |= (uint32_t)p << i;
          zm |= (uint32_t)z << i;
          const HT hv = sk_hash<HT>(z ? kmer1 : kmer0, mask);
          h[i] = p ? MAXV : hv;
        }
      } else {
        int i0 = k - 1 - pos0;  // first position of this lane whose k-mer is complete
        if (i0 < 0) i0 = 0;
        uint64_t bases = 0;
        if (i0 < SS_SPL && pos0 + i0 < len) {
          const int pos = pos0 + i0;
          const uint64_t v = fetch_fwd64(w, base0 + pos - k + 1) & mask64;
          kmer1 = (HT)((~v) & mask64);
          kmer0 = (HT)(rev2(v) >> (64 - 2 * k));
          bases = fetch_fwd64(w, base0 + pos + 1);
        }
#pragma unroll
        for (int i = 0; i < SS_SPL; i++) {
          h[i] = MAXV;  // past the read's end, or k-mer incomplete (sentinel slot, l < k)
          if (pos0 + i < len && i >= i0) {
            if (i > i0) {
              const HT c = (HT)((bases >> (2 * (i - i0 - 1))) & 3);
              kmer0 = (HT)((HT)(kmer0 << 2) | c) & mask;
              kmer1 = (HT)(kmer1 >> 2) | (HT)((HT)(3 ^ c) << shift1);
            }
            const bool z = !(kmer0 < kmer1);
            if (kmer0 == kmer1) palm |= 1u << i;
            else { zm |= (uint32_t)z << i; h[i] = sk_hash<HT>(z ? kmer1 : kmer0, mask); }
          }
        }
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
          if (q <= e_ff + 1 || n_pal >= SS_MAXPAL) flags |= SK_FLAG_PAL;  // before the first full window / too many: exact automaton
          if (lane == 0) sh.pal[n_pal & (SS_MAXPAL - 1)] = q;
          n_pal++;
          last_pal = q;
        }
      }
      __syncwarp();
    }
    const bool slow = last_pal >= cp - wsz - 1;  // some window of this strip may contain a palindromic k-mer
    // ---------------- 2. suffix minima of the lane's segment -> its row of the ring
    {
      const int rbase = ((pos0 >> 4) & (SS_ROWS - 1)) * SS_ROWPAD;
      SsMin<HT> run;
      run.v = h[SS_SPL - 1];
      run.a = (uint32_t)(pos0 + SS_SPL - 1) << 1 | ((zm >> (SS_SPL - 1)) & 1u);
      sh.sv[rbase + SS_SPL - 1] = run.v;
      sh.sa[rbase + SS_SPL - 1] = run.a;
#pragma unroll
      for (int j = SS_SPL - 2; j >= 0; j--) {
        SsMin<HT> c;
        c.v = h[j];
        c.a = (uint32_t)(pos0 + j) << 1 | ((zm >> j) & 1u);
        run = ss_combine(c, run);
        sh.sv[rbase + j] = run.v;
        sh.sa[rbase + j] = run.a;
      }
    }
    __syncwarp();
    // ---------------- 3. the windows that end in the lane's segment
    uint32_t n_st = 0;               // staged records of this lane
    uint32_t first_a = 0xFFFFFFFFu;  // arg-min of the lane's first evaluated window (its emission is decided after the shuffle)
    uint32_t last_a = 0xFFFFFFFFu;   // ... of its last evaluated window
    bool first_cond = false;         // stage entry 0 is that first window's record, valid only if it differs from the left neighbour's
    uint32_t tie = 0;
    if (pos0 < len && pos0 + SS_SPL - 1 >= s_eval) {
      const int lo0 = pos0 - wsz + 1, r_lo0 = lo0 >> 4, r_e = pos0 >> 4;
      SsMin<HT> m_short, m_long;
      m_short.v = MAXV; m_short.a = 0;
      auto row_total = [&](int r) -> SsMin<HT> {
        SsMin<HT> t;
        const int ix = (r & (SS_ROWS - 1)) * SS_ROWPAD;
        t.v = sh.sv[ix];
        t.a = sh.sa[ix];
        return t;
      };
      if (!slow) {
        for (int r = (r_lo0 + 2 > 0 ? r_lo0 + 2 : 0); r < r_e; r++) m_short = ss_combine(m_short, row_total(r));
        m_long = m_short;
        if (r_lo0 + 1 >= 0 && r_lo0 + 1 < r_e) m_long = ss_combine(row_total(r_lo0 + 1), m_short);
      }
      SsMin<HT> pre;
      pre.v = MAXV; pre.a = 0;
      uint32_t prev = 0xFFFFFFFFu;
      bool have_prev = false;
#pragma unroll
      for (int j = 0; j < SS_SPL; j++) {
        const int e = pos0 + j;
        SsMin<HT> c;
        c.v = h[j];
        c.a = (uint32_t)e << 1 | ((zm >> j) & 1u);
        pre = (j == 0) ? c : ss_combine(pre, c);
        if (e < len && e >= s_eval && !((palm >> j) & 1u)) {
          SsMin<HT> win;
          if (!slow) {
            const int lo = e - wsz + 1;
            const int ix = ((lo >> 4) & (SS_ROWS - 1)) * SS_ROWPAD + (lo & 15);
            SsMin<HT> s;
            s.v = sh.sv[ix];
            s.a = sh.sa[ix];
            win = ss_combine(ss_combine(s, (lo >> 4) == r_lo0 ? m_long : m_short), pre);
          } else {
            // general form: the window holds w SLOTS; a palindromic k-mer inside it is no slot, so the window reaches one
            // position further back (two of them: exact automaton)
            int cpal = 0;
            bool edge_pal = false;
            const int np = n_pal < SS_MAXPAL ? n_pal : SS_MAXPAL;
            for (int t = 0; t < np; t++) {
              const int q = sh.pal[t];
              cpal += (q >= e - wsz + 1 && q <= e);
              edge_pal |= (q == e - wsz);
            }
            if (cpal > 1 || (cpal == 1 && edge_pal)) flags |= SK_FLAG_PAL;
            const int lo = e - wsz + 1 - (cpal ? 1 : 0);
            const int r_lo = lo >> 4;
            const int ix = (r_lo & (SS_ROWS - 1)) * SS_ROWPAD + (lo & 15);
            win.v = sh.sv[ix];
            win.a = sh.sa[ix];
            for (int r = r_lo + 1; r < r_e; r++) win = ss_combine(win, row_total(r));
            win = ss_combine(win, pre);
          }
          const uint32_t a = win.a & ~SS_TIE;
          tie |= win.a >> 31;
          if (e >= e_ff) {
            const bool is_first = !have_prev;
            if (e == e_ff || is_first || a != prev) {
              if (n_st < SS_STAGE) {
                sh.stv[lane * SS_STAGE + n_st] = win.v;
                sh.sta[lane * SS_STAGE + n_st] = a;
              }
              if (is_first && e != e_ff) first_cond = true;
              n_st++;
            }
          }
          if (!have_prev) { first_a = a; have_prev = true; }
          prev = a;
          last_a = a;
        }
      }
    }
    // ---------------- 4. resolve the lanes' first windows against their left neighbours, then write in position order
    {
      const uint32_t have = __ballot_sync(FULL, last_a != 0xFFFFFFFFu);
      const uint32_t below = have & ((1u << lane) - 1u);
      const int src = below ? 31 - __clz((int)below) : 0;
      uint32_t left = __shfl_sync(FULL, last_a, src);
      if (!below) left = carry_a;
      uint32_t skip = 0;
      if (first_cond && first_a == left) { skip = 1; }  // same arg-min as the window before it: not a new minimizer
      const uint32_t top = have ? 31 - __clz((int)have) : 0;
      const uint32_t new_carry = __shfl_sync(FULL, last_a, top);
      if (have) carry_a = new_carry;
      if (__any_sync(FULL, n_st > SS_STAGE)) flags |= SK_FLAG_OVERFLOW;
      if (__any_sync(FULL, tie != 0)) flags |= SK_FLAG_TIE;
      flags = __reduce_or_sync(FULL, flags);
      const uint32_t cnt = n_st - skip;
      uint32_t inc = cnt;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, dlt);
        if (lane >= dlt) inc += o;
      }
      const uint32_t total = __shfl_sync(FULL, inc, 31);
      if ((uint64_t)n_out + total > cap) flags |= SK_FLAG_OVERFLOW;
      if (!flags) {
        uint32_t at = n_out + inc - cnt;
        for (uint32_t i = skip; i < n_st; i++) {
          const HT v = sh.stv[lane * SS_STAGE + i];
          const uint32_t a = sh.sta[lane * SS_STAGE + i];
          mm128 m;
          m.x = (uint64_t)v << 8 | (uint64_t)k;
          m.y = ridhi | (uint64_t)a;
          out[at++] = m;
        }
        n_out += total;
      }
    }
    __syncwarp();  // the stage and the ring rows are rewritten by the next strip
  }
  if (lane == 0) {
    row_flags[row] = flags;
    cnt_by_row[row] = flags ? 0u : n_out;
  }
}

template <class HT>
inline size_t ss_smem_bytes() { return sizeof(SsWarpSmem<HT>) * SS_WARPS; }

}  // namespace pgb
