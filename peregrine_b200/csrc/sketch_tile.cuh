// sketch_tile.cuh — tiled, data-parallel (w,k)-minimizer kernel (the fast path of mm_sketch, src/mm_sketch.c:70-151).
//
// One CTA of 256 threads sketches one TILE of one read out of shared memory:
//   phase 1  each thread hashes 17 consecutive positions (k-mer registers extracted once from the packed words, then
//            rolled) and notes palindromic k-mers (they occupy no window slot, mm_sketch.c:104-105);
//   phase 2  block scan of slot counts; slots (hash, pos<<1|strand) are written to shared memory in slot order;
//   phase 3  per 16-slot group: minimum (rightmost on ties, with a "tie seen" bit) and all suffix minima;
//   phase 4  for every window end s: rightmost arg-min of slots [s-w+1, s] = suffix(left group) + whole groups +
//            running prefix(own group)  (3 combines per window, no divergence);
//   phase 5  a slot is emitted when it becomes the window arg-min (first full window, or arg-min changed); block scan
//            of emit counts; records leave in position order.
// Exactness: on tie-free, N-free windows "rightmost arg-min of every full window" IS the reference's output (SURVEY
// App. A-5; re-checked against the reference by tests/hostsim, which runs these very functions on the CPU).  Every
// other case is detected and the whole read is handed to the exact automaton (k_sketch_exact): reads with N, reads
// shorter than one window (+ margin), any evaluated window whose minimum occurs twice (hash tie), regions with more
// than SK_PALPAD palindromic k-mers, tiles that overflow their record budget.
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

enum { SK_THREADS = 256, SK_G = 17, SK_R = SK_THREADS * SK_G /* 4352 region positions */, SK_GS = 16, SK_PALPAD = 16,
       SK_NG = (SK_R + SK_GS - 1) / SK_GS /* 272 slot groups */, SK_CAP = 512 /* records per tile */ };
enum { SK_FLAG_TIE = 1, SK_FLAG_PAL = 2, SK_FLAG_OVERFLOW = 4, SK_FLAG_SHORT = 8, SK_FLAG_N = 16 };

PGB_HD int sk_halo(int wsz) { return wsz + SK_PALPAD; }
PGB_HD int sk_tile_len(int wsz) { return SK_R - sk_halo(wsz); }
PGB_HD int sk_min_len(int wsz, int k) { return wsz + k + SK_PALPAD + 1; }  // shorter reads go to the exact automaton

struct SkParams {
  const uint64_t *w;  // packed reads
  uint64_t word_off;  // of this read
  int len;            // read length
  uint32_t rid;
  int wsz, k;
  int r0;             // region start position (negative for the first tile)
  int first_tile;
};

template <class HT>
struct SkShared {
  HT hv[SK_R];               // hash per slot (all ones = sentinel: k-mer not complete yet)
  HT sv[SK_R];               // suffix minimum value within the slot's 16-group
  uint16_t ps[SK_R];         // (region-relative position) << 1 | strand
  uint16_t sp[SK_R];         // suffix arg-min slot | tie << 15
  uint16_t amin[SK_R];       // window arg-min slot for the window ending at this slot
  HT gv[SK_NG];              // group minimum
  uint16_t gp[SK_NG];        // group arg-min slot | tie << 15
  uint32_t scan[2 * SK_THREADS + 2];
  uint32_t n_slots, n_pal, n_halo_slots, flags, n_emit;
};

template <class HT>
struct SkMin {
  HT v;
  uint32_t p;  // slot index | tie << 15
};
// b lies to the RIGHT of a: ties go to b (the reference keeps the newest of equal k-mers, mm_sketch.c:126,135-138)
template <class HT>
PGB_HD SkMin<HT> sk_combine(const SkMin<HT> &a, const SkMin<HT> &b) {
  SkMin<HT> r;
  if (b.v < a.v) r = b;
  else if (b.v == a.v) { r.v = b.v; r.p = b.p | 0x8000u; }
  else r = a;
  return r;
}

PGB_HD int sk_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// ---- phase 1: hash the thread's SK_G positions.  slot_mask bit i = position i is a window slot; halo_slots = slots whose
// region-relative position is below `halo`.
template <class HT>
PGB_HD void sk_phase1(int tid, const SkParams &p, int halo, HT *hv_out, uint16_t *ps_out, uint32_t *slot_mask, uint32_t *n_pal,
                      uint32_t *halo_slots) {
  const int k = p.k;
  const uint64_t mask = (1ULL << 2 * k) - 1, shift1 = 2 * (uint64_t)(k - 1);
  const int q0 = tid * SK_G;
  uint32_t sm = 0, np = 0, hs = 0;
  uint64_t kmer0 = 0, kmer1 = 0;
  bool have = false;
  const int64_t base0 = (int64_t)p.word_off * 32;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < SK_G; i++) {
    const int pos = p.r0 + q0 + i;
    hv_out[i] = (HT)~(HT)0;
    ps_out[i] = 0;
    if (pos < 0 || pos >= p.len) continue;  // does not exist
    if (pos < k - 1) {                     // exists, k-mer incomplete: sentinel slot (l < k)
      sm |= 1u << i;
      if (q0 + i < halo) hs++;
      continue;
    }
    if (!have) {
      uint64_t v = fetch_fwd64(p.w, base0 + pos - k + 1) & mask;  // bases pos-k+1 .. pos, earliest in the low bits
      kmer1 = (~v) & mask;
      kmer0 = rev2(v) >> (64 - 2 * k);
      have = true;
    } else {
      const int64_t a = base0 + pos;
      const uint64_t c = (p.w[a >> 5] >> (2 * (a & 31))) & 3;
      kmer0 = (kmer0 << 2 | c) & mask;
      kmer1 = (kmer1 >> 2) | (3ULL ^ c) << shift1;
    }
    if (kmer0 == kmer1) { np++; continue; }  // palindromic k-mer: no slot
    const int z = kmer0 < kmer1 ? 0 : 1;
    hv_out[i] = (HT)hash64(z ? kmer1 : kmer0, mask);
    ps_out[i] = (uint16_t)(((q0 + i) << 1) | z);
    sm |= 1u << i;
    if (q0 + i < halo) hs++;
  }
  *slot_mask = sm;
  *n_pal = np;
  *halo_slots = hs;
}

// ---- phase 2: write the thread's slots at their slot index
template <class HT>
PGB_HD void sk_phase2_write(SkShared<HT> &sh, const HT *hv, const uint16_t *ps, uint32_t slot_mask, uint32_t slot_base) {
  uint32_t s = slot_base;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int i = 0; i < SK_G; i++)
    if (slot_mask >> i & 1) {
      sh.hv[s] = hv[i];
      sh.ps[s] = ps[i];
      s++;
    }
}

// ---- phase 3: group minimum + suffix minima of group g
template <class HT>
PGB_HD void sk_phase3_group(int g, SkShared<HT> &sh) {
  const uint32_t ns = sh.n_slots;
  const uint32_t b = (uint32_t)g * SK_GS;
  if (b >= ns) return;
  const uint32_t e = b + SK_GS < ns ? b + SK_GS : ns;
  SkMin<HT> run;
  run.v = sh.hv[e - 1];
  run.p = e - 1;
  sh.sv[e - 1] = run.v;
  sh.sp[e - 1] = (uint16_t)run.p;
  for (uint32_t s = e - 1; s-- > b;) {
    SkMin<HT> c;
    c.v = sh.hv[s];
    c.p = s;
    run = sk_combine(c, run);  // c is to the LEFT of run
    sh.sv[s] = run.v;
    sh.sp[s] = (uint16_t)run.p;
  }
  sh.gv[g] = run.v;
  sh.gp[g] = (uint16_t)run.p;
}

// ---- phase 4: window arg-min for the window-end slots of group g (windows ending at s >= s_eval); returns tie bit.
// Requires wsz >= SK_GS + 1 so that the left window edge always lies in an earlier group.
template <class HT>
PGB_HD uint32_t sk_phase4_group(int g, SkShared<HT> &sh, int wsz, int s_eval) {
  const int ns = (int)sh.n_slots;
  const int b = g * SK_GS;
  if (b >= ns) return 0;
  const int e = b + SK_GS < ns ? b + SK_GS : ns;
  if (e - 1 < s_eval) return 0;
  int first = b > s_eval ? b : s_eval;          // first evaluated window end in this group
  const int ga = (first - wsz + 1) / SK_GS;      // group of the left edge for the first evaluated window (>= 0)
  // whole groups strictly between the left-edge group and this group: Mb = groups ga+2 .. g-1, Ma = ga+1 .. g-1
  SkMin<HT> Mb, Ma;
  bool hasMb = false, hasMa = false;
  Mb.v = 0; Mb.p = 0;
  for (int gi = ga + 2; gi < g; gi++) {
    SkMin<HT> c;
    c.v = sh.gv[gi];
    c.p = sh.gp[gi];
    if (!hasMb) { Mb = c; hasMb = true; } else Mb = sk_combine(Mb, c);
  }
  Ma = Mb;
  hasMa = hasMb;
  if (ga + 1 < g) {
    SkMin<HT> c;
    c.v = sh.gv[ga + 1];
    c.p = sh.gp[ga + 1];
    if (hasMb) Ma = sk_combine(c, Mb); else { Ma = c; hasMa = true; }
  }
  uint32_t tie = 0;
  SkMin<HT> pre;
  pre.v = 0; pre.p = 0;
  for (int s = b; s < e; s++) {
    SkMin<HT> c;
    c.v = sh.hv[s];
    c.p = (uint32_t)s;
    if (s == b) pre = c; else pre = sk_combine(pre, c);
    if (s < s_eval) continue;
    const int lo = s - wsz + 1;
    const int gl = lo / SK_GS;
    SkMin<HT> win;
    win.v = sh.sv[lo];
    win.p = sh.sp[lo];
    if (gl == ga) { if (hasMa) win = sk_combine(win, Ma); }
    else { if (hasMb) win = sk_combine(win, Mb); }
    win = sk_combine(win, pre);
    sh.amin[s] = (uint16_t)(win.p & 0x7FFF);
    tie |= win.p >> 15;
  }
  return tie;
}

// ---- phase 5: records emitted by the window ends of group g
template <class HT, bool WRITE>
PGB_HD uint32_t sk_phase5_group(int g, const SkShared<HT> &sh, const SkParams &p, int s_emit, int s_first_full, mm128 *out) {
  const int ns = (int)sh.n_slots;
  const int b = g * SK_GS;
  if (b >= ns) return 0;
  const int e = b + SK_GS < ns ? b + SK_GS : ns;
  uint32_t n = 0;
  for (int s = b > s_emit ? b : s_emit; s < e; s++) {
    const uint32_t a = sh.amin[s];
    const bool emit = (s == s_first_full) || (a != sh.amin[s - 1]);
    if (!emit) continue;
    if (WRITE) {
      const uint32_t pz = sh.ps[a];
      mm128 m;
      m.x = (uint64_t)sh.hv[a] << 8 | (uint64_t)p.k;
      m.y = (uint64_t)p.rid << 32 | (uint64_t)(uint32_t)(p.r0 + (int)(pz >> 1)) << 1 | (pz & 1);
      out[n] = m;
    }
    n++;
  }
  return n;
}

// window bookkeeping of a tile once n_slots / n_halo_slots are known
PGB_HD void sk_ranges(const SkParams &p, uint32_t n_halo_slots, int *s_eval, int *s_emit, int *s_first_full) {
  if (p.first_tile) {
    *s_first_full = p.wsz + p.k - 2;  // slot index at which l == w+k-1 (mm_sketch.c:116)
    *s_emit = *s_first_full;
    *s_eval = *s_first_full - 1;      // the prefix window is evaluated for its tie bit only (first-window special case)
  } else {
    *s_first_full = -1;
    *s_emit = (int)n_halo_slots;      // first slot at/after the tile's first own position
    *s_eval = *s_emit - 1;
  }
}

}  // namespace pgb
