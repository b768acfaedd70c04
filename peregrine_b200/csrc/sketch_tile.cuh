// sketch_tile.cuh — tiled, data-parallel (w,k)-minimizer kernel (the fast path of mm_sketch, src/mm_sketch.c:70-151).
//
// One CTA of 256 threads sketches one TILE (SK_R positions incl. a w+16 halo) of one read out of shared memory:
//   phase 1  each thread hashes SK_G consecutive positions: the k-mer pair is extracted once from the packed words and
//            then rolled out of one 64-bit register of bases; k-mer and hash arithmetic run in 32 bits when k <= 16 (the
//            reference's hash64 masks to 2k bits after every step, so it is arithmetic mod 2^2k); palindromic k-mers are
//            noted (they occupy no window slot, mm_sketch.c:104-105);
//   phase 2  block scan of slot counts; slots (hash, pos<<1|strand) are written to shared memory in slot order;
//   phase 3  van Herk / Gil-Werman with blocks of B = ceil(w/2)|1 slots: one thread per block runs the suffix minima
//            (right to left), then the prefix minima (left to right, in place).  Every minimum carries its arg-min slot
//            (rightmost on ties) and a "this minimum occurs twice" bit.  B is odd, so the per-thread sequential walks
//            (stride B between threads) are free of bank conflicts;
//   phase 4  each thread evaluates SK_G consecutive window ends: window [e-w+1, e] = suffix(left block) (+ one whole
//            block) + prefix(right block), i.e. at most two combines; a slot is emitted when it becomes the window
//            arg-min (first full window, or arg-min changed);
//   phase 5  block scan of emit counts; records leave in position order.
// Exactness: on tie-free, N-free windows "rightmost arg-min of every full window" IS the reference's output (SURVEY
// App. A-5; re-checked against the reference by tests/hostsim, which runs these very functions on the CPU).  Every
// other case is detected and the whole read is handed to the exact automaton (k_sketch_exact_seg): reads with N, reads
// shorter than one window (+ margin), any evaluated window whose minimum occurs twice (hash tie), regions with more
// than SK_PALPAD palindromic k-mers, tiles that overflow their record budget.
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

#ifndef PGB_SK_G
#define PGB_SK_G 13
#endif
enum { SK_THREADS = 256, SK_G = PGB_SK_G, SK_R = SK_THREADS * SK_G /* region positions */, SK_PALPAD = 16, SK_CAP = 512 /* records per tile */,
       SK_MINW = 17 /* smallest window the tiled kernel takes */ };
// reasons for redoing a whole read with the exact automaton; PARTIAL (strip kernel only): no such reason, but some strips of the read
// hold a window the fast path cannot decide (tie / palindrome): only those strips are redone (row_bad = their bit mask)
enum { SK_FLAG_TIE = 1, SK_FLAG_PAL = 2, SK_FLAG_OVERFLOW = 4, SK_FLAG_SHORT = 8, SK_FLAG_N = 16, SK_FLAG_PARTIAL = 32 };

PGB_HD int sk_halo(int wsz) { return wsz + SK_PALPAD; }
PGB_HD int sk_tile_len(int wsz) { return SK_R - sk_halo(wsz); }
PGB_HD int sk_min_len(int wsz, int k) { return wsz + k + SK_PALPAD + 1; }  // shorter reads go to the exact automaton
PGB_HD int sk_block_len(int wsz) { return ((wsz + 1) / 2) | 1; }  // odd; ceil(w/2) <= B <= w - 1 for w >= SK_MINW
PGB_HD int sk_padn() { return SK_R + 4; }
// exact s / B for 0 <= s < 2^13 and 9 <= B <= 129 (s * magic < 2^32; error term s / 2^22 < 1 / B)
PGB_HD uint32_t sk_div_magic(int B) { return (uint32_t)(((1u << 22) + (uint32_t)B - 1u) / (uint32_t)B); }
PGB_HD int sk_div(int s, uint32_t magic) { return (int)(((uint32_t)s * magic) >> 22); }

struct SkParams {
  const uint64_t *w;  // packed reads
  uint64_t word_off;  // of this read
  int len;            // read length
  uint32_t rid;
  int wsz, k;
  int r0;             // region start position (negative for the first tile)
  int first_tile;
};

// shared-memory image of a tile (pointers into one dynamic allocation; plain vectors in tests/hostsim)
template <class HT>
struct SkTile {
  HT *hv;        // [padn] hash per slot (all ones = sentinel: k-mer not complete yet); prefix minimum after phase 3
  HT *sv;        // [padn] suffix minimum within the slot's block
  uint16_t *pp;  // [padn] prefix arg-min slot | tie << 15
  uint16_t *sp;  // [padn] suffix arg-min slot | tie << 15
  uint16_t *ps;  // [SK_R] (region-relative position) << 1 | strand, by slot
  uint32_t *scan;  // [16]
  uint32_t *ctr;   // [4] n_slots, n_pal, n_halo_slots, flags
  int B;           // block length of the prefix / suffix scans
  uint32_t magic;  // sk_div_magic(B)
  int two_blocks;  // a window whose left edge has in-block offset >= two_blocks spans a whole block in between
};
enum { SK_N_SLOTS = 0, SK_N_PAL = 1, SK_N_HALO = 2, SK_FLAGS = 3 };
template <class HT>
PGB_HD size_t sk_smem_bytes() {
  return (size_t)sk_padn() * (2 * sizeof(HT) + 4) + (size_t)SK_R * 2 + 20 * 4 + 16;
}
template <class HT>
PGB_HD void sk_tile_layout(SkTile<HT> &t, unsigned char *base, int wsz) {
  const size_t n = (size_t)sk_padn();
  t.hv = reinterpret_cast<HT *>(base);
  t.sv = t.hv + n;
  t.scan = reinterpret_cast<uint32_t *>(t.sv + n);
  t.ctr = t.scan + 16;
  t.pp = reinterpret_cast<uint16_t *>(t.ctr + 4);
  t.sp = t.pp + n;
  t.ps = t.sp + n;
  t.B = sk_block_len(wsz);
  t.magic = sk_div_magic(t.B);
  t.two_blocks = 2 * t.B - wsz + 1;
}

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

// src/mm_sketch.c:23-32 in HT arithmetic: identical to the 64-bit original because every step is masked to 2k <= bits(HT)
template <class HT>
PGB_HD HT sk_hash(HT key, HT mask) {
  key = (HT)(~key + (HT)(key << 21)) & mask;
  key = key ^ (HT)(key >> 24);
  key = (HT)((HT)(key + (HT)(key << 3)) + (HT)(key << 8)) & mask;
  key = key ^ (HT)(key >> 14);
  key = (HT)((HT)(key + (HT)(key << 2)) + (HT)(key << 4)) & mask;
  key = key ^ (HT)(key >> 28);
  key = (HT)(key + (HT)(key << 31)) & mask;
  return key;
}

// ---- phase 1: hash the thread's SK_G positions.  slot_mask bit i = position i is a window slot; halo_slots = slots whose
// region-relative position is below `halo`.
template <class HT>
PGB_HD void sk_phase1(int tid, const SkParams &p, int halo, HT *hv_out, uint16_t *ps_out, uint32_t *slot_mask, uint32_t *n_pal,
                      uint32_t *halo_slots) {
  const int k = p.k;
  const uint64_t mask64 = ((uint64_t)1 << 2 * k) - 1;
  const HT mask = (HT)mask64;
  const int shift1 = 2 * (k - 1);
  const int q0 = tid * SK_G;
  const int pos0 = p.r0 + q0;
  const int64_t base0 = (int64_t)p.word_off * 32;
  uint32_t sm = 0, np = 0;
  HT kmer0 = 0, kmer1 = 0;
  if (pos0 >= k - 1 && pos0 + SK_G <= p.len) {
    // interior thread (all but the few at the ends of the read): every position has a complete k-mer
    const uint64_t v = fetch_fwd64(p.w, base0 + pos0 - k + 1) & mask64;  // bases pos-k+1 .. pos, earliest in the low bits
    kmer1 = (HT)((~v) & mask64);
    kmer0 = (HT)(rev2(v) >> (64 - 2 * k));
    const uint64_t bases = fetch_fwd64(p.w, base0 + pos0 + 1);  // the bases rolled in at i = 1 ..
    uint32_t pal = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < SK_G; i++) {
      if (i > 0) {
        const HT c = (HT)((bases >> (2 * (i - 1))) & 3);
        kmer0 = (HT)((HT)(kmer0 << 2) | c) & mask;
        kmer1 = (HT)(kmer1 >> 2) | (HT)((HT)(3 ^ c) << shift1);
      }
      const bool z = !(kmer0 < kmer1);
      pal |= (uint32_t)(kmer0 == kmer1) << i;  // palindromic k-mer: no slot
      hv_out[i] = sk_hash<HT>(z ? kmer1 : kmer0, mask);
      ps_out[i] = (uint16_t)(((q0 + i) << 1) | (int)z);
    }
    np = (uint32_t)sk_popc(pal);
    sm = ~pal & ((1u << SK_G) - 1u);
  } else {
    // first position of this thread whose k-mer is complete
    int i0 = k - 1 - pos0;
    if (i0 < 0) i0 = 0;
    uint64_t bases = 0;
    if (i0 < SK_G && pos0 + i0 < p.len) {
      const int pos = pos0 + i0;
      const uint64_t v = fetch_fwd64(p.w, base0 + pos - k + 1) & mask64;
      kmer1 = (HT)((~v) & mask64);
      kmer0 = (HT)(rev2(v) >> (64 - 2 * k));
      bases = fetch_fwd64(p.w, base0 + pos + 1);
    }
    for (int i = 0; i < SK_G; i++) {
      const int pos = pos0 + i;
      hv_out[i] = (HT)~(HT)0;
      ps_out[i] = 0;
      if (pos < 0 || pos >= p.len) continue;  // does not exist
      if (i < i0) {                          // exists, k-mer incomplete: sentinel slot (l < k)
        sm |= 1u << i;
        continue;
      }
      if (i > i0) {
        const HT c = (HT)((bases >> (2 * (i - i0 - 1))) & 3);
        kmer0 = (HT)((HT)(kmer0 << 2) | c) & mask;
        kmer1 = (HT)(kmer1 >> 2) | (HT)((HT)(3 ^ c) << shift1);
      }
      if (kmer0 == kmer1) { np++; continue; }
      const int z = kmer0 < kmer1 ? 0 : 1;
      hv_out[i] = sk_hash<HT>(z ? kmer1 : kmer0, mask);
      ps_out[i] = (uint16_t)(((q0 + i) << 1) | z);
      sm |= 1u << i;
    }
  }
  *slot_mask = sm;
  *n_pal = np;
  const int nh = halo - q0;  // positions of this thread inside the halo
  *halo_slots = nh <= 0 ? 0u : (uint32_t)sk_popc(nh >= SK_G ? sm : (sm & ((1u << nh) - 1u)));
}

// ---- phase 2: write the thread's slots at their slot index
template <class HT>
PGB_HD void sk_phase2_write(SkTile<HT> &sh, const HT *hv, const uint16_t *ps, uint32_t slot_mask, uint32_t slot_base) {
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

// ---- phase 3a: suffix minima of block b
template <class HT>
PGB_HD void sk_phase3_suffix(int b, SkTile<HT> &sh) {
  const int ns = (int)sh.ctr[SK_N_SLOTS], B = sh.B;
  const int first = b * B;
  if (first >= ns) return;
  const int last = (first + B < ns ? first + B : ns) - 1;
  SkMin<HT> run;
  run.v = sh.hv[last];
  run.p = (uint32_t)last;
  sh.sv[last] = run.v;
  sh.sp[last] = (uint16_t)run.p;
  for (int s = last - 1; s >= first; s--) {
    SkMin<HT> c;
    c.v = sh.hv[s];
    c.p = (uint32_t)s;
    run = sk_combine(c, run);  // c is to the LEFT of run
    sh.sv[s] = run.v;
    sh.sp[s] = (uint16_t)run.p;
  }
}
// ---- phase 3b: prefix minima of block b, in place over hv (run after ALL suffix scans)
template <class HT>
PGB_HD void sk_phase3_prefix(int b, SkTile<HT> &sh) {
  const int ns = (int)sh.ctr[SK_N_SLOTS], B = sh.B;
  const int first = b * B;
  if (first >= ns) return;
  const int last = (first + B < ns ? first + B : ns) - 1;
  SkMin<HT> run;
  run.v = sh.hv[first];
  run.p = (uint32_t)first;
  sh.pp[first] = (uint16_t)run.p;
  for (int s = first + 1; s <= last; s++) {
    SkMin<HT> c;
    c.v = sh.hv[s];
    c.p = (uint32_t)s;
    run = sk_combine(run, c);
    sh.hv[s] = run.v;
    sh.pp[s] = (uint16_t)run.p;
  }
}

// minimum of the window [lo, e], e - lo = w - 1, lo_off = offset of lo in its block: suffix of lo's block, the whole block
// in between if there is one (its total = prefix at its last slot), prefix of e's block
template <class HT>
PGB_HD SkMin<HT> sk_window(const SkTile<HT> &sh, int lo, int lo_off, int e) {
  SkMin<HT> win, c;
  win.v = sh.sv[lo];
  win.p = sh.sp[lo];
  if (lo_off >= sh.two_blocks) {
    const int mid = lo - lo_off + 2 * sh.B - 1;
    c.v = sh.hv[mid];
    c.p = sh.pp[mid];
    win = sk_combine(win, c);
  }
  c.v = sh.hv[e];
  c.p = sh.pp[e];
  return sk_combine(win, c);
}

// ---- phase 4: the thread's SK_G window ends.  Returns the number of records; bit i of *emit_mask = window end
// SK_G*tid + i emits; *tie = some evaluated window's minimum occurs twice.  Windows ending at e >= s_eval are evaluated
// (s_eval >= w - 1, so every evaluated window is full), windows ending at e >= s_emit may emit.
template <class HT>
PGB_HD uint32_t sk_phase4(int tid, const SkTile<HT> &sh, int wsz, int s_eval, int s_emit, int s_first_full, uint32_t *emit_mask, uint32_t *tie) {
  const int ns = (int)sh.ctr[SK_N_SLOTS];
  const int e0 = tid * SK_G, e1 = e0 + SK_G < ns ? e0 + SK_G : ns;
  *emit_mask = 0;
  *tie = 0;
  int e = e0 - 1 > s_eval ? e0 - 1 : s_eval;  // one window before the thread's own: its arg-min is the "previous" one
  if (e >= e1) return 0;
  int lo = e - wsz + 1;
  int lo_off = lo - sk_div(lo, sh.magic) * sh.B;
  uint32_t prev = 0xFFFFFFFFu, n = 0, em = 0, ti = 0;
  for (; e < e1; e++) {
    const SkMin<HT> win = sk_window(sh, lo, lo_off, e);
    const uint32_t a = win.p & 0x7FFFu;
    if (e >= e0) {
      ti |= win.p >> 15;
      if (e >= s_emit && (e == s_first_full || a != prev)) { em |= 1u << (e - e0); n++; }
    }
    prev = a;
    lo++;
    lo_off = lo_off + 1 == sh.B ? 0 : lo_off + 1;
  }
  *emit_mask = em;
  *tie = ti;
  return n;
}

// ---- phase 5: write the thread's records
template <class HT>
PGB_HD void sk_phase5_write(int tid, const SkTile<HT> &sh, const SkParams &p, uint32_t emit_mask, mm128 *out) {
  uint32_t n = 0;
  while (emit_mask) {
    const int i = ctz32(emit_mask);
    emit_mask &= emit_mask - 1;
    const int e = tid * SK_G + i, lo = e - p.wsz + 1;
    const SkMin<HT> win = sk_window(sh, lo, lo - sk_div(lo, sh.magic) * sh.B, e);
    const uint32_t pz = sh.ps[win.p & 0x7FFFu];
    mm128 m;
    m.x = (uint64_t)win.v << 8 | (uint64_t)p.k;
    m.y = (uint64_t)p.rid << 32 | (uint64_t)(uint32_t)(p.r0 + (int)(pz >> 1)) << 1 | (pz & 1);
    out[n++] = m;
  }
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
