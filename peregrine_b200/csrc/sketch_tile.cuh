// sketch_tile.cuh — tiled, data-parallel (w,k)-minimizer kernel (the fast path of mm_sketch, src/mm_sketch.c:70-151).
//
// One CTA of 256 threads sketches one TILE of one read from shared memory:
//   phase 1  each thread hashes 17 consecutive positions (k-mer registers extracted once from the packed words,
//            then rolled), notes palindromic k-mers (they do not occupy a window slot, mm_sketch.c:104-105);
//   phase 2  block scan of slot counts -> slots (hash, pos<<1|strand) written to shared memory in slot order;
//   phase 3  per 16-slot group minimum (rightmost on ties, with a "tie seen" bit);
//   phase 4  for every window end s the rightmost arg-min of slots [s-w+1, s] = suffix(left group) + whole groups
//            + prefix(own group);
//   phase 5  a slot is emitted when it becomes the window arg-min (first full window, or arg-min changed), block
//            scan of emit counts, records written in position order.
// Exactness: on tie-free, N-free windows "rightmost arg-min of every full window (+ final min)" IS the reference's
// output (proven there on 2,075 cases and re-checked by tests/hostsim).  Anything else is detected
// and the whole read is handed to the exact automaton (k_sketch_exact): reads with N, reads shorter than one window,
// any window whose minimum occurs twice (hash tie), regions with more than SK_PALPAD palindromic k-mers, tiles that
// overflow their record budget.
//
// The phase functions are __host__ __device__ so tests/hostsim can run the identical code for all 256 "threads"
// sequentially (block scans are the only part with separate host/device bodies).
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

enum { SK_THREADS = 256, SK_G = 17, SK_R = SK_THREADS * SK_G /* 4352 region positions */, SK_GS = 16, SK_PALPAD = 16,
       SK_NG = (SK_R + SK_GS - 1) / SK_GS /* 272 slot groups */, SK_CAP = 512 /* records per tile */ };

struct SkParams {
  const uint64_t *w;     // packed reads
  uint64_t word_off;     // of this read
  int len;               // read length
  uint32_t rid;
  int wsz, k;
  int r0;                // region start position (may be negative for the first tile)
  int t0;                // first window-end position this tile is responsible for ( = r0 + halo )
  int first_tile;        // 1 if the region starts at the read start
};

struct SkShared {
  uint64_t hv[SK_R];             // hash per slot (0xFFFF.. = sentinel: k-mer not complete yet)
  uint16_t ps[SK_R];             // (region-relative position) << 1 | strand
  uint16_t amin[SK_R];           // slot index of the window arg-min for the window ending at this slot
  uint64_t gv[SK_NG];            // group minimum value
  uint16_t gp[SK_NG];            // group arg-min slot (rightmost)
  uint8_t gt[SK_NG];             // group tie bit
  uint32_t cnt[SK_THREADS + 1];  // scan scratch
  uint32_t n_slots, n_pal, tie, n_emit;
};

struct SkMin {
  uint64_t v;
  uint32_t p;    // slot index
  uint32_t tie;
};
// b lies to the RIGHT of a: ties go to b (the reference keeps the newest of equal k-mers, mm_sketch.c:126,135-138)
PGB_HD SkMin sk_combine(const SkMin &a, const SkMin &b) {
  SkMin r;
  if (b.v < a.v) { r = b; }
  else if (b.v == a.v) { r = b; r.tie = 1; }
  else { r = a; }
  return r;
}

// ---- phase 1+2a: hash the thread's positions; returns the number of slots and palindromes among them.
// hv_out/ps_out: per-thread arrays of SK_G entries (registers / local).  slot_flag bit i set = position i is a slot.
PGB_HD void sk_phase1(int tid, const SkParams &p, uint64_t *hv_out, uint16_t *ps_out, uint32_t *slot_mask, uint32_t *n_pal) {
  const int k = p.k;
  const uint64_t mask = (1ULL << 2 * k) - 1, shift1 = 2 * (uint64_t)(k - 1);
  const int q0 = tid * SK_G;           // region-relative index of the first position
  uint32_t sm = 0, np = 0;
  uint64_t kmer0 = 0, kmer1 = 0;
  bool have = false;
  const int64_t base0 = (int64_t)p.word_off * 32;
  for (int i = 0; i < SK_G; i++) {
    const int pos = p.r0 + q0 + i;
    hv_out[i] = ~0ULL;
    ps_out[i] = 0;
    if (pos < 0 || pos >= p.len) continue;  // does not exist
    if (pos < k - 1) {                     // exists, k-mer incomplete: sentinel slot (l < k)
      sm |= 1u << i;
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
    hv_out[i] = hash64(z ? kmer1 : kmer0, mask);
    ps_out[i] = (uint16_t)(((q0 + i) << 1) | z);
    sm |= 1u << i;
  }
  *slot_mask = sm;
  *n_pal = np;
}

PGB_HD int sk_popc(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// ---- phase 2b: write the thread's slots at their slot index
PGB_HD void sk_phase2_write(int tid, SkShared &sh, const uint64_t *hv, const uint16_t *ps, uint32_t slot_mask, uint32_t slot_base) {
  uint32_t s = slot_base;
  for (int i = 0; i < SK_G; i++)
    if (slot_mask >> i & 1) {
      sh.hv[s] = hv[i];
      sh.ps[s] = ps[i];
      s++;
    }
}

// ---- phase 3: group minima
PGB_HD void sk_phase3(int tid, SkShared &sh) {
  const uint32_t ns = sh.n_slots;
  for (uint32_t g = (uint32_t)tid; g < (uint32_t)SK_NG; g += SK_THREADS) {
    const uint32_t b = g * SK_GS;
    if (b >= ns) break;
    SkMin m;
    m.v = sh.hv[b]; m.p = b; m.tie = 0;
    const uint32_t e = b + SK_GS < ns ? b + SK_GS : ns;
    for (uint32_t s = b + 1; s < e; s++) {
      SkMin c;
      c.v = sh.hv[s]; c.p = s; c.tie = 0;
      m = sk_combine(m, c);
    }
    sh.gv[g] = m.v; sh.gp[g] = (uint16_t)m.p; sh.gt[g] = (uint8_t)m.tie;
  }
}

// first slot index whose window this tile must evaluate, and the first it must emit for
struct SkRange { int s_eval, s_emit, s_first_full; };

// ---- phase 4: window arg-min for the window-end slots of group(s) tid (+256); returns the tie flag seen in evaluated windows
PGB_HD uint32_t sk_phase4(int tid, SkShared &sh, int wsz, int s_eval) {
  const int ns = (int)sh.n_slots;
  uint32_t tie = 0;
  for (int g = tid; g < SK_NG; g += SK_THREADS) {
    const int b = g * SK_GS;
    if (b >= ns) break;
    const int e = b + SK_GS < ns ? b + SK_GS : ns;
    if (e - 1 < s_eval) continue;             // nothing to evaluate in this group
    // suffix minima of the two groups the left window edge sweeps through
    const int lo0 = b - wsz + 1;              // left edge for j = 0 (may be negative: clipped, only for windows we skip)
    int ga = lo0 >= 0 ? lo0 / SK_GS : 0;      // group of the left edge for the first window
    SkMin sufA[SK_GS], sufB[SK_GS];
    {
      SkMin run;
      for (int gi = 0; gi < 2; gi++) {
        const int gb = (ga + gi) * SK_GS;
        SkMin *suf = gi ? sufB : sufA;
        bool has = false;
        for (int o = SK_GS - 1; o >= 0; o--) {
          const int s = gb + o;
          if (s >= ns || s > e - 1) { suf[o].v = ~0ULL; suf[o].p = 0; suf[o].tie = 0; continue; }
          SkMin c;
          c.v = sh.hv[s]; c.p = (uint32_t)s; c.tie = 0;
          if (!has) { run = c; has = true; }
          else run = sk_combine(c, run);     // c is to the LEFT of run
          suf[o] = run;
        }
      }
    }
    // whole groups strictly between the left-edge group and this group: M2 = groups ga+2 .. g-1, M1 = ga+1 .. g-1
    SkMin M2; M2.v = ~0ULL; M2.p = 0; M2.tie = 0;
    bool hasM2 = false;
    for (int gi = ga + 2; gi < g; gi++) {
      SkMin c;
      c.v = sh.gv[gi]; c.p = sh.gp[gi]; c.tie = sh.gt[gi];
      if (!hasM2) { M2 = c; hasM2 = true; } else M2 = sk_combine(M2, c);
    }
    SkMin M1 = M2;
    bool hasM1 = hasM2;
    if (ga + 1 < g) {
      SkMin c;
      c.v = sh.gv[ga + 1]; c.p = sh.gp[ga + 1]; c.tie = sh.gt[ga + 1];
      if (hasM2) M1 = sk_combine(c, M2); else { M1 = c; hasM1 = true; }
    }
    // sweep the window ends of this group
    SkMin pre;
    for (int j = 0; b + j < e; j++) {
      const int s = b + j;
      SkMin c;
      c.v = sh.hv[s]; c.p = (uint32_t)s; c.tie = 0;
      if (j == 0) pre = c; else pre = sk_combine(pre, c);
      if (s < s_eval) continue;
      const int lo = s - wsz + 1;             // >= 0 for every evaluated window
      const int gl = lo / SK_GS, ol = lo % SK_GS;
      SkMin win;
      if (gl == g) {                          // window inside this group (w <= 16)
        win.v = ~0ULL; win.p = 0; win.tie = 0;
        bool has = false;
        for (int t = lo; t <= s; t++) {
          SkMin d; d.v = sh.hv[t]; d.p = (uint32_t)t; d.tie = 0;
          if (!has) { win = d; has = true; } else win = sk_combine(win, d);
        }
      } else {
        // left partial: suffix of group gl from ol; then whole groups gl+1 .. g-1; then prefix of this group
        const SkMin &L = (gl == ga) ? sufA[ol] : sufB[ol];
        win = L;
        if (gl == ga) { if (hasM1) win = sk_combine(win, M1); }
        else { if (hasM2) win = sk_combine(win, M2); }
        if (gl + 1 == g && gl != ga) { /* no whole group in between: nothing */ }
        win = sk_combine(win, pre);
      }
      sh.amin[s] = (uint16_t)win.p;
      tie |= win.tie;
    }
  }
  return tie;
}

// ---- phase 5a: count the records the thread's window ends emit; 5b: write them
template <bool WRITE>
PGB_HD uint32_t sk_phase5(int tid, const SkShared &sh, const SkParams &p, int s_emit, int s_first_full, mm128 *out, uint32_t out_base) {
  const int ns = (int)sh.n_slots;
  uint32_t n = 0;
  for (int g = tid; g < SK_NG; g += SK_THREADS) {   // NOTE: emission order across the two strides is fixed up by the caller's scan
    const int b = g * SK_GS;
    if (b >= ns) break;
    const int e = b + SK_GS < ns ? b + SK_GS : ns;
    for (int s = b; s < e; s++) {
      if (s < s_emit) continue;
      const uint32_t a = sh.amin[s];
      const bool emit = (s == s_first_full) || (a != sh.amin[s - 1]);
      if (!emit) continue;
      if (WRITE) {
        const uint32_t pz = sh.ps[a];
        mm128 m;
        m.x = sh.hv[a] << 8 | (uint64_t)p.k;
        m.y = (uint64_t)p.rid << 32 | (uint64_t)(uint32_t)(p.r0 + (int)(pz >> 1)) << 1 | (pz & 1);
        out[out_base + n] = m;
      }
      n++;
    }
  }
  return n;
}

}  // namespace pgb
