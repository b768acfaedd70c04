// host_util.hpp — host-side (CPU) plumbing of libpgb200: file formats, the read table, and the keys-only
// emulation of klib's khash that reproduces the reference's bucket *visiting order* (SURVEY App. A-3).
// Nothing here computes sketches, counts, pairs or alignments; those are CUDA kernels.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <errno.h>
#include <glob.h>
#include <string>
#include <vector>
#include <algorithm>
#include "shimmer_core.cuh"

namespace pgb {

[[noreturn]] inline void die(const char *fmt, const char *a = "", const char *b = "") {
  fprintf(stderr, "pgb200: ");
  fprintf(stderr, fmt, a, b);
  fprintf(stderr, "\n");
  exit(1);
}

// ------------------------------------------------------------------------------------------------ .idx
// Text lines "%u %255s %u %lu" = rid name len offset (src/shmr_mkseqdb.c:112, parsed at src/shmr_utils.c:259 and
// src/shmr_index.c:155).  Rows are kept in file order; by_rid maps rid -> row (last row wins, as kh_put overwrites).
struct ReadTable {
  std::vector<uint32_t> rid, len;
  std::vector<uint64_t> off;
  std::vector<int64_t> by_rid;  // -1 when absent
  uint64_t total_bases = 0;
  size_t n() const { return rid.size(); }
};

inline bool load_read_table(const char *path, ReadTable *t) {
  FILE *f = fopen(path, "r");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)sz + 1);
  size_t got = fread(buf.data(), 1, (size_t)sz, f);
  fclose(f);
  buf[got] = 0;
  const char *p = buf.data(), *e = p + got;
  auto skipws = [&]() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; };
  auto num = [&](uint64_t *v) -> bool {
    skipws();
    if (p >= e || *p < '0' || *p > '9') return false;
    uint64_t x = 0;
    while (p < e && *p >= '0' && *p <= '9') x = x * 10 + (uint64_t)(*p++ - '0');
    *v = x;
    return true;
  };
  uint64_t max_rid = 0;
  for (;;) {
    uint64_t r, l, o;
    if (!num(&r)) break;
    skipws();
    while (p < e && !(*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;  // name token
    if (!num(&l) || !num(&o)) break;
    t->rid.push_back((uint32_t)r);
    t->len.push_back((uint32_t)l);
    t->off.push_back(o);
    t->total_bases += l;
    if (r > max_rid) max_rid = r;
  }
  if (t->n() && max_rid > 8 * (uint64_t)t->n() + (1u << 20)) {
    fprintf(stderr, "pgb200: read ids in %s are too sparse (max rid %lu for %zu reads)\n", path, (unsigned long)max_rid,
            t->n());
    exit(1);
  }
  t->by_rid.assign(t->n() ? (size_t)max_rid + 1 : 0, -1);
  for (size_t i = 0; i < t->n(); i++) t->by_rid[t->rid[i]] = (int64_t)i;
  return true;
}

// ------------------------------------------------------------------------------------------------ .dat files
// mm128 list: size_t n + n x {u64 x, u64 y}   (src/shmr_utils.c:98-123)
inline void write_mmlist_file(const char *fn, const mm128 *a, size_t n) {
  FILE *f = fopen(fn, "wb");
  if (!f) {
    fprintf(stderr, "file '%s' open error: %s\n", fn, strerror(errno));
    exit(1);
  }
  fwrite(&n, sizeof(size_t), 1, f);
  if (n) fwrite(a, sizeof(mm128), n, f);
  fclose(f);
}
inline void read_mmlist_file(const char *fn, std::vector<mm128> *out) {  // appends
  FILE *f = fopen(fn, "rb");
  if (!f) {
    fprintf(stderr, "file '%s' open error: %s\n", fn, strerror(errno));
    exit(1);
  }
  size_t n = 0;
  if (fread(&n, sizeof(size_t), 1, f) != 1) n = 0;
  size_t base = out->size();
  out->resize(base + n);
  if (n && fread(out->data() + base, sizeof(mm128), n, f) != n) die("short read on '%s'", fn);
  fclose(f);
}
// count table: size_t n + n x {u64 mer; u32 count; 4 pad}  (src/shmr_utils.c:178-203, src/shimmer.h:61-64)
struct mc_rec {
  uint64_t mer;
  uint32_t count;
  uint32_t pad;
};
inline void write_mc_file(const char *fn, const mc_rec *a, size_t n) {
  FILE *f = fopen(fn, "wb");
  if (!f) {
    fprintf(stderr, "file '%s' open error: %s\n", fn, strerror(errno));
    exit(1);
  }
  fwrite(&n, sizeof(size_t), 1, f);
  if (n) fwrite(a, sizeof(mc_rec), n, f);
  fclose(f);
}
inline void read_mc_file(const char *fn, std::vector<mc_rec> *out) {  // appends
  FILE *f = fopen(fn, "rb");
  if (!f) {
    fprintf(stderr, "file '%s' open error: %s\n", fn, strerror(errno));
    exit(1);
  }
  size_t n = 0;
  if (fread(&n, sizeof(size_t), 1, f) != 1) n = 0;
  size_t base = out->size();
  out->resize(base + n);
  if (n && fread(out->data() + base, sizeof(mc_rec), n, f) != n) die("short read on '%s'", fn);
  fclose(f);
}
// "<prefix>-[0-9]*-of-[0-9]*.dat" in the order wordexp() returns it (sorted), src/shmr_overlap.c:359-382
inline std::vector<std::string> glob_sorted(const std::string &pattern) {
  std::vector<std::string> r;
  glob_t g;
  memset(&g, 0, sizeof g);
  // GLOB_NOCHECK: with no match the pattern itself comes back, as wordexp does in the reference (src/shmr_overlap.c:362): opening
  // it then fails with the reference's message and exit 1 instead of silently producing an empty result
  if (glob(pattern.c_str(), GLOB_NOCHECK, NULL, &g) == 0)
    for (size_t i = 0; i < g.gl_pathc; i++) r.push_back(g.gl_pathv[i]);
  globfree(&g);
  return r;
}

// ------------------------------------------------------------------------------------------------ khash order emulation
// Keys-only model of klib khash (src/khash.h:218-343) for KHASH_MAP_INIT_INT64 tables that only ever see kh_put of
// 64-bit keys and no deletions: same hash (khash.h:373), same triangular probing, same 0.77 load-factor growth with
// the in-place kick-out rehash.  order() returns the keys' indices (in insertion order numbering) by ascending slot,
// which is the order `for (i = kh_begin; i != kh_end; ++i) if (kh_exist(i))` visits them.
// ---------------------------------------------------------------------------------------------- sketch fallback pieces
// A read of which only some 512-position strips must be redone by the exact automaton (sketch_strip.cuh) is cut into pieces in
// position order: kind 0 = redone by the automaton, kind 1 = the fast path's records with positions in the piece stand.  A bad
// strip c is redone as a whole, and so are the last `reach` positions before it (the positions a window that ends in strip c
// can contain: w plus the palindromic k-mers the window may skip); automaton pieces are at most `seg` positions long.
struct SketchPiece { uint32_t lo, hi; uint8_t kind; };
inline void build_sketch_pieces(uint32_t len, uint64_t bad, uint32_t strip, uint32_t reach, uint32_t seg, std::vector<SketchPiece> &out) {
  const uint32_t n_strips = (len + strip - 1) / strip;
  uint32_t at = 0;  // positions below `at` are covered by the pieces so far
  for (uint32_t cs = 0; cs < n_strips && cs < 64; cs++) {
    if (!((bad >> cs) & 1)) continue;
    const uint32_t s0 = cs * strip;
    uint32_t lo = s0 > reach ? s0 - reach : 0, hi = std::min<uint32_t>(s0 + strip, len);
    if (lo < at) lo = at;
    if (lo >= hi) continue;
    if (lo > at) out.push_back(SketchPiece{at, lo, 1});
    for (uint32_t x = lo; x < hi; x += seg) out.push_back(SketchPiece{x, std::min<uint32_t>(x + seg, hi), 0});
    at = hi;
  }
  if (at < len) out.push_back(SketchPiece{at, len, 1});
}

struct KhashEmu {
  // One 8-byte slot per bucket: the key's 32-bit khash value (placement depends on the key only through it; keys are
  // distinct by contract, so equality is never tested) and insertion number + 1 with a generation bit on top
  // (0 = empty).  A rehash flips the generation: slots still carrying the old one are khash's "not yet moved" keys.
  //
  // Probe shortcut.  SHIMMER keys are hash<<8 | span with small hash values (they are minimizers of minimizers), so
  // khash's hash function maps them onto very few home slots (the low 8 bits of key ^ key<<11 ^ key>>33 are constant):
  // on the 50 Mb benchmark 195 k outer keys share ~1000 homes and every insertion walks ~100 occupied slots - in the
  // reference as well.  All keys of one home walk the SAME probe sequence, and between two rehashes slots only fill up,
  // so the first free probe step of a home never decreases: next[home] remembers where the last search of that home
  // ended and the next one resumes there.  Same layout, O(1) probes per insertion.
  struct Slot { uint32_t h, t; };
  uint32_t nb = 0, size = 0, nocc = 0, ub = 0, gen = 0;
  std::vector<Slot> slots;
  std::vector<uint32_t> next;  // per home slot: probe step at which to resume
  static inline uint32_t H(uint64_t key) { return (uint32_t)(key >> 33 ^ key ^ key << 11); }
  // first slot of `home`'s probe sequence (i += ++step, khash.h:226,315) for which taken(slot) is false
  template <class Taken>
  inline uint32_t probe(uint32_t home, uint32_t mask, Taken &&taken) {
    uint32_t step = next[home];
    uint32_t i = (uint32_t)((home + (uint64_t)step * (step + 1) / 2) & mask);
    while (taken(slots[i])) i = (i + (++step)) & mask;
    next[home] = step;
    return i;
  }
  void resize(uint32_t m) {
    --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
    if (m < 4) m = 4;
    if (size >= (uint32_t)(m * 0.77 + 0.5)) return;
    if (nb < m) slots.resize(m, Slot{0, 0});
    next.assign(m, 0);
    const uint32_t nmask = m - 1, ngen = gen ^ 0x80000000u;
    for (uint32_t j = 0; j != nb; ++j) {
      Slot s = slots[j];
      if (s.t == 0 || (s.t & 0x80000000u) == ngen) continue;
      slots[j].t = 0;
      s.t = (s.t & 0x7FFFFFFFu) | ngen;
      for (;;) {
        // khash.h:262-264: skip the slots that already hold a moved key
        const uint32_t i = probe(s.h & nmask, nmask, [ngen](const Slot &q) { return q.t != 0 && (q.t & 0x80000000u) == ngen; });
        if (slots[i].t != 0) {  // an unmoved key lives here: kick it out (khash.h:265-270)
          Slot k = slots[i];
          slots[i] = s;
          s = k;
          s.t = (s.t & 0x7FFFFFFFu) | ngen;
        } else {
          slots[i] = s;
          break;
        }
      }
    }
    if (nb > m) slots.resize(m);
    gen = ngen;
    nb = m;
    nocc = size;
    ub = (uint32_t)(nb * 0.77 + 0.5);
  }
  // the caller guarantees `key` was not inserted before; t < 2^31 - 1
  void put_new(uint64_t key, uint32_t t) {
    if (nocc >= ub) {
      if (nb > (size << 1)) resize(nb - 1);
      else resize(nb + 1);
    }
    const uint32_t mask = nb - 1, h = H(key);
    const uint32_t i = probe(h & mask, mask, [](const Slot &q) { return q.t != 0; });
    slots[i].h = h;
    slots[i].t = (t + 1) | gen;
    ++size;
    ++nocc;
  }
  // put_new of keys[0..n) with tags 0..n-1
  void put_all(const uint64_t *keys, uint32_t n) {
    {  // final table size is known: allocate once
      uint32_t m = 4;
      while (n >= (uint32_t)(m * 0.77 + 0.5) && m < 0x80000000u) m <<= 1;
      slots.reserve(m);
      next.reserve(m);
    }
    for (uint32_t o = 0; o < n; o++) put_new(keys[o], o);
  }
  // kh_put of a key that is ALREADY present still runs the load-factor check first (khash.h:289-297), so a put of an
  // existing key that follows the insertion which filled the table to its upper bound rehashes it.  Only the last such
  // put matters for the final layout (a pending rehash would otherwise happen, identically, at the next new key).
  void touch_existing() {
    if (nocc >= ub) {
      if (nb > (size << 1)) resize(nb - 1);
      else resize(nb + 1);
    }
  }
  // f(tag) in ascending slot order = the order `for (i = kh_begin; i != kh_end; ++i) if (kh_exist(i))` visits the keys
  template <class F>
  void for_each_in_slot_order(F &&f) const {
    for (uint32_t i = 0; i < nb; i++)
      if (slots[i].t) f((uint64_t)slots[i].h, (slots[i].t & 0x7FFFFFFFu) - 1);
  }
  void clear() { nb = size = nocc = ub = gen = 0; slots.clear(); next.clear(); }  // keeps the capacity
};

}  // namespace pgb
