// khash_small.cuh — fixed-capacity keys-only model of klib khash (src/khash.h:218-343,373) for the INNER tables
// MMER1[x0] of build_map (src/shmr_utils.c:341-347): one GPU thread replays the kh_put sequence of one outer key to
// obtain the slot order in which process_overlaps (src/shmr_overlap.c:211) visits that key's buckets.
// Capacity 64 slots = at most 48 distinct inner keys (the 49th insertion would grow the table to 128); larger groups
// are replayed on the host with the unbounded KhashEmu of host_util.hpp.  Same algorithm as KhashEmu; see there.
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

enum { KHS_CAP = 64, KHS_MAX_KEYS = 48 };

struct KhSmall {
  uint32_t nb, size, nocc, ub;
  uint64_t used;  // bit i = slot i occupied
  uint64_t keys[KHS_CAP];
  uint8_t tag[KHS_CAP];

  PGB_HD void init() { nb = size = nocc = ub = 0; used = 0; }
  static PGB_HD uint32_t H(uint64_t key) { return (uint32_t)(key >> 33 ^ key ^ key << 11); }
  PGB_HD void resize(uint32_t m) {
    --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
    if (m < 4) m = 4;
    if (size >= (uint32_t)(m * 0.77 + 0.5)) return;
    uint64_t nused = 0;
    const uint32_t nmask = m - 1;
    for (uint32_t j = 0; j != nb; ++j) {
      if (!(used >> j & 1)) continue;
      uint64_t key = keys[j];
      uint8_t tg = tag[j];
      used &= ~(1ULL << j);
      for (;;) {
        uint32_t i = H(key) & nmask, step = 0;
        while (nused >> i & 1) i = (i + (++step)) & nmask;
        nused |= 1ULL << i;
        if (i < nb && (used >> i & 1)) {  // kick out the resident of the old table
          uint64_t tk = keys[i]; keys[i] = key; key = tk;
          uint8_t tt = tag[i]; tag[i] = tg; tg = tt;
          used &= ~(1ULL << i);
        } else {
          keys[i] = key;
          tag[i] = tg;
          break;
        }
      }
    }
    used = nused;
    nb = m;
    nocc = size;
    ub = (uint32_t)(nb * 0.77 + 0.5);
  }
  PGB_HD void grow_check() {
    if (nocc >= ub) {
      if (nb > (size << 1)) resize(nb - 1);
      else resize(nb + 1);
    }
  }
  // the caller guarantees `key` is new and that size stays <= KHS_MAX_KEYS
  PGB_HD void put_new(uint64_t key, uint8_t t) {
    grow_check();
    const uint32_t mask = nb - 1;
    uint32_t i = H(key) & mask, step = 0;
    while (used >> i & 1) i = (i + (++step)) & mask;
    keys[i] = key;
    tag[i] = t;
    used |= 1ULL << i;
    ++size;
    ++nocc;
  }
  PGB_HD void touch_existing() { grow_check(); }
};

// Slot-order rank of each of the n (<= KHS_MAX_KEYS) distinct inner keys, inserted in the given order; trailing = a put of
// an existing key follows the last new key (SURVEY App. A-3 addendum in DESIGN.md).  rank_out[i] = visiting position of key i.
PGB_HD void khs_order(const uint64_t *keys, uint32_t n, bool trailing, uint8_t *rank_out) {
  KhSmall h;
  h.init();
  for (uint32_t i = 0; i < n; i++) h.put_new(keys[i], (uint8_t)i);
  if (trailing) h.touch_existing();
  uint32_t r = 0;
  for (uint32_t s = 0; s < h.nb; s++)
    if (h.used >> s & 1) rank_out[h.tag[s]] = (uint8_t)r++;
}

}  // namespace pgb
