/*
 * shimmer_oracle.c — CPU restatement (plain C) of the reference's SHIMMER index + overlap path.
 * TEST INFRASTRUCTURE ONLY — see shimmer_oracle.h for who may use it and how it is pinned to the reference.
 *
 * Written from the algorithm descriptions in SURVEY.md (App. A, D) and the cited reference lines; it shares no code with
 * the product (libpgb200.so works on 2-bit packed words on the GPU; this file works byte-per-base on the CPU).
 */
#include "shimmer_oracle.h"
#include <stdlib.h>
#include <string.h>

#define U64MAX 0xFFFFFFFFFFFFFFFFULL

static void mmv_push(orc_mmv *v, uint64_t x, uint64_t y) {
  if (v->n == v->m) {
    v->m = v->m ? v->m * 2 : 256;
    v->a = (orc_mm128 *)realloc(v->a, v->m * sizeof(orc_mm128));
  }
  v->a[v->n].x = x;
  v->a[v->n].y = y;
  v->n++;
}
void orc_free(void *p) { free(p); }

/* ---------------------------------------------------------------------------------------------- hash (mm_sketch.c:23-32) */
static uint64_t mix64(uint64_t key, uint64_t mask) {
  key = (~key + (key << 21)) & mask;
  key ^= key >> 24;
  key = (key + (key << 3) + (key << 8)) & mask;
  key ^= key >> 14;
  key = (key + (key << 2) + (key << 4)) & mask;
  key ^= key >> 28;
  key = (key + (key << 31)) & mask;
  return key;
}

static int nib_code(uint8_t nib) { /* A1 C2 G4 T8 -> 0..3, anything else -> 4 (decode_biseq + seq_nt4_table) */
  switch (nib & 0xF) {
    case 1: return 0;
    case 2: return 1;
    case 4: return 2;
    case 8: return 3;
    default: return 4;
  }
}

/* ---------------------------------------------------------------------------------------------- mm_sketch
 * State machine of src/mm_sketch.c:84-150: a ring of w slots holding (x,y) of the last w
 * non-palindromic positions, the current minimum and its slot. */
typedef struct { uint64_t x, y; } slot_t;

static void emit_equal_to_min(const slot_t *ring, int from, int to, const slot_t *mn, orc_mmv *out) {
  for (int j = from; j < to; j++)
    if (ring[j].x == mn->x && ring[j].y != mn->y) mmv_push(out, ring[j].x, ring[j].y);
}

void orc_sketch(const uint8_t *nib, int len, int w, int k, uint32_t rid, orc_mmv *out) {
  const uint64_t mask = (1ULL << (2 * k)) - 1, top = 2ULL * (uint64_t)(k - 1);
  slot_t ring[256], mn = {U64MAX, U64MAX};
  uint64_t fw = 0, rv = 0;
  int run = 0, at = 0, mn_at = 0;
  for (int j = 0; j < w; j++) ring[j].x = ring[j].y = U64MAX;
  for (int i = 0; i < len; i++) {
    int c = nib_code(nib[i]);
    slot_t cur = {U64MAX, U64MAX};
    if (c < 4) {
      fw = ((fw << 2) | (uint64_t)c) & mask;
      rv = (rv >> 2) | ((uint64_t)(3 ^ c) << top);
      if (fw == rv) continue; /* palindromic k-mer: position does not exist for the window (:104-105) */
      int strand = fw < rv ? 0 : 1;
      run++;
      if (run >= k) {
        cur.x = (mix64(strand ? rv : fw, mask) << 8) | (uint64_t)k;
        cur.y = ((uint64_t)rid << 32) | ((uint64_t)(uint32_t)i << 1) | (uint64_t)strand;
      }
    } else {
      run = 0; /* ambiguous base: restart the run, the slot still gets the sentinel (:113-114) */
    }
    ring[at] = cur;
    if (run == w + k - 1 && mn.x != U64MAX) { /* first full window: older ties of the minimum (:116-125) */
      emit_equal_to_min(ring, at + 1, w, &mn, out);
      emit_equal_to_min(ring, 0, at, &mn, out);
    }
    if (cur.x <= mn.x) { /* new minimum (ties go to the newer k-mer) (:126-128) */
      if (run >= w + k && mn.x != U64MAX) mmv_push(out, mn.x, mn.y);
      mn = cur;
      mn_at = at;
    } else if (at == mn_at) { /* the minimum just left the window (:129-147) */
      if (run >= w + k - 1 && mn.x != U64MAX) mmv_push(out, mn.x, mn.y);
      mn.x = U64MAX;
      for (int j = at + 1; j < w; j++)
        if (ring[j].x <= mn.x) { mn = ring[j]; mn_at = j; }
      for (int j = 0; j <= at; j++)
        if (ring[j].x <= mn.x) { mn = ring[j]; mn_at = j; }
      if (run >= w + k - 1 && mn.x != U64MAX) {
        emit_equal_to_min(ring, at + 1, w, &mn, out);
        emit_equal_to_min(ring, 0, at + 1, &mn, out);
      }
    }
    if (++at == w) at = 0;
  }
  if (mn.x != U64MAX) mmv_push(out, mn.x, mn.y); /* :150 */
}

/* ---------------------------------------------------------------------------------------------- mm_reduce
 * src/shmr_reduce.c:53-90: ring of rs slots per read, pick = lowest slot among the smallest x>>8,
 * emitted when its y differs from the last emitted y. */
void orc_reduce(const orc_mmv *in, orc_mmv *out, int rs) {
  orc_mm128 ring[256];
  uint64_t last_y = U64MAX;
  uint32_t cur_rid = 0xFFFFFFFFu, seen = 0;
  int head = 0;
  for (size_t i = 0; i < in->n; i++) {
    uint32_t rid = (uint32_t)(in->a[i].y >> 32);
    if (rid != cur_rid) { cur_rid = rid; seen = 0; head = 0; }
    ring[head] = in->a[i];
    head = (head + 1) % rs;
    seen++;
    if (seen < (uint32_t)rs) continue;
    int best = 0;
    for (int s = 1; s < rs; s++)
      if ((ring[s].x >> 8) < (ring[best].x >> 8)) best = s;
    if (ring[best].y != last_y) {
      mmv_push(out, ring[best].x, ring[best].y);
      last_y = ring[best].y;
    }
  }
}

/* ---------------------------------------------------------------------------------------------- ovlp_match
 * src/DWmatch.c:66-204 with O(band) storage: prev[]/cur[] hold the furthest x of the previous / current
 * d on diagonals lo..hi (step 2); U = x + y = 2x - k. */
void orc_ovlp_match(const uint8_t *q, int q_len, int q_strand, const uint8_t *t, int t_len, int t_strand, int bw, orc_match *out) {
  orc_match r;
  memset(&r, 0, sizeof r);
  int qs = q_strand ? 4 : 0, ts = t_strand ? 4 : 0;
  int max_d = (int)(0.3 * (q_len + t_len));
  int cap = bw + 4;
  int *prev = (int *)calloc((size_t)cap, sizeof(int)), *cur = (int *)calloc((size_t)cap, sizeof(int));
  int lo = 0, hi = 0, prev_lo = 0, best = -1, started = 0, matched = 0, x = 0, y = 0, d;
  uint32_t longest = 0;
  for (d = 0; d < max_d && !matched; d++) {
    if (hi - lo > 2 * bw) break;
    if (lo > hi) break; /* never reached on real data: the best diagonal always stays inside the band */
    int n = 0;
    for (int k = lo; k <= hi; k += 2, n++) {
      if (d == 0) x = 0;
      else if (k == lo) x = prev[(k + 1 - prev_lo) / 2];
      else if (k == hi) x = prev[(k - 1 - prev_lo) / 2] + 1;
      else {
        int a = prev[(k - 1 - prev_lo) / 2], b = prev[(k + 1 - prev_lo) / 2];
        x = a < b ? b : a + 1;
      }
      y = x - k;
      int x0 = x, y0 = y;
      while (x < q_len && y < t_len && ((q[x] >> qs) & 0xF) == ((t[y] >> ts) & 0xF)) { x++; y++; }
      if (x - x0 > 16 && !started) { r.q_bgn = x0; r.t_bgn = y0; started = 1; }
      if ((uint32_t)(x - x0) > longest) { longest = (uint32_t)(x - x0); r.q_m_end = x; r.t_m_end = y; }
      cur[n] = x;
      if (x + y > best) best = x + y;
      if (x >= q_len || y >= t_len) { matched = 1; break; }
    }
    if (matched) {
      r.q_end = x; r.t_end = y; r.dist = d;
      r.m_size = (r.q_end - r.q_bgn + r.t_end - r.t_bgn + 2 * d) / 2;
      break;
    }
    int nlo = hi, nhi = lo;
    n = 0;
    for (int k = lo; k <= hi; k += 2, n++)
      if (2 * cur[n] - k >= best - bw) { if (k < nlo) nlo = k; if (k > nhi) nhi = k; }
    prev_lo = lo; lo = nlo - 1; hi = nhi + 1;
    int *sw = prev; prev = cur; cur = sw;
  }
  if (!matched) { r.q_bgn = 0; r.t_bgn = 0; }
  free(prev); free(cur);
  *out = r;
}

/* ---------------------------------------------------------------------------------------------- index chunk */
void orc_index_chunk(const uint8_t *seqdb, const uint32_t *rid, const uint32_t *len, const uint64_t *off, size_t n_reads,
                     uint32_t T, uint32_t c, int w, int k, int r, int levels, orc_mmv out[3]) {
  for (size_t i = 0; i < n_reads; i++) {
    if (rid[i] % T != c % T) continue; /* src/shmr_index.c:157 */
    if (len[i] == 0) continue;
    orc_sketch(seqdb + off[i], (int)len[i], w, k, rid[i], &out[0]);
  }
  if (levels >= 1) orc_reduce(&out[0], &out[1], r);
  if (levels >= 2) orc_reduce(&out[1], &out[2], r);
}

/* ---------------------------------------------------------------------------------------------- small u64 -> u32 map */
typedef struct { uint64_t *k; uint32_t *v; size_t cap, n; } map64;
static size_t m_slot(const map64 *m, uint64_t key) {
  uint64_t h = key * 0x9E3779B97F4A7C15ULL;
  size_t i = (size_t)(h >> 20) & (m->cap - 1);
  while (m->k[i] != U64MAX && m->k[i] != key) i = (i + 1) & (m->cap - 1);
  return i;
}
static void m_init(map64 *m, size_t expect) {
  m->cap = 1024;
  while (m->cap < expect * 2 + 16) m->cap <<= 1;
  m->k = (uint64_t *)malloc(m->cap * 8);
  m->v = (uint32_t *)calloc(m->cap, 4);
  memset(m->k, 0xFF, m->cap * 8);
  m->n = 0;
}
static uint32_t *m_at(map64 *m, uint64_t key) { /* insert-or-find; caller sized the table */
  size_t i = m_slot(m, key);
  if (m->k[i] == U64MAX) { m->k[i] = key; m->n++; }
  return &m->v[i];
}
static int m_get(const map64 *m, uint64_t key, uint32_t *v) {
  size_t i = m_slot(m, key);
  if (m->k[i] == U64MAX) return 0;
  *v = m->v[i];
  return 1;
}
static void m_free(map64 *m) { free(m->k); free(m->v); }

static int cmp_mc(const void *a, const void *b) {
  uint64_t x = ((const orc_mc *)a)->mer, y = ((const orc_mc *)b)->mer;
  return x < y ? -1 : x > y;
}
orc_mc *orc_count(const orc_mm128 *a, size_t n, size_t *n_out) {
  map64 m;
  m_init(&m, n);
  for (size_t i = 0; i < n; i++) (*m_at(&m, a[i].x >> 8))++;
  orc_mc *r = (orc_mc *)calloc(m.n ? m.n : 1, sizeof(orc_mc));
  size_t o = 0;
  for (size_t i = 0; i < m.cap; i++)
    if (m.k[i] != U64MAX) { r[o].mer = m.k[i]; r[o].count = m.v[i]; o++; }
  qsort(r, o, sizeof(orc_mc), cmp_mc);
  *n_out = o;
  m_free(&m);
  return r;
}

/* ---------------------------------------------------------------------------------------------- khash visiting order
 * Keys-only model of klib khash (src/khash.h:218-343,373) without deletions — SURVEY App. D-1. */
typedef struct { uint32_t nb, size, nocc, ub; uint64_t *keys; uint32_t *tag; uint8_t *used; } kemu;
static uint32_t kh_hash(uint64_t key) { return (uint32_t)(key >> 33 ^ key ^ key << 11); }
static void kemu_resize(kemu *h, uint32_t m) {
  --m; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; ++m;
  if (m < 4) m = 4;
  if (h->size >= (uint32_t)(m * 0.77 + 0.5)) return;
  uint8_t *nused = (uint8_t *)calloc(m, 1);
  if (h->nb < m) {
    h->keys = (uint64_t *)realloc(h->keys, (size_t)m * 8);
    h->tag = (uint32_t *)realloc(h->tag, (size_t)m * 4);
    h->used = (uint8_t *)realloc(h->used, m);
    memset(h->used + h->nb, 0, m - h->nb);
  }
  for (uint32_t j = 0; j != h->nb; j++) {
    if (!h->used[j]) continue;
    uint64_t key = h->keys[j];
    uint32_t tg = h->tag[j];
    h->used[j] = 0;
    for (;;) {
      uint32_t i = kh_hash(key) & (m - 1), step = 0;
      while (nused[i]) i = (i + (++step)) & (m - 1);
      nused[i] = 1;
      if (i < h->nb && h->used[i]) { /* kick out the resident of the old table */
        uint64_t tk = h->keys[i]; h->keys[i] = key; key = tk;
        uint32_t tt = h->tag[i]; h->tag[i] = tg; tg = tt;
        h->used[i] = 0;
      } else {
        h->keys[i] = key; h->tag[i] = tg;
        break;
      }
    }
  }
  free(h->used);
  h->used = nused;
  h->nb = m; h->nocc = h->size; h->ub = (uint32_t)(m * 0.77 + 0.5);
}
static void kemu_put_new(kemu *h, uint64_t key, uint32_t tag) {
  if (h->nocc >= h->ub) kemu_resize(h, h->nb > (h->size << 1) ? h->nb - 1 : h->nb + 1);
  uint32_t i = kh_hash(key) & (h->nb - 1), step = 0;
  while (h->used[i]) i = (i + (++step)) & (h->nb - 1);
  h->keys[i] = key; h->tag[i] = tag; h->used[i] = 1; h->size++; h->nocc++;
}
/* a kh_put of an already-present key still runs the load check (khash.h:289-297) and may rehash */
static void kemu_touch(kemu *h) {
  if (h->nocc >= h->ub) kemu_resize(h, h->nb > (h->size << 1) ? h->nb - 1 : h->nb + 1);
}
static void kemu_free(kemu *h) { free(h->keys); free(h->tag); free(h->used); memset(h, 0, sizeof *h); }

/* ---------------------------------------------------------------------------------------------- overlap chunk */
typedef struct { uint64_t k0, k1, y0; uint32_t seq; uint8_t dir; } prec;
static int cmp_prec(const void *a, const void *b) { /* group by (k0,k1), insertion order inside */
  const prec *p = (const prec *)a, *q = (const prec *)b;
  if (p->k0 != q->k0) return p->k0 < q->k0 ? -1 : 1;
  if (p->k1 != q->k1) return p->k1 < q->k1 ? -1 : 1;
  return p->seq < q->seq ? -1 : p->seq > q->seq;
}
typedef struct { uint32_t first, n, first_seq, last_seq; } bucket_t;
static const bucket_t *g_b;
static int cmp_bucket_seq(const void *a, const void *b) {
  uint32_t x = g_b[*(const uint32_t *)a].first_seq, y = g_b[*(const uint32_t *)b].first_seq;
  return x < y ? -1 : x > y;
}
static uint64_t rev_coord(uint64_t y, uint64_t x, uint32_t rlen) { /* src/shmr_utils.c:376-395 */
  uint32_t span = (uint32_t)(x & 0xFF), pos = (uint32_t)((y & 0xFFFFFFFFULL) >> 1) + 1;
  uint32_t rpos = rlen - pos + span - 1;
  return ((y & 0xFFFFFFFF00000001ULL) | (uint64_t)(uint32_t)(rpos << 1)) ^ 1ULL;
}
static uint32_t ypos(uint64_t y) { return (uint32_t)((y & 0xFFFFFFFFULL) >> 1); }

orc_ovlp *orc_overlap_chunk(const uint8_t *seqdb, const uint32_t *rid, const uint32_t *len, const uint64_t *off, size_t n_reads,
                            const orc_mm128 *mm, size_t n_mm, const orc_mc *mc, size_t n_mc, uint32_t T, uint32_t c,
                            uint32_t bestn, uint32_t mc_lower, uint32_t mc_upper, int bw, uint32_t ovlp_upper, size_t *n_out,
                            uint64_t *n_alignments) {
  *n_out = 0;
  if (n_alignments) *n_alignments = 0;
  /* read table by rid */
  uint32_t max_rid = 0;
  for (size_t i = 0; i < n_reads; i++) if (rid[i] > max_rid) max_rid = rid[i];
  uint32_t *rl = (uint32_t *)calloc((size_t)max_rid + 1, 4);
  uint64_t *ro = (uint64_t *)calloc((size_t)max_rid + 1, 8);
  for (size_t i = 0; i < n_reads; i++) { rl[rid[i]] = len[i]; ro[rid[i]] = off[i]; }
  /* aggregate counts (src/shmr_utils.c:162-176) */
  map64 cnt;
  m_init(&cnt, n_mc);
  for (size_t i = 0; i < n_mc; i++) *m_at(&cnt, mc[i].mer) += mc[i].count;
  /* build_map (src/shmr_utils.c:295-404) */
  prec *recs = (prec *)malloc((2 * n_mm + 2) * sizeof(prec));
  size_t nrec = 0, s = 0;
  uint32_t seq = 0, mcnt = 0;
  for (; s < n_mm; s++) {
    m_get(&cnt, mm[s].x >> 8, &mcnt);
    if (mcnt >= mc_lower && mcnt < mc_upper) break;
  }
  if (s < n_mm) {
    orc_mm128 m0 = mm[s];
    for (size_t i = s + 1; i < n_mm; i++) {
      orc_mm128 m1 = mm[i];
      mcnt = 0;
      m_get(&cnt, m1.x >> 8, &mcnt);
      if (mcnt < mc_lower || mcnt > mc_upper) continue;
      seq += 2; /* one sequence slot pair per consecutive kept pair */
      if ((m0.y >> 32) == (m1.y >> 32)) {
        if ((((m1.y >> 1) & 0xFFFFFFFULL) - ((m0.y >> 1) & 0xFFFFFFFULL)) < 100ULL) { m0 = m1; continue; }
        if ((m0.x >> 8) % T == c % T) { prec p = {m0.x, m1.x, m0.y, seq, 0}; recs[nrec++] = p; }
        if ((m1.x >> 8) % T == c % T) {
          prec p = {m1.x, m0.x, rev_coord(m1.y, m1.x, rl[(uint32_t)(m1.y >> 32)]), seq + 1, 1};
          recs[nrec++] = p;
        }
      }
      m0 = m1;
    }
  }
  m_free(&cnt);
  /* buckets = runs of equal (k0,k1) */
  qsort(recs, nrec, sizeof(prec), cmp_prec);
  bucket_t *b = (bucket_t *)malloc((nrec + 1) * sizeof(bucket_t));
  size_t nb = 0;
  for (size_t i = 0; i < nrec; i++) {
    if (i == 0 || recs[i].k0 != recs[i - 1].k0 || recs[i].k1 != recs[i - 1].k1) { b[nb].first = (uint32_t)i; b[nb].n = 0; b[nb].first_seq = recs[i].seq; nb++; }
    b[nb - 1].n++;
    b[nb - 1].last_seq = recs[i].seq;
  }
  /* visiting order: outer / inner khash slot order given first-insertion order */
  uint32_t *by_seq = (uint32_t *)malloc((nb + 1) * 4);
  for (size_t i = 0; i < nb; i++) by_seq[i] = (uint32_t)i;
  g_b = b;
  qsort(by_seq, nb, 4, cmp_bucket_seq);
  kemu outer;
  memset(&outer, 0, sizeof outer);
  map64 oid;
  m_init(&oid, nb);
  uint32_t *head = (uint32_t *)malloc((nb + 1) * 4), *tail = (uint32_t *)malloc((nb + 1) * 4), *next = (uint32_t *)malloc((nb + 1) * 4);
  uint32_t *olast = (uint32_t *)calloc(nb + 1, 4);
  uint32_t n_outer = 0, newest_outer = 0, last_all = 0;
  for (size_t i = 0; i < nb; i++) {
    uint32_t bi = by_seq[i];
    if (b[bi].last_seq > last_all) last_all = b[bi].last_seq;
    next[bi] = 0xFFFFFFFFu;
    uint64_t k0 = recs[b[bi].first].k0;
    uint32_t *id = m_at(&oid, k0);
    if (*id == 0) {
      *id = ++n_outer; /* ids are 1-based inside the map */
      kemu_put_new(&outer, k0, n_outer - 1);
      newest_outer = b[bi].first_seq;
      head[n_outer - 1] = tail[n_outer - 1] = bi;
      olast[n_outer - 1] = b[bi].last_seq;
    } else {
      next[tail[*id - 1]] = bi;
      tail[*id - 1] = bi;
      if (b[bi].last_seq > olast[*id - 1]) olast[*id - 1] = b[bi].last_seq;
    }
  }
  if (last_all > newest_outer) kemu_touch(&outer);
  m_free(&oid);
  /* process_overlaps (src/shmr_overlap.c:182-231) + shimmer_to_overlap (:52-180) */
  map64 pairs; /* rid_pairs: value = type + 1 */
  m_init(&pairs, nrec * 2 + 1024);
  size_t pairs_limit = pairs.cap / 2;
  orc_ovlp *out = NULL;
  size_t n = 0, m = 0;
  prec *sorted = (prec *)malloc((ovlp_upper + 1) * sizeof(prec));
  uint8_t *contained = (uint8_t *)malloc(ovlp_upper + 1);
  uint64_t n_aln = 0;
  for (uint32_t os = 0; os < outer.nb; os++) {
    if (!outer.used[os]) continue;
    kemu inner;
    memset(&inner, 0, sizeof inner);
    uint32_t newest_inner = 0;
    for (uint32_t bi = head[outer.tag[os]]; bi != 0xFFFFFFFFu; bi = next[bi]) { kemu_put_new(&inner, recs[b[bi].first].k1, bi); newest_inner = b[bi].first_seq; }
    if (olast[outer.tag[os]] > newest_inner) kemu_touch(&inner);
    for (uint32_t is = 0; is < inner.nb; is++) {
      if (!inner.used[is]) continue;
      const bucket_t *bk = &b[inner.tag[is]];
      uint32_t nn = bk->n;
      if (nn <= 2 || nn > ovlp_upper) continue; /* :216 */
      /* glibc qsort + boolean comparator == stable sort, descending position (SURVEY a-9): insertion sort keeps stability */
      for (uint32_t i = 0; i < nn; i++) {
        prec p = recs[bk->first + i];
        uint32_t j = i;
        while (j > 0 && ypos(sorted[j - 1].y0) < ypos(p.y0)) { sorted[j] = sorted[j - 1]; j--; }
        sorted[j] = p;
      }
      memset(contained, 0, nn);
      for (uint32_t i = nn - 1; i-- > 0;) {
        if (contained[i]) continue;
        uint32_t rid0 = (uint32_t)(sorted[i].y0 >> 32), pos0 = ypos(sorted[i].y0) + 1, rlen0 = rl[rid0], oc = 0;
        for (uint32_t j = i + 1; j < nn && oc < bestn; j++) {
          if (contained[j]) continue;
          uint32_t rid1 = (uint32_t)(sorted[j].y0 >> 32);
          if (rid0 == rid1) continue;
          uint64_t key = rid0 < rid1 ? ((uint64_t)rid0 << 32) | rid1 : ((uint64_t)rid1 << 32) | rid0;
          uint32_t seen = 0;
          if (m_get(&pairs, key, &seen) && seen) { if (seen - 1 == 0) oc++; continue; }
          uint32_t pos1 = ypos(sorted[j].y0) + 1, rlen1 = rl[rid1];
          uint32_t slen0 = rlen0 - pos0 + pos1, slen1 = rlen1;
          orc_match mt;
          orc_ovlp_match(seqdb + ro[rid0] + (pos0 - pos1), (int)slen0, sorted[i].dir, seqdb + ro[rid1], (int)slen1, sorted[j].dir, bw, &mt);
          n_aln++;
          long dq = (long)slen0 - mt.q_end, dt = (long)slen1 - mt.t_end;
          if (dq < 0) dq = -dq;
          if (dt < 0) dt = -dt;
          if (mt.q_bgn < 48 && mt.t_bgn < 48 && (dq < 48 || dt < 48) && mt.q_end > 500 && mt.t_end > 500) { /* :134-137 */
            long c0 = (long)rlen0 - (mt.q_end - mt.q_bgn), c1 = (long)rlen1 - (mt.t_end - mt.t_bgn);
            if (c0 < 0) c0 = -c0;
            if (c1 < 0) c1 = -c1;
            uint32_t type;
            if (c0 < 96 || c1 < 96) {
              if (rlen0 >= rlen1) { type = 1; contained[j] = 1; } else { type = 2; contained[i] = 1; }
            } else { type = 0; oc++; }
            if (pairs.n + 1 >= pairs_limit) { /* grow */
              map64 bigger;
              m_init(&bigger, pairs.cap);
              for (size_t q = 0; q < pairs.cap; q++) if (pairs.k[q] != U64MAX) *m_at(&bigger, pairs.k[q]) = pairs.v[q];
              m_free(&pairs);
              pairs = bigger;
              pairs_limit = pairs.cap / 2;
            }
            *m_at(&pairs, key) = type + 1;
            if (n == m) { m = m ? m * 2 : 1024; out = (orc_ovlp *)realloc(out, m * sizeof(orc_ovlp)); }
            orc_ovlp *o = &out[n++];
            memset(o, 0, sizeof *o);
            o->y0 = sorted[i].y0; o->y1 = sorted[j].y0; o->rl0 = rlen0; o->rl1 = rlen1;
            o->strand0 = sorted[i].dir; o->strand1 = sorted[j].dir; o->ovlp_type = (uint8_t)type; o->match = mt;
          }
          if (contained[i]) break; /* :176 */
        }
      }
    }
    kemu_free(&inner);
  }
  kemu_free(&outer);
  m_free(&pairs);
  free(sorted); free(contained); free(olast); free(head); free(tail); free(next); free(by_seq); free(b); free(recs); free(rl); free(ro);
  *n_out = n;
  if (n_alignments) *n_alignments = n_aln;
  if (!out) out = (orc_ovlp *)calloc(1, sizeof(orc_ovlp));
  return out;
}
