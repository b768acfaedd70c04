/*
 * stages_oracle.c — CPU restatement (plain C) of the stages next to the index/overlap path:
 * shmr_mkseqdb (FASTA/FASTQ records + encode_biseq), shmr_dedup, shmr_map.
 * TEST INFRASTRUCTURE ONLY (see shimmer_oracle.h).  Parity status: PINNED — tests/test_oracle_stages.py checks every
 * function against the unmodified reference binaries in oracle/_ref and against the committed golden vectors
 * (tests/golden/golden_stages.json, generated from the same reference build).
 * Shares no code with the product: records are scanned character by character like kseq does, text is formatted by libc's
 * printf like the reference does, the pair index is a sorted array.
 */
#include "shimmer_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ encode_biseq */
/* src/shmr_utils.c:18-51: byte p = code(seq[p]) | complement_code(seq[len-1-p]) << 4, A/a 1 C/c 2 G/g 4 T/t 8 */
static uint8_t code_f(char c) {
  switch (c) { case 'A': case 'a': return 1; case 'C': case 'c': return 2; case 'G': case 'g': return 4; case 'T': case 't': return 8; default: return 0; }
}
static uint8_t code_r(char c) {
  switch (c) { case 'A': case 'a': return 8; case 'C': case 'c': return 4; case 'G': case 'g': return 2; case 'T': case 't': return 1; default: return 0; }
}
void orc_encode_biseq(const char *seq, size_t len, uint8_t *out) {
  for (size_t p = 0; p < len; p++) out[p] = (uint8_t)(code_f(seq[p]) | (code_r(seq[len - 1 - p]) << 4));
}

/* ------------------------------------------------------------------------------------------------ FASTA / FASTQ records */
/* src/kseq.h:185-224 as a character-level state machine over a memory buffer (the reference streams through gzread).
 * Calls rec(name, name_len, seq, seq_len, user) per record, in file order; returns the number of records. */
typedef struct { const char *b; size_t n, pos; } cur_t;
static int cgetc(cur_t *c) { return c->pos < c->n ? (unsigned char)c->b[c->pos++] : -1; }
static int is_sp(int ch) { return ch == ' ' || (ch >= '\t' && ch <= '\r'); }
typedef struct { char *s; size_t l, m; } str_t;
static void sput(str_t *s, int ch) {
  if (s->l + 1 >= s->m) { s->m = s->m ? 2 * s->m : 256; s->s = (char *)realloc(s->s, s->m); }
  s->s[s->l++] = (char)ch;
}
/* rest of the line appended to s; 1 if anything (even an empty line) was consumed, 0 at end of input (ks_getuntil2 -> -1) */
static int append_line(cur_t *c, str_t *s) {
  if (c->pos >= c->n) return 0;
  int ch;
  while ((ch = cgetc(c)) != -1 && ch != '\n') sput(s, ch);
  if (s->l > 1 && s->s[s->l - 1] == '\r') s->l--;  /* kseq.h:138 */
  return 1;
}
size_t orc_fasta_records(const char *buf, size_t n, void (*rec)(const char *, size_t, const char *, size_t, void *), void *user) {
  cur_t c = {buf, n, 0};
  str_t name = {0, 0, 0}, seq = {0, 0, 0}, qual = {0, 0, 0};
  int last = 0, ch;
  size_t count = 0;
  for (;;) {
    if (last == 0) {  /* :189-192 */
      while ((ch = cgetc(&c)) != -1 && ch != '>' && ch != '@') {}
      if (ch == -1) break;
      last = ch;
    }
    if (c.pos >= c.n) break;  /* :195 ks_getuntil < 0 */
    name.l = seq.l = qual.l = 0;
    while ((ch = cgetc(&c)) != -1 && !is_sp(ch)) sput(&name, ch);
    if (ch != '\n' && ch != -1) while ((ch = cgetc(&c)) != -1 && ch != '\n') {}  /* comment, :196 */
    while ((ch = cgetc(&c)) != -1 && ch != '>' && ch != '+' && ch != '@') {  /* :201-205 */
      if (ch == '\n') continue;
      sput(&seq, ch);
      append_line(&c, &seq);
    }
    if (ch == '>' || ch == '@') last = ch;
    if (ch == '+') {  /* FASTQ, :215-223 */
      while ((ch = cgetc(&c)) != -1 && ch != '\n') {}
      if (ch == -1) break;  /* -2: no quality */
      while (append_line(&c, &qual) && qual.l < seq.l) {}
      last = 0;
      if (qual.l != seq.l) break;  /* -2: the caller's loop ends (src/shmr_mkseqdb.c:108) */
    }
    rec(name.s ? name.s : "", name.l, seq.s ? seq.s : "", seq.l, user);
    count++;
  }
  free(name.s); free(seq.s); free(qual.s);
  return count;
}

/* ------------------------------------------------------------------------------------------------ shmr_dedup */
/* src/shmr_dedup.c:19-101: first record of every unordered read pair, one text line each.  Returns malloc'd text. */
typedef struct { uint64_t key; size_t idx; } kidx_t;
static int cmp_kidx(const void *a, const void *b) {
  const kidx_t *x = (const kidx_t *)a, *y = (const kidx_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : x->idx > y->idx;
}
char *orc_dedup(const orc_ovlp *recs, size_t n, size_t *text_len) {
  kidx_t *ki = (kidx_t *)malloc((n + 1) * sizeof(kidx_t));
  uint8_t *keep = (uint8_t *)calloc(n + 1, 1);
  for (size_t i = 0; i < n; i++) {
    uint32_t r0 = (uint32_t)(recs[i].y0 >> 32), r1 = (uint32_t)(recs[i].y1 >> 32);
    ki[i].key = r0 < r1 ? ((uint64_t)r0 << 32) | r1 : ((uint64_t)r1 << 32) | r0;  /* :37-40 */
    ki[i].idx = i;
  }
  qsort(ki, n, sizeof(kidx_t), cmp_kidx);
  for (size_t i = 0; i < n; i++)
    if (i == 0 || ki[i].key != ki[i - 1].key) keep[ki[i].idx] = 1;  /* smallest stream index of the pair = first seen */
  size_t cap = 256 * (n + 1), len = 0;
  char *text = (char *)malloc(cap);
  for (size_t i = 0; i < n; i++) {
    if (!keep[i]) continue;
    const orc_ovlp *o = &recs[i];
    /* the reference's own declarations and statements, :22-25,46-95 */
    uint32_t a_bgn, a_end, b_bgn, b_end;
    uint32_t rid0 = (uint32_t)(o->y0 >> 32), rid1 = (uint32_t)(o->y1 >> 32);
    uint32_t pos0 = (uint32_t)((o->y0 & 0xFFFFFFFF) >> 1) + 1, rlen0 = o->rl0;
    uint32_t pos1 = (uint32_t)((o->y1 & 0xFFFFFFFF) >> 1) + 1, rlen1 = o->rl1;
    uint8_t strand0 = o->strand0, strand1 = o->strand1;
    int32_t q_bgn = o->match.q_bgn, q_end = o->match.q_end, t_bgn = o->match.t_bgn, t_end = o->match.t_end;
    q_bgn = (int32_t)((uint32_t)q_bgn - (uint32_t)t_bgn);
    t_bgn = 0;
    if (strand0 == 0) {
      a_bgn = (uint32_t)((int32_t)(pos0 - pos1)) + (uint32_t)q_bgn;
      a_end = (uint32_t)((int32_t)(pos0 - pos1)) + (uint32_t)q_end;
      a_end = a_end >= rlen0 ? rlen0 : a_end;
    } else {
      a_bgn = rlen0 - (uint32_t)((int32_t)(pos0 - pos1)) - (uint32_t)q_end;
      a_end = rlen0 - (uint32_t)((int32_t)(pos0 - pos1)) - (uint32_t)q_bgn;
      a_end = a_end >= rlen0 ? rlen0 : a_end;
    }
    if (strand1 == 0) {
      b_bgn = (uint32_t)t_bgn;
      b_end = (uint32_t)t_end;
      b_end = b_end >= rlen1 ? rlen1 : b_end;
    } else {
      b_bgn = rlen1 - (uint32_t)t_end;
      b_end = rlen1 - (uint32_t)t_bgn;
      b_end = b_end >= rlen1 ? rlen1 : b_end;
    }
    volatile double num = 100.0 * (double)(o->match.dist);
    double err_est = 100.0 - num / (double)(o->match.m_size);
    len += (size_t)snprintf(text + len, cap - len, "%09d %09d %d %0.1f %u %d %d %u %u %d %d %u %s\n", (int)rid0, (int)rid1,
                            (int)(0u - (uint32_t)o->match.m_size), err_est, 0u, (int)a_bgn, (int)a_end, rlen0,
                            (unsigned)(strand0 == 0 ? strand1 : 1 - strand1), (int)b_bgn, (int)b_end, rlen1,
                            o->ovlp_type == 0 ? "overlap" : (o->ovlp_type == 1 ? "contains" : "contained"));
  }
  free(ki); free(keep);
  *text_len = len;
  return text;
}

/* ------------------------------------------------------------------------------------------------ shmr_map */
/* build_map (src/shmr_utils.c:295-404) over the reads' shimmers + process_map (src/shmr_map.c:48-166) over the contigs'.
 * rlen_by_rid covers every read id.  Returns malloc'd text ("%u %u %u %u %u %u %d %u %u\n" per hit). */
typedef struct { uint64_t k0, k1, y0, y1; size_t seq; uint8_t dir; } mrec_t;
static int cmp_mrec(const void *a, const void *b) {
  const mrec_t *x = (const mrec_t *)a, *y = (const mrec_t *)b;
  if (x->k0 != y->k0) return x->k0 < y->k0 ? -1 : 1;
  if (x->k1 != y->k1) return x->k1 < y->k1 ? -1 : 1;
  return x->seq < y->seq ? -1 : x->seq > y->seq;
}
static int cmp_mc_mer(const void *a, const void *b) {
  uint64_t x = ((const orc_mc *)a)->mer, y = ((const orc_mc *)b)->mer;
  return x < y ? -1 : x > y;
}
static int count_of(const orc_mc *tab, size_t n, uint64_t mer, uint32_t *out) {
  size_t lo = 0, hi = n;
  while (lo < hi) { size_t mid = (lo + hi) / 2; if (tab[mid].mer < mer) lo = mid + 1; else hi = mid; }
  if (lo < n && tab[lo].mer == mer) { *out = tab[lo].count; return 1; }
  return 0;
}
static uint64_t map_rev(uint64_t y, uint64_t x, uint32_t rlen) {  /* src/shmr_utils.c:376-395 */
  uint32_t span = (uint32_t)(x & 0xFF), pos = (uint32_t)((y & 0xFFFFFFFFULL) >> 1) + 1;
  uint32_t rpos = rlen - pos + span - 1;
  return ((y & 0xFFFFFFFF00000001ULL) | (uint64_t)(uint32_t)(rpos << 1)) ^ 1ULL;
}
/* first record with k0 == x (and k1 == x1 if with_k1), or n */
static size_t find_rec(const mrec_t *r, size_t n, uint64_t x, int with_k1, uint64_t x1) {
  size_t lo = 0, hi = n;
  while (lo < hi) {
    size_t mid = (lo + hi) / 2;
    int less = r[mid].k0 < x || (with_k1 && r[mid].k0 == x && r[mid].k1 < x1);
    if (less) lo = mid + 1; else hi = mid;
  }
  if (lo < n && r[lo].k0 == x && (!with_k1 || r[lo].k1 == x1)) return lo;
  return n;
}
char *orc_map(const orc_mm128 *ref, size_t n_ref, const orc_mm128 *mm, size_t n_mm, const orc_mc *mc_in, size_t n_mc,
              const uint32_t *rlen_by_rid, uint32_t T, uint32_t c, uint32_t lower, uint32_t upper, size_t *text_len) {
  /* aggregate_mm_count, src/shmr_utils.c:162-176 */
  orc_mc *mc = (orc_mc *)malloc((n_mc + 1) * sizeof(orc_mc));
  memcpy(mc, mc_in, n_mc * sizeof(orc_mc));
  qsort(mc, n_mc, sizeof(orc_mc), cmp_mc_mer);
  size_t nm = 0;
  for (size_t i = 0; i < n_mc; i++) {
    if (nm && mc[nm - 1].mer == mc[i].mer) mc[nm - 1].count += mc[i].count;
    else mc[nm++] = mc[i];
  }
  /* build_map */
  mrec_t *recs = (mrec_t *)malloc((2 * n_mm + 2) * sizeof(mrec_t));
  size_t nrec = 0, s = 0, seq = 0;
  uint32_t cnt = 0;
  for (; s < n_mm; s++) {
    cnt = 0;
    count_of(mc, nm, mm[s].x >> 8, &cnt);
    if (cnt >= lower && cnt < upper) break;
  }
  if (s < n_mm) {
    orc_mm128 m0 = mm[s];
    for (size_t i = s + 1; i < n_mm; i++) {
      orc_mm128 m1 = mm[i];
      cnt = 0;
      count_of(mc, nm, m1.x >> 8, &cnt);
      if (cnt < lower || cnt > upper) continue;
      if ((m0.y >> 32) == (m1.y >> 32)) {
        if ((((m1.y >> 1) & 0xFFFFFFFULL) - ((m0.y >> 1) & 0xFFFFFFFULL)) < 100ULL) { m0 = m1; continue; }
        uint32_t rl = rlen_by_rid[(uint32_t)(m0.y >> 32)];
        if ((m0.x >> 8) % T == c % T) { mrec_t p = {m0.x, m1.x, m0.y, m1.y, seq++, 0}; recs[nrec++] = p; }
        if ((m1.x >> 8) % T == c % T) { mrec_t p = {m1.x, m0.x, map_rev(m1.y, m1.x, rl), map_rev(m0.y, m0.x, rl), seq++, 1}; recs[nrec++] = p; }
      }
      m0 = m1;
    }
  }
  qsort(recs, nrec, sizeof(mrec_t), cmp_mrec);
  /* process_map */
  size_t cap = 1 << 16, len = 0;
  char *text = (char *)malloc(cap);
  size_t st = 0;
  for (; st < n_ref; st++) if (find_rec(recs, nrec, ref[st].x, 0, 0) < nrec) break;  /* :85-91 */
  if (st < n_ref) {
    orc_mm128 m0 = ref[st];
    for (size_t i = st + 1; i < n_ref; i++) {
      orc_mm128 m1 = ref[i];
      uint32_t c1 = 0;
      if (!count_of(mc, nm, m1.x >> 8, &c1)) continue;  /* :96-97 */
      if (c1 < lower || c1 > upper) continue;
      size_t b = nrec;
      if ((m0.y >> 32) == (m1.y >> 32)) b = find_rec(recs, nrec, m0.x, 1, m1.x);
      if (b < nrec && (((m1.y >> 1) & 0xFFFFFFFULL) - ((m0.y >> 1) & 0xFFFFFFFULL)) >= 100ULL) {
        uint32_t c0 = 0;
        count_of(mc, nm, m0.x >> 8, &c0);
        for (size_t j = b; j < nrec && recs[j].k0 == m0.x && recs[j].k1 == m1.x; j++) {
          if (cap - len < 160) { cap *= 2; text = (char *)realloc(text, cap); }
          len += (size_t)snprintf(text + len, cap - len, "%u %u %u %u %u %u %d %u %u\n", (uint32_t)(m0.y >> 32),
                                  (uint32_t)((m0.y & 0xFFFFFFFF) >> 1), (uint32_t)((m1.y & 0xFFFFFFFF) >> 1), (uint32_t)(recs[j].y0 >> 32),
                                  (uint32_t)((recs[j].y0 & 0xFFFFFFFF) >> 1), (uint32_t)((recs[j].y1 & 0xFFFFFFFF) >> 1), recs[j].dir, c0, c1);
        }
      }
      m0 = m1;
    }
  }
  free(mc); free(recs);
  *text_len = len;
  return text;
}
