// hostsim — CPU simulator of the kernel logic in peregrine_b200/csrc/shimmer_core.cuh.
//
// TEST INFRASTRUCTURE ONLY.  There is no GPU in the development container, so the per-item device functions
// (sketch automaton, reduce pick, packed-word ovlp_match, bucket replay) are compiled for the host here and
// compared with the real reference (oracle/_ref/libshimmer_ref.so, built from the unmodified sources) before
// a GPU minute is spent.  libpgb200.so does not contain or call any of this.
//
//   hostsim sketch  <seqdb_prefix> <w> <k>            every read: sketch_exact vs reference mm_sketch, reduce x2
//   hostsim match   <seqdb_prefix> <n_pairs> <bw>     random + adversarial operand pairs vs reference ovlp_match
//   hostsim overlap <seqdb_prefix> <l2_prefix> <T> <c> <ref_ovlp_file> [jacobi]
//   hostsim fastaidx <file list>                       the product's FASTA/FASTQ scanner: prints the .idx lines shmr_mkseqdb would
//   hostsim dedup   < ovlp stream > text                the product's dedup_pair_key / dedup_format, first-seen per pair
#include <dlfcn.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <unordered_map>
#include <map>
#include <array>
#include <chrono>
#include "../../peregrine_b200/csrc/host_util.hpp"
#include "../../peregrine_b200/csrc/sketch_tile.cuh"
#include "../../peregrine_b200/csrc/khash_small.cuh"
#include "../../peregrine_b200/csrc/dedup.cuh"
#include "../../peregrine_b200/csrc/fasta_reader.hpp"
#include <unordered_set>
#include <random>

using namespace pgb;

// ---- reference ABI (src/shimmer.h) ----
typedef struct { size_t n, m; mm128 *a; } mm128_v;
typedef void (*ref_mm_sketch_t)(void *, const char *, int, int, int, uint32_t, int, mm128_v *);
typedef void (*ref_mm_reduce_t)(mm128_v *, mm128_v *, uint8_t);
typedef void (*ref_decode_t)(uint8_t *, char *, size_t, uint8_t);
typedef match_t *(*ref_ovlp_match_t)(uint8_t *, int32_t, uint8_t, uint8_t *, int32_t, uint8_t, int32_t);
static ref_mm_sketch_t ref_mm_sketch;
static ref_mm_reduce_t ref_mm_reduce;
static ref_decode_t ref_decode;
static ref_ovlp_match_t ref_ovlp_match;

static void load_ref() {
  const char *p = getenv("PGB_REF_LIB");
  void *h = dlopen(p ? p : "oracle/_ref/libshimmer_ref.so", RTLD_NOW);
  if (!h) { fprintf(stderr, "cannot load reference lib: %s\n", dlerror()); exit(2); }
  ref_mm_sketch = (ref_mm_sketch_t)dlsym(h, "mm_sketch");
  ref_mm_reduce = (ref_mm_reduce_t)dlsym(h, "mm_reduce");
  ref_decode = (ref_decode_t)dlsym(h, "decode_biseq");
  ref_ovlp_match = (ref_ovlp_match_t)dlsym(h, "ovlp_match");
}

struct Packed {
  ReadTable rt;
  std::vector<uint64_t> w, wrc;  // forward image and reverse-complement image (same word offsets)
  std::vector<uint32_t> nm;
  std::vector<uint64_t> woff;  // per row
  std::vector<uint8_t> has_n;
  const uint8_t *seqdb = nullptr;
  size_t seqdb_size = 0;
};

static void load_packed(const char *prefix, Packed *P) {
  std::string idx = std::string(prefix) + ".idx", db = std::string(prefix) + ".seqdb";
  if (!load_read_table(idx.c_str(), &P->rt)) die("cannot open %s", idx.c_str());
  int fd = open(db.c_str(), O_RDONLY);
  if (fd < 0) die("cannot open %s", db.c_str());
  struct stat sb; fstat(fd, &sb);
  P->seqdb = (const uint8_t *)mmap(0, sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
  P->seqdb_size = sb.st_size;
  size_t n = P->rt.n();
  uint64_t words = 2;
  P->woff.resize(n); P->has_n.resize(n);
  for (size_t i = 0; i < n; i++) { P->woff[i] = words; words += (P->rt.len[i] + 31) / 32; }
  words += 2;
  P->w.assign(words, 0); P->nm.assign(words, 0);
  for (size_t i = 0; i < n; i++) {
    const uint8_t *s = P->seqdb + P->rt.off[i];
    int hn = 0;
    for (uint32_t p = 0; p < P->rt.len[i]; p++) {
      uint8_t nib = s[p] & 0xF;
      uint64_t c = 0; int isn = 0;
      switch (nib) { case 1: c = 0; break; case 2: c = 1; break; case 4: c = 2; break; case 8: c = 3; break; default: isn = 1; }
      P->w[P->woff[i] + p / 32] |= c << (2 * (p & 31));
      if (isn) { P->nm[P->woff[i] + p / 32] |= 1u << (p & 31); hn = 1; }
    }
    P->has_n[i] = hn;
  }
  P->wrc.assign(words, 0);  // what k_make_rc builds on the device: rc base p = 3 - forward base (len-1-p)
  for (size_t i = 0; i < n; i++) {
    uint32_t L = P->rt.len[i];
    for (uint32_t p = 0; p < L; p++) {
      uint32_t f = L - 1 - p;
      uint64_t c = 3 - ((P->w[P->woff[i] + f / 32] >> (2 * (f & 31))) & 3);
      P->wrc[P->woff[i] + p / 32] |= c << (2 * (p & 31));
    }
  }
}
struct LeanCase { size_t r0; uint32_t start0; int s0; size_t r1; int s1; match_t want; };
static std::vector<LeanCase> g_lean_cases;
static bool lean_match(const Packed &P, size_t r0, uint32_t start0, int s0, size_t r1, int s1, int bw, match_t *m) {
  if (P.has_n[r0] || P.has_n[r1]) return false;  // reads with N take ovlp_match_flat in the product too
  std::vector<int> V(2 * (bw + 2));
  ovlp_match_lean((s0 ? P.wrc.data() : P.w.data()) + P.woff[r0], start0, (int)(P.rt.len[r0] - start0),
                  (s1 ? P.wrc.data() : P.w.data()) + P.woff[r1], 0, (int)P.rt.len[r1], bw, V.data(), bw + 2, m);
  return true;
}

static void sim_reduce(const std::vector<mm128> &in, std::vector<mm128> &out, uint32_t rs) {
  // per-element formulation (what the kernel does): element e with in-read offset o>=rs-1 emits its pick if it
  // differs from the pick of e-1 (or if e-1 had no full window / belongs to another read)
  size_t n = in.size();
  size_t start = 0;
  for (size_t e = 0; e < n; e++) {
    if (e == 0 || (in[e].y >> 32) != (in[e - 1].y >> 32)) start = e;
    uint32_t o = (uint32_t)(e - start);
    if (o + 1 < rs) continue;
    uint32_t t = reduce_pick(in.data() + start, o, rs);
    bool emit = true;
    if (o >= rs) {
      uint32_t tp = reduce_pick(in.data() + start, o - 1, rs);
      if (in[start + tp].y == in[start + t].y) emit = false;
    }
    if (emit) out.push_back(in[start + t]);
  }
}


// ---- CPU execution of the tiled sketch kernel's phases (peregrine_b200/csrc/sketch_tile.cuh): one "CTA" at a time,
// the 256 threads of each phase run sequentially, block scans are plain loops.  Returns flags != 0 if the read must
// take the exact automaton instead.
template <class HT>
static uint32_t sketch_tiled_host(const Packed &P, size_t row, int wsz, int k, std::vector<mm128> &out) {
  const int len = (int)P.rt.len[row];
  if (P.has_n[row]) return SK_FLAG_N;
  if (len < sk_min_len(wsz, k)) return SK_FLAG_SHORT;
  const int H = sk_halo(wsz), TILE = sk_tile_len(wsz);
  const int n_tiles = (len + TILE - 1) / TILE;
  std::vector<unsigned char> smem(sk_smem_bytes<HT>() + 16);
  SkTile<HT> sh;
  sk_tile_layout<HT>(sh, smem.data(), wsz);
  std::vector<mm128> rec;
  for (int j = 0; j < n_tiles; j++) {
    SkParams p;
    p.w = P.w.data(); p.word_off = P.woff[row]; p.len = len; p.rid = P.rt.rid[row]; p.wsz = wsz; p.k = k;
    p.r0 = j * TILE - H; p.first_tile = j == 0;
    static HT hv[SK_THREADS][SK_G];
    static uint16_t ps[SK_THREADS][SK_G];
    uint32_t mask[SK_THREADS], base[SK_THREADS];
    memset(smem.data(), 0xAB, smem.size());  // nothing may depend on stale shared memory
    for (int q = 0; q < 4; q++) sh.ctr[q] = 0;
    uint32_t tot = 0;
    for (int t = 0; t < SK_THREADS; t++) {
      uint32_t np, hs;
      sk_phase1<HT>(t, p, H, hv[t], ps[t], &mask[t], &np, &hs);
      sh.ctr[SK_N_PAL] += np; sh.ctr[SK_N_HALO] += hs;
      base[t] = tot; tot += (uint32_t)sk_popc(mask[t]);
    }
    sh.ctr[SK_N_SLOTS] = tot;
    if (sh.ctr[SK_N_PAL] > SK_PALPAD) return SK_FLAG_PAL;
    for (int t = 0; t < SK_THREADS; t++) sk_phase2_write<HT>(sh, hv[t], ps[t], mask[t], base[t]);
    const int n_blocks = ((int)tot + sh.B - 1) / sh.B;
    for (int b = 0; b < n_blocks; b++) sk_phase3_suffix<HT>(b, sh);
    for (int b = 0; b < n_blocks; b++) sk_phase3_prefix<HT>(b, sh);
    int s_eval, s_emit, s_first_full;
    sk_ranges(p, sh.ctr[SK_N_HALO], &s_eval, &s_emit, &s_first_full);
    uint32_t tie = 0, total = 0, em[SK_THREADS], cn[SK_THREADS];
    for (int t = 0; t < SK_THREADS; t++) { uint32_t ti; cn[t] = sk_phase4<HT>(t, sh, wsz, s_eval, s_emit, s_first_full, &em[t], &ti); tie |= ti; total += cn[t]; }
    if (tie) return SK_FLAG_TIE;
    if (total > SK_CAP) return SK_FLAG_OVERFLOW;
    size_t at = rec.size();
    rec.resize(at + total);
    uint32_t o = 0;
    for (int t = 0; t < SK_THREADS; t++) { if (cn[t]) sk_phase5_write<HT>(t, sh, p, em[t], rec.data() + at + o); o += cn[t]; }
  }
  out.insert(out.end(), rec.begin(), rec.end());
  return 0;
}

// Host MODEL of what the strip sketch kernel (peregrine_b200/csrc/sketch_strip.cuh) decides, used to check the splice argument of
// its fallback on the CPU: per full window the rightmost arg-min over the last w SLOTS (palindromic k-mers are no slots), a
// record when the minimum changes, and the same "cannot decide" conditions as the kernel (a minimum that occurs twice, a new
// element equal to the previous window's minimum, two palindromic k-mers in reach, a palindromic k-mer before the first full
// window), recorded per 512-position strip.  Returns false when the kernel would hand over the whole read.
static bool sketch_fast_model(const char *seq, uint32_t len, int w, int k, uint32_t rid, std::vector<mm128> &recs, uint64_t *bad_out) {
  const uint64_t MAXV = ~0ULL, mask = (1ULL << 2 * k) - 1, shift1 = 2 * (uint64_t)(k - 1);
  *bad_out = 0;
  if ((int)len < sk_min_len(w, k)) return false;
  struct Slot { uint64_t v; uint32_t pos, strand; };
  std::vector<Slot> slots;
  std::vector<uint32_t> pals;
  uint64_t kmer0 = 0, kmer1 = 0, bad = 0;
  int l = 0;
  const int e_ff = w + k - 2;
  for (uint32_t p = 0; p < len; p++) {
    int c;
    switch (seq[p]) { case 'A': case 'a': c = 0; break; case 'C': case 'c': c = 1; break; case 'G': case 'g': c = 2; break; case 'T': case 't': c = 3; break; default: return false; }
    kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
    kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
    if (kmer0 == kmer1) {
      if (p + 1 >= (uint32_t)k) {  // (the kernel looks at complete k-mers only)
        if ((int)p <= e_ff + 1) bad |= 1;
        if (pals.size() >= 8) return false;
        pals.push_back(p);
      }
      continue;
    }
    const uint32_t z = kmer0 < kmer1 ? 0 : 1;
    ++l;
    slots.push_back(Slot{l >= k ? hash64(z ? kmer1 : kmer0, mask) : MAXV, p, z});
  }
  const uint64_t cap = (6ull * len) / (uint32_t)(w + 1) + 64;
  uint64_t prev = MAXV;
  bool have_prev = false;
  std::vector<uint32_t> per16((len >> 4) + 2, 0);
  for (size_t s = 0; s < slots.size(); s++) {
    const int e = (int)slots[s].pos;
    if (e < e_ff - 1 || (int)s < w - 1) continue;  // evaluated from the window before the first full one (its tie check only)
    uint64_t v = MAXV;
    uint32_t arg = 0, strand = 0, cnt = 0;
    for (size_t t = s + 1 - (size_t)w; t <= s; t++) {
      if (slots[t].v < v) { v = slots[t].v; cnt = 0; }
      if (slots[t].v == v) { arg = slots[t].pos; strand = slots[t].strand; cnt++; }
    }
    bool undecided = cnt > 1 && v != MAXV;
    if (have_prev && slots[s].v == prev && prev != MAXV) undecided = true;
    uint32_t cpal = 0;
    for (uint32_t q : pals) cpal += (q >= slots[s + 1 - (size_t)w].pos && q <= (uint32_t)e);
    if (cpal > 1) undecided = true;
    if (undecided) { if ((e >> 9) >= 64) return false; bad |= 1ull << (e >> 9); }
    const bool full = (int)s >= w + k - 2;  // l >= w+k-1
    if (full && (!have_prev || v != prev || (int)s == w + k - 2)) {
      recs.push_back(mm128{v << 8 | (uint64_t)k, (uint64_t)rid << 32 | (uint64_t)(arg << 1 | strand)});
      if (++per16[e >> 4] > 8) return false;  // the kernel stages at most 8 records per lane and strip
    }
    prev = v;
    have_prev = true;
  }
  if (recs.size() > cap) return false;
  *bad_out = bad;
  return true;
}

static int cmd_sketch(int argc, char **argv) {
  if (argc < 5) return 1;
  for (int B = 9; B <= 129; B++)
    for (int q = 0; q < 8192; q++)
      if (sk_div(q, sk_div_magic(B)) != q / B) { fprintf(stderr, "sk_div(%d, %d) is wrong\n", q, B); return 4; }
  static_assert(SK_R <= 8192, "sk_div is exact below 2^13 only");
  Packed P; load_packed(argv[2], &P);
  int w = atoi(argv[3]), k = atoi(argv[4]);
  int rs = argc > 5 ? atoi(argv[5]) : 6;
  std::vector<uint64_t> rx(256); std::vector<uint32_t> rp(256);
  size_t bad = 0, total = 0, bad_tiled = 0, n_fallback = 0, bad_seg = 0, n_seg_retry = 0; uint32_t fallback_flags = 0;
  size_t n_model = 0, n_model_partial = 0, n_model_pieces = 0, bad_splice = 0;
  // PGB_SPLICE_REACH: positions before a bad strip that are redone as well (default: the product's w + 8); a value that is too small
  // must make the splice check fail (the test uses this to show that the check can fail)
  const int splice_reach = getenv("PGB_SPLICE_REACH") ? atoi(getenv("PGB_SPLICE_REACH")) : -1;
  std::vector<mm128> all_mine;
  mm128_v all_ref = {0, 0, 0};
  for (size_t i = 0; i < P.rt.n(); i++) {
    uint32_t len = P.rt.len[i];
    if (len == 0) continue;
    std::vector<char> seq(len + 1);
    ref_decode((uint8_t *)P.seqdb + P.rt.off[i], seq.data(), len, 0);
    size_t n0 = all_ref.n;
    ref_mm_sketch(NULL, seq.data(), len, w, k, P.rt.rid[i], 0, &all_ref);
    size_t m0 = all_mine.size();
    sketch_exact(P.w.data(), P.has_n[i] ? P.nm.data() : nullptr, P.woff[i], (int)len, w, k, P.rt.rid[i], rx.data(), rp.data(),
                 [&](uint64_t x, uint64_t y) { all_mine.push_back(mm128{x, y}); });
    size_t nr = all_ref.n - n0, nmine = all_mine.size() - m0;
    total += nr;
    {  // segment-parallel exact automaton: cut the read into segments, replay each independently, concatenate
      for (int seg : {257, 1024}) {
        std::vector<mm128> segout;
        for (int lo = 0; lo < (int)len; lo += seg) {
          int hi = std::min<int>(lo + seg, (int)len);
          int st = std::max(0, lo - sketch_warmup_len(w, k));
          auto em = [&](uint64_t x, uint64_t y) { segout.push_back(mm128{x, y}); };
          size_t mark = segout.size();
          if (!sketch_exact_range(P.w.data(), P.has_n[i] ? P.nm.data() : nullptr, P.woff[i], (int)len, w, k, P.rt.rid[i], st, lo, hi, rx.data(), rp.data(), em)) {
            segout.resize(mark); n_seg_retry++;
            sketch_exact_range(P.w.data(), P.has_n[i] ? P.nm.data() : nullptr, P.woff[i], (int)len, w, k, P.rt.rid[i], 0, lo, hi, rx.data(), rp.data(), em);
          }
        }
        if (segout.size() != nr || memcmp(all_ref.a + n0, segout.data(), nr * 16) != 0) {
          if (bad_seg < 5) fprintf(stderr, "SEGMENT sketch mismatch read row %zu len %u seg %d: ref %zu seg %zu\n", i, len, seg, nr, segout.size());
          bad_seg++;
        }
      }
    }
    if (w >= SK_MINW) {
      std::vector<mm128> tiled;
      uint32_t fl = (k <= 16) ? sketch_tiled_host<uint32_t>(P, i, w, k, tiled) : sketch_tiled_host<uint64_t>(P, i, w, k, tiled);
      if (fl) { n_fallback++; fallback_flags |= fl; }
      else if (tiled.size() != nr || memcmp(all_ref.a + n0, tiled.data(), nr * 16) != 0) {
        if (bad_tiled < 5) {
          fprintf(stderr, "TILED sketch mismatch read row %zu rid %u len %u: ref %zu tiled %zu\n", i, P.rt.rid[i], len, nr, tiled.size());
          for (size_t q = 0; q < std::min(nr, tiled.size()); q++) if (memcmp(&all_ref.a[n0 + q], &tiled[q], 16)) { fprintf(stderr, "  first diff at %zu: ref pos %u tiled pos %u\n", q, (uint32_t)(all_ref.a[n0+q].y & 0xFFFFFFFF) >> 1, (uint32_t)(tiled[q].y & 0xFFFFFFFF) >> 1); break; }
        }
        bad_tiled++;
      }
    }
    if (w >= SK_MINW && !P.has_n[i]) {  // strip-kernel model + per-strip fallback, spliced by position (sketch_strip.cuh, build_sketch_pieces)
      std::vector<mm128> fast, spliced;
      uint64_t badmask = 0;
      if (sketch_fast_model(seq.data(), len, w, k, P.rt.rid[i], fast, &badmask)) {
        n_model++;
        if (badmask) n_model_partial++;
        std::vector<SketchPiece> pieces;
        if (badmask) build_sketch_pieces(len, badmask, 512, splice_reach >= 0 ? (uint32_t)splice_reach : (uint32_t)w + 8, 256, pieces);
        else pieces.push_back(SketchPiece{0, len, 1});
        for (const SketchPiece &pc : pieces) {
          if (pc.kind) {
            for (const mm128 &m : fast) { const uint32_t pos = (uint32_t)m.y >> 1; if (pos >= pc.lo && pos < pc.hi) spliced.push_back(m); }
          } else {
            n_model_pieces++;
            int st = std::max(0, (int)pc.lo - sketch_warmup_len(w, k));
            auto em = [&](uint64_t x, uint64_t y) { spliced.push_back(mm128{x, y}); };
            size_t mark = spliced.size();
            if (!sketch_exact_range(P.w.data(), nullptr, P.woff[i], (int)len, w, k, P.rt.rid[i], st, (int)pc.lo, (int)pc.hi, rx.data(), rp.data(), em)) {
              spliced.resize(mark);
              sketch_exact_range(P.w.data(), nullptr, P.woff[i], (int)len, w, k, P.rt.rid[i], 0, (int)pc.lo, (int)pc.hi, rx.data(), rp.data(), em);
            }
          }
        }
        if (spliced.size() != nr || memcmp(all_ref.a + n0, spliced.data(), nr * 16) != 0) {
          if (bad_splice < 5) {
            fprintf(stderr, "SPLICE mismatch read row %zu rid %u len %u bad 0x%llx: ref %zu spliced %zu\n", i, P.rt.rid[i], len, (unsigned long long)badmask, nr,
                    spliced.size());
            for (size_t q = 0; q < std::min(nr, spliced.size()); q++)
              if (memcmp(&all_ref.a[n0 + q], &spliced[q], 16)) {
                fprintf(stderr, "  first diff at %zu: ref pos %u spliced pos %u\n", q, (uint32_t)(all_ref.a[n0 + q].y & 0xFFFFFFFF) >> 1, (uint32_t)(spliced[q].y & 0xFFFFFFFF) >> 1);
                break;
              }
          }
          bad_splice++;
        }
      }
    }
    if (nr != nmine || memcmp(all_ref.a + n0, all_mine.data() + m0, nr * 16) != 0) {
      if (bad < 5) fprintf(stderr, "sketch mismatch read row %zu rid %u len %u: ref %zu mine %zu\n", i, P.rt.rid[i], len, nr, nmine);
      bad++;
    }
  }
  printf("sketch: reads=%zu L0=%zu mismatching_reads=%zu ; tiled kernel: mismatching=%zu fallback_reads=%zu (flags 0x%x) ; segmented exact: mismatching=%zu retries=%zu\n", P.rt.n(), total, bad, bad_tiled, n_fallback, fallback_flags, bad_seg, n_seg_retry);
  printf("strip model + per-strip fallback: reads=%zu of them in part=%zu automaton pieces=%zu mismatching=%zu\n", n_model, n_model_partial, n_model_pieces, bad_splice);
  // reduce twice
  mm128_v r1 = {0, 0, 0}, r2 = {0, 0, 0};
  ref_mm_reduce(&all_ref, &r1, (uint8_t)rs);
  ref_mm_reduce(&r1, &r2, (uint8_t)rs);
  std::vector<mm128> m1, m2;
  sim_reduce(all_mine, m1, rs);
  sim_reduce(m1, m2, rs);
  bool ok1 = r1.n == m1.size() && (r1.n == 0 || memcmp(r1.a, m1.data(), r1.n * 16) == 0);
  bool ok2 = r2.n == m2.size() && (r2.n == 0 || memcmp(r2.a, m2.data(), r2.n * 16) == 0);
  printf("reduce: L1 ref=%zu mine=%zu %s ; L2 ref=%zu mine=%zu %s\n", r1.n, m1.size(), ok1 ? "OK" : "MISMATCH", r2.n, m2.size(),
         ok2 ? "OK" : "MISMATCH");
  return (bad || bad_tiled || bad_seg || bad_splice || !ok1 || !ok2) ? 3 : 0;
}

static uint64_t rng_state = 0x12345;
static inline uint64_t rnd() {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

static bool same_match(const match_t &a, const match_t &b) { return memcmp(&a, &b, sizeof(match_t)) == 0; }

static int one_match(const Packed &P, size_t r0, uint32_t start0, int s0, size_t r1, int s1, int bw, size_t *bad) {
  uint32_t rlen0 = P.rt.len[r0], rlen1 = P.rt.len[r1];
  if (start0 >= rlen0) start0 = 0;
  int qlen = (int)(rlen0 - start0), tlen = (int)rlen1;
  match_t *ref = ref_ovlp_match((uint8_t *)P.seqdb + P.rt.off[r0] + start0, qlen, (uint8_t)s0, (uint8_t *)P.seqdb + P.rt.off[r1], tlen,
                                (uint8_t)s1, bw);
  std::vector<int> Va(bw + 8), Vb(bw + 8);
  SeqView q = make_view(P.w.data(), P.nm.data(), P.woff[r0], rlen0, start0, s0, P.has_n[r0]);
  SeqView t = make_view(P.w.data(), P.nm.data(), P.woff[r1], rlen1, 0, s1, P.has_n[r1]);
  match_t mine; int err = 0;
  ovlp_match_core(q, qlen, t, tlen, bw, Va.data(), Vb.data(), bw + 8, &mine, &err);
  match_t flat; int err2 = 0;
  std::vector<int> Vf(2 * (bw + 8));
  ovlp_match_flat(q, qlen, t, tlen, bw, Vf.data(), bw + 8, &flat, &err2);
  int ok = same_match(*ref, mine) && err == 0 && same_match(*ref, flat) && err2 == 0;
  match_t lean;
  const bool has_lean = lean_match(P, r0, start0, s0, r1, s1, bw, &lean);
  if (has_lean) g_lean_cases.push_back(LeanCase{r0, start0, s0, r1, s1, *ref});
  if (has_lean && !same_match(*ref, lean)) {
    ok = 0;
    if (*bad < 8) fprintf(stderr, "LEAN differs: {%d %d %d %d %d %d %d %d}\n", lean.m_size, lean.dist, lean.q_bgn, lean.q_end, lean.t_bgn, lean.t_end, lean.t_m_end, lean.q_m_end);
  }
  if (!same_match(*ref, flat) && *bad < 8) fprintf(stderr, "FLAT differs: {%d %d %d %d %d %d %d %d}\n", flat.m_size, flat.dist, flat.q_bgn, flat.q_end, flat.t_bgn, flat.t_end, flat.t_m_end, flat.q_m_end);
  if (!ok) {
    if (*bad < 8)
      fprintf(stderr, "match mismatch rows %zu(+%u,s%d) %zu(s%d) bw=%d err=%d: ref{%d %d %d %d %d %d %d %d} mine{%d %d %d %d %d %d %d %d}\n", r0,
              start0, s0, r1, s1, bw, err, ref->m_size, ref->dist, ref->q_bgn, ref->q_end, ref->t_bgn, ref->t_end, ref->t_m_end, ref->q_m_end,
              mine.m_size, mine.dist, mine.q_bgn, mine.q_end, mine.t_bgn, mine.t_end, mine.t_m_end, mine.q_m_end);
    (*bad)++;
  }
  free(ref);
  return ok;
}

static int cmd_match(int argc, char **argv) {
  if (argc < 5) return 1;
  Packed P; load_packed(argv[2], &P);
  size_t npairs = strtoull(argv[3], 0, 10);
  int bw = atoi(argv[4]);
  size_t bad = 0, n = P.rt.n();
  for (size_t it = 0; it < npairs; it++) {
    size_t r0 = rnd() % n, r1 = rnd() % n;
    uint32_t st = (uint32_t)(rnd() % (P.rt.len[r0] ? P.rt.len[r0] : 1));
    if (it % 4 == 0) r1 = r0;  // self alignment at an offset / strand
    if (it % 4 == 1) st = 0;
    one_match(P, r0, st, (int)(rnd() & 1), r1, (int)(rnd() & 1), bw, &bad);
  }
  const size_t stream_bad = 0, stream_n = 0;  // (the streaming form of ovlp_match_lean was measured and dropped in round 2)
  printf("match: pairs=%zu mismatches=%zu stream_results=%zu stream_mismatches=%zu\n", npairs, bad, stream_n, stream_bad);
  return (bad || stream_bad) ? 3 : 0;
}

// ------------------------------------------------------------------------------------------------------------ overlap
struct PairRec { uint64_t k0, k1, y0, y1; uint32_t seq; uint8_t dir; };
struct Bucket { uint64_t k0, k1; uint32_t first_seq; std::vector<uint32_t> recs; uint32_t last_seq = 0; };

struct HostCtx {
  const Packed *P;
  std::unordered_map<uint64_t, uint64_t> *e_old, *e_new;
  std::unordered_map<uint64_t, match_t> *aln;  // key rank<<32 | i<<16 | j
  std::vector<std::array<uint32_t, 8>> *requests;
  std::vector<ovlp_rec> *out;
  uint32_t rank;
  bool request_enabled;
  uint32_t rlen(uint32_t rid) const { return P->rt.len[P->rt.by_rid[rid]]; }
  uint64_t pair_old(uint64_t p) const { auto it = e_old->find(p); return it == e_old->end() ? ~0ULL : it->second; }
  bool strict = getenv("PGB_SIM_STRICT") != nullptr;  // pure Jacobi: other buckets' writes of this pass stay invisible
  uint64_t pair_new(uint64_t p) const {
    auto it = e_new->find(p);
    if (it == e_new->end()) return ~0ULL;
    if (strict && (uint32_t)(it->second >> 2) != rank) return ~0ULL;
    return it->second;
  }
  void pair_get(uint64_t p, uint64_t *vo, uint64_t *vn) const { *vo = pair_old(p); *vn = pair_new(p); }
  void pair_set(uint64_t p, uint64_t v) { auto it = e_new->find(p); if (it == e_new->end() || v < it->second) (*e_new)[p] = v; }
  bool aln_get(uint32_t i, uint32_t j, match_t *m) const {
    auto it = aln->find(((uint64_t)rank << 32) | (i << 16) | j);
    if (it == aln->end()) return false;
    *m = it->second; return true;
  }
  void aln_request(uint32_t i, uint32_t j, uint32_t rid0, uint32_t start0, uint32_t s0, uint32_t rid1, uint32_t s1) {
    if (request_enabled) requests->push_back({rank, i, j, rid0, start0, s0, rid1, s1});
  }
  void emit(uint32_t, const ovlp_rec &o) { out->push_back(o); }
};

static match_t do_align(const Packed &P, uint32_t rid0, uint32_t start0, uint32_t s0, uint32_t rid1, uint32_t s1, int bw) {
  size_t r0 = P.rt.by_rid[rid0], r1 = P.rt.by_rid[rid1];
  uint32_t rlen0 = P.rt.len[r0], rlen1 = P.rt.len[r1];
  static std::vector<int> Va, Vb;
  Va.resize(bw + 8); Vb.resize(bw + 8);
  SeqView q = make_view(P.w.data(), P.nm.data(), P.woff[r0], rlen0, start0, s0, P.has_n[r0]);
  SeqView t = make_view(P.w.data(), P.nm.data(), P.woff[r1], rlen1, 0, s1, P.has_n[r1]);
  match_t m; int err = 0;
  ovlp_match_core(q, (int)(rlen0 - start0), t, (int)rlen1, bw, Va.data(), Vb.data(), bw + 8, &m, &err);
  if (err) fprintf(stderr, "ovlp_match_core err=%d\n", err);
  {  // the flattened (SIMT-friendly) form must give the same answer on every alignment the pipeline runs
    static std::vector<int> Vf; Vf.resize(2 * (bw + 8));
    match_t f; int e2 = 0;
    ovlp_match_flat(q, (int)(rlen0 - start0), t, (int)rlen1, bw, Vf.data(), bw + 8, &f, &e2);
    if (e2 || memcmp(&f, &m, sizeof f)) { fprintf(stderr, "ovlp_match_flat disagrees with ovlp_match_core (rid %u vs %u)\n", rid0, rid1); exit(5); }
    match_t l;
    if (lean_match(P, r0, start0, (int)s0, r1, (int)s1, bw, &l) && memcmp(&l, &m, sizeof l)) {
      fprintf(stderr, "ovlp_match_lean disagrees with ovlp_match_core (rid %u vs %u)\n", rid0, rid1); exit(5);
    }
  }
  return m;
}

static int cmd_overlap(int argc, char **argv) {
  if (argc < 7) return 1;
  Packed P; load_packed(argv[2], &P);
  std::string l2 = argv[3];
  uint32_t T = atoi(argv[4]), c = atoi(argv[5]);
  const char *ref_file = argv[6];
  bool jacobi = argc > 7 && !strcmp(argv[7], "jacobi");
  int dry_passes = argc > 8 ? atoi(argv[8]) : 3;
  uint32_t mc_lower = 2, mc_upper = 240, ovlp_upper = 120, bestn = 4; int bw = 100;
  std::vector<mm128> mm;
  for (auto &fn : glob_sorted(l2 + "-[0-9]*-of-[0-9]*.dat")) read_mmlist_file(fn.c_str(), &mm);
  std::unordered_map<uint64_t, uint32_t> cnt;
  for (auto &fn : glob_sorted(l2 + "-MC-[0-9]*-of-[0-9]*.dat")) {
    std::vector<mc_rec> v; read_mc_file(fn.c_str(), &v);
    for (auto &r : v) cnt[r.mer] += r.count;
  }
  printf("L2 mmers=%zu distinct=%zu\n", mm.size(), cnt.size());
  // kept set (src/shmr_utils.c:310-328)
  size_t n = mm.size(), s = 0;
  for (; s < n; s++) { uint32_t mc = cnt[mm[s].x >> 8]; if (mc >= mc_lower && mc < mc_upper) break; }
  std::vector<uint32_t> kept;
  if (s < n) kept.push_back((uint32_t)s);
  for (size_t i = s + 1; i < n; i++) { uint32_t mc = cnt[mm[i].x >> 8]; if (mc < mc_lower || mc > mc_upper) continue; kept.push_back((uint32_t)i); }
  // pair records
  std::vector<PairRec> recs;
  for (size_t t = 0; t + 1 < kept.size(); t++) {
    const mm128 &m0 = mm[kept[t]], &m1 = mm[kept[t + 1]];
    if ((m0.y >> 32) != (m1.y >> 32)) continue;
    if (!pair_far_enough(m0.y, m1.y)) continue;
    if ((m0.x >> 8) % T == c % T) recs.push_back({m0.x, m1.x, m0.y, m1.y, (uint32_t)(2 * t), 0});
    if ((m1.x >> 8) % T == c % T) {
      uint32_t rl0 = P.rt.len[P.rt.by_rid[m1.y >> 32]], rl1 = P.rt.len[P.rt.by_rid[m0.y >> 32]];
      recs.push_back({m1.x, m0.x, rev_y(m1.y, m1.x, rl0), rev_y(m0.y, m0.x, rl1), (uint32_t)(2 * t + 1), 1});
    }
  }
  printf("kept=%zu pair records=%zu\n", kept.size(), recs.size());
  // group into buckets (stand-in for the device hash tables), first-seq order
  std::map<std::pair<uint64_t, uint64_t>, uint32_t> bidx;
  std::vector<Bucket> buckets;
  for (uint32_t r = 0; r < recs.size(); r++) {
    auto key = std::make_pair(recs[r].k0, recs[r].k1);
    auto it = bidx.find(key);
    if (it == bidx.end()) { bidx[key] = (uint32_t)buckets.size(); buckets.push_back(Bucket{recs[r].k0, recs[r].k1, recs[r].seq, {}, 0}); it = bidx.find(key); }
    buckets[it->second].recs.push_back(r);
    buckets[it->second].last_seq = recs[r].seq;
  }
  // khash visiting order: outer keys by first insertion, inner keys per outer by first insertion
  // (buckets vector is already in first-seq order because recs are in seq order)
  KhashEmu outer;
  std::unordered_map<uint64_t, uint32_t> outer_id;
  std::vector<std::vector<uint32_t>> inner_lists;
  std::vector<uint32_t> o_last;
  uint32_t newest_outer_seq = 0, last_seq_all = 0;
  for (uint32_t b = 0; b < buckets.size(); b++) {
    last_seq_all = std::max(last_seq_all, buckets[b].last_seq);
    auto it = outer_id.find(buckets[b].k0);
    if (it == outer_id.end()) {
      uint32_t id = (uint32_t)inner_lists.size();
      outer_id[buckets[b].k0] = id;
      outer.put_new(buckets[b].k0, id);
      newest_outer_seq = buckets[b].first_seq;
      inner_lists.emplace_back();
      inner_lists[id].push_back(b);
      o_last.push_back(buckets[b].last_seq);
    } else { inner_lists[it->second].push_back(b); o_last[it->second] = std::max(o_last[it->second], buckets[b].last_seq); }
  }
  if (last_seq_all > newest_outer_seq) outer.touch_existing();
  size_t n_khs_checked = 0;
  std::vector<uint32_t> visit;  // eligible buckets in visiting order
  outer.for_each_in_slot_order([&](uint64_t, uint32_t id) {
    KhashEmu inner;
    uint32_t newest_inner_seq = 0;
    for (uint32_t b : inner_lists[id]) { inner.put_new(buckets[b].k1, b); newest_inner_seq = buckets[b].first_seq; }
    if (o_last[id] > newest_inner_seq) inner.touch_existing();
    if (inner_lists[id].size() <= KHS_MAX_KEYS) {  // the fixed-capacity device model must agree with the unbounded one
      uint64_t kk[KHS_MAX_KEYS]; uint8_t rk[KHS_MAX_KEYS];
      for (size_t q = 0; q < inner_lists[id].size(); q++) kk[q] = buckets[inner_lists[id][q]].k1;
      khs_order(kk, (uint32_t)inner_lists[id].size(), o_last[id] > newest_inner_seq, rk);
      uint32_t r = 0; bool okk = true;
      inner.for_each_in_slot_order([&](uint64_t, uint32_t b) { size_t q = 0; while (inner_lists[id][q] != b) q++; if (rk[q] != r) okk = false; r++; });
      if (!okk) { fprintf(stderr, "khs_order disagrees with KhashEmu for outer id %u (%zu keys)\n", id, inner_lists[id].size()); exit(4); }
      n_khs_checked++;
    }
    inner.for_each_in_slot_order([&](uint64_t, uint32_t b) {
      size_t nn = buckets[b].recs.size();
      if (nn <= 2 || nn > ovlp_upper) return;
      visit.push_back(b);
    });
  });
  printf("buckets=%zu outer=%zu eligible=%zu khs_checked=%zu\n", buckets.size(), inner_lists.size(), visit.size(), n_khs_checked);
  if (getenv("PGB_SIM_VISIT")) { FILE *f = fopen(getenv("PGB_SIM_VISIT"), "w"); for (uint32_t b : visit) fprintf(f, "VISIT %lu %lu %zu\n", (unsigned long)buckets[b].k0, (unsigned long)buckets[b].k1, buckets[b].recs.size()); fclose(f); }
  // sorted record arrays per eligible bucket: stable, descending position (glibc qsort with mp128_comp)
  std::vector<std::vector<uint64_t>> by0(visit.size());
  std::vector<std::vector<uint8_t>> bdir(visit.size());
  size_t cand = 0;
  for (size_t r = 0; r < visit.size(); r++) {
    std::vector<uint32_t> v = buckets[visit[r]].recs;
    std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) {
      return ((recs[a].y0 & 0xFFFFFFFFULL) >> 1) > ((recs[b].y0 & 0xFFFFFFFFULL) >> 1);
    });
    for (uint32_t a : v) { by0[r].push_back(recs[a].y0); bdir[r].push_back(recs[a].dir); }
    cand += v.size() * (v.size() - 1) / 2;
  }
  printf("candidate (i,j) pairs=%zu\n", cand);
  // reference output
  std::vector<ovlp_rec> ref;
  {
    FILE *f = fopen(ref_file, "rb");
    if (!f) die("cannot open %s", ref_file);
    ovlp_rec o;
    while (fread(&o, sizeof o, 1, f) == 1) { o.pad0 = 0; o.pad1 = 0; ref.push_back(o); }
    fclose(f);
  }
  std::unordered_map<uint64_t, uint64_t> eA, eB;
  std::unordered_map<uint64_t, match_t> aln;
  std::vector<std::array<uint32_t, 8>> requests;
  std::vector<ovlp_rec> out;
  std::vector<uint8_t> contained(65536);
  HostCtx ctx{&P, &eA, &eB, &aln, &requests, &out, 0, true};
  if (ctx.strict) printf("strict Jacobi simulation\n");
  size_t n_align = 0;
  if (!jacobi) {
    // sequential (Gauss-Seidel with immediate alignment): exactness check of the restated control path
    for (uint32_t r = 0; r < visit.size(); r++) {
      // iterate the bucket until it asks for nothing new
      for (;;) {
        ctx.rank = r; requests.clear(); out.clear();
        // e_new must not keep entries of this rank from the previous attempt
        std::unordered_map<uint64_t, uint64_t> scratch;
        ctx.e_old = &eA; ctx.e_new = &scratch;
        uint32_t unk = 0;
        replay_bucket(ctx, r, by0[r].data(), bdir[r].data(), (uint32_t)by0[r].size(), contained.data(), bestn, true, &unk);
        if (requests.empty()) { for (auto &kv : scratch) eA[kv.first] = kv.second; break; }
        for (auto &q : requests) { aln[((uint64_t)q[0] << 32) | (q[1] << 16) | q[2]] = do_align(P, q[3], q[4], q[5], q[6], q[7], bw); n_align++; }
      }
      static std::vector<ovlp_rec> all;
      all.insert(all.end(), out.begin(), out.end());
      if (r + 1 == visit.size()) out = all;
    }
  } else {
    int pass = 0;
    for (;; pass++) {
      bool dry = pass < dry_passes;
      ctx.request_enabled = !dry;
      ctx.e_old = &eA; ctx.e_new = &eB; eB.clear(); requests.clear(); out.clear();
      size_t unk_total = 0;
      for (uint32_t r = 0; r < visit.size(); r++) {
        ctx.rank = r; uint32_t unk = 0;
        replay_bucket(ctx, r, by0[r].data(), bdir[r].data(), (uint32_t)by0[r].size(), contained.data(), bestn, true, &unk);
        unk_total += unk;
      }
      size_t diffs = 0;
      for (auto &kv : eB) { auto it = eA.find(kv.first); if (it == eA.end() || it->second != kv.second) diffs++; }
      for (auto &kv : eA) if (eB.find(kv.first) == eB.end()) diffs++;
      for (auto &q : requests) { aln[((uint64_t)q[0] << 32) | (q[1] << 16) | q[2]] = do_align(P, q[3], q[4], q[5], q[6], q[7], bw); n_align++; }
      printf("pass %d %s: records=%zu unknown=%zu new_requests=%zu table_diffs=%zu total_aligned=%zu\n", pass, dry ? "dry" : "wet", out.size(),
             unk_total, requests.size(), diffs, n_align);
      eA.swap(eB);
      if (!dry && requests.empty() && diffs == 0 && unk_total == 0) break;
      if (pass > 200) { printf("no convergence\n"); break; }
    }
  }
  if (getenv("PGB_SIM_DUMP")) { FILE *f = fopen(getenv("PGB_SIM_DUMP"), "wb"); fwrite(out.data(), 64, out.size(), f); fclose(f); }
  size_t nb = 0;
  size_t m = std::min(out.size(), ref.size());
  for (size_t i = 0; i < m; i++) if (memcmp(&out[i], &ref[i], sizeof(ovlp_rec)) != 0) { if (nb < 5) fprintf(stderr, "record %zu differs\n", i); nb++; }
  printf("overlap: ref=%zu mine=%zu differing=%zu alignments=%zu\n", ref.size(), out.size(), nb, n_align);
  return (nb || out.size() != ref.size()) ? 3 : 0;
}

static int cmd_dedup() {
  std::vector<ovlp_rec> recs;
  ovlp_rec r;
  while (fread(&r, sizeof r, 1, stdin) == 1) recs.push_back(r);
  std::unordered_set<uint64_t> seen;
  char buf[192];
  for (const ovlp_rec &o : recs) {
    if (!seen.insert(dedup_pair_key(o)).second) continue;
    int n = dedup_format(o, buf);
    fwrite(buf, 1, (size_t)n, stdout);
  }
  return 0;
}

// block > 0: through GzRecordStream with that block size (what shmr_mkseqdb uses; tiny blocks stress the re-scan at block ends)
static int cmd_fastaidx(const char *lst, size_t block) {
  FILE *f = fopen(lst, "r");
  if (!f) return 1;
  char fn[8192];
  uint32_t rid = 0;
  size_t offset = 0;
  std::vector<char> buf;
  while (fscanf(f, "%8191s", fn) != EOF) {
    if (block) {
      GzRecordStream st(fn, block);
      if (!st.ok()) return 1;
      FastaRecord r;
      while (st.next(r)) {
        printf("%09d %s %u %lu\n", rid, r.name.c_str(), (unsigned)r.seq.size(), offset);
        rid++;
        offset += r.seq.size();
      }
      continue;
    }
    if (!slurp_gz(fn, buf)) return 1;
    FastaScanner sc(buf.data(), buf.size());
    FastaRecord r;
    while (sc.next(r)) {
      printf("%09d %s %u %lu\n", rid, r.name.c_str(), (unsigned)r.seq.size(), offset);
      rid++;
      offset += r.seq.size();
    }
  }
  fclose(f);
  return 0;
}

// ------------------------------------------------------------------------------------------------ lane-group row walk (model)
// k_replay_group (kernels.cuh) lets G lanes probe G candidates of a bucket row at once and then applies the row's sequential
// semantics (src/shmr_overlap.c:97-176: stop at bestn overlaps, an accepted CONTAINED ends the row, CONTAINS marks the candidate)
// with ballots.  This is the same decision logic on the CPU, against the product's sequential replay_bucket on random buckets in
// which every read occurs once: the set of visited candidates (table updates, requests, records) and their order must be equal.
struct RowEvent { uint32_t i, j, kind; };  // kind: 1 = table update (accepted, with its type << 4), 2 = alignment request
struct ModelCtx {
  std::unordered_map<uint64_t, uint64_t> E;       // pair -> rank << 2 | type (time-stamped rid_pairs)
  std::unordered_map<uint64_t, match_t> aln;      // (i, j) -> known alignment
  std::vector<uint32_t> rlen_by_rid;
  std::vector<RowEvent> ev;
  uint32_t rank = 0;
  uint32_t rlen(uint32_t rid) const { return rlen_by_rid[rid]; }
  void pair_get(uint64_t p, uint64_t *vold, uint64_t *vnew) const {
    *vnew = ~0ULL;
    auto it = E.find(p);
    *vold = it == E.end() ? ~0ULL : it->second;
  }
  void pair_set(uint64_t p, uint64_t v) { ev.push_back(RowEvent{(uint32_t)(p >> 32), (uint32_t)p, 1u | (uint32_t)(v & 3) << 4}); }
  bool aln_get(uint32_t i, uint32_t j, match_t *m) const {
    auto it = aln.find(((uint64_t)i << 32) | j);
    if (it == aln.end()) return false;
    *m = it->second;
    return true;
  }
  void aln_request(uint32_t i, uint32_t j, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t) { ev.push_back(RowEvent{i, j, 2}); }
  void emit(uint32_t, const ovlp_rec &) {}
};
static int cmd_grouprow(int argc, char **argv) {
  const int trials = argc > 2 ? atoi(argv[2]) : 20000;
  std::mt19937_64 rng(99);
  size_t n_rows_cut = 0, n_rows_ended = 0, n_contains = 0, n_events = 0;
  for (int trial = 0; trial < trials; trial++) {
    const int Gs[4] = {4, 8, 16, 32};
    const uint32_t G = (uint32_t)Gs[trial & 3], n = 2 + (uint32_t)(rng() % 62), bestn = 1 + (uint32_t)(rng() % 6), rank = 1000;
    // a bucket: n records of n different reads, descending positions as the reference's qsort leaves them
    ModelCtx c;
    c.rank = rank;
    c.rlen_by_rid.assign(n + 1, 0);
    std::vector<uint64_t> y0(n);
    std::vector<uint8_t> dir(n), cont_seq(n);
    for (uint32_t t = 0; t < n; t++) {
      c.rlen_by_rid[t] = 9000 + (uint32_t)(rng() % 9000);
      const uint32_t pos = 8000 - t * (uint32_t)(1 + rng() % 100) % 7000;
      y0[t] = ((uint64_t)t << 32) | ((uint64_t)pos << 1) | (rng() & 1);
      dir[t] = (uint8_t)(rng() & 1);
    }
    // what the tables hold: some pairs already in rid_pairs with a lower rank (hits), some alignments known (various outcomes)
    for (uint32_t i = 0; i < n; i++)
      for (uint32_t j = i + 1; j < n; j++) {
        const uint32_t r = (uint32_t)(rng() % 100);
        if (r < 25) c.E[((uint64_t)i << 32) | j] = ((uint64_t)(rng() % rank) << 2) | (rng() % 3);
        else if (r < 85) {
          match_t m;
          predict_match(c.rlen_by_rid[i], c.rlen_by_rid[j], 100, &m);
          const uint32_t o = (uint32_t)(rng() % 4);
          if (o == 0) m.q_end = m.t_end = 100;                                         // rejected (short)
          if (o == 1) { m.q_bgn = 0; m.q_end = (int)c.rlen_by_rid[i]; m.t_end = (int)c.rlen_by_rid[j]; }  // containment
          c.aln[((uint64_t)i << 32) | j] = m;
        }
      }
    // sequential walk (the product's replay_bucket)
    ModelCtx cs = c;
    uint32_t unk_s = 0;
    const uint32_t acc_s = replay_bucket(cs, rank, y0.data(), dir.data(), n, cont_seq.data(), bestn, false, &unk_s);
    // lane-group walk
    ModelCtx cg = c;
    std::vector<uint8_t> ct(n, 0);
    uint32_t acc_g = 0, unk_g = 0;
    for (uint32_t k0 = n - 1; k0 > 0; k0--) {
      const uint32_t i = k0 - 1;
      if (ct[i]) continue;
      const uint32_t rid0 = (uint32_t)(y0[i] >> 32), pos0 = (uint32_t)((y0[i] & 0xFFFFFFFFULL) >> 1) + 1, rlen0 = cg.rlen(rid0);
      uint32_t overlap_count = 0;
      bool row_done = false;
      for (uint32_t base = i + 1; base < n && overlap_count < bestn && !row_done; base += G) {
        uint32_t incm = 0, endm = 0;
        struct Lane { bool valid, hit, known, accepted; uint32_t type, j, rid1, pos1; uint64_t ridp; match_t m; } L[32];
        for (uint32_t gl = 0; gl < G; gl++) {
          Lane &l = L[gl];
          l = Lane();
          l.j = base + gl;
          l.valid = l.j < n && !ct[l.j];
          l.known = true;
          if (!l.valid) continue;
          l.rid1 = (uint32_t)(y0[l.j] >> 32);
          l.ridp = rid0 < l.rid1 ? ((uint64_t)rid0 << 32) | l.rid1 : ((uint64_t)l.rid1 << 32) | rid0;
          uint64_t v, vnew;
          cg.pair_get(l.ridp, &v, &vnew);
          l.hit = v != ~0ULL && (uint32_t)(v >> 2) < rank;
          if (l.hit) { l.type = (uint32_t)(v & 3); }
          else {
            l.pos1 = (uint32_t)((y0[l.j] & 0xFFFFFFFFULL) >> 1) + 1;
            const uint32_t rlen1 = cg.rlen(l.rid1);
            l.known = cg.aln_get(i, l.j, &l.m);
            if (!l.known) predict_match(rlen0, rlen1, pos0 - l.pos1, &l.m);
            l.accepted = classify_match(l.m, rlen0, rlen1, rlen0 - pos0 + l.pos1, rlen1, &l.type);
          }
          if (l.type == OVL_OVERLAP && (l.hit || l.accepted)) incm |= 1u << gl;
          if (!l.hit && l.accepted && l.type == OVL_CONTAINED) endm |= 1u << gl;
        }
        const uint32_t need = bestn - overlap_count;
        uint32_t cut = G;
        if ((uint32_t)__builtin_popcount(incm) >= need) {  // __fns(incm, 0, need): the need-th set bit
          uint32_t seen = 0;
          for (uint32_t b = 0; b < G; b++) if ((incm >> b) & 1u) { if (++seen == need) { cut = b; break; } }
        }
        if (endm) { const uint32_t e = (uint32_t)__builtin_ctz(endm); if (e <= cut) { cut = e; row_done = true; n_rows_ended++; } }
        if (cut < G) n_rows_cut++;
        for (uint32_t gl = 0; gl < G && gl <= cut; gl++) {
          Lane &l = L[gl];
          if (!l.valid || l.hit) continue;
          if (!l.known) { unk_g++; cg.aln_request(i, l.j, 0, 0, 0, 0, 0); }
          if (l.accepted) {
            if (l.type == OVL_CONTAINS) { ct[l.j] = 1; n_contains++; }
            cg.pair_set(l.ridp, ((uint64_t)rank << 2) | l.type);
            acc_g++;
          }
        }
        uint32_t vis = cut >= G - 1 ? (G == 32 ? ~0u : (1u << G) - 1) : ((2u << cut) - 1u);
        overlap_count += (uint32_t)__builtin_popcount(incm & vis);
        if (row_done) ct[i] = 1;
      }
    }
    // the sequential walk issues request and update of one candidate in the same order; compare the event streams
    bool same = acc_s == acc_g && unk_s == unk_g && cs.ev.size() == cg.ev.size() && memcmp(cont_seq.data(), ct.data(), n) == 0;
    for (size_t e = 0; same && e < cs.ev.size(); e++) same = cs.ev[e].i == cg.ev[e].i && cs.ev[e].j == cg.ev[e].j && cs.ev[e].kind == cg.ev[e].kind;
    n_events += cs.ev.size();
    if (!same) {
      fprintf(stderr, "group walk differs: trial %d G %u n %u bestn %u: accepted %u vs %u, unknown %u vs %u, events %zu vs %zu\n", trial, G, n, bestn, acc_s,
              acc_g, unk_s, unk_g, cs.ev.size(), cg.ev.size());
      return 3;
    }
  }
  printf("grouprow: %d buckets, %zu events, %zu row chunks ended by a cut-off (%zu by a CONTAINED result), %zu CONTAINS marks: identical to the sequential walk\n",
         trials, n_events, n_rows_cut, n_rows_ended, n_contains);
  return n_events > 0 && n_rows_cut > n_rows_ended && n_rows_ended > 0 && n_contains > 0 ? 0 : 4;
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: hostsim sketch|match|overlap|grouprow|dedup|fastaidx ...\n"); return 1; }
  if (!strcmp(argv[1], "dedup")) return cmd_dedup();
  if (!strcmp(argv[1], "grouprow")) return cmd_grouprow(argc, argv);
  if (!strcmp(argv[1], "fastaidx") && argc > 2) return cmd_fastaidx(argv[2], argc > 3 ? (size_t)strtoull(argv[3], 0, 10) : 0);
  load_ref();
  if (!strcmp(argv[1], "sketch")) return cmd_sketch(argc, argv);
  if (!strcmp(argv[1], "match")) return cmd_match(argc, argv);
  if (!strcmp(argv[1], "overlap")) return cmd_overlap(argc, argv);
  return 1;
}
