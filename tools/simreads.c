/*
 * simreads — synthetic read-set generator for the SHIMMER index/overlap benchmark.
 *
 * Re-implements the *semantics* of the reference's test simulator (test/ecoli_K12/simulate_reads.py:12-45):
 *   - genome made circular by appending its first 40 kb                      (:29)
 *   - read length int(mean + gauss(0, sd)), start uniform in [0, G]         (:38-39)
 *   - per-base edit with probability p drawn uniformly from the 9-way menu
 *     {A, C, G, T, deletion, c+A, c+C, c+G, c+T}                              (:13-14)
 *   - 50 % of reads reverse-complemented                                      (:42-43)
 *   - FASTA names ">{file:02d}/{i:06d}/0_{len}"                               (:40)
 * but with a counter-based RNG (splitmix64 keyed by seed and read number) instead of Python's Mersenne
 * twister, so that read i can be generated independently (OpenMP) and a 1.5 Gbase set takes seconds, not
 * minutes.  The genome itself is i.i.d. uniform ACGT.
 *
 * Output (what shmr_mkseqdb would produce from the FASTA, src/shmr_mkseqdb.c:99-121):
 *   <prefix>.seqdb : 1 byte/base, low nibble = base (A1 C2 G4 T8), high nibble = complement of the base at
 *                    len-1-p (src/shmr_utils.c:44-51)
 *   <prefix>.idx   : "%09d %s %u %lu\n" rid name len offset
 *   <prefix>.bed   : truth "name start end strand"
 *   optional -f    : also the FASTA (single file <prefix>.fa) so tests can push it through the real shmr_mkseqdb.
 *
 *   -m MOD -r RES  : emit only the reads with rid % MOD == RES (rids stay global, offsets are local to the output): one
 *                    rank's share of a sharded job without ever writing the whole set (read i depends on (seed, i) only).
 *
 * usage: simreads -g GENOME_BP -c COVERAGE [-l 15000] [-s 1500] [-e 0.005] [-S 42] [-n NREADS] [-m MOD -r RES] [-f] -p PREFIX
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

static inline uint64_t splitmix64(uint64_t *s) {
  uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char BASES[4] = {'A', 'C', 'G', 'T'};
static const uint8_t FWD[4] = {1, 2, 4, 8};  /* fourbit_map_f */
static const uint8_t REV[4] = {8, 4, 2, 1};  /* fourbit_map_r: code of the complement */

typedef struct {
  uint8_t *b;  /* 0..3 codes */
  uint32_t len;
  uint64_t start;
  uint32_t span;
  int strand;
} read_t;

static void make_read(const uint8_t *genome, uint64_t G, uint64_t seed, uint64_t i, double mean, double sd, double perr,
                      read_t *r) {
  uint64_t s = seed * 0x100000001B3ULL + i * 0xD6E8FEB86659FD93ULL + 0x1234567ULL;
  (void)splitmix64(&s);
  double u1 = u01(&s), u2 = u01(&s);
  if (u1 < 1e-300) u1 = 1e-300;
  double g = sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
  long rl = (long)(mean + g * sd);
  if (rl < 100) rl = 100;
  if (rl > 39000) rl = 39000;
  uint64_t st = splitmix64(&s) % (G + 1);
  uint8_t *out = (uint8_t *)malloc((size_t)rl * 2 + 16);
  uint32_t n = 0;
  for (long p = 0; p < rl; p++) {
    uint8_t c = genome[(st + p) % G];
    if (u01(&s) < perr) {
      uint32_t m = (uint32_t)(splitmix64(&s) % 9);
      if (m < 4) out[n++] = (uint8_t)m;
      else if (m == 4) { /* deletion */ }
      else { out[n++] = c; out[n++] = (uint8_t)(m - 5); }
    } else
      out[n++] = c;
  }
  int strand = (int)(splitmix64(&s) & 1);
  if (strand) {
    for (uint32_t a = 0, b = n - 1; a < b; a++, b--) {
      uint8_t t = out[a]; out[a] = 3 - out[b]; out[b] = 3 - t;
    }
    if (n & 1) out[n / 2] = 3 - out[n / 2];
  }
  r->b = out; r->len = n; r->start = st; r->span = (uint32_t)rl; r->strand = strand;
}

int main(int argc, char **argv) {
  uint64_t G = 0, seed = 42, nreads = 0, mod = 1, res = 0;
  double cov = 30, mean = 15000, sd = 1500, perr = 0.005;
  const char *prefix = NULL;
  int fasta = 0, c;
  while ((c = getopt(argc, argv, "g:c:l:s:e:S:n:p:fm:r:")) != -1) {
    switch (c) {
      case 'g': G = strtoull(optarg, 0, 10); break;
      case 'c': cov = atof(optarg); break;
      case 'l': mean = atof(optarg); break;
      case 's': sd = atof(optarg); break;
      case 'e': perr = atof(optarg); break;
      case 'S': seed = strtoull(optarg, 0, 10); break;
      case 'n': nreads = strtoull(optarg, 0, 10); break;
      case 'p': prefix = optarg; break;
      case 'f': fasta = 1; break;
      case 'm': mod = strtoull(optarg, 0, 10); break;
      case 'r': res = strtoull(optarg, 0, 10); break;
      default: fprintf(stderr, "bad option\n"); return 1;
    }
  }
  if (!G || !prefix) {
    fprintf(stderr, "usage: simreads -g GENOME_BP -c COVERAGE [-l 15000] [-s 1500] [-e 0.005] [-S 42] [-n NREADS] [-f] -p PREFIX\n");
    return 1;
  }
  if (!nreads) nreads = (uint64_t)(cov * (double)G / mean);
  if (mod == 0 || res >= mod) { fprintf(stderr, "need 0 <= RES < MOD\n"); return 1; }
  uint8_t *genome = (uint8_t *)malloc(G);
  {
    uint64_t s = seed ^ 0xA5A5A5A5DEADBEEFULL;
    for (uint64_t i = 0; i < G; i += 32) {
      uint64_t z = splitmix64(&s);
      for (int j = 0; j < 32 && i + j < G; j++) genome[i + j] = (z >> (2 * j)) & 3;
    }
  }
  char fn[8192];
  snprintf(fn, sizeof fn, "%s.seqdb", prefix); FILE *fdb = fopen(fn, "wb");
  snprintf(fn, sizeof fn, "%s.idx", prefix);   FILE *fidx = fopen(fn, "w");
  snprintf(fn, sizeof fn, "%s.bed", prefix);   FILE *fbed = fopen(fn, "w");
  FILE *ffa = NULL;
  if (fasta) { snprintf(fn, sizeof fn, "%s.fa", prefix); ffa = fopen(fn, "w"); }
  if (!fdb || !fidx || !fbed || (fasta && !ffa)) { perror("open output"); return 1; }
  setvbuf(fdb, NULL, _IOFBF, 1 << 22);

  const uint64_t BATCH = 4096;
  read_t *batch = (read_t *)malloc(sizeof(read_t) * BATCH);
  uint8_t *enc = (uint8_t *)malloc(80000);
  char *asc = (char *)malloc(80000);
  uint64_t per_file = (nreads + 7) / 8, offset = 0, total = 0;
  /* the selected rids are res, res + mod, ...; slot q of the selection is rid res + q * mod */
  const uint64_t nsel = nreads > res ? (nreads - res + mod - 1) / mod : 0;
  for (uint64_t b0 = 0; b0 < nsel; b0 += BATCH) {
    uint64_t nb = nsel - b0 < BATCH ? nsel - b0 : BATCH;
#pragma omp parallel for schedule(dynamic, 16)
    for (uint64_t j = 0; j < nb; j++) make_read(genome, G, seed, res + (b0 + j) * mod, mean, sd, perr, &batch[j]);
    for (uint64_t j = 0; j < nb; j++) {
      read_t *r = &batch[j];
      uint64_t rid = res + (b0 + j) * mod;
      char name[64];
      snprintf(name, sizeof name, "%02d/%06d/0_%u", (int)(rid / per_file), (int)(rid % per_file), r->len);
      for (uint32_t p = 0; p < r->len; p++) enc[p] = (uint8_t)(FWD[r->b[p]] | (REV[r->b[r->len - 1 - p]] << 4));
      if (fwrite(enc, 1, r->len, fdb) != r->len) { perror("simreads: write .seqdb"); return 1; }
      fprintf(fidx, "%09d %s %u %lu\n", (int)rid, name, r->len, (unsigned long)offset);
      fprintf(fbed, "%s\t%lu\t%lu\t%d\n", name, (unsigned long)r->start, (unsigned long)(r->start + r->span), r->strand);
      if (ffa) {
        for (uint32_t p = 0; p < r->len; p++) asc[p] = BASES[r->b[p]];
        asc[r->len] = 0;
        fprintf(ffa, ">%s\n%s\n", name, asc);
      }
      offset += r->len; total += r->len;
      free(r->b);
    }
  }
  if (ferror(fidx) || ferror(fbed) || (ffa && ferror(ffa))) { fprintf(stderr, "simreads: write error\n"); return 1; }
  if (fclose(fdb) | fclose(fidx) | fclose(fbed) | (ffa ? fclose(ffa) : 0)) { perror("simreads: close"); return 1; }
  fprintf(stderr, "simreads: genome=%lu reads=%lu bases=%lu err=%g seed=%lu -> %s.{seqdb,idx,bed}\n", (unsigned long)G,
          (unsigned long)nsel, (unsigned long)total, perr, (unsigned long)seed, prefix);
  return 0;
}
