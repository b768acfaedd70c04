// align_coop.cuh — ovlp_match (src/DWmatch.c:66-204), G cooperating lanes per alignment, branch-free inner loop.
//
// A warp holds 32/G independent alignments ("groups").  Lane l of a group owns the l-th diagonal of the current chunk of the
// band row (diagonals min_k, min_k+2, ... of edit distance d, DWmatch.c:124); the cells of a row depend only on the previous
// row (SURVEY A-8), so they are computed together.  ONE iteration of the loop is ONE 32-base compare per lane:
//   * a group that starts a chunk sets its cells up from the previous row (DWmatch.c:125-131) and every lane compares the first
//     32 bases of its own cell's snake;
//   * a group with a snake still running after its first word extends it COOPERATIVELY: lane j compares bases
//     [32 j, 32 j + 32) past the snake's head, a ballot finds the first mismatch (32 G bases per iteration, DWmatch.c:135-140);
//   * when no snake of the chunk is running any more, the row's order-dependent bookkeeping is resolved in diagonal order with
//     ballots / shuffles: first snake > 16 (:142-146), strictly-longest snake (:148-152), best_m (:157), the end test
//     (:161-164, it hides the cells after it), and after the row's last chunk the band trim to the hull of
//     { k : x+y >= best_m - band_tolerance } (:168-183).
// The loop body is straight-line code: the groups' situations (which row of which alignment, snake running or not) are
// predicates, every ballot / shuffle is issued by all 32 lanes with the full mask, and only rare events leave the stream
// through WARP-UNIFORM branches (a new longest snake, the end of an alignment, a row wider than G).  This is what the
// bulk-staged first version (align_quad.cuh) lacked: there the per-group window maintenance sat in per-group branches, the
// groups of a warp drifted apart and every section was executed ~2.7 times per iteration with 12 of 32 lanes
// (profiles/r2_align.md).  Operands are read straight from the packed images through L1 (a group's working set is the two
// 128-byte lines around its frontier, all groups of an SM together ~100 KB); band rows live in shared memory while a row has
// at most QC_VCAP diagonals, wider rows use a per-group global scratch row.
// Reads that contain N are not handled here (k_align with only_n = 1 does them, as for k_align_lean).
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

#define QC_VCAP 32
#define QC_THREADS 128

struct QcGroupSmem {
  int V[2][QC_VCAP];
};

__device__ __forceinline__ int qc_match_len(uint64_t df) {  // leading equal bases of two 32-base windows, given their XOR
  if (!df) return 32;
  const uint32_t lo = (uint32_t)df;
  return lo ? (ctz32(lo) >> 1) : 16 + (ctz32((uint32_t)(df >> 32)) >> 1);
}

template <int G>
__global__ void __launch_bounds__(QC_THREADS) k_align_coop(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n, const uint32_t *__restrict__ perm,
                                                            const uint64_t *__restrict__ w, const uint64_t *__restrict__ wrc,
                                                            const uint64_t *__restrict__ woff_by_rid, const uint32_t *__restrict__ rlen_by_rid,
                                                            const uint32_t *__restrict__ hasn_by_rid, int bw, match_t *results,
                                                            unsigned long long *bases_total, unsigned int *queue_head, int *vscratch, int vcap_g) {
  static_assert(G == 2 || G == 4 || G == 8 || G == 16, "G lanes per alignment");
  __shared__ QcGroupSmem qc_sm[QC_THREADS / G];
  constexpr uint32_t GBITS = (1u << G) - 1u;
  constexpr uint32_t FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u;
  const int gl = (int)(lane & (G - 1));
  const uint32_t gshift = lane & ~(uint32_t)(G - 1);
  const uint32_t gm = GBITS << gshift;
  QcGroupSmem *sm = &qc_sm[threadIdx.x / G];
  int *gV = vscratch + (size_t)(blockIdx.x * (QC_THREADS / G) + threadIdx.x / G) * 2 * (size_t)vcap_g;
  auto gbits = [&](uint32_t bal) -> uint32_t { return (bal >> gshift) & GBITS; };

  // ---- per-alignment state (group-uniform unless noted)
  const uint64_t *qarr = w, *tarr = w;  // first word of the operand's read in the image of its strand
  uint32_t qo = 0;                      // base offset of the query's logical base 0 in its read (the target starts at 0)
  uint32_t slot = 0;
  int q_len = 0, t_len = 0, max_d = 0;
  int d = 0, min_k = 0, max_k = 0, pbase = 0, best_m = -1, c0 = 0, nk = 1;
  uint32_t longest = 0;
  bool start = false;
  int q_bgn = 0, t_bgn = 0, q_m_end = 0, t_m_end = 0;
  int *Vp = sm->V[0], *Vc = sm->V[1];
  int cur = 1;  // Vc is buffer `cur`
  unsigned long long bases = 0;
  bool coop = false;  // a snake of the chunk is being extended by the whole group
  int lead = 0;       // ... the lane that owns it
  // lane-local cell state
  int k = 0, x = 0, x1 = 0;
  bool cell = false, pend = false;

  auto row_buf = [&](int buf, int cells) -> int * { return cells <= QC_VCAP ? sm->V[buf] : gV + (size_t)buf * vcap_g; };
  auto finish = [&](const match_t &r) {
    if (gl == 0) {
      int4 *dst = reinterpret_cast<int4 *>(&results[slot]);
      dst[0] = make_int4(r.m_size, r.dist, r.q_bgn, r.q_end);
      dst[1] = make_int4(r.t_bgn, r.t_end, r.t_m_end, r.q_m_end);
      bases += (unsigned long long)(r.q_end + r.t_end);
    }
  };
  // next alignment of the queue -> state; false when the queue is empty.  Runs per group (rare; per-group collectives).
  auto fetch = [&]() -> bool {
    cell = false; pend = false; coop = false; lead = 0;
    for (;;) {
      uint32_t i = 0;
      if (gl == 0) i = atomicAdd(queue_head, 1u);
      i = (uint32_t)__shfl_sync(gm, (int)i, 0, G);
      if (i >= n) return false;
      if (perm) i = perm[i];
      const AlnReq q = reqs[first + i];
      if (hasn_by_rid[q.rid0] | hasn_by_rid[q.rid1]) continue;  // left to k_align(only_n = 1)
      const uint32_t rl0 = rlen_by_rid[q.rid0], rl1 = rlen_by_rid[q.rid1];
      slot = q.slot;
      q_len = (int)(rl0 - q.start0); t_len = (int)rl1;
      max_d = (int)(0.3 * (double)(q_len + t_len));  // DWmatch.c:96
      if (max_d <= 0) {  // no row runs (DWmatch.c:118): all-zero result
        match_t r;
        r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = r.t_m_end = r.q_m_end = 0;
        finish(r);
        continue;
      }
      qarr = ((q.strands & 1) ? wrc : w) + woff_by_rid[q.rid0];
      tarr = ((q.strands & 2) ? wrc : w) + woff_by_rid[q.rid1];
      qo = q.start0;
      d = 0; min_k = 0; max_k = 0; pbase = 0; best_m = -1; c0 = 0; nk = 1;
      longest = 0; start = false; q_bgn = t_bgn = q_m_end = t_m_end = 0;
      cur = 1; Vp = sm->V[0]; Vc = sm->V[1];
      return true;
    }
  };

  bool active = fetch();
  uint32_t iters = 0;
  for (;;) {
    __syncwarp();  // band-row entries written below are read by other lanes of the group in the next set-up
    if (!__any_sync(FULL, active)) break;
    if (++iters > (1u << 27)) __trap();  // (a group runs ~10^5 iterations per launch)
    // ------------------------------------------------------------------ set-up of a chunk's cells (groups that are not extending a snake)
    if (active && !coop) {
      const int idx = c0 + gl;
      cell = idx < nk;
      k = min_k + 2 * idx;
      x = 0;
      if (cell && d > 0) {  // DWmatch.c:125-131
        const int i_lo = (k - 1 - pbase) >> 1;  // V[d-1][k-1]; V[d-1][k+1] is the next entry
        const int vm = Vp[k == min_k ? i_lo + 1 : i_lo], vp = Vp[k == max_k ? i_lo : i_lo + 1];
        x = (k == min_k) ? vp : ((k == max_k || !(vm < vp)) ? vm + 1 : vp);
      }
      x1 = x;
    }
    // ------------------------------------------------------------------ one 32-base compare per lane
    const int xl = __shfl_sync(FULL, x, lead, G);               // head of the group's running snake (if any)
    const int kk = coop ? min_k + 2 * (c0 + lead) : k;
    const int px = coop ? xl + 32 * gl : x, py = px - kk;
    const int rem = (q_len - px) < (t_len - py) ? (q_len - px) : (t_len - py);
    int m = 0;
    if (active && (coop || cell) && rem > 0) {
      const uint32_t qb = qo + (uint32_t)px, tb = (uint32_t)py;
      const uint64_t *qp = qarr + (qb >> 5), *tp = tarr + (tb >> 5);
      m = qc_match_len(window64(qp[0], qp[1], (qb & 31u) * 2u) ^ window64(tp[0], tp[1], (tb & 31u) * 2u));
      if (m > rem) m = rem;
    }
    // ---- interpret it
    const uint32_t stop = gbits(__ballot_sync(FULL, coop && m < 32));  // lanes of a cooperative round whose window holds the snake's end
    const int fs = stop ? __ffs((int)stop) - 1 : 0;
    const int mf = __shfl_sync(FULL, m, fs, G);
    if (coop) {
      if (gl == lead) {
        // rem of lane 0 of the group is the snake's own remainder
        x = stop ? xl + 32 * fs + mf : xl + 32 * G;
      }
      const int remc = rem + 32 * gl;
      if (gl == lead) pend = !stop && remc > 32 * G;
    } else {
      x += m;
      pend = (m == 32) && (rem > 32);
    }
    const uint32_t pm = gbits(__ballot_sync(FULL, pend));
    coop = pm != 0;
    lead = coop ? __ffs((int)pm) - 1 : 0;  // lowest lane with a running snake: the whole group extends it next
    // ------------------------------------------------------------------ the chunk's snakes have ended
    const bool s3 = active && !coop;
    const int y = x - k;
    const uint32_t em_w = __ballot_sync(FULL, s3 && cell && (x >= q_len || y >= t_len));  // DWmatch.c:161
    const uint32_t em = gbits(em_w);
    const int Lm = em ? __ffs((int)em) - 1 : G;  // first cell that reaches an end: later cells of the row are not visited
    const bool valid = s3 && cell && gl <= Lm;
    const int len = x - x1;
    {  // DWmatch.c:142-146 (rare after the first rows: warp-uniform branch)
      const uint32_t st_w = __ballot_sync(FULL, valid && !start && len > 16);
      if (st_w) {
        const uint32_t s16 = gbits(st_w);
        const int f16 = s16 ? __ffs((int)s16) - 1 : 0;
        const int xs = __shfl_sync(FULL, x1, f16, G);
        if (s16) { q_bgn = xs; t_bgn = xs - (min_k + 2 * (c0 + f16)); start = true; }
      }
    }
    {  // DWmatch.c:148-152: strictly longer than every snake before it; the first such cell in diagonal order
      const bool longer = valid && (uint32_t)len > longest;
      if (__ballot_sync(FULL, longer)) {  // warp-uniform
        int key = longer ? (len * G + (G - 1 - gl)) : -1;
#pragma unroll
        for (int s = 1; s < G; s <<= 1) { const int o = __shfl_xor_sync(FULL, key, s, G); key = o > key ? o : key; }
        const int fl = key >= 0 ? G - 1 - (key & (G - 1)) : 0;
        const int xm = __shfl_sync(FULL, x, fl, G);
        if (key >= 0) { longest = (uint32_t)(key / G); q_m_end = xm; t_m_end = xm - (min_k + 2 * (c0 + fl)); }
      }
    }
    {  // DWmatch.c:157
      int u = valid ? (x + y) : -1;
#pragma unroll
      for (int s = 1; s < G; s <<= 1) { const int o = __shfl_xor_sync(FULL, u, s, G); u = o > u ? o : u; }
      if (u > best_m) best_m = u;
    }
    if (s3 && cell) Vc[c0 + gl] = x;
    const int thr = best_m - bw;
    const uint32_t hm = gbits(__ballot_sync(FULL, s3 && cell && (x + y) >= thr));  // hull of a row that fits one chunk
    bool done = false;
    if (em_w) {  // warp-uniform: some group's alignment reached an end (DWmatch.c:185-194)
      const int xe = __shfl_sync(FULL, x, em ? Lm : 0, G);
      if (em) {
        match_t r;
        r.q_end = xe;
        r.t_end = xe - (min_k + 2 * (c0 + Lm));
        r.dist = d;
        r.q_bgn = q_bgn; r.t_bgn = t_bgn; r.q_m_end = q_m_end; r.t_m_end = t_m_end;
        r.m_size = (r.q_end - r.q_bgn + r.t_end - r.t_bgn + 2 * d) / 2;
        finish(r);
        done = true;
      }
    }
    const bool row_end = s3 && !done && !(c0 + G < nk);
    if (s3 && !done && !row_end) c0 += G;  // next chunk of this row
    const bool wide = row_end && nk > G;
    int new_min_k = min_k + 2 * (__ffs((int)hm) - 1), new_max_k = min_k + 2 * (31 - __clz((int)hm));  // (a row that fits one chunk)
    if (__any_sync(FULL, wide)) {  // warp-uniform, rare: scan the stored row of a wide band, per-group collectives
      if (wide) {
        __syncwarp(gm);
        int lo_i = 0x7fffffff, hi_i = -1;
        for (int i = gl; i < nk; i += G)
          if (2 * Vc[i] - (min_k + 2 * i) >= thr) { if (i < lo_i) lo_i = i; hi_i = i; }
#pragma unroll
        for (int s = 1; s < G; s <<= 1) {
          const int a = __shfl_xor_sync(gm, lo_i, s, G), b = __shfl_xor_sync(gm, hi_i, s, G);
          lo_i = a < lo_i ? a : lo_i;
          hi_i = b > hi_i ? b : hi_i;
        }
        new_min_k = min_k + 2 * lo_i;
        new_max_k = min_k + 2 * hi_i;
      }
    }
    if (row_end) {  // band trim (DWmatch.c:168-183), next row
      pbase = min_k;
      min_k = new_min_k - 1;
      max_k = new_max_k + 1;
      d++;
      c0 = 0;
      nk = ((max_k - min_k) >> 1) + 1;
      Vp = Vc;
      cur ^= 1;
      Vc = row_buf(cur, nk);
    }
    const bool fail = row_end && (d >= max_d || max_k - min_k > 2 * bw);  // DWmatch.c:118-122: not matched (:196-199)
    if (__any_sync(FULL, fail || done)) {  // warp-uniform, once per alignment
      if (fail) {
        match_t r;
        r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = 0;
        r.q_m_end = q_m_end; r.t_m_end = t_m_end;
        finish(r);
      }
      if (fail || done) active = fetch();
    }
  }
  if (gl == 0 && bases) atomicAdd(bases_total, bases);
}

}  // namespace pgb
