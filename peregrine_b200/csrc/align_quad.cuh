// align_quad.cuh — ovlp_match (src/DWmatch.c:66-204) with G cooperating lanes per alignment and the operand windows staged
// in shared memory by bulk-async copies (cp.async.bulk + mbarrier).
//
// Work split.  A warp holds 32/G independent alignments ("groups").  Within a group lane l owns the l-th diagonal of the
// current chunk of the band row (diagonals min_k, min_k+2, ... of edit distance d, DWmatch.c:124); the cells of a row only
// depend on the previous row (SURVEY A-8), so they are computed together:
//   S1  cell set-up from the previous row (DWmatch.c:125-131) + the first 32-base word-step of every cell's snake;
//   S2  a snake that is still running after its first word is extended COOPERATIVELY: lane j compares bases
//       [32 j, 32 j + 32) past its head, a ballot finds the first mismatch (up to 32 G bases per round, DWmatch.c:135-140);
//   S3  the row's order-dependent bookkeeping, resolved in diagonal order with group ballots / shuffles: first snake > 16
//       (:142-146), strictly-longest snake (:148-152), best_m (:157), the end test (:161-164) which hides the cells after it;
//       after the last chunk of a row: band trim to the hull of { k : x+y >= best_m - band_tolerance } (:168-183).
// The three sections form ONE loop body executed by the whole warp in lockstep: every warp-level primitive of the common path
// (ballots, shuffles) is issued with the FULL mask by all 32 lanes, the groups' different situations (different rows of
// different alignments, snake running or not) are predicates, not branches.  (A first version let each group branch on its
// own state and used per-group masks: results were right, but once groups diverge nothing merges them again inside the loop,
// every group ran its own instruction stream and the kernel took 127 ms instead of 20; profiles/r2_align.md.)  Only rare
// events branch per group: window maintenance, rows wider than G, the end of an alignment.  A group spends
// max(1, #cooperative rounds) iterations per chunk of G diagonals.
//
// Operand staging.  Each group owns, per operand, a ring of two 256-byte stages (2 x 1024 bases of the 2-bit image) in shared
// memory.  An elected lane moves whole 256-byte blocks of the packed read image with cp.async.bulk (1-D TMA) and an mbarrier
// per stage (expect_tx / complete_tx); lanes read the ring with LDS.  The band only moves forward, so the block after the
// window is requested as soon as the band's floor (a bound on every base a later row can touch, from best_m and the band
// limits) has left the window's first block: the copy has ~10 rows of the recurrence to complete before a snake reaches it.
// ensure() checks every access range against the window and waits for / advances / repositions it as needed, so correctness
// never depends on the prefetch heuristic.
//
// Band rows live in shared memory while a row has at most QA_VCAP diagonals; wider rows use a per-group global scratch row.
// Reads that contain N are not handled here (k_align with only_n = 1 does them, as for k_align_lean).
#pragma once
#include "shimmer_core.cuh"

namespace pgb {

#define QA_RING_WORDS 64   // per operand: 2 stages x 32 words
#define QA_BLK_SHIFT 10    // bases per block = 1024 (256 bytes)
#define QA_BLK_BYTES 256
#define QA_VCAP 32
#define QA_THREADS 128

struct QaGroupSmem {
  uint64_t ring[2][QA_RING_WORDS];  // [operand][word & 63]
  int V[2][QA_VCAP];
  unsigned long long bar[4];        // [operand * 2 + stage]
};

__device__ __forceinline__ uint32_t qa_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qa_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void qa_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qa_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void qa_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool qa_mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

// window of one operand of one group (all fields are identical in the G lanes of the group); positions are relative to
// the first byte of the 256-byte block that holds the operand's logical base 0, so that everything fits 32 bits
struct QaWin {
  const char *src0;  // that block, in the packed image the operand lives in (forward or reverse-complement)
  uint32_t a0;       // offset (bases) of logical base 0 inside it
  uint32_t rot;      // parity of its absolute block number: relative block b lives in stage (b + rot) & 1
  uint32_t nblk;     // relative blocks that may be read (the image ends there)
  uint32_t wstart;   // relative blocks wstart and wstart + 1 have been requested
  uint32_t nwaited;  // how many of them are known to have landed (0..2)
};

// BULK = true: a stage is filled by ONE cp.async.bulk (UBLKCP, complete_tx on the stage's mbarrier) issued by the group's first
// lane.  BULK = false: every lane of the group copies 16-byte pieces with cp.async (LDGSTS) and lets the mbarrier count the
// completion of its own copies (cp.async.mbarrier.arrive.noinc).  Measured on the bench workload (profiles/r2_align.md).
template <int G, bool BULK>
struct QaLane {
  uint32_t gmask, gl, gshift;
  uint32_t ring_s[2];  // shared address of the operand rings
  uint32_t bar_s;      // shared address of bar[0]
  uint32_t parity;     // bit (op * 2 + stage): parity the next wait on that barrier uses

  __device__ __forceinline__ void issue(int op, const QaWin &wn, uint32_t blk) {
    __syncwarp(gmask);  // every lane of the group is done reading the stage that is overwritten
    const uint32_t st = (blk + wn.rot) & 1u, bar = bar_s + 8u * (uint32_t)(op * 2 + st);
    if (BULK) {
      if (gl == 0) {
        if (blk < wn.nblk) {
          qa_mbar_expect_tx(bar, QA_BLK_BYTES);
          qa_bulk_g2s(ring_s[op] + st * QA_BLK_BYTES, wn.src0 + (size_t)blk * QA_BLK_BYTES, QA_BLK_BYTES, bar);
        } else {
          qa_mbar_arrive(bar);  // past the end of the image: nothing to fetch, complete the phase
        }
      }
    } else {
      if (blk < wn.nblk) {
        const char *src = wn.src0 + (size_t)blk * QA_BLK_BYTES;
        const uint32_t dst = ring_s[op] + st * QA_BLK_BYTES;
#pragma unroll
        for (uint32_t i = gl; i < QA_BLK_BYTES / 16; i += G)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + 16u * i) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");  // arrives when this lane's copies have landed
      } else {
        qa_mbar_arrive(bar);
      }
    }
  }
  __device__ __forceinline__ void wait(int op, const QaWin &wn, uint32_t blk) {
    const uint32_t b = (uint32_t)(op * 2) + ((blk + wn.rot) & 1u), bar = bar_s + 8u * b;
    const uint32_t par = (parity >> b) & 1u;
    uint32_t spins = 0;
    while (!qa_mbar_try_wait(bar, par))
      if (++spins > (1u << 22)) __trap();  // a copy that never lands is a bug in the window bookkeeping: fail loudly, do not hang the GPU
    parity ^= 1u << b;
  }
  // start a window at relative block b0 (both stages free)
  __device__ __forceinline__ void open(int op, QaWin &wn, uint32_t b0) {
    wn.wstart = b0; wn.nwaited = 0;
    issue(op, wn, b0);
    issue(op, wn, b0 + 1);
  }
  __device__ __forceinline__ void drain(int op, QaWin &wn) {
    for (; wn.nwaited < 2; wn.nwaited++) wait(op, wn, wn.wstart + wn.nwaited);
  }
  __device__ __forceinline__ void advance(int op, QaWin &wn) {  // drop block wstart, request wstart + 2
    if (wn.nwaited == 0) { wait(op, wn, wn.wstart); wn.nwaited = 1; }  // keeps the phase bookkeeping in step
    issue(op, wn, wn.wstart + 2);
    wn.wstart++; wn.nwaited--;
  }
  // make bases [lo, hi) (relative positions, hi - lo <= 1025) readable from the ring
  __device__ __forceinline__ void ensure(int op, QaWin &wn, uint32_t lo, uint32_t hi) {
    const uint32_t b0 = lo >> QA_BLK_SHIFT, b1 = (hi - 1) >> QA_BLK_SHIFT;
    if (b0 - wn.wstart >= 2u) {  // outside the window (before it, or beyond its second block): reposition (rare)
      drain(op, wn);
      open(op, wn, b0);
    } else if (b1 == wn.wstart + 2u) {
      advance(op, wn);
    }
    const uint32_t need = b1 - wn.wstart + 1u;
    while (wn.nwaited < need) { wait(op, wn, wn.wstart + wn.nwaited); wn.nwaited++; }
  }
  // prefetch hint: nothing below `floor` will be read again
  __device__ __forceinline__ void hint(int op, QaWin &wn, uint32_t floor) {
    if ((floor >> QA_BLK_SHIFT) > wn.wstart) advance(op, wn);
  }
};

// 32 bases starting at relative position r of the operand whose ring is at shared address ring_s (rot32 = 32 * QaWin::rot)
__device__ __forceinline__ uint64_t qa_fetch(uint32_t ring_s, uint32_t rot32, uint32_t r) {
  const uint32_t wi = (r >> 5) + rot32;
  uint64_t lo, hi;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(lo) : "r"(ring_s + 8u * (wi & (QA_RING_WORDS - 1))));
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(hi) : "r"(ring_s + 8u * ((wi + 1u) & (QA_RING_WORDS - 1))));
  return window64(lo, hi, (r & 31u) * 2u);
}
__device__ __forceinline__ int qa_match_len(uint64_t df) {  // leading equal bases of two 32-base windows, given their XOR
  if (!df) return 32;
  const uint32_t lo = (uint32_t)df;
  return lo ? (ctz32(lo) >> 1) : 16 + (ctz32((uint32_t)(df >> 32)) >> 1);
}

template <int G, bool BULK>
__global__ void __launch_bounds__(QA_THREADS) k_align_quad(const AlnReq *__restrict__ reqs, uint32_t first, uint32_t n, const uint32_t *__restrict__ perm,
                                                            const uint64_t *__restrict__ w, const uint64_t *__restrict__ wrc, uint64_t arr_words,
                                                            const uint64_t *__restrict__ woff_by_rid, const uint32_t *__restrict__ rlen_by_rid,
                                                            const uint32_t *__restrict__ hasn_by_rid, int bw, match_t *results,
                                                            unsigned long long *bases_total, unsigned int *queue_head, int *vscratch, int vcap_g) {
  static_assert(G == 2 || G == 4 || G == 8 || G == 16, "G lanes per alignment");
  extern __shared__ __align__(128) unsigned char qa_smem[];
  constexpr int GPW = 32 / G;
  constexpr uint32_t GBITS = (1u << G) - 1u;
  constexpr uint32_t FULL = 0xffffffffu;
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  QaLane<G, BULK> L;
  L.gl = lane & (G - 1);
  L.gshift = lane & ~(uint32_t)(G - 1);
  L.gmask = GBITS << L.gshift;
  QaGroupSmem *sm = reinterpret_cast<QaGroupSmem *>(qa_smem) + (wid * GPW + lane / G);
  L.ring_s[0] = qa_smem_u32(&sm->ring[0][0]);
  L.ring_s[1] = qa_smem_u32(&sm->ring[1][0]);
  L.bar_s = qa_smem_u32(&sm->bar[0]);
  L.parity = 0;
  const uint32_t n_blocks = (uint32_t)((arr_words * 8 + QA_BLK_BYTES - 1) / QA_BLK_BYTES);  // (the allocation is a multiple of 512 bytes)
  if (L.gl == 0) {
    for (int b = 0; b < 4; b++) qa_mbar_init(L.bar_s + 8u * b, BULK ? 1 : G);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  int *gV = vscratch + (size_t)(blockIdx.x * (QA_THREADS / G) + threadIdx.x / G) * 2 * (size_t)vcap_g;
  const uint32_t gm = L.gmask;
  const int gl = (int)L.gl;
  auto gbits = [&](uint32_t bal) -> uint32_t { return (bal >> L.gshift) & GBITS; };
  // per-group collectives for the rare paths that branch per group (window maintenance, wide rows, end of an alignment)
  auto gshfl = [&](int v, int src) -> int { return __shfl_sync(gm, v, src, G); };

  // ---- per-alignment state (group-uniform unless noted)
  QaWin wq, wt;
  wq.src0 = wt.src0 = (const char *)w; wq.a0 = wt.a0 = 0; wq.rot = wt.rot = 0; wq.nblk = wt.nblk = 0;
  wq.wstart = wt.wstart = 0; wq.nwaited = wt.nwaited = 2;  // "nothing outstanding"
  bool opened = false;
  uint32_t slot = 0;
  int q_len = 0, t_len = 0, max_d = 0;
  int d = 0, min_k = 0, max_k = 0, pbase = 0, best_m = -1, c0 = 0, nk = 1;
  uint32_t longest = 0;
  bool start = false;
  int q_bgn = 0, t_bgn = 0, q_m_end = 0, t_m_end = 0;
  int *Vp = sm->V[0], *Vc = sm->V[1];
  int cur = 1;              // Vc is buffer `cur`
  int lo_x = 0, hi_x = 66, lo_y = 0, hi_y = 66;  // access range of the row's first word-steps (logical bases)
  unsigned long long bases = 0;
  // lane-local cell state
  int k = 0, x = 0, x1 = 0;
  bool cell = false, pend = false;
  bool gpend = false;  // some lane of the group has a running snake

  auto row_buf = [&](int buf, int cells) -> int * { return cells <= QA_VCAP ? sm->V[buf] : gV + (size_t)buf * vcap_g; };
  auto finish = [&](const match_t &r) {
    if (gl == 0) {
      int4 *dst = reinterpret_cast<int4 *>(&results[slot]);
      dst[0] = make_int4(r.m_size, r.dist, r.q_bgn, r.q_end);
      dst[1] = make_int4(r.t_bgn, r.t_end, r.t_m_end, r.q_m_end);
      bases += (unsigned long long)(r.q_end + r.t_end);
    }
  };
  auto set_window = [&](QaWin &wn, const uint64_t *arr, uint64_t A) {
    const uint32_t blk0 = (uint32_t)(A >> QA_BLK_SHIFT);
    wn.src0 = (const char *)arr + (size_t)blk0 * QA_BLK_BYTES;
    wn.a0 = (uint32_t)A & ((1u << QA_BLK_SHIFT) - 1u);
    wn.rot = blk0 & 1u;
    wn.nblk = n_blocks - blk0;
  };
  // next alignment of the queue -> state; false when the queue is empty.  Runs per group (branches, per-group collectives).
  auto fetch = [&]() -> bool {
    cell = false; pend = false; gpend = false;
    for (;;) {
      uint32_t i = 0;
      if (gl == 0) i = atomicAdd(queue_head, 1u);
      i = (uint32_t)gshfl((int)i, 0);
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
      if (opened) { L.drain(0, wq); L.drain(1, wt); }
      set_window(wq, (q.strands & 1) ? wrc : w, woff_by_rid[q.rid0] * 32 + q.start0);
      set_window(wt, (q.strands & 2) ? wrc : w, woff_by_rid[q.rid1] * 32);
      L.open(0, wq, 0);
      L.open(1, wt, 0);
      opened = true;
      d = 0; min_k = 0; max_k = 0; pbase = 0; best_m = -1; c0 = 0; nk = 1;
      longest = 0; start = false; q_bgn = t_bgn = q_m_end = t_m_end = 0;
      cur = 1; Vp = sm->V[0]; Vc = sm->V[1];
      lo_x = 0; hi_x = 66; lo_y = 0; hi_y = 66;
      return true;
    }
  };

  bool active = fetch();
  uint32_t iters = 0;
  for (;;) {
    __syncwarp();  // band-row entries written in S3 are read by other lanes of the group in S1
    if (!__any_sync(FULL, active)) break;
    if (++iters > (1u << 27)) __trap();  // (a group runs ~10^5 iterations per launch)
    // ------------------------------------------------------------------ S1: start a chunk of the row
    if (active && !gpend) {
      const int idx = c0 + gl;
      cell = idx < nk;
      k = min_k + 2 * idx;
      x = 0;
      if (cell && d > 0) {  // DWmatch.c:125-131
        const int i_lo = (k - 1 - pbase) >> 1;  // V[d-1][k-1]; V[d-1][k+1] is the next entry
        if (k == min_k) x = Vp[i_lo + 1];
        else if (k == max_k) x = Vp[i_lo] + 1;
        else {
          const int vm = Vp[i_lo], vp = Vp[i_lo + 1];
          x = (vm < vp) ? vp : vm + 1;
        }
      }
      x1 = x;
      L.ensure(0, wq, wq.a0 + (uint32_t)lo_x, wq.a0 + (uint32_t)hi_x);
      L.ensure(1, wt, wt.a0 + (uint32_t)lo_y, wt.a0 + (uint32_t)hi_y);
      pend = false;
      if (cell) {
        const int y = x - k;
        const int rem = (q_len - x) < (t_len - y) ? (q_len - x) : (t_len - y);
        if (rem > 0) {
          int nn = qa_match_len(qa_fetch(L.ring_s[0], wq.rot << 5, wq.a0 + (uint32_t)x) ^ qa_fetch(L.ring_s[1], wt.rot << 5, wt.a0 + (uint32_t)y));
          if (nn > rem) nn = rem;
          x += nn;
          pend = (nn == 32) && (rem > 32);
        }
      }
    }
    // ------------------------------------------------------------------ S2: one cooperative round of a running snake
    const uint32_t pm = gbits(__ballot_sync(FULL, pend));
    const bool s2 = pm != 0;
    const int c = s2 ? __ffs((int)pm) - 1 : 0;  // lowest lane with a running snake: the whole group extends it
    const int xc = __shfl_sync(FULL, x, c, G);
    int m = 32, remc = 0;
    if (s2) {
      const int yc = xc - (min_k + 2 * (c0 + c));
      remc = (q_len - xc) < (t_len - yc) ? (q_len - xc) : (t_len - yc);  // > 0: the snake is running
      L.ensure(0, wq, wq.a0 + (uint32_t)xc, wq.a0 + (uint32_t)xc + 32 * G + 64);
      L.ensure(1, wt, wt.a0 + (uint32_t)yc, wt.a0 + (uint32_t)yc + 32 * G + 64);
      const int off = 32 * gl;
      m = 0;  // bases of this lane's 32-base window that extend the snake
      if (remc > off) {
        m = qa_match_len(qa_fetch(L.ring_s[0], wq.rot << 5, wq.a0 + (uint32_t)(xc + off)) ^
                         qa_fetch(L.ring_s[1], wt.rot << 5, wt.a0 + (uint32_t)(yc + off)));
        if (m > remc - off) m = remc - off;
      }
    }
    const uint32_t stop = gbits(__ballot_sync(FULL, s2 && m < 32));
    const int fs = stop ? __ffs((int)stop) - 1 : 0;
    const int mf = __shfl_sync(FULL, m, fs, G);
    if (s2 && gl == c) {
      if (stop) { x = xc + 32 * fs + mf; pend = false; }
      else { x = xc + 32 * G; pend = remc > 32 * G; }
    }
    gpend = gbits(__ballot_sync(FULL, pend)) != 0;
    // ------------------------------------------------------------------ S3: the chunk's snakes have ended
    const bool s3 = active && !gpend;
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
        const int xl = __shfl_sync(FULL, x, fl, G);
        if (key >= 0) { longest = (uint32_t)(key / G); q_m_end = xl; t_m_end = xl - (min_k + 2 * (c0 + fl)); }
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
    if (s3 && !done) {
      if (c0 + G < nk) {
        c0 += G;  // next chunk of this row
      } else {
        // ---- end of the row: band trim (DWmatch.c:168-183)
        int new_min_k, new_max_k;
        if (nk <= G) {  // never empty: the cell that holds best_m of this row qualifies
          new_min_k = min_k + 2 * (__ffs((int)hm) - 1);
          new_max_k = min_k + 2 * (31 - __clz((int)hm));
        } else {  // wide row (rare): scan the stored row, per-group collectives
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
        pbase = min_k;
        min_k = new_min_k - 1;
        max_k = new_max_k + 1;
        d++;
        c0 = 0;
        nk = ((max_k - min_k) >> 1) + 1;
        if (d >= max_d || max_k - min_k > 2 * bw) {  // DWmatch.c:118-122: not matched (:196-199)
          match_t r;
          r.m_size = r.dist = r.q_bgn = r.q_end = r.t_bgn = r.t_end = 0;
          r.q_m_end = q_m_end; r.t_m_end = t_m_end;
          finish(r);
          done = true;
        } else {
          Vp = Vc;
          cur ^= 1;
          Vc = row_buf(cur, nk);
          // every base a later row touches lies at or above these floors (hull cells have x+y >= thr), and the first
          // word-steps of the next row stay below the ceilings (x+y <= best_m in the row just finished)
          int fx = (thr + new_min_k) >> 1, fy = (thr - new_max_k) >> 1;
          fx = fx < 2 ? 0 : fx - 2;
          fy = fy < 2 ? 0 : fy - 2;
          lo_x = fx; lo_y = fy;
          hi_x = ((best_m + max_k) >> 1) + 2 + 66;
          hi_y = ((best_m - min_k) >> 1) + 2 + 66;
          L.hint(0, wq, wq.a0 + (uint32_t)fx);
          L.hint(1, wt, wt.a0 + (uint32_t)fy);
        }
      }
    }
    if (done) active = fetch();
  }
  if (opened) { L.drain(0, wq); L.drain(1, wt); }
  if (gl == 0 && bases) atomicAdd(bases_total, bases);
}

}  // namespace pgb
