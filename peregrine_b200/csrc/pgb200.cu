// pgb200.cu — host orchestration + C ABI of libpgb200.so (see include/pgb200.h).
// All results are produced by the kernels in kernels.cuh; the host code here moves bytes, sizes buffers, replays the
// keys-only khash model that fixes the bucket visiting order, and drives the replay/align fix-point loop.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>
#include <getopt.h>
#include <assert.h>
#include <chrono>
#include <cmath>
#include <unordered_map>
#include <mutex>
#include <functional>
#include "../../include/pgb200.h"
#include "host_util.hpp"
#include "fasta_reader.hpp"
#include "kernels.cuh"

using namespace pgb;

static_assert(sizeof(ovlp_t) == 64 && sizeof(ovlp_rec) == 64, "ovlp_t layout");
static_assert(sizeof(mm_count_t) == 16 && sizeof(mc_entry) == 16, "mm_count_t layout");
static_assert(sizeof(mm128_t) == 16 && sizeof(ovlp_match_t) == 32, "mm128_t / ovlp_match_t layout");

#define CU(call)                                                                                             \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess) {                                                                                 \
      fprintf(stderr, "pgb200: CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_));                        \
    }                                                                                                        \
  } while (0)

static inline uint32_t pow2_at_least(uint64_t v) {
  uint64_t p = 1024;
  while (p < v) p <<= 1;
  if (p > (1ull << 31)) throw std::runtime_error("hash table too large");
  return (uint32_t)p;
}
static inline unsigned nblk(size_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }
static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct pgb_ctx {
  int device = 0, sm_count = 148;
  cudaStream_t st = nullptr;
  std::string err;
  pgb_stats stats;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
  cudaEvent_t user_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_pass[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // fix-point pass: [0,1] pass, [2,3] k_replay, [4,5] alignment batch
  PassOut *h_pass = nullptr;  // page-locked mailbox of the fix-point loop
  unsigned long long *d_align_bases = nullptr;

  // ---- reads
  size_t n_rows = 0;        // selected rows
  uint32_t max_rid = 0;
  uint64_t n_words = 0;     // packed words incl. guards
  uint64_t sel_bases = 0;
  uint8_t *d_raw = nullptr; size_t raw_bytes = 0;
  size_t raw_ring = 0;  // != 0: d_raw is a ring of two staging windows of this many bytes, not the whole image (keep_raw = 0)
  uint64_t *d_w = nullptr; uint32_t *d_nm = nullptr;
  // deferred bulk copy (pgb_load_reads with PGB_LOAD_DEFER): pgb_index overlaps the host->device copy of the .seqdb image
  // with packing and sketching, chunk by chunk; any other consumer of the reads completes the load first (ensure_loaded)
  struct { const uint8_t *src = nullptr; bool active = false, keep_raw = false;
           const uint64_t *src_words = nullptr; bool packed = false; } pend;  // packed: src_words is the 2-bit image of the selected rows
  cudaStream_t st_copy = nullptr;
  std::vector<cudaEvent_t> ev_pool;
  uint64_t *d_wrc = nullptr;   // reverse-complement image of d_w (built on first use by the overlap / ovlp_match paths)
  uint32_t n_reads_with_n = 0;  // valid once d_wrc exists
  uint32_t *d_rlen_by_rid = nullptr, *d_hasn_by_rid = nullptr; uint64_t *d_woff_by_rid = nullptr;
  uint32_t *d_row_rid = nullptr, *d_row_len = nullptr; uint64_t *d_row_woff = nullptr, *d_row_raw_off = nullptr;
  uint32_t *d_sel_rows = nullptr;  // identity 0..n_rows-1 (rows are already the selected ones)
  std::vector<uint32_t> h_row_len;
  // ---- index levels
  mm128 *d_level[3] = {nullptr, nullptr, nullptr};
  uint64_t *d_level_off[3] = {nullptr, nullptr, nullptr};
  size_t level_n[3] = {0, 0, 0};
  std::vector<mc_entry> level_mc[3];
  // ---- overlap inputs
  mm128 *d_shm = nullptr; size_t n_shm = 0; bool shm_owned = false;
  uint64_t *d_mckeys = nullptr; uint32_t *d_mcvals = nullptr; uint32_t mcmask = 0;
  // ---- routed exchange (multi-GPU): dumped partial count table, per-mmer counts of the scan, records grouped by owner chunk
  mc_entry *d_mc_dump = nullptr; size_t n_mc_dump = 0;
  uint32_t *d_route_cnt = nullptr; unsigned long long route_first = ~0ULL;
  void *d_route = nullptr; size_t n_route = 0;
  // ---- map output (shmr_map text)
  char *d_map_text = nullptr; size_t map_bytes = 0, map_hits = 0;
  // ---- dedup output (preads.ovl text)
  char *d_dedup_text = nullptr; size_t dedup_bytes = 0, dedup_kept = 0;
  // pair table of a record stream that is deduplicated in batches (pgb_dedup_stream_*): persists between the pushes
  uint64_t *dd_keys = nullptr; unsigned long long *dd_first = nullptr; uint32_t dd_cap = 0; unsigned long long dd_base = 0, dd_occ = 0;
  bool dd_open = false;
  // ---- replay table sizing relative to the eligible record count (learned: doubled whenever a table overflowed)
  double ecap_ratio = 1.25, acap_ratio = 0.75;
  // ---- overlap output
  ovlp_rec *d_ovl = nullptr; size_t n_ovl = 0;
  void *h_ovl = nullptr; size_t h_ovl_cap = 0;  // page-locked staging of the records (pgb_overlap_host)
  uint64_t *h_okey = nullptr; size_t h_okey_cap = 0;  // page-locked staging of the outer khash keys (overlap_core)
  int *d_err = nullptr;
  KhashEmu outer_emu;  // host replay of the outer khash (visiting order), storage reused between calls

  // persistent device memory (reads, index levels, overlap output).  Blocks are recycled through a per-context cache: a
  // steady-state job asks for the same sizes every step, so after the first step no call reaches the driver's allocator
  // (the stream-ordered pool re-maps memory when its free list fragments, which showed up as 0.4-0.8 s host stalls).
  // Every user of a block runs on `st` (or, for d_raw, on st_copy fenced by events that `st` waits on) and every API call
  // ends with a synchronize of `st`, so handing a released block to the next palloc is safe.
  struct Blk { void *p; size_t bytes; };
  std::vector<Blk> blk_cache;
  std::unordered_map<void *, size_t> blk_live;
  size_t blk_cached_bytes = 0;
  void blk_flush() {
    if (blk_cache.empty()) return;
    cudaStreamSynchronize(st);
    for (auto &b : blk_cache) cudaFree(b.p);
    blk_cache.clear();
    blk_cached_bytes = 0;
  }
  void *blk_alloc(size_t bytes) {
    bytes = (bytes + 511) & ~(size_t)511;
    // size classes (8 per octave) above 1 MB: table sizes that follow data-dependent counts (e.g. the alignment cache, sized from a
    // pass's count of predicted alignments, which depends on the order in which concurrent buckets saw each other's entries)
    // differ by a few entries from step to step; without classes a request just above the cached block's size is a miss, and a
    // cudaMalloc in the middle of a step costs ~100 ms (bench.py e2e.device_mallocs shows them)
    if (bytes >= ((size_t)1 << 20)) {
      int lg = 63 - __builtin_clzll((unsigned long long)bytes);
      const size_t q = (size_t)1 << (lg - 3);
      bytes = (bytes + q - 1) & ~(q - 1);
    }
    int best = -1;
    for (int i = 0; i < (int)blk_cache.size(); i++)
      if (blk_cache[i].bytes >= bytes && blk_cache[i].bytes <= bytes + bytes / 4 + (1 << 20) && (best < 0 || blk_cache[i].bytes < blk_cache[best].bytes))
        best = i;
    void *p = nullptr;
    if (best >= 0) {
      p = blk_cache[best].p;
      bytes = blk_cache[best].bytes;
      blk_cached_bytes -= bytes;
      blk_cache[best] = blk_cache.back();
      blk_cache.pop_back();
    } else {
      // a miss means the job changed shape: stale blocks would only pile up
      if (blk_cached_bytes > ((size_t)8 << 30) || blk_cache.size() > 256) blk_flush();
      stats.n_device_mallocs++;
      if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        blk_flush();
        CU(cudaMalloc(&p, bytes));
      }
    }
    blk_live[p] = bytes;
    return p;
  }
  void blk_release(void *p) {
    auto it = blk_live.find(p);
    if (it == blk_live.end()) { cudaFreeAsync(p, st); return; }
    blk_cache.push_back(Blk{p, it->second});
    blk_cached_bytes += it->second;
    blk_live.erase(it);
  }
  template <class T> T *palloc(size_t n) { return (T *)blk_alloc((n ? n : 1) * sizeof(T)); }
  // stage temporaries: bump allocation out of slabs that are kept for the lifetime of the context, reset at the end of every
  // API call (no per-step cudaMalloc/cudaFree traffic once the slabs have reached their steady-state size)
  struct Slab { char *base; size_t cap, used; };
  std::vector<Slab> slabs;
  template <class T> T *alloc(size_t n) {
    size_t bytes = ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
    for (auto &sl : slabs)
      if (sl.cap - sl.used >= bytes) { T *p = (T *)(sl.base + sl.used); sl.used += bytes; return p; }
    size_t cap = std::max(bytes, (size_t)512 << 20);
    void *b = nullptr;
    stats.n_device_mallocs++;
    CU(cudaMalloc(&b, cap));
    slabs.push_back(Slab{(char *)b, cap, bytes});
    return (T *)b;
  }
  bool in_scratch(const void *p) const {
    for (auto &sl : slabs)
      if ((const char *)p >= sl.base && (const char *)p < sl.base + sl.cap) return true;
    return false;
  }
  void scratch_reset() {
    // coalesce into one slab once the total demand of a call is known, so later calls never allocate again
    size_t total = 0, used = 0;
    for (auto &sl : slabs) { total += sl.cap; used += sl.used; }
    if (slabs.size() > 1) {
      cudaStreamSynchronize(st);
      for (auto &sl : slabs) cudaFree(sl.base);
      slabs.clear();
      void *b = nullptr;
      size_t cap = used + (used >> 3) + ((size_t)64 << 20);
      if (cudaMalloc(&b, cap) == cudaSuccess) slabs.push_back(Slab{(char *)b, cap, 0});
    }
    for (auto &sl : slabs) sl.used = 0;
    (void)total;
  }
  void scratch_free() {
    for (auto &sl : slabs) cudaFree(sl.base);
    slabs.clear();
  }
  template <class T> void release(T *&p) {
    if (p && !in_scratch((const void *)p)) blk_release((void *)p);
    p = nullptr;
  }
  void sync() { CU(cudaStreamSynchronize(st)); }
  void tic() { CU(cudaEventRecord(ev0, st)); }
  double toc() {
    CU(cudaEventRecord(ev1, st));
    CU(cudaEventSynchronize(ev1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev0, ev1));
    return ms;
  }
  void ktic() { CU(cudaEventRecord(evk0, st)); }
  double ktoc() {
    CU(cudaEventRecord(evk1, st));
    CU(cudaEventSynchronize(evk1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, evk0, evk1));
    return ms;
  }
  void h2d(void *d, const void *h, size_t bytes) {
    if (!bytes) return;
    CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
    stats.h2d_bytes += bytes;
  }
  void d2h(void *h, const void *d, size_t bytes) {
    if (!bytes) return;
    CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    stats.d2h_bytes += bytes;
  }
  int check_err(const char *where) {
    int e = 0;
    d2h(&e, d_err, sizeof e);
    if (e) {
      char buf[256];
      snprintf(buf, sizeof buf, "device error flag 0x%x in %s (4=count table full, 8=mer missing from count table, 16=bucket table full, "
               "32=pair table full, 64=alignment queue full, 128=ovlp_match band state, 256=dedup pair table)", e, where);
      err = buf;
      int z = 0;
      h2d(d_err, &z, sizeof z);
      sync();
    }
    return e;
  }
  void free_reads() {
    pend.active = false;
    release(d_raw); release(d_w); release(d_nm); release(d_wrc); release(d_rlen_by_rid); release(d_hasn_by_rid); release(d_woff_by_rid);
    release(d_row_rid); release(d_row_len); release(d_row_woff); release(d_row_raw_off); release(d_sel_rows);
    n_rows = 0;
  }
  void free_index() {
    for (int l = 0; l < 3; l++) {
      if (d_shm == d_level[l]) { d_shm = nullptr; n_shm = 0; }
      release(d_level[l]); release(d_level_off[l]); level_n[l] = 0; level_mc[l].clear();
    }
  }
  void free_shimmers() {
    if (shm_owned) release(d_shm);
    d_shm = nullptr; n_shm = 0; shm_owned = false;
    release(d_mckeys); release(d_mcvals);
    release(d_mc_dump); n_mc_dump = 0;
    release(d_route_cnt);
    release(d_route); n_route = 0;
  }
};

#define LAUNCH_SMEM(ctx, kern, grid, block, smem, ...)           \
  do {                                                           \
    if ((grid) > 0) {                                            \
      kern<<<(grid), (block), (smem), (ctx)->st>>>(__VA_ARGS__); \
      (ctx)->stats.kernel_launches++;                            \
      CU(cudaGetLastError());                                    \
    }                                                            \
  } while (0)
#define LAUNCH(ctx, kern, grid, block, ...)                      \
  do {                                                           \
    if ((grid) > 0) {                                            \
      kern<<<(grid), (block), 0, (ctx)->st>>>(__VA_ARGS__);      \
      (ctx)->stats.kernel_launches++;                            \
      CU(cudaGetLastError());                                    \
    }                                                            \
  } while (0)

// exclusive scan of n+1 u32 (last element must be 0 on input) -> u64 offsets; returns total
static uint64_t scan_u32_to_u64(pgb_ctx *c, const uint32_t *d_in, uint64_t *d_out, size_t n_plus1) {
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveScan((void *)nullptr, tmp_bytes, d_in, d_out, cub::Sum(), (uint64_t)0, (int)n_plus1, c->st));
  uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
  CU(cub::DeviceScan::ExclusiveScan((void *)tmp, tmp_bytes, d_in, d_out, cub::Sum(), (uint64_t)0, (int)n_plus1, c->st));
  c->stats.kernel_launches += 2;
  c->release(tmp);
  uint64_t total = 0;
  c->d2h(&total, d_out + (n_plus1 - 1), sizeof total);
  return total;
}
static uint32_t scan_u32(pgb_ctx *c, const uint32_t *d_in, uint32_t *d_out, size_t n_plus1) {
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum((void *)nullptr, tmp_bytes, d_in, d_out, (int)n_plus1, c->st));
  uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
  CU(cub::DeviceScan::ExclusiveSum((void *)tmp, tmp_bytes, d_in, d_out, (int)n_plus1, c->st));
  c->stats.kernel_launches += 2;
  c->release(tmp);
  uint32_t total = 0;
  c->d2h(&total, d_out + (n_plus1 - 1), sizeof total);
  return total;
}

// stable LSD radix sort of (u32 key, u32 value) pairs (cub::DeviceRadixSort is stable)
static void sort_pairs_u32(pgb_ctx *c, const uint32_t *kin, uint32_t *kout, const uint32_t *vin, uint32_t *vout, uint32_t n) {
  if (!n) return;
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, tmp_bytes, kin, kout, vin, vout, (int)n, 0, 32, c->st));
  uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
  CU(cub::DeviceRadixSort::SortPairs((void *)tmp, tmp_bytes, kin, kout, vin, vout, (int)n, 0, 32, c->st));
  c->stats.kernel_launches += 9;  // histogram + 4 x (scan, downsweep) passes of 8 bits
  c->release(tmp);
}

// ================================================================================================ context
extern "C" int pgb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}
extern "C" pgb_ctx *pgb_create(int device) {
  int n = pgb_device_count();
  if (n <= 0) {
    fprintf(stderr, "pgb200: no CUDA device available; this library has no CPU path\n");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    fprintf(stderr, "pgb200: device %d out of range (have %d)\n", device, n);
    return nullptr;
  }
  try {
    CU(cudaSetDevice(device));
    pgb_ctx *c = new pgb_ctx();
    c->device = device;
    if (cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || c->sm_count <= 0) c->sm_count = 148;
    memset(&c->stats, 0, sizeof c->stats);
    CU(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreate(&c->evk0));
    CU(cudaEventCreate(&c->evk1));
    for (int i = 0; i < 8; i++) CU(cudaEventCreate(&c->user_ev[i]));
    for (int i = 0; i < 6; i++) CU(cudaEventCreate(&c->ev_pass[i]));
    CU(cudaMallocHost((void **)&c->h_pass, sizeof(PassOut)));
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = UINT64_MAX;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    c->d_err = c->palloc<int>(1);
    CU(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->st));
    c->d_align_bases = c->palloc<unsigned long long>(1);
    CU(cudaMemsetAsync(c->d_align_bases, 0, 8, c->st));
    c->sync();
    return c;
  } catch (std::exception &e) {
    fprintf(stderr, "pgb200: pgb_create failed: %s\n", e.what());
    return nullptr;
  }
}
extern "C" void pgb_destroy(pgb_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  c->free_shimmers();
  c->free_index();
  c->free_reads();
  c->release(c->d_ovl);
  c->release(c->d_dedup_text);
  c->release(c->dd_keys); c->release(c->dd_first);
  c->release(c->d_map_text);
  if (c->h_ovl) cudaFreeHost(c->h_ovl);
  if (c->h_okey) cudaFreeHost(c->h_okey);
  c->release(c->d_err);
  c->release(c->d_align_bases);
  cudaStreamSynchronize(c->st);
  c->blk_flush();
  c->scratch_free();
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaEventDestroy(c->evk0);
  cudaEventDestroy(c->evk1);
  for (int i = 0; i < 8; i++) cudaEventDestroy(c->user_ev[i]);
  for (int i = 0; i < 6; i++) cudaEventDestroy(c->ev_pass[i]);
  if (c->h_pass) cudaFreeHost(c->h_pass);
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaStreamDestroy(c->st_copy);
  cudaStreamDestroy(c->st);
  delete c;
}
extern "C" const char *pgb_last_error(pgb_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" void pgb_stats_reset(pgb_ctx *c) { memset(&c->stats, 0, sizeof c->stats); }
extern "C" void pgb_stats_get(pgb_ctx *c, pgb_stats *out) { *out = c->stats; }
extern "C" int pgb_event_record(pgb_ctx *c, int slot) {
  if (!c || slot < 0 || slot > 7) return -1;
  cudaSetDevice(c->device);
  return cudaEventRecord(c->user_ev[slot], c->st) == cudaSuccess ? 0 : -1;
}
extern "C" double pgb_event_elapsed_ms(pgb_ctx *c, int a, int b) {
  if (!c || a < 0 || a > 7 || b < 0 || b > 7) return -1.0;
  cudaSetDevice(c->device);
  float ms = -1.f;
  if (cudaEventSynchronize(c->user_ev[b]) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, c->user_ev[a], c->user_ev[b]) != cudaSuccess) return -1.0;
  return ms;
}

#define API_BEGIN(c)          \
  if (!(c)) return -1;        \
  (c)->err.clear();           \
  try {                       \
    CU(cudaSetDevice((c)->device));
#define API_END(c)                \
  }                               \
  catch (std::exception & e) {    \
    (c)->err = e.what();          \
    cudaStreamSynchronize((c)->st); \
    (c)->scratch_reset();         \
    return -1;                    \
  }                               \
  cudaStreamSynchronize((c)->st); \
  (c)->scratch_reset();           \
  return (c)->err.empty() ? 0 : -1;

// ================================================================================================ reads
static void do_pack(pgb_ctx *c) {
  c->tic();
  uint64_t body = c->n_words - 4;  // words [2, n_words-2)
  CU(cudaMemsetAsync(c->d_w, 0, c->n_words * 8, c->st));
  CU(cudaMemsetAsync(c->d_nm, 0, c->n_words * 4, c->st));
  CU(cudaMemsetAsync(c->d_hasn_by_rid, 0, ((size_t)c->max_rid + 1) * 4, c->st));
  if (body && c->n_rows)
    LAUNCH(c, k_pack_reads, nblk(body), 256, c->d_raw, c->d_row_raw_off, c->d_row_len, c->d_row_woff, c->d_row_rid,
           (uint32_t)c->n_rows, (uint64_t)2, body, c->d_w, c->d_nm, c->d_hasn_by_rid);
  c->stats.ms_pack += c->toc();
  c->stats.bases_packed += c->sel_bases;
}

// reverse-complement image + number of reads with N, built once per loaded read set
static void ensure_loaded(pgb_ctx *c);
// rc_launch only enqueues the kernels (returns the device counter, or null when the image exists); rc_finish reads the counter.
// overlap_core puts the host's khash replay between the two, so that the image is built while the CPU is busy.
static unsigned int *rc_launch(pgb_ctx *c) {
  if (c->d_wrc || !c->d_w) return nullptr;
  c->d_wrc = c->palloc<uint64_t>(c->n_words);
  CU(cudaMemsetAsync(c->d_wrc, 0, c->n_words * 8, c->st));
  uint64_t body = c->n_words - 4;
  if (body && c->n_rows) LAUNCH(c, k_make_rc, nblk(body), 256, c->d_w, c->d_row_len, c->d_row_woff, (uint32_t)c->n_rows, (uint64_t)2, body, c->d_wrc);
  unsigned int *d_n = c->alloc<unsigned int>(1);
  CU(cudaMemsetAsync(d_n, 0, 4, c->st));
  LAUNCH(c, k_count_nonzero_u32, nblk((size_t)c->max_rid + 1), 256, c->d_hasn_by_rid, (size_t)c->max_rid + 1, d_n);
  return d_n;
}
static void rc_finish(pgb_ctx *c, unsigned int *d_n) {
  if (!d_n) return;
  unsigned int h = 0;
  c->d2h(&h, d_n, 4);
  c->n_reads_with_n = h;
  c->release(d_n);
}
static void ensure_rc(pgb_ctx *c) {
  ensure_loaded(c);
  c->tic();
  rc_finish(c, rc_launch(c));
  c->stats.ms_pack += c->toc();
}

// completes a deferred bulk copy (no overlap with compute): whoever needs the packed reads before pgb_index ran
static void ensure_loaded(pgb_ctx *c) {
  if (!c->pend.active) return;
  if (c->pend.packed) {  // 2-bit image: nothing to pack
    const uint64_t body = c->n_words - 4;
    const size_t CHW = (size_t)32 << 20;
    for (size_t o = 0; o < body; o += CHW) c->h2d(c->d_w + 2 + o, c->pend.src_words + o, std::min(CHW, (size_t)body - o) * 8);
    c->sync();
    c->pend.active = false; c->pend.packed = false;
    c->stats.bases_packed += c->sel_bases;
    return;
  }
  if (c->raw_ring) {  // window by window through the ring (see pgb_load_reads)
    CU(cudaMemsetAsync(c->d_w, 0, c->n_words * 8, c->st));
    CU(cudaMemsetAsync(c->d_nm, 0, c->n_words * 4, c->st));
    CU(cudaMemsetAsync(c->d_hasn_by_rid, 0, ((size_t)c->max_rid + 1) * 4, c->st));
    uint64_t raw_o = 0, word_o = 2;
    size_t r0 = 0;
    int half = 0;
    const size_t ns = c->n_rows;
    while (r0 < ns) {
      size_t r1 = r0;
      uint64_t bytes = 0, wcount = 0;
      while (r1 < ns && bytes + c->h_row_len[r1] <= c->raw_ring) { bytes += c->h_row_len[r1]; wcount += ((uint64_t)c->h_row_len[r1] + 31) / 32; r1++; }
      uint8_t *win = c->d_raw + (size_t)half * c->raw_ring;
      c->h2d(win, c->pend.src + raw_o, bytes);
      if (wcount)
        LAUNCH(c, k_pack_reads, nblk(wcount), 256, win, c->d_row_raw_off, c->d_row_len, c->d_row_woff, c->d_row_rid, (uint32_t)c->n_rows, word_o, wcount, c->d_w,
               c->d_nm, c->d_hasn_by_rid, raw_o);
      raw_o += bytes; word_o += wcount; r0 = r1; half ^= 1;
    }
    c->stats.bases_packed += c->sel_bases;
    c->sync();
    c->pend.active = false;
    c->release(c->d_raw);
    return;
  }
  const size_t CH = (size_t)256 << 20;
  for (size_t o = 0; o < c->raw_bytes; o += CH) c->h2d(c->d_raw + o, c->pend.src + o, std::min(CH, c->raw_bytes - o));
  do_pack(c);
  c->sync();
  c->pend.active = false;
  if (!c->pend.keep_raw) c->release(c->d_raw);
}

extern "C" int pgb_load_reads(pgb_ctx *c, const uint8_t *seqdb, size_t seqdb_bytes, const uint32_t *rid, const uint32_t *len,
                              const uint64_t *offset, size_t n_reads, uint32_t T, uint32_t mychunk, int keep_raw) {
  API_BEGIN(c)
  if (T == 0 || mychunk == 0 || mychunk > T) throw std::runtime_error("bad chunk spec");
  c->free_reads();
  const bool defer = (keep_raw & PGB_LOAD_DEFER) != 0;
  keep_raw &= 1;
  // the by-rid tables cover every read (lengths are needed for any rid an index file mentions)
  uint32_t max_rid = 0;
  for (size_t i = 0; i < n_reads; i++) if (rid[i] > max_rid) max_rid = rid[i];
  if (n_reads && (uint64_t)max_rid > 64 * (uint64_t)n_reads + (1u << 24)) throw std::runtime_error("read ids too sparse");
  c->max_rid = max_rid;
  std::vector<uint32_t> rows;
  for (size_t i = 0; i < n_reads; i++) {
    if (offset[i] + len[i] > seqdb_bytes) throw std::runtime_error("read extends past the end of the seqdb image");
    if (rid[i] % T == mychunk % T) rows.push_back((uint32_t)i);
  }
  size_t nsel = rows.size();
  std::vector<uint32_t> h_rid(nsel), h_len(nsel), h_rlen_by_rid((size_t)max_rid + 1, 0);
  std::vector<uint64_t> h_woff(nsel), h_rawoff(nsel), h_woff_by_rid((size_t)max_rid + 1, 0);
  for (size_t i = 0; i < n_reads; i++) h_rlen_by_rid[rid[i]] = len[i];
  uint64_t words = 2, raw = 0;
  bool contiguous = true;
  for (size_t j = 0; j < nsel; j++) {
    uint32_t i = rows[j];
    h_rid[j] = rid[i]; h_len[j] = len[i]; h_woff[j] = words; h_rawoff[j] = raw;
    h_woff_by_rid[rid[i]] = words;
    if (j && offset[i] != offset[rows[j - 1]] + len[rows[j - 1]]) contiguous = false;
    words += ((uint64_t)len[i] + 31) / 32;
    raw += len[i];
  }
  words += 2;
  c->n_rows = nsel; c->n_words = words; c->sel_bases = raw; c->raw_bytes = raw;
  c->h_row_len = h_len;
  // keep_raw = 0 and one contiguous source range: the 1-byte/base image never sits whole in HBM, it passes through a ring of two
  // staging windows (the packed form is 3/8 byte per base; at 90 Gbase the whole image would be 90 GB of a 180 GB device)
  uint32_t longest = 0;
  for (size_t j = 0; j < nsel; j++) longest = std::max(longest, h_len[j]);
  const size_t RING = std::max<size_t>((size_t)256 << 20, (size_t)longest + 64);
  c->raw_ring = (!keep_raw && contiguous && raw > 2 * RING) ? RING : 0;
  c->d_raw = c->palloc<uint8_t>((c->raw_ring ? 2 * c->raw_ring : raw) + 64);
  c->d_w = c->palloc<uint64_t>(words); c->d_nm = c->palloc<uint32_t>(words);
  c->d_rlen_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1); c->d_hasn_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1);
  c->d_woff_by_rid = c->palloc<uint64_t>((size_t)max_rid + 1);
  c->d_row_rid = c->palloc<uint32_t>(nsel); c->d_row_len = c->palloc<uint32_t>(nsel);
  c->d_row_woff = c->palloc<uint64_t>(nsel); c->d_row_raw_off = c->palloc<uint64_t>(nsel); c->d_sel_rows = c->palloc<uint32_t>(nsel);
  std::vector<uint32_t> ident(nsel);
  for (size_t j = 0; j < nsel; j++) ident[j] = (uint32_t)j;
  c->h2d(c->d_rlen_by_rid, h_rlen_by_rid.data(), h_rlen_by_rid.size() * 4);
  c->h2d(c->d_woff_by_rid, h_woff_by_rid.data(), h_woff_by_rid.size() * 8);
  c->h2d(c->d_row_rid, h_rid.data(), nsel * 4); c->h2d(c->d_row_len, h_len.data(), nsel * 4);
  c->h2d(c->d_row_woff, h_woff.data(), nsel * 8); c->h2d(c->d_row_raw_off, h_rawoff.data(), nsel * 8);
  c->h2d(c->d_sel_rows, ident.data(), nsel * 4);
  // raw bytes of the selected reads
  bool packed_already = false;
  if (nsel) {
    if (contiguous && defer) {
      c->pend.src = seqdb + offset[rows[0]];
      c->pend.active = true;
      c->pend.keep_raw = keep_raw != 0;
      c->sync();
      return 0;  // (API_END's bookkeeping is not needed: nothing was taken from the scratch arena)
    } else if (contiguous && c->raw_ring) {
      // window by window: whole reads up to the window size, copy into one half of the ring, pack from there
      const uint8_t *src = seqdb + offset[rows[0]];
      CU(cudaMemsetAsync(c->d_w, 0, c->n_words * 8, c->st));
      CU(cudaMemsetAsync(c->d_nm, 0, c->n_words * 4, c->st));
      CU(cudaMemsetAsync(c->d_hasn_by_rid, 0, ((size_t)c->max_rid + 1) * 4, c->st));
      c->tic();
      uint64_t raw_o = 0, word_o = 2;
      size_t r0 = 0;
      int half = 0;
      while (r0 < nsel) {
        size_t r1 = r0;
        uint64_t bytes = 0, wcount = 0;
        while (r1 < nsel && bytes + h_len[r1] <= c->raw_ring) { bytes += h_len[r1]; wcount += ((uint64_t)h_len[r1] + 31) / 32; r1++; }
        uint8_t *win = c->d_raw + (size_t)half * c->raw_ring;
        c->h2d(win, src + raw_o, bytes);  // (same stream as the pack kernels: a window is reused only after its pack finished)
        if (wcount)
          LAUNCH(c, k_pack_reads, nblk(wcount), 256, win, c->d_row_raw_off, c->d_row_len, c->d_row_woff, c->d_row_rid, (uint32_t)c->n_rows, word_o, wcount,
                 c->d_w, c->d_nm, c->d_hasn_by_rid, raw_o);
        raw_o += bytes; word_o += wcount; r0 = r1; half ^= 1;
      }
      c->stats.ms_pack += c->toc();
      c->stats.bases_packed += c->sel_bases;
      c->sync();
      c->release(c->d_raw);
      packed_already = true;
    } else if (contiguous) {
      const uint8_t *src = seqdb + offset[rows[0]];
      const size_t CH = (size_t)256 << 20;
      for (size_t o = 0; o < raw; o += CH) c->h2d(c->d_raw + o, src + o, std::min(CH, raw - o));
    } else {
      // gather through two pinned staging buffers
      const size_t CH = (size_t)64 << 20;
      uint8_t *stage[2] = {nullptr, nullptr};
      cudaEvent_t done[2];
      for (int b = 0; b < 2; b++) { CU(cudaMallocHost((void **)&stage[b], CH)); CU(cudaEventCreate(&done[b])); }
      size_t j = 0, dev_o = 0; int b = 0; uint32_t part = 0;
      while (j < nsel) {
        CU(cudaEventSynchronize(done[b]));
        size_t fill = 0;
        while (j < nsel && fill < CH) {
          uint32_t i = rows[j];
          size_t take = std::min((size_t)len[i] - part, CH - fill);
          memcpy(stage[b] + fill, seqdb + offset[i] + part, take);
          fill += take; part += (uint32_t)take;
          if (part == len[i]) { part = 0; j++; }
        }
        c->h2d(c->d_raw + dev_o, stage[b], fill);
        CU(cudaEventRecord(done[b], c->st));
        dev_o += fill; b ^= 1;
      }
      c->sync();
      for (int q = 0; q < 2; q++) { cudaFreeHost(stage[q]); cudaEventDestroy(done[q]); }
    }
  }
  if (!packed_already) {
    do_pack(c);
    c->sync();
    if (!keep_raw) c->release(c->d_raw);
  }
  API_END(c)
}

// The 2-bit hand-off ("ideally emitting the 2-bit device image alongside").  pgb_pack_2bit turns .seqdb bytes into
// the packed form on the device and returns it to the host: per read ceil(len/32) words, reads in the order given, plus the
// N mask words (all zero for a read without N) and a per-read flag.  shmr_mkseqdb writes these next to the .seqdb
// (<prefix>.seq2b = the words, <prefix>.seq2n = {read index, word count, mask words} for the reads that contain N).
extern "C" int pgb_pack_2bit(pgb_ctx *c, const uint8_t *seqdb, size_t seqdb_bytes, const uint64_t *offset, const uint32_t *len, size_t n_reads,
                             uint64_t *words_out, uint32_t *nmask_out, uint8_t *hasn_out) {
  API_BEGIN(c)
  if (n_reads >= (1ull << 31)) throw std::runtime_error("too many reads in one pgb_pack_2bit call");
  std::vector<uint64_t> h_woff(n_reads), h_raw(n_reads);
  std::vector<uint32_t> h_rid(n_reads);
  uint64_t words = 0, lo = ~0ULL, hi = 0;
  for (size_t i = 0; i < n_reads; i++) {
    if (offset[i] + len[i] > seqdb_bytes) throw std::runtime_error("read extends past the end of the seqdb buffer");
    h_woff[i] = words; h_rid[i] = (uint32_t)i;
    words += ((uint64_t)len[i] + 31) / 32;
    lo = std::min(lo, offset[i]); hi = std::max(hi, offset[i] + len[i]);
  }
  if (n_reads && words) {
    for (size_t i = 0; i < n_reads; i++) h_raw[i] = offset[i] - lo;
    uint8_t *d_raw = c->alloc<uint8_t>(hi - lo + 64);
    uint64_t *d_w = c->alloc<uint64_t>(words), *d_woff = c->alloc<uint64_t>(n_reads), *d_rawoff = c->alloc<uint64_t>(n_reads);
    uint32_t *d_nm = c->alloc<uint32_t>(words), *d_len = c->alloc<uint32_t>(n_reads), *d_rid = c->alloc<uint32_t>(n_reads), *d_hasn = c->alloc<uint32_t>(n_reads);
    c->tic();
    c->h2d(d_raw, seqdb + lo, hi - lo);
    c->h2d(d_woff, h_woff.data(), n_reads * 8); c->h2d(d_rawoff, h_raw.data(), n_reads * 8);
    c->h2d(d_len, len, n_reads * 4); c->h2d(d_rid, h_rid.data(), n_reads * 4);
    CU(cudaMemsetAsync(d_hasn, 0, n_reads * 4, c->st));
    LAUNCH(c, k_pack_reads, nblk(words), 256, d_raw, d_rawoff, d_len, d_woff, d_rid, (uint32_t)n_reads, (uint64_t)0, words, d_w, d_nm, d_hasn);
    c->d2h(words_out, d_w, words * 8);
    c->d2h(nmask_out, d_nm, words * 4);
    std::vector<uint32_t> h_hasn(n_reads);
    c->d2h(h_hasn.data(), d_hasn, n_reads * 4);
    for (size_t i = 0; i < n_reads; i++) hasn_out[i] = h_hasn[i] != 0;
    c->stats.ms_pack += c->toc();
    c->stats.bases_packed += hi - lo;
  }
  API_END(c)
}

// Read set from the 2-bit image (what pgb_pack_2bit produced for ALL reads of the .idx table, in table order): rows with
// rid % total_chunk == mychunk % total_chunk are copied to the device as they are - a quarter of the .seqdb bytes over the
// bus and no packing kernel.  n_records: the .seq2n stream (u32: read index, word count, mask words; ...) or NULL.
// flags: PGB_LOAD_DEFER as for pgb_load_reads (the words must stay valid until the next pgb_index returns).
extern "C" int pgb_load_reads_2bit(pgb_ctx *c, const uint64_t *words, size_t n_words_total, const uint32_t *n_records, size_t n_record_words,
                                   const uint32_t *rid, const uint32_t *len, size_t n_reads, uint32_t T, uint32_t mychunk, int flags) {
  API_BEGIN(c)
  if (T == 0 || mychunk == 0 || mychunk > T) throw std::runtime_error("bad chunk spec");
  c->free_reads();
  const bool defer = (flags & PGB_LOAD_DEFER) != 0;
  uint32_t max_rid = 0;
  for (size_t i = 0; i < n_reads; i++) if (rid[i] > max_rid) max_rid = rid[i];
  if (n_reads && (uint64_t)max_rid > 64 * (uint64_t)n_reads + (1u << 24)) throw std::runtime_error("read ids too sparse");
  c->max_rid = max_rid;
  std::vector<uint64_t> src_woff(n_reads + 1);
  uint64_t acc = 0;
  for (size_t i = 0; i < n_reads; i++) { src_woff[i] = acc; acc += ((uint64_t)len[i] + 31) / 32; }
  src_woff[n_reads] = acc;
  if (acc != n_words_total) throw std::runtime_error("the 2-bit image does not match the read table (stale .seq2b?)");
  std::vector<uint32_t> rows;
  for (size_t i = 0; i < n_reads; i++) if (rid[i] % T == mychunk % T) rows.push_back((uint32_t)i);
  const size_t nsel = rows.size();
  std::vector<uint32_t> h_rid(nsel), h_len(nsel), h_rlen_by_rid((size_t)max_rid + 1, 0), h_hasn_by_rid((size_t)max_rid + 1, 0);
  std::vector<uint64_t> h_woff(nsel), h_woff_by_rid((size_t)max_rid + 1, 0);
  std::vector<int64_t> row_of_read(n_reads, -1);
  for (size_t i = 0; i < n_reads; i++) h_rlen_by_rid[rid[i]] = len[i];
  uint64_t wsum = 2, bases = 0;
  bool contiguous = true;
  for (size_t j = 0; j < nsel; j++) {
    const uint32_t i = rows[j];
    h_rid[j] = rid[i]; h_len[j] = len[i]; h_woff[j] = wsum;
    h_woff_by_rid[rid[i]] = wsum;
    row_of_read[i] = (int64_t)j;
    if (j && rows[j] != rows[j - 1] + 1) contiguous = false;
    wsum += ((uint64_t)len[i] + 31) / 32;
    bases += len[i];
  }
  wsum += 2;
  c->n_rows = nsel; c->n_words = wsum; c->sel_bases = bases; c->raw_bytes = 0;
  c->h_row_len = h_len;
  c->d_w = c->palloc<uint64_t>(wsum); c->d_nm = c->palloc<uint32_t>(wsum);
  c->d_rlen_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1); c->d_hasn_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1);
  c->d_woff_by_rid = c->palloc<uint64_t>((size_t)max_rid + 1);
  c->d_row_rid = c->palloc<uint32_t>(nsel); c->d_row_len = c->palloc<uint32_t>(nsel);
  c->d_row_woff = c->palloc<uint64_t>(nsel); c->d_sel_rows = c->palloc<uint32_t>(nsel);
  std::vector<uint32_t> ident(nsel);
  for (size_t j = 0; j < nsel; j++) ident[j] = (uint32_t)j;
  c->tic();
  CU(cudaMemsetAsync(c->d_nm, 0, wsum * 4, c->st));
  CU(cudaMemsetAsync(c->d_w, 0, 2 * 8, c->st));
  CU(cudaMemsetAsync(c->d_w + (wsum - 2), 0, 2 * 8, c->st));
  // N masks of the selected reads that have one (sparse)
  for (size_t p = 0; n_records && p + 2 <= n_record_words;) {
    const uint32_t i = n_records[p], nw = n_records[p + 1];
    if (i >= n_reads || p + 2 + nw > n_record_words || nw != ((uint64_t)len[i] + 31) / 32) throw std::runtime_error("malformed N-mask stream (.seq2n)");
    if (row_of_read[i] >= 0) {
      c->h2d(c->d_nm + h_woff[(size_t)row_of_read[i]], n_records + p + 2, (size_t)nw * 4);
      h_hasn_by_rid[rid[i]] = 1;
    }
    p += 2 + (size_t)nw;
  }
  c->h2d(c->d_rlen_by_rid, h_rlen_by_rid.data(), h_rlen_by_rid.size() * 4);
  c->h2d(c->d_hasn_by_rid, h_hasn_by_rid.data(), h_hasn_by_rid.size() * 4);
  c->h2d(c->d_woff_by_rid, h_woff_by_rid.data(), h_woff_by_rid.size() * 8);
  c->h2d(c->d_row_rid, h_rid.data(), nsel * 4); c->h2d(c->d_row_len, h_len.data(), nsel * 4);
  c->h2d(c->d_row_woff, h_woff.data(), nsel * 8);
  c->h2d(c->d_sel_rows, ident.data(), nsel * 4);
  if (nsel) {
    if (contiguous && defer) {
      c->pend.src_words = words + src_woff[rows[0]];
      c->pend.active = true; c->pend.packed = true; c->pend.keep_raw = false;
    } else if (contiguous) {
      const size_t CHW = (size_t)32 << 20;
      const uint64_t *src = words + src_woff[rows[0]];
      const uint64_t body = wsum - 4;
      for (size_t o = 0; o < body; o += CHW) c->h2d(c->d_w + 2 + o, src + o, std::min(CHW, (size_t)body - o) * 8);
      c->stats.bases_packed += bases;
    } else {
      // gather the selected reads' word runs through two pinned staging buffers
      const size_t CHW = (size_t)8 << 20;
      uint64_t *stage[2] = {nullptr, nullptr};
      cudaEvent_t done[2];
      for (int b = 0; b < 2; b++) { CU(cudaMallocHost((void **)&stage[b], CHW * 8)); CU(cudaEventCreate(&done[b])); }
      size_t j = 0, dev_o = 2; int b = 0; uint64_t part = 0;
      while (j < nsel) {
        CU(cudaEventSynchronize(done[b]));
        size_t fill = 0;
        while (j < nsel && fill < CHW) {
          const uint32_t i = rows[j];
          const uint64_t nw = src_woff[i + 1] - src_woff[i];
          const size_t take = (size_t)std::min<uint64_t>(nw - part, CHW - fill);
          memcpy(stage[b] + fill, words + src_woff[i] + part, take * 8);
          fill += take; part += take;
          if (part == nw) { part = 0; j++; }
        }
        c->h2d(c->d_w + dev_o, stage[b], fill * 8);
        CU(cudaEventRecord(done[b], c->st));
        dev_o += fill; b ^= 1;
      }
      c->sync();
      for (int q = 0; q < 2; q++) { cudaFreeHost(stage[q]); cudaEventDestroy(done[q]); }
      c->stats.bases_packed += bases;
    }
  }
  c->sync();
  c->stats.ms_pack += c->toc();
  API_END(c)
}

extern "C" int pgb_repack(pgb_ctx *c) {
  API_BEGIN(c)
  ensure_loaded(c);
  if (!c->d_raw) throw std::runtime_error("pgb_repack: raw image was not kept (keep_raw = 0)");
  do_pack(c);
  c->sync();
  API_END(c)
}

struct MappedFile {
  const uint8_t *p = nullptr; size_t n = 0; int fd = -1;
  bool open_ro(const char *path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (fstat(fd, &sb) < 0) return false;
    n = (size_t)sb.st_size;
    if (n) {
      void *m = mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0);
      if (m == MAP_FAILED) return false;
      p = (const uint8_t *)m;
    }
    return true;
  }
  ~MappedFile() { if (p) munmap((void *)p, n); if (fd >= 0) ::close(fd); }
};

extern "C" int pgb_load_reads_from_files(pgb_ctx *c, const char *prefix, uint32_t T, uint32_t mychunk, int keep_raw) {
  if (!c) return -1;
  ReadTable rt;
  std::string idx = std::string(prefix) + ".idx", db = std::string(prefix) + ".seqdb";
  if (!load_read_table(idx.c_str(), &rt)) { c->err = "cannot open " + idx; return -1; }
  // the 2-bit side files written by this library's shmr_mkseqdb (<prefix>.seq2b / .seq2n), when they match the read table:
  // a quarter of the bytes to read and to move over the bus, no packing kernel.  PGB_NO_SEQ2B=1 ignores them.
  if (!getenv("PGB_NO_SEQ2B") && !(keep_raw & PGB_LOAD_KEEP_RAW)) {
    uint64_t want_words = 0;
    for (size_t i = 0; i < rt.n(); i++) want_words += ((uint64_t)rt.len[i] + 31) / 32;
    MappedFile m2, mn;
    struct stat s_db, s_2b;
    const std::string f2b = std::string(prefix) + ".seq2b", f2n = std::string(prefix) + ".seq2n";
    if (stat(f2b.c_str(), &s_2b) == 0 && stat(db.c_str(), &s_db) == 0 && s_2b.st_mtime >= s_db.st_mtime && (uint64_t)s_2b.st_size == want_words * 8 &&
        m2.open_ro(f2b.c_str()) && mn.open_ro(f2n.c_str()) && mn.n % 4 == 0) {
      return pgb_load_reads_2bit(c, (const uint64_t *)m2.p, (size_t)want_words, (const uint32_t *)mn.p, mn.n / 4, rt.rid.data(), rt.len.data(), rt.n(), T,
                                 mychunk, 0);  // (no deferred copy: the mappings end with this call)
    }
  }
  MappedFile mf;
  if (!mf.open_ro(db.c_str())) { c->err = "cannot open " + db; return -1; }
  return pgb_load_reads(c, mf.p, mf.n, rt.rid.data(), rt.len.data(), rt.off.data(), rt.n(), T, mychunk, keep_raw);
}

// ================================================================================================ index
static void build_level_counts(pgb_ctx *c, int level) {
  c->tic();
  size_t n = c->level_n[level];
  c->level_mc[level].clear();
  if (!n) { c->stats.ms_count += c->toc(); return; }
  uint32_t cap = pow2_at_least(2 * (uint64_t)n);
  uint64_t *keys = c->alloc<uint64_t>(cap);
  uint32_t *vals = c->alloc<uint32_t>(cap);
  uint32_t *flags = c->alloc<uint32_t>((size_t)cap + 1), *pos = c->alloc<uint32_t>((size_t)cap + 1);
  LAUNCH(c, k_fill_u64, 1184, 256, keys, PGB_EMPTY, (size_t)cap);
  CU(cudaMemsetAsync(vals, 0, (size_t)cap * 4, c->st));
  CU(cudaMemsetAsync(flags, 0, ((size_t)cap + 1) * 4, c->st));
  LAUNCH(c, k_mc_insert, nblk(n), 256, c->d_level[level], n, keys, vals, cap - 1, c->d_err);
  LAUNCH(c, k_mc_flags, nblk(cap), 256, keys, (size_t)cap, flags);
  uint32_t distinct = scan_u32(c, flags, pos, (size_t)cap + 1);
  mc_entry *out = c->alloc<mc_entry>(distinct);
  LAUNCH(c, k_mc_dump, nblk(cap), 256, keys, vals, pos, (size_t)cap, out);
  c->level_mc[level].resize(distinct);
  c->d2h(c->level_mc[level].data(), out, (size_t)distinct * sizeof(mc_entry));
  c->release(out); c->release(keys); c->release(vals); c->release(flags); c->release(pos);
  c->stats.ms_count += c->toc();
}

extern "C" int pgb_index(pgb_ctx *c, int w, int k, int r, int levels, int with_counts) {
  API_BEGIN(c)
  if (!(w > 0 && w < 256 && k > 0 && k <= 28)) throw std::runtime_error("mm_sketch needs 0<w<256 and 0<k<=28");  // mm_sketch.c:77
  if (!(r >= 1 && r < 256)) throw std::runtime_error("reduction factor must be in [1,255]");
  if (levels < 0 || levels > 2) throw std::runtime_error("levels must be 0, 1 or 2");
  if (!c->d_w) throw std::runtime_error("no reads loaded");
  c->free_index();
  size_t ns = c->n_rows;
  uint32_t *counts = c->alloc<uint32_t>(ns + 1);
  // ---- L0
  c->tic();
  CU(cudaMemsetAsync(counts, 0, (ns + 1) * 4, c->st));
  c->d_level_off[0] = c->palloc<uint64_t>(ns + 1);
  const char *force = getenv("PGB_SKETCH");
  const bool force_exact = force && !strcmp(force, "exact"), force_tiled = force && !strcmp(force, "tiled");
  // fast paths: the strip kernel (a warp walks a read, sketch_strip.cuh; default) or round 1's block-tiled kernel (PGB_SKETCH=tiled,
  // kept as the A/B baseline); tiny windows, k = 1 and PGB_SKETCH=exact run the exact automaton on every read
  const bool use_strip = w >= SK_MINW && k >= 2 && !force_exact && !force_tiled && ns > 0;
  const bool use_tiled = w >= SK_MINW && !force_exact && !use_strip && ns > 0;
  // the reads the fast kernel hands back (flags in row_flags) are redone by the exact automaton, segment-parallel; both fast
  // paths share this tail: counts[] holds the fast rows' record counts, exact_flag[] marks the others
  // (strip path only) row_bad / fast_cnt / tmp_off / tmp: reads of which only some strips are redone (SK_FLAG_PARTIAL, sketch_strip.cuh)
  auto exact_tail = [&](uint32_t *row_flags, uint32_t *exact_flag, uint32_t *exact_pos, const std::function<void()> &place_fast, const uint64_t *row_bad,
                        const uint32_t *fast_cnt, const uint64_t *tmp_off, const mm128 *tmp) {
    uint32_t n_exact = scan_u32(c, exact_flag, exact_pos, ns + 1);
    uint32_t *exact_list = nullptr, *seg_row = nullptr, *seg_lo = nullptr, *seg_hi = nullptr, *seg_first = nullptr, *list_first = nullptr, *seg_cnt = nullptr,
             *seg_pos = nullptr, *seg_src = nullptr;
    uint8_t *seg_kind = nullptr;
    uint32_t n_seg = 0, stage_cap = 0;
    mm128 *seg_stage = nullptr;
    bool staged_ok = false;
    // positions per thread of the exact automaton (plus its warm-up of ~2(w+k) and w trailing slots)
    const int SEG = getenv("PGB_EXACT_SEG") ? std::max(16, atoi(getenv("PGB_EXACT_SEG"))) : 256;
    // one warp per CTA: its w-slot rings (12 B per slot and thread) live in shared memory, 30 KB at w = 80
    const unsigned SEG_THREADS = 32;
    unsigned seg_stride = 1;  // lanes per piece (only the first of them works, see k_sketch_exact_seg): sized once the pieces are known
    size_t seg_smem = 0;
    if (n_exact) {
      exact_list = c->alloc<uint32_t>(n_exact);
      LAUNCH(c, k_compact_idx, nblk(ns), 256, exact_flag, exact_pos, ns, exact_list);
      std::vector<uint32_t> h_list(n_exact), h_seg_row, h_seg_lo, h_seg_hi, h_seg_first, h_list_first(n_exact + 1);
      std::vector<uint8_t> h_seg_kind;
      std::vector<uint64_t> h_bad;
      c->d2h(h_list.data(), exact_list, (size_t)n_exact * 4);
      if (row_bad) { h_bad.resize(ns); c->d2h(h_bad.data(), row_bad, ns * 8); }
      size_t n_partial = 0;
      auto push = [&](uint32_t row, uint32_t lo, uint32_t hi, uint8_t kind, uint32_t first) {
        h_seg_row.push_back(row); h_seg_lo.push_back(lo); h_seg_hi.push_back(hi); h_seg_kind.push_back(kind); h_seg_first.push_back(first);
      };
      for (uint32_t i = 0; i < n_exact; i++) {
        const uint32_t row = h_list[i], len = c->h_row_len[row];
        const uint32_t first = (uint32_t)h_seg_row.size();
        h_list_first[i] = first;
        const uint64_t bad = row_bad ? h_bad[row] : 0;
        if (!bad) {  // the whole read
          for (uint32_t lo = 0; lo < len; lo += SEG) push(row, lo, std::min<uint32_t>(lo + SEG, len), 0, first);
          continue;
        }
        // pieces in position order (build_sketch_pieces, host_util.hpp): bad strips and the w + SS_MAXPAL positions before them are
        // redone, everything between keeps the fast path's records
        n_partial++;
        std::vector<SketchPiece> pieces;
        build_sketch_pieces(len, bad, SS_STRIP, (uint32_t)w + SS_MAXPAL, (uint32_t)SEG, pieces);
        for (const SketchPiece &pc : pieces) push(row, pc.lo, pc.hi, pc.kind, first);
      }
      n_seg = (uint32_t)h_seg_row.size();
      h_list_first[n_exact] = n_seg;
      while (seg_stride < 32 && (size_t)n_seg * seg_stride * 2 <= 100000) seg_stride *= 2;  // about 100 k lanes fill the GPU
      if (getenv("PGB_EXACT_STRIDE")) seg_stride = (unsigned)std::min(32, std::max(1, atoi(getenv("PGB_EXACT_STRIDE"))));
      while (32 % seg_stride) seg_stride--;
      seg_smem = (size_t)w * (SEG_THREADS / seg_stride) * 12;
      if (getenv("PGB_VERBOSE")) fprintf(stderr, "pgb200: exact automaton: %u reads (%zu of them in part), %u pieces\n", n_exact, n_partial, n_seg);
      seg_row = c->alloc<uint32_t>(n_seg); seg_lo = c->alloc<uint32_t>(n_seg); seg_hi = c->alloc<uint32_t>(n_seg); seg_first = c->alloc<uint32_t>(n_seg);
      seg_kind = c->alloc<uint8_t>(n_seg); seg_src = c->alloc<uint32_t>(n_seg);
      list_first = c->alloc<uint32_t>((size_t)n_exact + 1); seg_cnt = c->alloc<uint32_t>((size_t)n_seg + 1); seg_pos = c->alloc<uint32_t>((size_t)n_seg + 1);
      c->h2d(seg_row, h_seg_row.data(), (size_t)n_seg * 4); c->h2d(seg_lo, h_seg_lo.data(), (size_t)n_seg * 4); c->h2d(seg_hi, h_seg_hi.data(), (size_t)n_seg * 4);
      c->h2d(seg_kind, h_seg_kind.data(), (size_t)n_seg);
      c->h2d(seg_first, h_seg_first.data(), (size_t)n_seg * 4); c->h2d(list_first, h_list_first.data(), ((size_t)n_exact + 1) * 4);
      CU(cudaMemsetAsync(seg_cnt, 0, ((size_t)n_seg + 1) * 4, c->st));
      CU(cudaMemsetAsync(seg_src, 0, (size_t)n_seg * 4, c->st));
      // ONE pass of the automaton: the records are staged per piece (budget: 4x the expected 2/(w+1) density, at least 32) and
      // placed once the counts are scanned; only a piece that overflows its budget costs the second pass
      stage_cap = std::max<uint32_t>(32, (uint32_t)(8 * SEG / (w + 1)) + 8);
      seg_stage = c->alloc<mm128>((size_t)n_seg * stage_cap);
      int *d_ovf = c->alloc<int>(1);
      CU(cudaMemsetAsync(d_ovf, 0, 4, c->st));
      c->ktic();
      CU(cudaFuncSetAttribute(k_sketch_exact_seg<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem));
      LAUNCH_SMEM(c, k_sketch_exact_seg<2>, nblk((size_t)n_seg * seg_stride, SEG_THREADS), SEG_THREADS, seg_smem, c->d_w, c->d_nm, seg_row, seg_lo, seg_hi, seg_kind, seg_first, n_seg,
                  c->d_row_rid, c->d_row_len, c->d_row_woff, c->d_hasn_by_rid, w, k, seg_cnt, (const uint32_t *)nullptr, (const uint64_t *)nullptr, seg_stage,
                  stage_cap, d_ovf, seg_stride);
      if (n_partial)
        LAUNCH(c, k_seg_fast_count, nblk(n_seg, 128), 128, seg_row, seg_lo, seg_hi, seg_kind, n_seg, tmp_off, tmp, fast_cnt, seg_cnt, seg_src);
      c->stats.ms_k_sketch_count += c->ktoc(); c->stats.n_k_sketch_count++;
      int h_ovf = 0;
      c->d2h(&h_ovf, d_ovf, 4);
      staged_ok = h_ovf == 0;
      c->release(d_ovf);
      scan_u32(c, seg_cnt, seg_pos, (size_t)n_seg + 1);
      LAUNCH(c, k_seg_row_counts, nblk(n_exact), 256, exact_list, list_first, n_exact, seg_pos, counts);
    }
    c->stats.n_sketch_fallback_reads += n_exact;
    c->level_n[0] = scan_u32_to_u64(c, counts, c->d_level_off[0], ns + 1);
    c->d_level[0] = c->palloc<mm128>(c->level_n[0]);
    place_fast();
    if (n_exact) {
      c->ktic();
      if (!staged_ok) {
        CU(cudaFuncSetAttribute(k_sketch_exact_seg<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem));
        LAUNCH_SMEM(c, k_sketch_exact_seg<1>, nblk((size_t)n_seg * seg_stride, SEG_THREADS), SEG_THREADS, seg_smem, c->d_w, c->d_nm, seg_row, seg_lo, seg_hi, seg_kind, seg_first,
                    n_seg, c->d_row_rid, c->d_row_len, c->d_row_woff, c->d_hasn_by_rid, w, k, (uint32_t *)nullptr, seg_pos, c->d_level_off[0], c->d_level[0], 0u,
                    (int *)nullptr, seg_stride);
      }
      LAUNCH(c, k_seg_place, nblk((size_t)n_seg * 8, 256), 256, seg_row, seg_first, seg_kind, seg_src, n_seg, seg_pos, c->d_level_off[0], seg_stage, stage_cap,
             tmp_off, tmp, c->d_level[0], staged_ok ? 1 : 0);
      c->stats.ms_k_sketch_write += c->ktoc(); c->stats.n_k_sketch_write++;
      c->release(seg_stage);
      c->release(exact_list); c->release(seg_row); c->release(seg_lo); c->release(seg_hi); c->release(seg_first); c->release(list_first); c->release(seg_cnt);
      c->release(seg_pos); c->release(seg_kind); c->release(seg_src);
    }
  };
  // host->device copy of the .seqdb image in chunks of whole reads on the copy stream while the compute stream packs and
  // sketches the previous chunk (reads are independent up to the final placement); sketch_rows(r0, r1) launches the fast kernel
  auto deferred_load = [&](const std::function<void(size_t, size_t)> &sketch_rows) {
    const bool packed = c->pend.packed;  // 2-bit image on the host (pgb_load_reads_2bit): N masks and flags are already on the device
    if (!packed) {
      CU(cudaMemsetAsync(c->d_w, 0, c->n_words * 8, c->st));
      CU(cudaMemsetAsync(c->d_nm, 0, c->n_words * 4, c->st));
      CU(cudaMemsetAsync(c->d_hasn_by_rid, 0, ((size_t)c->max_rid + 1) * 4, c->st));
    }
    // chunk size in .seqdb bytes (= bases): a 2-bit chunk carries a quarter of that over the bus
    uint64_t CH = getenv("PGB_LOAD_CHUNK_MB") ? strtoull(getenv("PGB_LOAD_CHUNK_MB"), 0, 10) << 20 : (uint64_t)96 << 20;
    const bool ring = !packed && c->raw_ring != 0;  // the raw image passes through two staging windows (pgb_load_reads)
    if (ring) CH = std::min<uint64_t>(CH, c->raw_ring);
    std::vector<cudaEvent_t> packed_ev;  // ring: window h may be overwritten once the pack kernel of the chunk before last is done
    uint64_t raw_o = 0, word_o = 2;
    size_t r0 = 0, n_ev = 0, chunk_no = 0;
    while (r0 < ns) {
      size_t r1 = r0;
      uint64_t bytes = 0, wcount = 0;
      while (r1 < ns && (ring ? bytes + c->h_row_len[r1] <= CH : bytes < CH)) { bytes += c->h_row_len[r1]; wcount += ((uint64_t)c->h_row_len[r1] + 31) / 32; r1++; }
      if (n_ev == c->ev_pool.size()) { cudaEvent_t e; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_pool.push_back(e); }
      if (packed) {  // the words travel as they are
        if (wcount) {
          CU(cudaMemcpyAsync(c->d_w + word_o, c->pend.src_words + (word_o - 2), wcount * 8, cudaMemcpyHostToDevice, c->st_copy));
          c->stats.h2d_bytes += wcount * 8;
        }
      } else if (bytes) {
        uint8_t *dst = ring ? c->d_raw + (chunk_no & 1) * c->raw_ring : c->d_raw + raw_o;
        if (ring && chunk_no >= 2) CU(cudaStreamWaitEvent(c->st_copy, packed_ev[chunk_no - 2], 0));
        CU(cudaMemcpyAsync(dst, c->pend.src + raw_o, bytes, cudaMemcpyHostToDevice, c->st_copy));
        c->stats.h2d_bytes += bytes;
      }
      CU(cudaEventRecord(c->ev_pool[n_ev], c->st_copy));
      CU(cudaStreamWaitEvent(c->st, c->ev_pool[n_ev], 0));
      n_ev++;
      if (wcount && !packed)
        LAUNCH(c, k_pack_reads, nblk(wcount), 256, ring ? c->d_raw + (chunk_no & 1) * c->raw_ring : c->d_raw, c->d_row_raw_off, c->d_row_len, c->d_row_woff,
               c->d_row_rid, (uint32_t)c->n_rows, word_o, wcount, c->d_w, c->d_nm, c->d_hasn_by_rid, ring ? raw_o : (uint64_t)0);
      if (ring) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU(cudaEventRecord(e, c->st));
        packed_ev.push_back(e);
      }
      sketch_rows(r0, r1);
      raw_o += bytes; word_o += wcount; r0 = r1; chunk_no++;
    }
    if (!packed_ev.empty()) {
      CU(cudaStreamSynchronize(c->st));
      for (auto e : packed_ev) cudaEventDestroy(e);
    }
    c->pend.active = false;
    c->stats.bases_packed += c->sel_bases;
    if (packed) c->pend.packed = false;
    else if (!c->pend.keep_raw) c->release(c->d_raw);
  };
  if (use_strip) {
    uint32_t *caps = c->alloc<uint32_t>(ns + 1), *row_flags = c->alloc<uint32_t>(ns), *fast_cnt = c->alloc<uint32_t>(ns);
    uint32_t *exact_flag = c->alloc<uint32_t>(ns + 1), *exact_pos = c->alloc<uint32_t>(ns + 1);
    uint64_t *tmp_off = c->alloc<uint64_t>(ns + 1), *row_bad = c->alloc<uint64_t>(ns);
    LAUNCH(c, k_row_caps, nblk(ns + 1), 256, c->d_row_len, (uint32_t)ns, w, caps);
    const uint64_t tmp_n = scan_u32_to_u64(c, caps, tmp_off, ns + 1);
    mm128 *tmp = c->alloc<mm128>(tmp_n);
    CU(cudaMemsetAsync(exact_flag, 0, (ns + 1) * 4, c->st));
    const bool k32 = k <= 16;
    const size_t smem = k32 ? ss_smem_bytes<uint32_t>() : ss_smem_bytes<uint64_t>();
    if (k32) CU(cudaFuncSetAttribute(k_sketch_strip<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(k_sketch_strip<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto sketch_rows = [&](size_t r0, size_t r1) {
      if (r1 <= r0) return;
      const unsigned grid = nblk(r1 - r0, SS_WARPS);
      if (k32)
        k_sketch_strip<uint32_t><<<grid, SS_WARPS * 32, smem, c->st>>>(c->d_w, c->d_row_rid, c->d_row_len, c->d_row_woff, c->d_hasn_by_rid, (uint32_t)r0,
                                                                        (uint32_t)(r1 - r0), w, k, tmp_off, tmp, counts, row_flags, row_bad, fast_cnt);
      else
        k_sketch_strip<uint64_t><<<grid, SS_WARPS * 32, smem, c->st>>>(c->d_w, c->d_row_rid, c->d_row_len, c->d_row_woff, c->d_hasn_by_rid, (uint32_t)r0,
                                                                        (uint32_t)(r1 - r0), w, k, tmp_off, tmp, counts, row_flags, row_bad, fast_cnt);
      c->stats.kernel_launches++;
      CU(cudaGetLastError());
    };
    c->ktic();
    if (c->pend.active) deferred_load(sketch_rows);
    else sketch_rows(0, ns);
    c->stats.ms_k_sketch_tiled += c->ktoc(); c->stats.n_k_sketch_tiled++;
    LAUNCH(c, k_row_exact_flags, nblk(ns), 256, row_flags, (uint32_t)ns, exact_flag);
    if (getenv("PGB_VERBOSE")) {  // why reads were handed to the exact automaton
      std::vector<uint32_t> hf(ns);
      c->d2h(hf.data(), row_flags, ns * 4);
      size_t n_tie = 0, n_pal = 0, n_ovf = 0, n_short = 0, n_n = 0, n_any = 0, n_part = 0;
      for (uint32_t f : hf) { n_any += f != 0; n_part += (f & SK_FLAG_PARTIAL) != 0; n_tie += (f & SK_FLAG_TIE) != 0; n_pal += (f & SK_FLAG_PAL) != 0; n_ovf += (f & SK_FLAG_OVERFLOW) != 0;
                              n_short += (f & SK_FLAG_SHORT) != 0; n_n += (f & SK_FLAG_N) != 0; }
      fprintf(stderr, "pgb200: strip sketch: %zu of %zu reads to the exact automaton (single strips of %zu; whole reads: tie %zu, palindrome %zu, overflow %zu, "
                      "short %zu, N %zu)\n", n_any, ns, n_part, n_tie, n_pal, n_ovf, n_short, n_n);
      size_t by_bit[32] = {0};
      for (uint32_t f : hf) for (int b = 8; b < 32; b++) by_bit[b] += (f >> b) & 1u;
      for (int b = 8; b < 20; b++) if (by_bit[b]) fprintf(stderr, "pgb200:   tie check %d: %zu reads\n", b, by_bit[b]);
    }
    exact_tail(row_flags, exact_flag, exact_pos, [&]() {
      LAUNCH(c, k_row_gather, nblk(ns * 32, 256), 256, counts, row_flags, (uint32_t)ns, tmp_off, tmp, c->d_level_off[0], c->d_level[0]);
    }, row_bad, fast_cnt, tmp_off, tmp);
    c->release(row_bad); c->release(fast_cnt);
    c->release(caps); c->release(row_flags); c->release(exact_flag); c->release(exact_pos); c->release(tmp_off); c->release(tmp);
  } else
  if (!use_tiled) {
    // exact automaton for every read (tiny windows, or PGB_SKETCH=exact)
    ensure_loaded(c);
    c->ktic();
    LAUNCH(c, k_sketch_exact<false>, nblk(ns, 64), 64, c->d_w, c->d_nm, c->d_sel_rows, (uint32_t)ns, c->d_row_rid, c->d_row_len,
           c->d_row_woff, c->d_hasn_by_rid, w, k, counts, (const uint64_t *)nullptr, (mm128 *)nullptr);
    c->stats.ms_k_sketch_count += c->ktoc(); c->stats.n_k_sketch_count++;
    c->level_n[0] = scan_u32_to_u64(c, counts, c->d_level_off[0], ns + 1);
    c->d_level[0] = c->palloc<mm128>(c->level_n[0]);
    c->ktic();
    LAUNCH(c, k_sketch_exact<true>, nblk(ns, 64), 64, c->d_w, c->d_nm, c->d_sel_rows, (uint32_t)ns, c->d_row_rid, c->d_row_len,
           c->d_row_woff, c->d_hasn_by_rid, w, k, (uint32_t *)nullptr, c->d_level_off[0], c->d_level[0]);
    c->stats.ms_k_sketch_write += c->ktoc(); c->stats.n_k_sketch_write++;
  } else {
    // tiled fast path; reads it flags (N, ties, palindrome-dense, short, overflow) are redone by the exact automaton
    const int TILE = sk_tile_len(w);
    std::vector<uint32_t> h_tile_off(ns + 1);
    uint64_t nt = 0;
    for (size_t i = 0; i < ns; i++) { h_tile_off[i] = (uint32_t)nt; nt += (c->h_row_len[i] + (uint32_t)TILE - 1) / (uint32_t)TILE; }
    h_tile_off[ns] = (uint32_t)nt;
    if (nt >= (1ull << 31)) throw std::runtime_error("too many sketch tiles in one pgb_index call");
    const uint32_t n_tiles = (uint32_t)nt;
    uint32_t *tile_off = c->alloc<uint32_t>(ns + 1), *tile_cnt = c->alloc<uint32_t>(n_tiles), *row_flags = c->alloc<uint32_t>(ns);
    uint32_t *exact_flag = c->alloc<uint32_t>(ns + 1), *exact_pos = c->alloc<uint32_t>(ns + 1);
    // per-tile record budget: 3x the expected 2/(w+1) density, at most SK_CAP
    uint32_t tile_cap = (uint32_t)((3 * 2 * TILE) / (w + 1) + 31) & ~31u;
    if (tile_cap < 64) tile_cap = 64;
    if (tile_cap > SK_CAP) tile_cap = SK_CAP;
    mm128 *tmp = c->alloc<mm128>((size_t)n_tiles * tile_cap);
    c->h2d(tile_off, h_tile_off.data(), (ns + 1) * 4);
    TileDesc *tile_desc = c->alloc<TileDesc>(n_tiles);
    LAUNCH(c, k_tile_desc, nblk(ns), 256, tile_off, (uint32_t)ns, c->d_row_rid, c->d_row_len, c->d_row_woff, tile_desc);
    CU(cudaMemsetAsync(row_flags, 0, ns * 4, c->st));
    CU(cudaMemsetAsync(exact_flag, 0, (ns + 1) * 4, c->st));
    const bool k32 = k <= 16;
    const size_t smem = k32 ? sk_smem_bytes<uint32_t>() : sk_smem_bytes<uint64_t>();
    if (k32) CU(cudaFuncSetAttribute(k_sketch_tiled<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CU(cudaFuncSetAttribute(k_sketch_tiled<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    auto launch_tiles = [&](uint32_t t0, uint32_t t1) {  // tiles [t0, t1)
      if (t1 <= t0) return;
      if (k32)
        k_sketch_tiled<uint32_t><<<t1 - t0, SK_THREADS, smem, c->st>>>(c->d_w, tile_desc, c->d_hasn_by_rid, w, k, tile_cnt, row_flags, tmp, tile_cap, t0);
      else
        k_sketch_tiled<uint64_t><<<t1 - t0, SK_THREADS, smem, c->st>>>(c->d_w, tile_desc, c->d_hasn_by_rid, w, k, tile_cnt, row_flags, tmp, tile_cap, t0);
      c->stats.kernel_launches++;
      CU(cudaGetLastError());
    };
    c->ktic();
    if (c->pend.active) deferred_load([&](size_t r0, size_t r1) { launch_tiles(h_tile_off[r0], h_tile_off[r1]); });
    else launch_tiles(0, n_tiles);
    c->stats.ms_k_sketch_tiled += c->ktoc(); c->stats.n_k_sketch_tiled++;
    LAUNCH(c, k_row_counts, nblk(ns), 256, tile_off, tile_cnt, row_flags, (uint32_t)ns, counts, exact_flag);
    exact_tail(row_flags, exact_flag, exact_pos, [&]() {
      LAUNCH(c, k_tile_gather, nblk((size_t)n_tiles * 64, 256), 256, tile_off, tile_cnt, row_flags, (uint32_t)ns, n_tiles, c->d_level_off[0], tmp,
             tile_cap, c->d_level[0]);
    }, (const uint64_t *)nullptr, (const uint32_t *)nullptr, (const uint64_t *)nullptr, (const mm128 *)nullptr);
    c->release(tile_desc);
    c->release(tile_off); c->release(tile_cnt); c->release(row_flags); c->release(exact_flag); c->release(exact_pos); c->release(tmp);
  }
  c->stats.ms_sketch += c->toc();
  c->stats.bases_sketched += c->sel_bases;
  c->stats.n_l0 += c->level_n[0];
  // ---- L1, L2
  for (int l = 1; l <= levels; l++) {
    c->tic();
    CU(cudaMemsetAsync(counts, 0, (ns + 1) * 4, c->st));
    // warp per read (measured 0.93 ms per step against 2.03 ms for thread per read, profiles/r2_notes.md); PGB_REDUCE=thread: the old form
    const bool reduce_warp = !(getenv("PGB_REDUCE") && !strcmp(getenv("PGB_REDUCE"), "thread"));
    if (reduce_warp)
      LAUNCH(c, k_reduce_warp<false>, nblk(ns * 32, 128), 128, c->d_level[l - 1], c->d_level_off[l - 1], (uint32_t)ns, (uint32_t)r, counts,
             (const uint64_t *)nullptr, (mm128 *)nullptr);
    else
    LAUNCH(c, k_reduce<false>, nblk(ns, 128), 128, c->d_level[l - 1], c->d_level_off[l - 1], (uint32_t)ns, (uint32_t)r, counts,
           (const uint64_t *)nullptr, (mm128 *)nullptr);
    c->d_level_off[l] = c->palloc<uint64_t>(ns + 1);
    c->level_n[l] = scan_u32_to_u64(c, counts, c->d_level_off[l], ns + 1);
    c->d_level[l] = c->palloc<mm128>(c->level_n[l]);
    if (reduce_warp)
      LAUNCH(c, k_reduce_warp<true>, nblk(ns * 32, 128), 128, c->d_level[l - 1], c->d_level_off[l - 1], (uint32_t)ns, (uint32_t)r,
             (uint32_t *)nullptr, c->d_level_off[l], c->d_level[l]);
    else
    LAUNCH(c, k_reduce<true>, nblk(ns, 128), 128, c->d_level[l - 1], c->d_level_off[l - 1], (uint32_t)ns, (uint32_t)r,
           (uint32_t *)nullptr, c->d_level_off[l], c->d_level[l]);
    c->stats.ms_reduce += c->toc();
    if (l == 1) c->stats.n_l1 += c->level_n[1]; else c->stats.n_l2 += c->level_n[2];
  }
  c->release(counts);
  for (int l = 0; l <= levels; l++)
    if (with_counts & (1 << l)) build_level_counts(c, l);
  c->sync();
  c->check_err("pgb_index");
  API_END(c)
}

extern "C" size_t pgb_index_size(pgb_ctx *c, int level) { return (c && level >= 0 && level < 3) ? c->level_n[level] : 0; }
extern "C" int pgb_index_copy(pgb_ctx *c, int level, mm128_t *out) {
  API_BEGIN(c)
  if (level < 0 || level > 2) throw std::runtime_error("bad level");
  c->d2h(out, c->d_level[level], c->level_n[level] * sizeof(mm128));
  API_END(c)
}
extern "C" size_t pgb_index_count_size(pgb_ctx *c, int level) { return (c && level >= 0 && level < 3) ? c->level_mc[level].size() : 0; }
extern "C" int pgb_index_count_copy(pgb_ctx *c, int level, mm_count_t *out) {
  if (!c || level < 0 || level > 2) return -1;
  memcpy((void *)out, c->level_mc[level].data(), c->level_mc[level].size() * sizeof(mc_entry));
  return 0;
}

// ================================================================================================ overlap inputs
static void build_mc_table(pgb_ctx *c, const mc_entry *d_mc, size_t n_mc) {
  c->release(c->d_mckeys); c->release(c->d_mcvals);
  uint32_t cap = pow2_at_least(2 * (uint64_t)n_mc + 16);
  c->d_mckeys = c->palloc<uint64_t>(cap); c->d_mcvals = c->palloc<uint32_t>(cap); c->mcmask = cap - 1;
  LAUNCH(c, k_fill_u64, 1184, 256, c->d_mckeys, PGB_EMPTY, (size_t)cap);
  CU(cudaMemsetAsync(c->d_mcvals, 0, (size_t)cap * 4, c->st));
  LAUNCH(c, k_mc_add, nblk(n_mc), 256, d_mc, n_mc, c->d_mckeys, c->d_mcvals, cap - 1, c->d_err);
}

extern "C" int pgb_set_shimmers(pgb_ctx *c, const mm128_t *mmers, size_t n, const mm_count_t *counts, size_t n_counts) {
  API_BEGIN(c)
  c->free_shimmers();
  c->tic();
  c->d_shm = c->palloc<mm128>(n); c->n_shm = n; c->shm_owned = true;
  c->h2d(c->d_shm, mmers, n * sizeof(mm128));
  mc_entry *d_mc = c->alloc<mc_entry>(n_counts);
  c->h2d(d_mc, counts, n_counts * sizeof(mc_entry));
  build_mc_table(c, d_mc, n_counts);
  c->release(d_mc);
  c->stats.ms_count += c->toc();
  c->check_err("pgb_set_shimmers");
  API_END(c)
}

extern "C" int pgb_set_shimmers_from_index(pgb_ctx *c, int level) {
  API_BEGIN(c)
  if (level < 0 || level > 2 || !c->d_level[level]) throw std::runtime_error("level not built");
  c->free_shimmers();
  c->tic();
  c->d_shm = c->d_level[level]; c->n_shm = c->level_n[level]; c->shm_owned = false;
  // multiplicity table straight from the mmers (what aggregate_mm_count over the chunk's -MC- file yields)
  size_t n = c->n_shm;
  uint32_t cap = pow2_at_least(2 * (uint64_t)n + 16);
  c->d_mckeys = c->palloc<uint64_t>(cap); c->d_mcvals = c->palloc<uint32_t>(cap); c->mcmask = cap - 1;
  LAUNCH(c, k_fill_u64, 1184, 256, c->d_mckeys, PGB_EMPTY, (size_t)cap);
  CU(cudaMemsetAsync(c->d_mcvals, 0, (size_t)cap * 4, c->st));
  LAUNCH(c, k_mc_insert, nblk(n), 256, c->d_shm, n, c->d_mckeys, c->d_mcvals, cap - 1, c->d_err);
  c->stats.ms_count += c->toc();
  c->check_err("pgb_set_shimmers_from_index");
  API_END(c)
}

// ================================================================================================ multi-GPU plumbing
extern "C" size_t pgb_buffer_elems(pgb_ctx *c, int which) {
  if (!c) return 0;
  switch (which) {
    case PGB_BUF_WORDS: case PGB_BUF_NMASK: return c->d_w ? (size_t)c->n_words : 0;
    case PGB_BUF_ROW_RID: case PGB_BUF_ROW_LEN: case PGB_BUF_ROW_WOFF: case PGB_BUF_ROW_HASN: return c->n_rows;
    case PGB_BUF_LEVEL0: case PGB_BUF_LEVEL1: case PGB_BUF_LEVEL2: return c->level_n[which - PGB_BUF_LEVEL0];
    case PGB_BUF_COUNTS: return c->n_mc_dump;
    case PGB_BUF_ROUTE: return c->n_route;
    case PGB_BUF_OVLP: return c->n_ovl;
    default: return 0;
  }
}
extern "C" int pgb_buffer_copy_out(pgb_ctx *c, int which, void *dst) {
  API_BEGIN(c)
  ensure_loaded(c);
  size_t n = pgb_buffer_elems(c, which);
  const void *src = nullptr;
  size_t esz = 0;
  uint32_t *tmp = nullptr;
  switch (which) {
    case PGB_BUF_WORDS: src = c->d_w; esz = 8; break;
    case PGB_BUF_NMASK: src = c->d_nm; esz = 4; break;
    case PGB_BUF_ROW_RID: src = c->d_row_rid; esz = 4; break;
    case PGB_BUF_ROW_LEN: src = c->d_row_len; esz = 4; break;
    case PGB_BUF_ROW_WOFF: src = c->d_row_woff; esz = 8; break;
    case PGB_BUF_ROW_HASN:
      tmp = c->alloc<uint32_t>(n);
      LAUNCH(c, k_rows_hasn, nblk(n), 256, c->d_row_rid, c->d_hasn_by_rid, (uint32_t)n, tmp);
      src = tmp; esz = 4; break;
    case PGB_BUF_LEVEL0: case PGB_BUF_LEVEL1: case PGB_BUF_LEVEL2: src = c->d_level[which - PGB_BUF_LEVEL0]; esz = 16; break;
    case PGB_BUF_COUNTS: src = c->d_mc_dump; esz = 16; break;
    case PGB_BUF_ROUTE: src = c->d_route; esz = 40; break;
    case PGB_BUF_OVLP: src = c->d_ovl; esz = sizeof(ovlp_rec); break;
    default: throw std::runtime_error("unknown buffer id");
  }
  if (n) CU(cudaMemcpyAsync(dst, src, n * esz, cudaMemcpyDeviceToDevice, c->st));
  c->sync();
  API_END(c)
}
extern "C" int pgb_load_packed_device(pgb_ctx *c, const uint64_t *words, const uint32_t *nmask, size_t n_words, const uint32_t *row_rid,
                                      const uint32_t *row_len, const uint64_t *row_woff, const uint32_t *row_hasn, size_t n_rows) {
  API_BEGIN(c)
  c->free_reads();
  if (n_words < 4) throw std::runtime_error("packed word array must include its guard words");
  uint32_t *d_max = c->alloc<uint32_t>(1);
  CU(cudaMemsetAsync(d_max, 0, 4, c->st));
  LAUNCH(c, k_max_u32, nblk(n_rows), 256, row_rid, (uint32_t)n_rows, d_max);
  uint32_t max_rid = 0;
  c->d2h(&max_rid, d_max, 4);
  if (n_rows && (uint64_t)max_rid > 64 * (uint64_t)n_rows + (1u << 24)) throw std::runtime_error("read ids too sparse");
  c->max_rid = max_rid; c->n_rows = n_rows; c->n_words = n_words;
  c->d_w = c->palloc<uint64_t>(n_words); c->d_nm = c->palloc<uint32_t>(n_words);
  c->d_rlen_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1); c->d_hasn_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1);
  c->d_woff_by_rid = c->palloc<uint64_t>((size_t)max_rid + 1);
  c->d_row_rid = c->palloc<uint32_t>(n_rows); c->d_row_len = c->palloc<uint32_t>(n_rows); c->d_row_woff = c->palloc<uint64_t>(n_rows);
  c->d_sel_rows = c->palloc<uint32_t>(n_rows);
  CU(cudaMemcpyAsync(c->d_w, words, n_words * 8, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->d_nm, nmask, n_words * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->d_row_rid, row_rid, n_rows * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->d_row_len, row_len, n_rows * 4, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemcpyAsync(c->d_row_woff, row_woff, n_rows * 8, cudaMemcpyDeviceToDevice, c->st));
  CU(cudaMemsetAsync(c->d_rlen_by_rid, 0, ((size_t)max_rid + 1) * 4, c->st));
  CU(cudaMemsetAsync(c->d_hasn_by_rid, 0, ((size_t)max_rid + 1) * 4, c->st));
  CU(cudaMemsetAsync(c->d_woff_by_rid, 0, ((size_t)max_rid + 1) * 8, c->st));
  LAUNCH(c, k_rows_to_rid_tables, nblk(n_rows), 256, row_rid, row_len, row_woff, row_hasn, (uint32_t)n_rows, c->d_rlen_by_rid, c->d_woff_by_rid,
         c->d_hasn_by_rid);
  c->h_row_len.resize(n_rows);
  c->d2h(c->h_row_len.data(), c->d_row_len, n_rows * 4);
  std::vector<uint32_t> ident(n_rows);
  for (size_t j = 0; j < n_rows; j++) ident[j] = (uint32_t)j;
  c->h2d(c->d_sel_rows, ident.data(), n_rows * 4);
  c->sel_bases = 0;
  for (uint32_t l : c->h_row_len) c->sel_bases += l;
  c->sync();
  API_END(c)
}
extern "C" int pgb_set_shimmers_device(pgb_ctx *c, const mm128_t *mmers, size_t n) {
  API_BEGIN(c)
  c->free_shimmers();
  c->tic();
  c->d_shm = c->palloc<mm128>(n); c->n_shm = n; c->shm_owned = true;
  if (n) CU(cudaMemcpyAsync(c->d_shm, mmers, n * sizeof(mm128), cudaMemcpyDeviceToDevice, c->st));
  uint32_t cap = pow2_at_least(2 * (uint64_t)n + 16);
  c->d_mckeys = c->palloc<uint64_t>(cap); c->d_mcvals = c->palloc<uint32_t>(cap); c->mcmask = cap - 1;
  LAUNCH(c, k_fill_u64, 1184, 256, c->d_mckeys, PGB_EMPTY, (size_t)cap);
  CU(cudaMemsetAsync(c->d_mcvals, 0, (size_t)cap * 4, c->st));
  LAUNCH(c, k_mc_insert, nblk(n), 256, c->d_shm, n, c->d_mckeys, c->d_mcvals, cap - 1, c->d_err);
  c->stats.ms_count += c->toc();
  c->check_err("pgb_set_shimmers_device");
  API_END(c)
}

// ================================================================================================ overlap
// build_map's record stream for hash chunk `mychunk` of T from the context's shimmer list + multiplicity table
// (src/shmr_utils.c:295-404).  Returns the number of records; R is allocated from the scratch arena.
static uint32_t build_pair_records(pgb_ctx *c, uint32_t T, uint32_t mychunk, uint32_t mc_lower, uint32_t mc_upper, PairSoA &R) {
  size_t n = c->n_shm;
  c->tic();
  uint32_t *cnt = c->alloc<uint32_t>(n);
  uint32_t *flags = c->alloc<uint32_t>(n + 1), *pos = c->alloc<uint32_t>(n + 1);
  unsigned long long *d_first = c->alloc<unsigned long long>(1);
  CU(cudaMemsetAsync(d_first, 0xFF, 8, c->st));
  CU(cudaMemsetAsync(flags, 0, (n + 1) * 4, c->st));
  LAUNCH(c, k_count_lookup, nblk(n), 256, c->d_shm, n, c->d_mckeys, c->d_mcvals, c->mcmask, cnt, mc_lower, mc_upper, d_first, c->d_err);
  LAUNCH(c, k_kept_flags, nblk(n), 256, cnt, n, mc_lower, mc_upper, d_first, flags);
  uint32_t n_kept = scan_u32(c, flags, pos, n + 1);
  uint32_t *kept = c->alloc<uint32_t>(n_kept);
  LAUNCH(c, k_compact_idx, nblk(n), 256, flags, pos, n, kept);
  c->release(cnt); c->release(flags); c->release(pos); c->release(d_first);
  uint32_t *n_rec = c->alloc<uint32_t>((size_t)n_kept + 1), *rec_off = c->alloc<uint32_t>((size_t)n_kept + 1);
  CU(cudaMemsetAsync(n_rec, 0, ((size_t)n_kept + 1) * 4, c->st));
  LAUNCH(c, k_pair_count, nblk(n_kept), 256, c->d_shm, kept, n_kept, T, mychunk, n_rec);
  uint32_t nrec = scan_u32(c, n_rec, rec_off, (size_t)n_kept + 1);
  R.k0 = c->alloc<uint64_t>(nrec); R.k1 = c->alloc<uint64_t>(nrec); R.y0 = c->alloc<uint64_t>(nrec); R.y1 = c->alloc<uint64_t>(nrec);
  R.seq = c->alloc<uint32_t>(nrec); R.dir = c->alloc<uint8_t>(nrec);
  LAUNCH(c, k_pair_write, nblk(n_kept), 256, c->d_shm, kept, n_kept, T, mychunk, rec_off, c->d_rlen_by_rid, R);
  c->release(kept); c->release(n_rec); c->release(rec_off);
  c->stats.ms_pairs += c->toc();
  c->stats.n_pair_records += nrec;
  return nrec;
}

// process_overlaps (src/shmr_overlap.c:182-231) over a chunk's record stream (records in insertion order, seq ascending).
// Consumes R.  Returns 0, or -1 with c->err set; throws on resource errors.
static int overlap_core(pgb_ctx *c, PairSoA R, uint32_t nrec, uint32_t bestn, uint32_t bw, uint32_t ovlp_upper) {
  if (bw + 3 > PGB_MAXV) throw std::runtime_error("align bandwidth (-w) above 256 is not supported by this build");
  if (ovlp_upper > 65535) throw std::runtime_error("ovlp_upper (-n) above 65535 is not supported");
  bestn &= 0xFF;  // uint8_t bestn = atoi(), src/shmr_overlap.c:245,286
  ensure_loaded(c);
  auto free_R = [&]() { c->release(R.k0); c->release(R.k1); c->release(R.y0); c->release(R.y1); c->release(R.seq); c->release(R.dir); };
  if (nrec == 0) { free_R(); ensure_rc(c); c->sync(); c->check_err("pgb_overlap/build_map"); return c->err.empty() ? 0 : -1; }

  // ---------------- bucket tables
  c->tic();
  uint32_t xcap = pow2_at_least(4 * (uint64_t)nrec), bcap = pow2_at_least(2 * (uint64_t)nrec);
  uint64_t *xkeys = c->alloc<uint64_t>(xcap), *bkeys = c->alloc<uint64_t>(bcap);
  uint32_t *bcount = c->alloc<uint32_t>(bcap), *bfirst = c->alloc<uint32_t>(bcap), *blast = c->alloc<uint32_t>(bcap), *rec_bucket = c->alloc<uint32_t>(nrec);
  uint32_t *xfirst = c->alloc<uint32_t>(xcap), *xid = c->alloc<uint32_t>(xcap);
  LAUNCH(c, k_fill_u64, 1184, 256, xkeys, PGB_EMPTY, (size_t)xcap);
  LAUNCH(c, k_fill_u64, 1184, 256, bkeys, PGB_EMPTY, (size_t)bcap);
  CU(cudaMemsetAsync(bcount, 0, (size_t)bcap * 4, c->st));
  CU(cudaMemsetAsync(bfirst, 0xFF, (size_t)bcap * 4, c->st));
  CU(cudaMemsetAsync(blast, 0, (size_t)bcap * 4, c->st));
  CU(cudaMemsetAsync(xfirst, 0xFF, (size_t)xcap * 4, c->st));
  LAUNCH(c, k_bucket_insert, nblk(nrec), 256, R, nrec, xkeys, xcap - 1, bkeys, bcap - 1, bcount, bfirst, blast, xfirst, rec_bucket, c->d_err);
  uint32_t *bflags = c->alloc<uint32_t>((size_t)bcap + 1), *bpos = c->alloc<uint32_t>((size_t)bcap + 1);
  CU(cudaMemsetAsync(bflags, 0, ((size_t)bcap + 1) * 4, c->st));
  LAUNCH(c, k_mc_flags, nblk(bcap), 256, bkeys, (size_t)bcap, bflags);
  uint32_t n_buckets = scan_u32(c, bflags, bpos, (size_t)bcap + 1);
  auto free_tables = [&]() {
    c->release(bflags); c->release(bpos); c->release(xkeys); c->release(bkeys); c->release(bcount); c->release(bfirst); c->release(blast);
    c->release(xfirst); c->release(xid);
  };
  if (c->check_err("pgb_overlap/buckets")) { free_tables(); free_R(); c->release(rec_bucket); return -1; }
  c->stats.n_buckets += n_buckets;

  // ---------------- visiting order: everything on the GPU except the replay of the OUTER khash
  // (a) buckets in first-insertion order
  uint32_t *lslot = c->alloc<uint32_t>(n_buckets), *lkey = c->alloc<uint32_t>(n_buckets);
  uint32_t *sslot = c->alloc<uint32_t>(n_buckets), *skey = c->alloc<uint32_t>(n_buckets);
  LAUNCH(c, k_bucket_list, nblk(bcap), 256, bkeys, bfirst, bpos, (size_t)bcap, lslot, lkey);
  sort_pairs_u32(c, lkey, skey, lslot, sslot, n_buckets);
  // (b) outer ids in first-insertion order
  uint32_t *isfirst = c->alloc<uint32_t>((size_t)n_buckets + 1), *firstpos = c->alloc<uint32_t>((size_t)n_buckets + 1);
  CU(cudaMemsetAsync(isfirst, 0, ((size_t)n_buckets + 1) * 4, c->st));
  LAUNCH(c, k_outer_first, nblk(n_buckets), 256, sslot, n_buckets, bkeys, bfirst, xfirst, isfirst);
  uint32_t n_outer = scan_u32(c, isfirst, firstpos, (size_t)n_buckets + 1);
  LAUNCH(c, k_outer_id_first, nblk(n_buckets), 256, sslot, n_buckets, bkeys, isfirst, firstpos, xid);
  uint32_t *oid = lkey, *idx = lslot;  // reuse
  LAUNCH(c, k_outer_id_all, nblk(n_buckets), 256, sslot, n_buckets, bkeys, xid, oid, idx);
  // (c) group by outer id (stable: inner first-insertion order is kept inside a group)
  uint32_t *goid_sorted = skey, *gidx = c->alloc<uint32_t>(n_buckets);
  sort_pairs_u32(c, oid, goid_sorted, idx, gidx, n_buckets);
  GroupedBuckets G;
  G.slot = c->alloc<uint32_t>(n_buckets); G.first = c->alloc<uint32_t>(n_buckets); G.last = c->alloc<uint32_t>(n_buckets);
  G.count = c->alloc<uint32_t>(n_buckets); G.oid = c->alloc<uint32_t>(n_buckets); G.k1 = c->alloc<uint64_t>(n_buckets);
  uint32_t *goff = c->alloc<uint32_t>((size_t)n_outer + 1), *ipos = c->alloc<uint32_t>(n_buckets), *big_list = c->alloc<uint32_t>((size_t)n_outer + 1);
  uint64_t *okey = c->alloc<uint64_t>((size_t)n_outer + 1);
  unsigned int *d_small = c->alloc<unsigned int>(4);  // [0] last_seq_all, [1] n_big, [2] start of the last group, [3] its first sequence number
  unsigned long long *d_ncand = c->alloc<unsigned long long>(1);
  CU(cudaMemsetAsync(d_small, 0, 16, c->st));
  CU(cudaMemsetAsync(d_ncand, 0, 8, c->st));
  LAUNCH(c, k_group_gather, nblk(n_buckets), 256, gidx, goid_sorted, n_buckets, sslot, bkeys, xkeys, bfirst, blast, bcount, G, goff, n_outer, okey,
         d_small);
  // (d) inner khash replay, one thread per outer key
  LAUNCH(c, k_inner_order, nblk(n_outer, 128), 128, goff, n_outer, G, ipos, big_list, d_small + 1);
  // the outer keys go to a page-locked buffer; the reverse-complement image of the reads (needed by the alignments only) is
  // built behind that copy, while the host replays the outer khash
  if (c->h_okey_cap < (size_t)n_outer + 4) {
    if (c->h_okey) cudaFreeHost(c->h_okey);
    c->h_okey = nullptr; c->h_okey_cap = 0;
    const size_t cap = (size_t)n_outer + (n_outer >> 2) + 1024;
    CU(cudaMallocHost((void **)&c->h_okey, cap * 8));
    c->h_okey_cap = cap;
  }
  uint64_t *h_okey_p = c->h_okey;
  unsigned int *h_small = reinterpret_cast<unsigned int *>(h_okey_p + n_outer);  // 4 x u32 (see d_small)
  LAUNCH(c, k_last_group_first, 1, 1, goff, n_outer, G.first, (uint32_t *)(d_small + 2));
  CU(cudaMemcpyAsync(h_okey_p, okey, (size_t)n_outer * 8, cudaMemcpyDeviceToHost, c->st));
  CU(cudaMemcpyAsync(h_small, d_small, 16, cudaMemcpyDeviceToHost, c->st));
  c->stats.d2h_bytes += (size_t)n_outer * 8 + 16;
  CU(cudaEventRecord(c->ev_pass[0], c->st));
  unsigned int *d_rc_count = rc_launch(c);
  CU(cudaEventSynchronize(c->ev_pass[0]));
  const uint32_t newest_outer_seq = h_small[3];
  {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev0, c->ev_pass[0]));
    c->stats.ms_buckets += ms;
  }
  struct OkeyView { const uint64_t *p; const uint64_t *data() const { return p; } } h_okey{h_okey_p};
  // (e) host: outer khash replay (keys only) -> visiting rank of every outer key; inner replay of the few oversized groups
  double t_host0 = now_ms();
  std::vector<uint32_t> orank(n_outer);
  {
    KhashEmu &outer = c->outer_emu;  // kept across calls: its table (8 B per slot) is not re-faulted every step
    outer.clear();
    outer.put_all(h_okey.data(), n_outer);
    if (h_small[0] > newest_outer_seq) outer.touch_existing();  // a put of a known outer key followed the last new one
    const double t_put = now_ms();
    uint32_t r = 0;
    outer.for_each_in_slot_order([&](uint64_t, uint32_t o) { orank[o] = r++; });
    if (getenv("PGB_VERBOSE"))
      fprintf(stderr, "pgb200: outer khash replay: %u keys, table %u slots, put %.2f ms, slot order %.2f ms; %u oversized inner groups\n", n_outer, outer.nb,
              t_put - t_host0, now_ms() - t_put, h_small[1]);
  }
  if (h_small[1]) {
    std::vector<uint32_t> big(h_small[1]), h_goff((size_t)n_outer + 1);
    c->d2h(big.data(), big_list, (size_t)h_small[1] * 4);
    c->d2h(h_goff.data(), goff, ((size_t)n_outer + 1) * 4);
    KhashEmu inner;
    for (uint32_t o : big) {
      uint32_t b = h_goff[o], n = h_goff[o + 1] - b;
      std::vector<uint64_t> k1(n);
      std::vector<uint32_t> fi(n), la(n), pos(n);
      c->d2h(k1.data(), G.k1 + b, (size_t)n * 8);
      c->d2h(fi.data(), G.first + b, (size_t)n * 4);
      c->d2h(la.data(), G.last + b, (size_t)n * 4);
      inner.clear();
      uint32_t last = 0;
      for (uint32_t i = 0; i < n; i++) { inner.put_new(k1[i], i); last = std::max(last, la[i]); }
      if (last > fi[n - 1]) inner.touch_existing();
      uint32_t r = 0;
      inner.for_each_in_slot_order([&](uint64_t, uint32_t i) { pos[i] = r++; });
      c->h2d(ipos + b, pos.data(), (size_t)n * 4);
      c->sync();
    }
  }
  c->stats.ms_host_order += now_ms() - t_host0;
  rc_finish(c, d_rc_count);
  // (f) visiting positions, eligibility, ranks
  c->tic();
  uint32_t *d_orank = c->alloc<uint32_t>(n_outer), *size_by_rank = c->alloc<uint32_t>((size_t)n_outer + 1), *vstart = c->alloc<uint32_t>((size_t)n_outer + 1);
  c->h2d(d_orank, orank.data(), (size_t)n_outer * 4);
  CU(cudaMemsetAsync(size_by_rank, 0, ((size_t)n_outer + 1) * 4, c->st));
  LAUNCH(c, k_group_sizes_by_rank, nblk(n_outer), 256, goff, d_orank, n_outer, size_by_rank);
  scan_u32(c, size_by_rank, vstart, (size_t)n_outer + 1);
  uint32_t *vis_slot = c->alloc<uint32_t>(n_buckets), *vis_elig = c->alloc<uint32_t>((size_t)n_buckets + 1), *vis_cnt = c->alloc<uint32_t>((size_t)n_buckets + 1);
  uint32_t *rank_of = c->alloc<uint32_t>((size_t)n_buckets + 1), *off_of = c->alloc<uint32_t>((size_t)n_buckets + 1);
  CU(cudaMemsetAsync(vis_elig, 0, ((size_t)n_buckets + 1) * 4, c->st));
  CU(cudaMemsetAsync(vis_cnt, 0, ((size_t)n_buckets + 1) * 4, c->st));
  LAUNCH(c, k_visit_place, nblk(n_buckets), 256, G, n_buckets, d_orank, vstart, ipos, ovlp_upper, vis_slot, vis_elig, vis_cnt, d_ncand);
  uint32_t n_ranks = scan_u32(c, vis_elig, rank_of, (size_t)n_buckets + 1);
  uint32_t n_elig = scan_u32(c, vis_cnt, off_of, (size_t)n_buckets + 1);
  unsigned long long n_cand = 0;
  c->d2h(&n_cand, d_ncand, 8);
  c->stats.n_eligible_buckets += n_ranks; c->stats.n_candidates += n_cand;
  uint32_t *d_slot2rank = c->alloc<uint32_t>(bcap), *d_rank_off = c->alloc<uint32_t>((size_t)n_ranks + 1);
  LAUNCH(c, k_fill_u32, 1184, 256, d_slot2rank, PGB_NOSLOT, (size_t)bcap);
  LAUNCH(c, k_visit_rank, nblk(n_buckets), 256, vis_slot, vis_elig, rank_of, off_of, n_buckets, d_slot2rank, d_rank_off);
  CU(cudaMemcpyAsync(d_rank_off + n_ranks, &n_elig, 4, cudaMemcpyHostToDevice, c->st));
  c->sync();
  auto free_order = [&]() {
    c->release(lslot); c->release(lkey); c->release(sslot); c->release(skey); c->release(isfirst); c->release(firstpos); c->release(gidx);
    c->release(G.slot); c->release(G.first); c->release(G.last); c->release(G.count); c->release(G.oid); c->release(G.k1);
    c->release(goff); c->release(ipos); c->release(big_list); c->release(okey); c->release(d_small); c->release(d_ncand);
    c->release(d_orank); c->release(size_by_rank); c->release(vstart); c->release(vis_slot); c->release(vis_elig); c->release(vis_cnt);
    c->release(rank_of); c->release(off_of);
  };
  free_order();
  free_tables();
  if (n_ranks == 0) { free_R(); c->release(rec_bucket); c->release(d_slot2rank); c->release(d_rank_off); c->sync(); return 0; }

  // ---------------- scatter + per-bucket sort
  uint32_t *fill = c->alloc<uint32_t>(n_ranks);
  CU(cudaMemsetAsync(fill, 0, (size_t)n_ranks * 4, c->st));
  uint64_t *sy0 = c->alloc<uint64_t>(n_elig), *sy1 = c->alloc<uint64_t>(n_elig);
  uint32_t *sseq = c->alloc<uint32_t>(n_elig); uint8_t *sdir = c->alloc<uint8_t>(n_elig), *contained = c->alloc<uint8_t>(n_elig);
  LAUNCH(c, k_scatter, nblk(nrec), 256, R, nrec, rec_bucket, d_slot2rank, d_rank_off, fill, sy0, sy1, sseq, sdir);
  LAUNCH(c, k_sort_buckets, nblk(n_ranks, 64), 64, n_ranks, d_rank_off, sy0, sy1, sseq, sdir);
  free_R(); c->release(rec_bucket); c->release(d_slot2rank); c->release(fill);
  c->stats.ms_buckets += c->toc();

  // ---------------- replay / align fix-point (DESIGN.md "ordered greedy as a fix-point")
  if (n_ranks >= (1u << 30)) throw std::runtime_error("more than 2^30 eligible buckets in one overlap call");
  const bool incremental = !(getenv("PGB_REPLAY_FULL"));  // PGB_REPLAY_FULL=1: replay every bucket in every pass
  // buckets with >= BIG_N records are replayed by a CTA (k_replay_block), smaller ones by a thread (k_replay);
  // BIG_TAIL is the threshold of the incremental passes, where only the critical path of the biggest bucket matters
  const uint32_t BIG_N = getenv("PGB_REPLAY_BIG") ? (uint32_t)atoi(getenv("PGB_REPLAY_BIG")) : 48u;
  const uint32_t BIG_TAIL = getenv("PGB_REPLAY_BIG_TAIL") ? (uint32_t)atoi(getenv("PGB_REPLAY_BIG_TAIL")) : 8u;
  // buckets of the thread class with at least RW_MIN records are walked by RW_G lanes each (k_replay_group); PGB_REPLAY_WARP_MIN=64: none
  // Measured on config 2 (profiles/r2_replay.md): full passes are throughput bound (every bucket runs; a thread stops at bestn, a
  // group probes up to G-1 candidates too many) and stay thread-walked (64 = no group buckets); incremental passes are latency
  // bound (their time is the longest dependent probe chain of one bucket) and hand buckets of >= 8 records to lane groups.
  const uint32_t RW_MIN = getenv("PGB_REPLAY_WARP_MIN") ? (uint32_t)atoi(getenv("PGB_REPLAY_WARP_MIN")) : 64u;
  const uint32_t RW_MIN_INC = getenv("PGB_REPLAY_WARP_MIN_INC") ? (uint32_t)atoi(getenv("PGB_REPLAY_WARP_MIN_INC")) : (getenv("PGB_REPLAY_WARP_MIN") ? RW_MIN : 8u);
  const int RW_G = getenv("PGB_REPLAY_GROUP") ? atoi(getenv("PGB_REPLAY_GROUP")) : 8;
  const uint32_t TAIL_RUN = getenv("PGB_TAIL_RUN") ? (uint32_t)strtoul(getenv("PGB_TAIL_RUN"), 0, 10) : 40000u;
  // speculative passes before real alignments are computed: 2 cost ~2 % extra alignments and save two full passes
  const int MAX_DRY = getenv("PGB_DRY_PASSES") ? atoi(getenv("PGB_DRY_PASSES")) : 2;
  // alignment batches up to this size go to the warp-per-alignment kernel (k_align_warp), larger ones to k_align_lean
  const uint32_t ALIGN_WARP_MAX = getenv("PGB_ALIGN_WARP_MAX") ? (uint32_t)strtoul(getenv("PGB_ALIGN_WARP_MAX"), 0, 10) : 16384u;
  const bool verbose = getenv("PGB_VERBOSE") != nullptr;
  const int ALIGN_VARIANT = getenv("PGB_ALIGN_VARIANT") ? atoi(getenv("PGB_ALIGN_VARIANT")) : 7;
  const uint32_t CHANGED_CAP = 16384;
  // incremental passes: (rid, rank) index of the eligible records sorted by rid + per-bucket Bloom filter of read ids
  uint32_t *rid_sorted = c->alloc<uint32_t>(n_elig), *rank_sorted = c->alloc<uint32_t>(n_elig);
  uint64_t *bloom = c->alloc<uint64_t>(4 * (size_t)n_ranks), *changed = c->alloc<uint64_t>(CHANGED_CAP);
  uint8_t *unk_flag = c->alloc<uint8_t>(n_ranks), *dirty = c->alloc<uint8_t>(n_ranks);
  uint32_t *dflags = c->alloc<uint32_t>(2 * (size_t)n_ranks + 1), *dpos = c->alloc<uint32_t>(2 * (size_t)n_ranks + 1), *dlist = c->alloc<uint32_t>(n_ranks);
  uint32_t *all_list = c->alloc<uint32_t>(n_ranks);
  uint32_t *acc = c->alloc<uint32_t>((size_t)n_ranks + 1), *out_off = c->alloc<uint32_t>((size_t)n_ranks + 1);
  unsigned long long *d_ctr = c->alloc<unsigned long long>(2);  // [0] unknown alignments [1] table diffs
  {
    uint32_t *rr = c->alloc<uint32_t>(n_elig), *rk = c->alloc<uint32_t>(n_elig);
    LAUNCH(c, k_bucket_reads, nblk(n_ranks, 128), 128, n_ranks, d_rank_off, sy0, rr, rk, bloom);
    sort_pairs_u32(c, rr, rid_sorted, rk, rank_sorted, n_elig);
    c->release(rr); c->release(rk);
  }
  bool tail_mode = false;  // few buckets left to replay: latency of the biggest one is all that matters
  // (small..., big...) run list of a pass; returns the number of small buckets, *n_run = total
  auto class_lists = [&](const uint8_t *dirty_or_null, uint32_t *list, uint32_t *n_run) -> uint32_t {
    CU(cudaMemsetAsync(dflags + 2 * (size_t)n_ranks, 0, 4, c->st));
    LAUNCH(c, k_class_flags, nblk(n_ranks), 256, d_rank_off, n_ranks, dirty_or_null, (dirty_or_null && tail_mode) ? std::min(BIG_N, BIG_TAIL) : BIG_N, dflags);
    *n_run = scan_u32(c, dflags, dpos, 2 * (size_t)n_ranks + 1);
    uint32_t n_small = 0;
    c->d2h(&n_small, dpos + n_ranks, 4);
    LAUNCH(c, k_compact_classes, nblk(2 * (size_t)n_ranks), 256, dflags, dpos, n_ranks, list);
    return n_small;
  };
  uint32_t n_all = 0;
  const uint32_t n_all_small = class_lists(nullptr, all_list, &n_all);
  const uint32_t n_all_big = n_all - n_all_small;
  // upper bound of the CTA-replayed buckets of an incremental pass in tail mode (threshold BIG_TAIL instead of BIG_N)
  uint32_t n_tail_big = n_all_big;
  if (BIG_TAIL < BIG_N) {
    tail_mode = true;
    std::vector<uint8_t> ones(n_ranks, 1);
    c->h2d(dirty, ones.data(), n_ranks);
    uint32_t n_t = 0;
    const uint32_t n_t_small = class_lists(dirty, dlist, &n_t);
    n_tail_big = n_t - n_t_small;
    tail_mode = false;
  }
  // the same without a host round trip: the run list's sizes stay on the device (run_cnt), the kernels read them there
  uint32_t *run_cnt = c->alloc<uint32_t>(2), *all_cnt = c->alloc<uint32_t>(2);
  PassOut *d_pass = c->alloc<PassOut>(1);
  {
    const uint32_t h_all[2] = {n_all_small, n_all};
    c->h2d(all_cnt, h_all, 8);
  }
  auto class_lists_dev = [&](const uint8_t *dirty_flags, uint32_t *list) {
    CU(cudaMemsetAsync(dflags + 2 * (size_t)n_ranks, 0, 4, c->st));
    LAUNCH(c, k_class_flags, nblk(n_ranks), 256, d_rank_off, n_ranks, dirty_flags, tail_mode ? std::min(BIG_N, BIG_TAIL) : BIG_N, dflags);
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum((void *)nullptr, tmp_bytes, dflags, dpos, (int)(2 * (size_t)n_ranks + 1), c->st));
    uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
    CU(cub::DeviceScan::ExclusiveSum((void *)tmp, tmp_bytes, dflags, dpos, (int)(2 * (size_t)n_ranks + 1), c->st));
    c->stats.kernel_launches += 2;
    LAUNCH(c, k_run_counts, 1, 1, dpos, n_ranks, run_cnt);
    LAUNCH(c, k_compact_classes, nblk(2 * (size_t)n_ranks), 256, dflags, dpos, n_ranks, list);
  };
  auto free_common = [&]() {
    c->release(rid_sorted); c->release(rank_sorted); c->release(bloom); c->release(changed); c->release(unk_flag); c->release(dirty);
    c->release(dflags); c->release(dpos); c->release(dlist); c->release(all_list); c->release(acc); c->release(out_off); c->release(d_ctr);
    c->release(run_cnt); c->release(all_cnt); c->release(d_pass);
    c->release(sy0); c->release(sy1); c->release(sseq); c->release(sdir); c->release(contained); c->release(d_rank_off);
  };

  // table capacities: sized for what a chunk normally needs (pairs ever accepted ~ 0.3 n_elig, alignments ~ 0.25 n_elig);
  // a chunk that needs more raises the overflow flag and the fix-point restarts with doubled tables
  // The ratios are remembered by the context (a later call on a similar job starts with tables that fit) and grow with the
  // chunk count: rid_pairs is per chunk, so at T chunks a read pair is aligned in up to T of them (x4.5 at T=8).
  // PGB_TABLE_SCALE (tests): scale the initial capacities down to force the overflow / restart path on small inputs.
  const char *ts_env = getenv("PGB_TABLE_SCALE");
  const double tscale = ts_env ? atof(ts_env) : 1.0;
  uint64_t ecap64 = (uint64_t)(c->ecap_ratio * tscale * n_elig) + (ts_env ? 64 : 4096);
  // The alignment cache (key -> result slot + request queue, 60 B per slot) is only needed once real alignments are asked
  // for.  The speculative passes count the alignments they had to predict, which is (within a few %) the number the wet
  // passes will request, so the cache is sized from that count when the first wet pass starts; until then a 1024-slot
  // empty table answers every lookup with "unknown".  acap64 == 0 means "not sized yet".
  uint64_t acap64 = 0;
  bool converged = false;
  ReplayState S;
  memset(&S, 0, sizeof S);
  // S tables come from the block cache (not the bump arena): a restart hands them back and the next attempt reuses them
  auto free_S = [&]() { c->release(S.E); c->release(S.akeys); c->release(S.ares); c->release(S.reqs); c->release(S.n_req); };
  auto alloc_aln_tables = [&](uint64_t cap) {
    c->release(S.akeys); c->release(S.ares); c->release(S.reqs);
    S.acap = (uint32_t)cap; S.req_cap = S.acap;
    S.akeys = c->palloc<uint64_t>(S.acap); S.ares = c->palloc<match_t>(S.acap); S.reqs = c->palloc<AlnReq>(S.req_cap);
    LAUNCH(c, k_fill_u64, 1184, 256, S.akeys, PGB_EMPTY, (size_t)S.acap);
  };
  for (int attempt = 0; !converged; attempt++) {
    if (attempt > 24 || ecap64 >= (1ull << 32) || acap64 >= (1ull << 32)) { free_S(); free_common(); throw std::runtime_error("replay tables overflow"); }
    free_S();
    memset(&S, 0, sizeof S);
    S.ecap = (uint32_t)ecap64;
    S.E = c->palloc<EEntry>(S.ecap);
    S.n_req = c->palloc<uint32_t>(1); S.rlen_by_rid = c->d_rlen_by_rid; S.err = c->d_err;
    S.ctr = d_ctr;
    S.cur = 0;
    alloc_aln_tables(acap64 ? acap64 : 1024);
    LAUNCH(c, k_e_init, 1184, 256, S.E, (size_t)S.ecap);
    CU(cudaMemsetAsync(S.n_req, 0, 4, c->st));
    CU(cudaMemsetAsync(acc, 0, ((size_t)n_ranks + 1) * 4, c->st));
    CU(cudaMemsetAsync(unk_flag, 0, n_ranks, c->st));
    // run list of a pass: host-known sizes (cnt_dev == nullptr) or sizes left on the device by class_lists_dev
    auto launch_replay = [&](const uint32_t *list, uint32_t n_small, uint32_t n_total, int request, int emit, const uint32_t *ooff, ovlp_rec *out,
                             const uint32_t *cnt_dev) {
      const uint32_t nb = n_total - n_small;
      // of the "small" list: buckets below rw_min records are walked by a thread, the others by a group of RW_G lanes (the size
      // test is in the kernels); incremental passes (cnt_dev) are latency bound and use their own threshold
      const uint32_t rw_min = cnt_dev ? RW_MIN_INC : RW_MIN;
      LAUNCH(c, k_replay, nblk(n_small, 64), 64, S, n_small, list, d_rank_off, sy0, sdir, contained, bestn, request, emit, acc, ooff, out, unk_flag, cnt_dev,
             rw_min, (uint32_t)PGB_RW_MAXN);
      if (rw_min < PGB_RW_MAXN) {
#define PGB_RG(G)                                                                                                                                      \
  LAUNCH(c, k_replay_group<G>, nblk(n_small, PGB_RW_WARPS * (32 / G)), PGB_RW_WARPS * 32, S, n_small, list, d_rank_off, sy0, sdir, contained, bestn, \
         request, emit, acc, ooff, out, unk_flag, cnt_dev, rw_min, (uint32_t)PGB_RW_MAXN)
        if (RW_G == 4) PGB_RG(4);
        else if (RW_G == 16) PGB_RG(16);
        else if (RW_G == 32) PGB_RG(32);
        else PGB_RG(8);
#undef PGB_RG
      }
      LAUNCH(c, k_replay_block, nb, PGB_RB_THREADS, S, nb, cnt_dev ? list : list + n_small, d_rank_off, sy0, sdir, contained, bestn, request, emit, acc, ooff,
             out, unk_flag, cnt_dev);
    };
    bool wet = MAX_DRY <= 0, overflow = false;
    if (wet && acap64 == 0) {  // PGB_DRY_PASSES=0: no speculative count to size from
      acap64 = (uint64_t)(c->acap_ratio * tscale * n_elig) + (ts_env ? 64 : 4096);
      alloc_aln_tables(acap64);
    }
    tail_mode = false;
    uint64_t prev_diffs = ~0ULL, last_diffs = ~0ULL;
    uint32_t n_done = 0;
    int dry_passes = 0;
    bool align_pending = false;  // an alignment batch was launched whose event pair / base count has not been read yet
    auto read_align_timing = [&]() {
      if (!align_pending) return;
      float ms = 0;
      CU(cudaEventElapsedTime(&ms, c->ev_pass[4], c->ev_pass[5]));
      c->stats.ms_k_align += ms; c->stats.ms_align += ms;
      align_pending = false;
    };
    // ONE host synchronisation per pass: everything the host decides on (convergence, table overflow, the size of the next
    // alignment batch, the next run list's size class) arrives in one page-locked struct; event times are read after it
    for (int pass = 0; pass < 400; pass++) {
      CU(cudaEventRecord(c->ev_pass[0], c->st));
      CU(cudaMemsetAsync(d_ctr, 0, 16, c->st));
      // which buckets run in this pass
      const bool partial = incremental && pass > 0 && last_diffs <= CHANGED_CAP;
      if (partial) {
        LAUNCH(c, k_dirty_from_unknown, nblk(n_ranks), 256, unk_flag, n_ranks, dirty);
        if (last_diffs) LAUNCH(c, k_mark_dirty_pairs, nblk(last_diffs), 256, changed, (uint32_t)last_diffs, rid_sorted, rank_sorted, n_elig, bloom, dirty);
        class_lists_dev(dirty, dlist);
        LAUNCH(c, k_e_carry, 1184, 256, S.E, (size_t)S.ecap, S.cur, dirty);
      } else {
        LAUNCH(c, k_e_fill_new, 1184, 256, S.E, (size_t)S.ecap, S.cur);
        CU(cudaMemcpyAsync(run_cnt, all_cnt, 8, cudaMemcpyDeviceToDevice, c->st));
      }
      CU(cudaEventRecord(c->ev_pass[2], c->st));
      if (partial) launch_replay(dlist, n_ranks, n_ranks + (tail_mode ? n_tail_big : n_all_big), wet ? 1 : 0, 0, (const uint32_t *)nullptr, (ovlp_rec *)nullptr, run_cnt);
      else launch_replay(all_list, n_all_small, n_all, wet ? 1 : 0, 0, (const uint32_t *)nullptr, (ovlp_rec *)nullptr, (const uint32_t *)nullptr);
      CU(cudaEventRecord(c->ev_pass[3], c->st));
      LAUNCH(c, k_e_diff_list, 1184, 256, S.E, (size_t)S.ecap, d_ctr + 1, changed, CHANGED_CAP);
      LAUNCH(c, k_pass_out, 1, 1, d_ctr, S.n_req, c->d_err, run_cnt, c->d_align_bases, d_pass);
      CU(cudaMemcpyAsync(c->h_pass, d_pass, sizeof(PassOut), cudaMemcpyDeviceToHost, c->st));
      CU(cudaEventRecord(c->ev_pass[1], c->st));
      CU(cudaStreamSynchronize(c->st));
      c->stats.d2h_bytes += sizeof(PassOut);
      const PassOut po = *c->h_pass;
      {
        float ms_rp = 0, ms_pass = 0;
        CU(cudaEventElapsedTime(&ms_rp, c->ev_pass[2], c->ev_pass[3]));
        CU(cudaEventElapsedTime(&ms_pass, c->ev_pass[0], c->ev_pass[1]));
        c->stats.ms_k_replay += ms_rp; c->stats.n_k_replay++;
        c->stats.ms_replay += ms_pass;
        read_align_timing();  // (the batch launched after the previous pass finished before this pass's replay started)
        if (verbose)
          fprintf(stderr, "pgb200: replay pass %d %s: buckets=%u/%u unknown=%llu table_diffs=%llu requests=%u  k_replay %.3f ms, pass %.3f ms\n", pass,
                  wet ? "wet" : "dry", po.n_run, n_ranks, po.unknown, po.diffs, po.n_req, ms_rp, ms_pass);
      }
      c->stats.n_replay_passes++;
      const unsigned long long ctr[2] = {po.unknown, po.diffs};
      const uint32_t n_req = po.n_req;
      if (po.err & (32 | 64)) {  // a table filled up: grow it and start over
        if (po.err & 32) { ecap64 *= 2; if (!ts_env) c->ecap_ratio *= 2; }
        if (po.err & 64) {  // n_req counted every request of the pass, also those that no longer fitted: size for them at once
          const uint64_t want = std::max(2 * acap64, 2 * (uint64_t)n_req);
          if (!ts_env) c->acap_ratio *= (double)want / (double)std::max<uint64_t>(acap64, 1);
          acap64 = want;
        }
        int z = 0;
        c->h2d(c->d_err, &z, sizeof z);
        c->sync();
        if (verbose) fprintf(stderr, "pgb200: replay tables too small (flags %d): restarting with pair table %llu, alignment table %llu\n", po.err,
                             (unsigned long long)ecap64, (unsigned long long)acap64);
        overflow = true;
        c->stats.n_replay_restarts++;
        break;
      }
      if (po.err) { c->check_err("pgb_overlap/replay"); free_S(); free_common(); return -1; }
      if (n_req > n_done) {
        uint32_t nn = n_req - n_done;
        uint32_t *perm = nullptr;
        CU(cudaEventRecord(c->ev_pass[4], c->st));
        if (nn > 8192 && nn > ALIGN_WARP_MAX) {  // group alignments of similar predicted length into the same warps
          uint32_t *keys = c->alloc<uint32_t>(nn), *keys2 = c->alloc<uint32_t>(nn), *idx0 = c->alloc<uint32_t>(nn);
          perm = c->alloc<uint32_t>(nn);
          // PGB_ALIGN_SORT: 0 = by predicted length only, m in 1..9 = (length >> (m-1), target read); see k_align_keys
          int sort_mode = getenv("PGB_ALIGN_SORT") ? atoi(getenv("PGB_ALIGN_SORT")) : 4;
          if (sort_mode < 0 || sort_mode > 9) sort_mode = 4;
          const int key_bits = sort_mode ? 31 : 8;
          LAUNCH(c, k_align_keys, nblk(nn), 256, S.reqs, n_done, nn, c->d_rlen_by_rid, keys, idx0, sort_mode);
          size_t tmp_bytes = 0;
          CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, tmp_bytes, keys, keys2, idx0, perm, (int)nn, 0, key_bits, c->st));
          uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
          CU(cub::DeviceRadixSort::SortPairs((void *)tmp, tmp_bytes, keys, keys2, idx0, perm, (int)nn, 0, key_bits, c->st));
          c->stats.kernel_launches += 3;
        }
        if (nn <= ALIGN_WARP_MAX)  // small batch: latency-bound, one warp per alignment
          LAUNCH(c, k_align_warp, nblk(nn, PGB_AW_WARPS), PGB_AW_WARPS * 32, S.reqs, n_done, nn, c->d_w, c->d_wrc, c->d_woff_by_rid, c->d_rlen_by_rid,
                 c->d_hasn_by_rid, (int)bw, S.ares, c->d_align_bases);
        else {
#define PGB_LEAN(P, T, MB)                                                                                                             \
  LAUNCH(c, (k_align_lean<P, T, MB>), nblk(nn, PGB_ALIGN_THREADS), PGB_ALIGN_THREADS, S.reqs, n_done, nn, perm, c->d_w, c->d_wrc, \
         c->d_woff_by_rid, c->d_rlen_by_rid, c->d_hasn_by_rid, (int)bw, S.ares, c->d_align_bases)
          // production form: band-row prefetch + register-cached band trim at 12 CTAs/SM (75 registers, no spills): 20.3 ms per
          // step against 26.2 ms for the plain form at 16 CTAs/SM, whose 64-register cap forces spills (profiles/r1g_ncu.md);
          // PGB_ALIGN_VARIANT=0 selects the plain form (kept as the A/B baseline of the parity tests)
          if (ALIGN_VARIANT == 0) PGB_LEAN(false, false, 16);
          else if (ALIGN_VARIANT >= 20 && ALIGN_VARIANT <= 23) {
            // G lanes per alignment, operand windows staged in shared memory (align_quad.cuh): 20 / 21 = G 4 / 8 with cp.async.bulk,
            // 22 / 23 = G 4 / 8 with cp.async
            const int G = (ALIGN_VARIANT & 1) ? 8 : 4;
            const bool bulk = ALIGN_VARIANT < 22;
            const int vcap_g = (int)bw + 3;
            const size_t smem = (size_t)(QA_THREADS / G) * sizeof(QaGroupSmem);
            auto kern = G == 4 ? (bulk ? k_align_quad<4, true> : k_align_quad<4, false>) : (bulk ? k_align_quad<8, true> : k_align_quad<8, false>);
            int occ = 1;
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, QA_THREADS, smem));
            const unsigned grid = std::min(nblk(nn, QA_THREADS / G), (unsigned)(c->sm_count * std::max(occ, 1)));
            unsigned int *qhead = c->alloc<unsigned int>(1);
            int *vscratch = c->alloc<int>((size_t)grid * (QA_THREADS / G) * 2 * vcap_g);
            CU(cudaMemsetAsync(qhead, 0, 4, c->st));
            kern<<<grid, QA_THREADS, smem, c->st>>>(S.reqs, n_done, nn, perm, c->d_w, c->d_wrc, c->n_words, c->d_woff_by_rid, c->d_rlen_by_rid,
                                                    c->d_hasn_by_rid, (int)bw, S.ares, c->d_align_bases, qhead, vscratch, vcap_g);
            c->stats.kernel_launches++;
            CU(cudaGetLastError());
          } else if (ALIGN_VARIANT == 30 || ALIGN_VARIANT == 31) {
            // G lanes per alignment, branch-free loop, operands read through L1 (align_coop.cuh): 30 = G 4, 31 = G 8
            const int G = ALIGN_VARIANT == 30 ? 4 : 8;
            const int vcap_g = (int)bw + 3;
            auto kern = G == 4 ? k_align_coop<4> : k_align_coop<8>;
            int occ = 1;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, QC_THREADS, 0));
            const unsigned grid = std::min(nblk(nn, QC_THREADS / G), (unsigned)(c->sm_count * std::max(occ, 1)));
            unsigned int *qhead = c->alloc<unsigned int>(1);
            int *vscratch = c->alloc<int>((size_t)grid * (QC_THREADS / G) * 2 * vcap_g);
            CU(cudaMemsetAsync(qhead, 0, 4, c->st));
            kern<<<grid, QC_THREADS, 0, c->st>>>(S.reqs, n_done, nn, perm, c->d_w, c->d_wrc, c->d_woff_by_rid, c->d_rlen_by_rid, c->d_hasn_by_rid, (int)bw,
                                                 S.ares, c->d_align_bases, qhead, vscratch, vcap_g);
            c->stats.kernel_launches++;
            CU(cudaGetLastError());
          } else PGB_LEAN(true, true, 12);
#undef PGB_LEAN
        }
        if (c->n_reads_with_n)
          LAUNCH(c, k_align, nblk(nn, 64), 64, S.reqs, n_done, nn, perm, c->d_w, c->d_nm, c->d_woff_by_rid, c->d_rlen_by_rid, c->d_hasn_by_rid,
                 (int)bw, S.ares, c->d_err, c->d_align_bases, 1);
        CU(cudaEventRecord(c->ev_pass[5], c->st));
        align_pending = true;
        c->stats.n_k_align++;
        c->stats.n_alignments += nn;
      }
      const bool new_requests = n_req > n_done;
      n_done = n_req;
      S.cur ^= 1;
      last_diffs = ctr[1];
      tail_mode = po.n_run < TAIL_RUN;
      c->stats.n_replay_buckets += po.n_run;
      if (wet && !new_requests && ctr[0] == 0 && ctr[1] == 0) { converged = true; break; }
      if (!wet) {
        // speculative ("dry") passes settle the time-stamped pair table with predicted alignments only; switch to real
        // alignments once the table is nearly stable or stops improving
        dry_passes++;
        if (ctr[1] <= 16 + ctr[0] / 512 || ctr[1] >= prev_diffs || dry_passes >= MAX_DRY) wet = true;
        prev_diffs = ctr[1];
        if (wet && acap64 == 0) {  // first wet pass next: size the alignment cache from this pass's count of predicted alignments
          acap64 = (uint64_t)(2.0 * tscale * (double)ctr[0]) + (ts_env ? 64 : 8192);
          alloc_aln_tables(acap64);
        }
      }
    }
    if (align_pending) { c->sync(); read_align_timing(); }
    {
      unsigned long long ab = 0;  // bases compared along the final paths of this call's alignments (algorithmic-bytes accounting)
      c->d2h(&ab, c->d_align_bases, 8);
      c->stats.n_align_bases += ab;
      CU(cudaMemsetAsync(c->d_align_bases, 0, 8, c->st));
    }
    if (overflow) continue;
    if (!converged) { free_S(); free_common(); throw std::runtime_error("replay fix-point did not converge in 400 passes"); }

    // ---------------- emission pass in visiting order
    c->tic();
    uint32_t n_out = scan_u32(c, acc, out_off, (size_t)n_ranks + 1);
    c->d_ovl = c->palloc<ovlp_rec>(n_out); c->n_ovl = n_out;
    LAUNCH(c, k_e_fill_new, 1184, 256, S.E, (size_t)S.ecap, S.cur);
    CU(cudaMemsetAsync(d_ctr, 0, 16, c->st));
    launch_replay(all_list, n_all_small, n_all, 0, 1, out_off, c->d_ovl, (const uint32_t *)nullptr);
    c->sync();
    c->stats.ms_emit += c->toc();
    c->stats.n_overlaps += n_out;
    free_S();
  }
  free_common();
  c->sync();
  c->check_err("pgb_overlap/emit");
  return c->err.empty() ? 0 : -1;
}

// rid_pairs is per chunk: with T chunks a read pair is aligned (and accepted) in up to T of them, so the pair table of ONE
// chunk grows relative to its eligible records (measured x1.85 records at T=2, SURVEY 6.2: x4.5 at T=8).  Start large enough
// to spare the fix-point a restart; an underestimate only costs that restart.  (The alignment cache is sized from the
// speculative passes' own count of missing alignments, see overlap_core; acap_ratio only covers PGB_DRY_PASSES=0.)
static void size_tables_for_chunks(pgb_ctx *c, uint32_t T) {
  const double t = (double)std::min<uint32_t>(std::max<uint32_t>(T, 1), 16);
  c->acap_ratio = std::max(c->acap_ratio, 0.75 * pow(t, 0.75));
  c->ecap_ratio = std::max(c->ecap_ratio, 1.25 * std::max(1.0, t / 4.0));
}

extern "C" int pgb_overlap(pgb_ctx *c, uint32_t T, uint32_t mychunk, uint32_t bestn, uint32_t mc_lower, uint32_t mc_upper,
                           uint32_t bw, uint32_t ovlp_upper) {
  API_BEGIN(c)
  if (T == 0 || mychunk == 0 || mychunk > T) throw std::runtime_error("bad chunk spec");
  if (!c->d_w) throw std::runtime_error("no reads loaded");
  if (!c->d_shm && c->n_shm) throw std::runtime_error("no shimmers set");
  c->release(c->d_ovl); c->n_ovl = 0;
  if (c->n_shm >= (1ull << 31)) throw std::runtime_error("more than 2^31 shimmers in one overlap call");
  if (c->n_shm != 0) {
  ensure_loaded(c);
  size_tables_for_chunks(c, T);
  PairSoA R;
  uint32_t nrec = build_pair_records(c, T, mychunk, mc_lower, mc_upper, R);
  overlap_core(c, R, nrec, bestn, bw, ovlp_upper);  // failures are left in c->err; API_END syncs, resets the arena and reports them
  }
  API_END(c)
}

// ================================================================================================ routed exchange (multi-GPU)
// Rank r holds the shimmers of ITS reads only.  (1) pgb_counts_dump: its partial multiplicity table as (mer,count) entries;
// the ranks all-gather them and pgb_counts_set_device sums them (aggregate_mm_count over the -MC- files, src/shmr_utils.c:162-176).
// (2) pgb_route_scan / pgb_route_build: build_map over the rank's own list, emitting the records of EVERY hash chunk grouped by
// owner chunk; one all-to-all moves each group to its owner.  (3) pgb_overlap_routed: the owner's records, source ranks
// concatenated in rank order (= the insertion order of the reference's single scan over the concatenated chunk files).
extern "C" int pgb_counts_dump(pgb_ctx *c, size_t *n_out) {
  API_BEGIN(c)
  if (!c->d_mckeys) throw std::runtime_error("no shimmers set");
  c->release(c->d_mc_dump); c->n_mc_dump = 0;
  size_t cap = (size_t)c->mcmask + 1;
  uint32_t *flags = c->alloc<uint32_t>(cap + 1), *pos = c->alloc<uint32_t>(cap + 1);
  CU(cudaMemsetAsync(flags, 0, (cap + 1) * 4, c->st));
  LAUNCH(c, k_mc_flags, nblk(cap), 256, c->d_mckeys, cap, flags);
  uint32_t distinct = scan_u32(c, flags, pos, cap + 1);
  c->d_mc_dump = c->palloc<mc_entry>(distinct);
  LAUNCH(c, k_mc_dump, nblk(cap), 256, c->d_mckeys, c->d_mcvals, pos, cap, c->d_mc_dump);
  c->n_mc_dump = distinct;
  if (n_out) *n_out = distinct;
  API_END(c)
}
extern "C" int pgb_counts_set_device(pgb_ctx *c, const mm_count_t *entries_device, size_t n) {
  API_BEGIN(c)
  c->tic();
  build_mc_table(c, (const mc_entry *)entries_device, n);
  c->stats.ms_count += c->toc();
  c->check_err("pgb_counts_set_device");
  API_END(c)
}
extern "C" int pgb_route_scan(pgb_ctx *c, uint32_t mc_lower, uint32_t mc_upper, int *has_first) {
  API_BEGIN(c)
  if (!c->d_mckeys) throw std::runtime_error("no multiplicity table set");
  c->release(c->d_route_cnt);
  size_t n = c->n_shm;
  c->tic();
  c->d_route_cnt = c->palloc<uint32_t>(n);
  unsigned long long *d_first = c->alloc<unsigned long long>(1);
  CU(cudaMemsetAsync(d_first, 0xFF, 8, c->st));
  LAUNCH(c, k_count_lookup, nblk(n), 256, c->d_shm, n, c->d_mckeys, c->d_mcvals, c->mcmask, c->d_route_cnt, mc_lower, mc_upper, d_first, c->d_err);
  c->d2h(&c->route_first, d_first, 8);
  c->stats.ms_pairs += c->toc();
  if (has_first) *has_first = c->route_first != ~0ULL;
  c->check_err("pgb_route_scan");
  API_END(c)
}
extern "C" int pgb_route_build(pgb_ctx *c, uint32_t T, uint32_t mc_lower, uint32_t mc_upper, int first_found_before, uint64_t *n_per_chunk) {
  API_BEGIN(c)
  if (T == 0 || T > 4096) throw std::runtime_error("bad chunk count");
  if (!c->d_route_cnt && c->n_shm) throw std::runtime_error("pgb_route_scan has not run");
  if (!c->d_rlen_by_rid) throw std::runtime_error("no reads loaded");
  c->release(c->d_route); c->n_route = 0;
  for (uint32_t t = 0; t < T; t++) n_per_chunk[t] = 0;
  size_t n = c->n_shm;
  if (n != 0) {
  c->tic();
  uint32_t *flags = c->alloc<uint32_t>(n + 1), *pos = c->alloc<uint32_t>(n + 1);
  CU(cudaMemsetAsync(flags, 0, (n + 1) * 4, c->st));
  if (first_found_before) {
    LAUNCH(c, k_kept_flags_nonfirst, nblk(n), 256, c->d_route_cnt, n, mc_lower, mc_upper, flags);
  } else {
    unsigned long long *d_first = c->alloc<unsigned long long>(1);
    c->h2d(d_first, &c->route_first, 8);
    LAUNCH(c, k_kept_flags, nblk(n), 256, c->d_route_cnt, n, mc_lower, mc_upper, d_first, flags);
  }
  uint32_t n_kept = scan_u32(c, flags, pos, n + 1);
  uint32_t *kept = c->alloc<uint32_t>(n_kept);
  LAUNCH(c, k_compact_idx, nblk(n), 256, flags, pos, n, kept);
  uint32_t *n_rec = c->alloc<uint32_t>((size_t)n_kept + 1), *rec_off = c->alloc<uint32_t>((size_t)n_kept + 1);
  CU(cudaMemsetAsync(n_rec, 0, ((size_t)n_kept + 1) * 4, c->st));
  LAUNCH(c, k_pair_count_all, nblk(n_kept), 256, c->d_shm, kept, n_kept, n_rec);
  uint32_t nrec = scan_u32(c, n_rec, rec_off, (size_t)n_kept + 1);
  route_rec *raw = c->alloc<route_rec>(nrec);
  uint32_t *dest = c->alloc<uint32_t>(nrec), *dest2 = c->alloc<uint32_t>(nrec), *idx = c->alloc<uint32_t>(nrec), *perm = c->alloc<uint32_t>(nrec);
  uint32_t *per_dest = c->alloc<uint32_t>(T);
  CU(cudaMemsetAsync(per_dest, 0, (size_t)T * 4, c->st));
  LAUNCH(c, k_pair_write_all, nblk(n_kept), 256, c->d_shm, kept, n_kept, T, rec_off, c->d_rlen_by_rid, raw, dest, idx, per_dest);
  c->d_route = c->palloc<route_rec>(nrec);
  if (nrec) {
    int bits = 1;
    while ((1u << bits) < T) bits++;
    size_t tmp_bytes = 0;  // stable: records of one destination keep their scan (= insertion) order
    CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, tmp_bytes, dest, dest2, idx, perm, (int)nrec, 0, bits, c->st));
    uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
    CU(cub::DeviceRadixSort::SortPairs((void *)tmp, tmp_bytes, dest, dest2, idx, perm, (int)nrec, 0, bits, c->st));
    c->stats.kernel_launches += 3;
    LAUNCH(c, k_route_gather, nblk(nrec), 256, raw, perm, nrec, (route_rec *)c->d_route);
  }
  std::vector<uint32_t> h(T);
  c->d2h(h.data(), per_dest, (size_t)T * 4);
  for (uint32_t t = 0; t < T; t++) n_per_chunk[t] = h[t];
  c->n_route = nrec;
  c->release(c->d_route_cnt);
  c->stats.ms_pairs += c->toc();
  c->stats.n_pair_records += nrec;
  c->check_err("pgb_route_build");
  }
  API_END(c)
}
extern "C" int pgb_overlap_routed(pgb_ctx *c, const void *records_device, size_t n, uint32_t bestn, uint32_t bw, uint32_t ovlp_upper,
                                  uint32_t T) {
  API_BEGIN(c)
  if (!c->d_w) throw std::runtime_error("no reads loaded");
  if (n >= (1ull << 32)) throw std::runtime_error("more than 2^32 pair records in one overlap call");
  c->release(c->d_ovl); c->n_ovl = 0;
  ensure_loaded(c);
  size_tables_for_chunks(c, T);
  uint32_t nrec = (uint32_t)n;
  PairSoA R;
  c->tic();
  R.k0 = c->alloc<uint64_t>(nrec); R.k1 = c->alloc<uint64_t>(nrec); R.y0 = c->alloc<uint64_t>(nrec); R.y1 = c->alloc<uint64_t>(nrec);
  R.seq = c->alloc<uint32_t>(nrec); R.dir = c->alloc<uint8_t>(nrec);
  LAUNCH(c, k_route_unpack, nblk(nrec), 256, (const route_rec *)records_device, nrec, R);
  c->stats.ms_pairs += c->toc();
  overlap_core(c, R, nrec, bestn, bw, ovlp_upper);
  API_END(c)
}

extern "C" size_t pgb_overlap_size(pgb_ctx *c) { return c ? c->n_ovl : 0; }
extern "C" int pgb_overlap_copy(pgb_ctx *c, ovlp_t *out) {
  API_BEGIN(c)
  c->d2h(out, c->d_ovl, c->n_ovl * sizeof(ovlp_rec));
  API_END(c)
}
extern "C" int pgb_overlap_host(pgb_ctx *c, const ovlp_t **out, size_t *n) {
  API_BEGIN(c)
  const size_t bytes = c->n_ovl * sizeof(ovlp_rec);
  if (bytes > c->h_ovl_cap) {
    if (c->h_ovl) cudaFreeHost(c->h_ovl);
    c->h_ovl = nullptr; c->h_ovl_cap = 0;
    const size_t cap = bytes + bytes / 4 + (1 << 20);
    CU(cudaMallocHost((void **)&c->h_ovl, cap));
    c->h_ovl_cap = cap;
  }
  c->d2h(c->h_ovl, c->d_ovl, bytes);
  *out = (const ovlp_t *)c->h_ovl;
  *n = c->n_ovl;
  API_END(c)
}

// ================================================================================================ shmr_dedup
// first record of every unordered read pair, in stream order, as preads.ovl text (src/shmr_dedup.c:19-101)
static void dedup_table_free(pgb_ctx *c) {
  c->release(c->dd_keys); c->release(c->dd_first);
  c->dd_cap = 0; c->dd_base = 0; c->dd_occ = 0; c->dd_open = false;
}
// room for `n` more pairs at load <= 0.5: allocate, or grow and move the entries over
static void dedup_table_reserve(pgb_ctx *c, size_t n) {
  const uint64_t want = 2 * (c->dd_occ + (uint64_t)n) + 16;
  if (want > (1ull << 31)) throw std::runtime_error("shmr_dedup: more than 2^30 distinct read pairs in one stream");
  if (c->dd_cap && want <= c->dd_cap) return;
  const uint32_t cap = pow2_at_least(c->dd_cap ? std::max<uint64_t>(want, 2ull * c->dd_cap) : want);
  uint64_t *keys = c->palloc<uint64_t>(cap);
  unsigned long long *first = c->palloc<unsigned long long>(cap);
  LAUNCH(c, k_fill_u64, 1184, 256, keys, PGB_EMPTY, (size_t)cap);
  CU(cudaMemsetAsync(first, 0xFF, (size_t)cap * 8, c->st));
  if (c->dd_cap) LAUNCH(c, k_dedup_rehash, nblk(c->dd_cap), 256, c->dd_keys, c->dd_first, (size_t)c->dd_cap, keys, cap - 1, first, c->d_err);
  c->sync();
  c->release(c->dd_keys); c->release(c->dd_first);
  c->dd_keys = keys; c->dd_first = first; c->dd_cap = cap;
}
// one batch of the stream: the kept records' lines become the context's dedup text
static void dedup_push(pgb_ctx *c, const ovlp_rec *d_recs, size_t n) {
  c->release(c->d_dedup_text); c->dedup_bytes = 0; c->dedup_kept = 0;
  if (n == 0) return;
  if (n >= (1ull << 30)) throw std::runtime_error("more than 2^30 records in one dedup batch");
  c->tic();
  dedup_table_reserve(c, n);
  unsigned long long *d_kept = c->alloc<unsigned long long>(2);
  uint32_t *slot_of = c->alloc<uint32_t>(n), *len = c->alloc<uint32_t>(n + 1);
  uint64_t *off = c->alloc<uint64_t>(n + 1);
  CU(cudaMemsetAsync(d_kept, 0, 16, c->st));
  CU(cudaMemsetAsync(len + n, 0, 4, c->st));
  LAUNCH(c, k_dedup_insert, nblk(n), 256, d_recs, n, c->dd_keys, c->dd_cap - 1, c->dd_first, slot_of, c->d_err, c->dd_base);
  LAUNCH(c, k_dedup_len, nblk(n), 256, d_recs, n, c->dd_first, slot_of, len, d_kept, c->dd_base);
  const uint64_t bytes = scan_u32_to_u64(c, len, off, n + 1);
  c->d_dedup_text = c->palloc<char>(bytes);
  LAUNCH(c, k_dedup_write, nblk(n), 256, d_recs, n, len, off, c->d_dedup_text);
  unsigned long long kept = 0;
  c->d2h(&kept, d_kept, 8);
  c->dedup_bytes = bytes; c->dedup_kept = kept;
  c->dd_base += n;
  c->dd_occ += kept;  // every kept record is the first of a pair the table did not hold
  c->stats.ms_dedup += c->toc();
  c->stats.n_dedup_in += n; c->stats.n_dedup_kept += kept;
  c->check_err("pgb_dedup");
}
// a whole stream in one call
static void dedup_core(pgb_ctx *c, const ovlp_rec *d_recs, size_t n) {
  dedup_table_free(c);
  dedup_push(c, d_recs, n);
  dedup_table_free(c);
}
extern "C" int pgb_dedup(pgb_ctx *c, const ovlp_t *records, size_t n) {
  API_BEGIN(c)
  ovlp_rec *d = c->alloc<ovlp_rec>(n);
  c->h2d(d, records, n * sizeof(ovlp_rec));
  dedup_core(c, d, n);
  API_END(c)
}
extern "C" int pgb_dedup_device(pgb_ctx *c, const ovlp_t *records_device, size_t n) {
  API_BEGIN(c)
  dedup_core(c, (const ovlp_rec *)records_device, n);
  API_END(c)
}
extern "C" int pgb_dedup_overlaps(pgb_ctx *c) {
  API_BEGIN(c)
  dedup_core(c, c->d_ovl, c->n_ovl);
  API_END(c)
}
// the stream in bounded batches: begin, push x N (each push leaves the lines of ITS kept records, read with pgb_dedup_text_*), end
extern "C" int pgb_dedup_stream_begin(pgb_ctx *c) {
  API_BEGIN(c)
  dedup_table_free(c);
  c->dd_open = true;
  API_END(c)
}
extern "C" int pgb_dedup_stream_push(pgb_ctx *c, const ovlp_t *records, size_t n) {
  API_BEGIN(c)
  if (!c->dd_open) throw std::runtime_error("pgb_dedup_stream_push without pgb_dedup_stream_begin");
  ovlp_rec *d = c->alloc<ovlp_rec>(n);
  c->h2d(d, records, n * sizeof(ovlp_rec));
  dedup_push(c, d, n);
  API_END(c)
}
extern "C" int pgb_dedup_stream_end(pgb_ctx *c) {
  API_BEGIN(c)
  dedup_table_free(c);
  c->release(c->d_dedup_text); c->dedup_bytes = 0;
  API_END(c)
}
extern "C" size_t pgb_dedup_kept(pgb_ctx *c) { return c ? c->dedup_kept : 0; }
extern "C" size_t pgb_dedup_text_bytes(pgb_ctx *c) { return c ? c->dedup_bytes : 0; }
extern "C" int pgb_dedup_text_copy(pgb_ctx *c, char *out) {
  API_BEGIN(c)
  c->d2h(out, c->d_dedup_text, c->dedup_bytes);
  API_END(c)
}

// ================================================================================================ shmr_map
// read lengths only (build_map mirrors the coordinates of reverse records, src/shmr_utils.c:376-395); no sequence is loaded
extern "C" int pgb_set_read_lengths(pgb_ctx *c, const uint32_t *rid, const uint32_t *len, size_t n_reads) {
  API_BEGIN(c)
  c->free_reads();
  uint32_t max_rid = 0;
  for (size_t i = 0; i < n_reads; i++) if (rid[i] > max_rid) max_rid = rid[i];
  if (n_reads && (uint64_t)max_rid > 64 * (uint64_t)n_reads + (1u << 24)) throw std::runtime_error("read ids too sparse");
  std::vector<uint32_t> by_rid((size_t)max_rid + 1, 0);
  for (size_t i = 0; i < n_reads; i++) by_rid[rid[i]] = len[i];
  c->max_rid = max_rid;
  c->d_rlen_by_rid = c->palloc<uint32_t>(by_rid.size());
  c->h2d(c->d_rlen_by_rid, by_rid.data(), by_rid.size() * 4);
  API_END(c)
}
// process_map (src/shmr_map.c:48-166): contig shimmers against the pair index of the context's shimmers
extern "C" int pgb_map(pgb_ctx *c, const mm128_t *ref_mmers, size_t n_ref, uint32_t T, uint32_t mychunk, uint32_t mc_lower, uint32_t mc_upper) {
  API_BEGIN(c)
  if (T == 0 || mychunk == 0 || mychunk > T) throw std::runtime_error("bad chunk spec");
  if (!c->d_rlen_by_rid) throw std::runtime_error("no read lengths loaded");
  if (!c->d_shm && c->n_shm) throw std::runtime_error("no shimmers set");
  if (c->n_shm >= (1ull << 31) || n_ref >= (1ull << 31)) throw std::runtime_error("more than 2^31 shimmers in one map call");
  c->release(c->d_map_text); c->map_bytes = 0; c->map_hits = 0;
  if (n_ref != 0 && c->n_shm != 0) {
  // ---- pair index of the reads: records, X / B tables, records grouped by bucket in insertion order
  PairSoA R;
  const uint32_t nrec = build_pair_records(c, T, mychunk, mc_lower, mc_upper, R);
  c->tic();
  if (nrec) {
    const uint32_t xcap = pow2_at_least(4 * (uint64_t)nrec), bcap = pow2_at_least(2 * (uint64_t)nrec);
    uint64_t *xkeys = c->alloc<uint64_t>(xcap), *bkeys = c->alloc<uint64_t>(bcap);
    uint32_t *bcount = c->alloc<uint32_t>((size_t)bcap + 1), *bfirst = c->alloc<uint32_t>(bcap), *blast = c->alloc<uint32_t>(bcap);
    uint32_t *rec_bucket = c->alloc<uint32_t>(nrec), *xfirst = c->alloc<uint32_t>(xcap), *bstart = c->alloc<uint32_t>((size_t)bcap + 1);
    LAUNCH(c, k_fill_u64, 1184, 256, xkeys, PGB_EMPTY, (size_t)xcap);
    LAUNCH(c, k_fill_u64, 1184, 256, bkeys, PGB_EMPTY, (size_t)bcap);
    CU(cudaMemsetAsync(bcount, 0, ((size_t)bcap + 1) * 4, c->st));
    CU(cudaMemsetAsync(bfirst, 0xFF, (size_t)bcap * 4, c->st));
    CU(cudaMemsetAsync(blast, 0, (size_t)bcap * 4, c->st));
    CU(cudaMemsetAsync(xfirst, 0xFF, (size_t)xcap * 4, c->st));
    LAUNCH(c, k_bucket_insert, nblk(nrec), 256, R, nrec, xkeys, xcap - 1, bkeys, bcap - 1, bcount, bfirst, blast, xfirst, rec_bucket, c->d_err);
    scan_u32(c, bcount, bstart, (size_t)bcap + 1);
    uint32_t *iota = c->alloc<uint32_t>(nrec), *bk_sorted = c->alloc<uint32_t>(nrec), *by_bucket = c->alloc<uint32_t>(nrec);
    LAUNCH(c, k_iota_u32, nblk(nrec), 256, iota, nrec);
    sort_pairs_u32(c, rec_bucket, bk_sorted, iota, by_bucket, nrec);  // stable: records of a bucket stay in insertion order
    if (c->check_err("pgb_map/buckets")) throw std::runtime_error(c->err);
    // ---- contig walk
    mm128 *d_ref = c->alloc<mm128>(n_ref);
    c->h2d(d_ref, ref_mmers, n_ref * sizeof(mm128));
    uint32_t *cnt = c->alloc<uint32_t>(n_ref), *flags = c->alloc<uint32_t>(n_ref + 1), *pos = c->alloc<uint32_t>(n_ref + 1);
    unsigned long long *d_first = c->alloc<unsigned long long>(1);
    CU(cudaMemsetAsync(d_first, 0xFF, 8, c->st));
    CU(cudaMemsetAsync(flags, 0, (n_ref + 1) * 4, c->st));
    LAUNCH(c, k_map_ref_flags, nblk(n_ref), 256, d_ref, n_ref, c->d_mckeys, c->d_mcvals, c->mcmask, xkeys, xcap - 1, xfirst, cnt, d_first);
    LAUNCH(c, k_map_kept, nblk(n_ref), 256, cnt, n_ref, mc_lower, mc_upper, d_first, flags);
    const uint32_t n_kept = scan_u32(c, flags, pos, n_ref + 1);
    if (n_kept > 1) {
      uint32_t *kept = c->alloc<uint32_t>(n_kept), *n_hits = c->alloc<uint32_t>((size_t)n_kept + 1), *pair_bucket = c->alloc<uint32_t>(n_kept);
      uint64_t *hit_off = c->alloc<uint64_t>((size_t)n_kept + 1);
      LAUNCH(c, k_compact_idx, nblk(n_ref), 256, flags, pos, n_ref, kept);
      CU(cudaMemsetAsync(n_hits + n_kept, 0, 4, c->st));
      LAUNCH(c, k_map_pair_count, nblk(n_kept), 256, d_ref, kept, n_kept, xkeys, xcap - 1, bkeys, bcap - 1, bcount, n_hits, pair_bucket);
      const uint64_t nh = scan_u32_to_u64(c, n_hits, hit_off, (size_t)n_kept + 1);
      if (nh >= (1ull << 31)) throw std::runtime_error("more than 2^31 map hits in one call");
      if (nh) {
        map_hit *hits = c->alloc<map_hit>(nh);
        LAUNCH(c, k_map_emit, nblk(n_kept), 256, d_ref, kept, n_kept, cnt, pair_bucket, n_hits, hit_off, bstart, by_bucket, R, hits);
        uint32_t *len = c->alloc<uint32_t>(nh + 1);
        uint64_t *off = c->alloc<uint64_t>(nh + 1);
        CU(cudaMemsetAsync(len + nh, 0, 4, c->st));
        LAUNCH(c, k_map_len, nblk(nh), 256, hits, (size_t)nh, len);
        const uint64_t bytes = scan_u32_to_u64(c, len, off, nh + 1);
        c->d_map_text = c->palloc<char>(bytes);
        LAUNCH(c, k_map_write, nblk(nh), 256, hits, (size_t)nh, off, c->d_map_text);
        c->map_bytes = bytes; c->map_hits = nh;
      }
    }
  }
  c->stats.ms_map += c->toc();
  c->stats.n_map_hits += c->map_hits;
  c->check_err("pgb_map");
  }
  API_END(c)
}
extern "C" size_t pgb_map_hits(pgb_ctx *c) { return c ? c->map_hits : 0; }
extern "C" size_t pgb_map_text_bytes(pgb_ctx *c) { return c ? c->map_bytes : 0; }
extern "C" int pgb_map_text_copy(pgb_ctx *c, char *out) {
  API_BEGIN(c)
  c->d2h(out, c->d_map_text, c->map_bytes);
  API_END(c)
}

// ================================================================================================ shmr_mkseqdb
// encode_biseq over a batch of reads (src/shmr_utils.c:44-51): ascii -> .seqdb bytes at the same offsets
extern "C" int pgb_encode_biseq(pgb_ctx *c, const char *ascii, size_t total_bytes, const uint64_t *offset, const uint32_t *len, size_t n_reads,
                                uint8_t *seqdb_out) {
  API_BEGIN(c)
  std::vector<EncTile> tiles;
  for (size_t i = 0; i < n_reads; i++) {
    if (offset[i] + len[i] > total_bytes) throw std::runtime_error("read extends past the end of the ascii buffer");
    for (uint32_t p0 = 0; p0 < len[i]; p0 += ENC_TILE) tiles.push_back(EncTile{offset[i], len[i], p0});
  }
  if (tiles.size() >= (1ull << 31)) throw std::runtime_error("too many tiles in one pgb_encode_biseq call");
  c->tic();
  uint8_t *d_in = c->alloc<uint8_t>(total_bytes), *d_out = c->alloc<uint8_t>(total_bytes);
  EncTile *d_tiles = c->alloc<EncTile>(tiles.size());
  c->h2d(d_in, ascii, total_bytes);
  c->h2d(d_tiles, tiles.data(), tiles.size() * sizeof(EncTile));
  CU(cudaMemsetAsync(d_out, 0, total_bytes, c->st));  // bytes between reads (none when the reads are contiguous)
  c->ktic();
  LAUNCH(c, k_encode_biseq, (unsigned)tiles.size(), 256, d_in, d_tiles, d_out);
  c->stats.ms_k_encode += c->ktoc(); c->stats.n_k_encode++;
  c->d2h(seqdb_out, d_out, total_bytes);
  c->stats.ms_encode += c->toc();
  c->stats.bases_encoded += total_bytes;
  API_END(c)
}

// ================================================================================================ command-line tools
static pgb_ctx *cli_ctx() {
  int dev = 0;
  const char *e = getenv("PGB_DEVICE");
  if (e) dev = atoi(e);
  pgb_ctx *c = pgb_create(dev);
  if (!c) exit(1);
  return c;
}
#define CLI_CHECK(c, call)                                              \
  do {                                                                  \
    if ((call) != 0) {                                                  \
      fprintf(stderr, "pgb200: %s failed: %s\n", #call, pgb_last_error(c)); \
      exit(1);                                                          \
    }                                                                   \
  } while (0)

extern "C" int pgb_shmr_index_main(int argc, char **argv) {
  // option handling, defaults, assertions and messages follow src/shmr_index.c:37-131
  const char *seqdb_prefix = "seq_dataset", *shimmer_prefix = "shimmer";
  int total_chunk = 1, mychunk = 1, reduction_factor = 6, number_layers = 2, output_L0 = 1, window_size = 80, kmer_size = 16;
  int ch;
  opterr = 0; optind = 1;
  while ((ch = getopt(argc, argv, "p:o:t:c:l:r:m:w:k:")) != -1) {
    switch (ch) {
      case 'p': seqdb_prefix = optarg; break;
      case 'o': shimmer_prefix = optarg; break;
      case 't': total_chunk = atoi(optarg); break;
      case 'c': mychunk = atoi(optarg); break;
      case 'r': reduction_factor = atoi(optarg); break;
      case 'l': number_layers = atoi(optarg); break;
      case 'm': output_L0 = atoi(optarg); break;
      case 'w': window_size = atoi(optarg); break;
      case 'k': kmer_size = atoi(optarg); break;
      case '?':
        if (optopt == 'p') fprintf(stderr, "Option -%c not specified, using 'seq_dataset' as the sequence db prefix\n", optopt);
        else if (optopt == 'o') fprintf(stderr, "Option -%c not specified, using 'shimmer' as the output prefix\n", optopt);
        return 1;
      default: abort();
    }
  }
  assert(total_chunk > 0);
  assert(mychunk > 0 && mychunk <= total_chunk);
  assert(reduction_factor < 256);
  assert(window_size >= 24 && kmer_size >= 12 && window_size > kmer_size);
  fprintf(stderr, "reduction factor= %d\n", reduction_factor);
  char path[8400];
  snprintf(path, sizeof path, "%s.idx", seqdb_prefix);
  fprintf(stderr, "using index file: %s\n", path);
  snprintf(path, sizeof path, "%s.seqdb", seqdb_prefix);
  fprintf(stderr, "using seqdb file: %s\n", path);

  const bool vt = getenv("PGB_VERBOSE") != nullptr;
  const double t_0 = now_ms();
  pgb_ctx *c = cli_ctx();
  const double t_ctx = now_ms();
  CLI_CHECK(c, pgb_load_reads_from_files(c, seqdb_prefix, (uint32_t)total_chunk, (uint32_t)mychunk, 0));
  const double t_load = now_ms();
  int levels = number_layers == 1 ? 1 : (number_layers > 1 ? 2 : 0);
  int with_counts = (output_L0 == 1 ? 1 : 0) | (levels == 1 ? 2 : 0) | (levels == 2 ? 4 : 0);
  CLI_CHECK(c, pgb_index(c, window_size, kmer_size, reduction_factor, levels, with_counts));
  const double t_index = now_ms();
  auto write_level = [&](int level, const char *tag, bool data_msg_stderr) {
    std::vector<mm128_t> v(pgb_index_size(c, level));
    CLI_CHECK(c, pgb_index_copy(c, level, v.data()));
    snprintf(path, sizeof path, "%s-%s-%02d-of-%02d.dat", shimmer_prefix, tag, mychunk, total_chunk);
    if (data_msg_stderr) fprintf(stderr, "output data file: %s\n", path); else printf("output data file: %s\n", path);
    write_mmlist_file(path, (const mm128 *)v.data(), v.size());
    std::vector<mm_count_t> mc(pgb_index_count_size(c, level));
    pgb_index_count_copy(c, level, mc.data());
    snprintf(path, sizeof path, "%s-%s-MC-%02d-of-%02d.dat", shimmer_prefix, tag, mychunk, total_chunk);
    printf("output data file: %s\n", path);
    write_mc_file(path, (const mc_rec *)mc.data(), mc.size());
  };
  if (output_L0 == 1) write_level(0, "L0", true);   // src/shmr_index.c:165-197
  if (levels == 1) write_level(1, "L1", false);      // :201-214
  if (levels == 2) write_level(2, "L2", true);       // :216-232
  if (vt) fprintf(stderr, "pgb200: shmr_index timing: context %.0f ms, load %.0f ms, index %.0f ms, write %.0f ms\n", t_ctx - t_0, t_load - t_ctx,
                  t_index - t_load, now_ms() - t_index);
  pgb_destroy(c);
  return 0;
}

extern "C" int pgb_shmr_overlap_main(int argc, char **argv) {
  // option handling, defaults and messages follow src/shmr_overlap.c:233-392
  const char *seqdb_prefix = "seq_dataset", *shimmer_prefix = "shimmer-L2";
  char default_out[64];
  const char *ovlp_file_path = nullptr;
  uint8_t bestn = 4;
  uint32_t mc_lower = 2, mc_upper = 240, ovlp_upper = 120, align_bandwidth = 100, total_chunk = 1, mychunk = 1;
  int ch;
  opterr = 0; optind = 1;
  while ((ch = getopt(argc, argv, "p:l:t:c:b:o:m:M:w:n:")) != -1) {
    switch (ch) {
      case 'p': seqdb_prefix = optarg; break;
      case 'l': shimmer_prefix = optarg; break;
      case 't': total_chunk = atoi(optarg); break;
      case 'c': mychunk = atoi(optarg); break;
      case 'b': bestn = (uint8_t)atoi(optarg); break;
      case 'o': ovlp_file_path = optarg; break;
      case 'm': mc_lower = atoi(optarg); break;
      case 'M': mc_upper = atoi(optarg); break;
      case 'w': align_bandwidth = atoi(optarg); break;
      case 'n': ovlp_upper = atoi(optarg); break;
      case '?':
        if (optopt == 'p') fprintf(stderr, "Option -%c not specified, using 'seq_dataset' as the sequence db prefix\n", optopt);
        if (optopt == 'l') fprintf(stderr, "Option -%c not specified, using 'shimmer-L2' as the L2 index prefix\n", optopt);
        if (optopt == 'o') fprintf(stderr, "Option -%c not specified, using 'ovlp.# as default output' \n", optopt);
        return 1;
      default: abort();
    }
  }
  assert(total_chunk > 0);
  assert(mychunk > 0 && mychunk <= total_chunk);
  if (!ovlp_file_path) { snprintf(default_out, sizeof default_out, "ovlp.%02d", mychunk); ovlp_file_path = default_out; }
  char path[8400];
  snprintf(path, sizeof path, "%s.idx", seqdb_prefix);
  fprintf(stderr, "using index file: %s\n", path);
  snprintf(path, sizeof path, "%s.seqdb", seqdb_prefix);
  fprintf(stderr, "using seqdb file: %s\n", path);
  std::vector<mm128> mmers;
  for (auto &fn : glob_sorted(std::string(shimmer_prefix) + "-[0-9]*-of-[0-9]*.dat")) {
    fprintf(stderr, "using shimmer data file: %s\n", fn.c_str());
    read_mmlist_file(fn.c_str(), &mmers);
  }
  std::vector<mc_rec> mc;
  for (auto &fn : glob_sorted(std::string(shimmer_prefix) + "-MC-[0-9]*-of-[0-9]*.dat")) {
    fprintf(stderr, "using shimmer count file: %s\n", fn.c_str());
    read_mc_file(fn.c_str(), &mc);
  }
  FILE *out = fopen(ovlp_file_path, "w");
  if (!out) { fprintf(stderr, "file '%s' open error: %s\n", ovlp_file_path, strerror(errno)); exit(1); }
  pgb_ctx *c = cli_ctx();
  CLI_CHECK(c, pgb_load_reads_from_files(c, seqdb_prefix, 1, 1, 0));
  CLI_CHECK(c, pgb_set_shimmers(c, (const mm128_t *)mmers.data(), mmers.size(), (const mm_count_t *)mc.data(), mc.size()));
  CLI_CHECK(c, pgb_overlap(c, total_chunk, mychunk, bestn, mc_lower, mc_upper, align_bandwidth, ovlp_upper));
  const ovlp_t *recs_p = nullptr;
  size_t recs_n = 0;
  CLI_CHECK(c, pgb_overlap_host(c, &recs_p, &recs_n));
  if (recs_n) fwrite(recs_p, sizeof(ovlp_t), recs_n, out);
  fclose(out);
  pgb_destroy(c);
  return 0;
}

// shmr_dedup: raw ovlp_t stream on stdin -> preads.ovl text on stdout (src/shmr_dedup.c:19-101; no options)
extern "C" int pgb_shmr_dedup_main(int argc, char **argv) {
  (void)argc; (void)argv;
  // stdin is consumed in batches of PGB_DEDUP_BATCH records (default 8 M = 512 MB); the pair table stays on the device between them,
  // so host memory is bounded by one batch and device memory by the table (16 B per distinct pair at load <= 0.5) plus one batch
  size_t batch = getenv("PGB_DEDUP_BATCH") ? (size_t)strtoull(getenv("PGB_DEDUP_BATCH"), 0, 10) : ((size_t)8 << 20);
  if (batch < 1) batch = 1;
  if (batch > ((size_t)1 << 29)) batch = (size_t)1 << 29;
  std::vector<char> in(batch * sizeof(ovlp_t)), text;
  pgb_ctx *c = nullptr;
  size_t have = 0;  // bytes of an incomplete record carried over
  for (;;) {
    size_t got = have;
    while (got < in.size()) {
      const size_t r = fread(in.data() + got, 1, in.size() - got, stdin);
      if (r == 0) break;
      got += r;
    }
    const size_t n = got / sizeof(ovlp_t);  // a truncated trailing record is ignored
    if (n == 0) break;
    if (!c) {
      c = cli_ctx();
      if (!c) return 1;
      if (pgb_dedup_stream_begin(c) != 0) { fprintf(stderr, "shmr_dedup: %s\n", pgb_last_error(c)); pgb_destroy(c); return 1; }
    }
    if (pgb_dedup_stream_push(c, (const ovlp_t *)in.data(), n) != 0) {
      fprintf(stderr, "shmr_dedup: %s\n", pgb_last_error(c));
      pgb_destroy(c);
      return 1;
    }
    text.resize(pgb_dedup_text_bytes(c));
    if (!text.empty() && pgb_dedup_text_copy(c, text.data()) != 0) {
      fprintf(stderr, "shmr_dedup: %s\n", pgb_last_error(c));
      pgb_destroy(c);
      return 1;
    }
    fwrite(text.data(), 1, text.size(), stdout);
    have = got - n * sizeof(ovlp_t);
    if (have) memmove(in.data(), in.data() + n * sizeof(ovlp_t), have);
    if (got < in.size()) break;  // end of the stream
  }
  fflush(stdout);
  if (c) { pgb_dedup_stream_end(c); pgb_destroy(c); }
  return 0;
}

// shmr_mkseqdb -d seq_dataset.lst -p seq_dataset_prefix (src/shmr_mkseqdb.c:15-132): same options, defaults, messages and
// output files (<p>.idx text "%09d %s %u %lu", <p>.seqdb bytes); the FASTA/FASTQ(.gz) grammar is fasta_reader.hpp's
// restatement of kseq, the base encoding runs on the GPU in batches through page-locked staging buffers.
extern "C" int pgb_shmr_mkseqdb_main(int argc, char **argv) {
  const char *seq_dataset_path = nullptr, *seqdb_prefix = nullptr;
  int ch;
  opterr = 0;
  optind = 1;
  while ((ch = getopt(argc, argv, "d:p:")) != -1) {
    switch (ch) {
      case 'd': seq_dataset_path = optarg; break;
      case 'p': seqdb_prefix = optarg; break;
      case '?':
        if (optopt == 'd') fprintf(stderr, "Option -%c not specified, using 'seq_dataset.lst' as the input file\n", optopt);
        else if (optopt == 'p') fprintf(stderr, "Option -%c not specified, using 'seq_dataset' as the output prefix\n", optopt);
        else fprintf(stderr, "Usage: shmr_mkseqdb -d seq_dataset.lst -p seq_dataset_prefix\n");
        return 1;
      default: abort();
    }
  }
  if (!seq_dataset_path) seq_dataset_path = "seq_dataset.lst";
  if (!seqdb_prefix) seqdb_prefix = "seq_dataset";
  FILE *lst = fopen(seq_dataset_path, "r");
  printf("input sequence dataset file list: '%s'\n", seq_dataset_path);
  if (!lst) { fprintf(stderr, "file '%s' open error: %s\n", seq_dataset_path, strerror(errno)); exit(1); }
  std::string index_fn = std::string(seqdb_prefix) + ".idx", seqdb_fn = std::string(seqdb_prefix) + ".seqdb";
  printf("output index file: %s\n", index_fn.c_str());
  FILE *index_file = fopen(index_fn.c_str(), "w");
  if (!index_file) { fprintf(stderr, "file '%s' open error: %s\n", index_fn.c_str(), strerror(errno)); exit(1); }
  printf("output seqdb file: %s\n", index_fn.c_str());  // (sic) the reference prints the index name here, src/shmr_mkseqdb.c:92
  FILE *seqdb_file = fopen(seqdb_fn.c_str(), "wb");
  if (!seqdb_file) { fprintf(stderr, "file '%s' open error: %s\n", seqdb_fn.c_str(), strerror(errno)); exit(1); }
  // the 2-bit image of the same reads for this library's shmr_index / shmr_overlap (not part of the reference's outputs)
  const std::string f2b = std::string(seqdb_prefix) + ".seq2b", f2n = std::string(seqdb_prefix) + ".seq2n";
  FILE *w2b = getenv("PGB_NO_SEQ2B") ? nullptr : fopen(f2b.c_str(), "wb");
  FILE *w2n = w2b ? fopen(f2n.c_str(), "wb") : nullptr;
  std::vector<uint64_t> b_words;
  std::vector<uint32_t> b_nmask;
  std::vector<uint8_t> b_hasn;
  uint32_t batch_first_read = 0;
  pgb_ctx *c = cli_ctx();
  if (!c) return 1;
  // batch of reads staged in page-locked memory
  size_t cap = (size_t)256 << 20;
  char *stage_in = nullptr; uint8_t *stage_out = nullptr;
  if (cudaMallocHost((void **)&stage_in, cap) != cudaSuccess || cudaMallocHost((void **)&stage_out, cap) != cudaSuccess) {
    fprintf(stderr, "shmr_mkseqdb: cannot allocate staging buffers\n");
    return 1;
  }
  std::vector<uint64_t> b_off;
  std::vector<uint32_t> b_len;
  size_t fill = 0;
  auto flush = [&]() -> bool {
    if (b_len.empty()) return true;
    if (fill && pgb_encode_biseq(c, stage_in, fill, b_off.data(), b_len.data(), b_len.size(), stage_out) != 0) {
      fprintf(stderr, "shmr_mkseqdb: %s\n", pgb_last_error(c));
      return false;
    }
    fwrite(stage_out, 1, fill, seqdb_file);
    if (w2b && w2n) {
      uint64_t nw = 0;
      for (uint32_t l : b_len) nw += ((uint64_t)l + 31) / 32;
      b_words.resize(nw); b_nmask.resize(nw); b_hasn.resize(b_len.size());
      if (pgb_pack_2bit(c, stage_out, fill, b_off.data(), b_len.data(), b_len.size(), b_words.data(), b_nmask.data(), b_hasn.data()) != 0) {
        fprintf(stderr, "shmr_mkseqdb: %s\n", pgb_last_error(c));
        return false;
      }
      fwrite(b_words.data(), 8, nw, w2b);
      uint64_t o = 0;
      for (size_t i = 0; i < b_len.size(); i++) {
        const uint32_t n_i = (uint32_t)(((uint64_t)b_len[i] + 31) / 32);
        if (b_hasn[i]) {
          const uint32_t hdr[2] = {batch_first_read + (uint32_t)i, n_i};
          fwrite(hdr, 4, 2, w2n);
          fwrite(b_nmask.data() + o, 4, n_i, w2n);
        }
        o += n_i;
      }
    }
    batch_first_read += (uint32_t)b_len.size();
    b_off.clear(); b_len.clear(); fill = 0;
    return true;
  };
  char fn[8192];
  uint32_t rid = 0;
  size_t offset = 0;
  int rc = 0;
  while (rc == 0 && fscanf(lst, "%8191s", fn) != EOF) {
    GzRecordStream sc(fn);  // block-wise: a 100 GB fastq.gz does not have to fit in memory
    if (!sc.ok()) { fprintf(stderr, "file '%s' open error: %s\n", fn, strerror(errno)); exit(1); }
    FastaRecord r;
    while (sc.next(r)) {
      const size_t l = r.seq.size();
      if (l >= (1ull << 32)) { fprintf(stderr, "shmr_mkseqdb: sequence %s is longer than 2^32\n", r.name.c_str()); rc = 1; break; }
      if (fill + l > cap) {
        if (!flush()) { rc = 1; break; }
        if (l > cap) {  // one sequence larger than the staging buffers (a chromosome-sized contig)
          cudaFreeHost(stage_in); cudaFreeHost(stage_out);
          cap = l + (l >> 3);
          if (cudaMallocHost((void **)&stage_in, cap) != cudaSuccess || cudaMallocHost((void **)&stage_out, cap) != cudaSuccess) {
            fprintf(stderr, "shmr_mkseqdb: cannot allocate staging buffers\n");
            return 1;
          }
        }
      }
      memcpy(stage_in + fill, r.seq.data(), l);
      b_off.push_back(fill); b_len.push_back((uint32_t)l);
      fill += l;
      fprintf(index_file, "%09d %s %u %lu\n", rid, r.name.c_str(), (unsigned)l, offset);
      rid += 1;
      offset += l;
    }
  }
  if (rc == 0 && !flush()) rc = 1;
  fclose(lst);
  fclose(index_file);
  fclose(seqdb_file);
  if (w2b) fclose(w2b);
  if (w2n) fclose(w2n);
  cudaFreeHost(stage_in); cudaFreeHost(stage_out);
  pgb_destroy(c);
  return rc;
}

// shmr_map -r ref_prefix -m ref_shimmer_prefix -p seqdb_prefix -l shimmer_prefix [-M 240] [-n 1] [-t 1] [-c 1]
// (src/shmr_map.c:168-380): same options, defaults, messages; hits on stdout
extern "C" int pgb_shmr_map_main(int argc, char **argv) {
  const char *refdb_prefix = "ref", *ref_shimmer_prefix = "ref-L2", *seqdb_prefix = "seq_dataset", *shimmer_prefix = "shimmer-L2";
  uint32_t total_chunk = 1, mychunk = 1, mc_upper = 240, mc_lower = 1;
  int ch;
  opterr = 0; optind = 1;
  while ((ch = getopt(argc, argv, "r:m:p:l:M:n:t:c:b:")) != -1) {
    switch (ch) {
      case 'r': refdb_prefix = optarg; break;
      case 'm': ref_shimmer_prefix = optarg; break;
      case 'p': seqdb_prefix = optarg; break;
      case 'l': shimmer_prefix = optarg; break;
      case 'M': mc_upper = atoi(optarg); break;
      case 'n': mc_lower = atoi(optarg); break;
      case 't': total_chunk = atoi(optarg); break;
      case 'c': mychunk = atoi(optarg); break;
      case 'b': abort();  // accepted by the reference's getopt string but has no case: falls to default -> abort()
      case '?':
        if (optopt == 'r') fprintf(stderr, "Option -%c not specified, using 'ref' as the ref sequence db prefix\n", optopt);
        if (optopt == 'p') fprintf(stderr, "Option -%c not specified, using 'seq_dataset' as the sequence db prefix\n", optopt);
        if (optopt == 'l') fprintf(stderr, "Option -%c not specified, using 'shimmer-L2' as the L2 index prefix\n", optopt);
        return 1;
      default: abort();
    }
  }
  assert(total_chunk > 0);
  assert(mychunk > 0 && mychunk <= total_chunk);
  std::string ref_idx = std::string(refdb_prefix) + ".idx", seq_idx = std::string(seqdb_prefix) + ".idx";
  fprintf(stderr, "using ref index file: %s\n", ref_idx.c_str());
  ReadTable ref_rt, rt;
  if (!load_read_table(ref_idx.c_str(), &ref_rt)) { fprintf(stderr, "file '%s' open error: %s\n", ref_idx.c_str(), strerror(errno)); exit(1); }
  fprintf(stderr, "using ref seqdb file: %s.seqdb\n", refdb_prefix);
  std::vector<mm128> ref_mmers, mmers;
  for (auto &fn : glob_sorted(std::string(ref_shimmer_prefix) + "-[0-9]*-of-[0-9]*.dat")) {
    fprintf(stderr, "using ref shimmer data file: %s\n", fn.c_str());
    const size_t before = ref_mmers.size();
    read_mmlist_file(fn.c_str(), &ref_mmers);
    fprintf(stderr, "number of shimmers load: %lu\n", ref_mmers.size() - before);
  }
  fprintf(stderr, "using index file: %s\n", seq_idx.c_str());
  if (!load_read_table(seq_idx.c_str(), &rt)) { fprintf(stderr, "file '%s' open error: %s\n", seq_idx.c_str(), strerror(errno)); exit(1); }
  fprintf(stderr, "using seqdb file: %s.seqdb\n", seqdb_prefix);
  for (auto &fn : glob_sorted(std::string(shimmer_prefix) + "-[0-9]*-of-[0-9]*.dat")) {
    fprintf(stderr, "using shimmer data file: %s\n", fn.c_str());
    const size_t before = mmers.size();
    read_mmlist_file(fn.c_str(), &mmers);
    fprintf(stderr, "number of shimmers load: %lu\n", mmers.size() - before);
  }
  std::vector<mc_rec> mc;
  for (auto &fn : glob_sorted(std::string(shimmer_prefix) + "-MC-[0-9]*-of-[0-9]*.dat")) {
    fprintf(stderr, "using shimmer count file: %s\n", fn.c_str());
    read_mc_file(fn.c_str(), &mc);
  }
  assert(ref_mmers.size() > 0);  // src/shmr_map.c:84
  pgb_ctx *c = cli_ctx();
  CLI_CHECK(c, pgb_set_read_lengths(c, rt.rid.data(), rt.len.data(), rt.n()));
  CLI_CHECK(c, pgb_set_shimmers(c, (const mm128_t *)mmers.data(), mmers.size(), (const mm_count_t *)mc.data(), mc.size()));
  CLI_CHECK(c, pgb_map(c, (const mm128_t *)ref_mmers.data(), ref_mmers.size(), total_chunk, mychunk, mc_lower, mc_upper));
  std::vector<char> text(pgb_map_text_bytes(c));
  CLI_CHECK(c, pgb_map_text_copy(c, text.data()));
  fwrite(text.data(), 1, text.size(), stdout);
  fflush(stdout);
  pgb_destroy(c);
  return 0;
}

// ================================================================================================ reference cffi surface
extern "C" void decode_biseq(uint8_t *src, char *seq, size_t len, uint8_t strand) {
  // byte-format conversion helper for Python callers (src/shmr_utils.c:53-62); not part of the compute path
  static const char b2b[16] = {'N', 'A', 'C', 'N', 'G', 'N', 'N', 'N', 'T', 'N', 'N', 'N', 'N', 'N', 'N', 'N'};
  for (size_t p = 0; p < len; p++) seq[p] = strand == 0 ? b2b[src[p] & 0x0F] : b2b[src[p] >> 4];
}
extern "C" void free_ovlp_match(ovlp_match_t *m) { free(m); }
extern "C" mm128_v read_mmlist(char *fn) {
  mm128_v p = {0, 0, 0};
  std::vector<mm128> v;
  read_mmlist_file(fn, &v);
  p.n = p.m = v.size();
  p.a = (mm128_t *)malloc((v.size() ? v.size() : 1) * sizeof(mm128_t));
  memcpy((void *)p.a, v.data(), v.size() * sizeof(mm128_t));
  return p;
}

static pgb_ctx *g_ctx = nullptr;  // lazily created context behind the single-call cffi functions
// The reference's functions keep no state and may be called from several threads; ours share one device context, so every
// single-call entry point takes this lock for its duration (SharedCtx).
static std::recursive_mutex g_ctx_mu;
struct SharedCtx {
  std::lock_guard<std::recursive_mutex> guard;
  pgb_ctx *c;
  SharedCtx() : guard(g_ctx_mu) {
    if (!g_ctx) g_ctx = cli_ctx();
    c = g_ctx;
  }
};
static void mm128v_append(mm128_v *p, const mm128_t *src, size_t n) {
  if (p->n + n > p->m) {
    size_t m = p->m ? p->m : 16;
    while (m < p->n + n) m <<= 1;
    p->a = (mm128_t *)realloc(p->a, m * sizeof(mm128_t));
    p->m = m;
  }
  memcpy((void *)(p->a + p->n), src, n * sizeof(mm128_t));
  p->n += n;
}
static inline uint8_t ascii_to_nibble(char ch) {
  switch (ch) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 4;
    case 'T': case 't': case 'U': case 'u': return 8;  // seq_nt4_table maps U/u to 3 as well (src/mm_sketch.c:10-21)
    default: return 0;
  }
}
extern "C" void mm_sketch(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p) {
  (void)km;
  assert(len > 0 && (w > 0 && w < 256) && (k > 0 && k <= 28));  // src/mm_sketch.c:77-78
  if (is_hpc) { fprintf(stderr, "pgb200: mm_sketch with is_hpc != 0 is not supported (no reference caller uses it)\n"); exit(1); }
  SharedCtx shared_;
  pgb_ctx *c = shared_.c;
  std::vector<uint8_t> img((size_t)len);
  for (int i = 0; i < len; i++) img[i] = ascii_to_nibble(str[i]);
  uint32_t one_rid = 0, one_len = (uint32_t)len; uint64_t one_off = 0;
  CLI_CHECK(c, pgb_load_reads(c, img.data(), img.size(), &one_rid, &one_len, &one_off, 1, 1, 1, 0));
  CLI_CHECK(c, pgb_index(c, w, k, 1, 0, 0));
  std::vector<mm128_t> v(pgb_index_size(c, 0));
  CLI_CHECK(c, pgb_index_copy(c, 0, v.data()));
  for (auto &m : v) m.y = (m.y & 0xFFFFFFFFULL) | ((uint64_t)rid << 32);
  mm128v_append(p, v.data(), v.size());
}

extern "C" void mm_reduce(mm128_v *in, mm128_v *out, uint8_t rs) {
  // runs of equal rid are the "reads" of src/shmr_reduce.c:72-77 (the ring is reset whenever the rid changes)
  if (!in || in->n == 0) return;
  SharedCtx shared_;
  pgb_ctx *c = shared_.c;
  try {
    CU(cudaSetDevice(c->device));
    size_t n = in->n;
    std::vector<uint64_t> off;
    for (size_t i = 0; i < n; i++)
      if (i == 0 || (in->a[i].y >> 32) != (in->a[i - 1].y >> 32)) off.push_back(i);
    size_t runs = off.size();
    off.push_back(n);
    mm128 *d_in = c->alloc<mm128>(n);
    uint64_t *d_off = c->alloc<uint64_t>(runs + 1), *d_ooff = c->alloc<uint64_t>(runs + 1);
    uint32_t *counts = c->alloc<uint32_t>(runs + 1);
    c->h2d(d_in, in->a, n * sizeof(mm128));
    c->h2d(d_off, off.data(), (runs + 1) * 8);
    CU(cudaMemsetAsync(counts, 0, (runs + 1) * 4, c->st));
    LAUNCH(c, k_reduce<false>, nblk(runs, 128), 128, d_in, d_off, (uint32_t)runs, (uint32_t)rs, counts, (const uint64_t *)nullptr, (mm128 *)nullptr);
    uint64_t total = scan_u32_to_u64(c, counts, d_ooff, runs + 1);
    mm128 *d_out = c->alloc<mm128>(total);
    LAUNCH(c, k_reduce<true>, nblk(runs, 128), 128, d_in, d_off, (uint32_t)runs, (uint32_t)rs, (uint32_t *)nullptr, d_ooff, d_out);
    std::vector<mm128_t> v(total);
    c->d2h(v.data(), d_out, total * sizeof(mm128));
    c->release(d_in); c->release(d_off); c->release(d_ooff); c->release(counts); c->release(d_out);
    c->sync();
    c->scratch_reset();
    mm128v_append(out, v.data(), v.size());
  } catch (std::exception &e) {
    fprintf(stderr, "pgb200: mm_reduce failed: %s\n", e.what());
    exit(1);
  }
}

// ovlp_match for a batch of operand pairs given as .seqdb bytes (the cffi callers: py/scripts/path_to_contig.py:82-105,
// py/peregrine/utils.py).  Nothing of the context's loaded read set is touched: the operands are packed into scratch memory
// from the nibble their strand selects, one request per pair runs through the warp-per-alignment kernel (pairs with N
// through the generic one), and the results come back in one copy.
static void match_batch_core(pgb_ctx *c, const uint8_t *seq, size_t seq_bytes, size_t n, const uint64_t *q_off, const uint32_t *q_len, const uint8_t *q_strand,
                             const uint64_t *t_off, const uint32_t *t_len, const uint8_t *t_strand, int band_tolerance, ovlp_match_t *out) {
  static_assert(sizeof(AlnReqPOD) == sizeof(AlnReq), "request layout");
  if (band_tolerance + 3 > PGB_MAXV || band_tolerance < 0) throw std::runtime_error("ovlp_match: band_tolerance outside [0, 261]");
  if (n == 0) return;
  if (n >= (1ull << 30)) throw std::runtime_error("too many pairs in one batch");
  const size_t rows = 2 * n;
  // one host block: [raw offsets u64 x rows][word offsets u64 x rows][lengths u32 x rows][shifts u8 x rows]
  std::vector<uint64_t> meta(2 * rows + (rows + 1) / 2 + (rows + 7) / 8 + 1);
  uint64_t *h_raw = meta.data(), *h_woff = meta.data() + rows;
  uint32_t *h_len = reinterpret_cast<uint32_t *>(meta.data() + 2 * rows);
  uint8_t *h_shift = reinterpret_cast<uint8_t *>(meta.data() + 2 * rows + (rows + 1) / 2);
  uint64_t words = 2, lo = ~0ULL, hi = 0;
  for (size_t i = 0; i < n; i++) {
    for (int s_ = 0; s_ < 2; s_++) {
      const uint64_t off = s_ ? t_off[i] : q_off[i];
      const uint32_t len = s_ ? t_len[i] : q_len[i];
      if (off + len > seq_bytes) throw std::runtime_error("ovlp_match: operand extends past the end of the buffer");
      const size_t r = 2 * i + s_;
      h_raw[r] = off; h_len[r] = len; h_woff[r] = words; h_shift[r] = (s_ ? t_strand[i] : q_strand[i]) ? 4 : 0;
      words += ((uint64_t)len + 31) / 32;
      if (len) { lo = std::min(lo, off); hi = std::max(hi, off + len); }
    }
  }
  words += 2;
  if (hi <= lo) { lo = 0; hi = 0; }
  for (size_t r = 0; r < rows; r++) h_raw[r] -= std::min(h_raw[r], lo);
  uint8_t *d_raw = c->alloc<uint8_t>(hi - lo + 64);
  uint64_t *d_meta = c->alloc<uint64_t>(meta.size());
  uint64_t *d_w = c->alloc<uint64_t>(words);
  uint32_t *d_nm = c->alloc<uint32_t>(words), *d_hasn = c->alloc<uint32_t>(rows);
  AlnReq *d_req = c->alloc<AlnReq>(n);
  match_t *d_res = c->alloc<match_t>(n);
  c->h2d(d_raw, seq + lo, hi - lo);
  c->h2d(d_meta, meta.data(), meta.size() * 8);
  const uint64_t *d_rawoff = d_meta, *d_woff = d_meta + rows;
  const uint32_t *d_len = reinterpret_cast<const uint32_t *>(d_meta + 2 * rows);
  const uint8_t *d_shift = reinterpret_cast<const uint8_t *>(d_meta + 2 * rows + (rows + 1) / 2);
  CU(cudaMemsetAsync(d_hasn, 0, rows * 4, c->st));
  LAUNCH(c, k_pack_nibbles, nblk(words), 256, d_raw, d_rawoff, d_len, d_woff, d_shift, (uint32_t)rows, words, d_w, d_nm, d_hasn);
  LAUNCH(c, k_pair_requests, nblk(n), 256, (uint32_t)n, reinterpret_cast<AlnReqPOD *>(d_req));
  // (rid = row index: the by-rid tables of the kernels are the row tables)
  LAUNCH(c, k_align_warp, nblk(n, PGB_AW_WARPS), PGB_AW_WARPS * 32, d_req, 0u, (uint32_t)n, d_w, d_w, d_woff, d_len, d_hasn, band_tolerance, d_res,
         c->d_align_bases);
  LAUNCH(c, k_align, nblk(n, 64), 64, d_req, 0u, (uint32_t)n, (const uint32_t *)nullptr, d_w, d_nm, d_woff, d_len, d_hasn, band_tolerance, d_res, c->d_err,
         c->d_align_bases, 1);
  c->d2h(out, d_res, n * sizeof(match_t));
  CU(cudaMemsetAsync(c->d_align_bases, 0, 8, c->st));
  if (c->check_err("ovlp_match")) throw std::runtime_error(c->err);
}
extern "C" int pgb_ovlp_match_batch(pgb_ctx *c, const uint8_t *seq, size_t seq_bytes, size_t n, const uint64_t *q_off, const uint32_t *q_len,
                                    const uint8_t *q_strand, const uint64_t *t_off, const uint32_t *t_len, const uint8_t *t_strand, int band_tolerance,
                                    ovlp_match_t *out) {
  API_BEGIN(c)
  match_batch_core(c, seq, seq_bytes, n, q_off, q_len, q_strand, t_off, t_len, t_strand, band_tolerance, out);
  API_END(c)
}

extern "C" ovlp_match_t *ovlp_match(uint8_t *query_seq, seq_coor_t q_len, uint8_t q_strand, uint8_t *target_seq, seq_coor_t t_len,
                                    uint8_t t_strand, seq_coor_t band_tolerance) {
  // a batch of one through the same path (src/DWmatch.c:66-204); the result is calloc'd like the reference's (:105)
  ovlp_match_t *rtn = (ovlp_match_t *)calloc(1, sizeof(ovlp_match_t));
  if (q_len <= 0 || t_len <= 0) return rtn;
  SharedCtx shared_;
  pgb_ctx *c = shared_.c;
  // the two operands may live anywhere: stage them back to back
  std::vector<uint8_t> img((size_t)q_len + (size_t)t_len);
  memcpy(img.data(), query_seq, (size_t)q_len);
  memcpy(img.data() + q_len, target_seq, (size_t)t_len);
  const uint64_t qo = 0, to = (uint64_t)q_len;
  const uint32_t ql = (uint32_t)q_len, tl = (uint32_t)t_len;
  if (pgb_ovlp_match_batch(c, img.data(), img.size(), 1, &qo, &ql, &q_strand, &to, &tl, &t_strand, band_tolerance, rtn) != 0) {
    fprintf(stderr, "pgb200: ovlp_match failed: %s\n", pgb_last_error(c));
    exit(1);
  }
  return rtn;
}

static void idxv_push(mm_idx_v *v, mm_idx_t x) {
  if (v->n == v->m) {
    v->m = v->m ? v->m << 1 : 2;  // kvec growth (kvec.h:78-84)
    v->a = (mm_idx_t *)realloc(v->a, sizeof(mm_idx_t) * v->m);
  }
  v->a[v->n++] = x;
}
// n_pairs chaining problems in one go (kernels.cuh "shmr_aln"): the hits of pair p are out[hoff[p] .. hoff[p] + n_hits[p]) in the
// order the reference produces them; chains are numbered in order of creation
static void aln_batch_core(pgb_ctx *c, const mm128 *mm0, const uint64_t *off0, const mm128 *mm1, const uint64_t *off1, uint32_t n_pairs, uint32_t direction,
                           uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat, std::vector<uint64_t> &hoff, std::vector<uint32_t> &n_hits,
                           std::vector<uint32_t> &n_chains, std::vector<AlnHit> &hits) {
  hoff.assign((size_t)n_pairs + 1, 0); n_hits.assign(n_pairs, 0); n_chains.assign(n_pairs, 0); hits.clear();
  if (!n_pairs) return;
  if (n_pairs >= (1u << (64 - PGB_ALN_HBITS))) throw std::runtime_error("pgb_shmr_aln_batch: more than 2^24 pairs in one call");
  const uint64_t n0 = off0[n_pairs], n1 = off1[n_pairs];
  if (n0 >= (1ull << 31) || n1 >= (1ull << 31)) throw std::runtime_error("pgb_shmr_aln_batch: more than 2^31 minimizers in one call");
  if (!n0 || !n1) return;
  mm128 *d0 = c->alloc<mm128>(n0), *d1 = c->alloc<mm128>(n1);
  uint64_t *doff0 = c->alloc<uint64_t>((size_t)n_pairs + 1), *doff1 = c->alloc<uint64_t>((size_t)n_pairs + 1);
  c->h2d(d0, mm0, n0 * sizeof(mm128)); c->h2d(d1, mm1, n1 * sizeof(mm128));
  c->h2d(doff0, off0, ((size_t)n_pairs + 1) * 8); c->h2d(doff1, off1, ((size_t)n_pairs + 1) * 8);
  uint32_t *pid0 = c->alloc<uint32_t>(n0), *pid1 = c->alloc<uint32_t>(n1);
  LAUNCH(c, k_aln_pair_ids, nblk(n0), 256, doff0, n_pairs, n0, pid0);
  LAUNCH(c, k_aln_pair_ids, nblk(n1), 256, doff1, n_pairs, n1, pid1);
  uint64_t *keys = c->alloc<uint64_t>(n0), *skeys = c->alloc<uint64_t>(n0);
  uint32_t *idx = c->alloc<uint32_t>(n0), *sidx = c->alloc<uint32_t>(n0);
  LAUNCH(c, k_aln_keys0, nblk(n0), 256, d0, pid0, n0, keys, idx);
  size_t tmp_bytes = 0;
  CU(cub::DeviceRadixSort::SortPairs((void *)nullptr, tmp_bytes, keys, skeys, idx, sidx, (int)n0, 0, 64, c->st));
  uint8_t *tmp = c->alloc<uint8_t>(tmp_bytes);
  CU(cub::DeviceRadixSort::SortPairs((void *)tmp, tmp_bytes, keys, skeys, idx, sidx, (int)n0, 0, 64, c->st));
  c->stats.kernel_launches += 4;
  uint32_t *cnt = c->alloc<uint32_t>(n1 + 1);
  uint64_t *moff = c->alloc<uint64_t>(n1 + 1);
  CU(cudaMemsetAsync(cnt, 0, (n1 + 1) * 4, c->st));
  LAUNCH(c, k_aln_lookup<false>, nblk(n1, 128), 128, d0, d1, pid1, n1, doff0, skeys, sidx, cnt, (const uint64_t *)nullptr, (uint32_t *)nullptr);
  const uint64_t n_match = scan_u32_to_u64(c, cnt, moff, n1 + 1);
  uint32_t *midx = c->alloc<uint32_t>(n_match + 1);
  LAUNCH(c, k_aln_lookup<true>, nblk(n1, 128), 128, d0, d1, pid1, n1, doff0, skeys, sidx, (uint32_t *)nullptr, moff, midx);
  uint32_t *ub = c->alloc<uint32_t>((size_t)n_pairs + 1);
  uint64_t *dhoff = c->alloc<uint64_t>((size_t)n_pairs + 1);
  CU(cudaMemsetAsync(ub, 0, ((size_t)n_pairs + 1) * 4, c->st));
  LAUNCH(c, k_aln_hit_bound, nblk((size_t)n_pairs * 32, 128), 128, cnt, doff1, n_pairs, max_repeat, ub);
  const uint64_t n_bound = scan_u32_to_u64(c, ub, dhoff, (size_t)n_pairs + 1);
  AlnChain *chains = c->alloc<AlnChain>(n_bound + 1);
  AlnHit *dhits = c->alloc<AlnHit>(n_bound + 1);
  uint32_t *dnh = c->alloc<uint32_t>(n_pairs), *dnc = c->alloc<uint32_t>(n_pairs);
  LAUNCH(c, k_aln_chain_warp, nblk((size_t)n_pairs * 32, 128), 128, d0, d1, doff0, doff1, n_pairs, cnt, moff, midx, direction, max_diff, max_dist, max_repeat,
         dhoff, chains, dhits, dnh, dnc);
  hits.resize(n_bound);
  c->d2h(hoff.data(), dhoff, ((size_t)n_pairs + 1) * 8);
  c->d2h(n_hits.data(), dnh, (size_t)n_pairs * 4);
  c->d2h(n_chains.data(), dnc, (size_t)n_pairs * 4);
  if (n_bound) c->d2h(hits.data(), dhits, n_bound * sizeof(AlnHit));
  c->sync();
}

extern "C" int pgb_shmr_aln_batch(pgb_ctx *c, const mm128_t *mm0, const uint64_t *off0, const mm128_t *mm1, const uint64_t *off1, uint32_t n_pairs,
                                  uint8_t direction, uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat, uint64_t *hit_off, uint32_t *n_chains,
                                  pgb_aln_hit_t **hits_out) {
  API_BEGIN(c)
  static_assert(sizeof(pgb_aln_hit_t) == sizeof(AlnHit), "hit layout");
  if (hits_out) *hits_out = nullptr;
  std::vector<uint64_t> hoff;
  std::vector<uint32_t> nh, nc;
  std::vector<AlnHit> hits;
  aln_batch_core(c, (const mm128 *)mm0, off0, (const mm128 *)mm1, off1, n_pairs, direction, max_diff, max_dist, max_repeat, hoff, nh, nc, hits);
  uint64_t total = 0;
  for (uint32_t p = 0; p < n_pairs; p++) { hit_off[p] = total; total += nh[p]; if (n_chains) n_chains[p] = nc[p]; }
  hit_off[n_pairs] = total;
  pgb_aln_hit_t *o = (pgb_aln_hit_t *)malloc((total ? total : 1) * sizeof(pgb_aln_hit_t));
  if (!o) throw std::runtime_error("pgb_shmr_aln_batch: out of host memory");
  for (uint32_t p = 0; p < n_pairs; p++)
    if (nh[p]) memcpy(o + hit_off[p], hits.data() + hoff[p], (size_t)nh[p] * sizeof(AlnHit));
  *hits_out = o;
  API_END(c)
}
extern "C" void pgb_host_free(void *p) { free(p); }

extern "C" shmr_aln_v *shmr_aln(mm128_v *mmers0, mm128_v *mmers1, uint8_t direction, uint32_t max_diff, uint32_t max_dist, uint32_t max_repeat) {
  // a batch of one (src/shmr_align.c:21-160); the result is built with kvec's growth rule like the reference's
  shmr_aln_v *alns = (shmr_aln_v *)calloc(sizeof(shmr_aln_v), 1);
  const size_t n0 = mmers0 ? mmers0->n : 0, n1 = mmers1 ? mmers1->n : 0;
  if (!n0 || !n1) return alns;
  SharedCtx shared_;
  pgb_ctx *c = shared_.c;
  try {
    CU(cudaSetDevice(c->device));
    const uint64_t off0[2] = {0, n0}, off1[2] = {0, n1};
    std::vector<uint64_t> hoff;
    std::vector<uint32_t> nh, nc;
    std::vector<AlnHit> hits;
    aln_batch_core(c, (const mm128 *)mmers0->a, off0, (const mm128 *)mmers1->a, off1, 1, direction, max_diff, max_dist, max_repeat, hoff, nh, nc, hits);
    c->scratch_reset();
    alns->n = alns->m = nc[0];
    alns->a = (shmr_aln_t *)calloc(nc[0] ? nc[0] : 1, sizeof(shmr_aln_t));
    for (uint32_t h = 0; h < nh[0]; h++) {
      idxv_push(&alns->a[hits[h].chain].idx0, hits[h].i0);
      idxv_push(&alns->a[hits[h].chain].idx1, hits[h].i1);
    }
  } catch (std::exception &e) {
    fprintf(stderr, "pgb200: shmr_aln failed: %s\n", e.what());
    exit(1);
  }
  return alns;
}
extern "C" void free_shmr_alns(shmr_aln_v *alns) {
  if (!alns) return;
  for (size_t i = 0; i < alns->n; i++) {
    free(alns->a[i].idx0.a);
    free(alns->a[i].idx1.a);
  }
  free(alns->a);
  free(alns);
}

// ================================================================================================ shimmer4py index handle
struct PyMmerHandle {
  pgb_ctx *c = nullptr;
  mm128_v mmers = {0, 0, nullptr};
  std::vector<uint32_t> rid_first, rid_count;
  // device-resident pair index (persistent allocations of c)
  uint64_t *xkeys = nullptr; uint32_t xmask = 0; uint32_t *xfirst = nullptr, *xid = nullptr;
  uint32_t n_buckets = 0, n_outer = 0;
  uint32_t *goff = nullptr, *ipos = nullptr, *boff = nullptr;  // group offsets, in-group visiting position, record offset per grouped bucket
  uint64_t *gk1 = nullptr;
  uint64_t *sy0 = nullptr, *sy1 = nullptr; uint8_t *sdir = nullptr;
};

extern "C" void build_shimmer_map4py(py_mmer_t *py, char *seqdb_prefix, char *shimmer_prefix, uint32_t mychunk, uint32_t total_chunk,
                                     uint32_t lower, uint32_t upper) {
  std::lock_guard<std::recursive_mutex> lock_(g_ctx_mu);
  assert(total_chunk > 0);
  assert(mychunk > 0 && mychunk <= total_chunk);
  const char *sp = seqdb_prefix ? seqdb_prefix : "seq_dataset", *lp = shimmer_prefix ? shimmer_prefix : "shimmer-L2";
  PyMmerHandle *h = new PyMmerHandle();
  h->c = cli_ctx();
  pgb_ctx *c = h->c;
  ReadTable rt;
  std::string idx = std::string(sp) + ".idx";
  fprintf(stderr, "using index file: %s\n", idx.c_str());
  if (!load_read_table(idx.c_str(), &rt)) { fprintf(stderr, "file '%s' open error\n", idx.c_str()); exit(1); }
  fprintf(stderr, "using seqdb file: %s.seqdb\n", sp);
  std::vector<mm128> mm;
  for (auto &fn : glob_sorted(std::string(lp) + "-[0-9]*-of-[0-9]*.dat")) { fprintf(stderr, "using shimmer data file: %s\n", fn.c_str()); read_mmlist_file(fn.c_str(), &mm); }
  std::vector<mc_rec> mc;
  for (auto &fn : glob_sorted(std::string(lp) + "-MC-[0-9]*-of-[0-9]*.dat")) { fprintf(stderr, "using shimmer count file: %s\n", fn.c_str()); read_mc_file(fn.c_str(), &mc); }
  h->mmers.n = h->mmers.m = mm.size();
  h->mmers.a = (mm128_t *)malloc((mm.size() ? mm.size() : 1) * sizeof(mm128_t));
  memcpy((void *)h->mmers.a, mm.data(), mm.size() * sizeof(mm128_t));
  try {
    CU(cudaSetDevice(c->device));
    // read lengths by rid (the mirrored coordinates of reverse records need them, src/shmr_utils.c:376-395); no sequence is loaded
    uint32_t max_rid = 0;
    for (uint32_t r : rt.rid) max_rid = std::max(max_rid, r);
    for (auto &m : mm) max_rid = std::max(max_rid, (uint32_t)(m.y >> 32));
    std::vector<uint32_t> rl((size_t)max_rid + 1, 0);
    for (size_t i = 0; i < rt.n(); i++) rl[rt.rid[i]] = rt.len[i];
    c->free_reads();
    c->max_rid = max_rid;
    c->d_rlen_by_rid = c->palloc<uint32_t>((size_t)max_rid + 1);
    c->h2d(c->d_rlen_by_rid, rl.data(), rl.size() * 4);
    CLI_CHECK(c, pgb_set_shimmers(c, (const mm128_t *)mm.data(), mm.size(), (const mm_count_t *)mc.data(), mc.size()));
    const size_t n = c->n_shm;
    // ridmm
    h->rid_first.assign((size_t)max_rid + 1, 0xFFFFFFFFu);
    h->rid_count.assign((size_t)max_rid + 1, 0);
    if (n) {
      uint32_t *d_first = c->alloc<uint32_t>((size_t)max_rid + 1), *d_count = c->alloc<uint32_t>((size_t)max_rid + 1);
      CU(cudaMemsetAsync(d_first, 0xFF, ((size_t)max_rid + 1) * 4, c->st));
      CU(cudaMemsetAsync(d_count, 0, ((size_t)max_rid + 1) * 4, c->st));
      LAUNCH(c, k_ridmm, nblk(n), 256, c->d_shm, n, d_first, d_count);
      c->d2h(h->rid_first.data(), d_first, ((size_t)max_rid + 1) * 4);
      c->d2h(h->rid_count.data(), d_count, ((size_t)max_rid + 1) * 4);
    }
    // build_map -> records -> buckets -> groups -> inner order (same kernels as pgb_overlap, every bucket kept)
    uint32_t nrec = 0;
    PairSoA R = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (n) {
      uint32_t *cnt = c->alloc<uint32_t>(n), *flags = c->alloc<uint32_t>(n + 1), *pos = c->alloc<uint32_t>(n + 1);
      unsigned long long *d_first = c->alloc<unsigned long long>(1);
      CU(cudaMemsetAsync(d_first, 0xFF, 8, c->st));
      CU(cudaMemsetAsync(flags, 0, (n + 1) * 4, c->st));
      LAUNCH(c, k_count_lookup, nblk(n), 256, c->d_shm, n, c->d_mckeys, c->d_mcvals, c->mcmask, cnt, lower, upper, d_first, c->d_err);
      LAUNCH(c, k_kept_flags, nblk(n), 256, cnt, n, lower, upper, d_first, flags);
      uint32_t n_kept = scan_u32(c, flags, pos, n + 1);
      uint32_t *kept = c->alloc<uint32_t>(n_kept);
      LAUNCH(c, k_compact_idx, nblk(n), 256, flags, pos, n, kept);
      uint32_t *n_rec = c->alloc<uint32_t>((size_t)n_kept + 1), *rec_off = c->alloc<uint32_t>((size_t)n_kept + 1);
      CU(cudaMemsetAsync(n_rec, 0, ((size_t)n_kept + 1) * 4, c->st));
      LAUNCH(c, k_pair_count, nblk(n_kept), 256, c->d_shm, kept, n_kept, total_chunk, mychunk, n_rec);
      nrec = scan_u32(c, n_rec, rec_off, (size_t)n_kept + 1);
      R.k0 = c->alloc<uint64_t>(nrec); R.k1 = c->alloc<uint64_t>(nrec); R.y0 = c->alloc<uint64_t>(nrec); R.y1 = c->alloc<uint64_t>(nrec);
      R.seq = c->alloc<uint32_t>(nrec); R.dir = c->alloc<uint8_t>(nrec);
      LAUNCH(c, k_pair_write, nblk(n_kept), 256, c->d_shm, kept, n_kept, total_chunk, mychunk, rec_off, c->d_rlen_by_rid, R);
    }
    if (nrec) {
      uint32_t xcap = pow2_at_least(4 * (uint64_t)nrec), bcap = pow2_at_least(2 * (uint64_t)nrec);
      h->xkeys = c->palloc<uint64_t>(xcap); h->xmask = xcap - 1; h->xfirst = c->palloc<uint32_t>(xcap); h->xid = c->palloc<uint32_t>(xcap);
      uint64_t *bkeys = c->alloc<uint64_t>(bcap);
      uint32_t *bcount = c->alloc<uint32_t>(bcap), *bfirst = c->alloc<uint32_t>(bcap), *blast = c->alloc<uint32_t>(bcap), *rec_bucket = c->alloc<uint32_t>(nrec);
      LAUNCH(c, k_fill_u64, 1184, 256, h->xkeys, PGB_EMPTY, (size_t)xcap);
      LAUNCH(c, k_fill_u64, 1184, 256, bkeys, PGB_EMPTY, (size_t)bcap);
      CU(cudaMemsetAsync(bcount, 0, (size_t)bcap * 4, c->st));
      CU(cudaMemsetAsync(bfirst, 0xFF, (size_t)bcap * 4, c->st));
      CU(cudaMemsetAsync(blast, 0, (size_t)bcap * 4, c->st));
      CU(cudaMemsetAsync(h->xfirst, 0xFF, (size_t)xcap * 4, c->st));
      LAUNCH(c, k_bucket_insert, nblk(nrec), 256, R, nrec, h->xkeys, xcap - 1, bkeys, bcap - 1, bcount, bfirst, blast, h->xfirst, rec_bucket, c->d_err);
      uint32_t *bflags = c->alloc<uint32_t>((size_t)bcap + 1), *bpos = c->alloc<uint32_t>((size_t)bcap + 1);
      CU(cudaMemsetAsync(bflags, 0, ((size_t)bcap + 1) * 4, c->st));
      LAUNCH(c, k_mc_flags, nblk(bcap), 256, bkeys, (size_t)bcap, bflags);
      uint32_t nb = scan_u32(c, bflags, bpos, (size_t)bcap + 1);
      h->n_buckets = nb;
      uint32_t *lslot = c->alloc<uint32_t>(nb), *lkey = c->alloc<uint32_t>(nb), *sslot = c->alloc<uint32_t>(nb), *skey = c->alloc<uint32_t>(nb);
      LAUNCH(c, k_bucket_list, nblk(bcap), 256, bkeys, bfirst, bpos, (size_t)bcap, lslot, lkey);
      sort_pairs_u32(c, lkey, skey, lslot, sslot, nb);
      uint32_t *isfirst = c->alloc<uint32_t>((size_t)nb + 1), *firstpos = c->alloc<uint32_t>((size_t)nb + 1);
      CU(cudaMemsetAsync(isfirst, 0, ((size_t)nb + 1) * 4, c->st));
      LAUNCH(c, k_outer_first, nblk(nb), 256, sslot, nb, bkeys, bfirst, h->xfirst, isfirst);
      h->n_outer = scan_u32(c, isfirst, firstpos, (size_t)nb + 1);
      LAUNCH(c, k_outer_id_first, nblk(nb), 256, sslot, nb, bkeys, isfirst, firstpos, h->xid);
      uint32_t *oid = lkey, *idx0 = lslot, *goid_sorted = skey, *gidx = c->alloc<uint32_t>(nb);
      LAUNCH(c, k_outer_id_all, nblk(nb), 256, sslot, nb, bkeys, h->xid, oid, idx0);
      sort_pairs_u32(c, oid, goid_sorted, idx0, gidx, nb);
      GroupedBuckets G;
      G.slot = c->alloc<uint32_t>(nb); G.first = c->alloc<uint32_t>(nb); G.last = c->alloc<uint32_t>(nb);
      G.count = c->alloc<uint32_t>((size_t)nb + 1); G.oid = c->alloc<uint32_t>(nb);
      h->gk1 = c->palloc<uint64_t>(nb); G.k1 = h->gk1;
      h->goff = c->palloc<uint32_t>((size_t)h->n_outer + 1); h->ipos = c->palloc<uint32_t>(nb); h->boff = c->palloc<uint32_t>((size_t)nb + 1);
      uint32_t *big_list = c->alloc<uint32_t>((size_t)h->n_outer + 1);
      uint64_t *okey = c->alloc<uint64_t>((size_t)h->n_outer + 1);
      unsigned int *d_small = c->alloc<unsigned int>(2);
      CU(cudaMemsetAsync(d_small, 0, 16, c->st));
      CU(cudaMemsetAsync(G.count, 0, ((size_t)nb + 1) * 4, c->st));
      LAUNCH(c, k_group_gather, nblk(nb), 256, gidx, goid_sorted, nb, sslot, bkeys, h->xkeys, bfirst, blast, bcount, G, h->goff, h->n_outer, okey, d_small);
      LAUNCH(c, k_inner_order, nblk(h->n_outer, 128), 128, h->goff, h->n_outer, G, h->ipos, big_list, d_small + 1);
      unsigned int h_small[2];
      c->d2h(h_small, d_small, 8);
      if (h_small[1]) {  // oversized groups: inner replay on the host
        std::vector<uint32_t> big(h_small[1]), h_goff((size_t)h->n_outer + 1);
        c->d2h(big.data(), big_list, (size_t)h_small[1] * 4);
        c->d2h(h_goff.data(), h->goff, ((size_t)h->n_outer + 1) * 4);
        KhashEmu inner;
        for (uint32_t o : big) {
          uint32_t b = h_goff[o], m = h_goff[o + 1] - b;
          std::vector<uint64_t> k1(m);
          std::vector<uint32_t> fi(m), la(m), ps(m);
          c->d2h(k1.data(), G.k1 + b, (size_t)m * 8); c->d2h(fi.data(), G.first + b, (size_t)m * 4); c->d2h(la.data(), G.last + b, (size_t)m * 4);
          inner.clear();
          uint32_t last = 0;
          for (uint32_t i = 0; i < m; i++) { inner.put_new(k1[i], i); last = std::max(last, la[i]); }
          if (last > fi[m - 1]) inner.touch_existing();
          uint32_t r = 0;
          inner.for_each_in_slot_order([&](uint64_t, uint32_t i) { ps[i] = r++; });
          c->h2d(h->ipos + b, ps.data(), (size_t)m * 4);
          c->sync();
        }
      }
      // every bucket's records, grouped order, sorted like mp128_comp + glibc qsort (stable, descending position)
      uint32_t n_all = scan_u32(c, G.count, h->boff, (size_t)nb + 1);
      uint32_t *slot2j = c->alloc<uint32_t>(bcap), *fill = c->alloc<uint32_t>(nb), *sseq = c->alloc<uint32_t>(n_all);
      LAUNCH(c, k_fill_u32, 1184, 256, slot2j, PGB_NOSLOT, (size_t)bcap);
      LAUNCH(c, k_slot_to_group, nblk(nb), 256, G.slot, nb, slot2j);
      CU(cudaMemsetAsync(fill, 0, (size_t)nb * 4, c->st));
      h->sy0 = c->palloc<uint64_t>(n_all); h->sy1 = c->palloc<uint64_t>(n_all); h->sdir = c->palloc<uint8_t>(n_all);
      LAUNCH(c, k_scatter, nblk(nrec), 256, R, nrec, rec_bucket, slot2j, h->boff, fill, h->sy0, h->sy1, sseq, h->sdir);
      LAUNCH(c, k_sort_buckets, nblk(nb, 64), 64, nb, h->boff, h->sy0, h->sy1, sseq, h->sdir);
    }
    c->sync();
    if (c->check_err("build_shimmer_map4py")) { fprintf(stderr, "pgb200: %s\n", pgb_last_error(c)); exit(1); }
    c->scratch_reset();
  } catch (std::exception &e) {
    fprintf(stderr, "pgb200: build_shimmer_map4py failed: %s\n", e.what());
    exit(1);
  }
  py->mmers = &h->mmers;
  py->mmer0_map = h; py->rlmap = h; py->mcmap = h; py->ridmm = h;
}

extern "C" void get_shimmers_for_read(mm128_v *out, py_mmer_t *py, uint32_t rid) {
  PyMmerHandle *h = (PyMmerHandle *)py->ridmm;
  out->n = out->m = 0; out->a = nullptr;
  if (rid < h->rid_count.size() && h->rid_count[rid]) {
    out->n = out->m = h->rid_count[rid];
    out->a = h->mmers.a + h->rid_first[rid];
  }
}

extern "C" uint32_t get_mmer_count(py_mmer_t *py, uint64_t mhash) {
  std::lock_guard<std::recursive_mutex> lock_(g_ctx_mu);  // one device context behind the handle: calls are serialised
  PyMmerHandle *h = (PyMmerHandle *)py->mcmap;
  pgb_ctx *c = h->c;
  uint32_t v = 0;
  try {
    CU(cudaSetDevice(c->device));
    if (!c->d_mckeys) return 0;
    uint32_t *d = c->alloc<uint32_t>(1);
    LAUNCH(c, k_count_one, 1, 1, c->d_mckeys, c->d_mcvals, c->mcmask, mhash, d);
    c->d2h(&v, d, 4);
    c->scratch_reset();
  } catch (std::exception &e) { fprintf(stderr, "pgb200: get_mmer_count failed: %s\n", e.what()); exit(1); }
  return v;
}

extern "C" void get_shimmer_hits(mp256_v *out, py_mmer_t *py, uint64_t mhash0, uint32_t span) {
  std::lock_guard<std::recursive_mutex> lock_(g_ctx_mu);  // one device context behind the handle: calls are serialised
  PyMmerHandle *h = (PyMmerHandle *)py->mmer0_map;
  pgb_ctx *c = h->c;
  if (!h->xkeys) return;
  const uint64_t key = mhash0 << 8 | span;  // src/shimmer4py.c:172-173
  try {
    CU(cudaSetDevice(c->device));
    uint32_t *d = c->alloc<uint32_t>(1), o = PGB_NOSLOT;
    LAUNCH(c, k_outer_lookup, 1, 1, h->xkeys, h->xmask, h->xfirst, h->xid, key, d);
    c->d2h(&o, d, 4);
    c->scratch_reset();
    if (o == PGB_NOSLOT) return;
    uint32_t g[2];
    c->d2h(g, h->goff + o, 8);
    const uint32_t b = g[0], m = g[1] - g[0];
    std::vector<uint32_t> ipos(m), boff(m + 1);
    std::vector<uint64_t> k1(m);
    c->d2h(ipos.data(), h->ipos + b, (size_t)m * 4);
    c->d2h(boff.data(), h->boff + b, ((size_t)m + 1) * 4);
    c->d2h(k1.data(), h->gk1 + b, (size_t)m * 8);
    std::vector<uint32_t> order(m);
    for (uint32_t i = 0; i < m; i++) order[ipos[i]] = i;  // inner khash slot order (src/shimmer4py.c:181)
    const uint32_t r0 = boff[0], nr = boff[m] - boff[0];
    std::vector<uint64_t> y0(nr), y1(nr);
    std::vector<uint8_t> dir(nr);
    c->d2h(y0.data(), h->sy0 + r0, (size_t)nr * 8); c->d2h(y1.data(), h->sy1 + r0, (size_t)nr * 8); c->d2h(dir.data(), h->sdir + r0, nr);
    for (uint32_t q = 0; q < m; q++) {
      const uint32_t i = order[q];
      for (uint32_t t = boff[i] - r0; t < boff[i + 1] - r0; t++) {
        if (out->n == out->m) {
          out->m = out->m ? out->m << 1 : 2;
          out->a = (mp256_t *)realloc(out->a, out->m * sizeof(mp256_t));
        }
        mp256_t *e = &out->a[out->n++];
        memset(e, 0, sizeof *e);
        e->x0 = key; e->x1 = k1[i]; e->y0 = y0[t]; e->y1 = y1[t]; e->direction = dir[t];
      }
    }
  } catch (std::exception &e) { fprintf(stderr, "pgb200: get_shimmer_hits failed: %s\n", e.what()); exit(1); }
}
