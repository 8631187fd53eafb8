// libhnsw_b200.so — C ABI (include/hnsw_b200.h) over the sm_100a kernels.
// Host logic only: device memory layout, graph import/export, kernel launches, counters.
// There is no CPU compute path in this library; without a CUDA device every compute entry
// point fails with HNSWB200_ECUDA.
#include "../../include/hnsw_b200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <string>
#include <vector>

#include "search.cuh"
#include "build.cuh"
#include "bruteforce.cuh"
#include "bruteforce_tc.cuh"
#include "bruteforce_tc2.cuh"
#include "merge.cuh"
#include "stats.cuh"

namespace {

thread_local std::string g_err;

struct HbError : std::runtime_error {
  int code;
  HbError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] void fail(int code, const std::string& m) { throw HbError(code, m); }

#define CUDA_CHECK(x)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess) {                                                                       \
      int code_ = e_ == cudaErrorMemoryAllocation ? HNSWB200_ENOMEM : HNSWB200_ECUDA;              \
      fail(code_, std::string(#x) + ": " + cudaGetErrorString(e_));                                \
    }                                                                                              \
  } while (0)

template <class F>
int guard(F&& f) {
  try { f(); return HNSWB200_OK; }
  catch (const HbError& e) { g_err = e.what(); return e.code; }
  catch (const std::bad_alloc&) { g_err = "out of memory"; return HNSWB200_ENOMEM; }
  catch (const std::exception& e) { g_err = e.what(); return HNSWB200_ECUDA; }
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  // grow-only; contents are NOT preserved unless keep is set
  void reserve(size_t want, bool keep = false, cudaStream_t s = 0) {
    if (want <= n) return;
    T* q = nullptr;
    CUDA_CHECK(cudaMalloc(&q, want * sizeof(T)));
    if (keep && p && n) CUDA_CHECK(cudaMemcpyAsync(q, p, n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (keep && p) CUDA_CHECK(cudaStreamSynchronize(s));
    if (p) cudaFree(p);
    p = q; n = want;
  }
  // scratch that grows with the batch: geometric growth, contents dropped
  void reserve_geo(size_t want) { if (want > n) reserve(std::max(want, n + n / 2)); }
};

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct hnswb200_index {
  // configuration
  int dim = 0, metric = 0, M = 0, efC = 0, device = 0, flavour = HNSWB200_FLAVOUR_OHNSW;
  uint64_t seed = 0, rng_state = 0;
  int64_t param_hash_slots = 0, param_build_batch = 0, param_warps_per_cta = 0;
  // graph
  int64_t n = 0, cap = 0;
  int ld = 0, slots0 = 0, slotsU = 0, max_layer = 0;
  int64_t entry = -1;
  int64_t rowsU = 0, capU = 0;
  DevBuf<float> vec;
  DevBuf<int32_t> adj0, upper_off, adjU;
  DevBuf<int8_t> level;
  DevBuf<int32_t> row_owner;           // [rowsU] node owning each upper row (build)
  std::vector<int32_t> h_row_owner;
  std::vector<int8_t> h_level;        // host mirror of level (bookkeeping for build / export)
  std::vector<int32_t> h_upper_off;
  // scratch
  cudaStream_t stream = nullptr;
  cudaStream_t aux_stream[3] = {nullptr, nullptr, nullptr};   // host-buffer search: copy / compute overlap
  cudaStream_t copy_stream = nullptr;   // build: the new vectors are uploaded in pieces while the first batches run
  cudaEvent_t copy_event = nullptr;
  cudaEvent_t aux_event[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t param_host_chunks = 0;
  int64_t param_host_zero_copy = 1;     // pinned caller buffers are read / written by the search kernel itself (search_host)
  unsigned long long* h_evs = nullptr;  // pinned landing place of the event counters
  int64_t param_stage_rows = 0;         // 0 auto, -1 never stage, 4..32 rows in the ring
  int64_t param_stage_ahead = -1;       // rows beyond the ring prefetched to L2 (-1 auto, 0..31)
  int64_t param_gang = 0;               // warps per query: 0 auto (1 when the batch fills the GPU), 1, 2, 4
  int64_t param_build_ratio_early = 4;  // build: a batch is at most 1/this of the graph so far (and 1/build_ratio of the final graph)
  int64_t param_build_mates = 1;        // build: 1 = members of a batch are proposed to each other (build.cuh, mates), 0 = they never link
  int64_t param_build_qreg = 0;         // build: 1 = the new node's vector in registers (fewer resident warps), 0 = in shared memory
  int64_t param_hash_bits = 0;          // visited hash entries: 0 auto (16-bit quotiented when the id range allows), 16, 32
  unsigned int* h_ready = nullptr;      // pinned: the "pieces ready" values the copy stream writes after each piece
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  DevBuf<float> d_q, d_dists;
  DevBuf<int32_t> d_ids;
  DevBuf<uint32_t> d_counters, d_bitpool;
  DevBuf<int> d_pool_busy;
  DevBuf<unsigned int> d_next;
  DevBuf<unsigned long long> d_events;
  DevBuf<uint64_t> d_tie_pool;          // regions for tie lists that outgrow shared memory (search.cuh)
  DevBuf<int> d_tie_busy;
  int pool_size = 0, pool_words = 0;
  bool search_pending = false;          // ev1 marks the end of the last enqueued search (calls on one index are serialised on the device)
  bool poisoned = false;                // a build batch failed half-way: the graph may hold links to nodes that do not exist
  // build scratch
  DevBuf<uint64_t> b_req, b_req_sorted, b_rem, b_rem_sorted, b_mate, b_mate_sorted;
  DevBuf<unsigned int> b_heads, b_ctr;
  DevBuf<int32_t> b_order;
  std::vector<int32_t> h_order;
  DevBuf<unsigned long long> b_counters;
  DevBuf<unsigned char> b_cub;
  int64_t param_build_ratio = 64, param_max_warps_per_sm = 0, param_visited_mode = 0;   // 0 auto, 1 hash, 2 bitset
  bool layer_stats_dirty = true;        // Hgraph.Stats are recomputed only after the graph changed
  int num_sms = 0, max_smem_optin = 0;
  // stats
  hnswb200_stats st{};
  int64_t last_nq = 0;
  int last_k = 0, last_mode = 0;
  std::mutex mu;

  // float4 chunks that hold components: a row stride wider than the vector (row_floats: 128-byte aligned rows for
  // dim = 100) is padding the kernels never read
  int real_chunks() const { return (dim + 3) / 4; }

  hb::GraphView view() const {
    hb::GraphView g;
    g.vec = vec.p; g.adj0 = adj0.p; g.upper_off = upper_off.p; g.adjU = adjU.p;
    g.ld4 = ld / 4; g.chunks = real_chunks(); g.slots0 = slots0; g.slotsU = slotsU;
    g.max_layer = max_layer; g.entry = (int)entry; g.n = (int)n; g.metric = metric;
    return g;
  }
};

namespace {

void use_device(hnswb200_index* x) { CUDA_CHECK(cudaSetDevice(x->device)); }

void init_device(hnswb200_index* x) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(HNSWB200_ECUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
  if (x->device < 0 || x->device >= count) fail(HNSWB200_EINVAL, "device ordinal out of range");
  use_device(x);
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, x->device));
  x->num_sms = prop.multiProcessorCount;
  x->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  // a BLOCKING stream: work enqueued on it is ordered after everything already enqueued on the legacy
  // default stream (where a caller's copies and kernels usually run) and the other way round
  CUDA_CHECK(cudaStreamCreateWithFlags(&x->stream, cudaStreamDefault));
  CUDA_CHECK(cudaEventCreate(&x->ev0));
  CUDA_CHECK(cudaEventCreate(&x->ev1));
}

// Copy host rows [n][dim] into device rows [n][ld] (zero padded).
void upload_rows(float* dst, int ld, const float* src, int dim, int64_t n, cudaStream_t s) {
  if (n == 0) return;
  if (ld == dim) {
    CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)n * dim * sizeof(float), cudaMemcpyHostToDevice, s));
  } else {
    CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)n * ld * sizeof(float), s));
    CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(float), src, (size_t)dim * sizeof(float),
                                 (size_t)dim * sizeof(float), (size_t)n, cudaMemcpyHostToDevice, s));
  }
}

// ---- search ---------------------------------------------------------------------------------------
struct SearchPlan {
  int cpl, warps, ef_cap, hash_slots, q_chunks, smem_per_warp, grid, nb_cap;
  hb::HashCfg hc;        // visited hash geometry for hash_slots entries
  int stage_slots;      // bulk-copy ring of this many rows per warp (0: LDG gathers)
  int gang;             // warps per query (search.cuh, Gang)
  size_t smem;
};

// resident CTAs per SM of the persistent search kernel (registers and shared memory both count);
// cached: the query costs two driver calls
int search_resident(int cpl, int threads, size_t smem, bool gang = false) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, size_t>, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_tuple(cpl + (gang ? 100 : 0), threads, smem);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  int nb = 0;
  switch (cpl) {
#define HB_CASE(C)                                                                                              \
    case C:                                                                                                     \
      if (gang) {                                                                                               \
        CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb::search_kernel<C, true>, threads, smem)); \
      } else {                                                                                                  \
        CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb::search_kernel<C>, threads, smem));    \
      }                                                                                                         \
      break;
    HB_CASE(1) HB_CASE(2) HB_CASE(3) HB_CASE(4)
    default:
      if (gang) {
        CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb::search_kernel<0, true>, threads, smem));
      } else {
        CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hb::search_kernel<0>, threads, smem));
      }
#undef HB_CASE
  }
  cache[key] = std::max(1, nb);
  return std::max(1, nb);
}

// Geometry of a visited hash of `slots` entries over ids < n (common.cuh, HashCfg).  16-bit quotiented
// entries when the quotient leaves at least 5 bits of displacement and the division-by-multiplication
// is exact on the whole id range (verified here); 32-bit entries otherwise.
hb::HashCfg make_hash_cfg_uncached(const hnswb200_index* x, int slots, int64_t n) {
  hb::HashCfg hc{};
  hc.slots = (uint32_t)slots;
  hc.bytes = (uint32_t)slots * 4u;
  hc.mul = 2654435761u;
  hc.mul_inv = 1u;
  for (int i = 0; i < 5; i++) hc.mul_inv *= 2u - hc.mul * hc.mul_inv;      // Newton: inverse of an odd number mod 2^32
  if (slots <= 0 || x->param_hash_bits == 32) return hc;
  int b = 1;
  while (b < 31 && (int64_t(1) << b) < std::max<int64_t>(n, 2)) b++;
  const uint32_t mask = (uint32_t)((uint64_t(1) << b) - 1);
  const uint32_t qmax = mask / (uint32_t)slots;
  int db = 0;
  while (db < 12 && ((((uint64_t)qmax << (db + 1)) | ((1u << (db + 1)) - 1u)) <= 65534u)) db++;
  if (x->param_hash_bits < 0) db = std::min(db, (int)-x->param_hash_bits);      // tests: a short reach, so probes do run out
  else if (db < 5 && x->param_hash_bits != 16) return hc;
  if (db < 1) return hc;
  // x / slots for x <= mask: magic = ceil(2^(32+shift) / slots) when it fits 32 bits, checked at every multiple of slots
  int shift = 0;
  while ((uint64_t(1) << shift) < (uint64_t)slots) shift++;
  uint64_t magic = 0;
  for (; shift >= 0; shift--) {
    magic = ((uint64_t(1) << (32 + shift)) + (uint64_t)slots - 1) / (uint64_t)slots;
    if (magic <= 0xffffffffull) break;
  }
  if (shift < 0) return hc;
  auto div = [&](uint32_t v) { return (uint32_t)(((uint64_t)v * magic) >> 32) >> shift; };
  for (uint64_t m = 0; m <= mask; m += (uint64_t)slots) {
    if (div((uint32_t)m) != m / slots) return hc;
    if (m && div((uint32_t)(m - 1)) != (m - 1) / slots) return hc;
  }
  if (div(mask) != mask / (uint32_t)slots) return hc;
  hc.bits16 = 1; hc.bytes = (uint32_t)slots * 2u; hc.mask = mask; hc.magic = (uint32_t)magic; hc.shift = (uint32_t)shift; hc.db = (uint32_t)db;
  return hc;
}
hb::HashCfg make_hash_cfg(const hnswb200_index* x, int slots, int64_t n) {
  static std::mutex mu;
  static std::map<std::tuple<int, int64_t, int64_t>, hb::HashCfg> cache;      // the exactness check below walks the id range
  std::lock_guard<std::mutex> lk(mu);
  const auto key = std::make_tuple(slots, n, x->param_hash_bits);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  return cache[key] = make_hash_cfg_uncached(x, slots, n);
}
int hash_entry_bytes(const hnswb200_index* x, int slots, int64_t n) { return make_hash_cfg(x, slots, n).bits16 ? 2 : 4; }


// Visited set of a query: exact open-addressing hash in shared memory, or one n-bit set per warp
// in global memory.  The hash costs shared memory (fewer resident warps: throughput is linear in
// resident warps up to ~24 per SM); the bitset costs one global atomic round trip per expansion
// and an n/8-byte clear per query.  Bitset when the hash would leave fewer than 16 warps per SM
// (measured cross-over on the 1M x 128 shape: ef ~ 72)
// and the clear is small next to the vectors the query reads (~26 * ef of them).
bool use_bitset_visited(const hnswb200_index* x, int ef, int smem_per_warp_hash, int64_t n_nodes) {
  if (x->param_visited_mode == 1) return false;
  if (x->param_visited_mode == 2) return true;
  if (x->param_hash_slots > 0) return false;
  const double clear_bytes = (double)n_nodes / 8.0, query_bytes = 26.0 * ef * 4.0 * x->dim;
  return smem_per_warp_hash > (227 * 1024) / 16 - 256 && clear_bytes <= 0.3 * query_bytes;
}

// High-dimension rows (>= 1 KB, dim >= 256) are gathered with one bulk copy per row into a per-warp
// shared-memory ring (common.cuh, Stage).  Default depth: about 30 KB of rows in flight per warp
// (8 GIST rows), at least one round of 4 and at most 16 rows.
int stage_slots_for(const hnswb200_index* x, int cpl) {
  if (cpl != 0 || x->ld * 4 < hb::STAGE_MIN_BYTES || x->param_stage_rows < 0) return 0;
  if (x->param_stage_rows > 0) return (int)std::max<int64_t>(4, std::min<int64_t>(x->param_stage_rows, hb::STAGE_MAX_SLOTS));
  return std::max(4, std::min(16, (30 * 1024) / (x->ld * 4)));
}

int stage_ahead_for(const hnswb200_index* x) {
  return x->param_stage_ahead >= 0 ? (int)std::min<int64_t>(x->param_stage_ahead, 31) : 8;
}

SearchPlan plan_search(hnswb200_index* x, int ef, int64_t nq) {
  SearchPlan pl;
  const int chunks = x->ld / 4, used = x->real_chunks();      // row stride / chunks that hold components
  int cpl = (used + hb::TEAM - 1) / hb::TEAM;
  pl.cpl = cpl <= 4 ? cpl : 0;                        // register-resident query up to 128 dims
  pl.q_chunks = pl.cpl ? hb::TEAM * pl.cpl : round_up(used, 2);   // the target lives in shared memory
  pl.ef_cap = round_up(ef, 32);
  // visited hash: ~42 slots per beam entry (a query evaluates ~25-30 distances per beam entry on
  // the 1M-row shapes), kept under 75 % load; anything larger continues on a global bitset
  int hs = x->param_hash_slots > 0 ? round_up((int)x->param_hash_slots, 8) : round_up(std::max(1024, 42 * ef), 128);
  int eb = hash_entry_bytes(x, hs, x->n);                        // 2 (16-bit quotiented entries) or 4
  pl.nb_cap = std::max(x->slots0, x->slotsU) > 32 ? 64 : 32;     // list slots gathered per pass
  pl.stage_slots = stage_slots_for(x, pl.cpl);
  const int stage_bytes = pl.stage_slots ? hb::stage_smem_bytes(pl.stage_slots, chunks) : 0;
  int fixed = hb::search_smem_per_warp(pl.ef_cap, 0, pl.q_chunks, pl.nb_cap) + stage_bytes + hb::GANG_JOB_BYTES;
  if (fixed + 1024 * 4 > x->max_smem_optin) fail(HNSWB200_EINVAL, "ef too large for shared memory");
  // 16-bit entries cost a few more instructions per test-and-set: only where 32-bit ones would leave
  // fewer warps resident than the register file allows (ef >~ 42 at 128 dimensions)
  bool force32 = false;
  if (eb == 2 && x->param_hash_bits == 0 && (227 * 1024) / (fixed + hs * 4 + 256) >= 4 * HB_SEARCH_MINB) { eb = 4; force32 = true; }
  hs = std::min(hs, (x->max_smem_optin - fixed) / eb / 8 * 8);
  {
    // a table within ~12 % of what keeps the SM full is trimmed to fit (the few queries that
    // outgrow it continue on a global bitset)
    const int budget = (227 * 1024) / (4 * HB_SEARCH_MINB) - 256;
    const int hs_fit = (budget - fixed) / eb / 8 * 8;
    if (x->param_hash_slots == 0 && hs > hs_fit && hs_fit * 100 >= hs * 88) hs = hs_fit;
  }
  if (use_bitset_visited(x, ef, fixed + hs * eb, x->n)) hs = 0;
  pl.hash_slots = hs;
  pl.hc = make_hash_cfg(x, hs, x->n);
  if (force32 && pl.hc.bits16) { pl.hc.bits16 = 0; pl.hc.bytes = (uint32_t)hs * 4u; }
  pl.smem_per_warp = hb::search_smem_per_warp(pl.ef_cap, (int)pl.hc.bytes, pl.q_chunks, pl.nb_cap) + stage_bytes + hb::GANG_JOB_BYTES;
  // A batch small enough to be resident all at once (one warp per query, nq <= SMs x warps per SM) gets the
  // deepest ring with which it still is: every query then runs from the first cycle, in CTAs of one warp so
  // the SMs hold the same number of queries (1 000 GIST queries: 7 per SM with a 6-row ring).
  bool one_wave = false;
  if (pl.stage_slots && x->param_stage_rows == 0) {
    const int64_t per_sm = (nq + x->num_sms - 1) / x->num_sms;
    const int other = pl.smem_per_warp - stage_bytes, row_bytes = x->ld * 4;
    if (per_sm <= 4 * HB_SEARCH_MINB) {
      const int fit = (int)(((int64_t)(227 * 1024) / per_sm - 1024 - other - 8 * hb::STAGE_MAX_SLOTS) / row_bytes);
      if (fit >= 4) {
        one_wave = true;
        pl.stage_slots = std::min(fit, 16);
        pl.smem_per_warp = other + hb::stage_smem_bytes(pl.stage_slots, chunks);
      }
    }
  }
  // pack the SM: as many warps as shared memory and registers (64 per thread: 32 warps) allow, in CTAs of <= 4 warps
  int per_sm_warps = std::max(1, std::min(4 * HB_SEARCH_MINB, (int)((size_t)(227 * 1024) / (size_t)(pl.smem_per_warp + 256))));
  if (x->param_max_warps_per_sm > 0) per_sm_warps = std::max(1, std::min<int>(per_sm_warps, (int)x->param_max_warps_per_sm));
  int warps = x->param_warps_per_cta > 0 ? (int)std::min<int64_t>(x->param_warps_per_cta, 4) : (one_wave ? 1 : 0);
  if (warps <= 0) {                                   // CTA shape that keeps the most warps resident (1 KB reserved per CTA)
    int best = 0;
    for (int w = 4; w >= 1; w--) {
      int ctas = (int)((size_t)(227 * 1024) / ((size_t)w * pl.smem_per_warp + 1024));
      int tot = std::min(per_sm_warps / w, ctas) * w;
      if (tot > best) { best = tot; warps = w; }
    }
    warps = std::max(1, warps);
  }
  while (warps > 1 && (size_t)warps * pl.smem_per_warp > (size_t)x->max_smem_optin) warps--;
  // A batch that leaves most of the GPU's warps without a query gets a gang of 2 or 4 warps per query
  // (search.cuh, Gang): one CTA per query, the distance rounds of an expansion shared among its warps.
  pl.gang = 1;
  if (!pl.stage_slots) {
    const int64_t capacity = (int64_t)x->num_sms * per_sm_warps;
    if (x->param_gang > 0) pl.gang = (int)x->param_gang;
    else if (nq * 4 <= capacity) pl.gang = 4;
    else if (nq * 2 <= capacity) pl.gang = 2;
  }
  if (pl.gang > 1) {
    pl.warps = pl.gang;
    pl.smem = (size_t)pl.smem_per_warp;
    int per_sm = search_resident(pl.cpl, pl.warps * 32, pl.smem, true);
    per_sm = std::min(per_sm, std::max(1, per_sm_warps / pl.gang));
    pl.grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * per_sm, nq));
    return pl;
  }
  pl.warps = warps;
  pl.smem = (size_t)warps * pl.smem_per_warp;
  int per_sm = search_resident(pl.cpl, warps * 32, pl.smem);
  per_sm = std::min(per_sm, std::max(1, (per_sm_warps + warps - 1) / warps));
  int64_t need = (nq + warps - 1) / warps;
  pl.grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * per_sm, need));
  return pl;
}

template <int CPL>
void launch_search(const hb::SearchParams& p, const SearchPlan& pl, cudaStream_t s) {
  if (pl.gang > 1) {
    CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<CPL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    hb::search_kernel<CPL, true><<<pl.grid, pl.warps * 32, pl.smem, s>>>(p);
  } else {
    CUDA_CHECK(cudaFuncSetAttribute(hb::search_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    hb::search_kernel<CPL><<<pl.grid, pl.warps * 32, pl.smem, s>>>(p);
  }
  CUDA_CHECK(cudaGetLastError());
}

constexpr int TIE_SLOTS = 32, TIE_CAP = 1 << 15;     // 32 regions of 32k keys (8 MB)
void ensure_pool(hnswb200_index* x, int total_warps, int64_t n_nodes, cudaStream_t s) {
  if (!x->d_tie_pool.p) {
    x->d_tie_pool.reserve((size_t)TIE_SLOTS * TIE_CAP);
    x->d_tie_busy.reserve(TIE_SLOTS);
    CUDA_CHECK(cudaMemsetAsync(x->d_tie_busy.p, 0, TIE_SLOTS * sizeof(int), s));
  }
  int words = (int)((n_nodes + 31) / 32);
  words = round_up(std::max(words, 1), 4);
  // a spilled query (or, for large beams, every warp) borrows one n-bit set; cap the pool at 4 GiB
  int64_t max_sets = std::max<int64_t>(1, (int64_t(4) << 30) / ((int64_t)words * 4));
  int want = (int)std::min<int64_t>(total_warps, max_sets);
  if (want > x->pool_size || words > x->pool_words) {
    x->d_bitpool.release();
    x->d_bitpool.reserve((size_t)want * words);
    // cleared on the stream the kernel is launched on (a memset on another stream is not ordered before it)
    CUDA_CHECK(cudaMemsetAsync(x->d_bitpool.p, 0, (size_t)want * words * 4, s));
    x->d_pool_busy.reserve((size_t)want);
    CUDA_CHECK(cudaMemsetAsync(x->d_pool_busy.p, 0, (size_t)x->d_pool_busy.n * sizeof(int), s));
    x->pool_size = want; x->pool_words = words;
  }
}

// The device address of a pinned host buffer of `count` elements (registered with hnswb200_host_register /
// cudaHostRegister, or from cudaMallocHost): under unified addressing such memory is mapped into every device.
// Null for pageable memory, and for a buffer whose last byte is not pinned and mapped in line with its first
// (a registration that covers only part of it): those take the copies.
template <class T>
T* device_view_of_pinned(T* p, size_t count) {
  if (!p || count == 0) return nullptr;
  cudaPointerAttributes a, b;
  if (cudaPointerGetAttributes(&a, (const void*)p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  const size_t last = count * sizeof(T) - 1;
  if (cudaPointerGetAttributes(&b, (const void*)((const char*)p + last)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (b.type != cudaMemoryTypeHost || (const char*)b.devicePointer != (const char*)a.devicePointer + last) return nullptr;
  return reinterpret_cast<T*>(a.devicePointer);
}

// Device-resident queries are dense [nq][dim]; when rows are padded (ld != dim) they are re-laid
// into the index's own padded buffer first (one strided device-to-device copy).  A PINNED host buffer is
// accepted in their place: the kernel then reads each query over PCIe when a warp starts on it.
const float* padded_queries(hnswb200_index* x, const float* d_queries, int64_t nq, cudaStream_t s) {
  if (const float* m = device_view_of_pinned(d_queries, (size_t)nq * x->dim)) d_queries = m;
  if (x->ld == x->dim || nq == 0) return d_queries;
  x->d_q.reserve((size_t)nq * x->ld);
  CUDA_CHECK(cudaMemsetAsync(x->d_q.p, 0, (size_t)nq * x->ld * sizeof(float), s));
  CUDA_CHECK(cudaMemcpy2DAsync(x->d_q.p, (size_t)x->ld * sizeof(float), d_queries, (size_t)x->dim * sizeof(float),
                               (size_t)x->dim * sizeof(float), (size_t)nq, cudaMemcpyDefault, s));
  return x->d_q.p;
}

void check_search_args(hnswb200_index* x, int64_t nq, int k, int ef, int mode) {
  if (nq < 0 || k <= 0) fail(HNSWB200_EINVAL, "search: nq must be >= 0 and k > 0");
  if (ef < k) fail(HNSWB200_EINVAL, "search: ef must be >= k");
  if (mode != HNSWB200_MODE_PARITY) fail(HNSWB200_EINVAL, "search: unknown mode (only HNSWB200_MODE_PARITY exists)");
  if (x->poisoned) fail(HNSWB200_ECUDA, "the index is unusable: an earlier build/insert call failed half-way");
  if (x->n == 0 || x->entry < 0) fail(HNSWB200_EINVAL, "knn: empty hgraph");     // lib/ohnsw.ml:862
  if (ef > 4096) fail(HNSWB200_EINVAL, "search: ef > 4096 is not supported");
}

// One launch of the search kernel over `nq` queries; counters / work counter are the caller's.
void enqueue_search(hnswb200_index* x, const SearchPlan& pl, const float* d_queries, int64_t nq, int k, int ef,
                    int32_t* d_ids, float* d_dists, uint32_t* counters, unsigned int* next, cudaStream_t s,
                    int n_peer = 0, int32_t* const* peer_ids = nullptr, float* const* peer_dists = nullptr,
                    const unsigned int* ready = nullptr, unsigned int ready_step = 1, const hb::ShardTail* tail = nullptr) {
  hb::SearchParams p;
  p.ready = ready; p.ready_step = ready_step;
  if (tail) p.tail = *tail; else p.tail.n_shards = 0;
  p.g = x->view();
  p.queries = d_queries; p.nq = nq; p.ef = ef; p.k = k; p.ef_cap = pl.ef_cap;
  p.accept_ties = x->flavour == HNSWB200_FLAVOUR_HNSW_BA;
  p.pad_inf = x->flavour == HNSWB200_FLAVOUR_HNSW_BA;
  p.hc = pl.hc; p.q_smem_chunks = pl.q_chunks; p.smem_per_warp = pl.smem_per_warp; p.nb_cap = pl.nb_cap;
  p.stage_slots = pl.stage_slots;
  p.stage_ahead = stage_ahead_for(x);
  p.gang = pl.gang;
  p.out_ids = d_ids; p.out_dists = d_dists; p.counters = counters;
  p.n_peer_out = n_peer;
  for (int r = 0; r < n_peer; r++) { p.peer_ids[r] = peer_ids[r]; p.peer_dists[r] = peer_dists[r]; }
  p.next_query = next; p.bitset_pool = x->d_bitpool.p; p.pool_busy = x->d_pool_busy.p;
  p.pool_size = x->pool_size; p.words = x->pool_words; p.events = x->d_events.p;
  p.tie_pool = x->d_tie_pool.p; p.tie_busy = x->d_tie_busy.p; p.tie_slots = TIE_SLOTS; p.tie_cap = TIE_CAP;
  SearchPlan q = pl;
  q.grid = (int)std::max<int64_t>(1, std::min<int64_t>(pl.grid, pl.gang > 1 ? nq : (nq + pl.warps - 1) / pl.warps));
  static const bool trace = std::getenv("HNSWB200_TRACE") != nullptr;
  if (trace)
    fprintf(stderr, "[hnsw_b200 search] nq=%lld ef=%d cpl=%d grid=%d x %d warps, %d B smem/warp (hash %d slots, ring %d rows), gang %d, visited %s\n",
            (long long)nq, ef, pl.cpl, q.grid, pl.warps, pl.smem_per_warp, pl.hash_slots, pl.stage_slots, pl.gang,
            pl.hash_slots ? (pl.hc.bits16 ? "hash16" : "hash32") : "bitset");
  switch (pl.cpl) {
    case 1: launch_search<1>(p, q, s); break;
    case 2: launch_search<2>(p, q, s); break;
    case 3: launch_search<3>(p, q, s); break;
    case 4: launch_search<4>(p, q, s); break;
    default: launch_search<0>(p, q, s); break;
  }
  x->st.gpu_launches += 1;
}

void finish_search(hnswb200_index* x, const unsigned long long* evs, int mode) {
  x->st.search_visited_overflows = evs[0];
  x->st.search_tie_spills = evs[1];
  x->st.search_tie_overflows = evs[3];
  // Evicted candidates at exactly the beam's top distance (duplicate vectors) stay poppable: 32 of
  // them in shared memory, 32k more in a global region.  Beyond that — or when no region could be
  // had — the surplus is not revisited and the result may differ from the reference's: an error.
  if (evs[3] && mode == HNSWB200_MODE_PARITY)
    fail(HNSWB200_ECUDA, "search: the list of equal-distance candidates at the beam boundary overflowed for " +
                             std::to_string(evs[3]) + " queries (tens of thousands of duplicate vectors?); parity cannot be guaranteed");
}

void search_device(hnswb200_index* x, const float* d_queries, int64_t nq, int k, int ef, int mode,
                   int32_t* d_ids, float* d_dists, cudaStream_t s, bool own_stream, int n_peer = 0,
                   int32_t* const* peer_ids = nullptr, float* const* peer_dists = nullptr, const hb::ShardTail* tail = nullptr) {
  check_search_args(x, nq, k, ef, mode);
  if (nq == 0) return;
  SearchPlan pl = plan_search(x, ef, nq);
  // the work counter, events and per-query counters are per-index scratch: a search enqueued on another
  // stream waits for the previous one (calls on one index are serialised on the device as on the host)
  if (x->search_pending) CUDA_CHECK(cudaStreamWaitEvent(s, x->ev1, 0));
  x->d_counters.reserve((size_t)nq * 3);
  x->d_next.reserve(8);
  x->d_events.reserve(4);
  ensure_pool(x, pl.grid * pl.warps, x->n, s);
  if (pl.hash_slots == 0) pl.grid = std::max(1, std::min(pl.grid, x->pool_size / pl.warps));   // one set per warp
  CUDA_CHECK(cudaMemsetAsync(x->d_next.p, 0, sizeof(unsigned int), s));
  CUDA_CHECK(cudaMemsetAsync(x->d_events.p, 0, 4 * sizeof(unsigned long long), s));
  CUDA_CHECK(cudaEventRecord(x->ev0, s));
  enqueue_search(x, pl, d_queries, nq, k, ef, d_ids, d_dists, x->d_counters.p, x->d_next.p, s, n_peer, peer_ids, peer_dists,
                 nullptr, 1, tail);
  CUDA_CHECK(cudaEventRecord(x->ev1, s));
  x->search_pending = true;
  x->last_nq = nq;
  x->last_k = k;
  x->last_mode = mode;
  x->st.search_queries = (uint64_t)nq;
  if (own_stream) {
    CUDA_CHECK(cudaStreamSynchronize(s));
    unsigned long long evs[4];
    CUDA_CHECK(cudaMemcpy(evs, x->d_events.p, sizeof(evs), cudaMemcpyDeviceToHost));
    finish_search(x, evs, mode);
  }
}

// Ohnsw.knn_batch_bigarray with host buffers.  Pageable buffers: H2D of the queries, search, D2H of the rows.
// PINNED buffers (the Bigarray payloads after hnswb200_host_register) are not copied at all: a warp reads its
// 4*dim-byte query straight from host memory when it starts on it (one coalesced read over PCIe, ~2 us of a
// ~100 us query) and stores its k-entry row straight into the caller's ids / dists, so the transfers ride
// under the search instead of in front of and behind it (10k x 128 queries: 5.1 MB that cost 0.1 ms as a copy).
// "host_zero_copy" = 0 restores the copies.  With "host_chunks" = 2..8, batches of >= 4096 queries are streamed
// through device memory instead (see below; search_kernel_ms then includes the wait for the first piece) — off
// by default: measured on B200 it buys 1-2 % of the call (10k queries: 1.204 ms -> 1.185 ms).
constexpr int HOST_CHUNKS = 8;
void search_host(hnswb200_index* x, const float* queries, int64_t nq, int k, int ef, int mode, int32_t* ids, float* dists) {
  check_search_args(x, nq, k, ef, mode);
  if (nq == 0) return;
  SearchPlan pl = plan_search(x, ef, nq);
  if (x->search_pending) CUDA_CHECK(cudaStreamWaitEvent(x->stream, x->ev1, 0));
  ensure_pool(x, pl.grid * pl.warps, x->n, x->stream);
  if (pl.hash_slots == 0) pl.grid = std::max(1, std::min(pl.grid, x->pool_size / pl.warps));
  int C = x->param_host_chunks >= 2 ? (int)std::min<int64_t>(x->param_host_chunks, HOST_CHUNKS) : 1;
  if (nq < 4096) C = 1;
  const bool zc = C == 1 && x->param_host_zero_copy != 0;
  const float* q_map = zc && x->ld == x->dim ? device_view_of_pinned(queries, (size_t)nq * x->dim) : nullptr;   // padded rows are re-laid by the copy
  int32_t* ids_map = zc ? device_view_of_pinned(ids, (size_t)nq * k) : nullptr;
  float* d_map = zc ? device_view_of_pinned(dists, (size_t)nq * k) : nullptr;
  if (!q_map) x->d_q.reserve((size_t)nq * x->ld);
  if (ids && !ids_map) x->d_ids.reserve((size_t)nq * k);
  if (!d_map) x->d_dists.reserve((size_t)nq * k);
  x->d_counters.reserve((size_t)nq * 3);
  x->d_next.reserve(8);
  x->d_events.reserve(4);
  if (!x->h_evs) CUDA_CHECK(cudaMallocHost(&x->h_evs, 4 * sizeof(unsigned long long)));
  cudaStream_t s0 = x->stream;
  if (C > 1 && !x->aux_stream[0]) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&x->aux_stream[0], cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&x->aux_event[0], cudaEventDisableTiming));
    CUDA_CHECK(cudaMallocHost(&x->h_ready, 8 * sizeof(unsigned int)));
    for (int c = 0; c < 8; c++) x->h_ready[c] = (unsigned)(c + 1);
  }
  // d_next[0] = work counter, d_next[1] = pieces of the batch copied so far
  CUDA_CHECK(cudaMemsetAsync(x->d_next.p, 0, 8 * sizeof(unsigned int), s0));
  CUDA_CHECK(cudaMemsetAsync(x->d_events.p, 0, 4 * sizeof(unsigned long long), s0));
  int32_t* out_ids = ids ? (ids_map ? ids_map : x->d_ids.p) : nullptr;
  float* out_d = d_map ? d_map : x->d_dists.p;
  if (C == 1) {
    if (!q_map) upload_rows(x->d_q.p, x->ld, queries, x->dim, nq, s0);
    CUDA_CHECK(cudaEventRecord(x->ev0, s0));
    enqueue_search(x, pl, q_map ? q_map : x->d_q.p, nq, k, ef, out_ids, out_d, x->d_counters.p, x->d_next.p, s0);
  } else {
    // One kernel, started at once; the queries stream in behind it in C pieces on the copy stream,
    // each followed by a 4-byte "pieces ready" update the kernel's warps wait on (bounded) before
    // they read a query of that piece.  The H2D copy is hidden under the search except for piece 0.
    const unsigned step = (unsigned)((nq + C - 1) / C);
    cudaStream_t sc = x->aux_stream[0];
    CUDA_CHECK(cudaEventRecord(x->aux_event[0], s0));
    CUDA_CHECK(cudaStreamWaitEvent(sc, x->aux_event[0], 0));          // counters are zero before any piece lands
    CUDA_CHECK(cudaEventRecord(x->ev0, s0));
    enqueue_search(x, pl, x->d_q.p, nq, k, ef, out_ids, out_d, x->d_counters.p, x->d_next.p, s0, 0, nullptr, nullptr,
                   x->d_next.p + 1, step);
    for (int c = 0; c < C; c++) {
      const int64_t q0 = (int64_t)step * c, m = std::min<int64_t>(step, nq - q0);
      if (m <= 0) break;
      upload_rows(x->d_q.p + (size_t)q0 * x->ld, x->ld, queries + (size_t)q0 * x->dim, x->dim, m, sc);
      CUDA_CHECK(cudaMemcpyAsync(x->d_next.p + 1, &x->h_ready[c], sizeof(unsigned int), cudaMemcpyHostToDevice, sc));
    }
  }
  CUDA_CHECK(cudaEventRecord(x->ev1, s0));
  if (ids && !ids_map) CUDA_CHECK(cudaMemcpyAsync(ids, x->d_ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s0));
  if (!d_map) CUDA_CHECK(cudaMemcpyAsync(dists, x->d_dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s0));
  CUDA_CHECK(cudaMemcpyAsync(x->h_evs, x->d_events.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s0));
  x->search_pending = true;
  x->last_nq = nq;
  x->last_k = k;
  x->last_mode = mode;
  x->st.search_queries = (uint64_t)nq;
  x->st.search_zero_copy = (q_map ? 1u : 0u) | (d_map && (!ids || ids_map) ? 2u : 0u);
  CUDA_CHECK(cudaStreamSynchronize(s0));
  if (C > 1) CUDA_CHECK(cudaStreamSynchronize(x->aux_stream[0]));
  if (x->h_evs[2]) fail(HNSWB200_ECUDA, "search: the query copy did not arrive (streamed host buffers)");
  finish_search(x, x->h_evs, mode);
}

// ---- import / export ----------------------------------------------------------------------------------
void allocate_graph(hnswb200_index* x, int64_t cap, int64_t capU, bool keep) {
  use_device(x);
  x->vec.reserve((size_t)cap * x->ld, keep, x->stream);
  x->adj0.reserve((size_t)cap * x->slots0, keep, x->stream);
  x->upper_off.reserve((size_t)cap, keep, x->stream);
  x->level.reserve((size_t)cap, keep, x->stream);
  x->adjU.reserve((size_t)std::max<int64_t>(capU, 1) * x->slotsU, keep, x->stream);
  x->cap = cap; x->capU = std::max<int64_t>(capU, 1);
}

void import_graph(hnswb200_index* x, const float* data, int64_t n, int id_base, int max_layer, int64_t entry,
                  const int64_t* const* offs, const int32_t* const* nbrs) {
  if (n <= 0) fail(HNSWB200_EINVAL, "import_graph: n must be > 0");
  if (n >= (int64_t(1) << 31) - 1) fail(HNSWB200_EINVAL, "import_graph: n too large");
  if (max_layer < 0 || max_layer > 15) fail(HNSWB200_EINVAL, "import_graph: max_layer must be in 0..15");
  entry -= id_base;
  if (entry < 0 || entry >= n) fail(HNSWB200_EINVAL, "Hgraph.set_entry_point: invalid node");   // lib/ohnsw.ml:343
  int maxdeg0 = 0, maxdegU = 0;
  std::vector<int8_t> lvl((size_t)n, 0);
  for (int l = 0; l <= max_layer; l++) {
    const int64_t* o = offs[l];
    if (o[0] != 0) fail(HNSWB200_EINVAL, "import_graph: offsets must start at 0");
    for (int64_t i = 0; i < n; i++) {
      int64_t deg = o[i + 1] - o[i];
      if (deg < 0) fail(HNSWB200_EINVAL, "import_graph: offsets must be non-decreasing");
      if (deg > 4096) fail(HNSWB200_EINVAL, "import_graph: degree > 4096");
      if (l == 0) maxdeg0 = std::max<int>(maxdeg0, (int)deg);
      else { maxdegU = std::max<int>(maxdegU, (int)deg); if (deg > 0) lvl[i] = (int8_t)l; }
    }
  }
  lvl[(size_t)entry] = (int8_t)max_layer;
  x->slots0 = std::max(2 * x->M, maxdeg0);
  x->slotsU = std::max(x->M, maxdegU);
  std::vector<int32_t> uoff((size_t)n, -1);
  int64_t rows = 0;
  for (int64_t i = 0; i < n; i++) if (lvl[i] > 0) { uoff[i] = (int32_t)rows; rows += lvl[i]; }
  if (rows >= (int64_t(1) << 31)) fail(HNSWB200_EINVAL, "import_graph: too many upper-layer rows");
  std::vector<int32_t> a0((size_t)n * x->slots0, -1), aU((size_t)std::max<int64_t>(rows, 1) * x->slotsU, -1);
  std::vector<int32_t> seen;
  for (int l = 0; l <= max_layer; l++) {
    const int64_t* o = offs[l];
    for (int64_t i = 0; i < n; i++) {
      int64_t deg = o[i + 1] - o[i];
      if (!deg) continue;
      int32_t* row = l == 0 ? &a0[(size_t)i * x->slots0] : &aU[((size_t)uoff[i] + l - 1) * x->slotsU];
      seen.assign(nbrs[l] + o[i], nbrs[l] + o[i + 1]);
      for (int64_t j = 0; j < deg; j++) {
        int64_t v = (int64_t)seen[j] - id_base;
        if (v < 0 || v >= n) fail(HNSWB200_EINVAL, "import_graph: neighbour id out of range");
        row[j] = (int32_t)v;
      }
      std::sort(seen.begin(), seen.end());
      if (std::adjacent_find(seen.begin(), seen.end()) != seen.end())
        fail(HNSWB200_EINVAL, "import_graph: duplicate neighbour id in a row");
    }
  }
  x->vec.release(); x->adj0.release(); x->upper_off.release(); x->level.release(); x->adjU.release();
  allocate_graph(x, n, rows, false);
  upload_rows(x->vec.p, x->ld, data, x->dim, n, x->stream);
  CUDA_CHECK(cudaMemcpyAsync(x->adj0.p, a0.data(), a0.size() * 4, cudaMemcpyHostToDevice, x->stream));
  CUDA_CHECK(cudaMemcpyAsync(x->adjU.p, aU.data(), aU.size() * 4, cudaMemcpyHostToDevice, x->stream));
  CUDA_CHECK(cudaMemcpyAsync(x->upper_off.p, uoff.data(), uoff.size() * 4, cudaMemcpyHostToDevice, x->stream));
  CUDA_CHECK(cudaMemcpyAsync(x->level.p, lvl.data(), lvl.size(), cudaMemcpyHostToDevice, x->stream));
  CUDA_CHECK(cudaStreamSynchronize(x->stream));
  x->n = n; x->rowsU = rows; x->max_layer = max_layer; x->entry = entry;
  x->h_level = lvl; x->h_upper_off = uoff;
  x->h_row_owner.clear(); x->row_owner.release();   // rebuilt on the first insert
  x->layer_stats_dirty = true;
}

// host copy of one layer's rows: deg[i], and the row contents
void download_layer(hnswb200_index* x, int layer, std::vector<int32_t>& rows, int& slots) {
  use_device(x);
  if (layer == 0) {
    slots = x->slots0;
    rows.resize((size_t)x->n * slots);
    CUDA_CHECK(cudaMemcpy(rows.data(), x->adj0.p, rows.size() * 4, cudaMemcpyDeviceToHost));
  } else {
    slots = x->slotsU;
    rows.resize((size_t)std::max<int64_t>(x->rowsU, 1) * slots);
    CUDA_CHECK(cudaMemcpy(rows.data(), x->adjU.p, rows.size() * 4, cudaMemcpyDeviceToHost));
  }
}
const int32_t* layer_row(hnswb200_index* x, int layer, const std::vector<int32_t>& rows, int slots, int64_t i) {
  if (layer == 0) return &rows[(size_t)i * slots];
  if (x->h_level[(size_t)i] < layer) return nullptr;
  return &rows[((size_t)x->h_upper_off[(size_t)i] + layer - 1) * slots];
}
int row_degree(const int32_t* row, int slots) {
  int d = 0;
  while (row && d < slots && row[d] >= 0) d++;
  return d;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* hnswb200_last_error(void) { return g_err.c_str(); }
const char* hnswb200_version(void) { return "hnsw_b200 0.1 (sm_100a)"; }

int hnswb200_create(hnswb200_index** out, int dim, int metric, int M, int ef_construction, uint64_t seed, int device) {
  return guard([&] {
    if (!out) fail(HNSWB200_EINVAL, "create: out is NULL");
    *out = nullptr;
    if (dim <= 0 || dim > 65536) fail(HNSWB200_EINVAL, "create: dim must be in 1..65536");
    if (metric < 0 || metric > 2) fail(HNSWB200_EINVAL, "create: unknown metric");
    if (M < 2 || M > 512) fail(HNSWB200_EINVAL, "create: num_connections must be in 2..512 (level_mult = 1/ln M)");
    if (ef_construction < 1 || ef_construction > 4096) fail(HNSWB200_EINVAL, "create: num_nodes_search_construction must be in 1..4096");
    hnswb200_index* x = new hnswb200_index();
    x->dim = dim; x->metric = metric; x->M = M; x->efC = ef_construction; x->seed = seed; x->rng_state = seed;
    x->device = device;
    x->ld = round_up(dim, 4);
    x->slots0 = 2 * M; x->slotsU = M;
    try { init_device(x); } catch (...) { delete x; throw; }
    *out = x;
  });
}

int hnswb200_set_flavour(hnswb200_index* x, int flavour) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    if (flavour != HNSWB200_FLAVOUR_OHNSW && flavour != HNSWB200_FLAVOUR_HNSW_BA) fail(HNSWB200_EINVAL, "unknown flavour");
    x->flavour = flavour;
  });
}

int hnswb200_set_param(hnswb200_index* x, const char* name, int64_t value) {
  return guard([&] {
    if (!x || !name) fail(HNSWB200_EINVAL, "index or name is NULL");
    std::string s(name);
    if (s == "hash_slots") x->param_hash_slots = value;
    else if (s == "build_batch") x->param_build_batch = value;
    else if (s == "warps_per_cta") x->param_warps_per_cta = value;
    else if (s == "build_ratio") x->param_build_ratio = value;
    else if (s == "max_warps_per_sm") x->param_max_warps_per_sm = value;
    else if (s == "visited_mode") x->param_visited_mode = value;
    else if (s == "host_chunks") x->param_host_chunks = value;
    else if (s == "host_zero_copy") x->param_host_zero_copy = value;
    else if (s == "stage_rows") x->param_stage_rows = value;
    else if (s == "stage_ahead") x->param_stage_ahead = value;
    else if (s == "hash_bits") x->param_hash_bits = value;
    else if (s == "build_qreg") x->param_build_qreg = value;
    else if (s == "build_mates") x->param_build_mates = value;
    else if (s == "build_ratio_early") x->param_build_ratio_early = value;
    else if (s == "gang") {
      if (value != 0 && value != 1 && value != 2 && value != 4) fail(HNSWB200_EINVAL, "gang must be 0 (automatic), 1, 2 or 4");
      x->param_gang = value;
    }
    else if (s == "row_floats") {             // vector row stride in floats (multiple of 4, >= dim); only on an empty index
      if (x->n != 0) fail(HNSWB200_EINVAL, "row_floats can only be set on an empty index");
      if (value < x->dim || value % 4 != 0) fail(HNSWB200_EINVAL, "row_floats must be a multiple of 4 and >= dim");
      x->ld = (int)value;
    }
    else fail(HNSWB200_EINVAL, "unknown parameter: " + s);
  });
}

int hnswb200_destroy(hnswb200_index* x) {
  return guard([&] {
    if (!x) return;
    cudaSetDevice(x->device);
    if (x->stream) cudaStreamSynchronize(x->stream);
    if (x->ev0) cudaEventDestroy(x->ev0);
    if (x->ev1) cudaEventDestroy(x->ev1);
    if (x->h_ready) cudaFreeHost(x->h_ready);
    if (x->h_evs) cudaFreeHost(x->h_evs);
    if (x->copy_stream) { cudaStreamSynchronize(x->copy_stream); cudaStreamDestroy(x->copy_stream); }
    if (x->copy_event) cudaEventDestroy(x->copy_event);
    for (cudaStream_t a : x->aux_stream) if (a) cudaStreamDestroy(a);
    for (cudaEvent_t e : x->aux_event) if (e) cudaEventDestroy(e);
    if (x->stream) cudaStreamDestroy(x->stream);
    delete x;
  });
}

int hnswb200_search(hnswb200_index* x, const float* queries, int64_t nq, int k, int ef, int mode, int32_t* ids, float* dists) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    if (nq > 0 && (!queries || !dists)) fail(HNSWB200_EINVAL, "search: queries/dists is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    search_host(x, queries, nq, k, ef, mode, ids, dists);
  });
}

int hnswb200_search_device(hnswb200_index* x, const float* d_queries, int64_t nq, int k, int ef, int mode,
                           int32_t* d_ids, float* d_dists, void* stream) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    if (nq > 0 && (!d_queries || !d_dists)) fail(HNSWB200_EINVAL, "search_device: queries/dists is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    cudaStream_t s = stream ? (cudaStream_t)stream : x->stream;
    search_device(x, padded_queries(x, d_queries, nq, s), nq, k, ef, mode, d_ids, d_dists, s, stream == nullptr);
  });
}

int hnswb200_search_device_multi(hnswb200_index* x, const float* d_queries, int64_t nq, int k, int ef, int mode, int n_out,
                                 int32_t* const* d_ids_list, float* const* d_dists_list, void* stream) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    if (n_out < 1 || n_out > 8 || !d_ids_list || !d_dists_list) fail(HNSWB200_EINVAL, "search_device_multi: 1..8 destinations");
    for (int r = 0; r < n_out; r++) if (!d_ids_list[r] || !d_dists_list[r]) fail(HNSWB200_EINVAL, "search_device_multi: NULL destination");
    if (nq > 0 && !d_queries) fail(HNSWB200_EINVAL, "search_device_multi: queries is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    cudaStream_t s = stream ? (cudaStream_t)stream : x->stream;
    search_device(x, padded_queries(x, d_queries, nq, s), nq, k, ef, mode, nullptr, nullptr, s, stream == nullptr, n_out, d_ids_list, d_dists_list);
  });
}

int hnswb200_search_device_sharded(hnswb200_index* x, const float* d_queries, int64_t nq, int k, int ef, int mode, int shard,
                                   int n_shards, int64_t first_row, int32_t* g_ids, float* g_dists, uint32_t* arrive, int n_final,
                                   int32_t* const* f_ids, float* const* f_dists, void* stream) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    if (n_shards < 1 || n_shards > 32 || shard < 0 || shard >= n_shards) fail(HNSWB200_EINVAL, "search_device_sharded: shard / n_shards out of range");
    if (n_final < 1 || n_final > 8 || !f_ids || !f_dists || !g_ids || !g_dists || !arrive) fail(HNSWB200_EINVAL, "search_device_sharded: NULL buffer or n_final outside 1..8");
    if (first_row < 0 || first_row + x->n > (int64_t(1) << 31) - 1) fail(HNSWB200_EINVAL, "search_device_sharded: global ids exceed int32");
    if (nq > 0 && !d_queries) fail(HNSWB200_EINVAL, "search_device_sharded: queries is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    hb::ShardTail t{};
    t.n_shards = n_shards; t.shard = shard; t.id_offset = (int32_t)first_row;
    t.g_ids = g_ids; t.g_dists = g_dists; t.arrive = arrive; t.n_final = n_final;
    for (int r = 0; r < n_final; r++) {
      if (!f_ids[r] || !f_dists[r]) fail(HNSWB200_EINVAL, "search_device_sharded: NULL destination");
      t.f_ids[r] = f_ids[r]; t.f_dists[r] = f_dists[r];
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : x->stream;
    search_device(x, padded_queries(x, d_queries, nq, s), nq, k, ef, mode, nullptr, nullptr, s, stream == nullptr, 0, nullptr, nullptr, &t);
  });
}

int hnswb200_last_search_counters(hnswb200_index* x, uint32_t* out, int64_t nq) {
  return guard([&] {
    if (!x || !out) fail(HNSWB200_EINVAL, "index or out is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    if (nq != x->last_nq) fail(HNSWB200_EINVAL, "last_search_counters: nq does not match the last search");
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(out, x->d_counters.p, (size_t)nq * 3 * 4, cudaMemcpyDeviceToHost));
  });
}

int hnswb200_import_graph(hnswb200_index* x, const float* data, int64_t n, int id_base, int max_layer, int64_t entry,
                          const int64_t* const* layer_offsets, const int32_t* const* layer_nbrs) {
  return guard([&] {
    if (!x || !data || !layer_offsets || !layer_nbrs) fail(HNSWB200_EINVAL, "import_graph: NULL argument");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    import_graph(x, data, n, id_base, max_layer, entry, layer_offsets, layer_nbrs);
  });
}

int hnswb200_export_layer(hnswb200_index* x, int layer, int id_base, int64_t* offsets, int32_t* nbrs, int64_t* nnz) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    if (layer < 0 || layer > x->max_layer) fail(HNSWB200_EINVAL, "export_layer: no such layer");
    CUDA_CHECK(cudaStreamSynchronize(x->stream));
    std::vector<int32_t> rows; int slots = 0;
    download_layer(x, layer, rows, slots);
    int64_t o = 0;
    for (int64_t i = 0; i < x->n; i++) {
      const int32_t* row = layer_row(x, layer, rows, slots, i);
      int d = row_degree(row, slots);
      if (offsets) offsets[i] = o;
      if (nbrs) for (int j = 0; j < d; j++) nbrs[o + j] = row[j] + id_base;
      o += d;
    }
    if (offsets) offsets[x->n] = o;
    if (nnz) *nnz = o;
  });
}

int hnswb200_export_levels(hnswb200_index* x, int32_t* levels) {
  return guard([&] {
    if (!x || !levels) fail(HNSWB200_EINVAL, "index or levels is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    for (int64_t i = 0; i < x->n; i++) levels[i] = x->h_level[(size_t)i];
  });
}

int hnswb200_get_info(hnswb200_index* x, hnswb200_info* out) {
  return guard([&] {
    if (!x || !out) fail(HNSWB200_EINVAL, "index or out is NULL");
    out->n = x->n; out->dim = x->dim; out->metric = x->metric; out->M = x->M; out->ef_construction = x->efC;
    out->max_layer = x->max_layer; out->entry_point = x->n ? x->entry : -1; out->slots0 = x->slots0;
    out->slots_upper = x->slotsU; out->flavour = x->flavour; out->device = x->device;
  });
}

int hnswb200_get_stats(hnswb200_index* x, hnswb200_stats* out) {
  return guard([&] {
    if (!x || !out) fail(HNSWB200_EINVAL, "index or out is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    CUDA_CHECK(cudaDeviceSynchronize());
    hnswb200_stats& st = x->st;
    // search counters: reduce the per-query rows of the last call
    st.search_n_dist = st.search_n_exp0 = st.search_n_expU = 0;
    if (x->last_nq > 0) {
      std::vector<uint32_t> c((size_t)x->last_nq * 3);
      CUDA_CHECK(cudaMemcpy(c.data(), x->d_counters.p, c.size() * 4, cudaMemcpyDeviceToHost));
      for (int64_t i = 0; i < x->last_nq; i++) { st.search_n_dist += c[3 * i]; st.search_n_exp0 += c[3 * i + 1]; st.search_n_expU += c[3 * i + 2]; }
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, x->ev0, x->ev1) == cudaSuccess) st.search_kernel_ms = ms;
      else cudaGetLastError();
      // a search that was only enqueued on a caller's stream never went through finish_search
      unsigned long long evs[4];
      CUDA_CHECK(cudaMemcpy(evs, x->d_events.p, sizeof(evs), cudaMemcpyDeviceToHost));
      st.search_visited_overflows = evs[0]; st.search_tie_spills = evs[1]; st.search_tie_overflows = evs[3];
    }
    // lib/hnsw.ml:732-751 counts distance calls; bytes per SURVEY.md 8d
    st.search_algorithmic_bytes = (double)st.search_n_dist * 4.0 * x->dim + (double)st.search_n_exp0 * 4.0 * x->slots0 +
                                  (double)st.search_n_expU * 4.0 * x->slotsU +
                                  (double)x->last_nq * (4.0 * x->dim + 8.0 * x->last_k);
    // Hgraph.Stats (lib/hnsw.ml:353-375); recomputed only after the graph changed
    st.num_layers = x->n ? x->max_layer + 1 : 0;
    if (x->layer_stats_dirty && st.num_layers > 0) {
      // a reduction on the device (stats.cuh): 5 numbers per layer come back, not the adjacency arrays
      const int L = std::min(st.num_layers, 16);
      std::vector<hb::LayerStats> hs((size_t)L);
      for (auto& r : hs) { r.nodes = r.degree_sum = r.isolated = 0; r.min_degree = 0x7fffffff; r.max_degree = -1; }
      DevBuf<hb::LayerStats> ds;
      ds.reserve((size_t)L);
      CUDA_CHECK(cudaMemcpy(ds.p, hs.data(), hs.size() * sizeof(hb::LayerStats), cudaMemcpyHostToDevice));
      const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * 8, (x->n + 255) / 256));
      hb::layer_stats_kernel<<<grid, 256, 0, x->stream>>>(x->view(), x->level.p, L, ds.p);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaStreamSynchronize(x->stream));
      CUDA_CHECK(cudaMemcpy(hs.data(), ds.p, hs.size() * sizeof(hb::LayerStats), cudaMemcpyDeviceToHost));
      st.gpu_launches += 1;
      for (int l = 0; l < L; l++) {
        const hb::LayerStats& r = hs[(size_t)l];
        st.layer_nodes[l] = (int64_t)r.nodes; st.layer_min_degree[l] = r.nodes ? r.min_degree : 0; st.layer_max_degree[l] = r.nodes ? r.max_degree : 0;
        st.layer_mean_degree[l] = r.nodes ? (double)r.degree_sum / (double)r.nodes : 0.0; st.layer_isolated[l] = (int64_t)r.isolated;
      }
    }
    x->layer_stats_dirty = false;
    *out = st;
  });
}

int hnswb200_recall(const float* expected, const float* got, int64_t nq, int k, double epsilon, double* out) {
  return guard([&] {
    // Recall.compute (benchmark/dataset.ml:105-127): trivial host arithmetic on two k x nq mats
    if (!expected || !got || !out) fail(HNSWB200_EINVAL, "recall: NULL argument");
    if (nq <= 0 || k <= 0) fail(HNSWB200_EINVAL, "Recall.compute: arrrays have unequal shapes");
    double ret = 0.;
    for (int64_t q = 0; q < nq; q++) {
      int ok = 0;
      for (int i = 0; i < k; i++)
        if ((double)got[q * k + i] <= (double)expected[q * k + (k - 1)] + epsilon) ok++;
      ret += (double)ok / (double)k;
    }
    *out = ret / (double)nq;
  });
}

int hnswb200_host_register(const void* ptr, int64_t bytes) {
  return guard([&] {
    if (!ptr || bytes <= 0) fail(HNSWB200_EINVAL, "host_register: bad argument");
    CUDA_CHECK(cudaHostRegister(const_cast<void*>(ptr), (size_t)bytes, cudaHostRegisterDefault));
  });
}
int hnswb200_host_unregister(const void* ptr) {
  return guard([&] {
    if (!ptr) fail(HNSWB200_EINVAL, "host_unregister: NULL");
    CUDA_CHECK(cudaHostUnregister(const_cast<void*>(ptr)));
  });
}

}  // extern "C"

#include "api_build.inl"
#include "api_eval.inl"
#include "api_sharded.inl"
