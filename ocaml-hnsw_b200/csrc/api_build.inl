// C ABI: index construction (included by hnsw_b200.cu).  Host orchestration of build.cuh:
// level draw, storage growth, the batch schedule and the three phases per batch.
#include <cub/device/device_radix_sort.cuh>

namespace {

// lib/ohnsw.ml:781 — level = round_nearest(-ln U * mL), U uniform; the reference draws U from
// OCaml's global Random state, which cannot be reproduced: callers that compare against a
// reference/oracle build pass the levels in.  Same splitmix64 stream as the oracle's draw_level.
int draw_level(hnswb200_index* x) {
  x->rng_state += 0x9E3779B97F4A7C15ull;
  uint64_t z = x->rng_state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  double u = ((double)(z >> 11) + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
  double mL = 1.0 / std::log((double)x->M);                             // :844
  return (int)std::floor(-std::log(u) * mL + 0.5);
}

// resident CTAs per SM of a persistent kernel (registers and shared memory both count)
template <class K>
int resident_ctas(K kernel, int threads, size_t smem) {
  CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int nb = 0;
  CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem));
  return std::max(1, nb);
}
int build_search_resident(int cpl, bool qreg, int threads, size_t smem) {
  switch (cpl) {
#define HB_RES(C) case C: return qreg ? resident_ctas(hb::build_search_kernel<C, false, true>, threads, smem) \
                                      : resident_ctas(hb::build_search_kernel<C, false, false>, threads, smem);
    HB_RES(1) HB_RES(2) HB_RES(3) HB_RES(4)
#undef HB_RES
    default: return resident_ctas(hb::build_search_kernel<0>, threads, smem);
  }
}
int build_link_resident(int cpl, int threads, size_t smem) {
  switch (cpl) {
    case 1: return resident_ctas(hb::build_link_kernel<1>, threads, smem);
    case 2: return resident_ctas(hb::build_link_kernel<2>, threads, smem);
    case 3: return resident_ctas(hb::build_link_kernel<3>, threads, smem);
    case 4: return resident_ctas(hb::build_link_kernel<4>, threads, smem);
    default: return resident_ctas(hb::build_link_kernel<0>, threads, smem);
  }
}

struct BuildPlan {
  SearchPlan sp;
  int sel_cap, ucap, link_warps, link_smem_per_warp, link_grid_per_sm;
  int sel0, selU, cap0, capU, keep_all;
  int grid_regs;        // phase-1 grid of the variant that keeps the new node's vector in registers (fewer resident warps)
};

BuildPlan plan_build(hnswb200_index* x, int64_t n_total) {
  BuildPlan bp;
  const bool ba = x->flavour == HNSWB200_FLAVOUR_HNSW_BA;
  bp.sel0 = ba ? x->M : 2 * x->M;              // lib/ohnsw.ml:818 vs lib/hnsw.ml:753-758 (Q3)
  bp.selU = x->M;
  bp.cap0 = 2 * x->M; bp.capU = x->M;
  bp.keep_all = ba ? 1 : 0;
  int max_slots = std::max(x->slots0, x->slotsU);
  bp.sel_cap = round_up(std::max(max_slots, 4), 4);
  bp.ucap = round_up(max_slots + hb::LINK_MCAP, 4);
  // phase 1: the search plan for ef = efC plus the second target copy and the selected list
  SearchPlan& pl = bp.sp;
  int ef = x->efC;
  const int chunks = x->ld / 4, used = x->real_chunks();
  int cpl = (used + hb::TEAM - 1) / hb::TEAM;
  pl.cpl = cpl <= 4 ? cpl : 0;
  pl.q_chunks = pl.cpl ? hb::TEAM * pl.cpl : round_up(used, 2);      // (a gang reads the target from shared memory)
  pl.gang = 1;
  pl.ef_cap = round_up(ef, 32);
  pl.stage_slots = stage_slots_for(x, pl.cpl);
  const int stage_bytes = pl.stage_slots ? hb::stage_smem_bytes(pl.stage_slots, chunks) : 0;
  int extra = pl.q_chunks * 16 + bp.sel_cap * 4 + stage_bytes + hb::GANG_JOB_BYTES;
  int hs = x->param_hash_slots > 0 ? round_up((int)x->param_hash_slots, 8) : round_up(std::max(1024, 32 * ef), 128);
  const int eb = hash_entry_bytes(x, hs, n_total);
  pl.nb_cap = std::max(x->slots0, x->slotsU) > 32 ? 64 : 32;
  int fixed = hb::search_smem_per_warp(pl.ef_cap, 0, pl.q_chunks, pl.nb_cap) + extra;
  if (fixed + 1024 * 4 > x->max_smem_optin) fail(HNSWB200_EINVAL, "build: num_nodes_search_construction too large for shared memory");
  hs = std::min(hs, (x->max_smem_optin - fixed) / eb / 8 * 8);
  if (x->param_visited_mode != 1 && x->param_hash_slots == 0 && fixed + hs * eb > 14 * 1024 &&
      (double)n_total / 8.0 <= 0.3 * 26.0 * ef * 4.0 * x->dim) hs = 0;      // see use_bitset_visited
  if (x->param_visited_mode == 2) hs = 0;
  pl.hash_slots = hs;
  pl.hc = make_hash_cfg(x, hs, n_total);
  pl.smem_per_warp = hb::search_smem_per_warp(pl.ef_cap, (int)pl.hc.bytes, pl.q_chunks, pl.nb_cap) + extra;
  if (pl.smem_per_warp > x->max_smem_optin) fail(HNSWB200_EINVAL, "build: num_nodes_search_construction too large for shared memory");
  // pack the SM: as many warps as shared memory allows (<= 32), split into CTAs of <= 8 warps
  int per_sm_warps = std::max(1, std::min(32, (int)((size_t)(227 * 1024) / (size_t)(pl.smem_per_warp + 128))));
  int ctas = (per_sm_warps + 7) / 8;
  pl.warps = std::max(1, per_sm_warps / ctas);
  while (pl.warps > 1 && (size_t)pl.warps * pl.smem_per_warp > (size_t)x->max_smem_optin) pl.warps--;
  pl.smem = (size_t)pl.warps * pl.smem_per_warp;
  // two variants (build.cuh): vector in registers (lower latency per insert: batches smaller than the GPU) or in
  // shared memory only (more resident warps: batches that fill it); build_qreg = 1 / 2 forces one of them
  bp.grid_regs = x->num_sms * build_search_resident(pl.cpl, true, pl.warps * 32, pl.smem);
  pl.grid = x->param_build_qreg == 1 ? bp.grid_regs : x->num_sms * build_search_resident(pl.cpl, false, pl.warps * 32, pl.smem);
  // phase 2
  bp.link_smem_per_warp = hb::link_smem_per_warp(bp.ucap, bp.sel_cap, pl.q_chunks) + stage_bytes;
  bp.link_warps = 8;
  while (bp.link_warps > 1 && (size_t)bp.link_warps * bp.link_smem_per_warp > (size_t)x->max_smem_optin) bp.link_warps--;
  if ((size_t)bp.link_smem_per_warp > (size_t)x->max_smem_optin) fail(HNSWB200_EINVAL, "build: dimension / num_connections too large for shared memory");
  bp.link_grid_per_sm = build_link_resident(pl.cpl, bp.link_warps * 32, (size_t)bp.link_warps * bp.link_smem_per_warp);
  return bp;
}

template <int CPL, bool GANG, bool QREG>
void launch_build_search_v(const hb::BuildParams& p, const SearchPlan& pl, int grid, cudaStream_t s) {
  CUDA_CHECK(cudaFuncSetAttribute(hb::build_search_kernel<CPL, GANG, QREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  hb::build_search_kernel<CPL, GANG, QREG><<<grid, pl.warps * 32, pl.smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}
// qreg: the new node's vector in registers (128 registers, 16 warps per SM at dim 128) or in shared memory only
// (<= 85 registers, 24 warps per SM); dim > 128 (CPL = 0) has one variant
template <int CPL>
void launch_build_search(const hb::BuildParams& p, const SearchPlan& pl, int grid, bool qreg, cudaStream_t s) {
  if (CPL == 0) qreg = true;
  if (p.sp.gang > 1) { if (qreg) launch_build_search_v<CPL, true, true>(p, pl, grid, s); else launch_build_search_v<CPL, true, CPL == 0>(p, pl, grid, s); }
  else { if (qreg) launch_build_search_v<CPL, false, true>(p, pl, grid, s); else launch_build_search_v<CPL, false, CPL == 0>(p, pl, grid, s); }
}
template <int CPL>
void launch_build_link(const hb::BuildParams& p, int warps, size_t smem, int grid, cudaStream_t s) {
  CUDA_CHECK(cudaFuncSetAttribute(hb::build_link_kernel<CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  hb::build_link_kernel<CPL><<<grid, warps * 32, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

void sort_keys(hnswb200_index* x, const uint64_t* in, uint64_t* out, unsigned n, int end_bit) {
  size_t bytes = 0;
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, (int)n, 0, end_bit, x->stream));
  x->b_cub.reserve_geo(bytes + 16);
  CUDA_CHECK(cub::DeviceRadixSort::SortKeys(x->b_cub.p, bytes, in, out, (int)n, 0, end_bit, x->stream));
  x->st.gpu_launches += 3;
}

enum { CTR_REQ = 0, CTR_REM = 1, CTR_HEADS = 2, CTR_NEXT = 3, CTR_MATE = 4, CTR_N = 5 };

// HNSWB200_BUILD_TRACE=1: host wall time per phase (each ends at a stream synchronisation)
struct BuildTrace {
  bool on = std::getenv("HNSWB200_BUILD_TRACE") != nullptr;
  double t_search = 0, t_link = 0, t_rest = 0, t_sort = 0, t_alloc = 0, t_mates = 0;
  uint64_t n_mates = 0;
  int64_t batches = 0, small = 0;
  // per batch-size bucket (log2 B): batches, inserts, phase-1 seconds
  int64_t bk_n[32] = {0}, bk_ins[32] = {0};
  double bk_t[32] = {0};
  std::chrono::steady_clock::time_point t0;
  void start() { if (on) t0 = std::chrono::steady_clock::now(); }
  double lap() {
    if (!on) return 0;
    auto t1 = std::chrono::steady_clock::now();
    double d = std::chrono::duration<double>(t1 - t0).count();
    t0 = t1;
    return d;
  }
};
BuildTrace g_trace;

// The new vectors travel to the device in pieces on a copy stream while the first batches are built: a batch
// only needs the rows below its own end.  (The source is the caller's pageable memory: one 512 MB copy in
// front of the build costs ~50 ms in which the GPU does nothing.)
struct RowUploader {
  hnswb200_index* x = nullptr;
  const float* src = nullptr;     // [n_new][dim]
  int64_t base = 0, total = 0;    // device rows [base, base + total)
  int64_t sent = 0;               // rows handed to the copy stream so far
  bool unseen = false;            // pieces sent since the build stream last waited for the copy stream
  static constexpr int64_t PIECE_BYTES = 48ll << 20;
  int64_t piece_rows() const { return std::max<int64_t>(1024, PIECE_BYTES / ((int64_t)x->ld * 4)); }
  void send(int64_t upto) {       // make sure rows below `upto` (relative to base) are on their way
    upto = std::min(upto, total);
    if (upto <= sent) return;
    upload_rows(x->vec.p + (size_t)(base + sent) * x->ld, x->ld, src + (size_t)sent * x->dim, x->dim, upto - sent, x->copy_stream);
    CUDA_CHECK(cudaEventRecord(x->copy_event, x->copy_stream));
    sent = upto; unseen = true;
  }
  void need(int64_t upto, cudaStream_t s) {   // the work enqueued on `s` from here on reads rows below `upto`
    if (upto > sent) send(std::max(upto, std::min(total, sent + piece_rows())));
    if (unseen) { CUDA_CHECK(cudaStreamWaitEvent(s, x->copy_event, 0)); unseen = false; }
  }
  void idle_piece() { if (sent < total) send(sent + piece_rows()); }   // called while a long kernel runs
};

// One batch: nodes [n0, n0 + B) against the graph of nodes [0, n0).
void run_batch(hnswb200_index* x, const BuildPlan& bpl, int64_t n0, int64_t B, int64_t n_total, RowUploader* up = nullptr) {
  cudaStream_t s = x->stream;
  if (up) up->need(n0 + B - up->base, s);
  const SearchPlan& pl = bpl.sp;
  // request capacity: one per selected neighbour per layer
  size_t req_cap = 0;
  for (int64_t j = 0; j < B; j++)
    req_cap += (size_t)bpl.sel0 + (size_t)std::min<int>(x->h_level[(size_t)(n0 + j)], x->max_layer) * bpl.selU;
  x->b_req.reserve_geo(req_cap); x->b_req_sorted.reserve_geo(req_cap); x->b_heads.reserve_geo(req_cap);
  CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p, 0, CTR_N * sizeof(unsigned int), s));

  hb::BuildParams p{};
  hb::SearchParams& sp = p.sp;
  sp.g = x->view();
  sp.queries = nullptr; sp.nq = B; sp.ef = x->efC; sp.k = x->efC; sp.ef_cap = pl.ef_cap;
  sp.accept_ties = x->flavour == HNSWB200_FLAVOUR_HNSW_BA; sp.pad_inf = 0;
  sp.hc = pl.hc; sp.q_smem_chunks = pl.q_chunks; sp.smem_per_warp = pl.smem_per_warp; sp.nb_cap = pl.nb_cap;
  sp.stage_slots = pl.stage_slots; sp.stage_ahead = stage_ahead_for(x);
  // a batch that leaves most resident warps idle: a gang of warps per insert (the largest of 8, 4, 2 that
  // divides the CTA and still gives every insert of the batch a gang at once)
  const bool qreg = x->param_build_qreg == 1 || (x->param_build_qreg == 0 && B <= (int64_t)bpl.grid_regs * pl.warps);
  const int grid_cap = qreg ? std::min(bpl.grid_regs, pl.grid) : pl.grid;
  int gang = 1;
  if (!pl.stage_slots && x->param_build_batch != 1 && x->param_gang != 1) {
    const int64_t capacity = (int64_t)grid_cap * pl.warps;
    for (int P = 8; P >= 2; P >>= 1)
      if (pl.warps % P == 0 && B * P <= capacity && (x->param_gang == 0 || P <= x->param_gang)) { gang = P; break; }
  }
  sp.gang = gang;
  sp.out_ids = nullptr; sp.out_dists = nullptr; sp.counters = nullptr; sp.next_query = nullptr;
  sp.bitset_pool = x->d_bitpool.p; sp.pool_busy = x->d_pool_busy.p; sp.pool_size = x->pool_size; sp.words = x->pool_words;
  sp.events = x->d_events.p;
  sp.tie_pool = x->d_tie_pool.p; sp.tie_busy = x->d_tie_busy.p; sp.tie_slots = TIE_SLOTS; sp.tie_cap = TIE_CAP;
  p.adj0 = x->adj0.p; p.adjU = x->adjU.p; p.level = x->level.p; p.row_owner = x->row_owner.p;
  p.n0 = (int)n0; p.B = (int)B;
  p.sel0 = bpl.sel0; p.selU = bpl.selU; p.cap0 = bpl.cap0; p.capU = bpl.capU; p.keep_all = bpl.keep_all;
  p.sel_cap = bpl.sel_cap; p.ucap = bpl.ucap; p.smem_per_warp = pl.smem_per_warp;
  p.req = x->b_req.p; p.req_count = x->b_ctr.p + CTR_REQ;
  p.rem_count = x->b_ctr.p + CTR_REM; p.head_count = x->b_ctr.p + CTR_HEADS; p.next = x->b_ctr.p + CTR_NEXT;
  p.heads = x->b_heads.p; p.counters = x->b_counters.p;
  p.mate_mode = 0;
  // longest inserts first: a node of level l runs l + 1 beam searches, and one that starts in the last wave of a
  // batch holds the whole batch back (results do not depend on the order: every insert sees the snapshot)
  p.order = nullptr;
  if (B > (int64_t)pl.warps * 4 && x->param_build_batch != 1) {
    x->h_order.resize((size_t)B);
    int32_t* o = x->h_order.data();
    size_t start[17] = {0};                            // counting sort by level, highest first, stable
    for (int64_t j = 0; j < B; j++) start[15 - std::min<int>(x->h_level[(size_t)(n0 + j)], 15) + 1]++;
    for (int l = 0; l < 16; l++) start[l + 1] += start[l];
    for (int64_t j = 0; j < B; j++) o[start[15 - std::min<int>(x->h_level[(size_t)(n0 + j)], 15)]++] = (int32_t)j;
    x->b_order.reserve_geo((size_t)B);
    CUDA_CHECK(cudaMemcpyAsync(x->b_order.p, o, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    p.order = x->b_order.p;
  }

  // ---- phase 1
  g_trace.start();
  g_trace.batches++;
  if (B < (int64_t)grid_cap * pl.warps) g_trace.small++;
  const int per_cta = pl.warps / gang;             // inserts a CTA works on at a time
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(grid_cap, (B + per_cta - 1) / per_cta));
  switch (pl.cpl) {
    case 1: launch_build_search<1>(p, pl, grid, qreg, s); break;
    case 2: launch_build_search<2>(p, pl, grid, qreg, s); break;
    case 3: launch_build_search<3>(p, pl, grid, qreg, s); break;
    case 4: launch_build_search<4>(p, pl, grid, qreg, s); break;
    default: launch_build_search<0>(p, pl, grid, true, s); break;
  }
  x->st.gpu_launches += 1;
  if (up && B >= 2048) up->idle_piece();            // the host would only wait for phase 1 now: stage the next piece instead
  unsigned ctr[CTR_N];
  CUDA_CHECK(cudaMemcpyAsync(ctr, x->b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (g_trace.on) {
    const double dt = g_trace.lap();
    g_trace.t_search += dt;
    int bk = 0;
    while ((int64_t(2) << bk) <= B) bk++;
    g_trace.bk_n[bk]++; g_trace.bk_ins[bk] += B; g_trace.bk_t[bk] += dt;
  }
  unsigned n_req = ctr[CTR_REQ];
  if (n_req > req_cap) fail(HNSWB200_ECUDA, "build: request buffer overrun");
  if (n_req == 0) return;

  if (x->param_build_batch == 1) {
    // sequential inserts: the reference's own link order, one warp (build.cuh, build_link_seq_kernel)
    size_t smem = (size_t)hb::link_seq_smem(bpl.ucap, bpl.sel_cap, pl.q_chunks) +
                  (pl.stage_slots ? (size_t)hb::stage_smem_bytes(pl.stage_slots, x->ld / 4) : 0);
    switch (pl.cpl) {
#define HB_SEQ(C)                                                                                                   \
      case C:                                                                                                       \
        CUDA_CHECK(cudaFuncSetAttribute(hb::build_link_seq_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        hb::build_link_seq_kernel<C><<<1, 32, smem, s>>>(p);                                                        \
        break;
      HB_SEQ(1) HB_SEQ(2) HB_SEQ(3) HB_SEQ(4)
      default:
        CUDA_CHECK(cudaFuncSetAttribute(hb::build_link_seq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hb::build_link_seq_kernel<0><<<1, 32, smem, s>>>(p);
#undef HB_SEQ
    }
    CUDA_CHECK(cudaGetLastError());
    x->st.gpu_launches += 1;
    return;
  }

  // ---- mates: links between members of the batch (build.cuh)
  for (int round = 0; round < (int)x->param_build_mates && B > 1 && n_req > 0; round++) {
    sort_keys(x, x->b_req.p, x->b_req_sorted.p, n_req, 32 + hb::REQ_VBITS);
    CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p + CTR_MATE, 0, sizeof(unsigned int), s));
    x->b_mate.reserve_geo((size_t)n_req * hb::MATE_SPAN); x->b_mate_sorted.reserve_geo((size_t)n_req * hb::MATE_SPAN);
    hb::build_mates_kernel<<<(n_req + 255) / 256, 256, 0, s>>>(p, x->b_req_sorted.p, n_req, x->b_mate.p, x->b_ctr.p + CTR_MATE);
    CUDA_CHECK(cudaGetLastError());
    x->st.gpu_launches += 1;
    CUDA_CHECK(cudaMemcpyAsync(ctr, x->b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    const unsigned n_mate = ctr[CTR_MATE];
    if (n_mate > (size_t)n_req * hb::MATE_SPAN) fail(HNSWB200_ECUDA, "build: mate buffer overrun");
    if (n_mate > 0) {
      g_trace.n_mates += n_mate;
      sort_keys(x, x->b_mate.p, x->b_mate_sorted.p, n_mate, 32 + hb::REQ_VBITS);
      CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p + CTR_HEADS, 0, 2 * sizeof(unsigned int), s));
      x->b_heads.reserve_geo(n_mate);
      p.heads = x->b_heads.p;
      hb::segment_heads_kernel<<<(n_mate + 255) / 256, 256, 0, s>>>(x->b_mate_sorted.p, n_mate, hb::REQ_VBITS, x->b_heads.p,
                                                                    x->b_ctr.p + CTR_HEADS);
      CUDA_CHECK(cudaGetLastError());
      hb::BuildParams pm = p;
      pm.req = x->b_mate_sorted.p; pm.n_req = n_mate; pm.mate_mode = 1;
      pm.smem_per_warp = bpl.link_smem_per_warp;
      const size_t msmem = (size_t)bpl.link_warps * bpl.link_smem_per_warp;
      const int mgrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * bpl.link_grid_per_sm, (B * 2 + bpl.link_warps - 1) / bpl.link_warps));
      switch (pl.cpl) {
        case 1: launch_build_link<1>(pm, bpl.link_warps, msmem, mgrid, s); break;
        case 2: launch_build_link<2>(pm, bpl.link_warps, msmem, mgrid, s); break;
        case 3: launch_build_link<3>(pm, bpl.link_warps, msmem, mgrid, s); break;
        case 4: launch_build_link<4>(pm, bpl.link_warps, msmem, mgrid, s); break;
        default: launch_build_link<0>(pm, bpl.link_warps, msmem, mgrid, s); break;
      }
      // the link requests, again, from the final rows
      CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p + CTR_REQ, 0, sizeof(unsigned int), s));
      hb::build_requests_kernel<<<(int)std::min<int64_t>((B + 7) / 8, (int64_t)x->num_sms * 8), 256, 0, s>>>(p);
      CUDA_CHECK(cudaGetLastError());
      x->st.gpu_launches += 3;
      CUDA_CHECK(cudaMemcpyAsync(ctr, x->b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK(cudaStreamSynchronize(s));
      n_req = ctr[CTR_REQ];
      if (n_req > req_cap) fail(HNSWB200_ECUDA, "build: request buffer overrun");
    }
    g_trace.t_mates += g_trace.lap();
  }

  // ---- phase 2: sort by row, one warp per row
  sort_keys(x, x->b_req.p, x->b_req_sorted.p, n_req, 32 + hb::REQ_VBITS);
  CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p + CTR_HEADS, 0, 2 * sizeof(unsigned int), s));   // heads, next
  hb::segment_heads_kernel<<<(n_req + 255) / 256, 256, 0, s>>>(x->b_req_sorted.p, n_req, hb::REQ_VBITS, x->b_heads.p,
                                                               x->b_ctr.p + CTR_HEADS);
  CUDA_CHECK(cudaGetLastError());
  if (g_trace.on) { CUDA_CHECK(cudaStreamSynchronize(s)); g_trace.t_sort += g_trace.lap(); }
  size_t rem_cap = (size_t)n_req * (size_t)(std::max(x->slots0, x->slotsU) + 1);
  rem_cap = std::min<size_t>(rem_cap, (size_t)0xfffffff0u);
  x->b_rem.reserve_geo(rem_cap); x->b_rem_sorted.reserve_geo(rem_cap);
  if (g_trace.on) { CUDA_CHECK(cudaStreamSynchronize(s)); g_trace.t_alloc += g_trace.lap(); }
  p.req = x->b_req_sorted.p; p.n_req = n_req; p.heads = x->b_heads.p;
  p.rem = x->b_rem.p; p.rem_cap = (unsigned)rem_cap;
  p.smem_per_warp = bpl.link_smem_per_warp;
  size_t link_smem = (size_t)bpl.link_warps * bpl.link_smem_per_warp;
  int lgrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * bpl.link_grid_per_sm, (n_req + bpl.link_warps - 1) / bpl.link_warps));
  switch (pl.cpl) {
    case 1: launch_build_link<1>(p, bpl.link_warps, link_smem, lgrid, s); break;
    case 2: launch_build_link<2>(p, bpl.link_warps, link_smem, lgrid, s); break;
    case 3: launch_build_link<3>(p, bpl.link_warps, link_smem, lgrid, s); break;
    case 4: launch_build_link<4>(p, bpl.link_warps, link_smem, lgrid, s); break;
    default: launch_build_link<0>(p, bpl.link_warps, link_smem, lgrid, s); break;
  }
  x->st.gpu_launches += 2;
  CUDA_CHECK(cudaMemcpyAsync(ctr, x->b_ctr.p, sizeof(ctr), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  g_trace.t_link += g_trace.lap();
  unsigned n_rem = ctr[CTR_REM];
  if (n_rem > rem_cap) fail(HNSWB200_ECUDA, "build: removal buffer overrun");
  if (n_rem == 0) return;

  // ---- phase 3: sort removals by row, one warp per row
  sort_keys(x, x->b_rem.p, x->b_rem_sorted.p, n_rem, 32 + hb::REM_ABITS);
  CUDA_CHECK(cudaMemsetAsync(x->b_ctr.p + CTR_HEADS, 0, 2 * sizeof(unsigned int), s));
  x->b_heads.reserve_geo(n_rem);
  p.heads = x->b_heads.p;
  hb::segment_heads_kernel<<<(n_rem + 255) / 256, 256, 0, s>>>(x->b_rem_sorted.p, n_rem, hb::REM_ABITS, x->b_heads.p,
                                                               x->b_ctr.p + CTR_HEADS);
  CUDA_CHECK(cudaGetLastError());
  p.rem = x->b_rem_sorted.p; p.n_req = n_rem;
  int ugrid = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)x->num_sms * 8, (n_rem + 7) / 8));
  hb::build_unlink_kernel<<<ugrid, 256, 0, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
  x->st.gpu_launches += 2;
  g_trace.t_rest += g_trace.lap();
  (void)n_total;
}

// Append n_new vectors: Ohnsw.build_batch_bigarray on an empty index (lib/ohnsw.ml:840-857), a run
// of Ohnsw.insert calls otherwise (:766-837).
void append_nodes(hnswb200_index* x, const float* data, int64_t n_new, const int32_t* levels) {
  auto t0 = std::chrono::steady_clock::now();
  if (n_new < 0) fail(HNSWB200_EINVAL, "build: n < 0");
  if (n_new == 0) return;
  if (!data) fail(HNSWB200_EINVAL, "build: data is NULL");
  if (x->poisoned) fail(HNSWB200_ECUDA, "the index is unusable: an earlier build/insert call failed half-way");
  const int64_t n_old = x->n, n_tot = n_old + n_new;
  if (n_tot >= (int64_t(1) << 31) - 1) fail(HNSWB200_EINVAL, "build: too many nodes");
  cudaStream_t s = x->stream;

  // levels, upper-row bookkeeping (host mirrors are authoritative)
  std::vector<int8_t> lvl_new((size_t)n_new);
  for (int64_t i = 0; i < n_new; i++) {
    // a level the library drew itself is capped at the 16 layers the index holds (P(level > 15) is ~2e-5 per
    // node at M = 2 and nil at practical M); a caller-supplied level outside 0..15 is an argument error
    int l = levels ? levels[i] : std::min(draw_level(x), 15);
    if (n_old + i == 0) l = 0;                       // the first node is the entry of layer 0 (:774-778)
    if (l < 0 || l > 15) fail(HNSWB200_EINVAL, "build: level must be in 0..15");
    lvl_new[(size_t)i] = (int8_t)l;
  }
  const int64_t rows_old = x->rowsU;
  int64_t rows = rows_old;
  std::vector<int32_t> uoff_new((size_t)n_new, -1), owner_new;
  for (int64_t i = 0; i < n_new; i++)
    if (lvl_new[(size_t)i] > 0) {
      uoff_new[(size_t)i] = (int32_t)rows;
      rows += lvl_new[(size_t)i];
      owner_new.insert(owner_new.end(), (size_t)lvl_new[(size_t)i], (int32_t)(n_old + i));
    }
  if (rows >= (int64_t(1) << 31)) fail(HNSWB200_EINVAL, "build: too many upper-layer rows");
  if (x->h_row_owner.size() != (size_t)rows_old) {   // graph came from import_graph: rebuild the owner map
    x->h_row_owner.assign((size_t)rows_old, 0);
    for (int64_t i = 0; i < n_old; i++)
      for (int l = 0; l < x->h_level[(size_t)i]; l++) x->h_row_owner[(size_t)x->h_upper_off[(size_t)i] + l] = (int32_t)i;
    x->row_owner.release();
  }

  // storage: grow geometrically so a run of single inserts stays linear
  int64_t cap = x->cap, capU = x->capU;
  if (n_tot > cap) cap = std::max<int64_t>(n_tot, cap + cap / 2);
  if (rows > capU) capU = std::max<int64_t>(rows, capU + capU / 2);
  const bool fresh = n_old == 0;
  {
    const int64_t old_cap = x->vec.n / std::max(1, x->ld), old_capU = x->adjU.n / std::max(1, x->slotsU);
    x->vec.reserve((size_t)cap * x->ld, !fresh, s);
    x->adj0.reserve((size_t)cap * x->slots0, !fresh, s);
    x->upper_off.reserve((size_t)cap, !fresh, s);
    x->level.reserve((size_t)cap, !fresh, s);
    x->adjU.reserve((size_t)std::max<int64_t>(capU, 1) * x->slotsU, !fresh, s);
    bool owner_realloc = (size_t)std::max<int64_t>(capU, 1) > x->row_owner.n;
    x->row_owner.reserve((size_t)std::max<int64_t>(capU, 1), false, s);
    if (owner_realloc && rows_old > 0)
      CUDA_CHECK(cudaMemcpyAsync(x->row_owner.p, x->h_row_owner.data(), (size_t)rows_old * 4, cudaMemcpyHostToDevice, s));
    x->cap = cap; x->capU = std::max<int64_t>(capU, 1);
    (void)old_cap; (void)old_capU;
  }
  // empty rows for the new nodes (-1 terminated lists)
  CUDA_CHECK(cudaMemsetAsync(x->adj0.p + (size_t)n_old * x->slots0, 0xff, (size_t)n_new * x->slots0 * 4, s));
  if (rows > rows_old)
    CUDA_CHECK(cudaMemsetAsync(x->adjU.p + (size_t)rows_old * x->slotsU, 0xff, (size_t)(rows - rows_old) * x->slotsU * 4, s));
  RowUploader up;
  up.x = x; up.src = data; up.base = n_old; up.total = n_new;
  if (!x->copy_stream) {
    CUDA_CHECK(cudaStreamCreateWithFlags(&x->copy_stream, cudaStreamNonBlocking));
    CUDA_CHECK(cudaEventCreateWithFlags(&x->copy_event, cudaEventDisableTiming));
  }
  // the copy stream must not write into storage that `s` is still moving (reserve(keep) copies on s)
  CUDA_CHECK(cudaEventRecord(x->copy_event, s));
  CUDA_CHECK(cudaStreamWaitEvent(x->copy_stream, x->copy_event, 0));
  up.send(up.piece_rows());
  CUDA_CHECK(cudaMemcpyAsync(x->upper_off.p + n_old, uoff_new.data(), (size_t)n_new * 4, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaMemcpyAsync(x->level.p + n_old, lvl_new.data(), (size_t)n_new, cudaMemcpyHostToDevice, s));
  if (!owner_new.empty())
    CUDA_CHECK(cudaMemcpyAsync(x->row_owner.p + rows_old, owner_new.data(), owner_new.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  x->h_level.insert(x->h_level.end(), lvl_new.begin(), lvl_new.end());
  x->h_upper_off.insert(x->h_upper_off.end(), uoff_new.begin(), uoff_new.end());
  x->h_row_owner.insert(x->h_row_owner.end(), owner_new.begin(), owner_new.end());
  x->rowsU = rows;

  const double t_setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  // scratch
  BuildPlan bpl = plan_build(x, n_tot);
  const double t_plan_only = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - t_setup;
  x->b_ctr.reserve(CTR_N);
  x->b_counters.reserve(4);
  x->d_events.reserve(4);
  CUDA_CHECK(cudaMemsetAsync(x->b_counters.p, 0, 4 * sizeof(unsigned long long), s));
  CUDA_CHECK(cudaMemsetAsync(x->d_events.p, 0, 4 * sizeof(unsigned long long), s));
  ensure_pool(x, bpl.sp.grid * bpl.sp.warps, n_tot, s);
  if (bpl.sp.hash_slots == 0) bpl.sp.grid = std::max(1, std::min(bpl.sp.grid, x->pool_size / bpl.sp.warps));   // one set per warp

  // batch schedule
  int64_t max_batch = x->param_build_batch > 0 ? x->param_build_batch : 16384;
  max_batch = std::min<int64_t>(max_batch, (int64_t(1) << hb::REQ_VBITS) - 1);
  const int64_t ratio = std::max<int64_t>(1, x->param_build_ratio);
  const int64_t ratio_early = std::max<int64_t>(1, std::min<int64_t>(ratio, x->param_build_ratio_early));
  {
    // scratch for the largest batch this call will run, allocated once (growing it batch by
    // batch costs more in cudaMalloc / cudaFree than the kernels it feeds)
    int64_t bmax = std::max<int64_t>(1, std::min<int64_t>(max_batch, std::min<int64_t>(n_new, n_tot / ratio + 1)));
    size_t req_max = (size_t)bmax * (size_t)(bpl.sel0 + 2 * bpl.selU);
    size_t rem_max = std::min<size_t>(req_max * (size_t)(std::max(x->slots0, x->slotsU) + 1), (size_t)0xfffffff0u);
    x->b_req.reserve_geo(req_max); x->b_req_sorted.reserve_geo(req_max); x->b_heads.reserve_geo(req_max);
    x->b_rem.reserve_geo(rem_max); x->b_rem_sorted.reserve_geo(rem_max);
    if (x->param_build_mates != 0) { x->b_mate.reserve_geo(req_max * hb::MATE_SPAN); x->b_mate_sorted.reserve_geo(req_max * hb::MATE_SPAN); }
    size_t bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortKeys(nullptr, bytes, x->b_rem.p, x->b_rem_sorted.p, (int)std::min<size_t>(rem_max, 0x7fffffff), 0, 64, s));
    x->b_cub.reserve_geo(bytes + 16);
  }
  const double t_plan = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - t_setup;
  int64_t done = n_old;
  if (done == 0) { x->entry = 0; x->max_layer = 0; done = 1; x->n = 1; }     // :774-778
  // if a batch fails the bookkeeping is cut back to the nodes linked so far, but the failed batch may
  // already have prepended its nodes to existing rows: the index is marked unusable (search / insert
  // refuse it) rather than left answering with ids that do not exist
  auto rollback = [&]() {
    x->poisoned = true;
    x->h_level.resize((size_t)x->n);
    x->h_upper_off.resize((size_t)x->n);
    int64_t rows_kept = 0;
    for (int64_t i = 0; i < x->n; i++) rows_kept += x->h_level[(size_t)i];
    x->h_row_owner.resize((size_t)rows_kept);
    x->rowsU = rows_kept;
    x->layer_stats_dirty = true;
  };
  try {
  while (done < n_tot) {
    // a batch is at most 1/ratio of the FINAL graph (what bounds the links its members miss by not seeing each
    // other) and at most 1/ratio_early of the graph so far: the rows of the early nodes are re-selected many
    // times as the graph grows, so the coarser early batches leave no trace in the finished index, while they
    // cut the number of small, latency-bound batches from 64 ln(n / 64) to 4 ln(n / 256) + 60
    int64_t B = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(max_batch, n_tot / ratio), done / ratio_early));
    B = std::min<int64_t>(B, n_tot - done);
    // a node that raises max_layer becomes the entry point (:832-836) and must be seen by
    // every later insert: it closes its batch
    for (int64_t j = 0; j < B; j++)
      if (x->h_level[(size_t)(done + j)] > x->max_layer) { B = j + 1; break; }
    run_batch(x, bpl, done, B, n_tot, &up);
    const int64_t last = done + B - 1;
    if (x->h_level[(size_t)last] > x->max_layer) { x->max_layer = x->h_level[(size_t)last]; x->entry = last; }
    done += B;
    x->n = done;
  }
  up.need(n_new, s);
  CUDA_CHECK(cudaStreamSynchronize(s));
  } catch (...) { cudaStreamSynchronize(x->copy_stream); rollback(); throw; }
  unsigned long long c[4];
  CUDA_CHECK(cudaMemcpy(c, x->b_counters.p, sizeof(c), cudaMemcpyDeviceToHost));
  unsigned long long evs[2];
  CUDA_CHECK(cudaMemcpy(evs, x->d_events.p, sizeof(evs), cudaMemcpyDeviceToHost));
  if (c[3]) fail(HNSWB200_ECUDA, "build: removal buffer overflow");
  if (g_trace.on) {
    fprintf(stderr, "[hnsw_b200 build] storage + upload %.3fs, plan (occupancy queries: first use loads the kernels) %.3fs, scratch + visited pool %.3fs\n",
            t_setup, t_plan_only, t_plan - t_plan_only);
    fprintf(stderr, "[hnsw_b200 build] n=%lld batches=%lld (under one wave: %lld) search %.3fs mates %.3fs (%llu proposals) sort+heads %.3fs alloc %.3fs link %.3fs unlink-enqueue %.3fs dropped_incoming=%llu grid=%d x %d warps hash_slots=%d\n",
            (long long)n_new, (long long)g_trace.batches, (long long)g_trace.small, g_trace.t_search, g_trace.t_mates, (unsigned long long)g_trace.n_mates, g_trace.t_sort, g_trace.t_alloc, g_trace.t_link, g_trace.t_rest,
            c[2], bpl.sp.grid, bpl.sp.warps, bpl.sp.hash_slots);
    for (int b = 0; b < 32; b++)
      if (g_trace.bk_n[b])
        fprintf(stderr, "[hnsw_b200 build]   B in [%d, %d): %lld batches, %lld inserts, phase 1 %.3fs (%.3f ms per batch, %.2f us per insert)\n",
                1 << b, 2 << b, (long long)g_trace.bk_n[b], (long long)g_trace.bk_ins[b], g_trace.bk_t[b],
                1e3 * g_trace.bk_t[b] / g_trace.bk_n[b], 1e6 * g_trace.bk_t[b] / g_trace.bk_ins[b]);
    g_trace = BuildTrace();
  }
  x->st.build_inserts = (uint64_t)n_new;
  x->st.build_n_dist = c[0];
  x->st.build_n_exp = c[1];
  x->st.build_visited_overflows = evs[0];
  x->st.build_dropped_incoming = c[2];
  x->st.build_algorithmic_bytes = (double)c[0] * 4.0 * x->dim + (double)c[1] * 4.0 * x->slots0;
  x->st.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  x->last_nq = 0;
  x->layer_stats_dirty = true;
}

}  // namespace

extern "C" {

int hnswb200_build(hnswb200_index* x, const float* data, int64_t n, const int32_t* levels) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    if (x->n != 0) fail(HNSWB200_EINVAL, "build: the index is not empty (use insert)");
    x->slots0 = 2 * x->M; x->slotsU = x->M;
    append_nodes(x, data, n, levels);
  });
}

int hnswb200_insert(hnswb200_index* x, const float* data, int64_t n, const int32_t* levels) {
  return guard([&] {
    if (!x) fail(HNSWB200_EINVAL, "index is NULL");
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    append_nodes(x, data, n, levels);
  });
}

}  // extern "C"
