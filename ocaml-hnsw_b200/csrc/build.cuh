// GPU index construction: batched restatement of Ohnsw.insert (lib/ohnsw.ml:766-837).
//
// The reference inserts one node at a time; every insert sees every earlier one.  Here nodes
// are inserted in BATCHES against a snapshot of the graph (the host grows the batch with the
// graph, see api_build.inl), in three deterministic phases per batch:
//
//   1. build_search_kernel   one warp per new node: greedy descent (search_one_simple,
//      :492-508), then for layer = min(level, max_layer) .. 0 the efC beam search seeded with
//      the whole result set of the layer above (search_k, :543-588, :806-816), the selection
//      heuristic (select_neighbours, :647-663) and the new node's own adjacency rows
//      (set_connections_for_new_node, :198-202).  Every selected (row, new node) pair is
//      emitted as a 64-bit request.
//   2. requests are radix-sorted by row; build_link_kernel, one warp per distinct row, prepends
//      the incoming new nodes in id order (what the sequential loop would have done) and, when
//      the row outgrows its bound (2M on layer 0, M above; :818-823), re-selects it with the
//      same heuristic keyed on distances to the row's owner (:791-798, :824-826).  Entries that
//      fall out are emitted as removal requests (the symmetric half of Graph.set_connections,
//      :182-196).
//   3. removal requests are sorted by row; build_unlink_kernel, one warp per distinct row,
//      drops the owners that pruned it.
//
// Batch members do not see each other in phase 1 (they search the snapshot).  The links the sequential loop
// would have made between them are recovered by a "mates" pass between phases 1 and 2 (build_mates_kernel):
// two new nodes that selected the same neighbour on a layer are candidates for each other; every new node
// re-runs the selection heuristic over its own list plus its EARLIER mates (the later node is the one that
// would have found the earlier one, :806-819), and the link requests are then regenerated from the final
// rows, so that the earlier mate receives the reverse link like any other selected neighbour (:820).
//
// Links stay symmetric after every batch (Graph.Test.invariant, :217-225).  With a batch of
// one node and the sequential link kernel the phases collapse to the reference's order.
#pragma once
#include "search.cuh"

namespace hb {

constexpr int REQ_VBITS = 24;            // link request  = row id << 24 | index of the new node in its batch
constexpr int REM_ABITS = 31;            // unlink request = row id << 31 | owner that dropped the row's node
constexpr int LINK_MCAP = 96;            // incoming new nodes one row considers per batch
constexpr int MATE_SPAN = 8;             // earlier batch members proposed to a new node per shared neighbour
constexpr uint32_t ROW_UPPER = 0x80000000u;   // row id: node id (layer 0) or ROW_UPPER | upper row index

struct BuildParams {
  SearchParams sp;          // g = the snapshot (n = nodes already linked), ef = efC
  int32_t* adj0;            // writable aliases of g.adj0 / g.adjU
  int32_t* adjU;
  const int8_t* level;      // [n_total]
  const int32_t* row_owner; // [rowsU] node that owns each upper row
  int n0, B;                // this batch = nodes [n0, n0 + B)
  const int32_t* order;     // phase 1 takes the batch in this order (multi-layer nodes, the long inserts, first); null = as numbered
  int sel0, selU;           // neighbours selected for a new node on layer 0 / above (:818)
  int cap0, capU;           // degree bound of a row before it is re-selected (:822-823)
  int keep_all;             // Hnsw.Ba shortcut: #candidates <= n keeps all (hnsw_algo.ml:596-599)
  int sel_cap;              // uint32 slots reserved per warp for the selected list
  int ucap;                 // link kernel: max union size per row
  int smem_per_warp;
  uint64_t* req;            // phase 1 out / phase 2 in (sorted)
  unsigned int* req_count;
  uint64_t* rem;            // phase 2 out / phase 3 in (sorted)
  unsigned int* rem_count;
  unsigned int rem_cap;
  const unsigned int* heads;      // segment starts in the sorted request array
  const unsigned int* head_count;
  unsigned int n_req;             // number of sorted requests (phase 2 / 3)
  unsigned int* next;             // work counter
  int mate_mode;                  // link kernel: the rows are the new nodes' own, the incoming nodes their earlier mates
  unsigned long long* counters;   // [0] distance evaluations, [1] adjacency rows read, [2] dropped incoming, [3] rem overflow
};

__device__ __forceinline__ int32_t* row_ptr(const BuildParams& bp, uint32_t rid) {
  return (rid & ROW_UPPER) ? bp.adjU + (size_t)(rid & ~ROW_UPPER) * bp.sp.g.slotsU
                           : bp.adj0 + (size_t)rid * bp.sp.g.slots0;
}
__device__ __forceinline__ uint32_t row_id(const GraphView& g, uint32_t node, int layer) {
  return layer == 0 ? node : (ROW_UPPER | (uint32_t)(g.upper_off[node] + layer - 1));
}

// select_neighbours (lib/ohnsw.ml:647-663): candidates ascending by (distance to the base,
// id); e is kept iff it is strictly closer to the base than to every node already kept.
// Kept candidates get bit 0 of their key set and are appended to sel[] (oldest first).
// `qe`/`qs2` receive the candidate's vector.  The kept list is scanned oldest first, eight
// nodes per round, stopping at the first round that rejects (the reference's for_all scans
// newest first and stops at the first failure: same verdict, different distance count).
// QREG = false: the candidate's vector is kept in shared memory only (qs2, zero padded to `q_chunks`), as the
// search kernel keeps its query — fewer live registers, more resident warps.
template <int CPL, bool QREG = true>
__device__ __forceinline__ int select_neighbours(const GraphView& g, uint64_t* cand, int ncand, int want, int keep_all,
                                                 uint32_t* sel, float4* qe, float4* qs2, float* newd, int lane,
                                                 uint32_t& n_dist, Stage* st = nullptr, int q_chunks = 0) {
  int nsel = 0;
  if (want <= 0) return 0;
  if (keep_all && ncand <= want) {
    for (int i = lane; i < ncand; i += 32) { sel[i] = key_id(cand[i]); cand[i] |= 1ull; }
    __syncwarp();
    return ncand;
  }
  for (int i = 0; i < ncand; i++) {
    const uint64_t key = cand[i];
    const uint32_t e = key_id(key);
    const float de = key_dist(key);
    bool ok = true;
    if (nsel > 0) {
      if (QREG || CPL == 0) load_target<CPL>(g, reinterpret_cast<const float4*>(g.vec) + (size_t)e * g.ld4, qe, qs2, lane);
      else load_target_smem(g, reinterpret_cast<const float4*>(g.vec) + (size_t)e * g.ld4, qs2, q_chunks, lane);
      for (int base = 0; base < nsel; base += 8) {
        int cnt = min(8, nsel - base);
        batch_dist<CPL>(g, QREG ? qe : nullptr, qs2, sel + base, newd, cnt, lane, st);
        n_dist += cnt;
        bool bad = lane < cnt && !(de < newd[lane]);
        unsigned any_bad = __ballot_sync(FULL, bad);
        __syncwarp();
        if (any_bad) { ok = false; break; }
      }
    }
    if (ok) {
      if (lane == 0) { sel[nsel] = e; cand[i] = key | 1ull; }
      nsel++;
      __syncwarp();
      if (nsel >= want) break;
    }
  }
  return nsel;
}

// ---- phase 1 --------------------------------------------------------------------------------------
template <int CPL, bool GANG = false, bool QREG = true>
__global__ void __launch_bounds__(256, QREG ? 2 : 3) build_search_kernel(const BuildParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const SearchParams& p = bp.sp;
  const GraphView& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // batches smaller than the resident warps run a gang of p.gang warps per insert (search.cuh, Gang): the
  // distance rounds of every expansion are shared, the insert's latency — which is the batch's — drops
  unsigned char* my = smem_raw + (size_t)(GANG ? warp / p.gang : warp) * bp.smem_per_warp;
  WarpCtx<CPL, QREG> w;
  w.lane = lane;
  w.gang.P = GANG ? p.gang : 1; w.gang.rank = GANG ? warp % p.gang : 0; w.gang.bar = GANG ? 1 + warp / p.gang : 1;
  w.gang.job = reinterpret_cast<GangJob*>(my + bp.smem_per_warp - (int)sizeof(GangJob));
  if (GANG && w.gang.rank > 0) { gang_help<CPL>(g, w.gang, lane); return; }
  w.keys = reinterpret_cast<uint64_t*>(my);
  w.ties = w.keys + p.ef_cap;
  w.newid = reinterpret_cast<uint32_t*>(w.ties + TIES_CAP);
  w.newd = reinterpret_cast<float*>(w.newid + p.nb_cap);
  w.qs = reinterpret_cast<float4*>(w.newd + p.nb_cap);
  visited_init(w.vis, p, reinterpret_cast<uint32_t*>(w.qs + p.q_smem_chunks));
  float4* qs2 = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(w.vis.tab) + p.hc.bytes);
  uint32_t* sel = reinterpret_cast<uint32_t*>(qs2 + p.q_smem_chunks);
  stage_attach(w.st, reinterpret_cast<unsigned char*>(sel + bp.sel_cap), p.stage_slots, p.stage_ahead, g.ld4, lane);
  w.tie_spill = nullptr; w.tie_slot = -1;
  float4 qe[(CPL > 0 && QREG) ? CPL : 1];

  unsigned long long tot_dist = 0, tot_exp = 0;
  while (true) {
    unsigned b = 0;
    if (lane == 0) b = atomicAdd(bp.next, 1u);
    b = __shfl_sync(FULL, b, 0);
    if (b >= (unsigned)bp.B) break;
    if (bp.order) b = (unsigned)bp.order[b];
    const uint32_t v = (uint32_t)bp.n0 + b;
    if (QREG || CPL == 0) load_target<CPL>(g, reinterpret_cast<const float4*>(g.vec) + (size_t)v * g.ld4, w.q, w.qs, lane);
    if ((GANG || !QREG) && CPL > 0)                 // the gang (and the register-lean build) reads the target from shared memory
      load_target_smem(g, reinterpret_cast<const float4*>(g.vec) + (size_t)v * g.ld4, w.qs, p.q_smem_chunks, lane);
    const int lv = bp.level[v];
    uint32_t n_dist = 0, n_exp0 = 0, n_expU = 0;
    bool tie_overflow = false;

    // :783-789 entry point, greedy descent through the layers above the node's level
    uint32_t cur = (uint32_t)g.entry;
    if (lane == 0) w.newid[0] = cur;
    __syncwarp();
    batch_dist<CPL>(g, w.target_regs(), w.qs, w.newid, w.newd, 1, lane, &w.st);
    float d_cur = w.newd[0];
    __syncwarp();
    for (int layer = g.max_layer; layer > lv; layer--) greedy_layer(g, w, layer, cur, d_cur, n_dist, n_expU);

    // :801-802 w_queue = {node}
    n_dist++;
    if (lane == 0) w.keys[0] = make_key(d_cur, cur);
    int n = 1;
    __syncwarp();
    for (int layer = min(lv, g.max_layer); layer >= 0; layer--) {       // :806
      // search_k (:811): the beam is seeded with every element of w_queue, all unexpanded,
      // all marked visited (:555-557)
      visited_clear(w.vis, p.hc, lane);
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        uint64_t k = 0;
        if (i < n) { k = w.keys[i] & ~1ull; w.keys[i] = k; }
        visited_test_and_set(w.vis, p, i < n, key_id(k), lane);
      }
      w.vis.count = n;
      __syncwarp();
      layer_search<CPL, QREG, GANG>(p, w, layer, n, n_dist, layer == 0 ? n_exp0 : n_expU, tie_overflow);
      visited_release(w.vis, p, lane);
      __syncwarp();
      // select_neighbours (MinQueue.copy w_queue) nc (:818-819)
      const int want = layer == 0 ? bp.sel0 : bp.selU;
      int nsel = select_neighbours<CPL, QREG>(g, w.keys, n, want, bp.keep_all, sel, qe, qs2, w.newd, lane, n_dist, &w.st, p.q_smem_chunks);
      // set_connections_for_new_node (:820): Neighbours.add prepends, so the row head is the
      // last node selected; the reverse half is deferred to the link phase
      int32_t* row = row_ptr(bp, row_id(g, v, layer));
      const int slots = layer == 0 ? g.slots0 : g.slotsU;
      for (int j = lane; j < slots; j += 32) row[j] = j < nsel ? (int32_t)sel[nsel - 1 - j] : -1;
      unsigned base = 0;
      if (lane == 0 && nsel) base = atomicAdd(bp.req_count, (unsigned)nsel);
      base = __shfl_sync(FULL, base, 0);
      for (int j = lane; j < nsel; j += 32)
        bp.req[base + j] = ((uint64_t)row_id(g, sel[j], layer) << REQ_VBITS) | (uint64_t)b;
      __syncwarp();
    }
    tot_dist += n_dist;
    tot_exp += n_exp0 + n_expU;
    if (tie_overflow && lane == 0) atomicAdd(p.events + 3, 1ull);
  }
  if (GANG) gang_dismiss(w.gang, lane);
  if (lane == 0) {
    atomicAdd(bp.counters + 0, tot_dist);
    atomicAdd(bp.counters + 1, tot_exp);
  }
}

// ---- segment heads of a sorted request array ----------------------------------------------------
__global__ void segment_heads_kernel(const uint64_t* keys, unsigned n, int shift, unsigned int* heads,
                                     unsigned int* head_count) {
  unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  bool head = i < n && (i == 0 || (keys[i] >> shift) != (keys[i - 1] >> shift));
  unsigned m = __ballot_sync(FULL, head);
  if (!m) return;
  int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(head_count, (unsigned)__popc(m));
  base = __shfl_sync(FULL, base, __ffs(m) - 1);
  if (head) heads[base + __popc(m & ((1u << lane) - 1u))] = i;
}

// ---- batch mates -------------------------------------------------------------------------------------
// `sorted` = phase 1's requests ordered by (row, batch index).  Request i = (row of w, v_i): every one of
// the up to MATE_SPAN requests before it in the same segment names an earlier batch member v_j that
// selected the same w on the same layer; emit (row of v_i on that layer, v_j).  Duplicates (two shared
// neighbours) are dropped by the link kernel after the sort.
__global__ void build_mates_kernel(const BuildParams bp, const uint64_t* sorted, unsigned n_req, uint64_t* mates,
                                   unsigned int* mate_count) {
  const GraphView& g = bp.sp.g;
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const uint64_t vmask = (1ull << REQ_VBITS) - 1;
  uint64_t out[MATE_SPAN];
  int k = 0;
  if (i < n_req) {
    const uint64_t key = sorted[i];
    const uint32_t rid = (uint32_t)(key >> REQ_VBITS);
    int layer = 0;
    if (rid & ROW_UPPER) { const uint32_t r = rid & ~ROW_UPPER; layer = (int)r - g.upper_off[bp.row_owner[r]] + 1; }
    uint64_t mine = 0;
    for (unsigned t = 1; t <= (unsigned)MATE_SPAN && t <= i; t++) {
      const uint64_t kj = sorted[i - t];
      if ((uint32_t)(kj >> REQ_VBITS) != rid) break;
      if (k == 0) mine = (uint64_t)row_id(g, (uint32_t)bp.n0 + (uint32_t)(key & vmask), layer) << REQ_VBITS;
      out[k++] = mine | (kj & vmask);
    }
  }
  int incl = k;
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
  const int total = __shfl_sync(FULL, incl, 31);
  if (!total) return;
  unsigned base = 0;
  if (lane == 31) base = atomicAdd(mate_count, (unsigned)total);
  base = __shfl_sync(FULL, base, 31) + (unsigned)(incl - k);
#pragma unroll
  for (int t = 0; t < MATE_SPAN; t++) if (t < k) mates[base + t] = out[t];
}

// Link requests from the final rows of the batch's new nodes (one warp per node): (row of x on the layer, v)
// for every x in v's list.
__global__ void __launch_bounds__(256) build_requests_kernel(const BuildParams bp) {
  const GraphView& g = bp.sp.g;
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < bp.B; b += warps) {
    const uint32_t v = (uint32_t)bp.n0 + (uint32_t)b;
    for (int layer = min((int)bp.level[v], g.max_layer); layer >= 0; layer--) {
      const int32_t* row = row_ptr(bp, row_id(g, v, layer));
      const int slots = layer == 0 ? g.slots0 : g.slotsU;
      for (int r0 = 0; r0 < slots; r0 += 32) {
        const int nb = r0 + lane < slots ? row[r0 + lane] : -1;
        const unsigned m = __ballot_sync(FULL, nb >= 0);
        if (!m) break;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(bp.req_count, (unsigned)__popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (nb >= 0) bp.req[base + __popc(m & ((1u << lane) - 1u))] = ((uint64_t)row_id(g, (uint32_t)nb, layer) << REQ_VBITS) | (uint64_t)b;
        if (m != FULL) break;
      }
    }
  }
}

// Rank sort of n distinct keys (shared memory, one warp).
__device__ __forceinline__ void warp_rank_sort(const uint64_t* in, uint64_t* out, int n, int lane) {
  for (int i = lane; i < n; i += 32) {
    uint64_t k = in[i];
    int r = 0;
    for (int j = 0; j < n; j++) r += in[j] < k;
    out[r] = k;
  }
  __syncwarp();
}

// ---- phase 2 --------------------------------------------------------------------------------------
// Per-warp shared memory: ukey[ucap] u64, sorted[ucap] u64, uid[ucap] u32, ud[ucap] f32,
// sel[sel_cap] u32, newd[32] f32, qs / qs2.
__host__ __device__ inline int link_smem_per_warp(int ucap, int sel_cap, int q_chunks) {
  return ucap * 8 * 2 + ucap * 4 * 2 + sel_cap * 4 + 32 * 4 + 2 * q_chunks * 16;      // (+ the bulk-copy ring, when used)
}

template <int CPL>
__global__ void __launch_bounds__(256) build_link_kernel(const BuildParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GraphView& g = bp.sp.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* my = smem_raw + (size_t)warp * bp.smem_per_warp;
  uint64_t* ukey = reinterpret_cast<uint64_t*>(my);
  uint64_t* sorted = ukey + bp.ucap;
  float4* qs = reinterpret_cast<float4*>(sorted + bp.ucap);
  float4* qs2 = qs + bp.sp.q_smem_chunks;
  uint32_t* uid = reinterpret_cast<uint32_t*>(qs2 + bp.sp.q_smem_chunks);
  float* ud = reinterpret_cast<float*>(uid + bp.ucap);
  uint32_t* sel = reinterpret_cast<uint32_t*>(ud + bp.ucap);
  float* newd = reinterpret_cast<float*>(sel + bp.sel_cap);
  Stage st;
  stage_attach(st, reinterpret_cast<unsigned char*>(newd + 32), bp.sp.stage_slots, bp.sp.stage_ahead, g.ld4, lane);
  float4 qa[CPL > 0 ? CPL : 1], qe[CPL > 0 ? CPL : 1];
  const unsigned nheads = *bp.head_count;
  unsigned long long tot_dist = 0, tot_rows = 0, tot_dropped = 0;

  while (true) {
    unsigned s = 0;
    if (lane == 0) s = atomicAdd(bp.next, 1u);
    s = __shfl_sync(FULL, s, 0);
    if (s >= nheads) break;
    const unsigned h = bp.heads[s];
    const uint64_t k0 = bp.req[h];
    const uint32_t rid = (uint32_t)(k0 >> REQ_VBITS);
    // incoming new nodes of this row, ascending id (the order the sequential loop adds them)
    int m = 0;
    uint32_t* inc = reinterpret_cast<uint32_t*>(sorted);       // mate mode: the distinct incoming ids, ascending
    for (unsigned i = h;; i += 32) {
      unsigned idx = i + lane;
      const uint64_t key = idx < bp.n_req ? bp.req[idx] : 0ull;
      bool in = idx < bp.n_req && (uint32_t)(key >> REQ_VBITS) == rid;
      unsigned bm = __ballot_sync(FULL, in);
      if (bp.mate_mode) {                                      // the same mate can be proposed through several shared neighbours
        const bool fresh = in && (idx == h || bp.req[idx - 1] != key);
        const unsigned fm = __ballot_sync(FULL, fresh);
        const int pos = m + __popc(fm & ((1u << lane) - 1u));
        if (fresh && pos < LINK_MCAP) inc[pos] = (uint32_t)bp.n0 + (uint32_t)(key & ((1ull << REQ_VBITS) - 1));
        m += __popc(fm);
      } else m += __popc(bm);
      if (bm != FULL) break;
    }
    __syncwarp();
    int32_t* row = row_ptr(bp, rid);
    const bool upper = (rid & ROW_UPPER) != 0;
    const int slots = upper ? g.slotsU : g.slots0;
    const int nc = bp.mate_mode ? (upper ? bp.selU : bp.sel0) : (upper ? bp.capU : bp.cap0);
    uint32_t a;                                       // the row's owner
    int layer;
    if (upper) { uint32_t r = rid & ~ROW_UPPER; a = (uint32_t)bp.row_owner[r]; layer = (int)r - g.upper_off[a] + 1; }
    else { a = rid; layer = 0; }
    tot_rows++;
    // union, list order: newest incoming first, then the old row (:116-118 prepend)
    int dropped_inc = 0;
    int mm = m;
    if (mm > LINK_MCAP) { dropped_inc = bp.mate_mode ? 0 : mm - LINK_MCAP; mm = LINK_MCAP; }   // keeps the LINK_MCAP smallest ids
    for (int j = lane; j < mm; j += 32)
      uid[mm - 1 - j] = bp.mate_mode ? inc[j] : (uint32_t)bp.n0 + (uint32_t)(bp.req[h + j] & ((1ull << REQ_VBITS) - 1));
    int deg = 0;
    for (int r0 = 0; r0 < slots; r0 += 32) {
      int nb = r0 + lane < slots ? row[r0 + lane] : -1;
      unsigned bm = __ballot_sync(FULL, nb >= 0);
      if (bp.mate_mode) {
        // a later round: the row may already hold a proposed mate (a member of this batch) — it stays in its place once
        bool keep = nb >= 0;
        if (keep && nb >= bp.n0)
          for (int j = 0; j < mm; j++) if (inc[j] == (uint32_t)nb) { keep = false; break; }
        const unsigned km = __ballot_sync(FULL, keep);
        if (keep) uid[mm + deg + __popc(km & ((1u << lane) - 1u))] = (uint32_t)nb;
        deg += __popc(km);
      } else {
        if (nb >= 0) uid[mm + r0 + lane] = (uint32_t)nb;
        deg += __popc(bm);
      }
      if (bm != FULL) break;
    }
    __syncwarp();
    const int u = mm + deg;
    if (!bp.mate_mode && u <= nc && dropped_inc == 0) {
      for (int j = lane; j < u; j += 32) row[j] = (int32_t)uid[j];
      __syncwarp();
      continue;
    }
    // min_queue_of_neighbours (:791-798): distances from the owner to every member
    load_target<CPL>(g, reinterpret_cast<const float4*>(g.vec) + (size_t)a * g.ld4, qa, qs, lane);
    batch_dist<CPL>(g, qa, qs, uid, ud, u, lane, &st);
    uint32_t n_dist = (uint32_t)u;
    for (int j = lane; j < u; j += 32) ukey[j] = make_key(ud[j], uid[j]);
    __syncwarp();
    warp_rank_sort(ukey, sorted, u, lane);
    int nsel = select_neighbours<CPL>(g, sorted, u, nc, bp.mate_mode ? bp.keep_all : 0, sel, qe, qs2, newd, lane, n_dist, &st);   // :824-826
    // Graph.set_connections (:182-196): the row becomes the selected list (head = last kept) ...
    for (int j = lane; j < slots; j += 32) row[j] = j < nsel ? (int32_t)sel[nsel - 1 - j] : -1;
    if (bp.mate_mode) {            // nothing is linked to this row yet: the requests are regenerated from the rows
      tot_dist += n_dist;
      __syncwarp();
      continue;
    }
    // ... and every member that fell out loses its link to the owner
    const int nrem = u - nsel + dropped_inc;
    unsigned base = 0;
    if (lane == 0 && nrem) base = atomicAdd(bp.rem_count, (unsigned)nrem);
    base = __shfl_sync(FULL, base, 0);
    if (nrem && base + (unsigned)nrem > bp.rem_cap) {
      if (lane == 0) atomicAdd(bp.counters + 3, 1ull);
    } else if (nrem) {
      unsigned off = base;
      for (int j0 = 0; j0 < u; j0 += 32) {
        int j = j0 + lane;
        bool out = j < u && !(sorted[j] & 1ull);
        unsigned bm = __ballot_sync(FULL, out);
        if (out) {
          uint32_t x = key_id(sorted[j]);
          bp.rem[off + __popc(bm & ((1u << lane) - 1u))] = ((uint64_t)row_id(g, x, layer) << REM_ABITS) | (uint64_t)a;
        }
        off += __popc(bm);
      }
      for (int j = lane; j < dropped_inc; j += 32) {
        uint32_t x = (uint32_t)bp.n0 + (uint32_t)(bp.req[h + LINK_MCAP + j] & ((1ull << REQ_VBITS) - 1));
        bp.rem[off + j] = ((uint64_t)row_id(g, x, layer) << REM_ABITS) | (uint64_t)a;
      }
    }
    tot_dist += n_dist;
    tot_dropped += dropped_inc;
    __syncwarp();
  }
  if (lane == 0) {
    atomicAdd(bp.counters + 0, tot_dist);
    atomicAdd(bp.counters + 1, tot_rows);
    if (tot_dropped) atomicAdd(bp.counters + 2, tot_dropped);
  }
}

// ---- sequential link (build_batch = 1) -----------------------------------------------------------
// One new node, one warp, the reference's own order (lib/ohnsw.ml:820-829): prepend the node to
// every selected neighbour (set_connections_for_new_node, :198-202), then walk the neighbours in
// list order and re-select every list that outgrew its bound, applying Graph.set_connections
// (:182-196) at once — dropped members lose the pruned node immediately, their lists rebuilt in
// reverse as Neighbours.remove does (:119-124).  A full row that receives the node holds one
// entry more than its width until its turn comes: that entry is kept aside (`pend`).  With this
// kernel a GPU build reproduces the oracle's graph edge for edge (tests/test_build_parity.py).
__host__ __device__ inline int link_seq_smem(int ucap, int sel_cap, int q_chunks) {
  return link_smem_per_warp(ucap, sel_cap, q_chunks) + 3 * sel_cap * 4;
}

// list_r := reverse(filter(!= a, logical list of r)), where the logical list is [v] + row when pending
__device__ __forceinline__ void seq_remove(int32_t* row, int slots, bool pending, uint32_t v, uint32_t a, uint32_t* tmp, int lane) {
  int n = 0;
  if (pending) { if (lane == 0 && v != a) tmp[0] = v; n = v != a ? 1 : 0; }
  for (int r0 = 0; r0 < slots; r0 += 32) {
    int nb = r0 + lane < slots ? row[r0 + lane] : -1;
    unsigned valid = __ballot_sync(FULL, nb >= 0);
    bool keep = nb >= 0 && (uint32_t)nb != a;
    unsigned km = __ballot_sync(FULL, keep);
    if (keep) tmp[n + __popc(km & ((1u << lane) - 1u))] = (uint32_t)nb;
    n += __popc(km);
    if (valid != FULL) break;
  }
  __syncwarp();
  for (int j = lane; j < slots; j += 32) row[j] = j < n ? (int32_t)tmp[n - 1 - j] : -1;
  __syncwarp();
}

template <int CPL>
__global__ void __launch_bounds__(32) build_link_seq_kernel(const BuildParams bp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GraphView& g = bp.sp.g;
  const int lane = threadIdx.x & 31;
  uint64_t* ukey = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* sorted = ukey + bp.ucap;
  float4* qs = reinterpret_cast<float4*>(sorted + bp.ucap);
  float4* qs2 = qs + bp.sp.q_smem_chunks;
  uint32_t* uid = reinterpret_cast<uint32_t*>(qs2 + bp.sp.q_smem_chunks);
  float* ud = reinterpret_cast<float*>(uid + bp.ucap);
  uint32_t* sel = reinterpret_cast<uint32_t*>(ud + bp.ucap);
  float* newd = reinterpret_cast<float*>(sel + bp.sel_cap);
  uint32_t* nbs = reinterpret_cast<uint32_t*>(newd + 32);       // the new node's list on this layer
  uint32_t* pend = nbs + bp.sel_cap;                            // 1: the node is logically at the head of that neighbour's full row
  uint32_t* rem = pend + bp.sel_cap;                            // dropped members of one pruned list
  Stage st;
  stage_attach(st, reinterpret_cast<unsigned char*>(rem + bp.sel_cap), bp.sp.stage_slots, bp.sp.stage_ahead, g.ld4, lane);
  float4 qa[CPL > 0 ? CPL : 1], qe[CPL > 0 ? CPL : 1];
  const uint32_t v = (uint32_t)bp.n0;
  uint32_t n_dist = 0;
  for (int layer = min((int)bp.level[v], g.max_layer); layer >= 0; layer--) {
    const int slots = layer == 0 ? g.slots0 : g.slotsU;
    const int nc = layer == 0 ? bp.cap0 : bp.capU;
    const int32_t* row_v = row_ptr(bp, row_id(g, v, layer));
    int nsel = 0;
    for (int r0 = 0; r0 < slots; r0 += 32) {
      int nb = r0 + lane < slots ? row_v[r0 + lane] : -1;
      unsigned valid = __ballot_sync(FULL, nb >= 0);
      if (nb >= 0) nbs[r0 + lane] = (uint32_t)nb;
      nsel += __popc(valid);
      if (valid != FULL) break;
    }
    __syncwarp();
    // set_connections_for_new_node: Neighbours.add node on every selected neighbour
    for (int j = 0; j < nsel; j++) {
      int32_t* row = row_ptr(bp, row_id(g, nbs[j], layer));
      int deg = 0;
      for (int r0 = 0; r0 < slots; r0 += 32) {
        int nb = r0 + lane < slots ? row[r0 + lane] : -1;
        unsigned valid = __ballot_sync(FULL, nb >= 0);
        if (nb >= 0) uid[r0 + lane] = (uint32_t)nb;
        deg += __popc(valid);
        if (valid != FULL) break;
      }
      __syncwarp();
      if (deg < slots && deg < nc + 1 && deg + 1 <= slots) {
        for (int i = lane; i <= deg; i += 32) row[i] = i == 0 ? (int32_t)v : (int32_t)uid[i - 1];
        if (lane == 0) pend[j] = 0u;
      } else if (lane == 0) pend[j] = 1u;
      __syncwarp();
    }
    // re-select the lists that outgrew nc, in list order
    for (int j = 0; j < nsel; j++) {
      const uint32_t a = nbs[j];
      int32_t* row = row_ptr(bp, row_id(g, a, layer));
      const bool pj = pend[j] != 0u;
      int u = pj ? 1 : 0;
      if (pj && lane == 0) uid[0] = v;
      for (int r0 = 0; r0 < slots; r0 += 32) {
        int nb = r0 + lane < slots ? row[r0 + lane] : -1;
        unsigned valid = __ballot_sync(FULL, nb >= 0);
        if (nb >= 0) uid[u + __popc(valid & ((1u << lane) - 1u))] = (uint32_t)nb;
        u += __popc(valid);
        if (valid != FULL) break;
      }
      __syncwarp();
      if (u <= nc) continue;                                              // :823
      load_target<CPL>(g, reinterpret_cast<const float4*>(g.vec) + (size_t)a * g.ld4, qa, qs, lane);
      batch_dist<CPL>(g, CPL > 0 ? qa : nullptr, qs, uid, ud, u, lane, &st);   // min_queue_of_neighbours (:791-798)
      n_dist += u;
      for (int i = lane; i < u; i += 32) ukey[i] = make_key(ud[i], uid[i]);
      __syncwarp();
      warp_rank_sort(ukey, sorted, u, lane);
      int ns = select_neighbours<CPL>(g, sorted, u, nc, 0, sel, qe, qs2, newd, lane, n_dist, &st);   // :824-826
      for (int i = lane; i < slots; i += 32) row[i] = i < ns ? (int32_t)sel[ns - 1 - i] : -1;   // Graph.set_connections step 1
      if (lane == 0) pend[j] = 0u;
      // step 2: removed = old \ new, walked in ascending id (Set.iter)
      int nr = 0;
      for (int i0 = 0; i0 < u; i0 += 32) {
        int i = i0 + lane;
        bool out = i < u && !(sorted[i] & 1ull);
        unsigned bm = __ballot_sync(FULL, out);
        if (out) rem[nr + __popc(bm & ((1u << lane) - 1u))] = key_id(sorted[i]);
        nr += __popc(bm);
      }
      __syncwarp();
      for (int done = 0; done < nr; done++) {
        // next smallest id not yet handled (nr is small: selection sort by warp minimum)
        uint32_t best = 0xffffffffu;
        for (int i = lane; i < nr; i += 32) best = min(best, rem[i]);
        best = __reduce_min_sync(FULL, best);
        __syncwarp();
        for (int i = lane; i < nr; i += 32) if (rem[i] == best) rem[i] = 0xffffffffu;
        __syncwarp();
        bool pr = false;
        int jr = -1;
        for (int i = lane; i < nsel; i += 32) if (nbs[i] == best) jr = i;
        jr = __reduce_max_sync(FULL, jr);
        if (jr >= 0) pr = pend[jr] != 0u;
        int32_t* rrow = row_ptr(bp, row_id(g, best, layer));
        seq_remove(rrow, slots, pr, v, a, uid, lane);
        if (jr >= 0 && lane == 0) pend[jr] = 0u;
        __syncwarp();
      }
    }
  }
  if (lane == 0) atomicAdd(bp.counters + 0, (unsigned long long)n_dist);
}

// ---- phase 3 --------------------------------------------------------------------------------------
// One warp per row named in the sorted removal requests: drop every owner listed for it.
__global__ void __launch_bounds__(256) build_unlink_kernel(const BuildParams bp) {
  const GraphView& g = bp.sp.g;
  const int lane = threadIdx.x & 31;
  const unsigned nheads = *bp.head_count;
  while (true) {
    unsigned s = 0;
    if (lane == 0) s = atomicAdd(bp.next, 1u);
    s = __shfl_sync(FULL, s, 0);
    if (s >= nheads) break;
    const unsigned h = bp.heads[s];
    const uint32_t rid = (uint32_t)(bp.rem[h] >> REM_ABITS);
    int32_t* row = row_ptr(bp, rid);
    const int slots = (rid & ROW_UPPER) ? g.slotsU : g.slots0;
    int out = 0;
    for (int r0 = 0; r0 < slots; r0 += 32) {
      int nb = r0 + lane < slots ? row[r0 + lane] : -1;
      unsigned valid = __ballot_sync(FULL, nb >= 0);
      bool keep = nb >= 0;
      for (unsigned i = h; i < bp.n_req; i++) {          // the segment is short (owners that dropped this node)
        uint64_t k = bp.rem[i];
        if ((uint32_t)(k >> REM_ABITS) != rid) break;
        if ((uint32_t)(k & ((1ull << REM_ABITS) - 1)) == (uint32_t)nb) keep = false;
      }
      unsigned km = __ballot_sync(FULL, keep);
      __syncwarp();
      if (keep) row[out + __popc(km & ((1u << lane) - 1u))] = nb;
      out += __popc(km);
      __syncwarp();
      if (valid != FULL) break;
    }
    for (int j = out + lane; j < slots; j += 32) row[j] = -1;
    __syncwarp();
  }
}

}  // namespace hb
