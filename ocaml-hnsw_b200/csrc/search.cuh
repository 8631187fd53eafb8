// Batched k-NN search: one warp per query, persistent warps pulling queries from a counter.
//
// Restates, per query, Ohnsw.knn (lib/ohnsw.ml:859-875): greedy descent through layers
// max_layer..1 with search_one_simple (:492-508), then search_k on layer 0 (:543-588).
// The sequential best-first semantics are kept exactly (PARITY): one expansion per iteration,
// neighbours taken in list order, exact visited set, heaps ordered by (distance, id).  What is
// parallel is everything inside one expansion: the adjacency row is one coalesced load, the
// visited test-and-set runs one neighbour per lane, the distances of all unvisited neighbours
// are evaluated in one batch, eight vectors per round (a team of 8 lanes per vector, 128-bit
// loads, packed FFMA2), and the candidates are accepted and merged into the beam 32 at a time
// with the decisions the one-by-one loop would have taken.
//
// Per-warp shared memory: the beam (`near`, <= ef sorted keys), a tie list, the compacted
// neighbour ids/distances of the current expansion, the query, the visited hash (or, for large
// beams, one n-bit set per warp in global memory).
#pragma once
#include "common.cuh"

// Tuning knobs of search_kernel (measured on B200): minimum resident CTAs of 128
// threads per SM (register budget = 65536 / (128 * MINB)) and where the target vector lives.
#ifndef HB_SEARCH_MINB
#define HB_SEARCH_MINB 6
#endif
#ifndef HB_SEARCH_QREG
#define HB_SEARCH_QREG false
#endif

namespace hb {

// Row-sharded index: every shard answers every query; the exchange and the merge of the
// per-shard rows are the tail of the search kernel itself.  A warp stores its finished row (ids already
// global) into the gather block of the HOME device — plain stores through a peer mapping over NVLink —
// and bumps the query's arrival counter there with a system-scope atomic; the warp that finds itself
// last (on whichever GPU it runs) reads the n_shards rows back and writes the merged row to every
// final destination.  No all-gather, no barrier between the shards, no separate merge launch.
struct ShardTail {
  int n_shards;              // 0: not sharded
  int shard;
  int32_t id_offset;         // first global row of this shard
  int32_t* g_ids;            // home: [n_shards][nq][k]
  float* g_dists;
  unsigned int* arrive;      // home: [nq], zero before the call
  int n_final;
  int32_t* f_ids[8];         // [nq][k] each (any device)
  float* f_dists[8];
};

struct SearchParams {
  GraphView g;
  const float* queries;      // [nq][ld4*4]
  int64_t nq;
  int ef, k, ef_cap;
  int accept_ties;           // Hnsw.Ba flavour: accept d <= top (lib/hnsw.ml:494-506)
  int pad_inf;               // Hnsw.Ba flavour: +inf padding (lib/hnsw.ml:771)
  HashCfg hc;                // visited hash geometry (hc.slots == 0: global bitset per warp)
  int nb_cap;                // list slots gathered per pass: 32 or 64 (staging arrays hold this many)
  int q_smem_chunks;         // float4 slots reserved for the query copy
  int stage_slots;           // > 0: rows are staged through a per-warp bulk-copy ring of this many rows (common.cuh)
  int stage_ahead;           // rows beyond the ring prefetched to L2
  int gang;                  // warps per query (1, 2 or 4): see Gang
  int smem_per_warp;
  int32_t* out_ids;          // [nq][k] or null
  float* out_dists;          // [nq][k]
  // multi-GPU: the same result rows are also stored straight into every peer's gathered block
  // (peer-mapped device memory over NVLink), replacing the all-gather that would follow
  int n_peer_out;
  int32_t* peer_ids[8];
  float* peer_dists[8];
  ShardTail tail;
  uint32_t* counters;        // [nq][3]
  unsigned int* next_query;  // work counter
  uint32_t* bitset_pool;     // [pool_size][words]
  int* pool_busy;
  int pool_size;
  int words;
  unsigned long long* events; // [0] visited spills, [1] tie lists that left shared memory, [2] warps that gave up waiting for queries,
                              // [3] tie lists that outgrew their global region too (the only way PARITY can degrade)
  // evicted candidates tied with the beam's top beyond TIES_CAP continue in a region of this pool
  uint64_t* tie_pool;        // [tie_slots][tie_cap]
  int* tie_busy;
  int tie_slots, tie_cap;
  // streamed queries (host-buffer call): the batch arrives in pieces of `ready_step` queries while the
  // kernel runs; *ready = pieces copied so far (written by the copy stream after each piece)
  const unsigned int* ready;
  unsigned int ready_step;
};

__host__ __device__ inline int search_smem_per_warp(int ef_cap, int hash_bytes, int q_chunks, int nb_cap = 32) {
  return ef_cap * 8 + TIES_CAP * 8 + nb_cap * 4 + nb_cap * 4 + q_chunks * 16 + hash_bytes;
}
constexpr int GANG_JOB_BYTES = 32;      // sizeof(GangJob), at the very end of a gang's block
// the bulk-copy ring (when used) follows the fixed part of the warp's block
__device__ __forceinline__ void stage_attach(Stage& st, unsigned char* at, int slots, int ahead, int ld4, int lane) {
  float4* ring = slots > 0 ? reinterpret_cast<float4*>(at) : nullptr;
  uint64_t* bar = slots > 0 ? reinterpret_cast<uint64_t*>(at + (size_t)slots * ld4 * 16) : nullptr;
  stage_init(st, ring, bar, slots, ahead, lane);
}

// ---- a gang: several warps on one query / insert --------------------------------------------------------
// When a batch is smaller than the warps the GPU holds (a build batch early in the build, a replica's
// share of the queries, a single Ohnsw.knn) one warp per item leaves the machine idle and the item's
// dependent chain — row, visited, 3-4 distance rounds, merge — sets the time.  A gang of P warps then
// works on ONE item: warp 0 runs the traversal exactly as a lone warp does, and hands every distance
// batch of more than one round to all P warps (round r goes to warp r mod P); the others wait on a
// named barrier between batches.  Distances do not depend on who evaluates them, so the result is the
// lone warp's, bit for bit.
struct GangJob {
  const uint32_t* ids;
  float* d;
  const float4* target;
  int cnt;                   // < 0: no more work, the helpers leave
  int pad;
};
struct Gang {
  int P, rank, bar;          // warps in the gang, this warp's place, named barrier (1..15)
  GangJob* job;              // shared
};
__device__ __forceinline__ void gang_sync(const Gang& gg) {
  asm volatile("bar.sync %0, %1;" ::"r"(gg.bar), "r"(gg.P * 32) : "memory");
}
// helpers' loop: evaluate the rounds that fall to this warp until the leader says stop
template <int CPL>
__device__ __forceinline__ void gang_help(const GraphView& g, const Gang& gg, int lane) {
  while (true) {
    gang_sync(gg);
    const int cnt = gg.job->cnt;
    if (cnt < 0) return;
    batch_dist_rounds<CPL>(g, gg.job->target, gg.job->ids, gg.job->d, cnt, lane, gg.rank, gg.P);
    gang_sync(gg);
  }
}
__device__ __forceinline__ void gang_dismiss(const Gang& gg, int lane) {
  if (gg.P <= 1) return;
  if (lane == 0) gg.job->cnt = -1;
  __syncwarp();
  gang_sync(gg);
}

// QREG: the target vector lives in registers (CPL float4 per lane); otherwise in shared memory
// (qs, zero padded to TEAM * CPL chunks) — fewer live registers, more resident warps.
template <int CPL, bool QREG = true>
struct WarpCtx {
  uint64_t* keys;
  uint64_t* ties;
  uint32_t* newid;
  float* newd;
  float4* qs;
  float4 q[(CPL > 0 && QREG) ? CPL : 1];
  VisitedSet vis;
  Stage st;
  uint64_t* tie_spill;       // non-null while this warp holds a region of the tie pool
  int tie_slot;
  Gang gang;
  int lane;
  __device__ __forceinline__ const float4* target_regs() const { return (CPL > 0 && QREG) ? q : nullptr; }
};

__device__ __forceinline__ void visited_spill(VisitedSet& v, const SearchParams& p, int lane) {
  int s = -1;
  if (lane == 0) {
    unsigned start = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % (unsigned)p.pool_size;
    for (unsigned i = start;; i = (i + 1 == (unsigned)p.pool_size ? 0 : i + 1))
      if (atomicCAS(&p.pool_busy[i], 0, 1) == 0) { s = (int)i; break; }
    __threadfence();
    atomicAdd(p.events, 1ull);
  }
  s = __shfl_sync(FULL, s, 0);
  v.pool_slot = s;
  v.bits = p.bitset_pool + (size_t)s * p.words;
  for (uint32_t i = lane; i < p.hc.slots; i += 32) {
    const uint32_t id = hash_decode(v.tab, p.hc, i);
    if (id != 0xffffffffu) atomicOr(&v.bits[id >> 5], 1u << (id & 31));
  }
  __syncwarp();
}
__device__ __forceinline__ void visited_release(VisitedSet& v, const SearchParams& p, int lane) {
  if (v.bits) {
    uint4* b4 = reinterpret_cast<uint4*>(v.bits);
    for (int i = lane; i < p.words / 4; i += 32) b4[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    if (p.hc.slots == 0u) return;              // bitset is this warp's primary set: it keeps its slot
    __threadfence();
    if (lane == 0) atomicExch(&p.pool_busy[v.pool_slot], 0);
    v.bits = nullptr;
  }
}
// Large beams: the visited set of a query would not leave room for enough resident warps in
// shared memory, so every warp owns one n-bit set of the global pool for its whole life.
__device__ __forceinline__ void visited_init(VisitedSet& v, const SearchParams& p, uint32_t* tab) {
  v.tab = tab;
  v.limit = p.hc.slots / 4u * 3u;                                             // load <= 0.75
  v.bits = nullptr;
  v.pool_slot = -1;
  if (p.hc.slots == 0) {
    v.pool_slot = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);        // host guarantees pool_size >= warps
    v.bits = p.bitset_pool + (size_t)v.pool_slot * p.words;
    v.limit = 0xffffffffu;
  }
}
// The set belongs to one warp and the ids of one call are distinct, so the test is a plain (L2) load and the
// set a fire-and-forget reduction (RED, no return value to wait for): lanes that share a word each test their
// own bit of the same old value.  __syncwarp orders a call's reductions before the next call's loads.
__device__ __forceinline__ bool bitset_test_and_set(uint32_t* bits, uint32_t id) {
  const uint32_t bit = 1u << (id & 31);
#ifdef HB_BITSET_ATOMIC
  return !(atomicOr(&bits[id >> 5], bit) & bit);
#else
  const bool is_new = !(__ldcg(&bits[id >> 5]) & bit);
  if (is_new) atomicOr(&bits[id >> 5], bit);
  return is_new;
#endif
}
// Visited.mem / Visited.add for up to one id per lane (`active` lanes): true where the id was not yet
// visited (and marks it).  All 32 lanes must call.  A lane whose id cannot be placed in the 16-bit
// table makes the whole set move to the global bitset, where it is then placed.
__device__ __forceinline__ bool visited_test_and_set(VisitedSet& v, const SearchParams& p, bool active, uint32_t id, int lane) {
  if (v.bits) return active && bitset_test_and_set(v.bits, id);
  if (!p.hc.bits16) return active && hash_test_and_set(v.tab, p.hc, id) == 1;
  const int r = hash16_test_and_set_warp(v.tab, p.hc, active, id);
  if (!__any_sync(FULL, r == 2)) return r == 1;
  __syncwarp();
  visited_spill(v, p, lane);
  return r == 2 ? bitset_test_and_set(v.bits, id) : r == 1;
}

// ---- the tie list: evicted candidates whose distance equals the beam's top ------------------------
// Entries 0..TIES_CAP-1 live in shared memory, later ones in a region borrowed from a global pool,
// so the list never drops an entry the reference's visit_me queue would still pop (:565-568).
template <class W>
__device__ __forceinline__ uint64_t tie_get(const W& w, int i) { return i < TIES_CAP ? w.ties[i] : __ldcg(w.tie_spill + (i - TIES_CAP)); }
template <class W>
__device__ __forceinline__ void tie_set(W& w, int i, uint64_t v) { if (i < TIES_CAP) w.ties[i] = v; else __stcg(w.tie_spill + (i - TIES_CAP), v); }
// borrow a region (bounded wait: a region is held for the rest of one layer search only); false if none could be had
template <class W>
__device__ __forceinline__ bool tie_borrow(W& w, const SearchParams& p, int lane) {
  int s = -1;
  if (lane == 0 && p.tie_slots > 0) {
    unsigned i = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) % (unsigned)p.tie_slots;
    for (unsigned tries = 0; tries < (1u << 22); tries++) {
      if (atomicCAS(&p.tie_busy[i], 0, 1) == 0) { s = (int)i; break; }
      i = i + 1 == (unsigned)p.tie_slots ? 0 : i + 1;
      if ((tries & 31) == 31) __nanosleep(100);
    }
    if (s >= 0) { __threadfence(); atomicAdd(p.events + 1, 1ull); }
  }
  s = __shfl_sync(FULL, s, 0);
  if (s < 0) return false;
  w.tie_slot = s;
  w.tie_spill = p.tie_pool + (size_t)s * p.tie_cap;
  return true;
}
template <class W>
__device__ __forceinline__ void tie_release(W& w, const SearchParams& p, int lane) {
  if (!w.tie_spill) return;
  __syncwarp();
  __threadfence();
  if (lane == 0) atomicExch(&p.tie_busy[w.tie_slot], 0);
  w.tie_spill = nullptr;
}

// search_k (lib/ohnsw.ml:543-588) on layer `layer`, beam already seeded with `n` keys (all
// unexpanded, all marked visited).  Leaves the nearest set in keys[0..n).
template <int CPL, bool QREG, bool GANG = false>
__device__ __forceinline__ void layer_search(const SearchParams& p, WarpCtx<CPL, QREG>& w, int layer, int& n,
                                             uint32_t& n_dist, uint32_t& n_exp, bool& tie_overflow) {
  const GraphView& g = p.g;
  const int lane = w.lane;
  const int ef = p.ef;
  int fu = 0, ties_n = 0;
  float top_d = n == ef ? key_dist(w.keys[ef - 1]) : 0.f;
  while (true) {
    // ---- MinQueue.pop_min visit_me (:565): smallest unexpanded key of near + live tie list
    int pos = -1;
    while (fu < n) {
      int idx = fu + lane;
      bool un = idx < n && !(w.keys[idx] & 1ull);
      unsigned b = __ballot_sync(FULL, un);
      if (b) {
        pos = fu + __ffs(b) - 1;
        // the runner-up is the likely next pop: pull its adjacency row towards L2 now
        unsigned b2 = b & (b - 1u);
        if (b2 && lane == 0) {
          uint32_t nx = key_id(w.keys[fu + __ffs(b2) - 1]);
          if (layer == 0) prefetch_l2(g.adj0 + (size_t)nx * g.slots0);
        }
        break;
      }
      fu += 32;
    }
    if (fu > n) fu = n;
    uint64_t ck = pos >= 0 ? w.keys[pos] : KEY_INF;
    bool from_ties = false;
    if (ties_n > 0) {
      uint64_t mn = KEY_INF;
      for (int b = 0; b < ties_n; b += 32) {
        const uint64_t tk = b + lane < ties_n ? tie_get(w, b + lane) : KEY_INF;
        mn = tk < mn ? tk : mn;
      }
      for (int o = 16; o; o >>= 1) { uint64_t x = __shfl_xor_sync(FULL, mn, o); mn = x < mn ? x : mn; }
      if (mn < ck) {
        int sel = 0;
        for (int b = 0; b < ties_n; b += 32) {                 // keys are distinct: exactly one entry matches
          const unsigned who = __ballot_sync(FULL, b + lane < ties_n && tie_get(w, b + lane) == mn);
          if (who) { sel = b + __ffs(who) - 1; break; }
        }
        const uint64_t lastk = tie_get(w, ties_n - 1);
        __syncwarp();
        if (lane == 0) tie_set(w, sel, lastk);
        ties_n--;
        ck = mn; from_ties = true;
        __syncwarp();
      }
    }
    if (ck == KEY_INF) break;                      // visit_me empty (:566)
    if (!from_ties) {
      if (lane == 0) w.keys[pos] = ck | 1ull;
      fu = pos + 1;
      __syncwarp();
    }
    const uint32_t c = key_id(ck);
    n_exp++;

    // ---- Neighbours.iter (Graph.adjacent graph c.node) (:570), 32 list slots per round
    const int slots = layer == 0 ? g.slots0 : g.slotsU;
    const int32_t* row;
    if (layer == 0) row = g.adj0 + (size_t)c * g.slots0;
    else { int off = g.upper_off[c]; row = off < 0 ? nullptr : g.adjU + ((size_t)off + layer - 1) * g.slotsU; }
    // Up to nb_cap (32 or 64) list slots are gathered per pass: visited test-and-set chunk by chunk,
    // ONE batch of distance evaluations for everything new, then acceptance in list order, 32
    // candidates at a time (rows wider than 32 slots, M > 16, no longer pay a partial round and a
    // merge per 32-slot chunk).
    for (int s0 = 0; s0 < slots && row; s0 += p.nb_cap) {
      int total = 0;
      bool row_ended = false;
      // both 32-slot chunks of a pass are sent for at once (rows wider than 32 slots: M > 16)
      int nbs[2];
      nbs[0] = (s0 + lane < slots) ? __ldg(row + s0 + lane) : -1;
      nbs[1] = (p.nb_cap > 32 && s0 + 32 + lane < slots) ? __ldg(row + s0 + 32 + lane) : -1;
#pragma unroll
      for (int ci = 0; ci < 2; ci++) {
        const int r0 = s0 + 32 * ci;
        if (r0 >= min(slots, s0 + p.nb_cap)) break;
        const int nb = nbs[ci];
        unsigned valid = __ballot_sync(FULL, nb >= 0);
        if (!valid) { row_ended = true; break; }
        if (!w.vis.bits && w.vis.count + 32u > w.vis.limit) visited_spill(w.vis, p, lane);
        const bool is_new = visited_test_and_set(w.vis, p, nb >= 0, (uint32_t)nb, lane);   // Visited.mem / add (:571-572)
        unsigned m = __ballot_sync(FULL, is_new);
        if (is_new) w.newid[total + __popc(m & ((1u << lane) - 1u))] = (uint32_t)nb;
        total += __popc(m);
        if (valid != FULL) { row_ended = true; break; }                       // row ended inside this chunk
      }
      w.vis.count += total;
      __syncwarp();
      if (total) {
        // every vector beyond the first round of eight starts moving towards L2 now, so the
        // later rounds of batch_dist wait for L2, not for HBM
        if (GANG && total > 8) {
          // the whole gang evaluates this batch (w.qs holds the target); rounds beyond the first P are sent for first
          for (int j = lane; j < total; j += 32)
            if (j >= 8 * w.gang.P) {
              const char* vrow = reinterpret_cast<const char*>(g.vec) + (size_t)w.newid[j] * g.ld4 * 16;
              for (int b = 0; b < g.ld4 * 16; b += 128) prefetch_l2(vrow + b);
            }
          if (lane == 0) { w.gang.job->ids = w.newid; w.gang.job->d = w.newd; w.gang.job->target = w.qs; w.gang.job->cnt = total; }
          __syncwarp();
          gang_sync(w.gang);
          batch_dist_rounds<CPL>(g, w.qs, w.newid, w.newd, total, lane, 0, w.gang.P);
          gang_sync(w.gang);
        } else {
        for (int j = lane; j < total && !w.st.ring; j += 32)
          if (j >= 8) {
            const char* vrow = reinterpret_cast<const char*>(g.vec) + (size_t)w.newid[j] * g.ld4 * 16;
            for (int b = 0; b < g.ld4 * 16; b += 128) prefetch_l2(vrow + b);
          }
        batch_dist<CPL>(g, w.target_regs(), w.qs, w.newid, w.newd, total, lane, &w.st);   // MinQueue.element (:573)
        }
        n_dist += total;
      }
      for (int g0 = 0; g0 < total; g0 += 32) {
        const int cnt = min(32, total - g0);
        // ---- accept (:574-578).  The reference takes the candidates one at a time, in list
        // order: accept iff |near| < ef or t < top (t <= top in the Hnsw.Ba flavour), insert, evict
        // the maximum.  The same decisions for 32 candidates at once: with U_j = the beam at the
        // start of the group plus all earlier candidates, top_j is the ef-th smallest distance of
        // U_j, so candidate j is accepted iff fewer than ef members of U_j are at distance <= t_j
        // (< t_j when ties are accepted); the beam after the group is the ef smallest keys of
        // beam + accepted, and an evicted entry stays poppable only while its distance equals
        // the top (stop rule is a strict >, :568).
        const float t = lane < cnt ? w.newd[g0 + lane] : 0.f;
        const uint64_t key = make_key(t, lane < cnt ? w.newid[g0 + lane] : 0u);
        const bool pre = lane < cnt && (n < ef || (p.accept_ties ? t <= top_d : t < top_d));
        const unsigned pm = __ballot_sync(FULL, pre);
        if (pm) {
          int posK = 0;                               // beam keys smaller than mine
          if (pre) {
            int lo = 0, hi = n;
            while (lo < hi) { int mid = (lo + hi) >> 1; if (w.keys[mid] < key) lo = mid + 1; else hi = mid; }
            posK = lo;
          }
          bool acc = pre;
          if (n + __popc(pm) > ef) {
            int c = posK;                             // beam members at distance <= t (< t with ties)
            if (pre) {
              if (p.accept_ties) { while (c > 0 && key_dist(w.keys[c - 1]) == t) c--; }
              else { while (c < n && key_dist(w.keys[c]) == t) c++; }
            }
            for (unsigned m = pm; m; m &= m - 1u) {   // earlier candidates of this row
              const int i = __ffs(m) - 1;
              const float ti = __shfl_sync(FULL, t, i);
              c += (i < lane && (p.accept_ties ? ti < t : ti <= t)) ? 1 : 0;
            }
            acc = pre && c < ef;
          }
          const unsigned am = __ballot_sync(FULL, acc);
          if (am) {
            const bool was_full = n == ef;
            int rank = 0;
            for (unsigned m = am; m; m &= m - 1u) {
              const int i = __ffs(m) - 1;
              const uint64_t ki = __shfl_sync(FULL, key, i);
              rank += ki < key ? 1 : 0;
            }
            const int minpos = __reduce_min_sync(FULL, acc ? posK : 0x7fffffff);
            const int f = posK + rank;                // my key's place in the merged order
            // The merged beam, final position by final position, top chunk first: position p takes an accepted key
            // (written by its owner below) or the old entry p - (accepted keys placed below p).  The accepted
            // places of a chunk are one OR-reduction, the count below the chunk one ballot — no loop over the
            // accepted keys per chunk.  Old entries that no longer fit (at most one per accepted key) are read first.
            const int a_in = __popc(__ballot_sync(FULL, acc && f < ef));
            const int nkeep = min(ef, n + __popc(am));
            const int first_out = nkeep - a_in;       // old entries [first_out, n) fall off
            const bool ev = first_out + lane < n;
            const uint64_t ev_key = ev ? w.keys[first_out + lane] : 0ull;
            for (int base = (nkeep - 1) & ~31; base >= (minpos & ~31); base -= 32) {
              const unsigned word = __reduce_or_sync(FULL, (acc && (f & ~31) == base) ? 1u << (f & 31) : 0u);
              const int below = __popc(__ballot_sync(FULL, acc && f < base));
              const int pf = base + lane;
              const bool old_here = pf < nkeep && !((word >> lane) & 1u);
              const uint64_t kk = old_here ? w.keys[pf - below - __popc(word & ((1u << lane) - 1u))] : 0ull;
              __syncwarp();
              if (old_here) w.keys[pf] = kk;
            }
            __syncwarp();
            if (acc && f < ef) w.keys[f] = key;
            if (minpos < fu) fu = minpos;           // the lowest accepted key has rank 0: its place is minpos
            n = min(ef, n + __popc(am));
            __syncwarp();
            if (n == ef) {
              const float new_top = key_dist(w.keys[ef - 1]);
              if (was_full && new_top < top_d) ties_n = 0;
              top_d = new_top;
              // entries that fell off but tie with the top stay in visit_me (Heap.pop_exn nearest_maxq, :577)
              const bool t1 = ev && !(ev_key & 1ull) && key_dist(ev_key) == new_top;
              const bool t2 = acc && f >= ef && t == new_top;
              const unsigned m1 = __ballot_sync(FULL, t1), m2 = __ballot_sync(FULL, t2);
              const int c1 = __popc(m1), c2 = __popc(m2);
              bool room = ties_n + c1 + c2 <= TIES_CAP;
              if (!room) room = (w.tie_spill || tie_borrow(w, p, lane)) && ties_n + c1 + c2 <= TIES_CAP + p.tie_cap;
              if (!room) tie_overflow = true;
              else {
                if (t1) tie_set(w, ties_n + __popc(m1 & ((1u << lane) - 1u)), ev_key);
                if (t2) tie_set(w, ties_n + c1 + __popc(m2 & ((1u << lane) - 1u)), key);
                ties_n += c1 + c2;
              }
              __syncwarp();
            }
          }
        }
      }
      if (row_ended) break;
    }
  }
  tie_release(w, p, lane);
}


// target vector -> the team-distributed register copy (CPL > 0) or the shared copy (CPL == 0)
template <int CPL>
__device__ __forceinline__ void load_target(const GraphView& g, const float4* row, float4* q, float4* qs, int lane) {
  const int tl = lane & (TEAM - 1);
  if (CPL > 0) {
#pragma unroll
    for (int c = 0; c < (CPL > 0 ? CPL : 1); c++) {
      int ch = tl + TEAM * c;
      q[c] = ch < g.chunks ? __ldg(row + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    __syncwarp();
    for (int ch = lane; ch < g.chunks; ch += 32) qs[ch] = __ldg(row + ch);
    __syncwarp();
  }
}

// target vector -> shared memory, zero padded to `padded` chunks (the register-free variant)
__device__ __forceinline__ void load_target_smem(const GraphView& g, const float4* row, float4* qs, int padded, int lane) {
  __syncwarp();
  // (__ldcg, not the read-only path: with streamed queries the rows are written while the kernel runs)
  for (int ch = lane; ch < padded; ch += 32) qs[ch] = ch < g.chunks ? __ldcg(row + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
  __syncwarp();
}

// search_one_simple (lib/ohnsw.ml:492-508) on one upper layer: scan the current best's row,
// move to the row minimum if it is strictly closer, repeat until nothing moves.
template <int CPL, bool QREG>
__device__ __forceinline__ void greedy_layer(const GraphView& g, WarpCtx<CPL, QREG>& w, int layer, uint32_t& cur,
                                             float& d_cur, uint32_t& n_dist, uint32_t& n_expU) {
  const int lane = w.lane;
  n_dist++;                                       // best_distance = distance (value start) target (:496)
  bool changed = true;
  while (changed) {
    changed = false;
    int off = g.upper_off[cur];
    const int32_t* row = off < 0 ? nullptr : g.adjU + ((size_t)off + layer - 1) * g.slotsU;
    n_expU++;
    uint64_t best = KEY_INF;                      // (distance, list position) of the running minimum
    uint32_t best_id = 0;
    for (int r0 = 0; r0 < g.slotsU && row; r0 += 32) {
      int nb = (r0 + lane < g.slotsU) ? __ldg(row + r0 + lane) : -1;
      unsigned m = __ballot_sync(FULL, nb >= 0);
      int cnt = __popc(m);
      if (!cnt) break;
      if (nb >= 0) w.newid[__popc(m & ((1u << lane) - 1u))] = (uint32_t)nb;
      __syncwarp();
      batch_dist<CPL>(g, w.target_regs(), w.qs, w.newid, w.newd, cnt, lane, &w.st);
      n_dist += cnt;
      uint64_t mine = lane < cnt ? (((uint64_t)f2ord(w.newd[lane]) << 32) | (uint32_t)(r0 + lane)) : KEY_INF;
      uint64_t mn = mine;
      for (int o = 16; o; o >>= 1) { uint64_t x = __shfl_xor_sync(FULL, mn, o); mn = x < mn ? x : mn; }
      if (mn < best) {
        best = mn;
        int src = __ffs(__ballot_sync(FULL, mine == mn)) - 1;
        best_id = w.newid[src];
      }
      __syncwarp();
      if (m != FULL) break;
    }
    // the scan moves `best` on every strictly closer neighbour (:502): it ends on the first
    // occurrence of the row minimum, if that is strictly closer than the current node
    if (best != KEY_INF) {
      float bd = ord2f((uint32_t)(best >> 32));
      if (bd < d_cur) { d_cur = bd; cur = best_id; changed = true; }
    }
  }
}

// see ShardTail.  keys[0..n) = this shard's beam, ascending.
__device__ __forceinline__ void shard_tail(const SearchParams& p, unsigned qi, int lane, const uint64_t* keys, int n) {
  const ShardTail& t = p.tail;
  const GraphView& g = p.g;
  const size_t nqk = (size_t)p.nq * p.k;
  const size_t row = (size_t)qi * p.k;
  for (int i = lane; i < p.k; i += 32) {
    int32_t oid = -1;
    float od = __int_as_float(0x7fc00000);
    if (i < n) {
      const uint64_t key = keys[i];
      oid = (int32_t)key_id(key) + t.id_offset;
      const float d = key_dist(key);
      od = g.metric == 0 ? (float)sqrt((double)d) : d;
    }
    t.g_ids[(size_t)t.shard * nqk + row + i] = oid;
    t.g_dists[(size_t)t.shard * nqk + row + i] = od;
  }
  __syncwarp();
  __threadfence_system();
  unsigned prev = 0;
  if (lane == 0) prev = atomicAdd_system(t.arrive + qi, 1u);
  prev = __shfl_sync(FULL, prev, 0);
  if (prev + 1u != (unsigned)t.n_shards) return;
  // last to arrive: S-way merge, lane s holds the head of shard s's row (merge.cuh does the same after an all-gather)
  __threadfence_system();
  const volatile int32_t* gi = t.g_ids;
  const volatile float* gd = t.g_dists;
  int head = 0;
  const size_t base = (size_t)(lane < t.n_shards ? lane : 0) * nqk + row;
  for (int j = 0; j < p.k; j++) {
    uint64_t key = KEY_INF;
    float myd = 0.f;
    if (lane < t.n_shards && head < p.k) {
      const int32_t id = gi[base + head];
      myd = gd[base + head];
      // rows hold the OUTPUT distance (sqrt for L2): monotone in the beam's order, ties broken by global id
      if (id >= 0) key = make_key(myd, (uint32_t)id);
    }
    uint64_t mn = key;
    for (int o = 16; o; o >>= 1) { uint64_t x = __shfl_xor_sync(FULL, mn, o); mn = x < mn ? x : mn; }
    const unsigned who = __ballot_sync(FULL, key == mn && mn != KEY_INF);
    const int src = who ? __ffs(who) - 1 : 0;
    const float d = __shfl_sync(FULL, myd, src);
    if (who && lane == src) head++;
    if (lane < t.n_final) {
      t.f_ids[lane][row + j] = who ? (int32_t)key_id(mn) : -1;
      t.f_dists[lane][row + j] = who ? d : (p.pad_inf ? __int_as_float(0x7f800000) : __int_as_float(0x7fc00000));
    }
  }
}

// GANG: p.gang (2 or 4) warps per query, one query per CTA (the one-warp-per-query instance carries no gang code)
template <int CPL, bool GANG = false>
__global__ void __launch_bounds__(128, HB_SEARCH_MINB) search_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GraphView& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* my = smem_raw + (size_t)(GANG ? 0 : warp) * p.smem_per_warp;      // one block per warp, or one for the gang
  WarpCtx<CPL, HB_SEARCH_QREG> w;
  w.lane = lane;
  w.gang.P = GANG ? p.gang : 1; w.gang.rank = GANG ? warp : 0; w.gang.bar = 1;
  w.gang.job = reinterpret_cast<GangJob*>(my + p.smem_per_warp - (int)sizeof(GangJob));
  if (GANG && w.gang.rank > 0) { gang_help<CPL>(g, w.gang, lane); return; }
  w.keys = reinterpret_cast<uint64_t*>(my);
  w.ties = w.keys + p.ef_cap;
  w.newid = reinterpret_cast<uint32_t*>(w.ties + TIES_CAP);
  w.newd = reinterpret_cast<float*>(w.newid + p.nb_cap);
  w.qs = reinterpret_cast<float4*>(w.newd + p.nb_cap);
  visited_init(w.vis, p, reinterpret_cast<uint32_t*>(w.qs + p.q_smem_chunks));
  stage_attach(w.st, reinterpret_cast<unsigned char*>(w.vis.tab) + p.hc.bytes, p.stage_slots, p.stage_ahead, g.ld4, lane);
  w.tie_spill = nullptr; w.tie_slot = -1;

  while (true) {
    unsigned qi = 0;
    if (lane == 0) qi = atomicAdd(p.next_query, 1u);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= (unsigned)p.nq) break;
    if (p.ready) {
      // wait (bounded) until the piece holding this query has been copied in
      const unsigned need = qi / p.ready_step + 1u;
      int ok = 1;
      if (lane == 0) {
        unsigned spins = 0;
        while (*reinterpret_cast<const volatile unsigned int*>(p.ready) < need) {
          __nanosleep(200);
          if (++spins > (1u << 23)) { ok = 0; atomicAdd(p.events + 2, 1ull); break; }
        }
        __threadfence();
      }
      if (!__shfl_sync(FULL, ok, 0)) break;
    }

    // target -> registers (or shared for the generic path)
    if (HB_SEARCH_QREG) load_target<CPL>(g, reinterpret_cast<const float4*>(p.queries) + (size_t)qi * g.ld4, w.q, w.qs, lane);
    else load_target_smem(g, reinterpret_cast<const float4*>(p.queries) + (size_t)qi * g.ld4, w.qs, p.q_smem_chunks, lane);
    visited_clear(w.vis, p.hc, lane);

    uint32_t n_dist = 0, n_exp0 = 0, n_expU = 0;
    bool tie_overflow = false;

    // ---- knn (:859-875): entry point, greedy descent (search_one_simple :492-508)
    uint32_t cur = (uint32_t)g.entry;
    if (lane == 0) w.newid[0] = cur;
    __syncwarp();
    batch_dist<CPL>(g, w.target_regs(), w.qs, w.newid, w.newd, 1, lane, &w.st);
    float d_cur = w.newd[0];
    __syncwarp();
    for (int layer = g.max_layer; layer >= 1; layer--) greedy_layer(g, w, layer, cur, d_cur, n_dist, n_expU);

    // ---- search_k on layer 0 seeded with {node} (:870-873)
    n_dist++;                                       // MinQueue.add_node w_queue !node
    if (lane == 0) w.keys[0] = make_key(d_cur, cur);
    int n = 1;
    visited_test_and_set(w.vis, p, lane == 0, cur, lane);
    w.vis.count = 1;
    __syncwarp();
    layer_search<CPL, HB_SEARCH_QREG, GANG>(p, w, 0, n, n_dist, n_exp0, tie_overflow);

    // ---- pop ascending into the result rows (:886-893)
    for (int i = lane; i < p.k; i += 32) {
      int32_t oid = -1;
      float od = p.pad_inf ? __int_as_float(0x7f800000) : __int_as_float(0x7fc00000);
      if (i < n) {
        uint64_t key = w.keys[i];
        oid = (int32_t)key_id(key);
        float d = key_dist(key);
        od = g.metric == 0 ? (float)sqrt((double)d) : d;   // Float.sqrt in double, fp32 store (:889,:899)
      }
      if (p.out_ids) p.out_ids[(size_t)qi * p.k + i] = oid;
      if (p.out_dists) p.out_dists[(size_t)qi * p.k + i] = od;
      for (int r = 0; r < p.n_peer_out; r++) {
        p.peer_ids[r][(size_t)qi * p.k + i] = oid;
        p.peer_dists[r][(size_t)qi * p.k + i] = od;
      }
    }
    if (p.tail.n_shards > 0) shard_tail(p, qi, lane, w.keys, n);
    if (lane == 0) {
      if (p.counters) {
        p.counters[(size_t)qi * 3 + 0] = n_dist;
        p.counters[(size_t)qi * 3 + 1] = n_exp0;
        p.counters[(size_t)qi * 3 + 2] = n_expU;
      }
      if (tie_overflow) atomicAdd(p.events + 3, 1ull);
    }
    visited_release(w.vis, p, lane);
    __syncwarp();
  }
  if (GANG) gang_dismiss(w.gang, lane);
}

}  // namespace hb
