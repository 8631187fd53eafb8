// Device-side building blocks shared by the search and build kernels (sm_100a).
//
// Data layout in HBM (one index = one GPU):
//   vec   : float[cap][ld]      ld = dim rounded up to 4 floats, zero padded; one 16-byte
//                               aligned row per node -> every gather is 128-bit loads
//   adj0  : int32[cap][slots0]  layer-0 adjacency, fixed width (2M), list order kept
//                               (slot 0 = list head, lib/ohnsw.ml:116), -1 terminated
//   upper_off : int32[cap]      first upper-layer row of the node, -1 if level 0
//   adjU  : int32[rowsU][slotsU] rows of layers >= 1: node i, layer l -> row upper_off[i]+l-1
//   level : int8[cap]
//
// Distances are evaluated by a TEAM of 8 lanes per vector, one float4 per lane per step, with
// a fixed summation order (see oracle/ohnsw_oracle.hpp, SUM_TEAM8): lane t accumulates chunks
// t, t+8, ... in index order with fused multiply-add, then a butterfly over lane^4, ^2, ^1.
// A warp evaluates four vectors per load instruction, each team reading one whole 128-byte
// line per step.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TEAM = 8;
constexpr int TIES_CAP = 32;
constexpr uint64_t KEY_INF = ~0ull;

struct GraphView {
  const float* vec;
  const int32_t* adj0;
  const int32_t* upper_off;
  const int32_t* adjU;
  int ld4;        // float4 per vector row
  int chunks;     // float4 chunks that carry data ( = ld4 )
  int slots0;
  int slotsU;
  int max_layer;
  int entry;
  int n;
  int metric;     // 0 L2, 1 angular, 2 ip
};

// ---- keys: (distance, id) total order in one u64; bit 0 = "expanded" ---------------------------
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
  f = f + 0.0f;                                   // -0 -> +0: equal distances get equal bits
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
  return ((uint64_t)f2ord(d) << 32) | ((uint64_t)id << 1);
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return (uint32_t)(k >> 1) & 0x7fffffffu; }
__host__ __device__ __forceinline__ float key_dist(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }

// ---- distance ------------------------------------------------------------------------------------
// Packed fp32 pairs (FADD2 / FFMA2 on sm_100a): one instruction per two components, each
// component rounded exactly like the scalar operation.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rr; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.f32x2 rr, ra, rb; mov.b64 {%0, %1}, rr;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rr; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rr, ra, rb, rc; mov.b64 {%0, %1}, rr;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
// One float4 chunk into the lane's accumulator pair: components x,z feed the first accumulator,
// y,w the second (the SUM_TEAM8 order of oracle/ohnsw_oracle.hpp).
__device__ __forceinline__ float2 acc4(float2 acc, const float4& a, const float4& b, bool dot) {
  if (dot) {
    acc = fma2(make_float2(a.x, a.y), make_float2(b.x, b.y), acc);
    acc = fma2(make_float2(a.z, a.w), make_float2(b.z, b.w), acc);
  } else {
    float2 d = sub2(make_float2(a.x, a.y), make_float2(b.x, b.y));
    acc = fma2(d, d, acc);
    d = sub2(make_float2(a.z, a.w), make_float2(b.z, b.w));
    acc = fma2(d, d, acc);
  }
  return acc;
}
__device__ __forceinline__ float team_reduce(float2 acc2) {
  float acc = __fadd_rn(acc2.x, acc2.y);
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
  return acc;
}
// Two team sums with three shuffles instead of six: in the first butterfly step each lane keeps
// one of the two values (lanes 0-3 of the team the first, lanes 4-7 the second) and sends the
// other, so the remaining two steps reduce one value per lane.  Every partial sum is formed from
// the same pair of operands as in team_reduce, so the results are bit-identical; value 0 ends in
// team lanes 0-3, value 1 in lanes 4-7.
__device__ __forceinline__ float team_reduce2(float2 a2, float2 b2, int tl) {
  const float a = __fadd_rn(a2.x, a2.y), b = __fadd_rn(b2.x, b2.y);
  const bool upper = (tl & 4) != 0;
  const float send = upper ? a : b, keep = upper ? b : a;
  float acc = __fadd_rn(keep, __shfl_xor_sync(FULL, send, 4));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 2));
  acc = __fadd_rn(acc, __shfl_xor_sync(FULL, acc, 1));
  return acc;
}
__device__ __forceinline__ float finish_metric(float acc, int metric) {
  return metric == 0 ? acc : (metric == 1 ? __fsub_rn(1.0f, acc) : -acc);
}
__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Distances from the target (register copy q[CPL] when CPL > 0, shared copy qs otherwise) to
// NV nodes per team (NV * 4 vectors per warp), all loads issued before the first use.  Every
// node index must be valid (callers clamp); all 32 lanes must call.  Padding chunks hold zeros
// in both operands and add exactly +0.
template <int CPL, int NV>
__device__ __forceinline__ void team_dist(const GraphView& g, const float4* q, const float4* qs, const int (&node)[NV],
                                          int tl, float (&out)[NV]) {
  const bool dot = g.metric != 0;
  float2 acc[NV];
  const float4* r[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    acc[v] = make_float2(0.f, 0.f);
    r[v] = reinterpret_cast<const float4*>(g.vec) + (size_t)node[v] * g.ld4;
  }
  if (CPL > 0) {
    float4 x[NV][CPL > 0 ? CPL : 1];
    if (g.chunks == TEAM * CPL) {                 // every lane has CPL real chunks (dim 32, 64, 96, 128)
#pragma unroll
      for (int c = 0; c < CPL; c++)
#pragma unroll
        for (int v = 0; v < NV; v++) x[v][c] = ldg4(r[v] + tl + TEAM * c);
    } else {
#pragma unroll
      for (int c = 0; c < CPL; c++) {
        const bool in = tl + TEAM * c < g.chunks;
#pragma unroll
        for (int v = 0; v < NV; v++) x[v][c] = in ? ldg4(r[v] + tl + TEAM * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    // q == nullptr (a literal at the call site): the target stays in shared memory, padded with
    // zero chunks to TEAM * CPL, and is read chunk by chunk — 4 * CPL fewer live registers
#pragma unroll
    for (int c = 0; c < CPL; c++) {
      const float4 qq = q ? q[c] : qs[tl + TEAM * c];
#pragma unroll
      for (int v = 0; v < NV; v++) acc[v] = acc4(acc[v], qq, x[v][c], dot);
    }
  } else {
    for (int ch = tl; ch < g.chunks; ch += 4 * TEAM) {
      float4 x[NV][4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = ch + u * TEAM;
#pragma unroll
        for (int v = 0; v < NV; v++) x[v][u] = c < g.chunks ? ldg4(r[v] + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = ch + u * TEAM;
        const float4 qq = c < g.chunks ? qs[c] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int v = 0; v < NV; v++) acc[v] = acc4(acc[v], qq, x[v][u], dot);
      }
    }
  }
  if (NV == 2) {
    const float r = finish_metric(team_reduce2(acc[0], acc[NV - 1], tl), g.metric);
    out[0] = r; out[NV - 1] = r;                 // out[0] is valid in team lanes 0-3, out[1] in lanes 4-7
  } else {
#pragma unroll
    for (int v = 0; v < NV; v++) out[v] = finish_metric(team_reduce(acc[v]), g.metric);
  }
}

// ids[0..cnt) (shared) -> d[0..cnt) (shared).  cnt >= 1 is warp-uniform.  Rounds of eight
// vectors (two per team) while more than four remain, then one round of up to four.
struct Stage;
__device__ __forceinline__ void batch_dist_staged(const GraphView& g, const float4* qs, const uint32_t* ids, float* d,
                                                  int cnt, int lane, Stage& st);
__device__ __forceinline__ bool stage_on(const Stage* st);
template <int CPL>
__device__ __forceinline__ void batch_dist(const GraphView& g, const float4* q, const float4* qs,
                                           const uint32_t* ids, float* d, int cnt, int lane, Stage* st = nullptr) {
  if (CPL == 0 && stage_on(st)) { batch_dist_staged(g, qs, ids, d, cnt, lane, *st); return; }
  const int tl = lane & (TEAM - 1), team = lane >> 3;
  int base = 0;
  for (; base + 4 < cnt; base += 8) {
    const int j0 = base + team, j1 = base + 4 + team;
    const int node[2] = {(int)ids[j0], (int)ids[min(j1, cnt - 1)]};
    float o[2];
    team_dist<CPL, 2>(g, q, qs, node, tl, o);
    if (tl == 0) d[j0] = o[0];
    if (tl == 4 && j1 < cnt) d[j1] = o[1];
  }
  if (base < cnt) {
    const int j0 = base + team;
    const int node[1] = {(int)ids[min(j0, cnt - 1)]};
    float o[1];
    team_dist<CPL, 1>(g, q, qs, node, tl, o);
    if (tl == 0 && j0 < cnt) d[j0] = o[0];
  }
  __syncwarp();
}

// The rounds first, first + step, ... of batch_dist (round r = vectors 8r .. 8r+7): the share of one warp when
// several warps evaluate one batch (search.cuh, Gang).  The target is the shared copy qs.  Same arithmetic
// per vector as batch_dist, so which warp evaluates a vector does not change its distance.
template <int CPL>
__device__ __forceinline__ void batch_dist_rounds(const GraphView& g, const float4* qs, const uint32_t* ids, float* d,
                                                  int cnt, int lane, int first, int step) {
  const int tl = lane & (TEAM - 1), team = lane >> 3;
  for (int base = 8 * first; base < cnt; base += 8 * step) {
    if (base + 4 < cnt) {
      const int j0 = base + team, j1 = base + 4 + team;
      const int node[2] = {(int)ids[j0], (int)ids[min(j1, cnt - 1)]};
      float o[2];
      team_dist<CPL, 2>(g, nullptr, qs, node, tl, o);
      if (tl == 0) d[j0] = o[0];
      if (tl == 4 && j1 < cnt) d[j1] = o[1];
    } else {
      const int j0 = base + team;
      const int node[1] = {(int)ids[min(j0, cnt - 1)]};
      float o[1];
      team_dist<CPL, 1>(g, nullptr, qs, node, tl, o);
      if (tl == 0 && j0 < cnt) d[j0] = o[0];
    }
  }
  __syncwarp();
}

// ---- high-dimension rows: bulk-copy (1-D TMA) staged gather ---------------------------------------
// Rows of >= STAGE_MIN_BYTES (dim >= 256) are not fetched with per-lane LDG: one elected lane issues
// one `cp.async.bulk` per row (a single UBLKCP instruction moves a whole 3 840-byte GIST row) into a
// per-warp ring of `slots` rows in shared memory, each slot with its own mbarrier (complete_tx counts
// the bytes), and the teams consume the rows from shared memory — four rows per round, one per
// team — in the SAME summation order as the LDG path (SUM_TEAM8), so results stay bit-identical.
// Every slot is refilled with the row `slots` places further on as soon as its round is done, so
// `slots` rows (30 KB at 8 x 3 840 B) stay in flight per warp, against 4 KB for the LDG path.
constexpr int STAGE_MIN_BYTES = 1024;
constexpr int STAGE_MAX_SLOTS = 32;
struct Stage {
  float4* ring;        // slots rows of ld4 float4 each; null = rows are read with LDG
  uint64_t* bar;       // [slots] mbarriers
  int slots;
  int ahead;           // rows beyond the ring that are sent for with a bulk L2 prefetch (bounded: unbounded
                       // look-ahead of 100 KB per warp x 1000 warps evicts the rows from L2 before they are used)
  uint32_t parity;     // bit s = phase parity the next wait on slot s must observe (warp-uniform)
};
__device__ __forceinline__ bool stage_on(const Stage* st) { return st != nullptr && st->ring != nullptr; }
__host__ __device__ inline int stage_smem_bytes(int slots, int ld4) { return slots * ld4 * 16 + ((slots * 8 + 15) & ~15); }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// lane 0 of every warp initialises its own barriers; the fences make them visible to the async proxy
__device__ __forceinline__ void stage_init(Stage& st, float4* ring, uint64_t* bar, int slots, int ahead, int lane) {
  st.ring = ring; st.bar = bar; st.slots = slots; st.ahead = ahead; st.parity = 0u;
  if (ring && lane == 0) {
    for (int i = 0; i < slots; i++) mbar_init(bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
}

// batch_dist through the ring: ids[0..cnt) -> d[0..cnt).  Row j travels through slot j % slots; round
// r evaluates rows 4r .. 4r+3, one per team.  The target is the shared copy qs.
__device__ __forceinline__ void batch_dist_staged(const GraphView& g, const float4* qs, const uint32_t* ids, float* d,
                                                  int cnt, int lane, Stage& st) {
  const int tl = lane & (TEAM - 1), team = lane >> 3;
  const bool dot = g.metric != 0;
  const uint32_t row_bytes = (uint32_t)g.ld4 * 16u;
  const int R = st.slots;
  if (lane == 0) {
    const int m = min(cnt, R);
    for (int j = 0; j < m; j++) {
      mbar_expect_tx(st.bar + j, row_bytes);
      bulk_g2s(st.ring + (size_t)j * g.ld4, reinterpret_cast<const float4*>(g.vec) + (size_t)ids[j] * g.ld4, row_bytes, st.bar + j);
    }
  } else if (lane <= st.ahead && R + lane - 1 < cnt) {        // rows R .. R+ahead-1 start moving towards L2
    bulk_prefetch_l2(reinterpret_cast<const float4*>(g.vec) + (size_t)ids[R + lane - 1] * g.ld4, row_bytes);
  }
  int slot0 = 0;                                   // slot of row `base`
  for (int base = 0; base < cnt; base += 4) {
    const int m = min(4, cnt - base);
    int mine = slot0 + min(team, m - 1);           // teams beyond the tail re-read the last row of the round
    if (mine >= R) mine -= R;
    mbar_wait(st.bar + mine, (st.parity >> mine) & 1u);
    const float4* row = st.ring + (size_t)mine * g.ld4;
    float2 a0 = make_float2(0.f, 0.f);
    int c = tl;
    for (; c + 3 * TEAM < g.chunks; c += 4 * TEAM) {
      float4 x[4], qq[4];
#pragma unroll
      for (int u = 0; u < 4; u++) { x[u] = row[c + u * TEAM]; qq[u] = qs[c + u * TEAM]; }
#pragma unroll
      for (int u = 0; u < 4; u++) a0 = acc4(a0, qq[u], x[u], dot);
    }
    for (; c < g.chunks; c += TEAM) a0 = acc4(a0, qs[c], row[c], dot);
    const float o = finish_metric(team_reduce(a0), g.metric);
    if (tl == 0 && team < m) d[base + team] = o;
    __syncwarp();                                  // every lane is done with this round's rows
    if (lane == 0) {                               // the freed slots take the rows `R` places further on
      for (int i = 0; i < m; i++) {
        const int j = base + R + i;
        if (j >= cnt) break;
        int sl = slot0 + i; if (sl >= R) sl -= R;
        mbar_expect_tx(st.bar + sl, row_bytes);
        bulk_g2s(st.ring + (size_t)sl * g.ld4, reinterpret_cast<const float4*>(g.vec) + (size_t)ids[j] * g.ld4, row_bytes, st.bar + sl);
      }
    } else if (lane <= m && st.ahead > 0) {        // and the look-ahead window moves on by as many rows
      const int j = base + R + st.ahead + lane - 1;
      if (j < cnt) bulk_prefetch_l2(reinterpret_cast<const float4*>(g.vec) + (size_t)ids[j] * g.ld4, row_bytes);
    }
    for (int i = 0; i < m; i++) { int sl = slot0 + i; if (sl >= R) sl -= R; st.parity ^= 1u << sl; }
    slot0 += m; if (slot0 >= R) slot0 -= R;
  }
}

// ---- exact visited set ---------------------------------------------------------------------------
// Open-addressing hash in shared memory; when it would exceed its load bound (or a probe sequence
// its reach) the set moves to a bitset over all n nodes borrowed from a global pool — exactness is
// never traded.  Two entry formats:
//   32-bit  the id + 1, home slot by multiply-shift.
//   16-bit  "quotiented": ids are first scrambled by a bijection of [0, 2^b) (odd multiplier, b = bits
//           of n), id' = q * slots + home; the entry at displacement d from `home` stores
//           1 + (q << db | d).  (home, q) determine id' and d determines home, so membership is
//           exact with half the shared memory: q < 2^b / slots needs ~10 bits at 1M rows and 1 800
//           slots, which leaves 6 bits of displacement; a probe that would need more spills.
struct HashCfg {
  uint32_t slots;      // entries (a multiple of 8); 0 = the set is a global bitset from the start
  uint32_t bytes;      // table bytes, a multiple of 16
  uint32_t bits16;     // entry format
  uint32_t mask;       // 2^b - 1
  uint32_t magic, shift;   // x / slots == __umulhi(x, magic) >> shift for every x <= mask (checked by the host)
  uint32_t db;         // displacement bits
  uint32_t mul, mul_inv;   // odd multiplier and its inverse mod 2^32
};
struct VisitedSet {
  uint32_t* tab;        // shared
  uint32_t limit;       // max entries kept in shared memory
  uint32_t count;       // warp-uniform
  uint32_t* bits;       // non-null once spilled
  int pool_slot;
};
// 32-bit entries: 1 = newly added, 0 = was present
__device__ __forceinline__ int hash_test_and_set(uint32_t* tab, const HashCfg& hc, uint32_t id) {
  const uint32_t key = id + 1u;
  uint32_t h = __umulhi(id * 2654435761u, hc.slots);          // multiply-shift range reduction
  while (true) {
    const uint32_t old = atomicCAS(&tab[h], 0u, key);
    if (old == 0u) return 1;
    if (old == key) return 0;
    h = h + 1u == hc.slots ? 0u : h + 1u;
  }
}
// The 16-bit format for a whole warp at once (all 32 lanes call; the ids of the `active` lanes are distinct).
// sm_100a has no 16-bit shared-memory CAS: `atomicCAS(unsigned short*)` is a spin loop around a 32-bit CAS,
// ~18 instructions per probe (12 % of the search kernel's instructions on the GloVe shape, ncu).  The table
// belongs to this warp alone, so the lanes only have to agree among themselves: every pending lane reads its
// slot, an empty slot is claimed with a plain 16-bit store, and after a __syncwarp the lane whose value is still
// there owns it (distinct ids give distinct entries for one slot, so the survivor is unambiguous); everybody
// else moves one slot on.  Same table contents as the CAS version up to which of two racing ids takes the
// nearer slot — membership, and so every result, is identical.
// 1 = newly added, 0 = was present, 2 = could not be placed (displacement out of reach: the set spills).
__device__ __forceinline__ int hash16_test_and_set_warp(uint32_t* tab, const HashCfg& hc, bool active, uint32_t id) {
  volatile unsigned short* t16 = reinterpret_cast<volatile unsigned short*>(tab);
  const uint32_t x = (id * hc.mul) & hc.mask;
  const uint32_t q = __umulhi(x, hc.magic) >> hc.shift;
  uint32_t h = x - q * hc.slots;
  uint32_t want = 1u + (q << hc.db);                       // the entry for displacement 0; + d as the probe moves on
  const uint32_t last = want + (1u << hc.db) - 1u;
  int res = active ? -1 : 0;                               // -1: still looking
  while (true) {
    bool claimed = false;
    if (res < 0) {
      const uint32_t old = t16[h];
      if (old == want) res = 0;
      else if (old == 0u) { t16[h] = (unsigned short)want; claimed = true; }
    }
    __syncwarp();
    if (res < 0) {
      if (claimed && t16[h] == want) res = 1;
      else if (want == last) res = 2;
      else { want++; h = h + 1u == hc.slots ? 0u : h + 1u; }
    }
    if (!__any_sync(FULL, res < 0)) break;
  }
  return res;
}
// the id stored in slot `s` (0xffffffff if the slot is empty): used when the set moves to the bitset
__device__ __forceinline__ uint32_t hash_decode(const uint32_t* tab, const HashCfg& hc, uint32_t s) {
  if (!hc.bits16) { const uint32_t key = tab[s]; return key ? key - 1u : 0xffffffffu; }
  const uint32_t e = reinterpret_cast<const unsigned short*>(tab)[s];
  if (!e) return 0xffffffffu;
  const uint32_t v = e - 1u, d = v & ((1u << hc.db) - 1u), q = v >> hc.db;
  const uint32_t home = s >= d ? s - d : s + hc.slots - d;
  return ((q * hc.slots + home) * hc.mul_inv) & hc.mask;
}
__device__ __forceinline__ void visited_clear(VisitedSet& v, const HashCfg& hc, int lane) {
  uint4* t4 = reinterpret_cast<uint4*>(v.tab);
  for (uint32_t i = lane; i < hc.bytes / 16u; i += 32) t4[i] = make_uint4(0u, 0u, 0u, 0u);
  v.count = 0;
  __syncwarp();
}

// ---- the beam: `near` as a sorted array of keys in shared memory --------------------------------
// Insert K keeping ascending order; capacity ef (the largest key falls off when full).
// Returns the key that fell off (0 if none).  fu = index below which everything is expanded.
__device__ __forceinline__ uint64_t beam_insert(uint64_t* keys, int& n, int ef, uint64_t K, int lane, int& fu) {
  uint64_t evicted = 0;
  if (n == ef) {
    evicted = keys[ef - 1];
    if (K > evicted) return K;             // only reachable when ties are accepted (Hnsw.Ba flavour)
  }
  const int last = (n == ef) ? ef - 1 : n;
  int p = last;
  for (int base = last & ~31; base >= 0; base -= 32) {
    int idx = base + lane;
    uint64_t cur = idx < n ? keys[idx] : KEY_INF;
    bool left_less = (idx == 0) || ((idx - 1 < n ? keys[idx - 1] : KEY_INF) < K);
    uint64_t left = idx > 0 && idx - 1 < n ? keys[idx - 1] : KEY_INF;
    bool keep = cur < K;
    bool write = !keep && idx <= last;
    __syncwarp();
    if (write) keys[idx] = left_less ? K : left;
    unsigned here = __ballot_sync(FULL, write && left_less);
    if (here) p = base + __ffs(here) - 1;
    __syncwarp();
    if (__shfl_sync(FULL, (int)keep, 0)) break;
  }
  if (n < ef) n++;
  if (p < fu) fu = p;
  return evicted;
}

}  // namespace hb
