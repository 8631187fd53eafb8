// Exact k-NN by exhaustive scan on the 5th-generation tensor cores — the one dense contraction on
// this path (brute_force_knn_l2, benchmark/dataset.ml:15-30: all N distances per query, sort, keep k).
//
// ||q - x||^2 = ||q||^2 + ||x||^2 - 2 q.x.  The q.x part is a [nq x dim] x [dim x n] GEMM: it runs
// as tcgen05.mma (kind::f16, BF16 operands, FP32 accumulators in TMEM), with fp32 inputs split
// into bf16 hi + lo parts (q.x ~ qh.xh + ql.xh + qh.xl, three K segments of the same GEMM; one
// segment when every value is exactly a bf16, e.g. SIFT-like integer data).  Tensor-core distances
// only RANK candidates: each query keeps its TC_KP best per data split, fused into the epilogue
// (tcgen05.ld, one accumulator row per thread, running threshold), so the nq x n matrix never
// exists.  bruteforce_tc_finish_kernel then re-evaluates the candidates in exact fp32 (the same
// team-of-8 order as everything else here), sorts by (distance, id), and PROVES the result exact:
// the k-th exact distance must lie below every split's rejection bound by more than the
// tensor-core error bound, otherwise the query is flagged and recomputed by the fp32 kernel.
//
// One CTA = 128 queries (UMMA M) x its data split in tiles of 256 rows (UMMA N); operands reach
// shared memory with cp.async in the canonical K-major no-swizzle core-matrix layout
// (8 rows x 16 bytes contiguous), three stages; one thread issues the MMAs, tcgen05.commit
// releases stages and publishes the accumulator through mbarriers.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace hb {

constexpr int TC_M = 128;        // queries per CTA            (UMMA M)
constexpr int TC_N = 256;        // data rows per accumulator  (UMMA N)
constexpr int TC_KC = 64;        // bf16 elements per k-chunk: 128 bytes per row, four K=16 MMAs
constexpr int TC_STAGES = 3;
constexpr int TC_KP = 16;        // candidates kept per query per split (shorter sorted lists: the epilogue is the bottleneck)
constexpr int TC_THREADS = 256;   // 8 warps: warps w and w+4 share accumulator rows 32(w%4).., each takes half the columns
constexpr int TC_HALVES = TC_THREADS / 128;
constexpr int TC_A_BYTES = TC_M * TC_KC * 2;
constexpr int TC_B_BYTES = TC_N * TC_KC * 2;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_STAGE_CAP = 8;            // staged candidates per thread between list updates
constexpr int TC_LIST_LD = TC_KP + TC_STAGE_CAP + 1;   // one sorted list + stage per thread, odd stride (bank spread)
constexpr uint32_t TC_TMEM_COLS = 256;

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A = B = BF16 (bits 7, 10),
// both K-major (bits 15, 16 clear), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

struct TcParams {
  const __nv_bfloat16* x_hi;   // [n][kp]   kp = dim rounded up to TC_KC, zero padded
  const __nv_bfloat16* x_lo;
  const __nv_bfloat16* q_hi;   // [nq][kp]
  const __nv_bfloat16* q_lo;
  const float* x_norm;         // [n]  sum x^2 (fp32)
  int64_t n, nq;
  int kp;
  int segs;                    // 1: hi.hi only (inputs exactly bf16); 3: hi.hi + lo.hi + hi.lo
  int64_t split_len;           // data rows per gridDim.y slice (multiple of TC_N)
  uint64_t* partial;           // [gridDim.y * TC_HALVES][nq][TC_KP] keys (approximate distance, id)
  float* bound;                // [gridDim.y * TC_HALVES][nq] approximate distance below which nothing was rejected
  unsigned int* gthr;          // [nq] ordered-float: the tightest threshold any CTA has reached for the query (shared by all splits)
};

__host__ __device__ inline size_t tc_smem_bytes() {
  return (size_t)TC_STAGES * TC_STAGE_BYTES + (size_t)TC_THREADS * TC_LIST_LD * 8 + TC_N * 4 + 64;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared memory matrix descriptor (cute::UMMA::SmemDescriptor), SWIZZLE_NONE, K-major:
// start address, leading (K direction) and stride (M/N direction) byte offsets between 8x16-byte
// core matrices, all >> 4; version 1 at bit 46
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// fp32 rows -> bf16 hi / lo rows (zero padded to kp), squared norms, and whether any lo part is non-zero
__global__ void bf16_split_kernel(const float* src, int ld, int dim, int64_t n, int kp, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                  float* norm, int* any_lo, float* max_norm) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= n) return;
  float acc = 0.f;
  bool nz = false;
  for (int c = lane; c < kp; c += 32) {
    float v = c < dim ? src[row * ld + c] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi[row * kp + c] = h;
    lo[row * kp + c] = l;
    nz |= __bfloat162float(l) != 0.f;
    acc = fmaf(v, v, acc);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  if (__any_sync(FULL, nz) && lane == 0) atomicOr(any_lo, 1);
  if (lane == 0) {
    if (norm) norm[row] = acc;
    atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(acc));     // acc >= 0: int order = float order
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) bruteforce_tc_kernel(const TcParams p) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  unsigned char* stages = tc_smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(tc_smem + TC_STAGES * TC_STAGE_BYTES);
  float* xn = reinterpret_cast<float*>(lists + TC_THREADS * TC_LIST_LD);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xn + TC_N);         // [TC_STAGES] stage free, [TC_STAGES] accumulator ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + TC_STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = (warp & 3) * 32 + (tid & 31), half = warp >> 2;     // accumulator row (= TMEM lane) and column half of this thread
  const int64_t q0 = (int64_t)blockIdx.x * TC_M;
  const int64_t x_begin = (int64_t)blockIdx.y * p.split_len;
  const int64_t x_end = min(p.n, x_begin + p.split_len);
  const int kchunks = p.kp / TC_KC;             // per segment
  const int nchunks = kchunks * p.segs;         // per tile

  if (tid == 0) {
    for (int i = 0; i <= TC_STAGES; i++) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  uint64_t* my_list = lists + (size_t)tid * TC_LIST_LD;
  uint64_t* my_stage = my_list + TC_KP;
  int cnt = 0;
  float thr = __int_as_float(0x7f800000);

  const int64_t ntiles = (x_end - x_begin + TC_N - 1) / TC_N;
  const uint32_t total = (uint32_t)(ntiles * nchunks);       // k-chunks this CTA streams, tile after tile

  // stage a k-chunk: chunk index g -> (tile, segment, kc)
  // (32-bit index arithmetic: a split holds < 2^31 / TC_N tiles of at most a few dozen chunks)
  auto issue = [&](uint32_t gch) {
    const int s = (int)(gch % TC_STAGES);
    const uint32_t tile = gch / (uint32_t)nchunks;
    const int c = (int)(gch - tile * (uint32_t)nchunks), seg = c / kchunks, kc = c - seg * kchunks;
    const __nv_bfloat16* A = seg == 1 ? p.q_lo : p.q_hi;
    const __nv_bfloat16* B = seg == 2 ? p.x_lo : p.x_hi;
    const uint32_t sA = smem_u32(stages + (size_t)s * TC_STAGE_BYTES), sB = sA + TC_A_BYTES;
    const int64_t xb = x_begin + (int64_t)tile * TC_N;
#pragma unroll
    for (int i = 0; i < TC_M * 8 / TC_THREADS; i++) {
      const int idx = tid + TC_THREADS * i, row = idx >> 3, ch = idx & 7;
      const int64_t gq = min(q0 + row, p.nq - 1);
      cp_async16(sA + ch * (TC_M * 16) + (row >> 3) * 128 + (row & 7) * 16, A + gq * p.kp + kc * TC_KC + ch * 8);
    }
#pragma unroll
    for (int i = 0; i < TC_N * 8 / TC_THREADS; i++) {
      const int idx = tid + TC_THREADS * i, row = idx >> 3, ch = idx & 7;
      const int64_t gx = min(xb + row, p.n - 1);
      cp_async16(sB + ch * (TC_N * 16) + (row >> 3) * 128 + (row & 7) * 16, B + gx * p.kp + kc * TC_KC + ch * 8);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (uint32_t g = 0; g < TC_STAGES - 1; g++) {
    if (g < total) issue(g); else asm volatile("cp.async.commit_group;" ::: "memory");
  }

  int s = 0, c = 0;                                // stage and chunk-in-tile of chunk g, kept incrementally
  uint32_t tile = 0;
  for (uint32_t g = 0; g < total; g++, s = s + 1 == TC_STAGES ? 0 : s + 1) {
    // this thread's part of chunk g has landed; publish to the async proxy; everyone's part has landed
    asm volatile("cp.async.wait_group %0;" ::"n"(TC_STAGES - 2) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sA = smem_u32(stages + (size_t)s * TC_STAGE_BYTES), sB = sA + TC_A_BYTES;
#pragma unroll
      for (int j = 0; j < TC_KC / 16; j++) {
        const uint64_t ad = umma_desc(sA + j * 2 * (TC_M * 16), TC_M * 16, 128);
        const uint64_t bd = umma_desc(sB + j * 2 * (TC_N * 16), TC_N * 16, 128);
        umma_bf16(tmem, ad, bd, (c > 0 || j > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(bars + s));                         // stage s is free once these MMAs have read it
      if (c == nchunks - 1) umma_commit(smem_u32(bars + TC_STAGES));   // the tile's accumulator is complete
    }
    // refill: chunk g + STAGES - 1 goes into the stage chunk g - 1 used
    const uint32_t nx = g + TC_STAGES - 1;
    if (nx < total) {
      if (g >= 1) mbar_wait(smem_u32(bars + (s == 0 ? TC_STAGES - 1 : s - 1)), ((g - 1) / TC_STAGES) & 1u);
      issue(nx);
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }

    if (c == nchunks - 1) {
      // ---- epilogue of this tile: one accumulator row (query) per thread
      const int64_t xb = x_begin + (int64_t)tile * TC_N;
      for (int i = tid; i < TC_N; i += TC_THREADS) xn[i] = xb + i < x_end ? p.x_norm[xb + i] : __int_as_float(0x7f800000);
      // every split of a query tightens one shared threshold: a value rejected under it can never
      // be among the query's TC_KP best, whichever split it lives in (the bound stays valid because
      // thresholds only decrease)
      if (q0 + row < p.nq) thr = fminf(thr, ord2f(__ldcg(p.gthr + q0 + row)));
      mbar_wait(smem_u32(bars + TC_STAGES), (uint32_t)(tile & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncthreads();
      // Candidates that beat the thread's threshold are only STAGED while the columns stream by;
      // the sorted list is updated once per tile (and when the stage fills), so a warp pays for
      // the longest stage of its 32 threads per tile, not for every thread's every hit.
      int ns = 0;
      auto flush = [&]() {
        for (int e = 0; e < ns; e++) {
          const uint64_t key = my_stage[e];
          if (cnt == TC_KP && key >= my_list[TC_KP - 1]) continue;
          int pos = cnt < TC_KP ? cnt : TC_KP - 1;
          while (pos > 0 && my_list[pos - 1] > key) { my_list[pos] = my_list[pos - 1]; pos--; }
          my_list[pos] = key;
          if (cnt < TC_KP) cnt++;
        }
        ns = 0;
        if (cnt == TC_KP) {
          const float t16 = key_dist(my_list[TC_KP - 1]);
          if (t16 < thr) { thr = t16; if (q0 + row < p.nq) atomicMin(p.gthr + q0 + row, f2ord(t16)); }
        }
      };
#pragma unroll 1
      for (int cb = half * (TC_N / 32 / TC_HALVES); cb < (half + 1) * (TC_N / 32 / TC_HALVES); cb++) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32), r);
        const float4* xn4 = reinterpret_cast<const float4*>(xn + cb * 32);
        // fast path: does ANY of the 32 columns beat the threshold?  (two instructions per value,
        // no branch; after the first tiles the answer is almost always no)
        float dd[32];
        bool any = false;
#pragma unroll
        for (int j4 = 0; j4 < 8; j4++) {
          const float4 nv = xn4[j4];                                        // columns past the end carry +inf norms
          dd[j4 * 4 + 0] = fmaf(-2.f, __uint_as_float(r[j4 * 4 + 0]), nv.x);
          dd[j4 * 4 + 1] = fmaf(-2.f, __uint_as_float(r[j4 * 4 + 1]), nv.y);
          dd[j4 * 4 + 2] = fmaf(-2.f, __uint_as_float(r[j4 * 4 + 2]), nv.z);
          dd[j4 * 4 + 3] = fmaf(-2.f, __uint_as_float(r[j4 * 4 + 3]), nv.w);
          any |= (dd[j4 * 4 + 0] < thr) | (dd[j4 * 4 + 1] < thr) | (dd[j4 * 4 + 2] < thr) | (dd[j4 * 4 + 3] < thr);
        }
        if (any) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            if (dd[j] < thr) {
              my_stage[ns++] = make_key(dd[j], (uint32_t)(xb + cb * 32 + j));
              if (ns == TC_STAGE_CAP) flush();
            }
          }
        }
      }
      flush();
      // the next tile's first MMA overwrites the accumulator: order it after every thread's reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
    }
    if (++c == nchunks) { c = 0; tile++; }
  }

  // partial lists and the rejection bound of this split
  if (q0 + row < p.nq) {
    const size_t slot = ((size_t)blockIdx.y * TC_HALVES + half) * p.nq + (q0 + row);
    uint64_t* out = p.partial + slot * TC_KP;
    for (int j = 0; j < TC_KP; j++) out[j] = j < cnt ? my_list[j] : KEY_INF;
    p.bound[slot] = thr;                                         // +inf while fewer than TC_KP were seen
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
}

// Exact fp32 re-evaluation of every candidate, top-k by (distance, id), and the exactness proof.
// One warp per query.  Shared memory per warp: q (chunks float4), ids/dists staging, k_cap keys.
struct TcFinishParams {
  GraphView g;                 // vec = the fp32 data rows, metric L2
  const float* queries;        // [nq][ld]
  const uint64_t* partial;     // [S][nq][TC_KP]
  const float* bound;          // [S][nq]
  int S;
  int64_t nq;
  int k, k_cap, q_chunks, smem_per_warp;
  float eps;                   // tensor-core error bound, relative to ||q||^2 + max ||x||^2
  const float* max_norm;
  int32_t* ids;
  float* dists;
  int* flags;                  // [nq] 1 = not proven exact
};
__host__ __device__ inline int tc_finish_smem_per_warp(int k_cap, int q_chunks) { return q_chunks * 16 + 32 * 4 + 32 * 4 + k_cap * 8; }

__global__ void __launch_bounds__(256) bruteforce_tc_finish_kernel(const TcFinishParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= p.nq) return;
  unsigned char* my = smem_raw + (size_t)warp * p.smem_per_warp;
  float4* qs = reinterpret_cast<float4*>(my);
  uint32_t* ids = reinterpret_cast<uint32_t*>(qs + p.q_chunks);
  float* ds = reinterpret_cast<float*>(ids + 32);
  uint64_t* top = reinterpret_cast<uint64_t*>(ds + 32);
  const GraphView& g = p.g;
  const float4* qrow = reinterpret_cast<const float4*>(p.queries) + (size_t)q * g.ld4;
  float qn = 0.f;
  for (int ch = lane; ch < g.chunks; ch += 32) {
    float4 v = __ldg(qrow + ch);
    qs[ch] = v;
    qn += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int o = 16; o; o >>= 1) qn += __shfl_xor_sync(FULL, qn, o);
  __syncwarp();
  int n = 0, fu = 0;
  float min_bound = __int_as_float(0x7f800000);
  for (int s = 0; s < p.S; s++) {
    const uint64_t* list = p.partial + ((size_t)s * p.nq + q) * TC_KP;
    min_bound = fminf(min_bound, p.bound[(size_t)s * p.nq + q]);
    const uint64_t key = lane < TC_KP ? list[lane] : KEY_INF;
    const unsigned m = __ballot_sync(FULL, key != KEY_INF);
    const int cnt = __popc(m);
    if (!cnt) continue;
    if (key != KEY_INF) ids[__popc(m & ((1u << lane) - 1u))] = key_id(key);
    __syncwarp();
    batch_dist<0>(g, nullptr, qs, ids, ds, cnt, lane);
    for (int c = 0; c < cnt; c++) {
      const uint64_t K = make_key(ds[c], ids[c]);
      if (n == p.k && K > top[p.k - 1]) continue;
      beam_insert(top, n, p.k, K, lane, fu);
      __syncwarp();
    }
  }
  // every row that is not a candidate has approximate distance >= its split's bound, i.e. exact
  // distance >= ||q||^2 + bound - err; the result is exact if the k-th exact distance is below that
  const float err = p.eps * (qn + *p.max_norm);
  const bool proven = n > 0 && (n < p.k ? min_bound == __int_as_float(0x7f800000)
                                        : key_dist(top[n - 1]) + err < qn + min_bound);
  if (lane == 0) p.flags[q] = proven ? 0 : 1;
  for (int i = lane; i < p.k; i += 32) {
    int32_t oid = -1;
    float od = __int_as_float(0x7fc00000);
    if (i < n) { oid = (int32_t)key_id(top[i]); od = (float)sqrt((double)key_dist(top[i])); }
    if (p.ids) p.ids[q * p.k + i] = oid;
    p.dists[q * p.k + i] = od;
  }
}

}  // namespace hb
