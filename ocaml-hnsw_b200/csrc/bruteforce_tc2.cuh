// Exact k-NN by exhaustive scan on the 5th-generation tensor cores, warp-specialised (round 2).
//
// Same mathematics, candidate lists and exactness proof as bruteforce_tc.cuh (which documents them and
// provides the helpers and bruteforce_tc_finish_kernel); what changed is how the tile pipeline runs:
//
//   warp 0   PRODUCER  one elected lane moves operand tiles with cp.async.bulk (1-D TMA, UBLKCP): the bf16
//                      operands are written by bf16_tile_kernel in the exact core-matrix order the MMA
//                      reads from shared memory, so a whole 128-query x 64-element A block (16 KB) and a
//                      256-row B block (32 KB) are ONE bulk copy each; completion by complete_tx on the
//                      stage's `full` mbarrier; a stage is reused when its `empty` mbarrier (armed by
//                      tcgen05.commit) fires.  No thread of the CTA issues per-element copies, no
//                      __syncthreads per k-chunk.
//   warp 1   MMA       one elected lane issues tcgen05.mma (M = 128, N = 256, K = 16, bf16 -> fp32) into one
//                      of TWO 256-column accumulators (all 512 TMEM columns): tile t + 1 is multiplied while
//                      tile t is scanned.  tcgen05.commit publishes the finished accumulator (`acc_full`).
//   warps 2-9 EPILOGUE eight warps, each thread one accumulator row (= one query) and half the columns:
//                      tcgen05.ld of the NEXT 32 columns is in flight while the current 32 are scanned with
//                      packed FFMA2 + a running minimum (the common case — nothing beats the row's
//                      threshold — costs ~1.5 instructions per value); the accumulator is handed back through
//                      `acc_empty`.
#pragma once
#include "bruteforce_tc.cuh"

namespace hb {

constexpr int T2_STAGES = 3;
constexpr int T2_EPI_WARPS = 8;
constexpr int T2_THREADS = (2 + T2_EPI_WARPS) * 32;
constexpr int T2_EPI_THREADS = T2_EPI_WARPS * 32;
constexpr uint32_t T2_TMEM_COLS = 512;
constexpr int T2_STAGE_CAP = 6;              // qualifying values a thread parks between two list updates
constexpr int T2_LIST_LD = TC_KP + T2_STAGE_CAP + 1;   // sorted list + parked values per thread, odd stride (bank spread)
constexpr int T2_DD_LD = 9;                  // the 8 distances of the column group being examined, per thread, odd stride

__host__ __device__ inline size_t tc2_smem_bytes() {
  return (size_t)T2_STAGES * TC_STAGE_BYTES + (size_t)T2_EPI_THREADS * T2_LIST_LD * 8 + (size_t)T2_EPI_THREADS * T2_DD_LD * 4 +
         (size_t)T2_EPI_WARPS * 2 * (TC_N / 2) * 4 + 16 * 8 + 64;
}

// fp32 rows -> bf16 hi / lo in TILED order: tile T = row / RT, k-chunk kc = c / 64; inside the (RT x 64) block the
// shared-memory core-matrix order [c % 64 / 8][r / 8][r % 8][8 elements] (K-major, no swizzle), so a block is one
// contiguous bulk copy.  Buffers hold whole tiles and are zeroed by the caller (padding rows and columns).
__global__ void bf16_tile_kernel(const float* src, int ld, int dim, int64_t n, int kp, int RT, __nv_bfloat16* hi,
                                 __nv_bfloat16* lo, float* norm, int* any_lo, float* max_norm) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int64_t T = row / RT;
  const int r = (int)(row - T * RT), kchunks = kp / TC_KC;
  float acc = 0.f;
  bool nz = false;
  for (int c = lane; c < dim; c += 32) {
    const float v = src[row * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    const int kc = c / TC_KC, cc = c - kc * TC_KC;
    const size_t at = ((size_t)T * kchunks + kc) * ((size_t)RT * TC_KC) + (size_t)(cc >> 3) * (RT * 8) + (size_t)(r >> 3) * 64 + (r & 7) * 8 + (cc & 7);
    hi[at] = h;
    lo[at] = l;
    nz |= __bfloat162float(l) != 0.f;
    acc = fmaf(v, v, acc);
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
  if (__any_sync(FULL, nz) && lane == 0) atomicOr(any_lo, 1);
  if (lane == 0) {
    if (norm) norm[row] = acc;
    atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(acc));
  }
}

__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
// 32 columns of this thread's TMEM lane, asynchronous: the registers are valid after tmem_wait(r)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                 "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
// wait for every tcgen05.ld of this thread; the registers are operands so that no use of them moves above the wait
__device__ __forceinline__ void tmem_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

__global__ void __launch_bounds__(T2_THREADS, 1) bruteforce_tc2_kernel(const TcParams p) {
  extern __shared__ __align__(128) unsigned char tc_smem[];
  unsigned char* stages = tc_smem;
  uint64_t* lists = reinterpret_cast<uint64_t*>(tc_smem + T2_STAGES * TC_STAGE_BYTES);
  float* dds = reinterpret_cast<float*>(lists + T2_EPI_THREADS * T2_LIST_LD);           // [thread][T2_DD_LD]
  float* xn = dds + T2_EPI_THREADS * T2_DD_LD;                                          // [epilogue warp][2][TC_N / 2]: each warp's own copy
  uint64_t* bars = reinterpret_cast<uint64_t*>(xn + T2_EPI_WARPS * 2 * (TC_N / 2));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  // bars: [0..2] full (stage landed), [3..5] empty (stage read by the MMAs), [6..7] acc_full, [8..9] acc_empty
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + T2_STAGES), bar_accf = smem_u32(bars + 2 * T2_STAGES),
                 bar_acce = smem_u32(bars + 2 * T2_STAGES + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t q0 = (int64_t)blockIdx.x * TC_M;
  const int64_t x_begin = (int64_t)blockIdx.y * p.split_len;
  const int64_t x_end = min(p.n, x_begin + p.split_len);
  const int kchunks = p.kp / TC_KC, nchunks = kchunks * p.segs;
  const int64_t ntiles = (x_end - x_begin + TC_N - 1) / TC_N;
  const uint32_t total = (uint32_t)(ntiles * nchunks);

  if (tid == 0) {
    for (int i = 0; i < 2 * T2_STAGES + 2; i++) mbar_init(smem_u32(bars + i), 1);
    for (int i = 0; i < 2; i++) mbar_init(bar_acce + 8 * i, T2_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(T2_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    if (lane == 0) {
      const size_t a_block = (size_t)TC_M * TC_KC, b_block = (size_t)TC_N * TC_KC;      // elements
      const int64_t tile0 = x_begin / TC_N;
      int s = 0, c = 0;
      uint32_t ph = 0;
      int64_t tile = 0;
      for (uint32_t g = 0; g < total; g++) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u);
        const int seg = c / kchunks, kc = c - seg * kchunks;
        const __nv_bfloat16* A = (seg == 1 ? p.q_lo : p.q_hi) + ((size_t)blockIdx.x * kchunks + kc) * a_block;
        const __nv_bfloat16* B = (seg == 2 ? p.x_lo : p.x_hi) + ((size_t)(tile0 + tile) * kchunks + kc) * b_block;
        const uint32_t sA = smem_u32(stages + (size_t)s * TC_STAGE_BYTES), sB = sA + TC_A_BYTES;
        mbar_expect(bar_full + 8 * s, TC_STAGE_BYTES);
        bulk_load(sA, A, TC_A_BYTES, bar_full + 8 * s);
        bulk_load(sB, B, TC_B_BYTES, bar_full + 8 * s);
        if (++c == nchunks) { c = 0; tile++; }
        if (++s == T2_STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = 0; tile < ntiles; tile++) {
        const int a = (int)(tile & 1);
        mbar_wait(bar_acce + 8 * a, (uint32_t)((tile >> 1) & 1) ^ 1u);          // the epilogue is done with this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c = 0; c < nchunks; c++) {
          mbar_wait(bar_full + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sA = smem_u32(stages + (size_t)s * TC_STAGE_BYTES), sB = sA + TC_A_BYTES;
#pragma unroll
          for (int j = 0; j < TC_KC / 16; j++) {
            const uint64_t ad = umma_desc(sA + j * 2 * (TC_M * 16), TC_M * 16, 128);
            const uint64_t bd = umma_desc(sB + j * 2 * (TC_N * 16), TC_N * 16, 128);
            umma_bf16(tmem + (uint32_t)(a * TC_N), ad, bd, (c > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * s);
          if (c == nchunks - 1) umma_commit(bar_accf + 8 * a);
          if (++s == T2_STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: one accumulator row per thread
    const int et = tid - 64;                                   // 0..255
    const int quarter = warp & 3, half = (warp - 2) >> 2;      // TMEM lanes 32 * quarter .., columns 128 * half ..
    const int row = quarter * 32 + lane;
    uint64_t* my_list = lists + (size_t)et * T2_LIST_LD;
    float* my_dd = dds + (size_t)et * T2_DD_LD;
    uint64_t* my_stage = my_list + TC_KP;
    int cnt = 0, ns = 0;
    float thr = __int_as_float(0x7f800000);
    const bool live = q0 + row < p.nq;
    // Nothing in a tile's epilogue waits for another warp or for a global round trip: the 128 column norms a
    // warp needs are its own shared copy (one coalesced load, four per lane, under the wait for the accumulator),
    // the query's shared threshold is read one tile ahead, and its updates are fire-and-forget reductions.
    // A value below the row's threshold is first PARKED (one store); the sorted list of the TC_KP best is updated
    // for all 32 rows of the warp together, when some row's parking space is nearly full and once at the end:
    // an insertion is a serial, divergent loop, and a warp pays for the busiest of its rows each time it runs.
    auto insert = [&](uint64_t key) {
      if (cnt == TC_KP && key >= my_list[TC_KP - 1]) return;
      int pos = cnt < TC_KP ? cnt : TC_KP - 1;
      while (pos > 0 && my_list[pos - 1] > key) { my_list[pos] = my_list[pos - 1]; pos--; }
      my_list[pos] = key;
      if (cnt < TC_KP) cnt++;
      if (cnt == TC_KP) {
        const float t16 = key_dist(my_list[TC_KP - 1]);
        if (t16 < thr) {
          thr = t16;
          if (live) asm volatile("red.global.min.u32 [%0], %1;" ::"l"(p.gthr + q0 + row), "r"(f2ord(t16)) : "memory");
        }
      }
    };
    auto flush = [&]() {
      for (int e = 0; e < ns; e++) insert(my_stage[e]);
      ns = 0;
    };
    unsigned int gthr_next = live ? __ldcg(p.gthr + q0 + row) : 0xffffffffu;     // what other splits of the query have reached so far
    for (int64_t tile = 0; tile < ntiles; tile++) {
      const int a = (int)(tile & 1);
      const int64_t xb = x_begin + tile * TC_N;
      float* xna = xn + ((size_t)(warp - 2) * 2 + a) * (TC_N / 2) - half * (TC_N / 2);   // indexed by the tile's column like before
      {
        const int64_t c0 = xb + half * (TC_N / 2) + lane * 4;
        float4 nv;
        nv.x = c0 + 0 < x_end ? __ldg(p.x_norm + c0 + 0) : __int_as_float(0x7f800000);
        nv.y = c0 + 1 < x_end ? __ldg(p.x_norm + c0 + 1) : __int_as_float(0x7f800000);
        nv.z = c0 + 2 < x_end ? __ldg(p.x_norm + c0 + 2) : __int_as_float(0x7f800000);
        nv.w = c0 + 3 < x_end ? __ldg(p.x_norm + c0 + 3) : __int_as_float(0x7f800000);
        thr = fminf(thr, ord2f(gthr_next));
        if (live) gthr_next = __ldcg(p.gthr + q0 + row);                 // consumed at the next tile
        mbar_wait(bar_accf + 8 * a, (uint32_t)((tile >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        reinterpret_cast<float4*>(xna + half * (TC_N / 2))[lane] = nv;      // (this warp read the previous contents two tiles ago)
        __syncwarp();
      }
      // the 8 columns 8u .. 8u+7 of a 32-column block (u is a literal at every call)
      auto group = [&](uint32_t (&r)[32], int cb, int u) {
        unsigned m8 = 0;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
          const float dd = fmaf(-2.f, __uint_as_float(r[8 * u + jj]), xna[cb * 32 + 8 * u + jj]);
          my_dd[jj] = dd;
          m8 |= dd < thr ? (1u << jj) : 0u;
        }
#pragma unroll 1
        while (m8) {
          const int jj = __ffs(m8) - 1;
          m8 &= m8 - 1u;
          const float dd = my_dd[jj];
          if (dd < thr) {
            const uint64_t key = make_key(dd, (uint32_t)(xb + cb * 32 + 8 * u + jj));
            if (ns < T2_STAGE_CAP) my_stage[ns++] = key; else insert(key);
          }
        }
      };
      auto scan = [&](uint32_t (&r)[32], int cb) {
        const float4* xn4 = reinterpret_cast<const float4*>(xna + cb * 32);
        const float2 neg2 = make_float2(-2.f, -2.f);
        float g[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const float4 n0 = xn4[2 * u], n1 = xn4[2 * u + 1];                  // columns past the end carry +inf norms
          const float2 d0 = fma2(neg2, make_float2(__uint_as_float(r[8 * u + 0]), __uint_as_float(r[8 * u + 1])), make_float2(n0.x, n0.y));
          const float2 d1 = fma2(neg2, make_float2(__uint_as_float(r[8 * u + 2]), __uint_as_float(r[8 * u + 3])), make_float2(n0.z, n0.w));
          const float2 d2 = fma2(neg2, make_float2(__uint_as_float(r[8 * u + 4]), __uint_as_float(r[8 * u + 5])), make_float2(n1.x, n1.y));
          const float2 d3 = fma2(neg2, make_float2(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7])), make_float2(n1.z, n1.w));
          g[u] = fminf(fminf(fminf(d0.x, d0.y), fminf(d1.x, d1.y)), fminf(fminf(d2.x, d2.y), fminf(d3.x, d3.y)));
        }
        if (fminf(fminf(g[0], g[1]), fminf(g[2], g[3])) < thr) {               // some column of this block qualifies
          if (g[0] < thr) group(r, cb, 0);
          if (g[1] < thr) group(r, cb, 1);
          if (g[2] < thr) group(r, cb, 2);
          if (g[3] < thr) group(r, cb, 3);
        }
      };
      const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * TC_N + half * (TC_N / 2));
      uint32_t ra[32], rb[32];
      tmem_ld32_async(tbase, ra);
#pragma unroll 1
      for (int cb = 0; cb < TC_N / 2 / 32; cb += 2) {
        tmem_wait(ra);
        tmem_ld32_async(tbase + (uint32_t)((cb + 1) * 32), rb);
        scan(ra, half * (TC_N / 2 / 32) + cb);
        tmem_wait(rb);
        if (cb + 2 < TC_N / 2 / 32) tmem_ld32_async(tbase + (uint32_t)((cb + 2) * 32), ra);
        scan(rb, half * (TC_N / 2 / 32) + cb + 1);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(bar_acce + 8 * a);                              // this thread no longer reads accumulator a (nor xn[a])
      if (__any_sync(FULL, ns >= T2_STAGE_CAP - 1)) flush();
    }
    flush();
    if (live) {
      const size_t slot = ((size_t)blockIdx.y * TC_HALVES + half) * p.nq + (q0 + row);
      uint64_t* out = p.partial + slot * TC_KP;
      for (int j = 0; j < TC_KP; j++) out[j] = j < cnt ? my_list[j] : KEY_INF;
      p.bound[slot] = thr;                                        // +inf while fewer than TC_KP were seen
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(T2_TMEM_COLS) : "memory");
}

}  // namespace hb
