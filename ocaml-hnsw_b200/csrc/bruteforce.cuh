// Exact k-NN by exhaustive scan — replaces brute_force_knn_l2 (benchmark/dataset.ml:15-30):
// "all N distances per query, sort, keep the first k", without ever materialising the
// nq x N distance matrix.
//
// Tiled like a GEMM (64 queries x 128 data vectors per CTA step, 32 dims per stage, 4x8
// register tile per thread) but evaluated in exact fp32 on the CUDA cores as sum (q-x)^2 —
// ground truth must not carry TF32/BF16 product error — with the top-k fused in: each thread
// filters its 32 distances against the query's current k-th best, survivors go to a per-query
// candidate list in shared memory, and one warp per query folds them into a sorted k-list.
// The data set is split across gridDim.y CTAs per query block; merge_topk_kernel joins the
// partial lists.
#pragma once
#include "common.cuh"

namespace hb {

constexpr int BF_QT = 64;     // queries per CTA
constexpr int BF_XT = 128;    // data vectors per step
constexpr int BF_DK = 32;     // dims per stage
constexpr int BF_THREADS = 256;
constexpr int BF_XS_LD = BF_XT + 4;
constexpr int BF_QS_LD = BF_QT + 4;

struct BruteParams {
  const float* data;    // [n][ld]
  const float* queries; // [nq][ld]
  int64_t n, nq;
  int ld;               // floats per row (multiple of 4)
  int k, k_cap;         // k_cap = k rounded up to 32
  int metric;
  int64_t split_len;    // data rows per gridDim.y slice
  uint64_t* partial;    // [gridDim.y][nq][k] keys
};

__host__ __device__ inline size_t brute_smem_bytes(int k_cap) {
  return (size_t)BF_DK * BF_QS_LD * 4 + (size_t)BF_DK * BF_XS_LD * 4 + (size_t)BF_QT * BF_XT * 8 +
         (size_t)BF_QT * k_cap * 8 + BF_QT * 4 * 3;
}

__global__ void __launch_bounds__(BF_THREADS, 1) bruteforce_kernel(const BruteParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);                       // [DK][QS_LD]
  float* Xs = Qs + BF_DK * BF_QS_LD;                                    // [DK][XS_LD]
  uint64_t* cand = reinterpret_cast<uint64_t*>(Xs + BF_DK * BF_XS_LD);  // [QT][XT]
  uint64_t* topk = cand + BF_QT * BF_XT;                                // [QT][k_cap]
  int* cand_cnt = reinterpret_cast<int*>(topk + (size_t)BF_QT * p.k_cap);  // [QT]
  int* top_n = cand_cnt + BF_QT;                                        // [QT]
  float* thr = reinterpret_cast<float*>(top_n + BF_QT);                 // [QT]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;          // 16 x 16 threads: 4 queries x 8 vectors each
  const int64_t q0 = (int64_t)blockIdx.x * BF_QT;
  const int64_t x_begin = (int64_t)blockIdx.y * p.split_len;
  const int64_t x_end = min(p.n, x_begin + p.split_len);
  const bool dot = p.metric != 0;
  const float INF = __int_as_float(0x7f800000);

  for (int i = tid; i < BF_QT; i += BF_THREADS) { cand_cnt[i] = 0; top_n[i] = 0; thr[i] = INF; }
  __syncthreads();

  for (int64_t xb = x_begin; xb < x_end; xb += BF_XT) {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

    for (int d0 = 0; d0 < p.ld; d0 += BF_DK) {
      // stage loads: 8 lanes cover 32 consecutive floats of one row (one 128-byte line)
      {
        int r = tid >> 3, kq = tid & 7;            // 32 rows per pass
#pragma unroll
        for (int pass = 0; pass < BF_XT / 32; pass++) {
          int row = r + pass * 32;
          int64_t gx = xb + row;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gx < x_end && d0 + kq * 4 < p.ld) v = __ldg(reinterpret_cast<const float4*>(p.data + gx * p.ld + d0) + kq);
          Xs[(kq * 4 + 0) * BF_XS_LD + row] = v.x; Xs[(kq * 4 + 1) * BF_XS_LD + row] = v.y;
          Xs[(kq * 4 + 2) * BF_XS_LD + row] = v.z; Xs[(kq * 4 + 3) * BF_XS_LD + row] = v.w;
        }
#pragma unroll
        for (int pass = 0; pass < BF_QT / 32; pass++) {
          int row = r + pass * 32;
          int64_t gq = q0 + row;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gq < p.nq && d0 + kq * 4 < p.ld) v = __ldg(reinterpret_cast<const float4*>(p.queries + gq * p.ld + d0) + kq);
          Qs[(kq * 4 + 0) * BF_QS_LD + row] = v.x; Qs[(kq * 4 + 1) * BF_QS_LD + row] = v.y;
          Qs[(kq * 4 + 2) * BF_QS_LD + row] = v.z; Qs[(kq * 4 + 3) * BF_QS_LD + row] = v.w;
        }
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < BF_DK; kk++) {
        float4 qv = *reinterpret_cast<const float4*>(&Qs[kk * BF_QS_LD + ty * 4]);
        float4 xa = *reinterpret_cast<const float4*>(&Xs[kk * BF_XS_LD + tx * 8]);
        float4 xc = *reinterpret_cast<const float4*>(&Xs[kk * BF_XS_LD + tx * 8 + 4]);
        float qa[4] = {qv.x, qv.y, qv.z, qv.w};
        float xv[8] = {xa.x, xa.y, xa.z, xa.w, xc.x, xc.y, xc.z, xc.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (dot) acc[i][j] = fmaf(qa[i], xv[j], acc[i][j]);
            else { float t = qa[i] - xv[j]; acc[i][j] = fmaf(t, t, acc[i][j]); }
          }
      }
      __syncthreads();
    }

    // filter against the current k-th best of each query
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int q = ty * 4 + i;
      float t = thr[q];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        int64_t gx = xb + tx * 8 + j;
        float d = finish_metric(acc[i][j], p.metric);
        if (gx < x_end && q0 + q < p.nq && (d < t || (d == t))) {   // equal distance may still win on id
          int slot = atomicAdd(&cand_cnt[q], 1);
          cand[q * BF_XT + slot] = make_key(d, (uint32_t)gx);
        }
      }
    }
    __syncthreads();
    // fold candidates: warp w owns queries w*8 .. w*8+7
    for (int qi = 0; qi < BF_QT / 8; qi++) {
      int q = warp * (BF_QT / 8) + qi;
      int cnt = cand_cnt[q];
      if (cnt == 0) continue;
      uint64_t* list = topk + (size_t)q * p.k_cap;
      int n = top_n[q];
      int fu = 0;
      for (int c = 0; c < cnt; c++) {
        uint64_t K = cand[q * BF_XT + c];
        if (n == p.k && K > list[p.k - 1]) continue;
        beam_insert(list, n, p.k, K, lane, fu);
        __syncwarp();
      }
      if (lane == 0) {
        top_n[q] = n; cand_cnt[q] = 0;
        if (n == p.k) thr[q] = key_dist(list[p.k - 1]);
      }
    }
    __syncthreads();
  }

  // write the partial lists
  for (int i = tid; i < BF_QT * p.k; i += BF_THREADS) {
    int q = i / p.k, j = i % p.k;
    if (q0 + q < p.nq) {
      uint64_t key = j < top_n[q] ? topk[(size_t)q * p.k_cap + j] : KEY_INF;
      p.partial[((size_t)blockIdx.y * p.nq + (q0 + q)) * p.k + j] = key;
    }
  }
}

// partial keys [S][nq][k] -> ids/dists [nq][k]; L2 distances become sqrt_double (dataset.ml:23
// goes through Hnsw.EuclideanBa.distance = sqrt of the fp32 sum, lib/hnsw.ml:814).
__global__ void bruteforce_finish_kernel(const uint64_t* partial, int S, int64_t nq, int k, int metric,
                                         int32_t* ids, float* dists) {
  int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (q >= nq) return;
  // S-way merge of sorted lists, one list head per lane (S <= 32)
  int head = 0;
  const uint64_t* mine = lane < S ? partial + ((size_t)lane * nq + q) * k : nullptr;
  for (int j = 0; j < k; j++) {
    uint64_t key = (mine && head < k) ? mine[head] : KEY_INF;
    uint64_t mn = key;
    for (int o = 16; o; o >>= 1) { uint64_t x = __shfl_xor_sync(FULL, mn, o); mn = x < mn ? x : mn; }
    unsigned who = __ballot_sync(FULL, key == mn && mn != KEY_INF);
    if (who && lane == __ffs(who) - 1) head++;
    if (lane == 0) {
      if (mn == KEY_INF) { ids[q * k + j] = -1; dists[q * k + j] = __int_as_float(0x7fc00000); }
      else {
        float d = key_dist(mn);
        ids[q * k + j] = (int32_t)key_id(mn);
        dists[q * k + j] = metric == 0 ? (float)sqrt((double)d) : d;
      }
    }
  }
}

}  // namespace hb
