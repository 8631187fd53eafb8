// Per-shard top-k merge (K6).  After the all-gather of every shard's `[nq][k]` result rows
// (shard-local ids, made global here by adding the shard's row offset; rows ascending,
// -1/NaN padded) one warp per query runs an S-way merge:
// lane s holds the head of shard s's list, the warp minimum of (distance, id) is emitted k
// times.  80 bytes per query per shard at k = 10 — latency-bound, so it is a single small
// kernel on the stream right behind the collective.
#pragma once
#include "common.cuh"

namespace hb {

struct ShardOffsets { int32_t v[32]; };   // global id = shard-local id + v[shard]

__global__ void merge_topk_kernel(const int32_t* ids, const float* dists, int S, int64_t nq, int k, int64_t shard_stride,
                                  ShardOffsets offs, int32_t* out_ids, float* out_dists) {
  int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (q >= nq) return;
  int head = 0;
  const size_t base = lane < S ? (size_t)lane * shard_stride + (size_t)q * k : 0;
  for (int j = 0; j < k; j++) {
    uint64_t key = KEY_INF;
    float myd = 0.f;
    if (lane < S && head < k) {
      int32_t id = ids[base + head];
      myd = dists[base + head];
      if (id >= 0) key = make_key(myd, (uint32_t)(id + offs.v[lane]));
    }
    uint64_t mn = key;
    for (int o = 16; o; o >>= 1) { uint64_t x = __shfl_xor_sync(FULL, mn, o); mn = x < mn ? x : mn; }
    unsigned who = __ballot_sync(FULL, key == mn && mn != KEY_INF);
    int src = who ? __ffs(who) - 1 : 0;
    float d = __shfl_sync(FULL, myd, src);
    if (who && lane == src) head++;
    if (lane == 0) {
      out_ids[q * k + j] = who ? (int32_t)key_id(mn) : -1;
      out_dists[q * k + j] = who ? d : __int_as_float(0x7fc00000);
    }
  }
}

}  // namespace hb
