// Hgraph.Stats (lib/hnsw.ml:353-375) on the device: per layer the number of nodes that have a row,
// min / max / sum of their degrees and the number of isolated nodes — one pass over the adjacency
// arrays where they live, 5 numbers per layer back to the host (instead of downloading every layer).
#pragma once
#include "common.cuh"

namespace hb {

struct LayerStats {                     // one per layer, zero-initialised except min_degree (INT_MAX)
  unsigned long long nodes, degree_sum, isolated;
  int min_degree, max_degree;
};

// one thread per (node, layer) row; a warp reduces before touching the global counters
__global__ void layer_stats_kernel(GraphView g, const int8_t* level, int num_layers, LayerStats* out) {
  const int lane = threadIdx.x & 31;
  for (int l = 0; l < num_layers; l++) {
    unsigned long long nodes = 0, sum = 0, iso = 0;
    int mn = 0x7fffffff, mx = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n; i += (int64_t)gridDim.x * blockDim.x) {
      if (level[i] < l) continue;
      const int slots = l == 0 ? g.slots0 : g.slotsU;
      const int32_t* row = l == 0 ? g.adj0 + (size_t)i * g.slots0 : g.adjU + ((size_t)g.upper_off[i] + l - 1) * g.slotsU;
      int d = 0;
      while (d < slots && row[d] >= 0) d++;
      nodes++; sum += (unsigned)d; iso += d == 0;
      mn = min(mn, d); mx = max(mx, d);
    }
    for (int o = 16; o; o >>= 1) {
      nodes += __shfl_xor_sync(FULL, nodes, o); sum += __shfl_xor_sync(FULL, sum, o); iso += __shfl_xor_sync(FULL, iso, o);
      mn = min(mn, __shfl_xor_sync(FULL, mn, o)); mx = max(mx, __shfl_xor_sync(FULL, mx, o));
    }
    if (lane == 0 && nodes) {
      atomicAdd(&out[l].nodes, nodes); atomicAdd(&out[l].degree_sum, sum); atomicAdd(&out[l].isolated, iso);
      atomicMin(&out[l].min_degree, mn); atomicMax(&out[l].max_degree, mx);
    }
  }
}

}  // namespace hb
