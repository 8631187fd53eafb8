// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of lehy/ocaml-hnsw path B (lib/ohnsw.ml), used as the checker for the CUDA
// path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may link, load or execute anything under oracle/.  The product (libhnsw_b200.so) never
// routes through this code.
//
// Parity status:
//   * ALGORITHM: pinned.  Every inline golden vector the reference holds for this path
//     (ohnsw.ml:514-534 search_one, :593-644 search_k, :665-764 select_neighbours, the
//     container tests :85-108 :138-156 :204-251 :270-296 :353-400) is reproduced by
//     tests/test_oracle_golden.py through the Abs1D space below (distance = |a-b| on OCaml
//     floats = doubles, exactly Hgraph.Test.distance, ohnsw.ml:361).
//   * ARITHMETIC of distance_l2 (ohnsw.ml:899 -> Lacaml.S.Vec.ssqr_diff): PARITY UNPINNED.
//     Lacaml is a third-party dependency that is not vendored and not version-pinned
//     (hnsw.opam is empty; lib/jbuild names `lacaml`), and no reference test pins its
//     output.  Published behaviour: fp32 accumulation of (x-y)^2, summation order is
//     build-dependent (-O3 -ffast-math vectorises it), result widened to double, Float.sqrt
//     in double.  Two summation orders are offered here:
//        SUM_SEQUENTIAL : one fp32 accumulator, index order, mul+add   (a scalar Lacaml build)
//        SUM_TEAM8      : eight accumulator PAIRS; pair t takes the 4-float chunks t, t+8,
//                         t+16.. in index order, components x and z of each chunk going to
//                         the first accumulator of the pair, y and w to the second, each with
//                         a fused multiply-add; the pair is added, then a butterfly
//                         (t ^ 4, t ^ 2, t ^ 1).  This is one legal vectorised order, and it is
//                         the order the CUDA kernels use (8 lanes per vector, one float4 per
//                         lane per step, two packed f32x2 FMAs per float4), which makes
//                         GPU-vs-oracle distances bit-identical and the id comparison exact.
//   * TIE ORDER: Core_kernel.Heap (unpinned third party) leaves the order among equal keys
//     unspecified.  The oracle orders heap elements by the total order (distance, node id);
//     accept/stop rules still compare distances only, exactly as the reference does.
//
// Everything follows lib/ohnsw.ml; the line each function restates is cited beside it.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <set>
#include <stdexcept>
#include <vector>

namespace oracle {

struct Counters {
  uint64_t n_dist = 0;   // distance closure calls (hnsw.ml:732-751 counts the same thing)
  uint64_t n_exp0 = 0;   // adjacency rows read on layer 0 (search_k expansions)
  uint64_t n_expU = 0;   // adjacency rows read on layers >= 1 (greedy scans + search_k)
};

// ohnsw.ml:6-12  HeapElt = { node; distance }.  distance is an OCaml float = double.
struct HeapElt {
  int32_t node;
  double distance;
};
// compare_nearest / compare_farthest (ohnsw.ml:10-11) extended to a total order by node id.
struct NearestFirst {  // for std::priority_queue: top() = smallest (distance, node)
  bool operator()(const HeapElt& a, const HeapElt& b) const {
    return a.distance > b.distance || (a.distance == b.distance && a.node > b.node);
  }
};
struct FarthestFirst {  // top() = largest (distance, node)
  bool operator()(const HeapElt& a, const HeapElt& b) const {
    return a.distance < b.distance || (a.distance == b.distance && a.node < b.node);
  }
};
using MinHeap = std::priority_queue<HeapElt, std::vector<HeapElt>, NearestFirst>;
using MaxHeap = std::priority_queue<HeapElt, std::vector<HeapElt>, FarthestFirst>;

// ohnsw.ml:111-136.  An ordered id list; index 0 is the list head.
struct Neighbours {
  std::vector<int32_t> list;
  void add(int32_t node) { list.insert(list.begin(), node); }          // :116-118 prepend
  void remove(int32_t node) {                                          // :119-124
    std::vector<int32_t> ret;                                          // rebuilt by prepending,
    for (int32_t e : list)                                             // i.e. order reversed
      if (e != node) ret.insert(ret.begin(), e);
    list.swap(ret);
  }
  size_t length() const { return list.size(); }
  bool mem(int32_t a) const { return std::find(list.begin(), list.end(), a) != list.end(); }
};

// ohnsw.ml:163-202
struct Graph {
  std::vector<Neighbours> v;
  Graph() {}
  explicit Graph(size_t n) : v(n) {}
  size_t num_nodes() const { return v.size(); }
  Neighbours& adjacent(int32_t node) { return v.at(node); }
  const Neighbours& adjacent(int32_t node) const { return v.at(node); }
  void add_node() { v.emplace_back(); }

  // :182-196.  diff_both (:129-134) builds Int sets; Set.iter walks them in ascending id.
  void set_connections(int32_t node, const Neighbours& neighbours) {
    const Neighbours& old = v.at(node);
    std::set<int32_t> sa(old.list.begin(), old.list.end());
    std::set<int32_t> sb(neighbours.list.begin(), neighbours.list.end());
    std::vector<int32_t> added, removed;
    for (int32_t x : sb) if (!sa.count(x)) added.push_back(x);
    for (int32_t x : sa) if (!sb.count(x)) removed.push_back(x);
    v.at(node) = neighbours;                                   // 1.
    for (int32_t r : removed) v.at(r).remove(node);            // 2.
    for (int32_t a : added) v.at(a).add(node);                 // 3.
  }
  // :198-202
  void set_connections_for_new_node(int32_t node, const Neighbours& neighbours) {
    v.at(node) = neighbours;
    for (int32_t a : neighbours.list) v.at(a).add(node);
  }
  // Graph.Test.invariant :217-225
  bool invariant() const {
    for (size_t i = 0; i < v.size(); i++)
      for (int32_t nb : v[i].list)
        if (!v.at(nb).mem((int32_t)i)) return false;
    return true;
  }
};

// ohnsw.ml:256-268.  `max_epoch` is OCaml's Int.max_value (2^62-1); it is a member only so
// the wrap test (:285-295) can be driven.
struct Visited {
  std::vector<int64_t> visited;
  int64_t epoch = 1;
  static constexpr int64_t kIntMax = (int64_t(1) << 62) - 1;
  Visited() {}
  explicit Visited(size_t n) : visited(n, 0) {}
  bool mem(int32_t node) const { return visited.at(node) >= epoch; }
  void add(int32_t node) { visited.at(node) = epoch; }
  size_t card() const {
    size_t c = 0;
    for (int64_t e : visited) if (e >= epoch) c++;
    return c;
  }
  void clear() {
    if (epoch < kIntMax - 1) epoch++;
    else { std::fill(visited.begin(), visited.end(), 0); epoch = 1; }
  }
  void grow(size_t n) { if (visited.size() < n) visited.resize(n, 0); }
};

// ---------------------------------------------------------------------------------------------
// Spaces: the 'a distance / 'a value pair (ohnsw.ml:3-4).

enum SumOrder { SUM_SEQUENTIAL = 0, SUM_TEAM8 = 1 };
enum Metric { METRIC_L2 = 0, METRIC_ANGULAR = 1, METRIC_IP = 2 };

inline float ssqr_diff_sequential(const float* a, const float* b, int d) {
  float acc = 0.f;
  for (int i = 0; i < d; i++) { float x = a[i] - b[i]; x = x * x; acc = acc + x; }
  return acc;
}
inline float team8_butterfly(float* p) {
  for (int m = 4; m >= 1; m >>= 1) {
    float q[8];
    for (int t = 0; t < 8; t++) q[t] = p[t] + p[t ^ m];
    for (int t = 0; t < 8; t++) p[t] = q[t];
  }
  return p[0];
}
// Scalar statement of the TEAM8 order — this is the definition.
inline float ssqr_diff_team8_scalar(const float* a, const float* b, int d) {
  float pa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, p[8];
  for (int i = 0; i < d; i++) {
    int t = (i >> 2) & 7;
    float x = a[i] - b[i];
    if (i & 1) pb[t] = fmaf(x, x, pb[t]); else pa[t] = fmaf(x, x, pa[t]);
  }
  for (int t = 0; t < 8; t++) p[t] = pa[t] + pb[t];
  return team8_butterfly(p);
}
#if defined(__AVX2__) && defined(__FMA__)
}  // namespace oracle
#include <immintrin.h>
namespace oracle {
// Same arithmetic, the eight first accumulators in one register and the eight second ones in
// another.  A block of 32 floats is 8 chunks; a 4x4 transpose inside each 128-bit half gathers
// component j of the 8 chunks into one register (accumulator order 0,2,4,6,1,3,5,7 inside the
// register).  Every accumulator still sees its chunks in index order, components x,z (first)
// or y,w (second), one fused multiply-add each, so the result is bit-identical to the scalar
// statement (tests/test_oracle_arith.py checks it).
template <bool kDot>
inline float team8_avx2(const float* a, const float* b, int d) {
  __m256 PA = _mm256_setzero_ps(), PB = _mm256_setzero_ps();
  int i = 0;
  for (; i + 32 <= d; i += 32) {
    __m256 a0 = _mm256_loadu_ps(a + i), a1 = _mm256_loadu_ps(a + i + 8), a2 = _mm256_loadu_ps(a + i + 16), a3 = _mm256_loadu_ps(a + i + 24);
    __m256 b0 = _mm256_loadu_ps(b + i), b1 = _mm256_loadu_ps(b + i + 8), b2 = _mm256_loadu_ps(b + i + 16), b3 = _mm256_loadu_ps(b + i + 24);
    if (!kDot) { a0 = _mm256_sub_ps(a0, b0); a1 = _mm256_sub_ps(a1, b1); a2 = _mm256_sub_ps(a2, b2); a3 = _mm256_sub_ps(a3, b3); }
#define ORC_T4(v0, v1, v2, v3, X, Y, Z, W)                                       \
    __m256 X, Y, Z, W;                                                           \
    {                                                                            \
      __m256 t0 = _mm256_unpacklo_ps(v0, v1), t1 = _mm256_unpackhi_ps(v0, v1);   \
      __m256 t2 = _mm256_unpacklo_ps(v2, v3), t3 = _mm256_unpackhi_ps(v2, v3);   \
      X = _mm256_shuffle_ps(t0, t2, 0x44); Y = _mm256_shuffle_ps(t0, t2, 0xEE);  \
      Z = _mm256_shuffle_ps(t1, t3, 0x44); W = _mm256_shuffle_ps(t1, t3, 0xEE);  \
    }
    ORC_T4(a0, a1, a2, a3, ax, ay, az, aw)
    if (kDot) {
      ORC_T4(b0, b1, b2, b3, bx, by, bz, bw)
      PA = _mm256_fmadd_ps(ax, bx, PA); PB = _mm256_fmadd_ps(ay, by, PB);
      PA = _mm256_fmadd_ps(az, bz, PA); PB = _mm256_fmadd_ps(aw, bw, PB);
    } else {
      PA = _mm256_fmadd_ps(ax, ax, PA); PB = _mm256_fmadd_ps(ay, ay, PB);
      PA = _mm256_fmadd_ps(az, az, PA); PB = _mm256_fmadd_ps(aw, aw, PB);
    }
#undef ORC_T4
  }
  float ra[8], rb[8], pa[8], pb[8], p[8];
  _mm256_storeu_ps(ra, PA);
  _mm256_storeu_ps(rb, PB);
  // register slot s holds accumulator: low half chunks 0,2,4,6 ; high half chunks 1,3,5,7
  static const int slot[8] = {0, 2, 4, 6, 1, 3, 5, 7};
  for (int s = 0; s < 8; s++) { pa[slot[s]] = ra[s]; pb[slot[s]] = rb[s]; }
  for (; i < d; i++) {
    int t = (i >> 2) & 7;
    float* acc = (i & 1) ? pb : pa;
    if (kDot) acc[t] = fmaf(a[i], b[i], acc[t]);
    else { float x = a[i] - b[i]; acc[t] = fmaf(x, x, acc[t]); }
  }
  for (int t = 0; t < 8; t++) p[t] = pa[t] + pb[t];
  return team8_butterfly(p);
}
inline float ssqr_diff_team8(const float* a, const float* b, int d) { return team8_avx2<false>(a, b, d); }
#else
inline float ssqr_diff_team8(const float* a, const float* b, int d) { return ssqr_diff_team8_scalar(a, b, d); }
#endif
inline float dot_sequential(const float* a, const float* b, int d) {
  float acc = 0.f;
  for (int i = 0; i < d; i++) { float x = a[i] * b[i]; acc = acc + x; }
  return acc;
}
inline float dot_team8_scalar(const float* a, const float* b, int d) {
  float pa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, p[8];
  for (int i = 0; i < d; i++) {
    int t = (i >> 2) & 7;
    if (i & 1) pb[t] = fmaf(a[i], b[i], pb[t]); else pa[t] = fmaf(a[i], b[i], pa[t]);
  }
  for (int t = 0; t < 8; t++) p[t] = pa[t] + pb[t];
  return team8_butterfly(p);
}
#if defined(__AVX2__) && defined(__FMA__)
inline float dot_team8(const float* a, const float* b, int d) { return team8_avx2<true>(a, b, d); }
#else
inline float dot_team8(const float* a, const float* b, int d) { return dot_team8_scalar(a, b, d); }
#endif

// fp32 vectors stored row-major [n][dim] (= Lacaml.S.mat D x N, one vector per column).
struct VecSpace {
  using Target = const float*;
  int dim = 0;
  int metric = METRIC_L2;
  int order = SUM_TEAM8;
  std::vector<float> data;  // owned copy: `value i = Mat.col batch (i+1)` keeps the batch alive (:842)
  const float* value(int32_t node) const { return data.data() + (size_t)node * dim; }
  // The "work" value every comparison is made on.  For L2 this is the fp32 sum of squares:
  // ohnsw.ml:899 compares sqrt_double((double)s); sqrt is strictly monotonic and injective on
  // fp32-valued doubles, so ordering s is ordering the reference's distance.  report() gives
  // the value the reference would hand back.
  double distance(const float* a, const float* b) const {
    if (metric == METRIC_L2) {
      float s = order == SUM_TEAM8 ? ssqr_diff_team8(a, b, dim) : ssqr_diff_sequential(a, b, dim);
      return std::sqrt((double)s);                      // Float.sqrt @@ ssqr_diff a b
    }
    float dt = order == SUM_TEAM8 ? dot_team8(a, b, dim) : dot_sequential(a, b, dim);
    if (metric == METRIC_ANGULAR) return (double)(1.0f - dt);
    return (double)(-dt);
  }
};

// Hgraph.Test.distance / value (ohnsw.ml:361-362): OCaml floats, distance |a-b|.
struct Abs1DSpace {
  using Target = double;
  std::vector<double> data;
  double value(int32_t node) const { return data.at(node); }
  double distance(double a, double b) const { return std::fabs(a - b); }
};

// ---------------------------------------------------------------------------------------------
// ohnsw.ml:306-351 Hgraph + the algorithms.

template <class Space>
struct Hnsw {
  using Target = typename Space::Target;
  Space space;
  std::vector<Graph> layers;         // Hgraph.create pushes one empty layer (:323)
  bool has_entry = false;
  int32_t entry_point = -1;
  Counters counters;                 // charged by insert(); queries charge the Counters they are given
  std::vector<int32_t> levels;       // the level drawn for each insert (bookkeeping for export)
  uint64_t rng_state = 0x9E3779B97F4A7C15ull;
  // Hnsw.Ba's Nearest.insert_distance (lib/hnsw.ml:494-506) accepts a candidate whose distance EQUALS
  // the current maximum; path B (ohnsw.ml:574) does not.  Off = path B, the default.
  bool accept_ties = false;
  // The other Hnsw.Ba build parameters the GPU's HNSW_BA flavour follows (lib/hnsw.ml:753-758,
  // hnsw_algo.ml:596-599, 661-667): M links for a new node on layer 0 too, and a candidate set no
  // larger than that is kept whole.  Pruning stays path B's (w.r.t. the pruned node): path A's
  // prune-by-distance-to-the-inserted-point (Q6) and do_not_isolate are not restated.
  bool ba_build = false;

  Hnsw() { layers.emplace_back(); }

  // Every algorithm takes the Counters it charges, so query-parallel callers can keep one per thread.
  double dist(Target a, Target b, Counters& c) const { c.n_dist++; return space.distance(a, b); }
  Target value(int32_t n) const { return space.value(n); }

  int max_layer() const { return (int)layers.size() - 1; }              // :346
  size_t num_nodes() const { return layers[0].num_nodes(); }            // :335
  bool has_node(int64_t n) const { return n >= 0 && (size_t)n < num_nodes(); }
  int32_t add_node() {                                                  // :328-330
    for (Graph& g : layers) g.add_node();
    return (int32_t)layers[0].num_nodes() - 1;
  }
  void set_entry_point(int64_t n) {                                     // :341-344
    if (!has_node(n)) throw std::invalid_argument("Hgraph.set_entry_point: invalid node");
    has_entry = true; entry_point = (int32_t)n;
  }
  void set_max_layer(int n) {                                           // :347-351
    size_t nn = layers.back().num_nodes();
    for (int i = max_layer() + 1; i <= n; i++) layers.emplace_back(nn);
  }
  bool invariant() const {                                              // Hgraph.Test.invariant :354-359
    for (const Graph& g : layers) if (!g.invariant()) return false;
    for (const Graph& g : layers) if (g.num_nodes() != layers[0].num_nodes()) return false;
    return !has_entry || has_node(entry_point);
  }

  // HeapElt.create (:8-9): distance target (value node)
  HeapElt element(Target target, int32_t node, Counters& c) const { return HeapElt{node, dist(target, value(node), c)}; }

  // search_one_simple (:492-508); search_one = search_one_simple (:512)
  int32_t search_one(int layer, int32_t start_node, Target target, Counters& c) const {
    const Graph& graph = layers.at(layer);
    bool changed = true;
    int32_t best_node = start_node;
    double best_distance = dist(value(start_node), target, c);
    while (changed) {
      changed = false;
      const std::vector<int32_t> neighbours = graph.adjacent(best_node).list;  // captured before the scan
      if (layer == 0) c.n_exp0++; else c.n_expU++;
      for (int32_t nb : neighbours) {
        double d = dist(value(nb), target, c);
        if (d < best_distance) { best_node = nb; best_distance = d; changed = true; }
      }
    }
    return best_node;
  }

  // search_k (:543-588).  start_nodes is consumed (it becomes visit_me).  Returns the
  // nearest set as a min-queue.
  MinHeap search_k(int layer, Visited& visited, MinHeap& start_nodes, int k, Target target, Counters& c) const {
    const Graph& graph = layers.at(layer);
    visited.clear();                                                    // :553
    MaxHeap nearest_maxq;                                               // :554 (cleared)
    {
      MinHeap copy = start_nodes;                                       // MinQueue.iter (:555-557)
      while (!copy.empty()) {
        visited.add(copy.top().node);
        nearest_maxq.push(copy.top());
        copy.pop();
      }
    }
    MinHeap& visit_me = start_nodes;                                    // :559
    while (!visit_me.empty()) {                                         // aux (:564-581)
      HeapElt cur = visit_me.top(); visit_me.pop();
      if (cur.distance > nearest_maxq.top().distance) break;            // :568
      const std::vector<int32_t>& adj = graph.adjacent(cur.node).list;
      if (layer == 0) c.n_exp0++; else c.n_expU++;
      for (int32_t e : adj) {                                           // :570
        if (!visited.mem(e)) {
          visited.add(e);
          HeapElt he = element(target, e, c);                            // :573
          if ((int)nearest_maxq.size() < k ||
              (accept_ties ? he.distance <= nearest_maxq.top().distance : he.distance < nearest_maxq.top().distance)) {  // :574 / hnsw.ml:494-506
            visit_me.push(he);
            nearest_maxq.push(he);
            if ((int)nearest_maxq.size() > k) nearest_maxq.pop();       // :577
          }
        }
      }
    }
    MinHeap result;                                                     // :586-588
    while (!nearest_maxq.empty()) { result.push(nearest_maxq.top()); nearest_maxq.pop(); }
    return result;
  }

  // select_neighbours (:647-663).  Consumes the candidate min-queue.
  Neighbours select_neighbours(MinHeap& possible, int num_neighbours, Counters& c, bool keep_all = false) const {
    Neighbours selected;
    if (keep_all && (int)possible.size() <= num_neighbours) {            // hnsw_algo.ml:596-599
      while (!possible.empty()) { selected.add(possible.top().node); possible.pop(); }
      return selected;
    }
    while (!possible.empty()) {
      HeapElt e = possible.top(); possible.pop();
      bool all = true;                                                  // Neighbours.for_all, head first
      for (int32_t nb : selected.list) {
        if (!(e.distance < dist(value(nb), value(e.node), c))) { all = false; break; }
      }
      if (all) selected.add(e.node);
      if (!((int)selected.length() < num_neighbours)) break;           // :660
    }
    return selected;
  }

  // -ln(U)*mL rounded to nearest (ohnsw.ml:781; Base Float.round_nearest = floor(x + 0.5)).
  // The reference draws U from the global OCaml Random state, which cannot be reproduced
  // here; callers inject the level, or this splitmix64 stream is used.
  int draw_level(double level_mult) {
    rng_state += 0x9E3779B97F4A7C15ull;
    uint64_t z = rng_state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    double u = ((double)(z >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
    return (int)std::floor(-std::log(u) * level_mult + 0.5);
  }

  // insert (:766-837).  `level` < 0 draws from draw_level.
  void insert(Target target, int num_connections, int num_nodes_search_construction,
              double level_mult, Visited& visited, int level_in) {
    int32_t new_node = add_node();                                      // :771
    visited.grow(num_nodes());
    if (!has_entry) {                                                   // :774-778
      set_entry_point(new_node);
      set_max_layer(0);
      levels.push_back(0);
      return;
    }
    visited.clear();                                                    // :780
    int level = level_in >= 0 ? level_in : draw_level(level_mult);     // :781
    levels.push_back(level);
    int32_t node = entry_point;
    for (int layer = max_layer(); layer >= level + 1; layer--)          // :785-789
      node = search_one(layer, node, target, counters);

    MinHeap w_queue;                                                    // :801-802
    w_queue.push(element(target, node, counters));

    for (int layer = std::min(level, max_layer()); layer >= 0; layer--) {   // :806
      Graph& graph = layers.at(layer);
      MinHeap nearest = search_k(layer, visited, w_queue, num_nodes_search_construction, target, counters);  // :811
      w_queue = nearest;                                                // :814-816 swap
      int nc = layer == 0 ? 2 * num_connections : num_connections;      // :818
      int n_new = ba_build ? num_connections : nc;                      // lib/hnsw.ml:753-758
      MinHeap copy = w_queue;                                           // MinQueue.copy (:819)
      Neighbours neighbours = select_neighbours(copy, n_new, counters, ba_build);
      graph.set_connections_for_new_node(new_node, neighbours);         // :820
      const std::vector<int32_t> iter_list = neighbours.list;           // List.iter holds the old immutable list
      for (int32_t neighbour : iter_list) {                             // :821-829
        const Neighbours& nn = graph.adjacent(neighbour);
        if ((int)nn.length() > nc) {
          MinHeap neighbour_queue;                                      // min_queue_of_neighbours (:791-798)
          Target base = value(neighbour);
          for (int32_t x : nn.list) neighbour_queue.push(element(base, x, counters));
          Neighbours reduced = select_neighbours(neighbour_queue, nc, counters);
          graph.set_connections(neighbour, reduced);                    // :828
        }
      }
    }
    if (level > max_layer()) {                                          // :832-836
      set_max_layer(level);
      set_entry_point(new_node);
    }
  }

  // knn (:859-875): returns the result min-queue (<= k elements).
  MinHeap knn(Visited& visited, int k, Target target, Counters& c) const {
    if (!has_entry) throw std::invalid_argument("knn: empty hgraph");    // :862
    int32_t node = entry_point;
    for (int layer = max_layer(); layer >= 1; layer--) node = search_one(layer, node, target, c);  // :865-867
    MinHeap w_queue;
    w_queue.push(element(target, node, c));                             // :871
    return search_k(0, visited, w_queue, k, target, c);                 // :873
  }
};

using VecHnsw = Hnsw<VecSpace>;
using AbsHnsw = Hnsw<Abs1DSpace>;

}  // namespace oracle
