// ORACLE — TEST INFRASTRUCTURE ONLY (see ohnsw_oracle.hpp).  C ABI over the restatement so
// tests/ and bench.py's CPU-baseline legs can drive it through ctypes.  Nothing in the product
// library links or loads this file.
#include "ohnsw_oracle.hpp"

#include <chrono>
#include <cstdio>
#include <string>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

namespace {
thread_local std::string g_err;
struct VecHandle {
  VecHnsw h;
  Visited visited;   // one shared Visited for a whole build, as in ohnsw.ml:845
};
struct AbsHandle {
  AbsHnsw h;
  Visited visited;
};
template <class F>
int guard(F&& f) {
  try { f(); return 0; }
  catch (const std::invalid_argument& e) { g_err = e.what(); return 1; }
  catch (const std::out_of_range& e) { g_err = std::string("index out of range: ") + e.what(); return 1; }
  catch (const std::bad_alloc&) { g_err = "out of memory"; return 3; }
  catch (const std::exception& e) { g_err = e.what(); return 2; }
}
// knn_batch_bigarray (ohnsw.ml:877-897) for one query: pop ascending into row i of ids / dists.
// `ef` is the beam (the reference's ~k, Q4); the first k popped are kept.
void one_query(const VecHnsw& h, Visited& visited, const float* q, int k, int ef, int32_t* ids,
               float* dists, Counters& c) {
  for (int i = 0; i < k; i++) { ids[i] = -1; dists[i] = std::numeric_limits<float>::quiet_NaN(); }
  MinHeap nearest = h.knn(visited, ef, q, c);
  int i = 0;
  while (!nearest.empty() && i < k) {
    dists[i] = (float)nearest.top().distance;   // distances.{i,j} <- e.distance : fp32 store
    ids[i] = nearest.top().node;
    nearest.pop();
    i++;
  }
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- fp32 vector space -----------------------------------------------------------------------
void* orc_vec_create(int dim, int metric, int order) {
  VecHandle* v = new VecHandle();
  v->h.space.dim = dim; v->h.space.metric = metric; v->h.space.order = order;
  return v;
}
void orc_vec_destroy(void* p) { delete (VecHandle*)p; }
// 1: the Hnsw.Ba acceptance rule (ties with the current maximum are accepted, lib/hnsw.ml:494-506)
void orc_vec_set_accept_ties(void* p, int on) { ((VecHandle*)p)->h.accept_ties = on != 0; }
// 1: the Hnsw.Ba build parameters (M links for a new node on every layer, small candidate sets kept whole)
void orc_vec_set_ba_build(void* p, int on) { ((VecHandle*)p)->h.ba_build = on != 0; }

// build_batch_bigarray (ohnsw.ml:840-857): n sequential inserts in row order.  May be called
// again to keep inserting (the reference's `insert`, :766).  levels may be null.
int orc_vec_build(void* p, const float* data, int64_t n, int M, int efC, const int32_t* levels) {
  VecHandle* v = (VecHandle*)p;
  return guard([&] {
    if (M < 2) throw std::invalid_argument("num_connections must be >= 2 (level_mult = 1/ln M)");
    VecHnsw& h = v->h;
    int dim = h.space.dim;
    size_t base = h.num_nodes();
    h.space.data.resize((base + n) * (size_t)dim);
    std::memcpy(h.space.data.data() + base * dim, data, sizeof(float) * n * dim);
    double level_mult = 1.0 / std::log((double)M);                      // :844
    v->visited.grow(base + n);
    for (int64_t i = 0; i < n; i++)
      h.insert(h.space.value((int32_t)(base + i)), M, efC, level_mult, v->visited, levels ? levels[i] : -1);
  });
}

int orc_vec_search(void* p, const float* q, int64_t nq, int k, int ef, int32_t* ids, float* dists,
                   uint64_t* counters /* [nq][3] or null */) {
  VecHandle* v = (VecHandle*)p;
  return guard([&] {
    const VecHnsw& h = v->h;
    if (ef < k) throw std::invalid_argument("ef must be >= k");
    Visited visited(h.num_nodes());                                     // :882 one Visited per batch
    for (int64_t j = 0; j < nq; j++) {
      Counters c;
      one_query(h, visited, q + j * h.space.dim, k, ef, ids + j * k, dists + j * k, c);
      if (counters) { counters[3 * j] = c.n_dist; counters[3 * j + 1] = c.n_exp0; counters[3 * j + 2] = c.n_expU; }
    }
  });
}

// Query-parallel courtesy variant for the CPU baseline: each thread owns a Visited.  Results are
// identical to orc_vec_search (queries are independent).  Returns seconds spent in *seconds.
int orc_vec_search_mt(void* p, const float* q, int64_t nq, int k, int ef, int32_t* ids, float* dists,
                      int nthreads, double* seconds, uint64_t* total_counters /* [3] or null */) {
  VecHandle* v = (VecHandle*)p;
  return guard([&] {
    const VecHnsw& h = v->h;
    if (ef < k) throw std::invalid_argument("ef must be >= k");
    if (!h.has_entry) throw std::invalid_argument("knn: empty hgraph");
    uint64_t nd = 0, n0 = 0, nu = 0;
    auto t0 = std::chrono::steady_clock::now();
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads) reduction(+ : nd, n0, nu)
#endif
    {
      Visited visited(h.num_nodes());
      Counters c;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
      for (int64_t j = 0; j < nq; j++)
        one_query(h, visited, q + j * h.space.dim, k, ef, ids + j * k, dists + j * k, c);
      nd += c.n_dist; n0 += c.n_exp0; nu += c.n_expU;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    if (total_counters) { total_counters[0] = nd; total_counters[1] = n0; total_counters[2] = nu; }
  });
}

int orc_vec_info(void* p, int64_t* n, int* max_layer, int64_t* entry, int* dim) {
  VecHandle* v = (VecHandle*)p;
  if (n) *n = (int64_t)v->h.num_nodes();
  if (max_layer) *max_layer = v->h.max_layer();
  if (entry) *entry = v->h.has_entry ? v->h.entry_point : -1;
  if (dim) *dim = v->h.space.dim;
  return 0;
}
void orc_vec_counters(void* p, uint64_t* out3, int reset) {
  VecHandle* v = (VecHandle*)p;
  out3[0] = v->h.counters.n_dist; out3[1] = v->h.counters.n_exp0; out3[2] = v->h.counters.n_expU;
  if (reset) v->h.counters = Counters();
}
int orc_vec_levels(void* p, int32_t* out) {
  VecHandle* v = (VecHandle*)p;
  std::copy(v->h.levels.begin(), v->h.levels.end(), out);
  return 0;
}
int64_t orc_vec_layer_nnz(void* p, int layer) {
  VecHandle* v = (VecHandle*)p;
  if (layer < 0 || layer > v->h.max_layer()) return -1;
  int64_t nnz = 0;
  for (const Neighbours& nb : v->h.layers[layer].v) nnz += (int64_t)nb.length();
  return nnz;
}
// CSR of one layer, list order preserved (head first).  offsets has n+1 entries.
int orc_vec_export_layer(void* p, int layer, int64_t* offsets, int32_t* nbrs) {
  VecHandle* v = (VecHandle*)p;
  return guard([&] {
    const Graph& g = v->h.layers.at(layer);
    int64_t o = 0;
    for (size_t i = 0; i < g.num_nodes(); i++) {
      offsets[i] = o;
      for (int32_t x : g.v[i].list) nbrs[o++] = x;
    }
    offsets[g.num_nodes()] = o;
  });
}
// Replace the whole index by an imported graph (same exchange layout as the export).
int orc_vec_import(void* p, const float* data, int64_t n, int max_layer, int64_t entry,
                   const int64_t* const* layer_offsets, const int32_t* const* layer_nbrs) {
  VecHandle* v = (VecHandle*)p;
  return guard([&] {
    VecHnsw& h = v->h;
    int dim = h.space.dim;
    h.space.data.assign(data, data + (size_t)n * dim);
    h.layers.clear();
    for (int l = 0; l <= max_layer; l++) {
      h.layers.emplace_back((size_t)n);
      for (int64_t i = 0; i < n; i++) {
        Neighbours& nb = h.layers[l].v[i];
        nb.list.assign(layer_nbrs[l] + layer_offsets[l][i], layer_nbrs[l] + layer_offsets[l][i + 1]);
        for (int32_t x : nb.list) if (x < 0 || x >= n) throw std::invalid_argument("import: neighbour id out of range");
      }
    }
    h.levels.assign(n, 0);
    h.has_entry = false; h.entry_point = -1;
    if (n > 0) h.set_entry_point(entry);
    v->visited = Visited((size_t)n);
  });
}
int orc_vec_invariant(void* p) { return ((VecHandle*)p)->h.invariant() ? 1 : 0; }
double orc_vec_distance(void* p, const float* a, const float* b) { return ((VecHandle*)p)->h.space.distance(a, b); }
// The fp32 value comparisons are made on (sum of squares for L2, 1-dot / -dot otherwise).
float orc_work_distance(const float* a, const float* b, int dim, int metric, int order) {
  if (metric == METRIC_L2) return order == SUM_TEAM8 ? ssqr_diff_team8(a, b, dim) : ssqr_diff_sequential(a, b, dim);
  float dt = order == SUM_TEAM8 ? dot_team8(a, b, dim) : dot_sequential(a, b, dim);
  return metric == METRIC_ANGULAR ? 1.0f - dt : -dt;
}

// brute_force_knn_l2 (benchmark/dataset.ml:15-30): all n distances per query, full sort,
// first k; distances only in the reference — ids are returned as well here, ordered by
// (distance, id).  dists are fp32 stores of the double distance.
int orc_bruteforce(const float* data, int64_t n, const float* q, int64_t nq, int dim, int k, int metric,
                   int order, int32_t* ids, float* dists, int nthreads) {
  return guard([&] {
    VecSpace sp; sp.dim = dim; sp.metric = metric; sp.order = order;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 4)
#endif
    for (int64_t j = 0; j < nq; j++) {
      std::vector<std::pair<float, int32_t>> d((size_t)n);
      for (int64_t i = 0; i < n; i++)
        d[i] = {(float)sp.distance(data + i * dim, q + j * dim), (int32_t)i};   // dists.{j} <- ... (fp32 vec)
      size_t kk = std::min<size_t>(k, n);
      std::partial_sort(d.begin(), d.begin() + kk, d.end());
      for (int i = 0; i < k; i++) {
        if ((size_t)i < kk) { ids[j * k + i] = d[i].second; dists[j * k + i] = d[i].first; }
        else { ids[j * k + i] = -1; dists[j * k + i] = std::numeric_limits<float>::quiet_NaN(); }
      }
    }
  });
}

// Recall.compute (benchmark/dataset.ml:105-127).  expected/got are [nq][k] (= Lacaml k x nq).
int orc_recall(const float* expected, const float* got, int64_t nq, int k, double epsilon, double* out) {
  double ret = 0.;
  for (int64_t q = 0; q < nq; q++) {
    int num_ok = 0;
    for (int i = 0; i < k; i++)
      if ((double)got[q * k + i] <= (double)expected[q * k + (k - 1)] + epsilon) num_ok++;   // NaN never counts
    ret += (double)num_ok / (double)k;
  }
  *out = ret / (double)nq;
  return 0;
}

// ---- Abs1D space: drives the reference's inline golden tests ------------------------------------
void* orc_abs_create(const double* values, int64_t n) {
  AbsHandle* a = new AbsHandle();
  a->h.space.data.assign(values, values + n);
  return a;
}
void orc_abs_destroy(void* p) { delete (AbsHandle*)p; }
// Graph.create n on a layer (ohnsw.ml:166-167)
int orc_abs_layer_create(void* p, int layer, int64_t n) {
  AbsHandle* a = (AbsHandle*)p;
  return guard([&] { a->h.layers.at(layer) = Graph((size_t)n); a->visited = Visited((size_t)n); });
}
// Graph.Test.create_loop (ohnsw.ml:205-212) over all values
int orc_abs_layer_create_loop(void* p, int layer) {
  AbsHandle* a = (AbsHandle*)p;
  return guard([&] {
    int n = (int)a->h.space.data.size();
    Graph g((size_t)n);
    auto wrap = [n](int i) { return i < 0 ? i + n : (i >= n ? i - n : i); };
    for (int i = 0; i < n; i++) {
      Neighbours nb; nb.list = {wrap(i - 1), wrap(i + 1)};
      g.set_connections(i, nb);
    }
    a->h.layers.at(layer) = g;
    a->visited = Visited((size_t)n);
  });
}
int orc_abs_set_connections(void* p, int layer, int node, const int32_t* ids, int n) {
  AbsHandle* a = (AbsHandle*)p;
  return guard([&] { Neighbours nb; nb.list.assign(ids, ids + n); a->h.layers.at(layer).set_connections(node, nb); });
}
int orc_abs_adjacent(void* p, int layer, int node, int32_t* out, int cap) {
  AbsHandle* a = (AbsHandle*)p;
  const Neighbours& nb = a->h.layers.at(layer).adjacent(node);
  int n = (int)nb.length();
  for (int i = 0; i < n && i < cap; i++) out[i] = nb.list[i];
  return n;
}
int orc_abs_add_node(void* p) { return ((AbsHandle*)p)->h.add_node(); }
int orc_abs_graph_add_node(void* p, int layer) { ((AbsHandle*)p)->h.layers.at(layer).add_node(); return 0; }
int orc_abs_set_entry_point(void* p, int64_t n) { AbsHandle* a = (AbsHandle*)p; return guard([&] { a->h.set_entry_point(n); }); }
int64_t orc_abs_entry_point(void* p) { AbsHandle* a = (AbsHandle*)p; return a->h.has_entry ? a->h.entry_point : -1; }
int orc_abs_set_max_layer(void* p, int n) { ((AbsHandle*)p)->h.set_max_layer(n); return 0; }
int orc_abs_max_layer(void* p) { return ((AbsHandle*)p)->h.max_layer(); }
int64_t orc_abs_num_nodes(void* p) { return (int64_t)((AbsHandle*)p)->h.num_nodes(); }
int64_t orc_abs_layer_num_nodes(void* p, int layer) {
  AbsHandle* a = (AbsHandle*)p;
  if (layer < 0 || layer > a->h.max_layer()) return -1;
  return (int64_t)a->h.layers[layer].num_nodes();
}
int orc_abs_invariant(void* p) { return ((AbsHandle*)p)->h.invariant() ? 1 : 0; }
int orc_abs_graph_invariant(void* p, int layer) { return ((AbsHandle*)p)->h.layers.at(layer).invariant() ? 1 : 0; }
int orc_abs_search_one(void* p, int layer, int start, double target) {
  AbsHandle* a = (AbsHandle*)p; Counters c;
  return a->h.search_one(layer, start, target, c);
}
// start nodes given by id (MinQueue.add_node); returns count, ascending (distance, node).
int orc_abs_search_k(void* p, int layer, const int32_t* start, int n_start, int k, double target,
                     int32_t* out_nodes, double* out_dists, int cap) {
  AbsHandle* a = (AbsHandle*)p; Counters c;
  MinHeap s;
  for (int i = 0; i < n_start; i++) s.push(a->h.element(target, start[i], c));
  MinHeap r = a->h.search_k(layer, a->visited, s, k, target, c);
  int n = 0;
  while (!r.empty()) { if (n < cap) { out_nodes[n] = r.top().node; out_dists[n] = r.top().distance; } n++; r.pop(); }
  return n;
}
// returns count; out in list order (head first)
int orc_abs_select(void* p, double target, const int32_t* cands, int n_cands, int num, int32_t* out, int cap) {
  AbsHandle* a = (AbsHandle*)p; Counters c;
  MinHeap s;
  for (int i = 0; i < n_cands; i++) s.push(a->h.element(target, cands[i], c));
  Neighbours sel = a->h.select_neighbours(s, num, c);
  int n = (int)sel.length();
  for (int i = 0; i < n && i < cap; i++) out[i] = sel.list[i];
  return n;
}
// full insert / knn on the 1-D space (small end-to-end cases)
int orc_abs_insert_all(void* p, int M, int efC, const int32_t* levels) {
  AbsHandle* a = (AbsHandle*)p;
  return guard([&] {
    double level_mult = 1.0 / std::log((double)M);
    size_t n = a->h.space.data.size();
    a->visited.grow(n);
    for (size_t i = a->h.num_nodes(); i < n; i++)
      a->h.insert(a->h.space.data[i], M, efC, level_mult, a->visited, levels ? levels[i] : -1);
  });
}
int orc_abs_knn(void* p, double target, int k, int32_t* out_nodes, double* out_dists) {
  AbsHandle* a = (AbsHandle*)p;
  int n = 0;
  int rc = guard([&] {
    Counters c;
    a->visited.grow(a->h.num_nodes());
    MinHeap r = a->h.knn(a->visited, k, target, c);
    while (!r.empty()) { out_nodes[n] = r.top().node; out_dists[n] = r.top().distance; n++; r.pop(); }
  });
  return rc ? -rc : n;
}

// ---- containers ---------------------------------------------------------------------------------
void* orc_nb_create() { return new Neighbours(); }
void orc_nb_destroy(void* p) { delete (Neighbours*)p; }
void orc_nb_add(void* p, int node) { ((Neighbours*)p)->add(node); }
void orc_nb_remove(void* p, int node) { ((Neighbours*)p)->remove(node); }
int orc_nb_length(void* p) { return (int)((Neighbours*)p)->length(); }
int orc_nb_get(void* p, int32_t* out, int cap) {
  Neighbours* nb = (Neighbours*)p;
  for (int i = 0; i < (int)nb->length() && i < cap; i++) out[i] = nb->list[i];
  return (int)nb->length();
}
void* orc_visited_create(int64_t n) { return new Visited((size_t)n); }
void orc_visited_destroy(void* p) { delete (Visited*)p; }
int orc_visited_mem(void* p, int64_t node) {
  Visited* v = (Visited*)p;
  if (node < 0 || (size_t)node >= v->visited.size()) return -1;   // OCaml raises Invalid_argument "index out of bounds"
  return v->mem((int32_t)node) ? 1 : 0;
}
void orc_visited_add(void* p, int64_t node) { ((Visited*)p)->add((int32_t)node); }
void orc_visited_clear(void* p) { ((Visited*)p)->clear(); }
int64_t orc_visited_card(void* p) { return (int64_t)((Visited*)p)->card(); }
void orc_visited_set_epoch_near_max(void* p, int64_t below) { ((Visited*)p)->epoch = Visited::kIntMax - below; }
int64_t orc_visited_epoch(void* p) { return ((Visited*)p)->epoch; }

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}  // extern "C"

// The scalar definitions of the TEAM8 order, exported so a test can pin the AVX2 path to them.
extern "C" float orc_work_distance_scalar(const float* a, const float* b, int dim, int metric) {
  if (metric == oracle::METRIC_L2) return oracle::ssqr_diff_team8_scalar(a, b, dim);
  float dt = oracle::dot_team8_scalar(a, b, dim);
  return metric == oracle::METRIC_ANGULAR ? 1.0f - dt : -dt;
}
