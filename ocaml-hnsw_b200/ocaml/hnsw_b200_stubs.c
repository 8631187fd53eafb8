/* OCaml <-> libhnsw_b200.so stubs.  Bigarray payloads (Lacaml.S.mat = float32, Fortran layout,
 * dim x n == C float[n][dim]) are passed without a copy; the runtime lock is released around GPU
 * work; status codes map to the exceptions the reference raises (include/hnsw_b200.h).
 * NOT compiled in the build container (no caml headers there). */
#include <string.h>
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/bigarray.h>
#include <caml/threads.h>
#include "hnsw_b200.h"

#define Index_val(v) (*((hnswb200_index**)Data_custom_val(v)))

static void index_finalize(value v) {
  hnswb200_index* x = Index_val(v);
  if (x) { hnswb200_destroy(x); Index_val(v) = NULL; }
}
static struct custom_operations index_ops = {
  "hnsw_b200.index", index_finalize, custom_compare_default, custom_hash_default,
  custom_serialize_default, custom_deserialize_default, custom_compare_ext_default, custom_fixed_length_default };

static void check(int rc) {
  if (rc == HNSWB200_OK) return;
  const char* msg = hnswb200_last_error();
  if (rc == HNSWB200_EINVAL) caml_invalid_argument(msg);     /* e.g. "knn: empty hgraph", lib/ohnsw.ml:862 */
  if (rc == HNSWB200_ENOMEM) caml_raise_out_of_memory();
  caml_failwith(msg);
}

CAMLprim value hb_create(value dim, value metric, value m, value efc, value seed, value device) {
  CAMLparam5(dim, metric, m, efc, seed);
  CAMLxparam1(device);
  CAMLlocal1(v);
  hnswb200_index* x = NULL;
  check(hnswb200_create(&x, Int_val(dim), Int_val(metric), Int_val(m), Int_val(efc), (uint64_t)Long_val(seed), Int_val(device)));
  v = caml_alloc_custom(&index_ops, sizeof(hnswb200_index*), 0, 1);
  Index_val(v) = x;
  CAMLreturn(v);
}
CAMLprim value hb_create_byte(value* a, int n) { (void)n; return hb_create(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value hb_close(value v) { CAMLparam1(v); index_finalize(v); CAMLreturn(Val_unit); }
CAMLprim value hb_set_flavour(value v, value f) { CAMLparam2(v, f); check(hnswb200_set_flavour(Index_val(v), Int_val(f))); CAMLreturn(Val_unit); }

/* batch : Lacaml.S.mat (dim x n).  levels : (int32, c_layout) Array1 or empty for "draw". */
static value build_or_insert(value v, value batch, value levels, int insert) {
  CAMLparam3(v, batch, levels);
  hnswb200_index* x = Index_val(v);
  const float* data = (const float*)Caml_ba_data_val(batch);
  int64_t n = Caml_ba_array_val(batch)->dim[1];
  const int32_t* lv = Caml_ba_array_val(levels)->dim[0] > 0 ? (const int32_t*)Caml_ba_data_val(levels) : NULL;
  int rc;
  caml_release_runtime_system();            /* Bigarray payloads are off-heap: safe while released */
  rc = insert ? hnswb200_insert(x, data, n, lv) : hnswb200_build(x, data, n, lv);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value hb_build(value v, value batch, value levels) { return build_or_insert(v, batch, levels, 0); }
CAMLprim value hb_insert(value v, value batch, value levels) { return build_or_insert(v, batch, levels, 1); }

/* queries : Lacaml.S.mat (dim x nq); ids : (int32, c_layout) Array2 nq x k; dists : Lacaml.S.mat k x nq */
CAMLprim value hb_search(value v, value queries, value k, value ef, value ids, value dists) {
  CAMLparam5(v, queries, k, ef, ids);
  CAMLxparam1(dists);
  hnswb200_index* x = Index_val(v);
  const float* q = (const float*)Caml_ba_data_val(queries);
  int64_t nq = Caml_ba_array_val(queries)->dim[1];
  int32_t* pi = (int32_t*)Caml_ba_data_val(ids);
  float* pd = (float*)Caml_ba_data_val(dists);
  int kk = Int_val(k), e = Int_val(ef), rc;
  caml_release_runtime_system();
  rc = hnswb200_search(x, q, nq, kk, e, HNSWB200_MODE_PARITY, pi, pd);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value hb_search_byte(value* a, int n) { (void)n; return hb_search(a[0], a[1], a[2], a[3], a[4], a[5]); }

CAMLprim value hb_info(value v) {
  CAMLparam1(v);
  CAMLlocal1(r);
  hnswb200_info inf;
  check(hnswb200_get_info(Index_val(v), &inf));
  r = caml_alloc_tuple(3);
  Store_field(r, 0, Val_long(inf.n));
  Store_field(r, 1, Val_int(inf.max_layer));
  Store_field(r, 2, Val_long(inf.entry_point));
  CAMLreturn(r);
}

CAMLprim value hb_params(value v) {
  CAMLparam1(v);
  CAMLlocal1(r);
  hnswb200_info inf;
  check(hnswb200_get_info(Index_val(v), &inf));
  r = caml_alloc_tuple(2);
  Store_field(r, 0, Val_int(inf.M));
  Store_field(r, 1, Val_int(inf.ef_construction));
  CAMLreturn(r);
}

CAMLprim value hb_bruteforce(value train, value test, value k, value dists) {
  CAMLparam4(train, test, k, dists);
  const float* x = (const float*)Caml_ba_data_val(train);
  const float* q = (const float*)Caml_ba_data_val(test);
  int64_t n = Caml_ba_array_val(train)->dim[1], nq = Caml_ba_array_val(test)->dim[1];
  int dim = (int)Caml_ba_array_val(train)->dim[0], kk = Int_val(k), rc;
  float* pd = (float*)Caml_ba_data_val(dists);
  caml_release_runtime_system();
  rc = hnswb200_bruteforce_knn(x, n, q, nq, dim, kk, HNSWB200_L2, 0, NULL, pd);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}

CAMLprim value hb_pin(value ba) {
  CAMLparam1(ba);
  check(hnswb200_host_register(Caml_ba_data_val(ba), (int64_t)caml_ba_byte_size(Caml_ba_array_val(ba))));
  CAMLreturn(Val_unit);
}
CAMLprim value hb_unpin(value ba) { CAMLparam1(ba); check(hnswb200_host_unregister(Caml_ba_data_val(ba))); CAMLreturn(Val_unit); }

/* ---- graph exchange, statistics, multi-GPU (hnsw_b200_graph.ml) ------------------------------------ */

/* offsets : (int64, c_layout) Array1.t array (one per layer, n+1 entries each); nbrs : (int32, c_layout) Array1.t array */
CAMLprim value hb_import_graph(value v, value data, value id_base, value entry, value offsets, value nbrs) {
  CAMLparam5(v, data, id_base, entry, offsets);
  CAMLxparam1(nbrs);
  hnswb200_index* x = Index_val(v);
  int layers = (int)Wosize_val(offsets), rc, l;
  const int64_t* po[16];
  const int32_t* pn[16];
  if (layers < 1 || layers > 16 || (int)Wosize_val(nbrs) != layers) caml_invalid_argument("import_graph: 1..16 layers, one offsets and one nbrs array each");
  for (l = 0; l < layers; l++) {
    po[l] = (const int64_t*)Caml_ba_data_val(Field(offsets, l));
    pn[l] = (const int32_t*)Caml_ba_data_val(Field(nbrs, l));
  }
  const float* d = (const float*)Caml_ba_data_val(data);
  int64_t n = Caml_ba_array_val(data)->dim[1];
  int base = Int_val(id_base);
  int64_t e = Long_val(entry);
  caml_release_runtime_system();
  rc = hnswb200_import_graph(x, d, n, base, layers - 1, e, po, pn);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value hb_import_graph_byte(value* a, int n) { (void)n; return hb_import_graph(a[0], a[1], a[2], a[3], a[4], a[5]); }

/* number of links on a layer (first call of the two-call pattern) */
CAMLprim value hb_export_layer_nnz(value v, value layer) {
  CAMLparam2(v, layer);
  int64_t nnz = 0;
  check(hnswb200_export_layer(Index_val(v), Int_val(layer), 0, NULL, NULL, &nnz));
  CAMLreturn(Val_long(nnz));
}
/* offsets : (int64, c_layout) Array1.t of n+1; nbrs : (int32, c_layout) Array1.t of nnz */
CAMLprim value hb_export_layer(value v, value layer, value id_base, value offsets, value nbrs) {
  CAMLparam5(v, layer, id_base, offsets, nbrs);
  int64_t nnz = 0;
  check(hnswb200_export_layer(Index_val(v), Int_val(layer), Int_val(id_base), (int64_t*)Caml_ba_data_val(offsets),
                              (int32_t*)Caml_ba_data_val(nbrs), &nnz));
  CAMLreturn(Val_unit);
}
CAMLprim value hb_export_levels(value v, value levels) {
  CAMLparam2(v, levels);
  check(hnswb200_export_levels(Index_val(v), (int32_t*)Caml_ba_data_val(levels)));
  CAMLreturn(Val_unit);
}

/* Hgraph.Stats (lib/hnsw.ml:353-375) per layer: (nodes, min, max, mean, isolated) array */
static value layer_stats(const hnswb200_stats* st) {
  CAMLlocal2(arr, row);
  int l;
  arr = caml_alloc_tuple((uintnat)(st->num_layers > 0 ? st->num_layers : 0));
  for (l = 0; l < st->num_layers && l < 16; l++) {
    row = caml_alloc_tuple(5);
    Store_field(row, 0, Val_long(st->layer_nodes[l]));
    Store_field(row, 1, Val_int(st->layer_min_degree[l]));
    Store_field(row, 2, Val_int(st->layer_max_degree[l]));
    Store_field(row, 3, caml_copy_double(st->layer_mean_degree[l]));
    Store_field(row, 4, Val_long(st->layer_isolated[l]));
    Store_field(arr, l, row);
  }
  return arr;
}
/* (layers, distance evaluations of the last search, of the last build, build seconds, search kernel ms) */
static value stats_tuple(const hnswb200_stats* st) {
  CAMLlocal2(r, layers);
  layers = layer_stats(st);
  r = caml_alloc_tuple(5);
  Store_field(r, 0, layers);
  Store_field(r, 1, Val_long((intnat)st->search_n_dist));
  Store_field(r, 2, Val_long((intnat)st->build_n_dist));
  Store_field(r, 3, caml_copy_double(st->build_seconds));
  Store_field(r, 4, caml_copy_double(st->search_kernel_ms));
  return r;
}
CAMLprim value hb_stats(value v) {
  CAMLparam1(v);
  hnswb200_stats st;
  check(hnswb200_get_stats(Index_val(v), &st));
  CAMLreturn(stats_tuple(&st));
}

/* -- one process, every GPU (hnswb200_sharded_*) */
#define Sharded_val(v) (*((hnswb200_sharded**)Data_custom_val(v)))
static void sharded_finalize(value v) {
  hnswb200_sharded* s = Sharded_val(v);
  if (s) { hnswb200_sharded_destroy(s); Sharded_val(v) = NULL; }
}
static struct custom_operations sharded_ops = {
  "hnsw_b200.sharded", sharded_finalize, custom_compare_default, custom_hash_default,
  custom_serialize_default, custom_deserialize_default, custom_compare_ext_default, custom_fixed_length_default };

/* devices : int array */
CAMLprim value hb_sharded_create(value dim, value metric, value m, value efc, value seed, value devices) {
  CAMLparam5(dim, metric, m, efc, seed);
  CAMLxparam1(devices);
  CAMLlocal1(v);
  int dev[32], n = (int)Wosize_val(devices), i;
  hnswb200_sharded* s = NULL;
  if (n < 1 || n > 32) caml_invalid_argument("sharded_create: 1..32 devices");
  for (i = 0; i < n; i++) dev[i] = Int_val(Field(devices, i));
  check(hnswb200_sharded_create(&s, Int_val(dim), Int_val(metric), Int_val(m), Int_val(efc), (uint64_t)Long_val(seed), n, dev));
  v = caml_alloc_custom(&sharded_ops, sizeof(hnswb200_sharded*), 0, 1);
  Sharded_val(v) = s;
  CAMLreturn(v);
}
CAMLprim value hb_sharded_create_byte(value* a, int n) { (void)n; return hb_sharded_create(a[0], a[1], a[2], a[3], a[4], a[5]); }
CAMLprim value hb_sharded_close(value v) { CAMLparam1(v); sharded_finalize(v); CAMLreturn(Val_unit); }
CAMLprim value hb_sharded_set_flavour(value v, value f) { CAMLparam2(v, f); check(hnswb200_sharded_set_flavour(Sharded_val(v), Int_val(f))); CAMLreturn(Val_unit); }

CAMLprim value hb_sharded_build(value v, value batch, value levels) {
  CAMLparam3(v, batch, levels);
  hnswb200_sharded* s = Sharded_val(v);
  const float* data = (const float*)Caml_ba_data_val(batch);
  int64_t n = Caml_ba_array_val(batch)->dim[1];
  const int32_t* lv = Caml_ba_array_val(levels)->dim[0] > 0 ? (const int32_t*)Caml_ba_data_val(levels) : NULL;
  int rc;
  caml_release_runtime_system();
  rc = hnswb200_sharded_build(s, data, n, lv);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value hb_sharded_search(value v, value queries, value k, value ef, value ids, value dists) {
  CAMLparam5(v, queries, k, ef, ids);
  CAMLxparam1(dists);
  hnswb200_sharded* s = Sharded_val(v);
  const float* q = (const float*)Caml_ba_data_val(queries);
  int64_t nq = Caml_ba_array_val(queries)->dim[1];
  int32_t* pi = (int32_t*)Caml_ba_data_val(ids);
  float* pd = (float*)Caml_ba_data_val(dists);
  int kk = Int_val(k), e = Int_val(ef), rc;
  caml_release_runtime_system();
  rc = hnswb200_sharded_search(s, q, nq, kk, e, HNSWB200_MODE_PARITY, pi, pd);
  caml_acquire_runtime_system();
  check(rc);
  CAMLreturn(Val_unit);
}
CAMLprim value hb_sharded_search_byte(value* a, int n) { (void)n; return hb_sharded_search(a[0], a[1], a[2], a[3], a[4], a[5]); }
CAMLprim value hb_sharded_stats(value v) {
  CAMLparam1(v);
  hnswb200_stats st;
  check(hnswb200_sharded_get_stats(Sharded_val(v), &st));
  CAMLreturn(stats_tuple(&st));
}
CAMLprim value hb_sharded_num_nodes(value v) {
  CAMLparam1(v);
  hnswb200_info inf;
  int ns = 0;
  check(hnswb200_sharded_get_info(Sharded_val(v), &inf, &ns));
  CAMLreturn(Val_long(inf.n));
}
