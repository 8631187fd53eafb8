/* hnsw_b200.h — C ABI of libhnsw_b200.so, the B200 (sm_100a) HNSW build + k-NN search engine
 * that sits behind the OCaml API of lehy/ocaml-hnsw.
 *
 * The reference has no FFI seam of its own on this path (its only native boundary is the
 * per-distance Lacaml stub, lib/ohnsw.ml:899).  The seam is therefore the set of OCaml values
 * the benchmark and tests call; each entry point below names the reference interface it
 * replaces.  The OCaml `external` declarations and C stubs a maintainer adds are in
 * ocaml-hnsw_b200/ocaml/ and shown in INTEGRATION.md.
 *
 * Conventions
 *   - Vectors are fp32, row-major `float[n][dim]`, dense.  This is byte-for-byte the payload
 *     of a `Lacaml.S.mat` of dim1 = dim, dim2 = n (Fortran layout, one vector per column;
 *     lib/ohnsw.ml:840-846), so `Caml_ba_data_val` can be passed without a copy.
 *   - Node ids are 0-based int32 (path B, lib/ohnsw.ml:842).  `id_base` on import/export lets
 *     path A's 1-based graphs (lib/hnsw.ml:313-325) cross the boundary.
 *   - Result buffers are caller-allocated `[nq][k]` (= a Lacaml k x nq mat).  Missing results
 *     are padded with id -1 and distance NaN (lib/ohnsw.ml:879-881), or +inf when
 *     HNSWB200_FLAVOUR_HNSW_BA is selected (lib/hnsw.ml:770-771).
 *   - Every function returns a status; 0 is success.  The message for the last failure on the
 *     calling thread is hnswb200_last_error().  Status -> OCaml exception:
 *       1 -> Invalid_argument msg   (lib/ohnsw.ml:862 "knn: empty hgraph", :343, dataset.ml:112)
 *       2 -> Failure msg            (CUDA / NCCL error)
 *       3 -> Out_of_memory
 *   - Host pointers unless the name ends in `_device`.  Input buffers are borrowed for the
 *     duration of the call only.  Calls on one handle are serialised internally; the OCaml
 *     stubs release the runtime lock around them.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails with 2.
 */
#ifndef HNSW_B200_H
#define HNSW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HNSWB200_OK 0
#define HNSWB200_EINVAL 1
#define HNSWB200_ECUDA 2
#define HNSWB200_ENOMEM 3

/* distance: the reference's `'a distance` closure (lib/ohnsw.ml:3) narrowed to a tag — an
 * arbitrary OCaml closure cannot run on the GPU.  L2 is Ohnsw.distance_l2 (lib/ohnsw.ml:899,
 * true Euclidean: sqrt of the fp32 sum of squares).  ANGULAR is 1 - a.b, IP is -a.b. */
#define HNSWB200_L2 0
#define HNSWB200_ANGULAR 1
#define HNSWB200_IP 2

/* search mode: the one traversal this library has.  PARITY reproduces the reference's sequential
 * best-first search id for id (one expansion per iteration, exact visited set, (distance, id) tie order).
 * A relaxed mode (two expansions per iteration, no id parity) was built and measured in round 2: with
 * every SM full of queries the kernel is bound by instructions and bytes per expansion, not by the
 * length of a query's dependent chain, so it was no faster (1.21 ms vs 1.21 ms per 10k queries in the
 * same build) and its extra code cost the PARITY path 14 %; it was removed rather than shipped as a
 * second, slower name for the same thing.  Any other mode value is rejected. */
#define HNSWB200_MODE_PARITY 0

/* which reference code path's quirks are mirrored:
 * OHNSW   = lib/ohnsw.ml   (strict accept `d < top`, 2M links for a new node on layer 0,
 *                           -1/NaN padding)
 * HNSW_BA = lib/hnsw.ml + lib/hnsw_algo.ml `Hnsw.Ba` (ties accepted `d <= top`,
 *                           M links for a new node on every layer, small candidate sets kept whole,
 *                           +inf padding).
 * What HNSW_BA is NOT: an edge-for-edge restatement of path A's build.  It is path B's insert (pruning keyed
 * on distances to the node being pruned, the paper's rule) run with path A's parameters.  Path A prunes a
 * neighbour with distances to the INSERTED point (lib/hnsw_algo.ml:633-635,678-689), force-keeps candidates of
 * degree <= 1 (`do_not_isolate`, :591-592), descends upper layers with a heap (:393-437) and orders its lists
 * by pairing-heap / Base.Map fold order; none of that is pinned by a reference test and none is reproduced.
 * Why the prune is not: the overflowing list contains the inserted point itself at distance 0, so it is kept
 * first, and every other member e is then tested with `distance e point > e.distance_to_target` (:578-583) — the
 * same two vectors on both sides, never true — so a pruned node keeps ONLY the inserted point plus its
 * degree-<=1 members: each overflow wipes a node's list (the reason `do_not_isolate` was added, :584-587, and,
 * one presumes, why benchmark.ml moved to path B).  A drop-in that reproduced it would reproduce the recall loss.
 * Search on an imported path-A graph (id_base = 1) follows path A's acceptance rule exactly. */
#define HNSWB200_FLAVOUR_OHNSW 0
#define HNSWB200_FLAVOUR_HNSW_BA 1

typedef struct hnswb200_index hnswb200_index;

typedef struct hnswb200_info {
  int64_t n;            /* number of nodes                      Hgraph.num_nodes  ohnsw.ml:335 */
  int32_t dim;
  int32_t metric;
  int32_t M;            /* num_connections                      ohnsw.ml:767 */
  int32_t ef_construction; /* num_nodes_search_construction     ohnsw.ml:768 */
  int32_t max_layer;    /* Hgraph.max_layer                     ohnsw.ml:346 */
  int64_t entry_point;  /* -1 when empty                        ohnsw.ml:340 */
  int32_t slots0;       /* adjacency row width on layer 0 (2M)  ohnsw.ml:818 */
  int32_t slots_upper;  /* row width on layers >= 1 (M) */
  int32_t flavour;
  int32_t device;
} hnswb200_info;

/* Work counters of the last search call and the last build call, plus graph shape.
 * n_dist counts what the reference's distance-call counter counts (lib/hnsw.ml:732-751). */
typedef struct hnswb200_stats {
  uint64_t search_queries;
  uint64_t search_n_dist;     /* distance evaluations */
  uint64_t search_n_exp0;     /* adjacency rows read on layer 0 */
  uint64_t search_n_expU;     /* adjacency rows read on layers >= 1 */
  uint64_t search_visited_overflows; /* queries whose visited set left shared memory */
  double   search_algorithmic_bytes; /* n_dist*4*dim + n_exp0*4*slots0 + n_expU*4*slots_upper + nq*(4*dim + 8*k) */
  double   search_kernel_ms;  /* device time of the search kernel (CUDA events) */
  uint64_t build_inserts;
  uint64_t build_n_dist;
  uint64_t build_n_exp;
  double   build_algorithmic_bytes;
  double   build_seconds;     /* wall time of the last build/insert call, H2D included */
  uint64_t gpu_launches;      /* kernels launched by this handle since creation */
  /* Hgraph.Stats (lib/hnsw.ml:353-375): per layer size / min / max / mean degree / isolated */
  int32_t  num_layers;
  int64_t  layer_nodes[16];
  int32_t  layer_min_degree[16];
  int32_t  layer_max_degree[16];
  double   layer_mean_degree[16];
  int64_t  layer_isolated[16];
  uint64_t build_visited_overflows;  /* inserts whose visited set left shared memory */
  uint64_t search_tie_overflows;     /* queries whose list of evicted candidates tied at the beam's top distance outgrew
                                        shared memory (32) AND its global region (32k): the surplus was not revisited, the
                                        only way a PARITY search can differ from the reference; hnswb200_search fails on it */
  uint64_t search_tie_spills;        /* queries whose tie list continued in a global region (exact, just slower) */
  uint64_t build_dropped_incoming;   /* links dropped because one row received more than 96 new nodes in ONE build batch
                                        (the 96 smallest ids are kept and re-selected); 0 on every shape measured */
  uint64_t search_zero_copy;         /* last hnswb200_search: bit 0 = the queries were read from the caller's pinned buffer by
                                        the kernel itself, bit 1 = the result rows were stored straight into the caller's */
} hnswb200_stats;

/* ---- lifetime ------------------------------------------------------------------------------ */

/* Replaces Ohnsw.Hgraph.create (lib/ohnsw.ml:316-324) / Hnsw.Ba's Hgraph.create
 * (lib/hnsw.ml:388-394).  `seed` drives the level draw (lib/ohnsw.ml:781) when levels are not
 * supplied.  `device` is the CUDA ordinal this index lives on (one process per GPU). */
int hnswb200_create(hnswb200_index** out, int dim, int metric, int M, int ef_construction,
                    uint64_t seed, int device);
int hnswb200_set_flavour(hnswb200_index* idx, int flavour);
/* Tunables (0 = automatic): "hash_slots" (visited hash slots per query), "visited_mode" (1 = shared
 * memory hash, 2 = global bitset), "warps_per_cta", "max_warps_per_sm", "build_batch" (max inserts
 * per GPU batch, default 16384; 1 = sequential inserts: the reference's own order, edge for edge),
 * "build_ratio" (a batch is at most 1 / build_ratio of the graph the call will end with, default 64),
 * "build_ratio_early" (... and at most 1 / this of the graph so far, default 4: early rows are re-selected
 * many times as the graph grows, so coarse early batches leave no trace in the finished index),
 * "build_mates" (default 1: members of a batch that selected the same neighbour are proposed to each other,
 * which restores the links the sequential loop would have made between them; 0 = batch members never link),
 * "build_qreg" (phase 1 of the build: 0 = automatic, 1 = new node's vector in registers, 2 = in shared memory
 * only; same graph either way), "host_chunks" (2..8: hnswb200_search streams batches of
 * >= 4096 queries to the GPU in this many pieces behind ONE already running search kernel whose
 * warps wait for the piece that holds their query; default: copy first, then search),
 * "host_zero_copy" (default 1: hnswb200_search reads PINNED query buffers and writes PINNED result buffers from
 * inside the kernel instead of copying them — see hnswb200_host_register; 0 = always copy),
 * "stage_rows" (rows of >= 1 KB are gathered with cp.async.bulk into a per-warp shared-memory ring of
 * this many rows, 4..32; 0 = automatic, -1 = per-lane 128-bit loads instead),
 * "stage_ahead" (rows beyond that ring sent for with a bulk L2 prefetch, 0..31; -1 = automatic),
 * "hash_bits" (visited hash entries: 0 = automatic, 16 = quotiented 16-bit entries wherever the id range allows,
 * 32 = plain ids), "gang" (warps working on one query / one insert when a batch is smaller than the GPU:
 * 0 = automatic, 1 = never, 2 or 4),
 * "row_floats" (stride of a vector row in floats, a multiple of 4 >= dim; default dim rounded up to
 * 4; only on an empty index).  Apart from the three batch-schedule parameters of the build (build_batch, build_ratio*,
 * build_mates — a batched build is checked by recall, not edge by edge) none of them changes a result: only where
 * data sits and who computes it. */
int hnswb200_set_param(hnswb200_index* idx, const char* name, int64_t value);
int hnswb200_destroy(hnswb200_index* idx);

/* ---- build --------------------------------------------------------------------------------- */

/* Replaces Ohnsw.build_batch_bigarray (lib/ohnsw.ml:840-857) and Hnsw.Ba.build
 * (lib/hnsw.ml:753-761): index `n` vectors in row order.  `levels` (int32[n], may be NULL) are
 * the per-node levels the reference would have drawn (lib/ohnsw.ml:781) — supply them to
 * compare against a reference/oracle build on identical levels. */
int hnswb200_build(hnswb200_index* idx, const float* data, int64_t n, const int32_t* levels);

/* Replaces repeated Ohnsw.insert (lib/ohnsw.ml:766-837) / Hnsw_algo.BuildIncr.insert
 * (lib/hnsw_algo.ml:901): append `n` more vectors to an existing index. */
int hnswb200_insert(hnswb200_index* idx, const float* data, int64_t n, const int32_t* levels);

/* ---- query --------------------------------------------------------------------------------- */

/* Replaces Ohnsw.knn_batch_bigarray (lib/ohnsw.ml:877-897) and Hnsw.Ba.knn_batch
 * (lib/hnsw.ml:769-777); with nq = 1, Ohnsw.knn (lib/ohnsw.ml:859-875) / Hnsw.Ba.knn.
 * `ef` is the beam width: the reference's path B has no separate ef (search_k is called with
 * k, lib/ohnsw.ml:873), so pass ef = k for the literal behaviour, or ef > k for "search with
 * ~k:ef, keep the first k rows".  ids may be NULL (Hnsw.Ba.knn_batch returns distances only).
 * Fails with EINVAL "knn: empty hgraph" on an empty index (lib/ohnsw.ml:862). */
int hnswb200_search(hnswb200_index* idx, const float* queries, int64_t nq, int k, int ef, int mode,
                    int32_t* ids, float* dists);

/* Same, all buffers already in this index's device memory: queries `float[nq][dim]` dense (any dim),
 * outputs `[nq][k]`.  `stream` is a cudaStream_t (NULL = the index's own stream; the call then
 * synchronises before returning, otherwise it only enqueues).  `d_queries` may also be a PINNED host buffer
 * (hnswb200_host_register) here and in the _multi / _sharded variants: the kernel then reads every query from
 * host memory when a warp starts on it, no copy. */
int hnswb200_search_device(hnswb200_index* idx, const float* d_queries, int64_t nq, int k, int ef,
                           int mode, int32_t* d_ids, float* d_dists, void* stream);

/* Multi-GPU variant: the result rows are stored into `n_out` (1..8) destinations, each `[nq][k]`,
 * any of which may be peer-mapped memory of another GPU (symmetric memory over NVLink): every
 * rank writes its block straight into every peer's gather buffer, so no all-gather follows the
 * search — only a barrier and hnswb200_merge_topk_device. */
int hnswb200_search_device_multi(hnswb200_index* idx, const float* d_queries, int64_t nq, int k, int ef,
                                 int mode, int n_out, int32_t* const* d_ids_list,
                                 float* const* d_dists_list, void* stream);

/* Per-query work counters of the last search call: uint32[nq][3] = n_dist, n_exp0, n_expU. */
int hnswb200_last_search_counters(hnswb200_index* idx, uint32_t* out, int64_t nq);

/* ---- graph exchange (the parity vehicle) ------------------------------------ */

/* Load a graph built elsewhere (the reference through an exporter functor over
 * Hnsw_algo.KNN_HGRAPH / a walker over Ohnsw.Hgraph.t, or the oracle): per layer l in
 * 0..max_layer a CSR (`layer_offsets[l]` int64[n+1], `layer_nbrs[l]` int32[nnz_l]) whose rows
 * keep the reference's list order (head first, lib/ohnsw.ml:116-124).  Ids in the CSR and
 * `entry` are offset by `id_base` (0 for Ohnsw, 1 for Hnsw.Ba). */
int hnswb200_import_graph(hnswb200_index* idx, const float* data, int64_t n, int id_base,
                          int max_layer, int64_t entry, const int64_t* const* layer_offsets,
                          const int32_t* const* layer_nbrs);

/* Two-call pattern: with nbrs == NULL only *nnz is written.  offsets is int64[n+1]. */
int hnswb200_export_layer(hnswb200_index* idx, int layer, int id_base, int64_t* offsets,
                          int32_t* nbrs, int64_t* nnz);
/* int32[n]: highest layer each node has adjacency rows on. */
int hnswb200_export_levels(hnswb200_index* idx, int32_t* levels);

/* ---- evaluation helpers (benchmark/dataset.ml) ----------------------------------------------- */

/* Replaces brute_force_knn_l2 (benchmark/dataset.ml:15-30): exact k nearest of every query,
 * ascending by (distance, id); the reference returns distances only, ids are extra. */
int hnswb200_bruteforce_knn(const float* data, int64_t n, const float* queries, int64_t nq, int dim,
                            int k, int metric, int device, int32_t* ids, float* dists);

/* How the last hnswb200_bruteforce_knn call on this thread ran: -1 = fp32 CUDA-core kernel only;
 * >= 0 = tensor-core ranking + exact fp32 re-rank, the value being the number of queries whose
 * result it could not prove exact (when not zero the whole batch was recomputed in fp32). */
int64_t hnswb200_bruteforce_last_unproven(void);

/* Replaces Recall.compute (benchmark/dataset.ml:105-127) on `[nq][k]` arrays. */
int hnswb200_recall(const float* expected, const float* got, int64_t nq, int k, double epsilon,
                    double* out);

/* ---- multi-GPU: per-shard top-k merge ------------------------------------------ */

/* After an all-gather of per-shard results: d_ids/d_dists are `[n_shards][nq][k]` in device
 * memory, ids local to each shard, consecutive shards `shard_stride` elements apart (0 = nq * k;
 * 2 * nq * k when every rank contributes one packed `[ids | dists]` block to a single
 * all-gather); `shard_offsets` (host, int64[n_shards], NULL = all zero) is the first global row
 * of each shard.  Merged into the k best per query `[nq][k]` with global ids, ascending by
 * (distance, id), -1/NaN padded.  `stream` NULL = default stream, synchronous. */
int hnswb200_merge_topk_device(const int32_t* d_ids, const float* d_dists, int n_shards, int64_t nq,
                               int k, int64_t shard_stride, const int64_t* shard_offsets,
                               int32_t* d_out_ids, float* d_out_dists, void* stream);

/* ---- multi-GPU inside the library: one process drives every GPU --------------- */

/* The reference builds / queries everything with ONE call (lib/ohnsw.ml:840-841, :877); a host that is
 * not a torchrun job (the OCaml benchmark, examples/benchmark_c.c) reaches N GPUs through this handle.
 * The dataset is cut into `n_shards` contiguous row ranges [i*n/S, (i+1)*n/S); shard i lives on CUDA
 * device devices[i] (NULL = 0, 1, 2, ...; a device may be named more than once) and is an ordinary
 * hnswb200_index.  No torch, no NCCL: a search copies the queries to the first shard's device once and
 * from there GPU to GPU, every shard's kernel stores its finished rows (ids already global) straight
 * into a gather block on that device through the NVLink peer mapping, and the last warp to arrive
 * for a query merges its rows in place — no all-gather, no barrier kernel, no merge launch. */
typedef struct hnswb200_sharded hnswb200_sharded;
int hnswb200_sharded_create(hnswb200_sharded** out, int dim, int metric, int M, int ef_construction,
                            uint64_t seed, int n_shards, const int* devices);
int hnswb200_sharded_destroy(hnswb200_sharded* s);
/* hnswb200_set_param / hnswb200_set_flavour on every shard; plus "query_path" (0: one H2D copy then GPU-to-GPU
 * copies, the default; 1: one H2D copy per device). */
int hnswb200_sharded_set_param(hnswb200_sharded* s, const char* name, int64_t value);
int hnswb200_sharded_set_flavour(hnswb200_sharded* s, int flavour);
/* Ohnsw.build_batch_bigarray (lib/ohnsw.ml:840-857) over all shards at once, one host thread per shard.
 * `levels` (int32[n], may be NULL) as in hnswb200_build. */
int hnswb200_sharded_build(hnswb200_sharded* s, const float* data, int64_t n, const int32_t* levels);
/* Ohnsw.knn_batch_bigarray (lib/ohnsw.ml:877-897): host buffers, global ids, rows ascending by
 * (distance, id) over all shards, -1 / NaN padded.  Pinned buffers (hnswb200_host_register) are not copied:
 * every shard's kernel reads the batch from host memory over its own PCIe link, and the warp that merges a
 * query stores its row into the caller's arrays. */
int hnswb200_sharded_search(hnswb200_sharded* s, const float* queries, int64_t nq, int k, int ef, int mode,
                            int32_t* ids, float* dists);
/* Same with every buffer on the FIRST shard's device (queries dense [nq][dim]); `stream` as in
 * hnswb200_search_device (a stream of that device; NULL = synchronous). */
int hnswb200_sharded_search_device(hnswb200_sharded* s, const float* d_queries, int64_t nq, int k, int ef, int mode,
                                   int32_t* d_ids, float* d_dists, void* stream);
/* Shard i as a plain index (import / export / stats / per-query counters) and its first global row. */
int hnswb200_sharded_shard(hnswb200_sharded* s, int i, hnswb200_index** out, int64_t* first_row);
/* n = total rows; max_layer = the highest over the shards; entry_point = -1 (per shard). */
int hnswb200_sharded_get_info(hnswb200_sharded* s, hnswb200_info* out, int* n_shards);
/* Counters summed over the shards; search_kernel_ms = device time of the last search step (first query
 * copy to the last shard's kernel end, merge included); build_seconds = the slowest shard's. */
int hnswb200_sharded_get_stats(hnswb200_sharded* s, hnswb200_stats* out);

/* One process per GPU (torchrun) with caller-provided peer-mapped buffers (torch symmetric memory): shard
 * `shard` of `n_shards` searches its own index `idx` and takes part in the same fused exchange + merge.
 * g_ids / g_dists ([n_shards][nq][k]) and arrive (uint32[nq], zero before the call) live on one home rank
 * and are mapped on every rank; the merged rows are stored into each of the n_final (1..8) destinations.
 * The caller separates consecutive calls that reuse `arrive` with a barrier. */
int hnswb200_search_device_sharded(hnswb200_index* idx, const float* d_queries, int64_t nq, int k, int ef, int mode,
                                   int shard, int n_shards, int64_t first_row, int32_t* g_ids, float* g_dists,
                                   uint32_t* arrive, int n_final, int32_t* const* f_ids, float* const* f_dists,
                                   void* stream);

/* ---- misc ------------------------------------------------------------------------------------ */

int hnswb200_get_info(hnswb200_index* idx, hnswb200_info* out);
int hnswb200_get_stats(hnswb200_index* idx, hnswb200_stats* out);
/* Pin / unpin a caller buffer (a Bigarray payload).  Copies from / to it are asynchronous DMA, and hnswb200_search
 * does not copy it at all: the search kernel reads each query from it when a warp starts on that query and stores
 * each result row into it when the query is done, so the PCIe traffic rides under the search.  Register whole
 * buffers (the address range the call touches must lie inside one registration). */
int hnswb200_host_register(const void* ptr, int64_t bytes);
int hnswb200_host_unregister(const void* ptr);
const char* hnswb200_last_error(void);
const char* hnswb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HNSW_B200_H */
