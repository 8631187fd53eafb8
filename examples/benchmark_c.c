/* benchmark_c.c — the recipe of the reference's benchmark/benchmark.ml:76-99 (build an index on a
 * random dataset, query it, compare with the exact neighbours, print recall) written in plain C
 * against include/hnsw_b200.h.  It is the C a maintainer's OCaml stubs boil down to: no torch, no
 * Python, pointers and sizes only.
 *
 *   gcc -O2 -Iinclude examples/benchmark_c.c -Locaml-hnsw_b200 -lhnsw_b200 \
 *       -Wl,-rpath,$PWD/ocaml-hnsw_b200 -lm -o build/benchmark_c
 *   build/benchmark_c [--gpus N | --shards S] [n] [dim] [nq] [M] [efC] [k] [ef]
 *
 * --gpus N cuts the rows into N shards, one per GPU (hnswb200_sharded_*: one process drives every GPU,
 * per-shard rows exchanged by peer stores and merged inside the search kernel); --shards S places S
 * shards on device 0 (the same kernels on a one-GPU box).
 *
 * Exit status: 0 on success, 3 when the library reports an error (e.g. no CUDA device: there is
 * no CPU fallback), 4 when recall is below 0.9. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "hnsw_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != HNSWB200_OK) {                                                        \
      fprintf(stderr, "%s -> status %d: %s\n", #call, rc_, hnswb200_last_error());   \
      return 3;                                                                      \
    }                                                                                \
  } while (0)

/* xorshift64*: uniform [-1, 1), the range of Lacaml.S.Mat.random (benchmark/dataset.ml:48) */
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static float uniform_pm1(void) {
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) * (2.0f / 16777216.0f) - 1.0f;
}

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char** argv) {
  int shards = 0, spread = 0;                          /* 0: a single hnswb200_index */
  if (argc > 2 && (!strcmp(argv[1], "--gpus") || !strcmp(argv[1], "--shards"))) {
    spread = !strcmp(argv[1], "--gpus");
    shards = atoi(argv[2]);
    argv += 2; argc -= 2;
  }
  int64_t n = argc > 1 ? atoll(argv[1]) : 10000;       /* benchmark.ml's default shape: 10k x 128 */
  int dim = argc > 2 ? atoi(argv[2]) : 128;
  int64_t nq = argc > 3 ? atoll(argv[3]) : 10;         /* benchmark.ml queries 10 points */
  int M = argc > 4 ? atoi(argv[4]) : 16;
  int efc = argc > 5 ? atoi(argv[5]) : 100;
  int k = argc > 6 ? atoi(argv[6]) : 10;
  int ef = argc > 7 ? atoi(argv[7]) : 50;

  /* low intrinsic dimension (8 latent coordinates) so that recall is a meaningful check */
  float* basis = malloc(sizeof(float) * 8 * dim);
  float* train = malloc(sizeof(float) * n * dim);
  float* test = malloc(sizeof(float) * nq * dim);
  for (int i = 0; i < 8 * dim; ++i) basis[i] = uniform_pm1();
  for (int64_t r = 0; r < n + nq; ++r) {
    float z[8];
    float* row = r < n ? train + r * dim : test + (r - n) * dim;
    for (int j = 0; j < 8; ++j) z[j] = uniform_pm1();
    for (int d = 0; d < dim; ++d) {
      float s = 0.05f * uniform_pm1();
      for (int j = 0; j < 8; ++j) s += z[j] * basis[j * dim + d];
      row[d] = s;
    }
  }

  printf("%s\n", hnswb200_version());
  hnswb200_index* h = NULL;
  hnswb200_sharded* sh = NULL;
  int32_t* ids = malloc(sizeof(int32_t) * nq * k);
  float* dists = malloc(sizeof(float) * nq * k);
  float* exact = malloc(sizeof(float) * nq * k);
  double t0, t_build, t_search;
  if (shards > 0) {
    int devices[32];
    if (shards > 32) shards = 32;
    for (int i = 0; i < shards; ++i) devices[i] = spread ? i : 0;
    CHECK(hnswb200_sharded_create(&sh, dim, HNSWB200_L2, M, efc, /*seed*/ 0, shards, devices));
    t0 = now();
    CHECK(hnswb200_sharded_build(sh, train, n, NULL));
    t_build = now() - t0;
    CHECK(hnswb200_sharded_search(sh, test, nq, k, ef, HNSWB200_MODE_PARITY, ids, dists));   /* warm-up: buffers, peer mappings */
    t0 = now();
    CHECK(hnswb200_sharded_search(sh, test, nq, k, ef, HNSWB200_MODE_PARITY, ids, dists));
    t_search = now() - t0;
  } else {
    CHECK(hnswb200_create(&h, dim, HNSWB200_L2, M, efc, /*seed*/ 0, /*device*/ 0));
    t0 = now();
    CHECK(hnswb200_build(h, train, n, NULL));              /* Ohnsw.build_batch_bigarray, benchmark.ml:76 */
    t_build = now() - t0;
    t0 = now();
    CHECK(hnswb200_search(h, test, nq, k, ef, HNSWB200_MODE_PARITY, ids, dists));   /* knn_batch_bigarray, :91 */
    t_search = now() - t0;
  }
  CHECK(hnswb200_bruteforce_knn(train, n, test, nq, dim, k, HNSWB200_L2, 0, NULL, exact));   /* dataset.ml:15 */
  double recall = 0;
  CHECK(hnswb200_recall(exact, dists, nq, k, 1e-8, &recall));                     /* dataset.ml:105 */

  hnswb200_info inf;
  hnswb200_stats st;
  if (sh) {
    int ns = 0;
    CHECK(hnswb200_sharded_get_info(sh, &inf, &ns));
    CHECK(hnswb200_sharded_get_stats(sh, &st));
    printf("%d shards on %s\n", ns, spread ? "one GPU each" : "device 0");
  } else {
    CHECK(hnswb200_get_info(h, &inf));
    CHECK(hnswb200_get_stats(h, &st));
  }
  printf("n=%lld dim=%d M=%d efC=%d: build %.3f s, max_layer %d, entry %lld\n", (long long)inf.n, inf.dim, inf.M,
         inf.ef_construction, t_build, inf.max_layer, (long long)inf.entry_point);
  printf("nq=%lld k=%d ef=%d: search %.3f ms (kernel %.3f ms), %.1f distance evaluations per query\n", (long long)nq, k, ef,
         1e3 * t_search, st.search_kernel_ms, (double)st.search_n_dist / (double)nq);
  printf("first query: ");
  for (int i = 0; i < k; ++i) printf("%d:%.4f ", ids[i], dists[i]);
  printf("\nrecall %.4f\n", recall);
  if (sh) CHECK(hnswb200_sharded_destroy(sh));
  else CHECK(hnswb200_destroy(h));
  free(basis); free(train); free(test); free(ids); free(dists); free(exact);
  return recall >= 0.9 ? 0 : 4;
}
