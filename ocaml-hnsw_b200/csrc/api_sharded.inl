// C ABI: a row-sharded index over several GPUs of one box, driven from ONE process (included by
// hnsw_b200.cu).  SURVEY.md 8e: contiguous row shards, each GPU builds and searches its own sub-graph,
// queries go to every shard, per-shard top-k rows are merged.  No torch, no NCCL: shards are plain
// hnswb200_index handles; the exchange is peer stores over NVLink into a gather block on the home
// device and the merge is done by the last warp to arrive (search.cuh, ShardTail).  One host thread
// per shard launches its device's work, so N devices start within microseconds of each other.
#include <condition_variable>
#include <functional>
#include <thread>

namespace {

struct Worker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::function<void()> job;
  bool has_job = false, stop = false, done = true;
  int err_code = 0;
  std::string err;
  void loop() {
    std::unique_lock<std::mutex> lk(m);
    while (true) {
      cv.wait(lk, [&] { return has_job || stop; });
      if (stop) return;
      std::function<void()> f = std::move(job);
      has_job = false;
      lk.unlock();
      int code = 0; std::string msg;
      try { f(); }
      catch (const HbError& e) { code = e.code; msg = e.what(); }
      catch (const std::bad_alloc&) { code = HNSWB200_ENOMEM; msg = "out of memory"; }
      catch (const std::exception& e) { code = HNSWB200_ECUDA; msg = e.what(); }
      lk.lock();
      err_code = code; err = msg; done = true;
      cv.notify_all();
    }
  }
  void post(std::function<void()> f) {
    std::lock_guard<std::mutex> lk(m);
    job = std::move(f); has_job = true; done = false; err_code = 0;
    cv.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return done; });
  }
};

}  // namespace

struct hnswb200_sharded {
  int n_shards = 0, dim = 0, ld = 0;
  std::vector<hnswb200_index*> shard;
  std::vector<int> device;
  std::vector<int64_t> offset;             // first global row of each shard (n_shards + 1 entries)
  std::vector<std::unique_ptr<Worker>> worker;
  std::vector<cudaStream_t> stream;        // per shard, on its device
  std::vector<cudaEvent_t> done;           // per shard: its search of the current call has been enqueued up to here
  std::vector<float*> q_local;             // per shard: the padded queries on its device (null on the home device)
  std::vector<size_t> q_local_n;
  std::vector<char> peer_ok;               // per shard: its device can store into the home device
  int home = 0;                            // device of shard 0: gather block, arrival counters, merged rows
  cudaStream_t hs = nullptr;               // home stream
  cudaEvent_t ev_q = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  DevBuf<float> q_home, out_dists, g_dists;
  DevBuf<int32_t> out_ids, g_ids;
  DevBuf<unsigned int> arrive;             // 2 x nq (double buffered: the other half is cleared for the next call)
  size_t arrive_half = 0;
  uint64_t calls = 0;
  int64_t param_query_path = 0;            // 0: H2D once to the home device, then GPU-to-GPU copies; 1: one H2D per device
  double last_search_ms = 0;
  std::mutex mu;
};

namespace {

void sharded_run(hnswb200_sharded* s, const std::function<void(int)>& f, int first = 0) {
  for (int i = first; i < s->n_shards; i++) s->worker[i]->post([=] { f(i); });
  int code = 0; std::string msg;
  for (int i = first; i < s->n_shards; i++) {
    s->worker[i]->wait();
    if (s->worker[i]->err_code && !code) { code = s->worker[i]->err_code; msg = "shard " + std::to_string(i) + ": " + s->worker[i]->err; }
  }
  if (code) fail(code, msg);
}

void sharded_destroy(hnswb200_sharded* s) {
  for (auto& w : s->worker) {
    if (!w) continue;
    { std::lock_guard<std::mutex> lk(w->m); w->stop = true; w->cv.notify_all(); }
    if (w->th.joinable()) w->th.join();
  }
  for (int i = 0; i < (int)s->shard.size(); i++) {
    if (i < (int)s->device.size()) cudaSetDevice(s->device[i]);
    if (i < (int)s->q_local.size() && s->q_local[i]) cudaFree(s->q_local[i]);
    if (i < (int)s->stream.size() && s->stream[i]) { cudaStreamSynchronize(s->stream[i]); cudaStreamDestroy(s->stream[i]); }
    if (i < (int)s->done.size() && s->done[i]) cudaEventDestroy(s->done[i]);
    if (s->shard[i]) hnswb200_destroy(s->shard[i]);
  }
  cudaSetDevice(s->home);
  if (s->hs) { cudaStreamSynchronize(s->hs); cudaStreamDestroy(s->hs); }
  for (cudaEvent_t e : {s->ev_q, s->ev_t0, s->ev_t1}) if (e) cudaEventDestroy(e);
  s->q_home.release(); s->out_dists.release(); s->g_dists.release(); s->out_ids.release(); s->g_ids.release(); s->arrive.release();
  delete s;
}

// The search of one batch on every shard.  Queries: `h_queries` (host, dense [nq][dim]) or `d_queries`
// (home device, dense [nq][dim], ordered after `after` on the home device).  The merged rows land in
// f_ids / f_dists on the home device; `hs` is ordered after every shard's kernel on return.
void sharded_enqueue(hnswb200_sharded* s, const float* h_queries, const float* d_queries, int64_t nq, int k, int ef, int mode,
                     int32_t* f_ids, float* f_dists) {
  CUDA_CHECK(cudaSetDevice(s->home));
  const int S = s->n_shards;
  const size_t nqk = (size_t)nq * k;
  s->g_ids.reserve((size_t)S * nqk); s->g_dists.reserve((size_t)S * nqk);
  if ((size_t)nq > s->arrive_half) {
    s->arrive.release(); s->arrive.reserve(2 * (size_t)nq); s->arrive_half = (size_t)nq;
    CUDA_CHECK(cudaMemsetAsync(s->arrive.p, 0, 2 * (size_t)nq * sizeof(unsigned int), s->hs));
    s->calls = 0;
  }
  unsigned int* arrive = s->arrive.p + (s->calls & 1) * s->arrive_half;
  unsigned int* arrive_next = s->arrive.p + ((s->calls + 1) & 1) * s->arrive_half;
  s->calls++;
  // queries on the home device, padded to the row stride
  const float* q_home;
  // a pinned host batch is read in place by every shard's kernel (each GPU over its own PCIe link): no copy at all
  const float* q_pinned = h_queries && s->ld == s->dim && s->param_query_path != 1 && s->shard[0]->param_host_zero_copy != 0
                              ? device_view_of_pinned(h_queries, (size_t)nq * s->dim) : nullptr;
  if (q_pinned) q_home = q_pinned;
  else if (h_queries) {
    s->q_home.reserve((size_t)nq * s->ld);
    upload_rows(s->q_home.p, s->ld, h_queries, s->dim, nq, s->hs);
    q_home = s->q_home.p;
  } else if (s->ld != s->dim) {
    s->q_home.reserve((size_t)nq * s->ld);
    CUDA_CHECK(cudaMemsetAsync(s->q_home.p, 0, (size_t)nq * s->ld * sizeof(float), s->hs));
    CUDA_CHECK(cudaMemcpy2DAsync(s->q_home.p, (size_t)s->ld * sizeof(float), d_queries, (size_t)s->dim * sizeof(float),
                                 (size_t)s->dim * sizeof(float), (size_t)nq, cudaMemcpyDeviceToDevice, s->hs));
    q_home = s->q_home.p;
  } else q_home = d_queries;
  CUDA_CHECK(cudaMemsetAsync(arrive_next, 0, (size_t)nq * sizeof(unsigned int), s->hs));   // ready for the next call
  CUDA_CHECK(cudaEventRecord(s->ev_t0, s->hs));
  CUDA_CHECK(cudaEventRecord(s->ev_q, s->hs));
  const bool per_device_h2d = h_queries && s->param_query_path == 1;
  auto one = [&](int i) {
    hnswb200_index* x = s->shard[i];
    std::lock_guard<std::mutex> lk(x->mu);
    use_device(x);
    cudaStream_t st = s->stream[i];
    const float* q = q_home;
    CUDA_CHECK(cudaStreamWaitEvent(st, s->ev_q, 0));
    if (s->device[i] != s->home && !q_pinned) {
      const size_t want = (size_t)nq * s->ld;
      if (want > s->q_local_n[i]) {
        if (s->q_local[i]) cudaFree(s->q_local[i]);
        s->q_local[i] = nullptr; s->q_local_n[i] = 0;
        CUDA_CHECK(cudaMalloc(&s->q_local[i], want * sizeof(float)));
        s->q_local_n[i] = want;
      }
      if (per_device_h2d) upload_rows(s->q_local[i], s->ld, h_queries, s->dim, nq, st);
      else CUDA_CHECK(cudaMemcpyPeerAsync(s->q_local[i], s->device[i], q_home, s->home, want * sizeof(float), st));
      q = s->q_local[i];
    }
    hb::ShardTail t{};
    t.n_shards = S; t.shard = i; t.id_offset = (int32_t)s->offset[i];
    t.g_ids = s->g_ids.p; t.g_dists = s->g_dists.p; t.arrive = arrive;
    t.n_final = 1; t.f_ids[0] = f_ids; t.f_dists[0] = f_dists;
    search_device(x, q, nq, k, ef, mode, nullptr, nullptr, st, false, 0, nullptr, nullptr, &t);
    CUDA_CHECK(cudaEventRecord(s->done[i], st));
  };
  // shard 0 is launched by this thread, the others by their own
  for (int i = 1; i < S; i++) s->worker[i]->post([=] { one(i); });
  int code = 0; std::string msg;
  try { one(0); } catch (const HbError& e) { code = e.code; msg = std::string("shard 0: ") + e.what(); }
  for (int i = 1; i < S; i++) {
    s->worker[i]->wait();
    if (s->worker[i]->err_code && !code) { code = s->worker[i]->err_code; msg = "shard " + std::to_string(i) + ": " + s->worker[i]->err; }
  }
  CUDA_CHECK(cudaSetDevice(s->home));
  if (code) { cudaDeviceSynchronize(); fail(code, msg); }
  for (int i = 0; i < S; i++) CUDA_CHECK(cudaStreamWaitEvent(s->hs, s->done[i], 0));
  CUDA_CHECK(cudaEventRecord(s->ev_t1, s->hs));
}

void sharded_check(hnswb200_sharded* s, int64_t nq, int k) {
  if (!s) fail(HNSWB200_EINVAL, "sharded index is NULL");
  if (nq < 0 || k <= 0) fail(HNSWB200_EINVAL, "search: nq must be >= 0 and k > 0");
  if (s->offset.empty() || s->offset[s->n_shards] == 0) fail(HNSWB200_EINVAL, "knn: empty hgraph");     // lib/ohnsw.ml:862
}

}  // namespace

extern "C" {

int hnswb200_sharded_create(hnswb200_sharded** out, int dim, int metric, int M, int ef_construction, uint64_t seed,
                            int n_shards, const int* devices) {
  return guard([&] {
    if (!out) fail(HNSWB200_EINVAL, "sharded_create: out is NULL");
    *out = nullptr;
    if (n_shards < 1 || n_shards > 32) fail(HNSWB200_EINVAL, "sharded_create: n_shards must be in 1..32");
    hnswb200_sharded* s = new hnswb200_sharded();
    try {
      s->n_shards = n_shards; s->dim = dim;
      s->shard.assign(n_shards, nullptr); s->device.resize(n_shards); s->stream.assign(n_shards, nullptr);
      s->done.assign(n_shards, nullptr); s->q_local.assign(n_shards, nullptr); s->q_local_n.assign(n_shards, 0);
      s->peer_ok.assign(n_shards, 1);
      for (int i = 0; i < n_shards; i++) {
        s->device[i] = devices ? devices[i] : i;
        int rc = hnswb200_create(&s->shard[i], dim, metric, M, ef_construction, seed + (uint64_t)i, s->device[i]);
        if (rc != HNSWB200_OK) fail(rc, "shard " + std::to_string(i) + ": " + g_err);
      }
      s->ld = s->shard[0]->ld;
      s->home = s->device[0];
      for (int i = 0; i < n_shards; i++) {
        CUDA_CHECK(cudaSetDevice(s->device[i]));
        CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream[i], cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming));
        if (s->device[i] != s->home) {
          int can = 0;
          CUDA_CHECK(cudaDeviceCanAccessPeer(&can, s->device[i], s->home));
          if (!can) fail(HNSWB200_ECUDA, "device " + std::to_string(s->device[i]) + " cannot store into device " + std::to_string(s->home) +
                                             " (no peer access): the sharded index needs NVLink / PCIe peer mapping");
          cudaError_t e = cudaDeviceEnablePeerAccess(s->home, 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CUDA_CHECK(e);
        }
      }
      CUDA_CHECK(cudaSetDevice(s->home));
      for (int i = 0; i < n_shards; i++)          // the home device reads the other devices' query copies only through copies; it stores into none
        if (s->device[i] != s->home) {
          cudaError_t e = cudaDeviceEnablePeerAccess(s->device[i], 0);
          if (e == cudaErrorPeerAccessAlreadyEnabled || e == cudaErrorPeerAccessUnsupported) cudaGetLastError(); else CUDA_CHECK(e);
        }
      CUDA_CHECK(cudaStreamCreateWithFlags(&s->hs, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&s->ev_q, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreate(&s->ev_t0));
      CUDA_CHECK(cudaEventCreate(&s->ev_t1));
      s->offset.assign(n_shards + 1, 0);
      for (int i = 0; i < n_shards; i++) {
        s->worker.emplace_back(new Worker());
        Worker* w = s->worker.back().get();
        w->th = std::thread([w] { w->loop(); });
      }
    } catch (...) { sharded_destroy(s); throw; }
    *out = s;
  });
}

int hnswb200_sharded_destroy(hnswb200_sharded* s) {
  return guard([&] { if (s) sharded_destroy(s); });
}

int hnswb200_sharded_set_param(hnswb200_sharded* s, const char* name, int64_t value) {
  return guard([&] {
    if (!s || !name) fail(HNSWB200_EINVAL, "sharded index or name is NULL");
    if (std::string(name) == "query_path") { s->param_query_path = value; return; }
    for (int i = 0; i < s->n_shards; i++) {
      int rc = hnswb200_set_param(s->shard[i], name, value);
      if (rc != HNSWB200_OK) fail(rc, g_err);
    }
    s->ld = s->shard[0]->ld;
  });
}

int hnswb200_sharded_set_flavour(hnswb200_sharded* s, int flavour) {
  return guard([&] {
    if (!s) fail(HNSWB200_EINVAL, "sharded index is NULL");
    for (int i = 0; i < s->n_shards; i++) {
      int rc = hnswb200_set_flavour(s->shard[i], flavour);
      if (rc != HNSWB200_OK) fail(rc, g_err);
    }
  });
}

int hnswb200_sharded_shard(hnswb200_sharded* s, int i, hnswb200_index** out, int64_t* first_row) {
  return guard([&] {
    if (!s || !out) fail(HNSWB200_EINVAL, "sharded index or out is NULL");
    if (i < 0 || i >= s->n_shards) fail(HNSWB200_EINVAL, "sharded_shard: no such shard");
    *out = s->shard[i];
    if (first_row) *first_row = s->offset[i];
  });
}

int hnswb200_sharded_build(hnswb200_sharded* s, const float* data, int64_t n, const int32_t* levels) {
  return guard([&] {
    if (!s) fail(HNSWB200_EINVAL, "sharded index is NULL");
    if (n < s->n_shards) fail(HNSWB200_EINVAL, "sharded_build: fewer rows than shards");
    if (!data) fail(HNSWB200_EINVAL, "build: data is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->offset[s->n_shards] != 0) fail(HNSWB200_EINVAL, "build: the index is not empty");
    auto t0 = std::chrono::steady_clock::now();
    std::vector<int64_t> off(s->n_shards + 1);
    for (int i = 0; i <= s->n_shards; i++) off[i] = n * i / s->n_shards;       // contiguous row ranges
    sharded_run(s, [&](int i) {
      hnswb200_index* x = s->shard[i];
      std::lock_guard<std::mutex> lk2(x->mu);
      use_device(x);
      if (x->n != 0) fail(HNSWB200_EINVAL, "build: the index is not empty (use insert)");
      x->slots0 = 2 * x->M; x->slotsU = x->M;
      append_nodes(x, data + (size_t)off[i] * s->dim, off[i + 1] - off[i], levels ? levels + off[i] : nullptr);
    });
    s->offset = off;
    (void)t0;
  });
}

int hnswb200_sharded_search(hnswb200_sharded* s, const float* queries, int64_t nq, int k, int ef, int mode, int32_t* ids, float* dists) {
  return guard([&] {
    sharded_check(s, nq, k);
    if (nq == 0) return;
    if (!queries || !dists) fail(HNSWB200_EINVAL, "search: queries/dists is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    CUDA_CHECK(cudaSetDevice(s->home));
    // pinned result buffers take the merged rows straight from the warp that merges them (whichever GPU it runs on)
    const bool zc = s->shard[0]->param_host_zero_copy != 0;
    int32_t* ids_map = zc ? device_view_of_pinned(ids, (size_t)nq * k) : nullptr;
    float* d_map = zc ? device_view_of_pinned(dists, (size_t)nq * k) : nullptr;
    if (!ids_map) s->out_ids.reserve((size_t)nq * k);
    if (!d_map) s->out_dists.reserve((size_t)nq * k);
    sharded_enqueue(s, queries, nullptr, nq, k, ef, mode, ids_map ? ids_map : s->out_ids.p, d_map ? d_map : s->out_dists.p);
    if (ids && !ids_map) CUDA_CHECK(cudaMemcpyAsync(ids, s->out_ids.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s->hs));
    if (!d_map) CUDA_CHECK(cudaMemcpyAsync(dists, s->out_dists.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s->hs));
    CUDA_CHECK(cudaStreamSynchronize(s->hs));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1) == cudaSuccess) s->last_search_ms = ms; else cudaGetLastError();
  });
}

int hnswb200_sharded_search_device(hnswb200_sharded* s, const float* d_queries, int64_t nq, int k, int ef, int mode,
                                   int32_t* d_ids, float* d_dists, void* stream) {
  return guard([&] {
    sharded_check(s, nq, k);
    if (nq == 0) return;
    if (!d_queries || !d_ids || !d_dists) fail(HNSWB200_EINVAL, "search_device: NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    CUDA_CHECK(cudaSetDevice(s->home));
    cudaStream_t cs = (cudaStream_t)stream;
    if (cs) {                                             // the home stream continues after the caller's work ...
      CUDA_CHECK(cudaEventRecord(s->ev_q, cs));
      CUDA_CHECK(cudaStreamWaitEvent(s->hs, s->ev_q, 0));
    } else CUDA_CHECK(cudaDeviceSynchronize());
    sharded_enqueue(s, nullptr, d_queries, nq, k, ef, mode, d_ids, d_dists);
    if (cs) {                                             // ... and the caller's stream after the merged rows
      CUDA_CHECK(cudaEventRecord(s->ev_q, s->hs));
      CUDA_CHECK(cudaStreamWaitEvent(cs, s->ev_q, 0));
    } else {
      CUDA_CHECK(cudaStreamSynchronize(s->hs));
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, s->ev_t0, s->ev_t1) == cudaSuccess) s->last_search_ms = ms; else cudaGetLastError();
    }
  });
}

int hnswb200_sharded_get_info(hnswb200_sharded* s, hnswb200_info* out, int* n_shards) {
  return guard([&] {
    if (!s || !out) fail(HNSWB200_EINVAL, "sharded index or out is NULL");
    int rc = hnswb200_get_info(s->shard[0], out);
    if (rc != HNSWB200_OK) fail(rc, g_err);
    out->n = s->offset[s->n_shards];
    out->max_layer = 0; out->entry_point = -1;            // per shard: see hnswb200_sharded_shard
    for (int i = 0; i < s->n_shards; i++) out->max_layer = std::max(out->max_layer, s->shard[i]->max_layer);
    if (n_shards) *n_shards = s->n_shards;
  });
}

int hnswb200_sharded_get_stats(hnswb200_sharded* s, hnswb200_stats* out) {
  return guard([&] {
    if (!s || !out) fail(HNSWB200_EINVAL, "sharded index or out is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    hnswb200_stats acc{};
    for (int i = 0; i < s->n_shards; i++) {
      hnswb200_stats st;
      int rc = hnswb200_get_stats(s->shard[i], &st);
      if (rc != HNSWB200_OK) fail(rc, g_err);
      acc.search_queries = st.search_queries;
      acc.search_n_dist += st.search_n_dist; acc.search_n_exp0 += st.search_n_exp0; acc.search_n_expU += st.search_n_expU;
      acc.search_visited_overflows += st.search_visited_overflows; acc.search_tie_overflows += st.search_tie_overflows;
      acc.search_tie_spills += st.search_tie_spills;
      acc.search_algorithmic_bytes += st.search_algorithmic_bytes;
      acc.search_kernel_ms = std::max(acc.search_kernel_ms, st.search_kernel_ms);
      acc.build_inserts += st.build_inserts; acc.build_n_dist += st.build_n_dist; acc.build_n_exp += st.build_n_exp;
      acc.build_algorithmic_bytes += st.build_algorithmic_bytes; acc.build_seconds = std::max(acc.build_seconds, st.build_seconds);
      acc.build_visited_overflows += st.build_visited_overflows; acc.build_dropped_incoming += st.build_dropped_incoming;
      acc.gpu_launches += st.gpu_launches;
      acc.num_layers = std::max(acc.num_layers, st.num_layers);
      for (int l = 0; l < st.num_layers && l < 16; l++) {
        const int64_t before = acc.layer_nodes[l];
        acc.layer_mean_degree[l] = (acc.layer_mean_degree[l] * (double)before + st.layer_mean_degree[l] * (double)st.layer_nodes[l]) /
                                   std::max<double>(1.0, (double)(before + st.layer_nodes[l]));
        acc.layer_min_degree[l] = before ? std::min(acc.layer_min_degree[l], st.layer_min_degree[l]) : st.layer_min_degree[l];
        acc.layer_max_degree[l] = std::max(acc.layer_max_degree[l], st.layer_max_degree[l]);
        acc.layer_nodes[l] += st.layer_nodes[l]; acc.layer_isolated[l] += st.layer_isolated[l];
      }
    }
    CUDA_CHECK(cudaSetDevice(s->home));
    // the step of the last search on the home stream: first query copy to the last shard's kernel end (merge included)
    if (s->last_search_ms > 0) acc.search_kernel_ms = s->last_search_ms;
    *out = acc;
  });
}

}  // extern "C"
