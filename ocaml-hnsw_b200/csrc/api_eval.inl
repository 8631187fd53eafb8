// C ABI: brute-force ground truth and shard merge (included by hnsw_b200.cu).
namespace {
thread_local int64_t g_last_bruteforce_unproven = -1;   // -1: tensor-core path not taken

// HNSWB200_TRACE=1: device time of the brute-force kernels (CUDA events) on stderr
struct EvTimer {
  bool on = std::getenv("HNSWB200_TRACE") != nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  explicit EvTimer(cudaStream_t st) : s(st) { if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, s); } }
  void lap(const char* what) {
    if (!on) return;
    cudaEventRecord(b, s); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    fprintf(stderr, "[hnsw_b200 bruteforce] %s %.3f ms\n", what, ms);
    cudaEventRecord(a, s);
  }
  ~EvTimer() { if (on) { cudaEventDestroy(a); cudaEventDestroy(b); } }
};
void launch_bruteforce(const float* d_data, int64_t n, const float* d_q, int64_t nq, int ld, int k, int metric,
                       int num_sms, int32_t* d_ids, float* d_dists, cudaStream_t s, uint64_t* launches) {
  int k_cap = round_up(k, 32);
  size_t smem = hb::brute_smem_bytes(k_cap);
  int qblocks = (int)((nq + hb::BF_QT - 1) / hb::BF_QT);
  // split the data so the grid covers the SMs about twice; at most 32 slices (one per merge lane)
  int splits = std::max(1, std::min(32, (2 * num_sms + qblocks - 1) / qblocks));
  int64_t split_len = ((n + splits - 1) / splits + hb::BF_XT - 1) / hb::BF_XT * hb::BF_XT;
  splits = (int)((n + split_len - 1) / split_len);
  DevBuf<uint64_t> partial;
  partial.reserve((size_t)splits * nq * k);
  hb::BruteParams p;
  p.data = d_data; p.queries = d_q; p.n = n; p.nq = nq; p.ld = ld; p.k = k; p.k_cap = k_cap; p.metric = metric;
  p.split_len = split_len; p.partial = partial.p;
  CUDA_CHECK(cudaFuncSetAttribute(hb::bruteforce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  EvTimer tm(s);
  hb::bruteforce_kernel<<<dim3(qblocks, splits), hb::BF_THREADS, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
  tm.lap("fp32 bruteforce_kernel");
  int wpb = 8;
  hb::bruteforce_finish_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, s>>>(partial.p, splits, nq, k, metric, d_ids, d_dists);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (launches) *launches += 2;
}
// Tensor-core path (bruteforce_tc.cuh): returns the number of queries whose result could not be
// proven exact (the caller recomputes with the fp32 kernel when that is not zero).
int64_t launch_bruteforce_tc(const float* d_data, int64_t n, const float* d_q, int64_t nq, int ld, int dim, int k,
                             int num_sms, int32_t* d_ids, float* d_dists, cudaStream_t s, uint64_t* launches) {
  const int kp = round_up(dim, hb::TC_KC);
  // round 2: operands pre-tiled for bulk copies + the warp-specialised kernel (bruteforce_tc2.cuh);
  // HNSWB200_BRUTEFORCE=tc1 keeps the round-1 kernel (cp.async by every thread) for comparison
  const char* which = std::getenv("HNSWB200_BRUTEFORCE");
  const bool tc2 = !(which && std::string(which) == "tc1");
  const int64_t n_pad = (n + hb::TC_N - 1) / hb::TC_N * hb::TC_N, nq_pad = (nq + hb::TC_M - 1) / hb::TC_M * hb::TC_M;
  DevBuf<__nv_bfloat16> x_hi, x_lo, q_hi, q_lo;
  DevBuf<float> x_norm, scal, bound;
  DevBuf<int> flags;
  DevBuf<unsigned int> gthr;
  DevBuf<uint64_t> partial;
  x_hi.reserve((size_t)n_pad * kp); x_lo.reserve((size_t)n_pad * kp); q_hi.reserve((size_t)nq_pad * kp); q_lo.reserve((size_t)nq_pad * kp);
  x_norm.reserve((size_t)n); scal.reserve(4); flags.reserve((size_t)nq + 4);
  CUDA_CHECK(cudaMemsetAsync(scal.p, 0, 4 * sizeof(float), s));     // [0] max ||x||^2, [1] any lo (int), [2] max ||q||^2 (unused)
  int wpb = 8;
  EvTimer tm(s);
  if (tc2) {
    for (DevBuf<__nv_bfloat16>* b : {&x_hi, &x_lo}) CUDA_CHECK(cudaMemsetAsync(b->p, 0, (size_t)n_pad * kp * 2, s));
    for (DevBuf<__nv_bfloat16>* b : {&q_hi, &q_lo}) CUDA_CHECK(cudaMemsetAsync(b->p, 0, (size_t)nq_pad * kp * 2, s));
    hb::bf16_tile_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(d_data, ld, dim, n, kp, hb::TC_N, x_hi.p, x_lo.p, x_norm.p,
                                                                             reinterpret_cast<int*>(scal.p + 1), scal.p);
    hb::bf16_tile_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, s>>>(d_q, ld, dim, nq, kp, hb::TC_M, q_hi.p, q_lo.p, nullptr,
                                                                              reinterpret_cast<int*>(scal.p + 1), scal.p + 2);
  } else {
    hb::bf16_split_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(d_data, ld, dim, n, kp, x_hi.p, x_lo.p, x_norm.p,
                                                                              reinterpret_cast<int*>(scal.p + 1), scal.p);
    hb::bf16_split_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, s>>>(d_q, ld, dim, nq, kp, q_hi.p, q_lo.p, nullptr,
                                                                               reinterpret_cast<int*>(scal.p + 1), scal.p + 2);
  }
  CUDA_CHECK(cudaGetLastError());
  tm.lap("bf16_split_kernel x2");
  float h_scal[4];
  CUDA_CHECK(cudaMemcpyAsync(h_scal, scal.p, sizeof(h_scal), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  int any_lo; memcpy(&any_lo, &h_scal[1], 4);

  const int qblocks = (int)((nq + hb::TC_M - 1) / hb::TC_M);
  int splits = std::max(1, std::min(32, (8 * num_sms + qblocks - 1) / qblocks));       // >= ~8 waves of CTAs
  int64_t split_len = ((n + splits - 1) / splits + hb::TC_N - 1) / hb::TC_N * hb::TC_N;
  splits = (int)((n + split_len - 1) / split_len);
  gthr.reserve((size_t)nq);
  CUDA_CHECK(cudaMemsetAsync(gthr.p, 0xff, (size_t)nq * sizeof(unsigned int), s));     // ordered-float +NaN/inf side: no threshold yet
  partial.reserve((size_t)splits * hb::TC_HALVES * nq * hb::TC_KP);
  bound.reserve((size_t)splits * hb::TC_HALVES * nq);
  hb::TcParams p{};
  p.x_hi = x_hi.p; p.x_lo = x_lo.p; p.q_hi = q_hi.p; p.q_lo = q_lo.p; p.x_norm = x_norm.p;
  p.n = n; p.nq = nq; p.kp = kp; p.segs = any_lo ? 3 : 1; p.split_len = split_len; p.partial = partial.p; p.bound = bound.p; p.gthr = gthr.p;
  tm.lap("(sync, alloc)");
  if (tc2) {
    size_t smem = hb::tc2_smem_bytes();
    CUDA_CHECK(cudaFuncSetAttribute(hb::bruteforce_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hb::bruteforce_tc2_kernel<<<dim3(qblocks, splits), hb::T2_THREADS, smem, s>>>(p);
  } else {
    size_t smem = hb::tc_smem_bytes();
    CUDA_CHECK(cudaFuncSetAttribute(hb::bruteforce_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hb::bruteforce_tc_kernel<<<dim3(qblocks, splits), hb::TC_THREADS, smem, s>>>(p);
  }
  CUDA_CHECK(cudaGetLastError());
  tm.lap(p.segs == 3 ? "bruteforce_tc kernel (3 segments)" : "bruteforce_tc kernel (1 segment)");

  hb::TcFinishParams f{};
  f.g.vec = d_data; f.g.ld4 = ld / 4; f.g.chunks = ld / 4; f.g.metric = 0; f.g.n = (int)n;
  f.queries = d_q; f.partial = partial.p; f.bound = bound.p; f.S = splits * hb::TC_HALVES; f.nq = nq; f.k = k; f.k_cap = round_up(k, 32);
  f.q_chunks = round_up(ld / 4, 2);
  f.smem_per_warp = hb::tc_finish_smem_per_warp(f.k_cap, f.q_chunks);
  f.eps = any_lo ? 1.0f / 4096.0f : 1.0f / 65536.0f;
  f.max_norm = scal.p; f.ids = d_ids; f.dists = d_dists; f.flags = flags.p;
  int fw = 8;
  while (fw > 1 && (size_t)fw * f.smem_per_warp > 200 * 1024) fw--;
  size_t fsmem = (size_t)fw * f.smem_per_warp;
  CUDA_CHECK(cudaFuncSetAttribute(hb::bruteforce_tc_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
  hb::bruteforce_tc_finish_kernel<<<(unsigned)((nq + fw - 1) / fw), fw * 32, fsmem, s>>>(f);
  CUDA_CHECK(cudaGetLastError());
  tm.lap("bruteforce_tc_finish_kernel");
  std::vector<int> h_flags((size_t)nq);
  CUDA_CHECK(cudaMemcpyAsync(h_flags.data(), flags.p, (size_t)nq * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (launches) *launches += 4;
  int64_t bad = 0;
  for (int v : h_flags) bad += v != 0;
  return bad;
}
}  // namespace

extern "C" {

int hnswb200_bruteforce_knn(const float* data, int64_t n, const float* queries, int64_t nq, int dim, int k, int metric,
                            int device, int32_t* ids, float* dists) {
  return guard([&] {
    if (!data || !queries || !dists) fail(HNSWB200_EINVAL, "bruteforce_knn: NULL argument");
    if (n <= 0 || nq <= 0 || dim <= 0) fail(HNSWB200_EINVAL, "bruteforce_knn: n, nq, dim must be > 0");
    if (k <= 0 || k > 1024) fail(HNSWB200_EINVAL, "bruteforce_knn: k must be in 1..1024");
    if (metric < 0 || metric > 2) fail(HNSWB200_EINVAL, "bruteforce_knn: unknown metric");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      fail(HNSWB200_ECUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    int ld = round_up(dim, 4);
    DevBuf<float> d_data, d_q, d_d;
    DevBuf<int32_t> d_i;
    d_data.reserve((size_t)n * ld); d_q.reserve((size_t)nq * ld); d_d.reserve((size_t)nq * k); d_i.reserve((size_t)nq * k);
    upload_rows(d_data.p, ld, data, dim, n, 0);
    upload_rows(d_q.p, ld, queries, dim, nq, 0);
    // L2 and k <= 16: tensor cores rank the candidates, fp32 re-ranks and proves exactness; anything it
    // cannot prove (and every other case) goes through the fp32 CUDA-core kernel.
    const char* force = std::getenv("HNSWB200_BRUTEFORCE");
    bool tc = metric == HNSWB200_L2 && k <= hb::TC_KP && n >= hb::TC_N && !(force && std::string(force) == "fp32");
    int64_t unproven = tc ? launch_bruteforce_tc(d_data.p, n, d_q.p, nq, ld, dim, k, prop.multiProcessorCount, d_i.p, d_d.p, 0, nullptr) : -1;
    g_last_bruteforce_unproven = unproven;
    if (unproven != 0)
      launch_bruteforce(d_data.p, n, d_q.p, nq, ld, k, metric, prop.multiProcessorCount, d_i.p, d_d.p, 0, nullptr);
    if (ids) CUDA_CHECK(cudaMemcpy(ids, d_i.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(dists, d_d.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
  });
}

int64_t hnswb200_bruteforce_last_unproven(void) { return g_last_bruteforce_unproven; }

int hnswb200_merge_topk_device(const int32_t* d_ids, const float* d_dists, int n_shards, int64_t nq, int k,
                               int64_t shard_stride, const int64_t* shard_offsets, int32_t* d_out_ids, float* d_out_dists,
                               void* stream) {
  return guard([&] {
    if (!d_ids || !d_dists || !d_out_ids || !d_out_dists) fail(HNSWB200_EINVAL, "merge_topk: NULL argument");
    if (n_shards < 1 || n_shards > 32) fail(HNSWB200_EINVAL, "merge_topk: n_shards must be in 1..32");
    if (nq <= 0 || k <= 0) fail(HNSWB200_EINVAL, "merge_topk: nq and k must be > 0");
    if (shard_stride == 0) shard_stride = nq * k;
    if (shard_stride < nq * k) fail(HNSWB200_EINVAL, "merge_topk: shard_stride smaller than nq * k");
    hb::ShardOffsets offs{};
    for (int s = 0; s < n_shards; s++) {
      int64_t o = shard_offsets ? shard_offsets[s] : 0;
      if (o < 0 || o >= (int64_t(1) << 31)) fail(HNSWB200_EINVAL, "merge_topk: shard offset out of range");
      offs.v[s] = (int32_t)o;
    }
    int wpb = 8;
    hb::merge_topk_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        d_ids, d_dists, n_shards, nq, k, shard_stride, offs, d_out_ids, d_out_dists);
    CUDA_CHECK(cudaGetLastError());
    if (!stream) CUDA_CHECK(cudaStreamSynchronize(0));
  });
}

}  // extern "C"
