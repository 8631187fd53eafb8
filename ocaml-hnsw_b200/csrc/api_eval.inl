// C ABI: brute-force ground truth and shard merge (included by hnsw_b200.cu).
namespace {
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
  hb::bruteforce_kernel<<<dim3(qblocks, splits), hb::BF_THREADS, smem, s>>>(p);
  CUDA_CHECK(cudaGetLastError());
  int wpb = 8;
  hb::bruteforce_finish_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, s>>>(partial.p, splits, nq, k, metric, d_ids, d_dists);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (launches) *launches += 2;
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
    launch_bruteforce(d_data.p, n, d_q.p, nq, ld, k, metric, prop.multiProcessorCount, d_i.p, d_d.p, 0, nullptr);
    if (ids) CUDA_CHECK(cudaMemcpy(ids, d_i.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(dists, d_d.p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost));
  });
}

int hnswb200_merge_topk_device(const int32_t* d_ids, const float* d_dists, int n_shards, int64_t nq, int k,
                               const int64_t* shard_offsets, int32_t* d_out_ids, float* d_out_dists, void* stream) {
  return guard([&] {
    if (!d_ids || !d_dists || !d_out_ids || !d_out_dists) fail(HNSWB200_EINVAL, "merge_topk: NULL argument");
    if (n_shards < 1 || n_shards > 32) fail(HNSWB200_EINVAL, "merge_topk: n_shards must be in 1..32");
    if (nq <= 0 || k <= 0) fail(HNSWB200_EINVAL, "merge_topk: nq and k must be > 0");
    hb::ShardOffsets offs{};
    for (int s = 0; s < n_shards; s++) {
      int64_t o = shard_offsets ? shard_offsets[s] : 0;
      if (o < 0 || o >= (int64_t(1) << 31)) fail(HNSWB200_EINVAL, "merge_topk: shard offset out of range");
      offs.v[s] = (int32_t)o;
    }
    int wpb = 8;
    hb::merge_topk_kernel<<<(unsigned)((nq + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        d_ids, d_dists, n_shards, nq, k, offs, d_out_ids, d_out_dists);
    CUDA_CHECK(cudaGetLastError());
    if (!stream) CUDA_CHECK(cudaStreamSynchronize(0));
  });
}

}  // extern "C"
