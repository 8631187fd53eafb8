// C ABI: index construction (included by hnsw_b200.cu).
extern "C" {

int hnswb200_build(hnswb200_index* x, const float* data, int64_t n, const int32_t* levels) {
  return guard([&] { fail(HNSWB200_EINVAL, "build: not implemented yet"); });
}
int hnswb200_insert(hnswb200_index* x, const float* data, int64_t n, const int32_t* levels) {
  return guard([&] { fail(HNSWB200_EINVAL, "insert: not implemented yet"); });
}

}  // extern "C"
