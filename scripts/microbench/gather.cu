// Random-row gather ceiling: what HBM delivers for independent 512-byte (dim*4) row reads,
// the access pattern of the search kernel's distance evaluations.  Not part of the product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// each team of 8 lanes reads VPT rows per round (float4 per lane x 4 chunks = 512 B per row)
template <int VPT, bool DEP>
__global__ void gather_kernel(const float4* __restrict__ data, uint32_t nrows, int rounds, float* out) {
  const int lane = threadIdx.x & 31, tl = lane & 7;
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) / 8 * 2654435761u + 12345u;   // same per team
  float acc = 0.f;
  for (int r = 0; r < rounds; r++) {
    float4 v[VPT][4];
#pragma unroll
    for (int j = 0; j < VPT; j++) {
      uint32_t row = rng(s) % nrows;
      if (DEP) row = (row + (uint32_t)(int)acc) % nrows;     // next address depends on the data: no overlap across rounds
      const float4* p = data + (size_t)row * 32;
#pragma unroll
      for (int c = 0; c < 4; c++) v[j][c] = __ldg(p + tl + 8 * c);
    }
#pragma unroll
    for (int j = 0; j < VPT; j++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc += v[j][c].x + v[j][c].y + v[j][c].z + v[j][c].w;
    if (DEP) acc = acc * 1e-30f;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int VPT, bool DEP>
void run(const float4* d, uint32_t nrows, int warps_per_sm, int sms, float* out) {
  int threads = 256, ctas = sms * warps_per_sm * 32 / threads;
  int rounds = 400;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_kernel<VPT, DEP><<<ctas, threads>>>(d, nrows, 20, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  gather_kernel<VPT, DEP><<<ctas, threads>>>(d, nrows, rounds, out);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = (double)ctas * threads / 8 * VPT * rounds * 512.0;
  printf("rows/team/round=%d dependent=%d warps/SM=%2d : %7.1f GB/s  (%.3f ms)\n", VPT, (int)DEP, warps_per_sm, bytes / ms / 1e6, ms);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  uint32_t nrows = 1000000;
  float4* d; float* out;
  CK(cudaMalloc(&d, (size_t)nrows * 512)); CK(cudaMemset(d, 0, (size_t)nrows * 512)); CK(cudaMalloc(&out, 4));
  printf("%s, %d SMs, table %u rows x 512 B\n", prop.name, prop.multiProcessorCount, nrows);
  for (int w : {8, 16, 24, 32, 48, 64}) run<2, true>(d, nrows, w, prop.multiProcessorCount, out);
  for (int w : {8, 16, 32, 64}) run<4, true>(d, nrows, w, prop.multiProcessorCount, out);
  for (int w : {8, 16, 32, 64}) run<2, false>(d, nrows, w, prop.multiProcessorCount, out);
  for (int w : {8, 16, 32}) run<8, false>(d, nrows, w, prop.multiProcessorCount, out);
  return 0;
}
