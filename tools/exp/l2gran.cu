// experiment: random 4-byte probes over a 1 GiB array under different L2 fetch granularities
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const uint32_t* a, uint64_t mask, uint32_t* out, int iters) {
  uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9e3779b97f4a7c15ull + 12345;
  uint32_t acc = 0;
  for (int i = 0; i < iters; i += 4) {
    uint64_t h[4];
    for (int j = 0; j < 4; j++) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 29; h[j] = x & mask; x += 0x632be59bd9b4e019ull; }
    for (int j = 0; j < 4; j++) acc += __ldg(a + h[j]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  const uint64_t words = 1ull << 28;  // 1 GiB
  uint32_t *a, *out;
  cudaMalloc(&a, words * 4); cudaMemset(a, 1, words * 4);
  const int grid = 148 * 16, block = 256, iters = 256;
  cudaMalloc(&out, grid * block * 4);
  for (size_t gran : {128, 64, 32}) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
    size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    probe<<<grid, block>>>(a, words - 1, out, iters);
    cudaEventRecord(t0);
    for (int r = 0; r < 5; r++) probe<<<grid, block>>>(a, words - 1, out, iters);
    cudaEventRecord(t1); cudaEventSynchronize(t1);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    double probes = 5.0 * grid * block * iters;
    printf("set %zu (%s) -> limit %zu: %.2f G probes/s, %.1f GB/s at 32B/probe\n", gran, cudaGetErrorString(e), got, probes / ms / 1e6, probes * 32 / ms / 1e6);
  }
  return 0;
}
