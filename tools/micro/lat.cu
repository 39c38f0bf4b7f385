// Micro-benchmark: dependent random 32-byte gathers over a footprint of F bytes, T threads in flight.
// Prints cycles per dependent gather (mean over the chain) - how footprint (TLB reach) and load change the latency
// that a level-synchronous step pays per round trip.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a lat.cu -o lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void chase(const uint4 *buf, uint64_t mask, int steps, unsigned long long *cyc, int atomic_mode, unsigned *abuf) {
  uint64_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL;
  unsigned long long t0 = clock64();
  uint32_t acc = 0;
  for (int i = 0; i < steps; i++) {
    x ^= x >> 31; x *= 0x7fb5d329728ea185ULL; x ^= x >> 27;
    uint64_t idx = (x + acc) & mask;
    if (atomic_mode) acc = atomicMax(abuf + idx * 8, (unsigned)i + 1) & 1;
    else { uint4 v = __ldcg(buf + idx * 2); acc = v.x & 1; }
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / steps + (acc > 5);
}
int main() {
  size_t maxF = 16ULL << 30;
  uint4 *buf; cudaMalloc(&buf, maxF); cudaMemset(buf, 0, maxF);
  unsigned long long *cyc; cudaMallocManaged(&cyc, 8 * 4096);
  for (int am = 0; am < 2; am++)
  for (size_t F : {32ULL << 20, 256ULL << 20, 1ULL << 30, 4ULL << 30, 16ULL << 30}) {
    for (int blocks : {1, 16, 58, 148, 592}) {
      for (int threads : {32, 1024}) {
        uint64_t mask = F / 32 - 1;
        chase<<<blocks, threads>>>(buf, mask, 200, cyc, am, (unsigned *)buf);
        cudaDeviceSynchronize();
        chase<<<blocks, threads>>>(buf, mask, 400, cyc, am, (unsigned *)buf);
        cudaDeviceSynchronize();
        unsigned long long s = 0, mx = 0; for (int b = 0; b < blocks; b++) { s += cyc[b]; if (cyc[b] > mx) mx = cyc[b]; }
        printf("%s F=%6zu MB blocks=%4d threads=%5d  cycles/gather mean %6llu max %6llu\n", am ? "atomicMax" : "ldcg32B  ", F >> 20, blocks, threads, s / blocks, mx);
      }
    }
  }
  return 0;
}
