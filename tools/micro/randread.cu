// randread.cu -- what does a random small read cost on this GPU as the footprint grows?  (cfg-5 probe design input:
// Bloom words and table slots are random 8/16-byte reads over 0.1 - 4 GB.)
//   throughput : every thread issues independent loads at hashed addresses (as the Bloom stage does)
//   latency    : one dependent chain per thread, low occupancy (as a slot -> posting -> key chain does)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/randread tools/micro/randread.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
template <int BYTES>
__global__ void throughput(const uint64_t* __restrict__ buf, uint64_t words_mask, int iters, uint64_t* sink) {
  uint64_t acc = 0, s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 1;
  for (int i = 0; i < iters; ++i) {
    s = mix(s + i);
    const uint64_t w = (s & words_mask) & ~(uint64_t)(BYTES / 8 - 1);
    if (BYTES == 8) acc += __ldg(buf + w);
    else if (BYTES == 16) { const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(buf + w)); acc += v.x ^ v.y; }
    else { const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(buf + w); acc += v.x ^ v.y ^ v.z ^ v.w; }
  }
  if (acc == 0x1234567) *sink = acc;
}
__global__ void latency(const uint64_t* __restrict__ buf, uint64_t words_mask, int iters, uint64_t* sink) {
  uint64_t s = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 1;
  for (int i = 0; i < iters; ++i) s = mix(s + __ldg(buf + (s & words_mask)));
  if (s == 0x1234567) *sink = s;
}
int main(int argc, char** argv) {
  cudaSetDevice(0);
  if (argc > 1) {  // L2 fetch granularity hint in bytes (32, 64 or 128; the default is the driver's)
    const cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
    size_t got = 0;
    cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity requested %s -> %s, now %zu\n", argv[1], cudaGetErrorString(e), got);
  }
  const size_t maxbytes = (size_t)16 << 30;
  uint64_t* buf; uint64_t* sink;
  if (cudaMalloc(&buf, maxbytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMalloc(&sink, 8);
  cudaMemset(buf, 1, maxbytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("%10s %14s %14s %14s %14s\n", "MB", "G 8B-reads/s", "G 16B-reads/s", "G 32B-reads/s", "chain ns");
  for (size_t mb = 32; mb <= 16384; mb *= 2) {
    const uint64_t mask = (mb << 20) / 8 - 1;
    float r[4];
    for (int v = 0; v < 4; ++v) {
      const int blocks = v == 3 ? 148 : 148 * 8, threads = v == 3 ? 64 : 256, iters = v == 3 ? 2000 : 400;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (v == 0) throughput<8><<<blocks, threads>>>(buf, mask, iters, sink);
        if (v == 1) throughput<16><<<blocks, threads>>>(buf, mask, iters, sink);
        if (v == 2) throughput<32><<<blocks, threads>>>(buf, mask, iters, sink);
        if (v == 3) latency<<<blocks, threads>>>(buf, mask, iters, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double n = (double)blocks * threads * iters;
      r[v] = v == 3 ? (float)(ms * 1e6 / iters) : (float)(n / (ms * 1e-3) / 1e9);
    }
    printf("%10zu %14.2f %14.2f %14.2f %14.1f\n", mb, r[0], r[1], r[2], r[3]);
  }
  return 0;
}
