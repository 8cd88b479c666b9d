// scratch: host->device copy bandwidth, default pinned vs write-combined pinned, one vs two streams, chunk sizes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <vector>
static double run(void* h, void* d, size_t bytes, size_t chunk, int nstreams, cudaStream_t* st) {
  cudaDeviceSynchronize();
  auto t0 = std::chrono::steady_clock::now();
  int k = 0;
  for (size_t off = 0; off < bytes; off += chunk, k++) {
    size_t n = std::min(chunk, bytes - off);
    cudaMemcpyAsync((char*)d + off, (char*)h + off, n, cudaMemcpyHostToDevice, st[k % nstreams]);
  }
  cudaDeviceSynchronize();
  double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return bytes / s / 1e9;
}
int main() {
  const size_t bytes = 800ull << 20;
  void *h0, *h1, *d;
  cudaMalloc(&d, bytes);
  cudaHostAlloc(&h0, bytes, cudaHostAllocDefault);
  cudaHostAlloc(&h1, bytes, cudaHostAllocWriteCombined);
  memset(h0, 1, bytes);
  memset(h1, 1, bytes);
  cudaStream_t st[4];
  for (auto& s : st) cudaStreamCreate(&s);
  for (int rep = 0; rep < 2; rep++)
    for (size_t chunk : {bytes, size_t(64) << 20, size_t(8) << 20, size_t(1) << 20})
      for (int ns : {1, 2, 4}) {
        if (chunk == bytes && ns > 1) continue;
        printf("chunk %4zu MB streams %d : default %.1f GB/s  write-combined %.1f GB/s\n", chunk >> 20, ns, run(h0, d, bytes, chunk, ns, st),
               run(h1, d, bytes, chunk, ns, st));
      }
  return 0;
}
