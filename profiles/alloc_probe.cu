// alloc_probe.cu -- how long does this box take to hand out 8 GB of device memory?  cudaMalloc vs the stream-ordered
// pool (cudaMallocAsync), first and second time, in either order.  nvcc -O2 -o alloc_probe alloc_probe.cu
//   ./alloc_probe malloc-first | pool-first
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void t_malloc(size_t bytes, const char *tag) {
  void *p = nullptr;
  double t0 = now();
  cudaMalloc(&p, bytes);
  double t1 = now();
  cudaMemset(p, 0, bytes);
  cudaDeviceSynchronize();
  double t2 = now();
  cudaFree(p);
  double t3 = now();
  printf("{\"what\": \"cudaMalloc %s\", \"gb\": %.1f, \"alloc_s\": %.4f, \"touch_s\": %.4f, \"free_s\": %.4f}\n", tag, bytes / 1e9, t1 - t0, t2 - t1, t3 - t2);
}
static void t_pool(size_t bytes, const char *tag, cudaStream_t s) {
  void *p = nullptr;
  double t0 = now();
  cudaMallocAsync(&p, bytes, s);
  cudaStreamSynchronize(s);
  double t1 = now();
  cudaMemsetAsync(p, 0, bytes, s);
  cudaStreamSynchronize(s);
  double t2 = now();
  cudaFreeAsync(p, s);
  cudaStreamSynchronize(s);
  double t3 = now();
  printf("{\"what\": \"cudaMallocAsync %s\", \"gb\": %.1f, \"alloc_s\": %.4f, \"touch_s\": %.4f, \"free_s\": %.4f}\n", tag, bytes / 1e9, t1 - t0, t2 - t1, t3 - t2);
}
int main(int argc, char **argv) {
  const size_t bytes = (size_t)8 << 30;
  cudaFree(0);
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaMemPool_t pool;
  cudaDeviceGetDefaultMemPool(&pool, 0);
  unsigned long long keep = ~0ull;
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  const bool pool_first = argc > 1 && strcmp(argv[1], "pool-first") == 0;
  for (int rep = 0; rep < 2; rep++) {
    if (pool_first) {
      t_pool(bytes, rep ? "second" : "first", s);
      t_malloc(bytes, rep ? "second" : "first");
    } else {
      t_malloc(bytes, rep ? "second" : "first");
      t_pool(bytes, rep ? "second" : "first", s);
    }
  }
  return 0;
}
