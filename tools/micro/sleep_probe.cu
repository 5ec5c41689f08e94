// development probe: how long does __nanosleep(N) really suspend a warp on this GPU, and how long does a
// relaxed gpu-scope load of a line another SM keeps rewriting take?  nvcc -arch=sm_100a -o sleep_probe sleep_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe_sleep(unsigned ns, long long* out, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) __nanosleep(ns);
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
__global__ void probe_pingpong(volatile unsigned long long* flag, long long* out, int iters) {
  // block 0 and block 1 bounce a counter through L2: round trip = 2 x (store -> visible to a polling load)
  unsigned long long me = blockIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    unsigned long long want = 2ull * i + me;
    unsigned long long v;
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    } while (v != want);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(want + 1) : "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
int main() {
  long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
  long long h[1024];
  unsigned vals[] = {0, 20, 50, 100, 200, 400, 800, 1000, 2000, 4000, 20000};
  for (unsigned ns : vals) {
    for (int blocks : {1, 148 * 12}) {
      probe_sleep<<<blocks, 64>>>(ns, d, 200);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, sizeof(long long) * (blocks > 1024 ? 1024 : blocks), cudaMemcpyDeviceToHost);
      long long mn = h[0], mx = h[0]; double s = 0; int n = blocks > 1024 ? 1024 : blocks;
      for (int i = 0; i < n; i++) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; s += h[i]; }
      printf("nanosleep(%5u) blocks %4d: cycles/iter min %lld avg %.0f max %lld\n", ns, blocks, mn, s / n, mx);
    }
  }
  unsigned long long* f; cudaMalloc(&f, 8); cudaMemset(f, 0, 8);
  probe_pingpong<<<2, 1>>>(f, d, 2000);
  cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("L2 ping-pong one-way (store -> polled load sees it): %lld cycles\n", h[0] / 2);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
