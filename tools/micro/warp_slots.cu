// Probe: which hardware warp slot (%warpid) and SM each warp of a 64-thread CTA gets when 10 such CTAs share an SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o warp_slots warp_slots.cu ; run on a B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64, 10) probe(unsigned* out) {
  __shared__ char pad[18000];
  pad[threadIdx.x] = 0;
  unsigned w, sm;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2] = sm;
    out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2 + 1] = w;
  }
  // stay resident so that all CTAs coexist
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {
  }
  if (pad[threadIdx.x] == 1) out[0] = 0;
}
int main() {
  const int grid = 1480;
  unsigned* d;
  cudaMalloc(&d, grid * 4 * sizeof(unsigned));
  cudaFuncSetAttribute(probe, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  probe<<<grid, 64>>>(d);
  unsigned* h = new unsigned[grid * 4];
  cudaMemcpy(h, d, grid * 4 * sizeof(unsigned), cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int sm = 0; sm < 2; sm++) {
    printf("SM %d:", sm);
    for (int b = 0; b < grid; b++)
      if (h[b * 4] == (unsigned)sm) printf(" cta%d:(%u,%u)", b, h[b * 4 + 1], h[b * 4 + 3]);
    printf("\n");
  }
  int hist[4][2] = {};
  for (int b = 0; b < grid; b++) {
    hist[h[b * 4 + 1] & 3][0]++;
    hist[h[b * 4 + 3] & 3][1]++;
  }
  for (int q = 0; q < 4; q++) printf("slot%%4 == %d: %d first warps, %d second warps\n", q, hist[q][0], hist[q][1]);
  return 0;
}
