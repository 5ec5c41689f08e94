// int_issue.cu — issue-rate micro-benchmark for the integer formulation of the reconstruction path (SURVEY §7
// "Integer-ALU budget", §8(d) "secondary bounds"): how many warp-instructions per clock one SM sustains for the
// instruction classes the kernels are made of, alone and mixed. The reconstruction kernels are issue bound, so
// these numbers (not the HBM peak) are the roof their instruction counts have to be read against.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_issue int_issue.cu ; run on a B200
// output: one line per instruction class: warp-instructions / clock / SM (4 schedulers: 4.0 = one per scheduler per clock)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;  // independent dependency chains per thread: latency 4-5 cycles is covered by 8 chains x warps

#define DEFINE_KERNEL(name, BODY)                                                          \
  __global__ void __launch_bounds__(256) name(uint32_t* out, uint32_t seed, long long* cyc) { \
    uint32_t r[kChains];                                                                   \
    const uint32_t k1 = seed * 3u + 1u, k2 = seed ^ 0x00ff00ffu;                             \
    _Pragma("unroll") for (int i = 0; i < kChains; i++) r[i] = seed + threadIdx.x * 977u + i * 131u; \
    __syncthreads();                                                                       \
    const long long t0 = clock64();                                                        \
    _Pragma("unroll 1") for (int it = 0; it < kIters; it++) {                              \
      _Pragma("unroll") for (int u = 0; u < 4; u++) {                                       \
        _Pragma("unroll") for (int i = 0; i < kChains; i++) { BODY }                       \
      }                                                                                    \
    }                                                                                      \
    const long long t1 = clock64();                                                        \
    uint32_t acc = 0;                                                                      \
    _Pragma("unroll") for (int i = 0; i < kChains; i++) acc ^= r[i];                       \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                      \
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                       \
  }

// one instruction per BODY unless noted (ops_per_body below)
DEFINE_KERNEL(k_iadd3, asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(k1));)
DEFINE_KERNEL(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(k1), "r"(k2));)
DEFINE_KERNEL(k_shf, asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(k1));)
DEFINE_KERNEL(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(r[i]) : "r"(k1));)
DEFINE_KERNEL(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));)
DEFINE_KERNEL(k_viadd16x2, asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k1));)
DEFINE_KERNEL(k_viaddmnmx, r[i] = __viaddmin_s16x2_relu(r[i], k1, 0x00ff00ffu);)
DEFINE_KERNEL(k_vimnmx3, r[i] = __vimax3_s16x2_relu(r[i], k1, k2);)
DEFINE_KERNEL(k_vimnmx_s32, r[i] = (uint32_t)__vimin_s32_relu((int)r[i], (int)k2);)
DEFINE_KERNEL(k_dp2a, asm volatile("dp2a.lo.s32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));)
DEFINE_KERNEL(k_dp4a, asm volatile("dp4a.u32.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));)
// mixes: two instructions per BODY
DEFINE_KERNEL(k_mix_iadd_imad, if (i & 1) { asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(k1)); } else {
  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));
})
DEFINE_KERNEL(k_mix_lop_dp2a, if (i & 1) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(k1), "r"(k2)); } else {
  asm volatile("dp2a.lo.s32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));
})
DEFINE_KERNEL(k_mix_viadd_imad, if (i & 1) { asm volatile("add.s16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(k1)); } else {
  asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(k1), "r"(k2));
})
DEFINE_KERNEL(k_mix_iadd_ffma, if (i & 1) { asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(k1)); } else {
  float f = __uint_as_float(r[i]);
  asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0000001f), "f"(0.5f));
  r[i] = __float_as_uint(f);
})
DEFINE_KERNEL(k_ffma, float f = __uint_as_float(r[i]);
              asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0000001f), "f"(0.5f));
              r[i] = __float_as_uint(f);)

// shared-memory loads: 32-bit conflict-free, one per BODY, address chained through the loaded value
__global__ void __launch_bounds__(256) k_lds(uint32_t* out, uint32_t seed, long long* cyc) {
  __shared__ uint32_t sm[256 * kChains];
  for (int i = 0; i < kChains; i++) sm[i * 256 + threadIdx.x] = (uint32_t)((i * 256 + threadIdx.x) * 4);  // holds its own byte offset
  __syncthreads();
  uint32_t r[kChains];
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
  for (int i = 0; i < kChains; i++) r[i] = base + (uint32_t)((i * 256 + threadIdx.x) * 4);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < kChains; i++) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(r[i]));
        r[i] = base + v;
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = seed;
#pragma unroll
  for (int i = 0; i < kChains; i++) acc ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(uint32_t*, uint32_t, long long*);

static void run(const char* name, kern_t k, int ops_per_body, int ctas_per_sm, int sms) {
  const int grid = sms * ctas_per_sm;
  uint32_t* d_out;
  long long* d_cyc;
  cudaMalloc(&d_out, (size_t)grid * 256 * 4);
  cudaMalloc(&d_cyc, (size_t)grid * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<<<grid, 256>>>(d_out, 12345u, d_cyc);  // warm-up
  cudaEventRecord(e0);
  k<<<grid, 256>>>(d_out, 12345u, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long* h = new long long[grid];
  cudaMemcpy(h, d_cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; i++) mean += (double)h[i];
  mean /= grid;
  // per SM: ctas_per_sm CTAs x 8 warps, each warp issues kIters * 4 * kChains * ops instructions in `mean` cycles
  const double warp_instr = (double)ctas_per_sm * 8.0 * kIters * 4.0 * kChains * ops_per_body;
  printf("%-18s ctas/SM %d  %.3f warp-instr/clk/SM (SM clocks)  %.3f ms  => %.2f T warp-instr/s chip, %.0f MHz effective\n", name,
         ctas_per_sm, warp_instr / mean, ms, warp_instr * sms / (ms * 1e-3) / 1e12, mean / (ms * 1e-3) / 1e6);
  delete[] h;
  cudaFree(d_out);
  cudaFree(d_cyc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("# %s, %d SMs, int_issue micro-benchmark: %d chains/thread, 8 warps/CTA\n", p.name, p.multiProcessorCount, kChains);
  const int sms = p.multiProcessorCount;
  for (int c = 1; c <= 2; c++) {
    run("IADD", k_iadd3, 1, c, sms);
    run("LOP3", k_lop3, 1, c, sms);
    run("SHF", k_shf, 1, c, sms);
    run("PRMT", k_prmt, 1, c, sms);
    run("IMAD", k_imad, 1, c, sms);
    run("VIADD.16x2", k_viadd16x2, 1, c, sms);
    run("VIADDMNMX.S16x2", k_viaddmnmx, 1, c, sms);
    run("VIMNMX3.S16x2", k_vimnmx3, 1, c, sms);
    run("VIMNMX.S32.RELU", k_vimnmx_s32, 1, c, sms);
    run("IDP.2A", k_dp2a, 1, c, sms);
    run("IDP.4A", k_dp4a, 1, c, sms);
    run("FFMA", k_ffma, 1, c, sms);
    run("LDS.32", k_lds, 1, c, sms);
    run("mix IADD+IMAD", k_mix_iadd_imad, 1, c, sms);
    run("mix LOP3+IDP.2A", k_mix_lop_dp2a, 1, c, sms);
    run("mix VIADD+IMAD", k_mix_viadd_imad, 1, c, sms);
    run("mix IADD+FFMA", k_mix_iadd_ffma, 1, c, sms);
  }
  return 0;
}
