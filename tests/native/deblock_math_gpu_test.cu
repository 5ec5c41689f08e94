// GPU check of the packed edge filters: the device build of dryv_b200/csrc/deblock_packed.cuh (real VIADDMNMX / VIMNMX / PRMT
// instructions) against the host build of the same functions (tests/native/deblock_math_test.cpp pins that one to 8.7.2).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../dryv_b200/csrc/deblock_packed.cuh"

static const int kAlpha[52] = DRYV_DB_ALPHA;
static const int kBeta[52] = DRYV_DB_BETA;
static const int kTc0[52] = DRYV_DB_TC0_BS3;

struct Case {
  uint32_t r[8];
  dryv::EdgeConst k;
  uint32_t off, chroma;
  int strong;
};

__host__ __device__ inline void run_case(Case& c) {
  if (c.strong) dryv::filter_edge_strong(c.r[0], c.r[1], c.r[2], c.r[3], c.r[4], c.r[5], c.r[6], c.r[7], c.k, c.off, c.chroma);
  else dryv::filter_edge_normal(c.r[1], c.r[2], c.r[3], c.r[4], c.r[5], c.r[6], c.k, c.off, c.chroma);
}

__global__ void run_cases(Case* c, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) run_case(c[i]);
}

static unsigned rng_state = 777u;
static unsigned rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

int main() {
  const int n = 1 << 20;
  std::vector<Case> h(n), want(n);
  for (int i = 0; i < n; i++) {
    Case& c = h[i];
    const int ia = rnd() % 52, ib = rnd() % 52;
    const bool chroma = rnd() & 1;
    c.k = dryv::make_edge_const(kAlpha[ia], kBeta[ib], kTc0[ia], chroma);
    c.chroma = chroma ? 0xffffffffu : 0u;
    c.off = (rnd() % 8 == 0) ? 0xffff0000u : 0u;
    c.strong = rnd() & 1;
    int base[2] = {(int)(rnd() % 256), (int)(rnd() % 256)};
    const int noise = (rnd() % 3 == 0) ? 40 : 4;
    for (int j = 0; j < 8; j++) {
      uint32_t v = 0;
      for (int hh = 0; hh < 2; hh++) {
        int s = base[hh] + (j >= 4 ? (int)(rnd() % 21) - 10 : 0) + (int)(rnd() % (2 * noise + 1)) - noise;
        s = s < 0 ? 0 : (s > 255 ? 255 : s);
        v |= (uint32_t)s << (16 * hh);
      }
      c.r[j] = v;
    }
    want[i] = c;
    run_case(want[i]);
  }
  Case* d = nullptr;
  if (cudaMalloc(&d, n * sizeof(Case)) != cudaSuccess) return 2;
  cudaMemcpy(d, h.data(), n * sizeof(Case), cudaMemcpyHostToDevice);
  run_cases<<<(n + 255) / 256, 256>>>(d, n);
  if (cudaDeviceSynchronize() != cudaSuccess) return 3;
  std::vector<Case> got(n);
  cudaMemcpy(got.data(), d, n * sizeof(Case), cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < 8; j++)
      if (got[i].r[j] != want[i].r[j]) {
        printf("MISMATCH case %d sample %d: device %08x host %08x (strong %d chroma %08x off %08x)\n  in:", i, j, got[i].r[j],
               want[i].r[j], h[i].strong, h[i].chroma, h[i].off);
        for (int q = 0; q < 8; q++) printf(" %08x", h[i].r[q]);
        printf("\n  k: ka %08x kb %08x ks %08x tcb %08x lo1 %08x\n", h[i].k.ka, h[i].k.kb, h[i].k.ks, h[i].k.tcb, h[i].k.lo1);
        return 1;
      }
  printf("ok: %d packed edge filters, device == host\n", n);
  return 0;
}
