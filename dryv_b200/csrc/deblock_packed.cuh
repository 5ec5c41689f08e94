// deblock_packed.cuh — the H.264 8.7.2.3 / 8.7.2.4 edge filters on TWO lines of samples per register (16-bit fields), without
// a branch. The deblocking kernel (deblock_kernel.cuh) is bound by instruction issue like the reconstruction kernels, and the
// edges of one macroblock are filtered one after the other by definition, so what can be saved is instructions per edge:
// every lane filters two lines at once (luma and chroma lanes run the same instruction stream, the differences are masks),
// thresholds are pre-biased constants, decisions are sign bits spread to field masks by one PRMT.
//
// A packed sample register holds sample c of line A in bits 0..7 and of line B in bits 16..23; every field of every such
// register is always in [0, 255], and every intermediate below keeps its fields inside [0, 65536) so that plain 32-bit
// add / subtract / multiply are exact per field (no carry or borrow crosses bit 16). Signed quantities travel biased.
// Host-compilable: tests/native/deblock_math_test.cpp checks these functions against a scalar statement of 8.7.2.
#pragma once
#include <stdint.h>

#include "residual_stage.cuh"  // DRYV_HD, prmt, viaddmin_relu_s16x2, pk2

// Tables 8-16 (alpha', beta') and 8-17 (tC0' for bS 3) by indexA / indexB, and 8-15 (QPc by qPI)
#define DRYV_DB_ALPHA                                                                                                     \
  {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40,   \
   45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255}
#define DRYV_DB_BETA                                                                                                      \
  {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11,  \
   12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18}
#define DRYV_DB_TC0_BS3                                                                                                   \
  {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, \
   7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25}
#define DRYV_DB_QPC                                                                                                       \
  {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30,  \
   31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39}

namespace dryv {

// per 16-bit field, wraparound
DRYV_HD uint32_t vadd2(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vadd2(a, b);
#else
  return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
DRYV_HD uint32_t vmin2s(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vmins2(a, b);
#else
  const int16_t l = (int16_t)a < (int16_t)b ? (int16_t)a : (int16_t)b;
  const int16_t h = (int16_t)(a >> 16) < (int16_t)(b >> 16) ? (int16_t)(a >> 16) : (int16_t)(b >> 16);
  return (uint16_t)l | ((uint32_t)(uint16_t)h << 16);
#endif
}
DRYV_HD uint32_t vmax2s(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __vmaxs2(a, b);
#else
  const int16_t l = (int16_t)a > (int16_t)b ? (int16_t)a : (int16_t)b;
  const int16_t h = (int16_t)(a >> 16) > (int16_t)(b >> 16) ? (int16_t)(a >> 16) : (int16_t)(b >> 16);
  return (uint16_t)l | ((uint32_t)(uint16_t)h << 16);
#endif
}
// 0xffff in every field whose bit 15 is set (one PRMT: the sign-replicating byte selectors)
DRYV_HD uint32_t signmask2(uint32_t x) {
#ifdef __CUDA_ARCH__
  uint32_t m;  // not __byte_perm: that one drops bit 3 of the selectors, which is what asks for the sign
  asm("prmt.b32 %0, %1, 0, 0xbb99;" : "=r"(m) : "r"(x));
  return m;
#else
  return ((x & 0x8000u) ? 0xffffu : 0u) | ((x & 0x80000000u) ? 0xffff0000u : 0u);
#endif
}
// a where the mask is set, b elsewhere
DRYV_HD uint32_t sel2(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

// bit 15 of a field is set iff |a - b| >= T there, with k = 0x8000 - T per field (a, b in [0, 255], T in [0, 255 + 66])
DRYV_HD uint32_t fail_abs(uint32_t a, uint32_t b, uint32_t k) { return ((a + k) - b) | ((b + k) - a); }

// Thresholds of one edge kind (left / top macroblock edge, inner edges) of one lane, pre-biased.
struct EdgeConst {
  uint32_t ka;   // 0x8000 - alpha
  uint32_t kb;   // 0x8000 - beta
  uint32_t ks;   // 0x8000 - ((alpha >> 2) + 2)                          (bS 4)
  uint32_t tcb;  // tc0 + 2; chroma: tc0 + 3                              (bS < 4)
  uint32_t lo1;  // 512 - tc0: lower clip of the biased p1 / q1 delta; the upper one is 1024 - lo1
};
// index in [0, 51] -> the constants (both fields alike); alpha, beta, tc0 (bS 3): Tables 8-16 / 8-17
DRYV_HD EdgeConst make_edge_const(int alpha, int beta, int tc0, bool chroma) {
  EdgeConst k;
  k.ka = pk2(0x8000u - (uint32_t)alpha);
  k.kb = pk2(0x8000u - (uint32_t)beta);
  k.ks = pk2(0x8000u - (uint32_t)((alpha >> 2) + 2));
  k.tcb = pk2((uint32_t)tc0 + (chroma ? 3u : 2u));
  k.lo1 = pk2(512u - (uint32_t)tc0);
  return k;
}

// bS < 4 (8.7.2.3). `off`: 0xffff per field that must stay untouched (edge not filtered on this lane); `chroma`: 0xffff per
// chroma field (p1 / q1 are never changed, tc = tc0 + 1). p3 / q3 are not used.
DRYV_HD void filter_edge_normal(uint32_t p2, uint32_t& p1, uint32_t& p0, uint32_t& q0, uint32_t& q1, uint32_t q2,
                                const EdgeConst& k, uint32_t off, uint32_t chroma) {
  const uint32_t keep = signmask2(fail_abs(p0, q0, k.ka) | fail_abs(p1, p0, k.kb) | fail_abs(q1, q0, k.kb)) | off;
  const uint32_t nap = signmask2(fail_abs(p2, p0, k.kb)) | chroma;  // 0xffff: a_p >= beta (or chroma)
  const uint32_t naq = signmask2(fail_abs(q2, q0, k.kb)) | chroma;
  const uint32_t tc = vadd2(vadd2(k.tcb, nap), naq);               // tc0 + (a_p < beta) + (a_q < beta); chroma tc0 + 1
  // delta + 256 = (((q0 - p0) << 2) + (p1 - q1) + 4 + 2048) >> 3, clipped to 256 +- tc
  const uint32_t x = (q0 * 4u + p1 + pk2(2048u + 4u)) - (p0 * 4u + q1);
  uint32_t dd = (x >> 3) & 0x1fff1fffu;
  dd = vmax2s(vmin2s(dd, pk2(256u) + tc), pk2(256u) - tc);
  const uint32_t np0 = viaddmin_relu_s16x2(p0 + dd, pk2(0x10000u - 256u), pk2(255u));
  const uint32_t nq0 = viaddmin_relu_s16x2(q0 + pk2(512u) - dd, pk2(0x10000u - 256u), pk2(255u));
  // p1 + Clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1), the inner term biased by 1024 before the shift
  const uint32_t avg = ((p0 + q0 + pk2(1u)) >> 1) & 0x01ff01ffu;
  const uint32_t hi1 = pk2(1024u) - k.lo1;
  uint32_t yp = ((p2 + avg + pk2(1024u) - p1 * 2u) >> 1) & 0x03ff03ffu;
  yp = vmax2s(vmin2s(yp, hi1), k.lo1);
  uint32_t yq = ((q2 + avg + pk2(1024u) - q1 * 2u) >> 1) & 0x03ff03ffu;
  yq = vmax2s(vmin2s(yq, hi1), k.lo1);
  const uint32_t np1 = p1 + yp - pk2(512u);
  const uint32_t nq1 = q1 + yq - pk2(512u);
  p1 = sel2(keep | nap, p1, np1);
  q1 = sel2(keep | naq, q1, nq1);
  p0 = sel2(keep, p0, np0);
  q0 = sel2(keep, q0, nq0);
}

// bS 4 (8.7.2.4)
DRYV_HD void filter_edge_strong(uint32_t p3, uint32_t& p2, uint32_t& p1, uint32_t& p0, uint32_t& q0, uint32_t& q1, uint32_t& q2,
                                uint32_t q3, const EdgeConst& k, uint32_t off, uint32_t chroma) {
  const uint32_t keep = signmask2(fail_abs(p0, q0, k.ka) | fail_abs(p1, p0, k.kb) | fail_abs(q1, q0, k.kb)) | off;
  const uint32_t big = fail_abs(p0, q0, k.ks);                            // |p0 - q0| >= (alpha >> 2) + 2
  const uint32_t wp = signmask2(fail_abs(p2, p0, k.kb) | big) | chroma;   // 0xffff: the weak (one-sample) form on the p side
  const uint32_t wq = signmask2(fail_abs(q2, q0, k.kb) | big) | chroma;
  const uint32_t t = p0 + q0, u = t + p1, v = t + q1;
  const uint32_t p0s = ((u * 2u + p2 + q1 + pk2(4u)) >> 3) & 0x00ff00ffu;
  const uint32_t p1s = ((p2 + u + pk2(2u)) >> 2) & 0x00ff00ffu;
  const uint32_t p2s = ((p3 * 2u + p2 * 3u + u + pk2(4u)) >> 3) & 0x00ff00ffu;
  const uint32_t p0w = ((p1 * 2u + p0 + q1 + pk2(2u)) >> 2) & 0x00ff00ffu;
  const uint32_t q0s = ((v * 2u + q2 + p1 + pk2(4u)) >> 3) & 0x00ff00ffu;
  const uint32_t q1s = ((q2 + v + pk2(2u)) >> 2) & 0x00ff00ffu;
  const uint32_t q2s = ((q3 * 2u + q2 * 3u + v + pk2(4u)) >> 3) & 0x00ff00ffu;
  const uint32_t q0w = ((q1 * 2u + q0 + p1 + pk2(2u)) >> 2) & 0x00ff00ffu;
  p0 = sel2(keep, p0, sel2(wp, p0w, p0s));
  p1 = sel2(keep | wp, p1, p1s);
  p2 = sel2(keep | wp, p2, p2s);
  q0 = sel2(keep, q0, sel2(wq, q0w, q0s));
  q1 = sel2(keep | wq, q1, q1s);
  q2 = sel2(keep | wq, q2, q2s);
}

}  // namespace dryv
