// residual_stage.cuh — levels -> residual samples (dequantisation, DC Hadamards, 4x4 / 8x8 inverse transforms) for a GROUP of
// up to four consecutive macroblocks per warp pass, and the packed "prediction + residual, clip" step of the consumers.
//
// Reference behaviour reproduced (src/video/frame/): transform.rs:116-191 (4x4 scaling + transform), pred8x8.rs:51-150 (8x8),
// pred16x16.rs:428-482 (Intra16x16 luma DC), trans_chroma.rs:369-456 (chroma DC), transform.rs:194-216 (get_qpc).
//
// Why it looks the way it does: the kernels are bound by instruction issue, not by HBM (profiles/r02_int_issue.txt:
// one warp instruction per scheduler per clock only when ALU-pipe and FMA-pipe instructions alternate), so the stage is written
// to execute as few warp instructions per macroblock as the arithmetic allows and to keep all 32 lanes busy:
//   * one 4x4 block per lane, and the blocks of several macroblocks share a pass: the luma blocks of two macroblocks
//     (16 + 16 lanes), the chroma blocks of four (4 x 8 lanes); an Intra8x8 macroblock's luma is one pass of its own
//     (4 blocks x 8 lanes);
//   * dequantisation = one dp2a per level where the LevelScale factor fits a byte (DeviceTables::t4b): the instruction
//     extracts the int16 level from its packed word, multiplies and adds the rounding / bias constant at once;
//   * the row pass runs on 32-bit scalars, its outputs are packed two columns per register as biased 16-bit fields and the
//     column pass, the final (x + 32) >> 6 and the clip all run on two samples per instruction (IADD3 on packed words is exact
//     as long as no field leaves [0, 65536), which the biases below guarantee for every block whose row-pass outputs lie in
//     [-8192, 8192); a block outside that range — no conforming stream produces one — takes the 32-bit path);
//   * residuals are handed on as biased 16-bit fields r + 512 in [0, 1023], two horizontally adjacent samples per word, so
//     the consumer's "clip(pred + r)" is one VIADDMNMX.S16x2.RELU per two samples.
#pragma once
#include <stdint.h>

#include "recon_tables.h"

#if defined(__CUDACC__)
#define DRYV_HD __host__ __device__ __forceinline__
#else
#define DRYV_HD inline
#endif

namespace dryv {

// ---------------------------------------------------------------------------------------------------------------------
// instruction wrappers (device: one SASS instruction each; host: the same arithmetic, for the CPU unit test of the
// lane-local math, tests/native/block_math_test.cpp)
// ---------------------------------------------------------------------------------------------------------------------
// c + s16(a.lo) * u8(b.byte0) + s16(a.hi) * u8(b.byte1)
DRYV_HD int dp2a_lo_su(uint32_t a, uint32_t b, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return c + (int)(int16_t)(a & 0xffffu) * (int)(b & 0xffu) + (int)(int16_t)(a >> 16) * (int)((b >> 8) & 0xffu);
#endif
}
// c + s16(a.lo) * u8(b.byte2) + s16(a.hi) * u8(b.byte3)
DRYV_HD int dp2a_hi_su(uint32_t a, uint32_t b, int c) {
#ifdef __CUDA_ARCH__
  int d;
  asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return c + (int)(int16_t)(a & 0xffffu) * (int)((b >> 16) & 0xffu) + (int)(int16_t)(a >> 16) * (int)(b >> 24);
#endif
}
DRYV_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  return __byte_perm(a, b, sel);
#else
  const uint64_t v = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
  return r;
#endif
}
// per 16-bit field: max(min(a + b, c), 0), signed
DRYV_HD uint32_t viaddmin_relu_s16x2(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  return __viaddmin_s16x2_relu(a, b, c);
#else
  uint32_t r = 0;
  for (int h = 0; h < 2; h++) {
    int s = (int)(int16_t)((a >> (16 * h)) + (b >> (16 * h)));
    const int m = (int)(int16_t)(c >> (16 * h));
    s = s < m ? s : m;
    s = s > 0 ? s : 0;
    r |= ((uint32_t)s & 0xffffu) << (16 * h);
  }
  return r;
#endif
}
// max(min(a, b), 0)
DRYV_HD int vimin_relu_s32(int a, int b) {
#ifdef __CUDA_ARCH__
  return __vimin_s32_relu(a, b);
#else
  const int m = a < b ? a : b;
  return m > 0 ? m : 0;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// residual hand-off format
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kResBias = 512;                 // a residual travels as the 16-bit field clamp(r, -512, 511) + 512
constexpr uint32_t kResBiasPair = 0x02000200u;
constexpr int kRowBias = 8192;                // bias carried by the row-pass outputs of the packed 4x4 path
DRYV_HD constexpr uint32_t pk2(uint32_t v) { return v * 0x10001u; }

// two biased residual fields + two prediction samples -> two final samples (one per field), clip(pred + r, 0, 255).
//   p: the prediction samples as bytes of a word, sel picks two of them into the low bytes of the fields, the constant
//   0xfe bytes make each field pred - 512 so the bias of the residual cancels inside the add.
DRYV_HD uint32_t add_clip_pair(uint32_t res2, uint32_t pred_bytes, uint32_t sel) {
  return viaddmin_relu_s16x2(res2, prmt(pred_bytes, 0xfefefefeu, sel), 0x00ff00ffu);
}
// four samples: residual words r01 (samples 0, 1) and r23, prediction bytes p -> four clipped bytes
DRYV_HD uint32_t add_clip4(uint32_t r01, uint32_t r23, uint32_t p) {
  const uint32_t a = add_clip_pair(r01, p, 0x4140u), b = add_clip_pair(r23, p, 0x4342u);
  return prmt(a, b, 0x6420u);
}

// ---------------------------------------------------------------------------------------------------------------------
// 4x4 block, lane-local. Levels in the reference's zig-zag array order (frame/mod.rs:185-210):
//   row 0: k = 0 1 5 6 | row 1: 2 4 7 12 | row 2: 3 8 11 13 | row 3: 9 10 14 15
// ---------------------------------------------------------------------------------------------------------------------
// transform.rs:159-169, one row; the outputs inherit any constant added to `a` (each contains +a exactly once)
DRYV_HD void row4(int a, int b, int c, int d, int& o0, int& o1, int& o2, int& o3) {
  const int e2 = (b >> 1) - d, e3 = b + (d >> 1);
  const int s = a + c, t = a - c;
  o0 = s + e3;
  o1 = t + e2;
  o2 = t - e2;
  o3 = s - e3;
}

// transform.rs:171-181 on two columns at once. F0..F3: rows 0..3 of two neighbouring columns, 16-bit fields f + 8192 in
// [0, 16384). O0..O3: fields (column-pass output) + 32768. Every intermediate stays inside [0, 65536):
//   S, D in [0, 32768]; F1s, F3s in [0, 8192); E*, NE* = +-e + 16384 in [4096, 28672]; O in [4096, 61440].
DRYV_HD void col4_packed(uint32_t F0, uint32_t F1, uint32_t F2, uint32_t F3, uint32_t& O0, uint32_t& O1, uint32_t& O2,
                         uint32_t& O3) {
  const uint32_t F1s = (F1 >> 1) & 0x7fff7fffu, F3s = (F3 >> 1) & 0x7fff7fffu;  // (f >> 1) + 4096 per field
  const uint32_t S = F0 + F2;                      // f0 + f2 + 16384
  const uint32_t D = F0 - F2 + pk2(16384);         // f0 - f2 + 16384
  const uint32_t E3 = F1 + F3s + pk2(4096);        // e3 + 16384,  e3 = f1 + (f3 >> 1)
  const uint32_t NE3 = pk2(28672) - F1 - F3s;      // -e3 + 16384
  const uint32_t E2 = F1s - F3 + pk2(20480);       // e2 + 16384,  e2 = (f1 >> 1) - f3
  const uint32_t NE2 = F3 - F1s + pk2(12288);      // -e2 + 16384
  O0 = S + E3;
  O1 = D + E2;
  O2 = D + NE2;
  O3 = S + NE3;
}

// Row pass on the 16 dequantised values d[k] (zig-zag order; d[0], d[2], d[3], d[9] carry +8192, d[0] also the +32 of the
// final rounding) -> f[i][j] + 8192. Returns the OR of all outputs: a bit at or above bit 14 means some f is outside
// [-8192, 8192) and the packed column pass must not be used.
DRYV_HD uint32_t rows4x4(const int d[16], int f[4][4]) {
  row4(d[0], d[1], d[5], d[6], f[0][0], f[0][1], f[0][2], f[0][3]);
  row4(d[2], d[4], d[7], d[12], f[1][0], f[1][1], f[1][2], f[1][3]);
  row4(d[3], d[8], d[11], d[13], f[2][0], f[2][1], f[2][2], f[2][3]);
  row4(d[9], d[10], d[14], d[15], f[3][0], f[3][1], f[3][2], f[3][3]);
  uint32_t g = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) g |= (uint32_t)f[i][0] | (uint32_t)f[i][1] | (uint32_t)f[i][2] | (uint32_t)f[i][3];
  return g;
}

// Column pass + rounding shift, packed. f: row-pass outputs in units of U = 2^e, biased by 8192 >> e, all inside
// [0, 16384 >> e); packing scales them back (two IMADs per word), so the fields are f_true + 8192.
// out[2 * i + jp]: samples (2 jp, 2 jp + 1) of row i as biased residual fields.
DRYV_HD void cols4x4_packed(const int f[4][4], uint32_t U, uint32_t out[8]) {
  const uint32_t U16 = U << 16;
#pragma unroll
  for (int jp = 0; jp < 2; jp++) {
    uint32_t F[4], O[4];
#pragma unroll
    for (int i = 0; i < 4; i++) F[i] = (uint32_t)f[i][2 * jp] * U + (uint32_t)f[i][2 * jp + 1] * U16;
    col4_packed(F[0], F[1], F[2], F[3], O[0], O[1], O[2], O[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) out[2 * i + jp] = (O[i] >> 6) & 0x03ff03ffu;  // ((o + 32768) >> 6) = (o >> 6) + 512
  }
}

// The same on 32-bit scalars (any int32-representable input); f unbiased.
DRYV_HD void cols4x4_wide(const int f[4][4], uint32_t out[8]) {
#pragma unroll
  for (int jp = 0; jp < 2; jp++) {
    int r[2][4];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int j = 2 * jp + q;
      row4(f[0][j], f[1][j], f[2][j], f[3][j], r[q][0], r[q][1], r[q][2], r[q][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int lo = vimin_relu_s32((r[0][i] >> 6) + kResBias, 1023), hi = vimin_relu_s32((r[1][i] >> 6) + kResBias, 1023);
      out[2 * i + jp] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
  }
}

// One whole block on one lane, byte-scale dequantisation (DeviceTables::t4b / t4b_e): levels cw (two per word, as loaded)
// -> eight words of biased residual fields. Everything up to the packing runs in units of 2^e (exact: the factors are
// multiples of 4 when e > 0, and the DC, which may be anything, meets no shift on its way — dropping its low e bits
// cannot change floor((h + 32) / 64)).
//   dc_pass: the DC comes from the Intra16x16 / chroma DC transform (already dequantised, transform.rs:145-146).
// Returns false when a row-pass output lies outside [-8192, 8192): `out` is then meaningless and the block has to go
// through block4x4_wide (no conforming stream produces such a block).
DRYV_HD bool block4x4_fast(const uint32_t cw[8], const uint32_t bs[8], int e, bool dc_pass, int dcv, uint32_t out[8]) {
  const int rb = kRowBias >> e, rb0 = rb + (32 >> e);
  int d[16], f[4][4];
#pragma unroll
  for (int w = 0; w < 8; w++) {
    const int k0 = 2 * w, k1 = 2 * w + 1;
    d[k0] = dp2a_lo_su(cw[w], bs[w], k0 == 0 ? rb0 : (k0 == 2 ? rb : 0));
    d[k1] = dp2a_hi_su(cw[w], bs[w], (k1 == 3 || k1 == 9) ? rb : 0);
  }
  if (dc_pass) d[0] = ((dcv + 32) >> e) + rb;
  const uint32_t g = rows4x4(d, f);
  cols4x4_packed(f, 1u << e, out);
  return (g & ~((16384u >> e) - 1u)) == 0;
}

// The whole block in 32-bit arithmetic with the general dequantisation, transform.rs:143-155: tt = t4[qP]
// (LevelScale << max(qP/6 - 4, 0)), shr = max(4 - qP/6, 0). Any qP, any scaling list, any level whose products fit int32.
DRYV_HD void block4x4_wide(const uint32_t cw[8], const int tt[16], int shr, bool dc_pass, int dcv, uint32_t out[8]) {
  const int rnd = shr > 0 ? (1 << (shr - 1)) : 0;
  int d[16], f[4][4];
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int v = (k & 1) ? ((int)cw[k >> 1] >> 16) : (int)(int16_t)(cw[k >> 1] & 0xffffu);
    d[k] = (v * tt[k] + rnd) >> shr;
  }
  if (dc_pass) d[0] = dcv;
  d[0] += 32;
  rows4x4(d, f);
  cols4x4_wide(f, out);
}

// 8-point inverse transform, pred8x8.rs:85-112
DRYV_HD void idct8(int* d) {
  const int e0 = d[0] + d[4];
  const int e1 = -d[3] + d[5] - d[7] - (d[7] >> 1);
  const int e2 = d[0] - d[4];
  const int e3 = d[1] + d[7] - d[3] - (d[3] >> 1);
  const int e4 = (d[2] >> 1) - d[6];
  const int e5 = -d[1] + d[7] + d[5] + (d[5] >> 1);
  const int e6 = d[2] + (d[6] >> 1);
  const int e7 = d[3] + d[5] + d[1] + (d[1] >> 1);
  const int f0 = e0 + e6, f1 = e1 + (e7 >> 2), f2 = e2 + e4, f3 = e3 + (e5 >> 2);
  const int f4 = e2 - e4, f5 = (e3 >> 2) - e5, f6 = e0 - e6, f7 = e7 - (e1 >> 2);
  d[0] = f0 + f7;
  d[1] = f2 + f5;
  d[2] = f4 + f3;
  d[3] = f6 + f1;
  d[4] = f6 - f1;
  d[5] = f4 - f3;
  d[6] = f2 - f5;
  d[7] = f0 - f7;
}

#if defined(__CUDACC__)
// =====================================================================================================================
// warp-level passes (device only)
// =====================================================================================================================
constexpr int kGroupMbs = 4;             // macroblocks per group
constexpr int kScratchWords = 4 * 72;    // 8x8 transposes: four blocks of 8 rows x 8 words (+ 8 so that blocks land on different banks)
// Residual tiles in shared memory (16-bit fields). The strides are chosen for the writers, which store 8 bytes per lane
// with one 4x4 block per lane: a luma row stride of 40 bytes puts the four block rows of a macroblock 32 bytes apart
// modulo 128, a chroma plane stride of 144 bytes and a macroblock stride that is 32 modulo 128 (kResChromaMb) do the
// same for the eight chroma blocks of two macroblocks: every store instruction is free of bank conflicts.
// kResLumaStride = 20 fields per luma row (16 used): recon_tables.h
constexpr int kResLumaTile = 16 * kResLumaStride;    // fields per macroblock
constexpr int kResChromaPlane = 72;                  // fields per chroma plane (8 rows of 8, + 8)
constexpr int kResChromaMb = 2 * kResChromaPlane;    // fields per macroblock: 288 bytes

// Lane constants of the residual passes.
struct ResLane {
  int res_off_luma;      // field offset of luma block (lane & 15) inside a luma residual tile
  int res_off_chroma;    // field offset of chroma block (lane & 7) inside a chroma residual tile
  uint32_t dc_partners;  // Intra16x16 luma-DC Hadamard: 4 x 5 bits partner lanes (within the half-warp) + routing source
  uint32_t dc_signs;
  uint32_t zz8_lo, zz8_hi;  // zig-zag indices of row (lane & 7) of an 8x8 block, one byte per column
};

__device__ __forceinline__ ResLane make_res_lane(int lane, const DeviceTables& tab) {
  ResLane lc;
  {
    const int b = lane & 15;  // 4x4 block position (spec block order, pred4x4.rs:14-17)
    const int bx = ((b >> 2) & 1) * 8 + (b & 1) * 4, by = (b >> 3) * 8 + ((b >> 1) & 1) * 4;
    lc.res_off_luma = by * kResLumaStride + bx;
  }
  {
    const int b = lane & 3, pl = (lane >> 2) & 1;
    lc.res_off_chroma = pl * kResChromaPlane + (b >> 1) * 32 + (b & 1) * 4;
  }
  // luma DC: lane L (< 16) holds c[i][j] with (i, j) = zig-zag position of L (pred16x16.rs:436-444)
  const int zi[16] = {0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 2, 3, 3};
  const int zj[16] = {0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 1, 2, 3, 3, 2, 3};
  const int inv[4][4] = {{0, 1, 5, 6}, {2, 4, 7, 12}, {3, 8, 11, 13}, {9, 10, 14, 15}};
  const int L = lane & 15;
  const int i = zi[L], j = zj[L];
  uint32_t partners = (uint32_t)inv[i][j ^ 1] | ((uint32_t)inv[i][j ^ 2] << 5) | ((uint32_t)inv[i ^ 1][j] << 10) |
                      ((uint32_t)inv[i ^ 2][j] << 15);
  uint32_t signs = 0;
  if (j & 1) signs |= 1u << 0;   // stage 0: own negative
  if (j == 2) signs |= 1u << 2;  // stage 1: own negative
  if (j == 1) signs |= 1u << 3;  // stage 1: other negative
  if (i & 1) signs |= 1u << 4;
  if (i == 2) signs |= 1u << 6;
  if (i == 1) signs |= 1u << 7;
  {
    // after the four stages the lane at (i, j) holds f[s(i)][s(j)], s = swap(1, 2); block L wants dcY[by][bx]
    // (pred16x16.rs:27-31) -> source lane inv[s(by)][s(bx)]
    const int gx = ((L >> 2) & 1) * 2 + (L & 1), gy = (L >> 3) * 2 + ((L >> 1) & 1);
    const int s[4] = {0, 2, 1, 3};
    partners |= (uint32_t)inv[s[gy]][s[gx]] << 20;
  }
  lc.dc_partners = partners;
  lc.dc_signs = signs;
  const int r = lane & 7;
  uint32_t lo = 0, hi = 0;
  for (int c = 0; c < 4; c++) lo |= (uint32_t)tab.zz8inv[r][c] << (8 * c);
  for (int c = 0; c < 4; c++) hi |= (uint32_t)tab.zz8inv[r][4 + c] << (8 * c);
  lc.zz8_lo = lo;
  lc.zz8_hi = hi;
  return lc;
}

// The rare lanes of a pass: blocks outside the packed range and qP without a byte-scale form. A real function, so that it
// exists once and stays out of the instruction stream of the passes (the kernels are sensitive to code size); everything
// is recomputed from the levels, so nothing but registers crosses the call.
struct Out8 {
  uint4 a, b;
};
__device__ __noinline__ Out8 pass4x4_wide(const DeviceTables* gtab, uint4 c0, uint4 c1, int qpl, bool dc_pass, int dcv) {
  const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const int4* tp = reinterpret_cast<const int4*>(&gtab->t4[qpl][0]);
  const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2), t3 = __ldg(tp + 3);
  const int tt[16] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w};
  const int qpd = qpl / 6;
  uint32_t out[8];
  block4x4_wide(cw, tt, qpd < 4 ? 4 - qpd : 0, dc_pass, dcv, out);
  Out8 r;
  r.a = make_uint4(out[0], out[1], out[2], out[3]);
  r.b = make_uint4(out[4], out[5], out[6], out[7]);
  return r;
}

// One pass of 4x4 blocks, one per lane, result in registers: out[2 * i + jp] = samples (2 jp, 2 jp + 1) of row i of the
// lane's block as biased residual fields.
//   c0, c1   the lane's 16 levels;  qpl its qP (QP'Y or QPc);  dc_pass / dcv: see block4x4_fast;  active: the lane has a block
//   tab: the shared-memory copy (residual part), gtab: the whole table in global memory (t4, for qP without a byte form)
__device__ __forceinline__ void pass4x4_regs(const DeviceTables& tab, const DeviceTables* gtab, uint4 c0, uint4 c1, int qpl,
                                             bool dc_pass, int dcv, bool active, uint32_t out[8]) {
  const uint32_t cw[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
  const int e = tab.t4b_e[qpl];
  const uint4* bp = reinterpret_cast<const uint4*>(&tab.t4b[qpl][0]);
  const uint4 b0 = bp[0], b1 = bp[1];
  const uint32_t bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const bool ok = block4x4_fast(cw, bs, e & 7, dc_pass, dcv, out) && e != 0xff;
  if (__any_sync(0xffffffffu, active && !ok)) {
    const Out8 w = pass4x4_wide(gtab, c0, c1, qpl, dc_pass, dcv);
    if (!ok) {
      out[0] = w.a.x; out[1] = w.a.y; out[2] = w.a.z; out[3] = w.a.w;
      out[4] = w.b.x; out[5] = w.b.y; out[6] = w.b.z; out[7] = w.b.w;
    }
  }
}

// The same with the result stored to a residual tile in shared memory.
//   dst      the block's first residual field, `stride` fields per sample row (20 luma, 8 chroma); null: lane idle
__device__ __forceinline__ void pass4x4_body(const DeviceTables& tab, const DeviceTables* gtab, uint4 c0, uint4 c1, int qpl,
                                             bool dc_pass, int dcv, uint16_t* dst, int stride) {
  uint32_t out[8];
  pass4x4_regs(tab, gtab, c0, c1, qpl, dc_pass, dcv, dst != nullptr, out);
  if (dst) {
#pragma unroll
    for (int i = 0; i < 4; i++) *reinterpret_cast<uint2*>(dst + i * stride) = make_uint2(out[2 * i], out[2 * i + 1]);
  }
}
// The wavefront kernel calls the pass (one copy for the luma and the chroma passes: the kernel sits at the capacity of the
// instruction cache, profiles/r02_ifetch.txt); the residual-only kernel inlines it (a call drains the scoreboard, which
// would expose the latency of the prediction loads issued ahead of the passes).
__device__ __noinline__ void pass4x4_call(const DeviceTables& tab, const DeviceTables* gtab, uint4 c0, uint4 c1, int qpl,
                                          bool dc_pass, int dcv, uint16_t* dst, int stride) {
  pass4x4_body(tab, gtab, c0, c1, qpl, dc_pass, dcv, dst, stride);
}
template <bool INLINE>
__device__ __forceinline__ void pass4x4(const DeviceTables& tab, const DeviceTables* gtab, uint4 c0, uint4 c1, int qpl,
                                        bool dc_pass, int dcv, uint16_t* dst, int stride) {
  if (INLINE) pass4x4_body(tab, gtab, c0, c1, qpl, dc_pass, dcv, dst, stride);
  else pass4x4_call(tab, gtab, c0, c1, qpl, dc_pass, dcv, dst, stride);
}

// Intra16x16 luma DC (pred16x16.rs:428-482) for the 16 lanes of a half-warp: v0 = level 0 of the lane's block.
// Returns the dequantised DC of the lane's block.
__device__ __forceinline__ int luma_dc16(const DeviceTables& tab, const ResLane& lc, int lane, int v0, int qp) {
  const int hb = lane & 16;
  int t = v0;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int o = __shfl_sync(0xffffffffu, t, hb | (int)((lc.dc_partners >> (5 * s)) & 31u));
    const int so = (lc.dc_signs >> (2 * s)) & 1, sp = (lc.dc_signs >> (2 * s + 1)) & 1;
    t = (so ? -t : t) + (sp ? -o : o);
  }
  const int qpm = qp % 6, qpd = qp / 6;
  const int ls00 = tab.ls00[qpm];
  const int dq = qp >= 36 ? ((t * ls00) << (qpd - 6)) : ((t * ls00 + (1 << (5 - qpd))) >> (6 - qpd));
  return __shfl_sync(0xffffffffu, dq, hb | (int)((lc.dc_partners >> 20) & 31u));
}

// Chroma DC (trans_chroma.rs:389-415) over the four lanes of a plane (lane & 3 = block): f = H c H, ((f * LS) << (qP/6)) >> 5
__device__ __forceinline__ int chroma_dc(const DeviceTables& tab, int lane, int v0, int qpc) {
  int o = __shfl_xor_sync(0xffffffffu, v0, 1);
  int t = (lane & 1) ? o - v0 : v0 + o;
  o = __shfl_xor_sync(0xffffffffu, t, 2);
  t = (lane & 2) ? o - t : t + o;
  return ((t * tab.ls00[qpc % 6]) << (qpc / 6)) >> 5;
}

// Intra8x8 luma of one macroblock (pred8x8.rs:51-150): 4 blocks x 8 lanes, lane i of a block = row i, then column i.
//   lv: the macroblock's 256 luma levels (shared memory), scratch: kScratchWords words, res: luma residual tile
__device__ __forceinline__ void pass8x8(const DeviceTables& tab, const ResLane& lc, int lane, const int16_t* lv,
                                        int* scratch, int qp, uint16_t* res) {
  const int blk = lane >> 3, i = lane & 7;
  const int16_t* slab = lv + blk * 64;
  const int qpm = qp % 6, qpd = qp / 6;
  const uint4 lsv = *reinterpret_cast<const uint4*>(&tab.ls8[qpm][i * 8]);
  const uint32_t lsw[4] = {lsv.x, lsv.y, lsv.z, lsv.w};
  int d[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const uint32_t zw = j < 4 ? lc.zz8_lo : lc.zz8_hi;
    const int k = (zw >> (8 * (j & 3))) & 0xff;
    const int c = slab[k];
    const int ls = (j & 1) ? (int)(lsw[j >> 1] >> 16) : (int)(lsw[j >> 1] & 0xffffu);
    d[j] = qp >= 36 ? ((c * ls) << (qpd - 6)) : ((c * ls + (1 << (5 - qpd))) >> (6 - qpd));  // pred8x8.rs:71-80
  }
  // d00 reaches every output sample with weight 1 and through no shift: fold the final (m + 32) >> 6 rounding and the
  // residual bias (512 << 6) into it
  if (i == 0) d[0] += 32 + (kResBias << 6);
  idct8(d);
  // transpose through shared memory; the two halves of rows 4..7 change places so that the 128-bit stores of a
  // quarter-warp (rows 0..7 of one block, 32 bytes apart) fall on different banks
  int* tb = scratch + blk * 72;
  const int sw = (i & 4);
  *reinterpret_cast<int4*>(tb + i * 8 + sw) = make_int4(d[0], d[1], d[2], d[3]);
  *reinterpret_cast<int4*>(tb + i * 8 + (sw ^ 4)) = make_int4(d[4], d[5], d[6], d[7]);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 8; r++) d[r] = tb[r * 8 + (r < 4 ? i : (i ^ 4))];
  idct8(d);
  uint16_t* rl = res + ((blk >> 1) * 8) * kResLumaStride + (blk & 1) * 8 + i;
#pragma unroll
  for (int r = 0; r < 8; r++) rl[r * kResLumaStride] = (uint16_t)vimin_relu_s32(d[r] >> 6, 1023);
  __syncwarp();
}

// The residual stage of one group.
//   hdr[m]    header of macroblock m of the group: mb class | qp << 8 (shared memory, kGroupMbs words)
//   m4, m8    which macroblocks of the group take the 4x4 / the 8x8 luma transform (bit m)
//   lv        the group's levels as they lie in HBM (384 int16 per macroblock), in shared memory
//   n_mb      macroblocks in the group (1..4)
//   luma / chroma + m * stride: macroblock m's luma and chroma residual tiles (kResLumaTile / kResChromaMb fields)
template <bool INLINE_PASS>
__device__ __forceinline__ void residual_group(const DeviceTables& tab, const DeviceTables* gtab, const ResLane& lc, int lane, const uint32_t* hdr,
                                               uint32_t m4, uint32_t m8, const int16_t* lv, int n_mb, int* scratch,
                                               uint16_t* luma, int luma_mb_stride, uint16_t* chroma,
                                               int chroma_mb_stride, int cb_off, int cr_off) {
  // ---- luma, 4x4 transform: two macroblocks per pass ----
  {
    const uint32_t list = tab.setbits4[m4];
    const int n4 = __popc(m4);
    for (int p = 0; 2 * p < n4; p++) {
      const uint32_t m = (list >> (4 * (2 * p + (lane >> 4)))) & 15u;
      const bool active = m < (uint32_t)kGroupMbs;
      const uint32_t mm = active ? m : 0u;
      const uint32_t h = hdr[mm];
      const int qp = (int)((h >> 8) & 0xffu);
      const bool i16 = (h & 0xffu) == 2u;
      const uint4* src = reinterpret_cast<const uint4*>(lv + mm * DRYV_COEFFS_PER_MB + (lane & 15) * 16);
      const uint4 c0 = src[0], c1 = src[1];
      int dcv = 0;
      if (__any_sync(0xffffffffu, active && i16)) dcv = luma_dc16(tab, lc, lane, (int)(int16_t)(c0.x & 0xffffu), qp);
      pass4x4<INLINE_PASS>(tab, gtab, c0, c1, qp, i16, dcv, active ? luma + mm * luma_mb_stride + lc.res_off_luma : nullptr, kResLumaStride);
    }
  }
  // ---- luma, 8x8 transform: one macroblock per pass ----
  while (m8) {
    const int m = __ffs(m8) - 1;
    m8 &= m8 - 1;
    pass8x8(tab, lc, lane, lv + m * DRYV_COEFFS_PER_MB, scratch, (int)((hdr[m] >> 8) & 0xffu), luma + m * luma_mb_stride);
  }
  // ---- chroma: the 8 blocks of each of the four macroblocks in one pass ----
  {
    const int m = lane >> 3;
    const bool active = m < n_mb;
    const int mm = active ? m : 0;
    int q = (int)((hdr[mm] >> 8) & 0xffu) + ((lane & 4) ? cr_off : cb_off);
    q = min(max(q, 0), 51);
    const int qpc = tab.qpc[q];  // transform.rs:194-216
    const uint4* src = reinterpret_cast<const uint4*>(lv + mm * DRYV_COEFFS_PER_MB + 256 + (lane & 7) * 16);
    const uint4 c0 = src[0], c1 = src[1];
    const int dcv = chroma_dc(tab, lane, (int)(int16_t)(c0.x & 0xffffu), qpc);
    pass4x4<INLINE_PASS>(tab, gtab, c0, c1, qpc, true, dcv, active ? chroma + mm * chroma_mb_stride + lc.res_off_chroma : nullptr, 8);
  }
  __syncwarp();
}
#endif  // __CUDACC__

}  // namespace dryv
