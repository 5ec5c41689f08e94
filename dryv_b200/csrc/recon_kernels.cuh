// recon_kernels.cuh — sm_100a device code of the AVC intra reconstruction path.
//
// One warp owns one macroblock row of one picture ("row walker") and walks it left to right. Per MB:
//   1. residual stage  (no dependencies): 128-bit loads of the MB's 768 B of int16 levels, inverse
//      zig-zag as a register permutation, dequant, luma-DC / chroma-DC Hadamard (warp shuffles), the
//      4x4 transform with one block per lane entirely in registers (24 lanes = 16 luma + 8 chroma
//      blocks) or the 8x8 transform with 8 lanes per block and a shared-memory transpose; int16
//      residuals land in a per-warp shared-memory tile.
//   2. wait until the row above has finished MB x+1 (x+2y wavefront, per-row progress counter,
//      ld.acquire / st.release at gpu scope).
//   3. prediction + residual add + clip into a shared-memory pixel tile (left column carried over in
//      shared memory from the previous MB, top strip re-read from the frame).
//   4. 128-bit row stores of the finished MB, then the row's progress counter is published.
// No tensor cores: the H.264 transforms are shift/add butterflies with exact integer rounding.
//
// Reference behaviour reproduced (paths relative to the reference root, src/video/frame/):
//   transform.rs:116-191, pred8x8.rs:51-150, pred16x16.rs:428-482, trans_chroma.rs:369-456 (residual),
//   pred4x4.rs:10-427, pred8x8.rs:152-764, pred16x16.rs:79-425, trans_chroma.rs:96-366 (prediction),
//   including the reference's deviations from the H.264 text listed in SURVEY.md §8 (Q1-Q5).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "recon_tables.h"

namespace dryv {

constexpr int kWarpsPerCta = 4;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;

constexpr int kLumaStride = 48;          // pixel (x, y) at (y + 1) * 48 + 16 + x, x in -4..31 (row -1), y in -1..15
constexpr int kLumaTileBytes = 17 * kLumaStride;
constexpr int kChromaStride = 24;        // pixel (x, y) at (y + 1) * 24 + 8 + x, x in -4..15 (row -1), y in -1..7
constexpr int kChromaTileBytes = 9 * kChromaStride;
constexpr int kScratchBytes = 1152;      // 8x8 coefficient slab (4 * 144 B) aliased with the transpose buffer (4 * 72 words)

// residual-only kernel: one warp per macroblock
struct ResidWarpSmem {
  alignas(16) int16_t res[384];  // luma [16][16] | cb [8][8] | cr [8][8]
  alignas(16) uint8_t scratch[kScratchBytes];
};
struct ResidCtaSmem {
  DeviceTables tab;
  ResidWarpSmem warp[kWarpsPerCta];
};

// wavefront kernel: a "row team" = one CTA of two warps walking one macroblock row.
//   front warp: residual (luma + chroma) and prediction-mode derivation
//   pixel warp: luma + chroma prediction, stores
// The front warp hands each macroblock to the luma warp through a ring of kSlots slots.
constexpr int kSlots = 4;
constexpr int kTeamThreads = 64;
struct Slot {
  alignas(16) int16_t res[256];  // luma residual [16][16]
  alignas(16) int16_t cres[128]; // chroma residual [2][8][8]
  uint32_t modes_lo, modes_hi;   // resolved Intra4x4/8x8 modes of the raster 4x4 grid cells 0..7 / 8..15, 4 bits each
  int32_t frame, row, x;         // row < 0: no more work
  int32_t mbcls, mode16;         // 0/1/2 = Intra4x4/8x8/16x16; Intra16x16 prediction mode | intra_chroma_pred_mode << 8
  int32_t pad[1];
};
static_assert(sizeof(Slot) % 16 == 0, "slot alignment");
struct TeamSmem {
  DeviceTables tab;
  Slot slot[kSlots];
  alignas(16) uint8_t luma[kLumaTileBytes];         // luma pixel tile (pixel warp)
  alignas(16) uint8_t chroma[2 * kChromaTileBytes]; // chroma pixel tiles (pixel warp)
  alignas(16) uint8_t scratch[kScratchBytes];
  alignas(8) unsigned long long full[kSlots];       // mbarriers: slot filled by the front warp
  alignas(8) unsigned long long empty[kSlots];      // mbarriers: slot released by the pixel warp
};

enum { STATUS_OK = 0, STATUS_UNSUPPORTED = 1, STATUS_WATCHDOG = 2 };

// Bottom line a macroblock hands to the row below: 4 luma words (16 px), 2 Cb, 2 Cr words (8 px each)
// and one word with the resolved prediction modes of its bottom 4x4 blocks. Every 32-bit payload
// travels with the launch tag in one 64-bit word, so a single relaxed 64-bit load both fetches the
// data and proves it is there: no fence, no separate flag, no second round trip.
constexpr int kLineWords = 9;

struct KernelArgs {
  const uint8_t* mb_type;
  const uint8_t* t8x8;
  const uint8_t* chroma_mode;
  const uint8_t* qp;
  const uint8_t* pred_syntax;
  const int16_t* coeff;
  uint8_t* out;             // n_frames pictures, Y | Cb | Cr each
  const uint8_t* pred_in;   // residual-add kernel only
  const DeviceTables* tables;
  unsigned long long* line; // [n_frames * H * W][kLineWords] bottom line of each MB: payload | tag << 32
  uint32_t tag;             // launch tag: a line word is valid when its upper half equals it
  unsigned int* ticket;     // row ticket counter
  unsigned long long* prof; // stage clocks (development builds), may be null
  int* status;
  int W, H, n_frames;
  int cb_off, cr_off;
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_relaxed_gpu_s32(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
// two residuals -> one s16x2 word, saturating: the final sample is clip(pred + r, 0, 255) with pred in 0..255,
// so any r beyond +-255 already pins the result and saturating at +-32767 keeps every input exact
__device__ __forceinline__ uint32_t pack2(int lo, int hi) {
  uint32_t d;
  asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
  return d;
}
__device__ __forceinline__ int16_t sat16(int v) { return (int16_t)min(max(v, -32768), 32767); }
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
  return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t w) { return ((int)w) >> 16; }

// 4-point inverse core transform, transform.rs:159-181
__device__ __forceinline__ void idct4(int& a, int& b, int& c, int& d) {
  int e0 = a + c, e1 = a - c, e2 = (b >> 1) - d, e3 = b + (d >> 1);
  a = e0 + e3;
  b = e1 + e2;
  c = e1 - e2;
  d = e0 - e3;
}
// 8-point inverse transform, pred8x8.rs:85-112
__device__ __forceinline__ void idct8(int* d) {
  int e0 = d[0] + d[4];
  int e1 = -d[3] + d[5] - d[7] - (d[7] >> 1);
  int e2 = d[0] - d[4];
  int e3 = d[1] + d[7] - d[3] - (d[3] >> 1);
  int e4 = (d[2] >> 1) - d[6];
  int e5 = -d[1] + d[7] + d[5] + (d[5] >> 1);
  int e6 = d[2] + (d[6] >> 1);
  int e7 = d[3] + d[5] + d[1] + (d[1] >> 1);
  int f0 = e0 + e6, f1 = e1 + (e7 >> 2), f2 = e2 + e4, f3 = e3 + (e5 >> 2);
  int f4 = e2 - e4, f5 = (e3 >> 2) - e5, f6 = e0 - e6, f7 = e7 - (e1 >> 2);
  d[0] = f0 + f7;
  d[1] = f2 + f5;
  d[2] = f4 + f3;
  d[3] = f6 + f1;
  d[4] = f6 - f1;
  d[5] = f4 - f3;
  d[6] = f2 - f5;
  d[7] = f0 - f7;
}

// Per-lane constants that do not change over the kernel.
struct LaneConst {
  int res_off;      // int16 offset of this lane's 4x4 block inside the luma (lanes 0..15) / chroma (16..23) residual tile
  int res_stride;   // 16 (luma) or 8 (chroma)
  // Intra16x16 luma-DC Hadamard: partner lanes / signs of the four butterfly stages + final routing
  uint32_t dc_partners;  // 5 x 5 bits: stage0..3 partner lane, then routing source lane
  uint32_t dc_signs;     // bit 2s: own sign negative, bit 2s+1: other sign negative
  uint32_t zz8_lo, zz8_hi;  // zig-zag indices of row (lane & 7) of an 8x8 block, one byte per column
};

__device__ __forceinline__ LaneConst make_lane_const(int lane, const DeviceTables& tab) {
  LaneConst lc;
  // 4x4 block position (spec block order, pred4x4.rs:14-17)
  if (lane < 16) {
    int bx = ((lane >> 2) & 1) * 8 + (lane & 1) * 4;
    int by = (lane >> 3) * 8 + ((lane >> 1) & 1) * 4;
    lc.res_off = by * 16 + bx;
    lc.res_stride = 16;
  } else {
    int b = lane & 3, pl = (lane >> 2) & 1;
    lc.res_off = pl * 64 + (b >> 1) * 32 + (b & 1) * 4;
    lc.res_stride = 8;
  }
  // luma DC: lane L (< 16) holds c[i][j] with (i, j) = zig-zag position of L.
  const int zi[16] = {0, 0, 1, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 2, 3, 3};
  const int zj[16] = {0, 1, 0, 0, 1, 2, 3, 2, 1, 0, 1, 2, 3, 3, 2, 3};
  const int inv[4][4] = {{0, 1, 5, 6}, {2, 4, 7, 12}, {3, 8, 11, 13}, {9, 10, 14, 15}};
  int L = lane & 15;
  int i = zi[L], j = zj[L];
  uint32_t partners = 0, signs = 0;
  // stage 0: j ^ 1 ; stage 1: j ^ 2 ; stage 2: i ^ 1 ; stage 3: i ^ 2
  int p0 = inv[i][j ^ 1], p1 = inv[i][j ^ 2], p2 = inv[i ^ 1][j], p3 = inv[i ^ 2][j];
  partners = (uint32_t)p0 | ((uint32_t)p1 << 5) | ((uint32_t)p2 << 10) | ((uint32_t)p3 << 15);
  // stage "first" (pairs 0-1, 2-3): even index: own + other ; odd index: other - own
  // stage "second" (pairs 0-2, 1-3): idx0: own+other, idx2: other-own, idx1: own-other, idx3: own+other
  if (j & 1) signs |= 1u << 0;       // stage0 own negative
  if (j == 2) signs |= 1u << 2;      // stage1 own negative
  if (j == 1) signs |= 1u << 3;      // stage1 other negative
  if (i & 1) signs |= 1u << 4;       // stage2 own negative
  if (i == 2) signs |= 1u << 6;      // stage3 own negative
  if (i == 1) signs |= 1u << 7;      // stage3 other negative
  // after the four stages the lane at (i, j) holds f[s(i)][s(j)], s = swap(1, 2).
  // block b (= lane) wants dcY[by][bx] (pred16x16.rs:27-31) -> source lane inv[s(by)][s(bx)].
  {
    int gx = ((L >> 2) & 1) * 2 + (L & 1), gy = (L >> 3) * 2 + ((L >> 1) & 1);
    const int s[4] = {0, 2, 1, 3};
    partners |= (uint32_t)inv[s[gy]][s[gx]] << 20;
  }
  lc.dc_partners = partners;
  lc.dc_signs = signs;
  int r = lane & 7;
  uint32_t lo = 0, hi = 0;
  for (int c = 0; c < 4; c++) lo |= (uint32_t)tab.zz8inv[r][c] << (8 * c);
  for (int c = 0; c < 4; c++) hi |= (uint32_t)tab.zz8inv[r][4 + c] << (8 * c);
  lc.zz8_lo = lo;
  lc.zz8_hi = hi;
  return lc;
}

// ------------------------------------------------------------------------------------------------
// Residual stage: levels (registers c0, c1 of lanes 0..23) -> int16 residual tile.
//   mbcls: 0 Intra4x4, 1 Intra8x8, 2 Intra16x16.  qp: QP'Y of the MB.
// ------------------------------------------------------------------------------------------------
//   res_luma: int16 [16][16]; res_chroma: int16 [2][8][8]; scratch: kScratchBytes, 16-byte aligned.
__device__ __forceinline__ void residual_stage(const DeviceTables& tab, uint8_t* scratch, int16_t* res_luma,
                                               int16_t* res_chroma, const LaneConst& lc, int lane, uint4 c0, uint4 c1,
                                               int mbcls, int qp, int cb_off, int cr_off) {
  // ---- Intra8x8 luma: 8 lanes per block -----------------------------------------------------
  if (mbcls == 1) {
    if (lane < 16) {
      uint8_t* dst = scratch + (lane >> 2) * 144 + (lane & 3) * 32;
      *reinterpret_cast<uint4*>(dst) = c0;
      *reinterpret_cast<uint4*>(dst + 16) = c1;
    }
    __syncwarp();
    const int blk = lane >> 3, i = lane & 7;
    const uint8_t* slab = scratch + blk * 144;
    const int qpm = qp % 6, qpd = qp / 6;
    const uint4 lsv = *reinterpret_cast<const uint4*>(&tab.ls8[qpm][i * 8]);
    const uint32_t lsw[4] = {lsv.x, lsv.y, lsv.z, lsv.w};
    int d[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint32_t zw = j < 4 ? lc.zz8_lo : lc.zz8_hi;
      int k = (zw >> (8 * (j & 3))) & 0xff;
      int c = *reinterpret_cast<const int16_t*>(slab + 2 * k);
      int ls = (j & 1) ? (int)(lsw[j >> 1] >> 16) : (int)(lsw[j >> 1] & 0xffffu);
      // pred8x8.rs:71-80
      d[j] = qp >= 36 ? ((c * ls) << (qpd - 6)) : ((c * ls + (1 << (5 - qpd))) >> (6 - qpd));
    }
    if (i == 0) d[0] += 32;  // folds the final (m + 32) >> 6 rounding: d00 reaches every output with weight 1
    idct8(d);
    __syncwarp();
    int* tb = reinterpret_cast<int*>(scratch) + blk * 72;
    *reinterpret_cast<int4*>(tb + i * 8) = make_int4(d[0], d[1], d[2], d[3]);
    *reinterpret_cast<int4*>(tb + i * 8 + 4) = make_int4(d[4], d[5], d[6], d[7]);
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 8; r++) d[r] = tb[r * 8 + i];
    idct8(d);
    int16_t* rl = res_luma + ((blk >> 1) * 8) * 16 + (blk & 1) * 8 + i;
#pragma unroll
    for (int r = 0; r < 8; r++) rl[r * 16] = sat16(d[r] >> 6);
  }

  // ---- 4x4 path: one block per lane (luma lanes 0..15 unless Intra8x8, chroma lanes 16..23) --
  const bool is_chroma_lane = lane >= 16;
  int qpl = qp;
  if (is_chroma_lane) {
    int q = qp + (lane < 20 ? cb_off : cr_off);
    q = min(max(q, 0), 51);
    qpl = tab.qpc[q];  // transform.rs:194-216
  }
  int v[16];
  v[0] = lo16(c0.x); v[1] = hi16(c0.x); v[2] = lo16(c0.y); v[3] = hi16(c0.y);
  v[4] = lo16(c0.z); v[5] = hi16(c0.z); v[6] = lo16(c0.w); v[7] = hi16(c0.w);
  v[8] = lo16(c1.x); v[9] = hi16(c1.x); v[10] = lo16(c1.y); v[11] = hi16(c1.y);
  v[12] = lo16(c1.z); v[13] = hi16(c1.z); v[14] = lo16(c1.w); v[15] = hi16(c1.w);

  const int qpm = qpl % 6, qpd = qpl / 6;
  const int ls00 = tab.t4[qpm][0];  // LevelScale4x4[qP%6][0][0] (rows 0..5 of t4 carry no pre-shift)

  // chroma DC, trans_chroma.rs:389-415: f = H c H over the 4 lanes of a plane, then ((f*LS) << (qP/6)) >> 5
  int dcv;
  {
    int o = __shfl_xor_sync(0xffffffffu, v[0], 1);
    int t = (lane & 1) ? o - v[0] : v[0] + o;
    o = __shfl_xor_sync(0xffffffffu, t, 2);
    t = (lane & 2) ? o - t : t + o;
    dcv = ((t * ls00) << qpd) >> 5;
  }
  // Intra16x16 luma DC, pred16x16.rs:428-482 (warp-uniform branch)
  if (mbcls == 2) {
    int t = v[0];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      int o = __shfl_sync(0xffffffffu, t, (lc.dc_partners >> (5 * s)) & 31);
      int so = (lc.dc_signs >> (2 * s)) & 1, sp = (lc.dc_signs >> (2 * s + 1)) & 1;
      t = (so ? -t : t) + (sp ? -o : o);
    }
    // here qpl == qp for the luma lanes
    int dq = qp >= 36 ? ((t * ls00) << (qpd - 6)) : ((t * ls00 + (1 << (5 - qpd))) >> (6 - qpd));
    int routed = __shfl_sync(0xffffffffu, dq, (lc.dc_partners >> 20) & 31);
    if (!is_chroma_lane) dcv = routed;
  }
  const bool dc_pass = is_chroma_lane || mbcls == 2;  // transform.rs:145-146

  if (lane < 24 && (mbcls != 1 || is_chroma_lane)) {
    const int4* tp = reinterpret_cast<const int4*>(&tab.t4[qpl][0]);
    int4 t0 = tp[0], t1 = tp[1], t2 = tp[2], t3 = tp[3];
    const int tt[16] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z, t3.w};
    const int shr = max(4 - qpd, 0);
    const int rnd = qpd < 4 ? (1 << (3 - qpd)) : 0;
    // transform.rs:143-155; t4 carries LevelScale << max(qP/6-4, 0), so one form covers both branches
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (v[k] * tt[k] + rnd) >> shr;
    if (dc_pass) v[0] = dcv;
    v[0] += 32;  // folds (h + 32) >> 6: d00 reaches every output sample with weight 1 and no shift
    // zig-zag: (i, j) <- k   row0: 0 1 5 6 | row1: 2 4 7 12 | row2: 3 8 11 13 | row3: 9 10 14 15
    idct4(v[0], v[1], v[5], v[6]);
    idct4(v[2], v[4], v[7], v[12]);
    idct4(v[3], v[8], v[11], v[13]);
    idct4(v[9], v[10], v[14], v[15]);
    idct4(v[0], v[2], v[3], v[9]);
    idct4(v[1], v[4], v[8], v[10]);
    idct4(v[5], v[7], v[11], v[14]);
    idct4(v[6], v[12], v[13], v[15]);
    int16_t* r = (is_chroma_lane ? res_chroma : res_luma) + lc.res_off;
    const int st = lc.res_stride;
    *reinterpret_cast<uint2*>(r) = make_uint2(pack2(v[0] >> 6, v[1] >> 6), pack2(v[5] >> 6, v[6] >> 6));
    *reinterpret_cast<uint2*>(r + st) = make_uint2(pack2(v[2] >> 6, v[4] >> 6), pack2(v[7] >> 6, v[12] >> 6));
    *reinterpret_cast<uint2*>(r + 2 * st) = make_uint2(pack2(v[3] >> 6, v[8] >> 6), pack2(v[11] >> 6, v[13] >> 6));
    *reinterpret_cast<uint2*>(r + 3 * st) = make_uint2(pack2(v[9] >> 6, v[10] >> 6), pack2(v[14] >> 6, v[15] >> 6));
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Prediction stage helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int luma_at(int x, int y) { return (y + 1) * kLumaStride + 16 + x; }
__device__ __forceinline__ int chroma_at(int x, int y) { return (y + 1) * kChromaStride + 8 + x; }

// legal-mode mask of the nine 4x4/8x8 modes given neighbour availability (the reference writes no
// prediction when the mode's neighbours are missing, so the prediction stays 0: SURVEY quirk Q4)
__device__ __forceinline__ uint32_t legal_mask(bool t, bool l, bool c) {
  return 0x004u | (t ? 0x089u : 0u) | (l ? 0x102u : 0u) | ((t && l && c) ? 0x070u : 0u);
}

// Intra4x4 luma, pred4x4.rs:10-360 + transform.rs:98-110. Ten dependency steps (DeviceTables::i4step),
// two blocks per step where the decode-order availability rules allow it; one pixel per lane, 16 lanes
// per block. Kept as a rolled loop: the kernel is instruction-fetch sensitive.
//   av = A | B<<1 | C<<2 | D<<3 (macroblock availability), modes_lo/hi = 4-bit modes per raster cell.
__device__ __forceinline__ void predict_i4x4(const DeviceTables& tab, uint8_t* lt, const int16_t* res_luma, int lane,
                                             uint32_t modes_lo, uint32_t modes_hi, int av) {
  const int half = lane >> 4, p = lane & 15, px = p & 3, py = p >> 2;
  // edge sample fetched by this lane, relative to the block origin: 0..7 top / top-right (4..7 fall back to
  // sample 3 when top-right is missing), 8..11 left, 12 corner, 13..15 contribute 0
  const int edge_off = p < 8 ? (p - kLumaStride) : (p < 12 ? (p - 8) * kLumaStride - 1 : -kLumaStride - 1);
  const int edge_off_notr = (p >= 4 && p < 8) ? (3 - kLumaStride) : edge_off;
  const int guard_bit = p < 8 ? 25 : (p < 12 ? 26 : (p == 12 ? 27 : 31));  // availability bit guarding the sample
  const int pix_off = py * kLumaStride + px;
  const int res_lane = py * 16 + px;
  const uint32_t* steps = &tab.i4step[av][0][half];
  uint32_t nxt = steps[0];
#pragma unroll 1
  for (int s = 0; s < 10; s++) {
    const uint32_t cur = nxt;
    nxt = steps[s < 9 ? 2 * s + 2 : 18];
    const int org = cur & 1023;
    const int msh = (cur >> 10) & 31;
    const int mode = (((cur & 0x8000u) ? modes_hi : modes_lo) >> msh) & 15;
    int ev = lt[org + ((cur & (1u << 28)) ? edge_off : edge_off_notr)];
    if (!((cur >> guard_bit) & 1u)) ev = 0;  // unavailable samples (and lanes 13..15) count as 0
    const uint32_t taps = tab.lut4[mode > 8 ? 2 : mode][p];
    // residual of this lane's pixel: cell -> (cell >> 2) * 64 + (cell & 3) * 4 = msh-derived
    const int cell = (msh >> 2) | ((cur >> 12) & 8);
    const int res = res_luma[(cell >> 2) * 64 + (cell & 3) * 4 + res_lane];
    // __shfl_sync with width 16 takes the source lane modulo 16: no masking of the tap fields needed
    const int e0 = __shfl_sync(0xffffffffu, ev, taps, 16);
    const int e1 = __shfl_sync(0xffffffffu, ev, taps >> 4, 16);
    const int e2 = __shfl_sync(0xffffffffu, ev, taps >> 8, 16);
    const int t3 = e0 + e1 + e2;
    int pred = (t3 + e1 + 2) >> 2;
    if (__any_sync(0xffffffffu, mode == 2)) {
      // DC (pred4x4.rs:116-167): second summation round over the three partial sums held by lanes 0..2
      const int q = __shfl_sync(0xffffffffu, t3, 0, 16) + __shfl_sync(0xffffffffu, t3, 1, 16) +
                    __shfl_sync(0xffffffffu, t3, 2, 16);
      const int nav = ((cur >> 25) & 1) + ((cur >> 26) & 1);  // available sides: top, left
      const int dcv = nav == 2 ? ((q + 4) >> 3) : (nav == 1 ? ((q + 2) >> 2) : 128);
      if (mode == 2) pred = dcv;
    }
    if (!((cur >> (16 + mode)) & 1u) || mode > 8) pred = 0;  // mode needs a missing neighbour: prediction stays 0 (Q4)
    if (cur & (1u << 29)) lt[org + pix_off] = (uint8_t)clip255(pred + res);
    __syncwarp();
  }
}

// Intra8x8 luma, pred8x8.rs:152-696 + pred8x8.rs:34-46. Four sequential blocks, two pixels per lane.
__device__ __forceinline__ void predict_i8x8(const DeviceTables& tab, uint8_t* lt, const int16_t* res_luma, int lane,
                                             uint32_t modes_lo, uint32_t modes_hi, bool availA, bool availB,
                                             bool availC, bool availD) {
  const int py = lane >> 2, px = (lane & 3) * 2;
#pragma unroll 1
  for (int b = 0; b < 4; b++) {
    const int bx = (b & 1) * 8, by = (b >> 1) * 8;
    const int mode = (((b & 2) ? modes_hi : modes_lo) >> ((b & 1) * 8)) & 15;  // cells 0, 2, 8, 10
    const bool aL = bx > 0 || availA;
    const bool aT = by > 0 || availB;
    const bool aTL = b == 0 ? availD : (b == 1 ? availB : (b == 2 ? availA : true));
    const bool aTR = b == 0 ? availB : (b == 1 ? availC : (b == 2));
    const int m = mode > 8 ? 2 : mode;
    const uint32_t tw = *reinterpret_cast<const uint32_t*>(&tab.lut8[m][py * 8 + px]);
    const uint32_t rw = *reinterpret_cast<const uint32_t*>(&res_luma[(by + py) * 16 + bx + px]);
    // raw edge sample of this lane: 0..15 top, 16..23 left, 24 corner
    int ex, ey;
    if (lane < 16) { ex = bx + ((lane >= 8 && !aTR) ? 7 : lane); ey = by - 1; }
    else if (lane < 24) { ex = bx - 1; ey = by + (lane - 16); }
    else { ex = bx - 1; ey = by - 1; }
    const int raw = lt[luma_at(ex, ey)];
    // reference sample filter, pred8x8.rs:222-288 (with the x = 0 overwrite of quirk Q2)
    int srcp, srcn;  // lanes supplying the previous / next sample of the 3-tap filter
    if (lane < 16) { srcp = lane == 0 ? E8_CORNER : lane - 1; srcn = lane == 15 ? 15 : lane + 1; }
    else if (lane < 24) { srcp = lane == 16 ? (aTL ? E8_CORNER : 16) : lane - 1; srcn = lane == 23 ? 23 : lane + 1; }
    else { srcp = aT ? 0 : lane; srcn = aL ? 16 : lane; }
    int pv = __shfl_sync(0xffffffffu, raw, srcp & 31);
    const int nv = __shfl_sync(0xffffffffu, raw, srcn & 31);
    if (lane == 0 && !aTL) pv = -1;  // Q2: raw p[-1,-1] sentinel enters the filter
    int ev = (pv + 2 * raw + nv + 2) >> 2;
    if (mode == 2) {  // DC, pred8x8.rs:326-425 (warp-uniform): sums of the filtered top 0..7 and left 0..7
      int sum = ev + __shfl_xor_sync(0xffffffffu, ev, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      sum += __shfl_xor_sync(0xffffffffu, sum, 4);
      const int sumT = __shfl_sync(0xffffffffu, sum, 0);
      const int sumL = __shfl_sync(0xffffffffu, sum, 16);
      int dc;
      if (aT && aL) dc = (sumT + sumL + 8) >> 4;
      else if (aL) dc = (sumL + 4) >> 3;
      else if (aT) dc = (sumT + 4) >> 3;
      else dc = 128;
      if (lane == E8_DC) ev = dc;
    }
    const bool ok = mode <= 8 && ((legal_mask(aT, aL, aTL) >> mode) & 1u);
    int pr[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const uint32_t taps = (tw >> (16 * q)) & 0xffffu;
      const int e0 = __shfl_sync(0xffffffffu, ev, taps & 31);
      const int e1 = __shfl_sync(0xffffffffu, ev, (taps >> 5) & 31);
      const int e2 = __shfl_sync(0xffffffffu, ev, (taps >> 10) & 31);
      pr[q] = ok ? ((e0 + 2 * e1 + e2 + 2) >> 2) : 0;
    }
    const int o0 = clip255(pr[0] + lo16(rw)), o1 = clip255(pr[1] + hi16(rw));
    *reinterpret_cast<uint16_t*>(&lt[luma_at(bx + px, by + py)]) = (uint16_t)(o0 | (o1 << 8));
    __syncwarp();
  }
}

// Intra16x16 luma, pred16x16.rs:79-425 + pred16x16.rs:64-75. Lane = (row, half): 8 pixels.
__device__ __forceinline__ void predict_i16x16(uint8_t* lt, const int16_t* res_luma, int lane, int mode, bool availA,
                                               bool availB) {
  const int row = lane >> 1, h = lane & 1;
  const uint32_t t0 = *reinterpret_cast<const uint32_t*>(&lt[luma_at(8 * h, -1)]);
  const uint32_t t1 = *reinterpret_cast<const uint32_t*>(&lt[luma_at(8 * h + 4, -1)]);
  const int left = lt[luma_at(-1, row)];
  int pr[8];
  if (mode == 0) {  // vertical
#pragma unroll
    for (int k = 0; k < 4; k++) { pr[k] = (t0 >> (8 * k)) & 0xff; pr[4 + k] = (t1 >> (8 * k)) & 0xff; }
    if (!availB) {
#pragma unroll
      for (int k = 0; k < 8; k++) pr[k] = 0;
    }
  } else if (mode == 1) {  // horizontal
#pragma unroll
    for (int k = 0; k < 8; k++) pr[k] = availA ? left : 0;
  } else if (mode == 2) {  // DC
    int st = dp4a_us(t0, 0x01010101, 0);
    st = dp4a_us(t1, 0x01010101, st);
    st += __shfl_xor_sync(0xffffffffu, st, 1);
    int sl = left;
    sl += __shfl_xor_sync(0xffffffffu, sl, 2);
    sl += __shfl_xor_sync(0xffffffffu, sl, 4);
    sl += __shfl_xor_sync(0xffffffffu, sl, 8);
    sl += __shfl_xor_sync(0xffffffffu, sl, 16);
    int dc;
    if (availA && availB) dc = (st + sl + 16) >> 5;
    else if (availA) dc = (sl + 8) >> 4;
    else if (availB) dc = (st + 8) >> 4;
    else dc = 128;
#pragma unroll
    for (int k = 0; k < 8; k++) pr[k] = dc;
  } else {  // plane (needs A and B; the corner is read unchecked like pred16x16.rs:404, quirk Q5)
    const int corner = lt[luma_at(-1, -1)];
    // H = sum_{x'=0..7} (x'+1) * (p[8+x',-1] - p[6-x',-1]),  p[-1,-1] = corner
    int hp;
    if (h) hp = dp4a_us(t1, 0x08070605, dp4a_us(t0, 0x04030201, 0));
    else hp = dp4a_us(t1, 0x00ffFEFD, dp4a_us(t0, 0xFCFBFAF9, -8 * corner));  // -7..-4 | -3,-2,-1,0
    const int H = hp + __shfl_xor_sync(0xffffffffu, hp, 1);
    // V likewise over the left column: weight(row) = row - 7 (row 7 -> 0), corner weight -8
    int vp = (row - 7) * left;  // both halves reduce over their own 16 rows
    vp += __shfl_xor_sync(0xffffffffu, vp, 2);
    vp += __shfl_xor_sync(0xffffffffu, vp, 4);
    vp += __shfl_xor_sync(0xffffffffu, vp, 8);
    vp += __shfl_xor_sync(0xffffffffu, vp, 16);
    const int V = vp - 8 * corner;
    const int l15 = __shfl_sync(0xffffffffu, left, 30);
    const int t15 = __shfl_sync(0xffffffffu, (int)(t1 >> 24), 1);
    const int a = 16 * (l15 + t15);
    const int bb = (5 * H + 32) >> 6;
    const int cc = (5 * V + 32) >> 6;
    const bool ok = availA && availB;
    int base = a + bb * (8 * h - 7) + cc * (row - 7) + 16;
#pragma unroll
    for (int k = 0; k < 8; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
  }
  const uint4 rv = *reinterpret_cast<const uint4*>(&res_luma[row * 16 + 8 * h]);
  const int o0 = clip255(pr[0] + lo16(rv.x)), o1 = clip255(pr[1] + hi16(rv.x));
  const int o2 = clip255(pr[2] + lo16(rv.y)), o3 = clip255(pr[3] + hi16(rv.y));
  const int o4 = clip255(pr[4] + lo16(rv.z)), o5 = clip255(pr[5] + hi16(rv.z));
  const int o6 = clip255(pr[6] + lo16(rv.w)), o7 = clip255(pr[7] + hi16(rv.w));
  __syncwarp();  // every lane has read the neighbours it needs before the tile is overwritten
  *reinterpret_cast<uint2*>(&lt[luma_at(8 * h, row)]) = make_uint2(pack4(o0, o1, o2, o3), pack4(o4, o5, o6, o7));
  __syncwarp();
}

// Chroma Cb + Cr, trans_chroma.rs:96-366 + trans_chroma.rs:81-92. Lane = (plane, row, half): 4 pixels.
//   ct: two chroma tiles of kChromaTileBytes each; res_chroma: int16 [2][8][8]
__device__ __forceinline__ void predict_chroma(uint8_t* ct, const int16_t* res_chroma, int lane, int mode, bool availA,
                                               bool availB, bool availD) {
  const int pl = lane >> 4, row = (lane >> 1) & 7, h = lane & 1;
  uint8_t* tile = ct + pl * kChromaTileBytes;
  const uint32_t tw = *reinterpret_cast<const uint32_t*>(&tile[chroma_at(4 * h, -1)]);
  const int left = tile[chroma_at(-1, row)];
  int pr[4];
  if (mode == 0) {
    // DC per 4x4 chroma block with the reference's ">= 0" / "> 0" tests (quirk Q3)
    const int sumT = dp4a_us(tw, 0x01010101, 0);
    int sumL = left;
    sumL += __shfl_xor_sync(0xffffffffu, sumL, 2);
    sumL += __shfl_xor_sync(0xffffffffu, sumL, 4);  // over the 4 rows of this block row
    // "> 0" variants: unavailable or zero-valued
    const bool t_all_gt = availB && !((tw - 0x01010101u) & ~tw & 0x80808080u);
    const bool t3_gt = availB && (tw >> 24) != 0;
    const unsigned lz = __ballot_sync(0xffffffffu, left > 0);
    const int gbase = (lane & ~6) & ~1;                 // lane of row (row & 4), half 0 of this plane
    const unsigned grp = (lz >> gbase) & 0x55u;          // rows r0..r0+3 at bit 2*k
    const bool l_all_gt = availA && grp == 0x55u;
    const bool l3_gt = availA && ((grp >> 6) & 1u);
    const int by4 = row >> 2;
    int val;
    if (h == by4) {  // blocks 0 and 3, trans_chroma.rs:174-226
      if (availB && availA) val = (sumT + sumL + 4) >> 3;
      else if (!availB && availA) val = (sumL + 2) >> 2;
      else if (t_all_gt && !l_all_gt) val = (sumT + 2) >> 2;
      else val = 128;
    } else if (h == 1) {  // block 1 (x > 0, y == 0), trans_chroma.rs:227-252
      if (availB) val = (sumT + 2) >> 2;
      else if (l3_gt) val = (sumL + 2) >> 2;
      else val = 128;
    } else {  // block 2 (x == 0, y > 0), trans_chroma.rs:253-279
      if (l3_gt) val = (sumL + 2) >> 2;
      else if (t3_gt) val = (sumT + 2) >> 2;
      else val = 128;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = val;
  } else if (mode == 1) {
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = availA ? left : 0;
  } else if (mode == 2) {
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = availB ? (int)((tw >> (8 * k)) & 0xff) : 0;
  } else {  // plane, trans_chroma.rs:319-364 (needs top, left and the corner)
    const int corner = tile[chroma_at(-1, -1)];
    int hp = h ? dp4a_us(tw, 0x04030201, 0) : dp4a_us(tw, 0x00ffFEFD, -4 * corner);  // -3,-2,-1,0
    const int H = hp + __shfl_xor_sync(0xffffffffu, hp, 1);
    int vp = (row - 3) * left;  // both halves reduce over their own 8 rows
    vp += __shfl_xor_sync(0xffffffffu, vp, 2);
    vp += __shfl_xor_sync(0xffffffffu, vp, 4);
    vp += __shfl_xor_sync(0xffffffffu, vp, 8);
    const int V = vp - 4 * corner;
    const int l7 = __shfl_sync(0xffffffffu, left, (lane & 16) | 14);
    const int t7 = __shfl_sync(0xffffffffu, (int)(tw >> 24), (lane & 16) | 1);
    const int a = 16 * (l7 + t7);
    const int bb = (34 * H + 32) >> 6;
    const int cc = (34 * V + 32) >> 6;
    const bool ok = availA && availB && availD;
    const int base = a + bb * (4 * h - 3) + cc * (row - 3) + 16;
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
  }
  const uint2 rv = *reinterpret_cast<const uint2*>(&res_chroma[pl * 64 + row * 8 + 4 * h]);
  const int o0 = clip255(pr[0] + lo16(rv.x)), o1 = clip255(pr[1] + hi16(rv.x));
  const int o2 = clip255(pr[2] + lo16(rv.y)), o3 = clip255(pr[3] + hi16(rv.y));
  __syncwarp();
  *reinterpret_cast<uint32_t*>(&tile[chroma_at(4 * h, row)]) = pack4(o0, o1, o2, o3);
  __syncwarp();
}

// Prediction-mode derivation for Intra4x4 / Intra8x8 MBs, pred4x4.rs:363-427 / pred8x8.rs:698-764.
// Works on the 4x4 grid of 4x4 blocks (raster): lane g < 16 owns grid cell (g & 3, g >> 2). An Intra8x8
// MB stores each block's mode in the four cells it covers, which makes "A is Intra8x8 -> its 8x8 mode"
// and "A is Intra4x4 -> block 4*blk8+1" (and the B rules) plain cell look-ups.
//   syn      : this lane's prev/rem byte (pred_syntax entry of the block covering the cell)
//   a_col    : mode of the cell left of grid column 0 in this lane's grid row (from the previous MB; 2 if not NxN)
//   b_row    : mode of the cell above grid row 0 in this lane's grid column (from the row above; 2 if not NxN)
__device__ __forceinline__ int resolve_modes(int lane, int mbcls, int syn, int a_col, int b_row, bool availA,
                                             bool availB) {
  const int g = lane & 15, gx = g & 3, gy = g >> 2;
  int m = 2;
  if (mbcls == 2) return 2;  // warp-uniform
  const int step = mbcls == 1 ? 2 : 1;  // Intra8x8: cells move in 2x2 groups
  const int ox = gx & ~(step - 1), oy = gy & ~(step - 1);  // origin cell of the block covering this cell
  const bool haveA = ox > 0 || availA, haveB = oy > 0 || availB;
  // the block covering this cell takes A from its origin row and B from its origin column
  a_col = __shfl_sync(0xffffffffu, a_col, (oy * 4) | (lane & 16));
  b_row = __shfl_sync(0xffffffffu, b_row, ox | (lane & 16));
  const int prev = (syn >> 3) & 1, rem = syn & 7;
  const int ndiag = mbcls == 1 ? 3 : 7;
#pragma unroll 1
  for (int d = 0; d < ndiag; d++) {
    // neighbour cells: left of the block origin in the origin's row, above the origin in its column
    int a = __shfl_sync(0xffffffffu, m, (oy * 4 + max(ox - 1, 0)) | (lane & 16));
    int b = __shfl_sync(0xffffffffu, m, (max(oy - 1, 0) * 4 + ox) | (lane & 16));
    if (ox == 0) a = a_col;
    if (oy == 0) b = b_row;
    const int pred = (haveA && haveB) ? min(a, b) : 2;
    const int cand = prev ? pred : (rem < pred ? rem : rem + 1);
    m = (((ox + oy) >> (step - 1)) == d) ? cand : m;
  }
  return m;
}

}  // namespace dryv
