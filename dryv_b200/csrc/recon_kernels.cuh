// recon_kernels.cuh — sm_100a device code of the AVC intra reconstruction path.
//
// Three kernels (entry points in recon.cu):
//   resolve_modes_kernel   pixel-independent pre-pass: Intra4x4/8x8 prediction-mode derivation for a
//                          whole picture as a lock-step anti-diagonal walk over the 4x4-cell grid.
//   recon_wavefront_kernel persistent "row teams" (one CTA = two warps) walking macroblock rows in
//                          x+2y wavefront order:
//       front warp  128-bit loads of the MB's 768 B of int16 levels, inverse zig-zag as a register
//                   permutation, dequant, luma-DC / chroma-DC Hadamard (warp shuffles), 4x4 transform
//                   with one block per lane in registers (24 lanes = 16 luma + 8 chroma blocks) or the
//                   8x8 transform with 8 lanes per block; int16 residuals land in a ring slot. It has no
//                   dependency on any other macroblock and runs ahead of the pixels.
//       pixel warp  waits for the bottom line of the row above (64-bit payload|tag words, relaxed
//                   loads), predicts + adds the residual + clips into a shared-memory pixel tile, stores
//                   128-bit rows, publishes its own bottom line. Prediction is a gather from the tile:
//                   no shuffles, per-lane sample addresses come from host-built tap tables.
//   recon_residual_add_kernel  dequant + transform + add to a supplied prediction picture.
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
#include "residual_stage.cuh"

namespace dryv {

constexpr int kWarpsPerCta = 4;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;

constexpr int kLumaStride = kLumaTileStride;  // pixel (x, y) at (y + 1) * 48 + 16 + x, x in -4..31 (row -1), y in -1..15
constexpr int kLumaTileBytes = 17 * kLumaStride;
constexpr int kChromaStride = 24;        // pixel (x, y) at (y + 1) * 24 + 8 + x, x in -4..15 (row -1), y in -1..7
constexpr int kChromaTileBytes = 9 * kChromaStride;
constexpr int kScratchBytes = 1152;      // 8x8 coefficient slab (4 * 144 B) aliased with the transpose buffer (4 * 72 words)

// wavefront kernel: a "row team" = one CTA of two warps walking one macroblock row.
// The front warp hands each macroblock to the pixel warp through a ring of kSlots slots.
// Ring depth, measured (64 / 16 x 1080p, ms per step): 2 slots 0.900 / 0.504, 3: 0.917 / 0.504, 4: 0.903 / 0.498,
// 5: 0.924 / 0.505, 6 (9 teams per SM: one named barrier per slot) 0.926 / 0.500, 8 (7 teams) 0.979 / 0.477. Flat: the
// depth of the ring is not what lets the classes overlap. Again at eight teams per SM with three batches in flight
// (64 x 1080p, one batch at a time / overlapped): 3 slots 0.879 / 0.738, 4: 0.866 / 0.725, 5: 0.873 / 0.741, 6: 0.884 / 0.752.
#ifndef DRYV_SLOTS
#define DRYV_SLOTS 4
#endif
constexpr int kSlots = DRYV_SLOTS;
constexpr int kTeamThreads = 64;
struct Slot {
  alignas(16) int16_t res[256];  // luma residual [16][16]
  uint32_t modes_lo, modes_hi;   // resolved Intra4x4/8x8 modes in schedule order (resolve_modes_kernel)
  alignas(16) uint8_t rows[32];  // Intra4x4: tap rows, half-warp A's steps 0..9 then half-warp B's steps 2..7; rest 0
  int32_t frame, row, x;         // row < 0: no more work
  int32_t mbcls, mode16;         // 0/1/2 = Intra4x4/8x8/16x16; Intra16x16 prediction mode | intra_chroma_pred_mode << 8
  int32_t pad[1];
};
static_assert(sizeof(Slot) % 16 == 0, "slot alignment");
struct TeamSmem {
  DeviceTables tab;
  Slot slot[kSlots];
  alignas(16) int16_t coef[4][DRYV_COEFFS_PER_MB];  // level ring filled by cp.async (front warp)
  alignas(16) int16_t cres[128];                    // chroma residual [2][8][8] of the front warp's current MB
  alignas(16) uint8_t luma[kLumaTileBytes];         // luma pixel tile (pixel warp)
  alignas(16) uint8_t chroma[2 * kChromaTileBytes]; // chroma pixel tiles (front warp)
  alignas(16) uint8_t lcol[16];                     // right-most luma column of the previous MB, contiguous
  alignas(16) uint8_t ccol[16];                     // right-most Cb | Cr columns of the previous MB
  alignas(16) uint8_t e8[32];                       // filtered edge vector p' of the current Intra8x8 block
  alignas(16) uint8_t scratch[kScratchBytes];
  alignas(8) unsigned long long full[kSlots];       // mbarriers: slot filled by the front warp
  uint32_t pace;                                    // holds its own address: pacing chain of the poll loops
};

enum { STATUS_OK = 0, STATUS_UNSUPPORTED = 1, STATUS_WATCHDOG = 2 };

// Bottom line a macroblock hands to the row below: 4 luma words (16 px), 2 Cb, 2 Cr words (8 px each).
// Every 32-bit payload travels with the launch tag in one 64-bit word, so a single relaxed 64-bit load
// both fetches the data and proves it is there: no fence, no separate flag, no second round trip.
constexpr int kLineWords = 8;

// Resolved prediction modes of one macroblock (written by resolve_modes_kernel): one nibble per 4x4 block in the
// order the Intra4x4 schedule consumes them,
//   lo word: half-warp A, steps 0..7;  hi word: bits 0..7 half-warp A, steps 8, 9; bits 8..31 half-warp B, steps 2..7
// (an Intra8x8 macroblock stores each block's mode in the four cells it covers). Like the line words, each 32-bit
// half travels with the launch tag in one 64-bit word, so the wavefront kernel can consume the record while the
// pre-pass is still running (the two kernels overlap through programmatic dependent launch) without any fence.
constexpr int kModeWords = 2;

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
  unsigned long long* modes; // [n_frames * H * W][kModeWords]: payload | tag << 32
  uint32_t tag;             // launch tag: a line word is valid when its upper half equals it
  unsigned int* ticket;     // row ticket counter
  unsigned long long* prof; // stage clocks (development builds), may be null
  unsigned int* trace;      // per-macroblock timeline of picture 0 (development builds), may be null
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
// Prediction stage (pixel warp). Every neighbour sample is read from the shared-memory pixel tiles:
// the tile is the gather network, there are no shuffles on the pixel path.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int luma_at(int x, int y) { return (y + 1) * kLumaStride + 16 + x; }
__host__ __device__ constexpr int chroma_at(int x, int y) { return (y + 1) * kChromaStride + 8 + x; }

constexpr int kTap4Row = 16 * 4 * 4;          // bytes per row of DeviceTables::tap4
constexpr int kTap8Row = 32 * 8;              // bytes per mode row of DeviceTables::tap8

// legal-mode mask of the nine 4x4/8x8 modes given neighbour availability (the reference writes no
// prediction when the mode's neighbours are missing, so the prediction stays 0: SURVEY quirk Q4).
// Bit 0 (Vertical) doubles as "top available", bit 1 (Horizontal) as "left available".
__device__ __forceinline__ uint32_t legal_mask(bool t, bool l, bool c) {
  return 0x004u | (t ? 0x089u : 0u) | (l ? 0x102u : 0u) | ((t && l && c) ? 0x070u : 0u);
}

// Per-lane constants of the pixel warp.
struct PixLane {
  int half;                // Intra4x4: half-warp A (0) / B (1)
  int i4_pix;              // tile offset of this lane's pixel relative to its block origin
  int i4_res2;             // byte offset of this lane's residual relative to its block's
  int i4_tab;              // byte offset of this lane's entry inside a tap4 row
  int e8_s, e8_p, e8_n;    // Intra8x8 reference filter: tile offsets (relative to the block origin) of this
                           // lane's raw edge sample, its predecessor and its successor
  int i8_pix, i8_res2, i8_tab;
};

__device__ __forceinline__ PixLane make_pix_lane(int lane) {
  PixLane pl;
  const int half = lane >> 4, p = lane & 15, px = p & 3, py = p >> 2;
  pl.half = half;
  pl.i4_pix = py * kLumaStride + px;
  pl.i4_res2 = 2 * (py * 16 + px);
  pl.i4_tab = p * 16;
  // edge sample `lane` of an 8x8 block: 0..15 top, 16..23 left, 24 corner (pred8x8.rs:166-200)
  int s, pv, nx;
  if (lane < 16) {
    s = -kLumaStride + lane;
    pv = lane == 0 ? -kLumaStride - 1 : s - 1;
    nx = lane == 15 ? s : s + 1;
  } else if (lane < 24) {
    const int k = lane - 16;
    s = kLumaStride * k - 1;
    pv = k == 0 ? -kLumaStride - 1 : s - kLumaStride;
    nx = k == 7 ? s : s + kLumaStride;
  } else if (lane == 24) {
    s = -kLumaStride - 1;
    pv = -kLumaStride;  // p[0,-1]
    nx = -1;            // p[-1,0]
  } else {
    s = pv = nx = -kLumaStride - 1;
  }
  pl.e8_s = s;
  pl.e8_p = pv;
  pl.e8_n = nx;
  pl.i8_pix = (lane >> 2) * kLumaStride + (lane & 3) * 2;
  pl.i8_res2 = 2 * ((lane >> 2) * 16 + (lane & 3) * 2);
  pl.i8_tab = lane * 8;
  return pl;
}

// ---- Intra4x4 luma, pred4x4.rs:10-360 + transform.rs:98-110 -------------------------------------------
// Ten dependency steps, two blocks per step where the decode-order availability rules allow it
// (kI4BlkA / kI4BlkB); one pixel per lane, 16 lanes per block. A rolled loop: the kernel is instruction-fetch
// bound when a dozen teams share an SM (an unrolled version with immediates was measured: faster on Intra4x4-only
// pictures, slower on the mix), so code size matters more than the handful of instructions an unrolled
// version saves. Everything that depends on the mode or on availability was folded into the tap row chosen by the front
// warp (i4_tap_row), the schedule entries and tap records are 32-bit fields used as loaded, and a half-warp without a
// block in a step works on a dummy block in the tile's padding: a step is loads, three adds and a clamp.
//   rows: this half-warp's ten tap-row bytes (Slot::rows + 0 or 8)
struct I4Regs {
  uint4 tap;      // three sample offsets (biased), kind
  uint32_t org;   // tile offset of the block origin
  int r;          // residual
};
__device__ __forceinline__ void i4_fetch(I4Regs& q, const I4Step* e, const uint8_t* tap4, int row, const uint8_t* resp) {
  q.tap = *reinterpret_cast<const uint4*>(tap4 + row * kTap4Row);
  const uint2 st = *reinterpret_cast<const uint2*>(e);
  q.org = st.x;
  q.r = *reinterpret_cast<const int16_t*>(resp + st.y);
}
__device__ __forceinline__ void predict_i4x4(const DeviceTables& tab, uint8_t* lt, const int16_t* res_luma,
                                             const PixLane& pl, const uint8_t* rows) {
  const I4Step* st = &tab.i4tab[0][pl.half];
  const uint8_t* tap4 = reinterpret_cast<const uint8_t*>(&tab.tap4[0][0][0]) + pl.i4_tab;
  const uint8_t* resp = reinterpret_cast<const uint8_t*>(res_luma) + pl.i4_res2;
  const uint8_t* ltb = lt - kTap4Bias;
  // software pipeline: the table look-ups of step s+1 (which do not depend on any pixel) are issued under the latency
  // of the pixel loads of step s, so a step's dependent chain is pixel load -> three adds -> clamp -> store
  I4Regs cur;
  i4_fetch(cur, st, tap4, rows[0], resp);
#pragma unroll 1
  for (int s = 0; s < 10; s++) {
    const uint8_t* ob = ltb + cur.org;
    const int e0 = ob[cur.tap.x], e1 = ob[cur.tap.y], e2 = ob[cur.tap.z];
    int kind = (int)cur.tap.w;
    const int r = cur.r;
    uint8_t* dst = lt + cur.org + pl.i4_pix;
    i4_fetch(cur, st + 2 * (s + 1), tap4, rows[s + 1], resp);  // step 10 = step 9 again (table and Slot::rows padding)
    int pred = (e0 + 2 * e1 + e2 + 2) >> 2;
    if (kind >= kI4KindDc) {  // DC, pred4x4.rs:116-167
      const uint8_t* eb = dst - pl.i4_pix;
      const int sT = dp4a_us(*reinterpret_cast<const uint32_t*>(eb - kLumaStride), 0x01010101, 0);
      const int sL = eb[-1] + eb[kLumaStride - 1] + eb[2 * kLumaStride - 1] + eb[3 * kLumaStride - 1];
      pred = kind == kI4KindDc ? ((sT + sL + 4) >> 3)
                               : (kind == kI4KindDcTop ? ((sT + 2) >> 2) : (kind == kI4KindDcLeft ? ((sL + 2) >> 2) : 128));
      kind = 1;
    }
    *dst = (uint8_t)clip255(pred * kind + r);
    __syncwarp();
  }
}

// Front-warp side of predict_i4x4: lane k < 16 turns the mode of the block behind Slot::rows[k] into its tap
// row. modes_lo / modes_hi: see kModeWords; av = A | B<<1 | C<<2 | D<<3.
__device__ __forceinline__ int i4_tap_row(const DeviceTables& tab, int k, uint32_t modes_lo, uint32_t modes_hi, int av) {
  const uint32_t m = ((k < 8 ? modes_lo : modes_hi) >> (4 * (k & 7))) & 15u;
  const uint32_t info = tab.i4row[av][k];
  if (!((info >> m) & 1u)) return kI4RowIllegal;
  if (m == 2) return (info & 1u) ? ((info & 2u) ? 2 : kI4RowDcTop) : ((info & 2u) ? kI4RowDcLeft : kI4RowDcNone);
  return (int)(m + 9u * ((info >> 9) & 1u));
}

// ---- Intra8x8 luma, pred8x8.rs:152-696 + pred8x8.rs:34-46 ---------------------------------------------
// Four sequential blocks (a rolled loop, see predict_i4x4). Phase 1: lanes 0..24 filter one reference sample
// each (pred8x8.rs:222-288, with the x = 0 overwrite of quirk Q2) into e8[]. Phase 2: two pixels per lane
// gathered from e8[].
__device__ __forceinline__ void predict_i8x8(const DeviceTables& tab, uint8_t* lt, uint8_t* e8,
                                             const int16_t* res_luma, const PixLane& pl, int lane, uint32_t modes_lo,
                                             uint32_t modes_hi, bool A, bool B, bool C, bool D) {
  const uint8_t* resp8 = reinterpret_cast<const uint8_t*>(res_luma) + pl.i8_res2;
  constexpr int t7 = -kLumaStride + 7;
  // availability of (top, left, corner, top-right), 4 bits per block
  //   block 0: B A D B   block 1: B 1 B C   block 2: 1 A A 1   block 3: 1 1 1 0
  const uint32_t avw = (B ? 0x0059u : 0u) | (A ? 0x0602u : 0u) | (D ? 0x0004u : 0u) | (C ? 0x0080u : 0u) | 0x7920u;
#pragma unroll 1
  for (int blk = 0; blk < 4; blk++) {
    const int o8 = luma_at((blk & 1) * 8, (blk >> 1) * 8);
    const uint32_t fl = avw >> (4 * blk);
    const bool aT = fl & 1u, aL = fl & 2u, aTL = fl & 4u, aTR = fl & 8u;
    // block modes: cells (0,0), (2,0), (0,2), (2,2) = A step 0, B step 2, A step 4, B step 6
    const int mode = (int)((((blk & 1) ? modes_hi : modes_lo) >> (8 * blk)) & 15u);
    int oS = pl.e8_s, oP = pl.e8_p, oN = pl.e8_n;
    if (!aTR) {  // p[8..15,-1] unavailable: replicate p[7,-1] (pred8x8.rs:202-220)
      if (lane >= 8 && lane < 16) oS = oP = oN = t7;
      if (lane == 7) oN = t7;
    }
    if (!aTL && lane == 16) oP = oS;
    if (lane == 24) {
      if (!aT) oP = oS;
      if (!aL) oN = oS;
    }
    const uint8_t* b = lt + o8;
    const int raw = b[oS], nv = b[oN];
    int pv = b[oP];
    if (!aTL && lane == 0) pv = -1;  // Q2: the raw p[-1,-1] sentinel enters the filter
    e8[lane] = (uint8_t)((pv + 2 * raw + nv + 2) >> 2);
    const uint8_t* tp = &tab.tap8[0][0][0] + mode * kTap8Row + pl.i8_tab;
    const uint32_t i0 = tp[0], i1 = tp[1], i2 = tp[2], i3 = tp[3], i4 = tp[4], i5 = tp[5];
    const uint32_t rw = *reinterpret_cast<const uint32_t*>(resp8 + 2 * (((blk >> 1) * 8) * 16 + (blk & 1) * 8));
    __syncwarp();
    int pr0, pr1;
    if (mode == 2) {  // DC, pred8x8.rs:326-425 (warp-uniform): sums of the filtered top 0..7 and left 0..7
      const uint32_t* ew = reinterpret_cast<const uint32_t*>(e8);
      const int sT = dp4a_us(ew[1], 0x01010101, dp4a_us(ew[0], 0x01010101, 0));
      const int sL = dp4a_us(ew[5], 0x01010101, dp4a_us(ew[4], 0x01010101, 0));
      pr0 = (aT && aL) ? ((sT + sL + 8) >> 4) : (aL ? ((sL + 4) >> 3) : (aT ? ((sT + 4) >> 3) : 128));
      pr1 = pr0;
    } else {
      pr0 = ((int)e8[i0] + 2 * (int)e8[i1] + (int)e8[i2] + 2) >> 2;
      pr1 = ((int)e8[i3] + 2 * (int)e8[i4] + (int)e8[i5] + 2) >> 2;
    }
    if (!((legal_mask(aT, aL, aTL) >> mode) & 1u)) pr0 = pr1 = 0;
    const int o0 = clip255(pr0 + lo16(rw)), o1 = clip255(pr1 + hi16(rw));
    *reinterpret_cast<uint16_t*>(&lt[o8 + pl.i8_pix]) = (uint16_t)(o0 | (o1 << 8));
    __syncwarp();
  }
}

// ---- Intra16x16 luma, pred16x16.rs:79-425 + pred16x16.rs:64-75. Lane = (row, half): 8 pixels ----------
//   lcol: the left neighbour column as 16 contiguous bytes.
__device__ __forceinline__ void predict_i16x16(uint8_t* lt, const uint8_t* lcol, const int16_t* res_luma, int lane,
                                               int mode, bool availA, bool availB) {
  const int row = lane >> 1, h = lane & 1;
  const uint4 tv = *reinterpret_cast<const uint4*>(&lt[luma_at(0, -1)]);
  const uint4 rv = *reinterpret_cast<const uint4*>(&res_luma[row * 16 + 8 * h]);
  int pr[8];
  if (mode == 0) {  // vertical
    const uint32_t t0 = h ? tv.z : tv.x, t1 = h ? tv.w : tv.y;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      pr[k] = availB ? (int)((t0 >> (8 * k)) & 0xff) : 0;
      pr[4 + k] = availB ? (int)((t1 >> (8 * k)) & 0xff) : 0;
    }
  } else if (mode == 1) {  // horizontal
    const int left = lcol[row];
#pragma unroll
    for (int k = 0; k < 8; k++) pr[k] = availA ? left : 0;
  } else {
    const uint4 lv = *reinterpret_cast<const uint4*>(lcol);
    if (mode == 2) {  // DC
      const int st = dp4a_us(tv.w, 0x01010101, dp4a_us(tv.z, 0x01010101, dp4a_us(tv.y, 0x01010101, dp4a_us(tv.x, 0x01010101, 0))));
      const int sl = dp4a_us(lv.w, 0x01010101, dp4a_us(lv.z, 0x01010101, dp4a_us(lv.y, 0x01010101, dp4a_us(lv.x, 0x01010101, 0))));
      int dc;
      if (availA && availB) dc = (st + sl + 16) >> 5;
      else if (availA) dc = (sl + 8) >> 4;
      else if (availB) dc = (st + 8) >> 4;
      else dc = 128;
#pragma unroll
      for (int k = 0; k < 8; k++) pr[k] = dc;
    } else {  // plane (needs A and B; the corner is read unchecked like pred16x16.rs:404, quirk Q5)
      const int corner = lt[luma_at(-1, -1)];
      // H = sum_{x'=0..7} (x'+1) * (p[8+x',-1] - p[6-x',-1]),  p[-1,-1] = corner ; V likewise over the left column
      const int H = dp4a_us(tv.w, 0x08070605, dp4a_us(tv.z, 0x04030201, dp4a_us(tv.y, 0x00ffFEFD, dp4a_us(tv.x, 0xFCFBFAF9, -8 * corner))));
      const int V = dp4a_us(lv.w, 0x08070605, dp4a_us(lv.z, 0x04030201, dp4a_us(lv.y, 0x00ffFEFD, dp4a_us(lv.x, 0xFCFBFAF9, -8 * corner))));
      const int a = 16 * ((int)(lv.w >> 24) + (int)(tv.w >> 24));
      const int bb = (5 * H + 32) >> 6;
      const int cc = (5 * V + 32) >> 6;
      const bool ok = availA && availB;
      const int base = a + bb * (8 * h - 7) + cc * (row - 7) + 16;
#pragma unroll
      for (int k = 0; k < 8; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
    }
  }
  const int o0 = clip255(pr[0] + lo16(rv.x)), o1 = clip255(pr[1] + hi16(rv.x));
  const int o2 = clip255(pr[2] + lo16(rv.y)), o3 = clip255(pr[3] + hi16(rv.y));
  const int o4 = clip255(pr[4] + lo16(rv.z)), o5 = clip255(pr[5] + hi16(rv.z));
  const int o6 = clip255(pr[6] + lo16(rv.w)), o7 = clip255(pr[7] + hi16(rv.w));
  // rows 0..15 / columns 0..15 are written, row -1 and the left column vector are read: no hazard
  *reinterpret_cast<uint2*>(&lt[luma_at(8 * h, row)]) = make_uint2(pack4(o0, o1, o2, o3), pack4(o4, o5, o6, o7));
  __syncwarp();
}

// ---- Chroma Cb + Cr, trans_chroma.rs:96-366 + trans_chroma.rs:81-92. Lane = (plane, row, half): 4 px ---
//   ct: two chroma tiles of kChromaTileBytes each; ccol: left neighbour columns, 8 bytes per plane;
//   res_chroma: int16 [2][8][8]
__device__ __forceinline__ void predict_chroma(uint8_t* ct, const uint8_t* ccol, const int16_t* res_chroma, int lane,
                                               int mode, bool availA, bool availB, bool availD) {
  const int pl = lane >> 4, row = (lane >> 1) & 7, h = lane & 1;
  uint8_t* tile = ct + pl * kChromaTileBytes;
  const uint8_t* cv = ccol + pl * 8;
  const uint32_t tw = *reinterpret_cast<const uint32_t*>(&tile[chroma_at(4 * h, -1)]);
  const uint2 rv = *reinterpret_cast<const uint2*>(&res_chroma[pl * 64 + row * 8 + 4 * h]);
  int pr[4];
  if (mode == 0) {
    // DC per 4x4 chroma block with the reference's ">= 0" / "> 0" tests (quirk Q3)
    const int by4 = row >> 2;
    const uint32_t lw = *reinterpret_cast<const uint32_t*>(cv + 4 * by4);  // left samples of this block row
    const int sumT = dp4a_us(tw, 0x01010101, 0), sumL = dp4a_us(lw, 0x01010101, 0);
    // "> 0" variants: unavailable or zero-valued
    const bool t_all_gt = availB && !((tw - 0x01010101u) & ~tw & 0x80808080u);
    const bool t3_gt = availB && (tw >> 24) != 0;
    const bool l_all_gt = availA && !((lw - 0x01010101u) & ~lw & 0x80808080u);
    const bool l3_gt = availA && (lw >> 24) != 0;
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
    const int left = cv[row];
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = availA ? left : 0;
  } else if (mode == 2) {
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = availB ? (int)((tw >> (8 * k)) & 0xff) : 0;
  } else {  // plane, trans_chroma.rs:319-364 (needs top, left and the corner)
    const int corner = tile[chroma_at(-1, -1)];
    const uint2 tv = *reinterpret_cast<const uint2*>(&tile[chroma_at(0, -1)]);
    const uint2 lv = *reinterpret_cast<const uint2*>(cv);
    const int H = dp4a_us(tv.y, 0x04030201, dp4a_us(tv.x, 0x00ffFEFD, -4 * corner));
    const int V = dp4a_us(lv.y, 0x04030201, dp4a_us(lv.x, 0x00ffFEFD, -4 * corner));
    const int a = 16 * ((int)(lv.y >> 24) + (int)(tv.y >> 24));
    const int bb = (34 * H + 32) >> 6;
    const int cc = (34 * V + 32) >> 6;
    const bool ok = availA && availB && availD;
    const int base = a + bb * (4 * h - 3) + cc * (row - 3) + 16;
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
  }
  const int o0 = clip255(pr[0] + lo16(rv.x)), o1 = clip255(pr[1] + hi16(rv.x));
  const int o2 = clip255(pr[2] + lo16(rv.y)), o3 = clip255(pr[3] + hi16(rv.y));
  *reinterpret_cast<uint32_t*>(&tile[chroma_at(4 * h, row)]) = pack4(o0, o1, o2, o3);
  __syncwarp();
}

}  // namespace dryv
