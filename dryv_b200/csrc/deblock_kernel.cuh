// deblock_kernel.cuh — optional in-loop deblocking post-pass (H.264 8.7) over reconstructed pictures, SURVEY.md §8(f) next-4.
// The reference has no deblocking filter (README.md:15; the slice-header fields are parsed, src/video/slice/header.rs:609-640),
// so this never runs on the dryv-parity path: it is what a consumer of streams with disable_deblocking_filter_idc != 1 calls
// after the reconstruction. Intra pictures only: bS = 4 on macroblock edges, 3 on the transform edges inside.
//
// The standard filters macroblocks in raster order, vertical edges then horizontal edges, each macroblock reading and
// changing up to three samples of its left and upper neighbours. That makes macroblock (x, y) depend on (x-1, y), (x, y-1)
// and (x+1, y-1) — the same x + 2y wavefront as intra prediction. One warp walks one macroblock row (rows dealt by an atomic
// ticket in row-major order over the pictures, so a row only ever waits on a lower ticket); a per-row progress counter in
// global memory says how many macroblocks of the row are finished. Inside a macroblock: lanes 0..15 take the sixteen luma
// lines, lanes 16..31 the eight lines of Cb and of Cr; the macroblock and its margins sit in a shared-memory tile.
#pragma once
#include <stdint.h>

namespace dryv {

struct DeblockArgs {
  uint8_t* yuv;             // pictures, in place
  const uint8_t* qp;        // per macroblock
  const uint8_t* t8x8;      // per macroblock
  int* progress;            // [n_frames * H] finished macroblocks per row, zeroed before the launch
  unsigned int* ticket;     // row ticket, zeroed before the launch
  int* status;              // sticky STATUS_*
  int W, H, n_frames;
  int cb_off, cr_off;       // chroma_qp_index_offset, second_chroma_qp_index_offset
  int off_a, off_b;         // FilterOffsetA / FilterOffsetB (2 * slice_*_offset_div2)
};

__constant__ uint8_t kDbAlpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13,
                                     15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144,
                                     162, 182, 203, 226, 255, 255};
__constant__ uint8_t kDbBeta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4,
                                    6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16,
                                    17, 17, 18, 18};
__constant__ uint8_t kDbTc0[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                   1, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25};  // bS = 3
__constant__ uint8_t kDbQpc[52] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25,
                                   26, 27, 28, 29, 29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38,
                                   39, 39, 39, 39};

__device__ __forceinline__ int db_clip(int v, int lo, int hi) { return min(max(v, lo), hi); }

// one line of samples across one edge, 8.7.2.3 / 8.7.2.4: s points at q0, `step` is the distance between samples across the edge
__device__ __forceinline__ void deblock_line(uint8_t* s, int step, bool strong, bool chroma, int qp_av, int off_a, int off_b) {
  const int ia = db_clip(qp_av + off_a, 0, 51), ib = db_clip(qp_av + off_b, 0, 51);
  const int alpha = kDbAlpha[ia], beta = kDbBeta[ib];
  const int p0 = s[-step], p1 = s[-2 * step], q0 = s[0], q1 = s[step];
  if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
  if (chroma) {
    if (strong) {
      s[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
      s[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
      const int tc = kDbTc0[ia] + 1;
      const int d = db_clip((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
      s[-step] = (uint8_t)db_clip(p0 + d, 0, 255);
      s[0] = (uint8_t)db_clip(q0 - d, 0, 255);
    }
    return;
  }
  const int p2 = s[-3 * step], q2 = s[2 * step];
  const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
  if (strong) {
    const bool small = abs(p0 - q0) < ((alpha >> 2) + 2);
    if (ap && small) {
      const int p3 = s[-4 * step];
      s[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      s[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
      s[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    } else {
      s[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    }
    if (aq && small) {
      const int q3 = s[3 * step];
      s[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      s[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
      s[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
    } else {
      s[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
  } else {
    const int tc0 = kDbTc0[ia];
    const int tc = tc0 + (ap ? 1 : 0) + (aq ? 1 : 0);
    const int d = db_clip((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
    s[-step] = (uint8_t)db_clip(p0 + d, 0, 255);
    s[0] = (uint8_t)db_clip(q0 - d, 0, 255);
    if (ap) s[-2 * step] = (uint8_t)(p1 + db_clip((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0, tc0));
    if (aq) s[step] = (uint8_t)(q1 + db_clip((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0, tc0));
  }
}

constexpr int kDbLumaStride = 20;    // 4 margin + 16
constexpr int kDbChromaStride = 12;  // 4 margin (2 used) + 8
constexpr int kDbWarps = 4;          // rows (warps) per CTA

struct DeblockTile {
  alignas(16) uint8_t luma[20 * kDbLumaStride];
  alignas(16) uint8_t chroma[2][12 * kDbChromaStride];
};

__global__ void __launch_bounds__(32 * kDbWarps, 8) deblock_wavefront_kernel(const DeblockArgs a) {
  __shared__ DeblockTile tiles[kDbWarps];
  __shared__ unsigned int s_row[kDbWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  DeblockTile& t = tiles[wid];
  const int W = a.W, H = a.H;
  const size_t lw = 16 * (size_t)W, cw = 8 * (size_t)W;
  const size_t luma_bytes = lw * 16 * H, frame_bytes = luma_bytes * 3 / 2;
  const unsigned total_rows = (unsigned)a.n_frames * (unsigned)H;
  for (;;) {
    if (lane == 0) s_row[wid] = atomicAdd(a.ticket, 1u);
    __syncwarp();
    const unsigned row = s_row[wid];
    __syncwarp();
    if (row >= total_rows) return;
    const int f = (int)(row / (unsigned)H), my = (int)(row % (unsigned)H);
    uint8_t* Y = a.yuv + (size_t)f * frame_bytes;
    uint8_t* C[2] = {Y + luma_bytes, Y + luma_bytes + luma_bytes / 4};
    const uint8_t* qp_row = a.qp + ((size_t)f * H + my) * W;
    const uint8_t* t8_row = a.t8x8 + ((size_t)f * H + my) * W;
    const volatile int* above = my > 0 ? a.progress + row - 1 : nullptr;
    bool dead = false;
    for (int mx = 0; mx < W; mx++) {
      // the row above must have finished the upper-right neighbour (its left edge changes samples this macroblock reads)
      if (above) {
        const int need = min(mx + 2, W);
        if (lane == 0 && !dead) {
          unsigned spins = 0;
          while (*above < need) {
            if (++spins > (1u << 24) || ((spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(a.status) == STATUS_WATCHDOG)) {
              atomicExch(a.status, STATUS_WATCHDOG);
              dead = true;
              break;
            }
          }
        }
        dead = __shfl_sync(0xffffffffu, dead ? 1 : 0, 0) != 0;
        __threadfence();
      }
      const int q = qp_row[mx];
      const int q_left = mx > 0 ? qp_row[mx - 1] : -1;
      const int q_up = my > 0 ? qp_row[mx - W] : -1;
      const int step = t8_row[mx] ? 8 : 4;
      // ---- load the macroblock with a 4-sample (chroma: 2 of 4) margin on the left and on top; L2 loads: the margins were
      // written by other SMs
      // (32-bit words of luma, 16-bit words of chroma: the margins start 4 / 2 samples left of the macroblock)
      for (int i = lane; i < 20 * 5; i += 32) {
        const int r = i / 5, wc = i % 5;
        const int gy = 16 * my + r - 4, gx = 16 * mx + 4 * wc - 4;
        *reinterpret_cast<uint32_t*>(t.luma + r * kDbLumaStride + 4 * wc) =
            (gy >= 0 && gx >= 0) ? __ldcg(reinterpret_cast<const uint32_t*>(Y + (size_t)gy * lw + gx)) : 0u;
      }
      for (int i = lane; i < 2 * 10 * 5; i += 32) {
        const int pl = i / 50, r = (i % 50) / 5, hc = i % 5;
        const int gy = 8 * my + r - 2, gx = 8 * mx + 2 * hc - 2;
        *reinterpret_cast<uint16_t*>(t.chroma[pl] + (r + 2) * kDbChromaStride + 2 * hc + 2) =
            (gy >= 0 && gx >= 0) ? __ldcg(reinterpret_cast<const uint16_t*>(C[pl] + (size_t)gy * cw + gx)) : (uint16_t)0;
      }
      __syncwarp();
      // ---- vertical edges, left to right: lane = line
      if (lane < 16) {
        uint8_t* line = t.luma + (4 + lane) * kDbLumaStride + 4;
        for (int e = 0; e < 16; e += step) {
          if (e == 0 && q_left < 0) continue;
          deblock_line(line + e, 1, e == 0, false, e == 0 ? (q + q_left + 1) >> 1 : q, a.off_a, a.off_b);
        }
      } else {
        const int pl = (lane - 16) >> 3, ln = lane & 7;
        const int off = pl ? a.cr_off : a.cb_off;
        const int qc = kDbQpc[db_clip(q + off, 0, 51)];
        uint8_t* line = t.chroma[pl] + (4 + ln) * kDbChromaStride + 4;
        for (int e = 0; e < 8; e += 4) {
          if (e == 0 && q_left < 0) continue;
          const int qa = e == 0 ? (qc + kDbQpc[db_clip(q_left + off, 0, 51)] + 1) >> 1 : qc;
          deblock_line(line + e, 1, e == 0, true, qa, a.off_a, a.off_b);
        }
      }
      __syncwarp();
      // ---- horizontal edges, top to bottom: lane = column
      if (lane < 16) {
        uint8_t* col = t.luma + 4 * kDbLumaStride + 4 + lane;
        for (int e = 0; e < 16; e += step) {
          if (e == 0 && q_up < 0) continue;
          deblock_line(col + e * kDbLumaStride, kDbLumaStride, e == 0, false, e == 0 ? (q + q_up + 1) >> 1 : q, a.off_a, a.off_b);
        }
      } else {
        const int pl = (lane - 16) >> 3, cn = lane & 7;
        const int off = pl ? a.cr_off : a.cb_off;
        const int qc = kDbQpc[db_clip(q + off, 0, 51)];
        uint8_t* col = t.chroma[pl] + 4 * kDbChromaStride + 4 + cn;
        for (int e = 0; e < 8; e += 4) {
          if (e == 0 && q_up < 0) continue;
          const int qa = e == 0 ? (qc + kDbQpc[db_clip(q_up + off, 0, 51)] + 1) >> 1 : qc;
          deblock_line(col + e * kDbChromaStride, kDbChromaStride, e == 0, true, qa, a.off_a, a.off_b);
        }
      }
      __syncwarp();
      // ---- write back what may have changed: the macroblock, three columns of the left neighbour (this macroblock's rows),
      // three rows of the upper neighbour (this macroblock's columns); chroma: one column / one row
      // (whole words: the outermost margin sample of a word is written back unchanged, and nobody else touches it before
      // this macroblock is published)
      for (int i = lane; i < 16 * 5; i += 32) {
        const int r = i / 5, wc = i % 5;
        if (wc == 0 && mx == 0) continue;
        *reinterpret_cast<uint32_t*>(Y + (size_t)(16 * my + r) * lw + 16 * mx + 4 * wc - 4) =
            *reinterpret_cast<const uint32_t*>(t.luma + (4 + r) * kDbLumaStride + 4 * wc);
      }
      if (my > 0 && lane < 12) {
        const int r = lane / 4 - 3, wc = lane % 4;
        *reinterpret_cast<uint32_t*>(Y + (size_t)(16 * my + r) * lw + 16 * mx + 4 * wc) =
            *reinterpret_cast<const uint32_t*>(t.luma + (4 + r) * kDbLumaStride + 4 + 4 * wc);
      }
      for (int i = lane; i < 2 * 8 * 5; i += 32) {
        const int pl = i / 40, r = (i % 40) / 5, hc = i % 5;
        if (hc == 0 && mx == 0) continue;
        *reinterpret_cast<uint16_t*>(C[pl] + (size_t)(8 * my + r) * cw + 8 * mx + 2 * hc - 2) =
            *reinterpret_cast<const uint16_t*>(t.chroma[pl] + (4 + r) * kDbChromaStride + 2 + 2 * hc);
      }
      if (my > 0 && lane < 8) {
        const int pl = lane >> 2, hc = lane & 3;
        *reinterpret_cast<uint16_t*>(C[pl] + (size_t)(8 * my - 1) * cw + 8 * mx + 2 * hc) =
            *reinterpret_cast<const uint16_t*>(t.chroma[pl] + 3 * kDbChromaStride + 4 + 2 * hc);
      }
      // ---- publish: every lane's stores are visible before the counter moves
      __threadfence();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile int*>(a.progress + row) = mx + 1;
    }
  }
}

}  // namespace dryv
