// deblock_kernel.cuh — optional in-loop deblocking post-pass (H.264 8.7) over reconstructed pictures, SURVEY.md §8(f) next-4.
// The reference has no deblocking filter (README.md:15; the slice-header fields are parsed, src/video/slice/header.rs:609-640),
// so this never runs on the dryv-parity path: it is what a consumer of streams with disable_deblocking_filter_idc != 1 calls
// after the reconstruction. Intra pictures only: bS = 4 on macroblock edges, 3 on the transform edges inside.
//
// The standard filters macroblocks in raster order, vertical edges then horizontal edges, each macroblock reading and
// changing up to three samples of its left and upper neighbours. That makes macroblock (x, y) depend on (x-1, y), (x, y-1)
// and (x+1, y-1) — the same x + 2y wavefront as intra prediction. One warp walks one macroblock row (rows dealt by an atomic
// ticket in row-major order over the pictures, so a row only ever waits on a lower ticket). Data flow of a row:
//   * its own macroblocks come from global memory exactly as the reconstruction left them, one macroblock ahead of use;
//   * the right four columns of a macroblock stay in the shared-memory tile as the left margin of the next one;
//   * the bottom four luma rows and two chroma rows of a macroblock go DOWN to the next row as 24 tagged 64-bit words
//     (payload | launch tag << 32), published once the macroblock on the right has filtered its left edge — the words
//     carry the samples themselves, so there is no counter, no fence and no second read of global memory;
//   * the lower row writes the upper macroblock's last three luma rows / last chroma row after filtering across the edge;
//     the upper row never writes them (the last picture row writes its own), so every byte has one writer per phase and
//     the two writes a row makes to the same word (macroblock, then its right columns one macroblock later) come from the
//     same thread in program order.
// Inside a macroblock: lanes 0..15 take the sixteen luma lines, lanes 16..31 the eight lines of Cb and of Cr.
#pragma once
#include <stdint.h>

namespace dryv {

constexpr int kDbLineWords = 24;  // 4 luma rows x 4 words, then Cb and Cr: 2 rows x 2 words each

struct DeblockArgs {
  uint8_t* yuv;             // pictures, in place
  const uint8_t* qp;        // per macroblock
  const uint8_t* t8x8;      // per macroblock
  unsigned long long* line; // [n_frames * H * W][kDbLineWords] tagged bottom rows handed to the row below
  unsigned int* ticket;     // row ticket, zeroed before the launch
  uint32_t tag;             // this launch's tag (never 0; the buffer starts zeroed)
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
  uint32_t pend[kDbLineWords];  // bottom rows of the previous macroblock, waiting for this one's left-edge filter
};

__device__ __forceinline__ uint32_t& db_word(uint8_t* p) { return *reinterpret_cast<uint32_t*>(p); }

__global__ void __launch_bounds__(32 * kDbWarps, 8) deblock_wavefront_kernel(const DeblockArgs a) {
  __shared__ DeblockTile tiles[kDbWarps];
  __shared__ unsigned int s_row[kDbWarps];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  DeblockTile& t = tiles[wid];
  const int W = a.W, H = a.H;
  const size_t lw = 16 * (size_t)W, cw = 8 * (size_t)W;
  const size_t luma_bytes = lw * 16 * H, frame_bytes = luma_bytes * 3 / 2;
  const unsigned total_rows = (unsigned)a.n_frames * (unsigned)H;
  const unsigned long long tag64 = (unsigned long long)a.tag << 32;
  // lane roles: luma line `lane` (lanes 0..15); chroma plane cpl, line cln (lanes 16..31)
  const bool is_luma = lane < 16;
  const int cpl = (lane >> 3) & 1, cln = lane & 7;
  // line-word roles (lanes 0..23): luma row lane>>2 word lane&3; chroma plane wpl row wr word wh
  const int wk = lane - 16, wpl = (wk >> 2) & 1, wr = (wk >> 1) & 1, wh = wk & 1;
  for (;;) {
    if (lane == 0) s_row[wid] = atomicAdd(a.ticket, 1u);
    __syncwarp();
    const unsigned row = s_row[wid];
    __syncwarp();
    if (row >= total_rows) return;
    const int f = (int)(row / (unsigned)H), my = (int)(row % (unsigned)H);
    const bool last_row = my == H - 1;
    uint8_t* Y = a.yuv + (size_t)f * frame_bytes;
    uint8_t* Cp = Y + luma_bytes + (cpl ? luma_bytes / 4 : 0);
    uint8_t* own = is_luma ? Y + (size_t)(16 * my + lane) * lw : Cp + (size_t)(8 * my + cln) * cw;  // this lane's line
    uint8_t* own_tile = is_luma ? t.luma + (4 + lane) * kDbLumaStride : t.chroma[cpl] + (4 + cln) * kDbChromaStride;
    const bool own_stored = last_row || (is_luma ? lane < 13 : cln < 7);
    const uint8_t* qp_row = a.qp + (size_t)row * W;
    const uint8_t* t8_row = a.t8x8 + (size_t)row * W;
    unsigned long long* line_mine = a.line + (size_t)row * W * kDbLineWords + lane;
    const unsigned long long* line_above = line_mine - (size_t)W * kDbLineWords;
    // where this lane's line word lives in the tile: top margin (read side) and bottom rows (write side)
    uint8_t* top_dst = is_luma ? t.luma + (lane >> 2) * kDbLumaStride + 4 + 4 * (lane & 3)
                               : t.chroma[wpl] + (2 + wr) * kDbChromaStride + 4 + 4 * wh;
    uint8_t* bot_src = is_luma ? t.luma + (16 + (lane >> 2)) * kDbLumaStride + 4 + 4 * (lane & 3)
                               : t.chroma[wpl] + (10 + wr) * kDbChromaStride + 4 + 4 * wh;
    // the same rows' left-margin word: the previous macroblock's last word after this macroblock's left-edge filter
    uint8_t* bot_left = is_luma ? t.luma + (16 + (lane >> 2)) * kDbLumaStride : t.chroma[wpl] + (10 + wr) * kDbChromaStride;
    const bool patched = is_luma ? (lane & 3) == 3 : wh == 1;
    bool dead = false;
    uint4 cy = make_uint4(0, 0, 0, 0);
    int nq = 0, nt8 = 0, nqup = -1;
    unsigned long long tv = 0;
    auto fetch = [&](int mx) {  // macroblock mx of this row: samples, QPs, and a first look at the words from above
      if (is_luma) cy = __ldcg(reinterpret_cast<const uint4*>(own + 16 * mx));
      else {
        const uint2 c = __ldcg(reinterpret_cast<const uint2*>(own + 8 * mx));
        cy.x = c.x;
        cy.y = c.y;
      }
      nq = qp_row[mx];
      nt8 = t8_row[mx];
      if (my > 0) {
        nqup = qp_row[mx - W];
        if (lane < kDbLineWords) tv = ld_relaxed_gpu_u64(line_above + (size_t)mx * kDbLineWords);
      }
    };
    fetch(0);
    int q_left = -1;
    for (int mx = 0; mx < W; mx++) {
      const int q = nq, q_up = nqup, step = nt8 ? 8 : 4;
      // ---- this macroblock into the tile; the words from above into the top margin
      db_word(own_tile + 4) = cy.x;
      db_word(own_tile + 8) = cy.y;
      if (is_luma) {
        db_word(own_tile + 12) = cy.z;
        db_word(own_tile + 16) = cy.w;
      }
      if (my > 0 && lane < kDbLineWords) {
        const unsigned long long* p = line_above + (size_t)mx * kDbLineWords;
        unsigned spins = 0;
        while ((uint32_t)(tv >> 32) != a.tag) {
          if (++spins > (1u << 22) || ((spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(a.status) == STATUS_WATCHDOG)) {
            atomicExch(a.status, STATUS_WATCHDOG);
            dead = true;
            break;
          }
          tv = ld_relaxed_gpu_u64(p);
        }
        db_word(top_dst) = (uint32_t)tv;
      }
      if (mx + 1 < W) fetch(mx + 1);
      __syncwarp();
      // ---- vertical edges, left to right: lane = line
      if (is_luma) {
        uint8_t* ln = own_tile + 4;
        for (int e = 0; e < 16; e += step) {
          if (e == 0 && q_left < 0) continue;
          deblock_line(ln + e, 1, e == 0, false, e == 0 ? (q + q_left + 1) >> 1 : q, a.off_a, a.off_b);
        }
      } else {
        const int off = cpl ? a.cr_off : a.cb_off;
        const int qc = kDbQpc[db_clip(q + off, 0, 51)];
        uint8_t* ln = own_tile + 4;
        for (int e = 0; e < 8; e += 4) {
          if (e == 0 && q_left < 0) continue;
          const int qa = e == 0 ? (qc + kDbQpc[db_clip(q_left + off, 0, 51)] + 1) >> 1 : qc;
          deblock_line(ln + e, 1, e == 0, true, qa, a.off_a, a.off_b);
        }
      }
      __syncwarp();
      // ---- the previous macroblock is finished down to its last rows: hand those to the row below
      if (mx > 0 && !last_row && lane < kDbLineWords)
        st_relaxed_gpu_u64(line_mine + (size_t)(mx - 1) * kDbLineWords, tag64 | (patched ? db_word(bot_left) : t.pend[lane]));
      // ---- horizontal edges, top to bottom: lane = column
      if (is_luma) {
        uint8_t* col = t.luma + 4 * kDbLumaStride + 4 + lane;
        for (int e = 0; e < 16; e += step) {
          if (e == 0 && q_up < 0) continue;
          deblock_line(col + e * kDbLumaStride, kDbLumaStride, e == 0, false, e == 0 ? (q + q_up + 1) >> 1 : q, a.off_a, a.off_b);
        }
      } else {
        const int off = cpl ? a.cr_off : a.cb_off;
        const int qc = kDbQpc[db_clip(q + off, 0, 51)];
        uint8_t* col = t.chroma[cpl] + 4 * kDbChromaStride + 4 + cln;
        for (int e = 0; e < 8; e += 4) {
          if (e == 0 && q_up < 0) continue;
          const int qa = e == 0 ? (qc + kDbQpc[db_clip(q_up + off, 0, 51)] + 1) >> 1 : qc;
          deblock_line(col + e * kDbChromaStride, kDbChromaStride, e == 0, true, qa, a.off_a, a.off_b);
        }
      }
      __syncwarp();
      // ---- stores: this lane's line of the macroblock (not the rows the next picture row still filters), the left
      // neighbour's last word of the same line, and the upper neighbour's last rows
      if (own_stored) {
        if (is_luma)
          *reinterpret_cast<uint4*>(own + 16 * mx) =
              make_uint4(db_word(own_tile + 4), db_word(own_tile + 8), db_word(own_tile + 12), db_word(own_tile + 16));
        else
          *reinterpret_cast<uint2*>(own + 8 * mx) = make_uint2(db_word(own_tile + 4), db_word(own_tile + 8));
        if (mx > 0) db_word(own + (is_luma ? 16 : 8) * mx - 4) = db_word(own_tile);
      }
      if (my > 0 && lane < 16) {
        if (lane < 12)
          db_word(Y + (size_t)(16 * my - 3 + (lane >> 2)) * lw + 16 * mx + 4 * (lane & 3)) =
              db_word(t.luma + (1 + (lane >> 2)) * kDbLumaStride + 4 + 4 * (lane & 3));
        else {
          const int pl = (lane >> 1) & 1, h = lane & 1;
          db_word(Y + luma_bytes + (pl ? luma_bytes / 4 : 0) + (size_t)(8 * my - 1) * cw + 8 * mx + 4 * h) =
              db_word(t.chroma[pl] + 3 * kDbChromaStride + 4 + 4 * h);
        }
      }
      // ---- keep the last rows for the hand-off, and the right columns as the next macroblock's left margin
      if (lane < kDbLineWords) t.pend[lane] = db_word(bot_src);
      db_word(own_tile) = db_word(own_tile + (is_luma ? 16 : 8));
      q_left = q;
      __syncwarp();
    }
    if (!last_row && lane < kDbLineWords) st_relaxed_gpu_u64(line_mine + (size_t)(W - 1) * kDbLineWords, tag64 | t.pend[lane]);
    if (__any_sync(0xffffffffu, dead)) return;
  }
}

}  // namespace dryv
