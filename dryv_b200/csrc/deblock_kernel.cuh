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
// A warp walks the same macroblock row of TWO pictures, one per half-warp (the pictures are independent, the control flow
// is identical): the kernel is bound by instruction issue, and the packed filters of deblock_packed.cuh need only sixteen
// lanes per macroblock — eight for luma, four for Cb, four for Cr, two lines (vertical edges) or two columns (horizontal
// edges) per lane in the two 16-bit fields of a register. Per macroblock: the two lines of a lane are unpacked from the
// row words (PRMT), the four vertical edges filtered in registers, the lines written to the shared-memory tile; then every
// lane reads its two columns of the tile (rows -4..15), filters the four horizontal edges and writes them back.
#pragma once
#include <stdint.h>

#include "deblock_packed.cuh"

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

__device__ __forceinline__ int db_clip(int v, int lo, int hi) { return min(max(v, lo), hi); }

__constant__ uint8_t kDbAlpha[52] = DRYV_DB_ALPHA;
__constant__ uint8_t kDbBeta[52] = DRYV_DB_BETA;
__constant__ uint8_t kDbTc0[52] = DRYV_DB_TC0_BS3;
__constant__ uint8_t kDbQpc[52] = DRYV_DB_QPC;

constexpr int kDbWarps = 4;  // warps per CTA

// edge constants by table index, and the QP each lane kind filters with (luma: QPY; Cb / Cr: QPC of QPY + the offset)
struct DbTables {
  uint4 a[52];     // by indexA: ka, ks, tcb (luma), lo1 — see EdgeConst
  uint32_t b[52];  // by indexB: kb
  uint8_t qmap[3][52];
};

// one picture's macroblock and its margins: every row 16 bytes apart, rows -4..-1 are the upper neighbour's last rows
struct DbHalf {
  alignas(16) uint8_t luma[20 * 16];
  alignas(16) uint8_t chroma[2][12 * 16];  // 8 bytes of a row used; rows -4, -3 unused
  uint32_t patch[8];                       // the left neighbour's last word of its bottom rows after this macroblock's left edge
  uint32_t pad[24];                        // the two halves of a warp 64 bytes apart modulo 128: their accesses share no bank
};
static_assert(sizeof(DbHalf) % 128 == 64, "half tiles must not alias in the banks");

__device__ __forceinline__ uint32_t& db_word(uint8_t* p) { return *reinterpret_cast<uint32_t*>(p); }

// two row words (4 samples of line A, of line B) -> four packed registers
__device__ __forceinline__ void db_unpack4(uint32_t wa, uint32_t wb, uint32_t* r) {
  const uint32_t lo = prmt(wa, wb, 0x6240), hi = prmt(wa, wb, 0x7351);
  r[0] = prmt(lo, 0, 0x4140);
  r[1] = prmt(hi, 0, 0x4140);
  r[2] = prmt(lo, 0, 0x4342);
  r[3] = prmt(hi, 0, 0x4342);
}
__device__ __forceinline__ void db_pack4(const uint32_t* r, uint32_t& wa, uint32_t& wb) {
  const uint32_t x01 = prmt(r[0], r[1], 0x6420), x23 = prmt(r[2], r[3], 0x6420);
  wa = prmt(x01, x23, 0x6420);
  wb = prmt(x01, x23, 0x7531);
}

__device__ __forceinline__ EdgeConst db_edge_const(const DbTables& tb, int qp_av, int off_a, int off_b, uint32_t chroma_inc) {
  const int ia = db_clip(qp_av + off_a, 0, 51), ib = db_clip(qp_av + off_b, 0, 51);
  const uint4 a = tb.a[ia];
  EdgeConst k;
  k.ka = a.x;
  k.ks = a.y;
  k.tcb = a.z + chroma_inc;
  k.lo1 = a.w;
  k.kb = tb.b[ib];
  return k;
}

// the four edges of one direction over the twenty packed registers of a lane (r[4e - 4 .. 4e + 3] around edge e)
__device__ __forceinline__ void db_filter_edges(uint32_t* r, const EdgeConst& k_mb, const EdgeConst& k_in, uint32_t off0,
                                                uint32_t off_odd, uint32_t off_hi, uint32_t chroma) {
  filter_edge_strong(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], k_mb, off0, chroma);
  filter_edge_normal(r[5], r[6], r[7], r[8], r[9], r[10], k_in, off_odd, chroma);
  filter_edge_normal(r[9], r[10], r[11], r[12], r[13], r[14], k_in, off_hi, chroma);
  filter_edge_normal(r[13], r[14], r[15], r[16], r[17], r[18], k_in, off_odd | off_hi, chroma);
}

__global__ void __launch_bounds__(32 * kDbWarps, 4) deblock_wavefront_kernel(const DeblockArgs a) {
  __shared__ DbTables tb;
  __shared__ DbHalf tiles[kDbWarps][2];
  __shared__ unsigned int s_row[kDbWarps];
  for (int i = threadIdx.x; i < 52; i += blockDim.x) {
    const EdgeConst k = make_edge_const(kDbAlpha[i], kDbBeta[i], kDbTc0[i], false);
    tb.a[i] = make_uint4(k.ka, k.ks, k.tcb, k.lo1);
    tb.b[i] = k.kb;
    tb.qmap[0][i] = (uint8_t)i;
    tb.qmap[1][i] = kDbQpc[db_clip(i + a.cb_off, 0, 51)];
    tb.qmap[2][i] = kDbQpc[db_clip(i + a.cr_off, 0, 51)];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, half = lane >> 4, hl = lane & 15;
  DbHalf& t = tiles[wid][half];
  const int W = a.W, H = a.H;
  const size_t lw = 16 * (size_t)W, cw = 8 * (size_t)W;
  const size_t luma_bytes = lw * 16 * H, frame_bytes = luma_bytes * 3 / 2;
  const unsigned total_rows = (unsigned)((a.n_frames + 1) / 2) * (unsigned)H;
  const unsigned long long tag64 = (unsigned long long)a.tag << 32;
  // ---- lane roles inside the half-warp
  const bool is_luma = hl < 8;
  const int cpl = (hl >> 2) & 1, cj = hl & 3;  // chroma lanes: plane, line / column pair
  const uint32_t chroma_mask = is_luma ? 0u : 0xffffffffu;
  const uint8_t* qmap = tb.qmap[is_luma ? 0 : 1 + cpl];
  // vertical edges: lines la and la + lstep of the macroblock (tile rows 4 + la, 4 + la + lstep)
  const int la = is_luma ? hl : cj, lstep = is_luma ? 8 : 4;
  uint8_t* const plane_tile = is_luma ? t.luma : t.chroma[cpl];
  uint8_t* const v_rowA = plane_tile + (4 + la) * 16;
  uint8_t* const v_rowB = v_rowA + lstep * 16;
  // horizontal edges: two columns of the tile
  uint8_t* const h_col = is_luma ? t.luma + 2 * hl : t.chroma[cpl] + 2 * cj;
  // hand-off words: luma word hl; lanes 0..7 also chroma word hl (plane hl >> 2, row (hl >> 1) & 1, word hl & 1)
  uint8_t* const cw_tile = t.chroma[(hl >> 2) & 1] + ((hl >> 1) & 1) * 16 + 4 * (hl & 1);
  for (;;) {
    if (lane == 0) s_row[wid] = atomicAdd(a.ticket, 1u);
    __syncwarp();
    const unsigned row = s_row[wid];
    __syncwarp();
    if (row >= total_rows) return;
    const int f = 2 * (int)(row / (unsigned)H) + half, my = (int)(row % (unsigned)H);
    const bool active = f < a.n_frames;
    const bool last_row = my == H - 1;
    const size_t frow = (size_t)(active ? f : 0) * H + my;  // this half's macroblock row among all rows
    uint8_t* const Y = a.yuv + (size_t)(active ? f : 0) * frame_bytes;
    uint8_t* const plane = is_luma ? Y : Y + luma_bytes + (cpl ? luma_bytes / 4 : 0);
    const size_t pitch = is_luma ? lw : cw;
    const int mbw = is_luma ? 16 : 8;  // bytes of a macroblock's line in this lane's plane
    uint8_t* const gA = plane + (size_t)(mbw * my + la) * pitch;
    uint8_t* const gB = gA + (size_t)lstep * pitch;
    const bool storeA = active, storeB = active && (last_row || la + lstep < mbw - 3 + (is_luma ? 0 : 2));
    const bool store_top = active && my > 0 && (is_luma ? hl < 3 : cj == 0);
    uint8_t* const g_top = is_luma ? Y + (size_t)(16 * my - 3 + hl) * lw : plane + (size_t)(8 * my - 1) * cw;
    const uint8_t* const top_tile = is_luma ? t.luma + (1 + hl) * 16 : t.chroma[cpl] + 3 * 16;
    const uint8_t* const qp_row = a.qp + frow * W;
    const uint8_t* const t8_row = a.t8x8 + frow * W;
    unsigned long long* const line_mine = a.line + frow * W * kDbLineWords;
    const unsigned long long* const line_above = line_mine - (size_t)W * kDbLineWords;
    const bool poll0 = active && my > 0, poll1 = poll0 && hl < 8;
    const bool pub0 = active && !last_row, pub1 = pub0 && hl < 8;
    bool dead = false;
    uint4 cyA = make_uint4(0, 0, 0, 0), cyB = cyA;
    int nq = 0, nt8 = 0, nqup = 0;
    unsigned long long ftv0 = 0, ftv1 = 0;
    uint32_t pend0 = 0, pend1 = 0;
    auto fetch = [&](int mx) {  // macroblock mx of this row: samples, QPs, and a first look at the words from above
      if (active) {
        if (is_luma) {
          cyA = __ldcg(reinterpret_cast<const uint4*>(gA + 16 * mx));
          cyB = __ldcg(reinterpret_cast<const uint4*>(gB + 16 * mx));
        } else {
          const uint2 c0 = __ldcg(reinterpret_cast<const uint2*>(gA + 8 * mx));
          const uint2 c1 = __ldcg(reinterpret_cast<const uint2*>(gB + 8 * mx));
          cyA.x = c0.x;
          cyA.y = c0.y;
          cyB.x = c1.x;
          cyB.y = c1.y;
        }
        nq = qp_row[mx];
        nt8 = t8_row[mx];
        if (my > 0) nqup = qp_row[mx - W];
      }
      if (poll0) ftv0 = ld_relaxed_gpu_u64(line_above + (size_t)mx * kDbLineWords + hl);
      if (poll1) ftv1 = ld_relaxed_gpu_u64(line_above + (size_t)mx * kDbLineWords + 16 + hl);
    };
    fetch(0);
    int q_left = 0;
    for (int mx = 0; mx < W; mx++) {
      const int q = nq, q_up = nqup;
      unsigned long long tv0 = ftv0, tv1 = ftv1;  // the first look at this macroblock's words; fetch() moves on to the next
      const uint32_t off_odd = (is_luma && nt8) ? 0xffffffffu : 0u;  // luma edges 4 and 12 of an 8x8-transform macroblock
      const uint32_t off_hi = is_luma ? 0u : 0xffffffffu;            // chroma has no edges 8 and 12
      // ---- the lane's edge constants: inner edges, left edge, top edge
      const int qe = qmap[q], qel = qmap[q_left], qeu = qmap[q_up];
      const uint32_t cinc = chroma_mask & 0x00010001u;
      const EdgeConst k_in = db_edge_const(tb, qe, a.off_a, a.off_b, cinc);
      const EdgeConst k_left = db_edge_const(tb, (qe + qel + 1) >> 1, a.off_a, a.off_b, cinc);
      const EdgeConst k_top = db_edge_const(tb, (qe + qeu + 1) >> 1, a.off_a, a.off_b, cinc);
      uint32_t r[20];
      // ---- vertical edges: the left margin is what the previous macroblock left in the tile's last word
      db_unpack4(db_word(v_rowA + mbw - 4), db_word(v_rowB + mbw - 4), r);
      db_unpack4(cyA.x, cyB.x, r + 4);
      db_unpack4(cyA.y, cyB.y, r + 8);
      if (is_luma) {
        db_unpack4(cyA.z, cyB.z, r + 12);
        db_unpack4(cyA.w, cyB.w, r + 16);
      } else {
#pragma unroll
        for (int i = 12; i < 20; i++) r[i] = 0;
      }
      if (mx + 1 < W) fetch(mx + 1);
      db_filter_edges(r, k_left, k_in, mx == 0 ? 0xffffffffu : 0u, off_odd, off_hi, chroma_mask);
      {
        uint32_t mA, mB, a0, b0, a1, b1;
        db_pack4(r, mA, mB);
        db_pack4(r + 4, a0, b0);
        db_pack4(r + 8, a1, b1);
        if (is_luma) {
          uint32_t a2, b2, a3, b3;
          db_pack4(r + 12, a2, b2);
          db_pack4(r + 16, a3, b3);
          *reinterpret_cast<uint4*>(v_rowA) = make_uint4(a0, a1, a2, a3);
          *reinterpret_cast<uint4*>(v_rowB) = make_uint4(b0, b1, b2, b3);
          if (hl >= 4) t.patch[hl - 4] = mB;
        } else {
          *reinterpret_cast<uint2*>(v_rowA) = make_uint2(a0, a1);
          *reinterpret_cast<uint2*>(v_rowB) = make_uint2(b0, b1);
          if (cj >= 2) t.patch[4 + 2 * cpl + (cj - 2)] = mB;
        }
        // the left neighbour's last word of these two lines is final now (its bottom rows go to the row below instead)
        if (mx > 0) {
          if (storeA) db_word(gA + mbw * mx - 4) = mA;
          if (storeB) db_word(gB + mbw * mx - 4) = mB;
        }
      }
      // ---- the words from above into the top margin
      if (poll0) {
        const unsigned long long* p = line_above + (size_t)mx * kDbLineWords + hl;
        unsigned spins = 0;
        while ((uint32_t)(tv0 >> 32) != a.tag || (poll1 && (uint32_t)(tv1 >> 32) != a.tag)) {
          if (++spins > (1u << 22) || ((spins & 1023u) == 0 && *reinterpret_cast<volatile int*>(a.status) == STATUS_WATCHDOG)) {
            atomicExch(a.status, STATUS_WATCHDOG);
            dead = true;
            break;
          }
          tv0 = ld_relaxed_gpu_u64(p);
          if (poll1) tv1 = ld_relaxed_gpu_u64(p + 16);
        }
        db_word(t.luma + 4 * hl) = (uint32_t)tv0;
        if (poll1) db_word(cw_tile + 2 * 16) = (uint32_t)tv1;
      }
      __syncwarp();
      // ---- the previous macroblock is finished down to its last rows: hand those to the row below
      if (mx > 0) {
        if (pub0)
          st_relaxed_gpu_u64(line_mine + (size_t)(mx - 1) * kDbLineWords + hl, tag64 | ((hl & 3) == 3 ? t.patch[hl >> 2] : pend0));
        if (pub1)
          st_relaxed_gpu_u64(line_mine + (size_t)(mx - 1) * kDbLineWords + 16 + hl,
                             tag64 | ((hl & 1) ? t.patch[4 + (hl >> 1)] : pend1));
      }
      // ---- horizontal edges: two columns per lane, rows -4..15 (chroma: -4..7)
#pragma unroll
      for (int i = 0; i < 20; i++) {
        if (i < 12 || is_luma) r[i] = prmt(*reinterpret_cast<const uint16_t*>(h_col + i * 16), 0, 0x4140);
        else r[i] = 0;
      }
      db_filter_edges(r, k_top, k_in, my == 0 ? 0xffffffffu : 0u, off_odd, off_hi, chroma_mask);
#pragma unroll
      for (int i = 1; i < 18; i++)
        if (i < 12 || is_luma) *reinterpret_cast<uint16_t*>(h_col + i * 16) = (uint16_t)prmt(r[i], 0, 0x4420);
      __syncwarp();
      // ---- stores: the lane's two lines of the macroblock (not the rows the next picture row still filters) and the
      // upper neighbour's last rows; the last rows of this macroblock wait in registers for the hand-off
      if (is_luma) {
        if (storeA) *reinterpret_cast<uint4*>(gA + 16 * mx) = *reinterpret_cast<const uint4*>(v_rowA);
        if (storeB) *reinterpret_cast<uint4*>(gB + 16 * mx) = *reinterpret_cast<const uint4*>(v_rowB);
        if (store_top) *reinterpret_cast<uint4*>(g_top + 16 * mx) = *reinterpret_cast<const uint4*>(top_tile);
      } else {
        if (storeA) *reinterpret_cast<uint2*>(gA + 8 * mx) = *reinterpret_cast<const uint2*>(v_rowA);
        if (storeB) *reinterpret_cast<uint2*>(gB + 8 * mx) = *reinterpret_cast<const uint2*>(v_rowB);
        if (store_top) *reinterpret_cast<uint2*>(g_top + 8 * mx) = *reinterpret_cast<const uint2*>(top_tile);
      }
      pend0 = db_word(t.luma + 16 * 16 + 4 * hl);
      pend1 = db_word(cw_tile + 10 * 16);
      q_left = q;
      __syncwarp();
    }
    if (pub0) st_relaxed_gpu_u64(line_mine + (size_t)(W - 1) * kDbLineWords + hl, tag64 | pend0);
    if (pub1) st_relaxed_gpu_u64(line_mine + (size_t)(W - 1) * kDbLineWords + 16 + hl, tag64 | pend1);
    if (__any_sync(0xffffffffu, dead)) return;
  }
}

}  // namespace dryv
