// recon_kernels.cuh — sm_100a device code of the AVC intra reconstruction path: data structures of the row teams and the
// prediction stage. (The residual stage — dequantisation, DC Hadamards, inverse transforms — is residual_stage.cuh; the
// kernels themselves are in recon.cu.)
//
// Three kernels:
//   resolve_modes_kernel   pixel-independent pre-pass: Intra4x4/8x8 prediction-mode derivation for a whole picture.
//   recon_wavefront_kernel persistent "row teams" (one CTA = two warps) walking macroblock rows in x+2y wavefront order,
//                          four macroblocks (a group) at a time:
//       front warp  one bulk copy (cp.async.bulk + mbarrier) of the group's 3072 B of levels into a ring, the residual
//                   stage of the group (luma of two macroblocks / chroma of four per pass), tap rows of the Intra4x4
//                   blocks, then chroma prediction of each macroblock. It has no luma dependency on other macroblocks
//                   and runs ahead of the pixels.
//       pixel warp  waits for the bottom line of the row above (64-bit payload|tag words, relaxed loads), predicts + adds
//                   the residual + clips (two samples per VIADDMNMX) into a shared-memory pixel tile, stores 128-bit
//                   rows, publishes its own bottom line. Prediction is a gather from the tile: no shuffles, per-lane
//                   sample addresses come from host-built tap tables.
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

constexpr int kWarpsPerCta = 4;  // residual-only kernel and level expansion
constexpr int kThreadsPerCta = kWarpsPerCta * 32;

constexpr int kLumaStride = kLumaTileStride;  // pixel (x, y) at (y + 1) * 48 + 16 + x, x in -4..31 (row -1), y in -1..15
constexpr int kLumaTileBytes = 17 * kLumaStride;
constexpr int kChromaStride = 24;        // pixel (x, y) at (y + 1) * 24 + 8 + x, x in -4..15 (row -1), y in -1..7
constexpr int kChromaTileBytes = 9 * kChromaStride;

// wavefront kernel: a "row team" = one CTA of two warps walking one macroblock row, a group of kGroupMbs macroblocks at
// a time. The front warp hands each group to the pixel warp through a ring of kGroupSlots group slots.
#ifndef DRYV_GROUP_SLOTS
#define DRYV_GROUP_SLOTS 2
#endif
constexpr int kGroupSlots = DRYV_GROUP_SLOTS;
#ifndef DRYV_CLUSTER
#define DRYV_CLUSTER 1
#endif
constexpr int kCluster = DRYV_CLUSTER;  // rows per cluster (1: every row team on its own, line hand-off through L2 only)
#ifndef DRYV_RING
#define DRYV_RING 16
#endif
constexpr int kRingEntries = DRYV_RING;  // line ring of the cluster mode, macroblocks
constexpr int kMailSlots = 4;
constexpr int kLineWords = 8;  // words of a macroblock's bottom line, see below
static_assert((kRingEntries & (kRingEntries - 1)) == 0 && kRingEntries >= 4, "ring size: a power of two");
// level ring: with two stages the bulk copy of group g + 1 may be issued before the residual stage of group g; the kernel
// issues it after that stage (DRYV_FETCH_LATE), where one stage is enough (the levels of group g have been consumed)
#ifndef DRYV_LV_STAGES
#define DRYV_LV_STAGES 2
#endif
constexpr int kLvStages = DRYV_LV_STAGES;
constexpr int kTeamThreads = 64;
struct MbSlot {
  alignas(16) uint16_t res[kResLumaTile];  // luma residual fields (residual_stage.cuh)
  alignas(16) uint8_t rows[32];  // Intra4x4: tap rows, half-warp A's steps 0..9 then half-warp B's steps 2..7; rest 0
  uint32_t modes_lo, modes_hi;   // resolved Intra4x4/8x8 modes in schedule order (resolve_modes_kernel)
  int32_t mbcls, mode16;         // 0/1/2 = Intra4x4/8x8/16x16; Intra16x16 prediction mode
};
static_assert(sizeof(MbSlot) % 16 == 0, "slot alignment");
struct GroupSlot {
  MbSlot mb[kGroupMbs];
  int32_t frame, row, x0, n;     // row < 0: no more work
};
struct TeamSmem {
  GroupSlot grp[kGroupSlots];
  alignas(16) int16_t lv[kLvStages][kGroupMbs * DRYV_COEFFS_PER_MB];  // level ring (bulk-copy destinations)
  // chroma residual fields of the front warp's current group; the 8x8 passes, which run before the chroma pass writes
  // them, transpose through the same bytes
  alignas(16) uint16_t cres[kGroupMbs][kResChromaMb];
  alignas(16) uint32_t hdr[kGroupMbs];                      // class | qp << 8 | chroma mode << 16 | mb_type << 24
  alignas(16) uint8_t luma[kLumaTileBytes];                 // luma pixel tile (pixel warp)
  alignas(16) uint8_t chroma[2 * kChromaTileBytes];         // chroma pixel tiles (front warp)
  alignas(16) uint8_t lcol[16];                             // right-most luma column of the previous MB, contiguous
  alignas(16) uint8_t ccol[16];                             // right-most Cb | Cr columns of the previous MB
  alignas(16) uint8_t e8[32];                               // filtered edge vector p' of the current Intra8x8 block
  alignas(8) unsigned long long full[kGroupSlots];          // mbarriers: group slot filled by the front warp
  alignas(8) unsigned long long lvfull[kLvStages];          // mbarriers: level stage landed
  uint32_t pace;                                            // holds its own address: pacing chain of the poll loops
#if DRYV_CLUSTER > 1
  // Cluster mode (a cluster of DRYV_CLUSTER row teams walks DRYV_CLUSTER consecutive rows of one picture): the bottom
  // lines of the row above arrive in this ring through distributed shared memory instead of the tagged words in L2.
  alignas(16) unsigned long long ring[kRingEntries * kLineWords];  // entry s & (kRingEntries - 1): payload | (s + 1) << 32
  alignas(8) unsigned long long mail[kMailSlots];                  // tickets from rank 0: t | (k + 1) << 32
  volatile uint32_t cons[2];                                       // entries the row below has consumed: luma, chroma
  volatile uint32_t ack[DRYV_CLUSTER];                             // rank 0 only: tickets rank r has taken from its mailbox
#endif
};
static_assert(sizeof(uint16_t) * kGroupMbs * kResChromaMb >= sizeof(int) * kScratchWords, "scratch aliases the chroma tiles");
// A CTA holds several row teams that share one copy of the tables: shared memory, not registers, is what limits the
// number of teams per SM, and every team in flight hides latency for the others (each team is a serial chain).
#ifndef DRYV_TEAMS_PER_CTA
#define DRYV_TEAMS_PER_CTA 1
#endif
constexpr int kTeamsPerCta = DRYV_TEAMS_PER_CTA;
constexpr int kWaveThreads = kTeamThreads * kTeamsPerCta;
struct WaveCtaSmem {
  alignas(16) unsigned char tab[kTeamTableBytes];           // DeviceTables without t4
  TeamSmem team[kTeamsPerCta];
};

constexpr int kResidMbFields = kResLumaTile + kResChromaMb;  // 464 fields = 928 bytes per macroblock (a multiple of 16)
static_assert((kResidMbFields * 2) % 16 == 0 && (kResLumaTile * 2) % 16 == 0, "bulk-copy granularity");
enum { STATUS_OK = 0, STATUS_UNSUPPORTED = 1, STATUS_WATCHDOG = 2 };

// Bottom line a macroblock hands to the row below: 4 luma words (16 px), 2 Cb, 2 Cr words (8 px each).
// Every 32-bit payload travels with the launch tag in one 64-bit word, so a single relaxed 64-bit load
// both fetches the data and proves it is there: no fence, no separate flag, no second round trip.
// (kLineWords = 8 is defined above TeamSmem)

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
  uint16_t* resid;          // split path: [n_mbs][kResidMbFields] biased residual fields (recon_residual_fields_kernel)
  unsigned int* ticket_c;   // split path: row ticket counter of the chroma walkers
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
// In cluster mode a line word lives either in global memory or in the shared memory of a CTA of the cluster (the ring):
// the accesses are generic there, so that one instruction stream serves both.
#if DRYV_CLUSTER > 1
#define DRYV_LINE_SPACE ""
#else
#define DRYV_LINE_SPACE ".global"
#endif
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu" DRYV_LINE_SPACE ".u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu" DRYV_LINE_SPACE ".u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// ---- thread-block cluster helpers (cluster mode) ----
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic address of `p` (a generic pointer into this CTA's shared memory) in the CTA with rank `rank` of the cluster
template <typename T>
__device__ __forceinline__ T* map_rank(T* p, unsigned rank) {
  unsigned long long r;
  asm volatile("mapa.u64 %0, %1, %2;" : "=l"(r) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
  return reinterpret_cast<T*>(r);
}
__device__ __forceinline__ void st_relaxed_cluster_u32(volatile uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.cluster.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int clip255(int v) { return vimin_relu_s32(v, 255); }
__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d) {
  return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24);
}
__device__ __forceinline__ uint32_t splat4(int v) { return (uint32_t)v * 0x01010101u; }

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
  pl.i4_res2 = 2 * (py * kResLumaStride + px);
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
  pl.i8_res2 = 2 * ((lane >> 2) * kResLumaStride + (lane & 3) * 2);
  pl.i8_tab = lane * 8;
  return pl;
}

// ---- Intra4x4 luma, pred4x4.rs:10-360 + transform.rs:98-110 -------------------------------------------
// Ten dependency steps, two blocks per step where the decode-order availability rules allow it (kI4BlkA / kI4BlkB:
// half-warp B's block always sits 8 px right / 4 px up of half-warp A's); one pixel per lane, 16 lanes per block.
// Everything that depends on the mode or on availability was folded into the tap row chosen by the front warp
// (i4_tap_row), so a step is: tap record (one 128-bit load), three sample loads, three adds, a shift, the residual and a
// clamp. The ten steps are unrolled with the block positions as immediates (about half the instructions of a rolled
// loop that reads them from a table); the DC flavours, a quarter of the blocks, share one out-of-line function.
//   rows: the slot's tap-row bytes (MbSlot::rows)
#ifndef DRYV_I4_ROLLED
#define DRYV_I4_ROLLED 1
#endif
#ifndef DRYV_I4_DC_INLINE
#define DRYV_I4_DC_INLINE 0
#endif
#if DRYV_I4_ROLLED
#ifndef DRYV_I4_UNROLL
#define DRYV_I4_UNROLL 1
#endif
// steps per loop iteration (development knob). Measured: 2 makes an Intra4x4-only batch 12 % faster (0.970 -> 0.856 ms,
// the register moves of the software pipeline go) for 64 more static instructions, and the mixed batch 4 % slower with
// batches in flight (0.713 -> 0.741; unchanged one batch at a time); 5: Intra4x4-only 0.848, mixed 0.947 — the kernel's hot
// code sits at the instruction-cache capacity. 2 together with the one-stage level ring (-24 static
// instructions): mixed batch 0.718 in flight (0.702), 0.779 one at a time (0.803).
constexpr int kI4Unroll = DRYV_I4_UNROLL;
struct I4Regs {
  uint4 tap;      // three sample offsets (biased), kind
  uint32_t org;   // tile offset of the block origin
  int r;          // residual field (r + 512)
};
__device__ __forceinline__ void i4_fetch(I4Regs& q, const I4Step* e, const uint8_t* tap4, int row, const uint8_t* resp) {
  q.tap = *reinterpret_cast<const uint4*>(tap4 + row * kTap4Row);
  const uint2 st = *reinterpret_cast<const uint2*>(e);
  q.org = st.x;
  q.r = *reinterpret_cast<const uint16_t*>(resp + st.y);
}
__device__ __forceinline__ void predict_i4x4(const DeviceTables& tab, uint8_t* lt, const uint16_t* res_luma,
                                             const PixLane& pl, const uint8_t* rows0) {
  const uint8_t* rows = rows0 + 8 * pl.half;
  const I4Step* st = &tab.i4tab[0][pl.half];
  const uint8_t* tap4 = reinterpret_cast<const uint8_t*>(&tab.tap4[0][0][0]) + pl.i4_tab;
  const uint8_t* resp = reinterpret_cast<const uint8_t*>(res_luma) + pl.i4_res2;
  const uint8_t* ltb = lt - kTap4Bias;
  // software pipeline: the table look-ups of step s+1 (which do not depend on any pixel) are issued under the latency
  // of the pixel loads of step s, so a step's dependent chain is pixel load -> three adds -> clamp -> store
  I4Regs cur;
  i4_fetch(cur, st, tap4, rows[0], resp);
#pragma unroll kI4Unroll
  for (int s = 0; s < 10; s++) {
    const uint8_t* ob = ltb + cur.org;
    const int e0 = ob[cur.tap.x], e1 = ob[cur.tap.y], e2 = ob[cur.tap.z];
    int kind = (int)cur.tap.w;
    const int r = cur.r - kResBias;
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

#else
#if DRYV_I4_DC_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
int i4_dc_pred(const uint8_t* eb, int kind) {  // pred4x4.rs:116-167; eb = block origin in the tile
  const int sT = dp4a_us(*reinterpret_cast<const uint32_t*>(eb - kLumaStride), 0x01010101, 0);
  const int sL = eb[-1] + eb[kLumaStride - 1] + eb[2 * kLumaStride - 1] + eb[3 * kLumaStride - 1];
  return kind == kI4KindDc ? ((sT + sL + 4) >> 3)
                           : (kind == kI4KindDcTop ? ((sT + 2) >> 2) : (kind == kI4KindDcLeft ? ((sL + 2) >> 2) : 128));
}
struct I4Tap {
  uint4 t;  // three sample offsets (biased by kTap4Bias, relative to the block origin), kind
  int r;    // residual field (r + 512)
};
__host__ __device__ constexpr int i4_org(int blk) { return luma_at(((blk >> 2) & 1) * 8 + (blk & 1) * 4, (blk >> 3) * 8 + ((blk >> 1) & 1) * 4); }
__host__ __device__ constexpr int i4_res2(int blk) {
  return 2 * (((blk >> 3) * 8 + ((blk >> 1) & 1) * 4) * kResLumaStride + ((blk >> 2) & 1) * 8 + (blk & 1) * 4);
}
constexpr int kI4HalfTile = 8 - 4 * kLumaStride;          // half-warp B's block relative to A's, tile bytes
constexpr int kI4HalfRes2 = 2 * (8 - 4 * kResLumaStride);  // ... residual bytes
template <int S>
__device__ __forceinline__ I4Tap i4_fetch(const uint8_t* tapl, uint2 rb, uint32_t rb2, const uint8_t* resl) {
  // tap row of step S = byte S of the half-warp's row bytes; times kTap4Row (256): one PRMT puts it into byte 1
  const uint32_t w = S < 4 ? rb.x : (S < 8 ? rb.y : rb2);
  const uint32_t rowoff = prmt(w, 0u, 0x4404u | ((uint32_t)(S < 8 ? (S & 3) : S - 8) << 4));
  I4Tap q;
  q.t = *reinterpret_cast<const uint4*>(tapl + rowoff);
  q.r = *reinterpret_cast<const uint16_t*>(resl + i4_res2(kI4BlkA[S]));
  return q;
}
//   ltl: sample gather base of the half-warp (tile - kTap4Bias + half-warp displacement); dstl: block origin base
template <int S>
__device__ __forceinline__ void i4_step(const I4Tap& q, const uint8_t* ltl, uint8_t* dstl, int i4_pix, bool has_blk) {
  constexpr int org = i4_org(kI4BlkA[S]);
  if (has_blk) {
    const int e0 = ltl[q.t.x + org], e1 = ltl[q.t.y + org], e2 = ltl[q.t.z + org];
    int kind = (int)q.t.w;
    int pred = (e0 + 2 * e1 + e2 + 2) >> 2;
    if (kind >= kI4KindDc) {
      pred = i4_dc_pred(dstl + org, kind);
      kind = 1;
    }
    dstl[org + i4_pix] = (uint8_t)clip255(pred * kind + q.r - kResBias);
  }
}
static_assert(kTap4Row == 256, "i4_fetch multiplies the row index by placing it in byte 1");
__device__ __forceinline__ void predict_i4x4(const DeviceTables& tab, uint8_t* lt, const uint16_t* res_luma,
                                             const PixLane& pl, const uint8_t* rows) {
  const bool isA = pl.half == 0;
  const int hoff = isA ? 0 : kI4HalfTile;
  const uint8_t* tapl = reinterpret_cast<const uint8_t*>(&tab.tap4[0][0][0]) + pl.i4_tab;
  const uint8_t* resl = reinterpret_cast<const uint8_t*>(res_luma) + pl.i4_res2 + (isA ? 0 : kI4HalfRes2);
  const uint8_t* ltl = lt - kTap4Bias + hoff;
  uint8_t* dstl = lt + hoff;
  // the half-warp's ten row bytes: A rows[0..9], B rows[8..17] (B's steps 2..7 are bytes 10..15, 16.. stay zero)
  const uint2 rb = *reinterpret_cast<const uint2*>(rows + 8 * pl.half);
  const uint32_t rb2 = *reinterpret_cast<const uint16_t*>(rows + 8 * pl.half + 8);
  // software pipeline: the table look-ups of step s+1 (which do not depend on any pixel) are issued before the warp
  // barrier that closes step s, so a step's dependent chain is pixel load -> three adds -> clamp -> store
#define DRYV_I4_STEP(S, HASB)                                                  \
  {                                                                            \
    const I4Tap nx = i4_fetch<((S) < 9 ? (S) + 1 : 9)>(tapl, rb, rb2, resl);   \
    i4_step<(S)>(cur, ltl, dstl, pl.i4_pix, (HASB) || isA);                    \
    __syncwarp();                                                              \
    cur = nx;                                                                  \
  }
  I4Tap cur = i4_fetch<0>(tapl, rb, rb2, resl);
  DRYV_I4_STEP(0, false)
  DRYV_I4_STEP(1, false)
  DRYV_I4_STEP(2, true)
  DRYV_I4_STEP(3, true)
  DRYV_I4_STEP(4, true)
  DRYV_I4_STEP(5, true)
  DRYV_I4_STEP(6, true)
  DRYV_I4_STEP(7, true)
  DRYV_I4_STEP(8, false)
  DRYV_I4_STEP(9, false)
#undef DRYV_I4_STEP
}

#endif

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
                                             const uint16_t* res_luma, const PixLane& pl, int lane, uint32_t modes_lo,
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
    const uint32_t rw = *reinterpret_cast<const uint32_t*>(resp8 + 2 * (((blk >> 1) * 8) * kResLumaStride + (blk & 1) * 8));
    __syncwarp();
    int pr0, pr1;
    if (mode == 2) {  // DC, pred8x8.rs:326-425 (warp-uniform): sums of the filtered top 0..7 and left 0..7
      const uint32_t* ew = reinterpret_cast<const uint32_t*>(e8);
      const int sT = dp4a_us(ew[1], 0x01010101, dp4a_us(ew[0], 0x01010101, 0));
      const int sL = dp4a_us(ew[5], 0x01010101, dp4a_us(ew[4], 0x01010101, 0));
      pr0 = (aT && aL) ? ((sT + sL + 8) >> 4) : (aL ? ((sL + 4) >> 3) : (aT ? ((sT + 4) >> 3) : 128));
      pr0 -= kResBias;
      pr1 = pr0;
    } else {
      // the - 512 that cancels the bias of the residual field rides on the rounding constant (2 - 4 * 512)
      pr0 = ((int)e8[i0] + 2 * (int)e8[i1] + (int)e8[i2] + 2 - 4 * kResBias) >> 2;
      pr1 = ((int)e8[i3] + 2 * (int)e8[i4] + (int)e8[i5] + 2 - 4 * kResBias) >> 2;
    }
    if (!((legal_mask(aT, aL, aTL) >> mode) & 1u)) pr0 = pr1 = -kResBias;
    const uint32_t o = viaddmin_relu_s16x2(rw, prmt((uint32_t)pr0, (uint32_t)pr1, 0x5410u), 0x00ff00ffu);
    *reinterpret_cast<uint16_t*>(&lt[o8 + pl.i8_pix]) = (uint16_t)prmt(o, 0u, 0x4420u);
    __syncwarp();
  }
}

// ---- Intra16x16 luma, pred16x16.rs:79-425 + pred16x16.rs:64-75. Lane = (row, half): 8 pixels ----------
//   lcol: the left neighbour column as 16 contiguous bytes.
__device__ __forceinline__ void predict_i16x16(uint8_t* lt, const uint8_t* lcol, const uint16_t* res_luma, int lane,
                                               int mode, bool availA, bool availB) {
  const int row = lane >> 1, h = lane & 1;
  const uint4 tv = *reinterpret_cast<const uint4*>(&lt[luma_at(0, -1)]);
  const uint2* rp = reinterpret_cast<const uint2*>(&res_luma[row * kResLumaStride + 8 * h]);
  const uint2 r0 = rp[0], r1 = rp[1];
  uint32_t p0, p1;  // prediction, four samples per word
  if (mode == 0) {  // vertical
    p0 = availB ? (h ? tv.z : tv.x) : 0u;
    p1 = availB ? (h ? tv.w : tv.y) : 0u;
  } else if (mode == 1) {  // horizontal
    p0 = p1 = availA ? splat4(lcol[row]) : 0u;
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
      p0 = p1 = splat4(dc);
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
      int pr[8];
#pragma unroll
      for (int k = 0; k < 8; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
      p0 = pack4(pr[0], pr[1], pr[2], pr[3]);
      p1 = pack4(pr[4], pr[5], pr[6], pr[7]);
    }
  }
  // rows 0..15 / columns 0..15 are written, row -1 and the left column vector are read: no hazard
  *reinterpret_cast<uint2*>(&lt[luma_at(8 * h, row)]) = make_uint2(add_clip4(r0.x, r0.y, p0), add_clip4(r1.x, r1.y, p1));
  __syncwarp();
}

// ---- Chroma Cb + Cr, trans_chroma.rs:96-366 + trans_chroma.rs:81-92. Lane = (plane, row, half): 4 px ---
//   ct: two chroma tiles of kChromaTileBytes each; ccol: left neighbour columns, 8 bytes per plane;
//   res_chroma: the macroblock's chroma residual tile (residual_stage.cuh)
__device__ __forceinline__ void predict_chroma(uint8_t* ct, const uint8_t* ccol, const uint16_t* res_chroma, int lane,
                                               int mode, bool availA, bool availB, bool availD) {
  const int pl = lane >> 4, row = (lane >> 1) & 7, h = lane & 1;
  uint8_t* tile = ct + pl * kChromaTileBytes;
  const uint8_t* cv = ccol + pl * 8;
  const uint32_t tw = *reinterpret_cast<const uint32_t*>(&tile[chroma_at(4 * h, -1)]);
  const uint2 rv = *reinterpret_cast<const uint2*>(&res_chroma[pl * kResChromaPlane + row * 8 + 4 * h]);
  uint32_t p;  // prediction, four samples
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
    p = splat4(val);
  } else if (mode == 1) {
    p = availA ? splat4(cv[row]) : 0u;
  } else if (mode == 2) {
    p = availB ? tw : 0u;
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
    int pr[4];
#pragma unroll
    for (int k = 0; k < 4; k++) pr[k] = ok ? clip255((base + bb * k) >> 5) : 0;
    p = pack4(pr[0], pr[1], pr[2], pr[3]);
  }
  *reinterpret_cast<uint32_t*>(&tile[chroma_at(4 * h, row)]) = add_clip4(rv.x, rv.y, p);
  __syncwarp();
}

}  // namespace dryv
