// recon_tables.h — host-built lookup tables consumed by the reconstruction kernels.
//
// Everything here is derived from formulas, never from measured data:
//   * LevelScale4x4/8x8  (reference: Frame::scaling, src/video/frame/transform.rs:8-78)
//   * chroma QP table     (reference: get_qpc, src/video/frame/transform.rs:194-216)
//   * zig-zag scans       (reference: inverse_scanner4x4 / inverse_scanner_8x8, src/video/frame/mod.rs:185-284)
//   * tap tables for the nine Intra4x4 / Intra8x8 modes (reference: pred4x4.rs:92-359, pred8x8.rs:294-692)
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/dryv_recon.h"

namespace dryv {

// Prediction as a gather: every predicted sample of every Intra4x4 / Intra8x8 mode other than DC is
// (E[i0] + 2*E[i1] + E[i2] + 2) >> 2 for a triple of edge samples (a 2-tap average (a + b + 1) >> 1 is the
// triple (a, b, a), a copy is (a, a, a)). The tables hold, per mode and lane, WHERE those three samples are:
//   tap4: byte offsets into the shared-memory luma tile, relative to the block origin, biased by +256.
//         Edge samples of a 4x4 block: top i (0..7) at -48 + i, left k (0..3) at 48 k - 1, corner at -49.
//         Variant 1 ("no top-right") maps top 4..7 onto top 3 (pred4x4.rs:66-76).
//   tap8: indices into the filtered edge vector p' of an 8x8 block: 0..15 top, 16..23 left, 24 corner.
enum { E4_LEFT = 8, E4_CORNER = 12, E8_LEFT = 16, E8_CORNER = 24 };
constexpr int kLumaTileStride = 48;
constexpr int kResLumaStride = 20;  // 16-bit fields per row of a luma residual tile (16 used; see residual_stage.cuh)
constexpr int kTap4Bias = 256;

// Intra4x4 schedule: ten dependency steps, blocks (spec 4x4 block order) of half-warp A / B per step.
constexpr int kI4BlkA[10] = {0, 1, 2, 3, 8, 9, 10, 11, 14, 15};
constexpr int kI4BlkB[10] = {-1, -1, 4, 5, 6, 7, 12, 13, -1, -1};

// One Intra4x4 schedule entry (see DeviceTables::i4tab): 32-bit fields, used as they are loaded.
// A half-warp without a block in a step (B in steps 0, 1, 8, 9) gets kI4DummyOrg: it predicts into an unused corner of the
// luma tile (columns 16.. of rows 0..15 are padding), so the step needs no "active" predicate.
struct I4Step {
  uint32_t org;     // tile offset of the block origin
  uint32_t res2;    // byte offset of the block's first residual field inside the luma residual tile
};
constexpr uint32_t kI4DummyOrg = (4 + 1) * kLumaTileStride + 16 + 20;  // pixel (20, 4): rows 4..7, columns 20..23
// Rows of DeviceTables::tap4. A row is chosen per block by the front warp from the block's mode and its
// neighbour availability, so the pixel warp does no legality or availability arithmetic at all.
//   0..8   modes 0..8, top-right available      9..17  modes 0..8, top-right replaced by top sample 3
//   18     a mode whose neighbours are missing: prediction 0 (the reference writes nothing, quirk Q4)
//   2, 11  DC with top and left      19  DC, top only      20  DC, left only      21  DC, neither (128)
// The fourth entry of a pixel's tap record is the row's kind: 0 illegal, 1 three-tap gather, 2..5 the DC flavours.
enum { kI4RowIllegal = 18, kI4RowDcTop = 19, kI4RowDcLeft = 20, kI4RowDcNone = 21, kI4Rows = 22 };
enum { kI4KindIllegal = 0, kI4KindTaps = 1, kI4KindDc = 2, kI4KindDcTop = 3, kI4KindDcLeft = 4, kI4KindDcNone = 5 };

// Laid out by who reads what: the residual passes need the first part only (recon_residual_add_kernel copies just
// that to shared memory), the row teams of the wavefront kernel the first two; t4 (the general dequantisation, used
// only for qP without a byte-scale form) stays in global memory.
struct DeviceTables {
  // ---- residual stage ----
  uint16_t ls8[6][64];        // [qP%6][i*8+j]   = LevelScale8x8
  // Byte-scale form of t4 for the packed residual stage (residual_stage.cuh). Where the dequantisation of qP
  // (transform.rs:143-155) is an exact multiplication d = level * m * 2^e with byte factors m (and m a multiple of 4 when
  // e > 0, so that the two halvings of the row and column passes stay exact in units of 2^e), word w of t4b[qP] holds
  // the factor of level 2w in byte 0 and the factor of level 2w+1 in byte 3: the operand of one dp2a.lo / dp2a.hi per
  // level, which extracts the int16 level from its packed word and multiplies it in a single instruction. Flat lists:
  // LevelScale << (qP/6 - 4) = normAdjust << qP/6, so m = normAdjust << qP/6 (e = 0) below qP 12 and m = 4 * normAdjust,
  // e = qP/6 - 2 from there on. t4b_e[qP] = e, or 0xff when qP has no such form (the stage then uses t4).
  uint32_t t4b[52][8];
  int32_t ls00[8];            // [qP%6] = LevelScale4x4[qP%6][0][0] (the DC transforms), 6 used
  uint8_t t4b_e[52];
  uint8_t qpc[52];            // qPI -> QPC
  uint8_t zz8inv[8][8];       // [i][j] -> zig-zag index
  // positions of the set bits of a 4-bit mask, lowest first, one nibble each (0xf = none): pairs the macroblocks of a
  // group of four that take the same residual pass
  uint16_t setbits4[16];
  uint8_t pad0[8];
  // ---- prediction (row teams) ----
  // [av][k], av = A | B<<1 | C<<2 | D<<3: legal-mode mask (9 bits; bit 0 doubles as "top available", bit 1 as
  // "left available") | no-top-right variant << 9 of the block whose tap row is byte k of the slot's row bytes
  // (k = 0..9: half-warp A's steps, k = 10..15: half-warp B's steps 2..7)
  uint16_t i4row[16][16];
  uint32_t tap4[kI4Rows][16][4];  // [row][pixel y*4+x] = three sample offsets, kind (one 128-bit load, no unpacking)
  uint8_t tap8[9][32][8];     // [mode][lane][pixel q (0,1) * 3 + tap], 2 pad bytes
  I4Step i4tab[11][2];        // Intra4x4 schedule: [step][half-warp]; entry 10 repeats 9 (look-ahead of the last step)
  uint8_t i4sched[10][2];     // copy of kI4BlkA / kI4BlkB (0xff = none), for the host-side tests
  uint8_t pad1[12];
  // ---- global memory only ----
  int32_t t4[52][16];         // [qP][zig-zag k] = LevelScale4x4[qP%6][pos(k)] << max(qP/6 - 4, 0)
};
static_assert(sizeof(DeviceTables) % 16 == 0, "DeviceTables is copied with 128-bit loads");
static_assert(offsetof(DeviceTables, t4b) % 16 == 0 && offsetof(DeviceTables, tap4) % 16 == 0 &&
                  offsetof(DeviceTables, ls8) % 16 == 0 && offsetof(DeviceTables, i4row) % 16 == 0 &&
                  offsetof(DeviceTables, t4) % 16 == 0 && offsetof(DeviceTables, i4tab) % 8 == 0,
              "tables read with vector loads / copied in 128-bit pieces");
constexpr size_t kResidTableBytes = offsetof(DeviceTables, i4row);  // what the residual passes read
constexpr size_t kTeamTableBytes = offsetof(DeviceTables, t4);      // what a row team reads

void build_device_tables(const dryv_pic_params& pp, DeviceTables* t);

}  // namespace dryv
