// recon_tables.h — host-built lookup tables consumed by the reconstruction kernels.
//
// Everything here is derived from formulas, never from measured data:
//   * LevelScale4x4/8x8  (reference: Frame::scaling, src/video/frame/transform.rs:8-78)
//   * chroma QP table     (reference: get_qpc, src/video/frame/transform.rs:194-216)
//   * zig-zag scans       (reference: inverse_scanner4x4 / inverse_scanner_8x8, src/video/frame/mod.rs:185-284)
//   * tap tables for the nine Intra4x4 / Intra8x8 modes (reference: pred4x4.rs:92-359, pred8x8.rs:294-692)
#pragma once
#include <stdint.h>

#include "../../include/dryv_recon.h"

namespace dryv {

// Edge-sample numbering used by the tap tables.
//   4x4: 0..7 = p[0..7,-1] (top, top-right), 8..11 = p[-1,0..3] (left), 12 = p[-1,-1], 13 = always 0;
//        mode 2 (DC) holds the first of two summation rounds: lanes 0..2 add three edge samples each
//   8x8: 0..15 = p'[0..15,-1], 16..23 = p'[-1,0..7], 24 = p'[-1,-1], 25 = DC value
// Every predicted sample of every mode is (E[i0] + 2*E[i1] + E[i2] + 2) >> 2 for a triple of edge
// indices: a 2-tap average (a + b + 1) >> 1 is the triple (a, b, a), a copy is (a, a, a).
enum { E4_LEFT = 8, E4_CORNER = 12, E4_ZERO = 13, E8_LEFT = 16, E8_CORNER = 24, E8_DC = 25 };

struct DeviceTables {
  int32_t t4[52][16];      // [qP][zig-zag k] = LevelScale4x4[qP%6][pos(k)] << max(qP/6 - 4, 0)
  uint16_t ls8[6][64];     // [qP%6][i*8+j]   = LevelScale8x8
  uint16_t lut4[9][16];    // [mode][y*4+x]   = i0 | i1<<4 | i2<<8
  uint16_t lut8[9][64];    // [mode][y*8+x]   = i0 | i1<<5 | i2<<10
  uint8_t zz8inv[8][8];    // [i][j] -> zig-zag index
  uint8_t qpc[52];         // qPI -> QPC
  uint8_t pad[12];
  // Intra4x4 dependency schedule, specialised per macroblock availability av = A | B<<1 | C<<2 | D<<3:
  // [av][step][half] -> tile origin (10 bits) | mode nibble shift (0..28) << 10 | modes-hi-word flag << 15 |
  // legal-mode mask (9 bits) << 16 | top << 25 | left << 26 | corner << 27 | top-right << 28 | active << 29
  uint32_t i4step[16][10][2];
};

void build_device_tables(const dryv_pic_params& pp, DeviceTables* t);

}  // namespace dryv
