/*
 * dryv_oracle.c — CPU restatement of dryv's AVC intra macroblock reconstruction (src/video/frame/).
 *
 * TEST INFRASTRUCTURE ONLY. This file is the parity checker for the CUDA path and the "port" CPU
 * baseline of bench.py. Nothing under dryv_b200/ (the product) may link, import or call it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY UNPINNED: the reference (Stuff7/dryv, Rust) ships no tests, golden vectors or fixtures for
 * this path and cannot be compiled in this environment (no cargo/rustc). This restatement was written
 * by reading the reference sources cited below; it is cross-checked against an independent
 * restatement of the H.264 text (oracle/spec_model.py) and against hand-computed vectors under
 * tests/golden/, but not against output of the reference binary itself.
 * What IS pinned, to an independent conformant decoder: tests/test_libavcodec_crosscheck.py writes real CABAC
 * High-profile streams from the syntax buffers (tests/avc/stream.py), libavcodec (cv2) decodes them, and this
 * oracle's luma equals libavcodec's bit for bit wherever the reference follows the standard (everywhere except
 * Intra8x8 macroblocks in column 0, quirk Q2); a stream + libavcodec luma fixture is committed under
 * tests/golden/avc/. The reference's deviations themselves (Q1-Q5) rest on reading its source.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root;
 * "frame/x.rs" = src/video/frame/x.rs, "slice/x.rs" = src/video/slice/x.rs). All arithmetic is
 * int64_t where the reference uses isize. Unavailable neighbour samples carry the sentinel -1 exactly
 * as in the reference, including the sites that test "> 0" instead of ">= 0"  and
 * the Intra8x8 reference-filter overwrite (quirk Q2).
 *
 * Plane storage: the reference keeps column-major planes plane[x][y] (frame/mod.rs:17-19); here a
 * plane is one row-major array and P(x,y) indexes it, which yields the same file bytes in
 * write_to_yuv_file order (frame/mod.rs:48-70).
 */
#include "dryv_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t isz;

/* ---- src/math.rs:109-125 ------------------------------------------------------------------- */
static isz clampz(isz v, isz lo, isz hi) { return v < lo ? lo : (v > hi ? hi : v); }

static isz inverse_raster_scan(isz a, isz b, isz c, isz d, isz e) {
  return e == 0 ? (a % (d / b)) * b : (a / (d / b)) * c;
}

/* ---- slice/macroblock.rs:434-477 (MbPosition) ---------------------------------------------- */
enum { POS_NONE = 0, POS_THIS, POS_A, POS_B, POS_C, POS_D };

static int from_coords(isz x, isz y, isz max_w, isz max_h) {
  if (x < 0 && y < 0) return POS_D;
  if (x < 0 && (y >= 0 && y < max_h)) return POS_A;
  if ((x >= 0 && x < max_w) && y < 0) return POS_B;
  if (x > max_w - 1 && y < 0) return POS_C;
  if ((x >= 0 && x < max_w) && (y >= 0 && y < max_h)) return POS_THIS;
  return POS_NONE;
}
static void pos_coords(isz x, isz y, isz max_w, isz max_h, isz* xw, isz* yw) {
  *xw = (x + max_w) % max_w;
  *yw = (y + max_h) % max_h;
}
static isz pos_blk_idx4x4(isz x, isz y, isz max_w, isz max_h) {
  isz xw, yw;
  pos_coords(x, y, max_w, max_h, &xw, &yw);
  return 8 * (yw / 8) + 4 * (xw / 8) + 2 * ((yw % 8) / 4) + ((xw % 8) / 4);
}
static isz pos_blk_idx8x8(isz x, isz y, isz max_w, isz max_h) {
  isz xw, yw;
  pos_coords(x, y, max_w, max_h, &xw, &yw);
  return 2 * (yw / 8) + (xw / 8);
}

/* ---- slice/macroblock.rs:21-129 (Macroblock), recon-relevant fields ------------------------- */
enum { MODE_I4x4 = 0, MODE_I8x8 = 1, MODE_I16x16 = 2 };

/* What a neighbour lookup reads from another macroblock: its part-pred mode and resolved
 * prediction modes (pred4x4.rs:386-412, pred8x8.rs:723-751). */
typedef struct {
  int mode;
  isz intra4x4_pred_mode[16];
  isz intra8x8_pred_mode[4];
} NbMb;

/* Full record of the macroblock being decoded (zeroed per MB like Macroblock::empty,
 * slice/macroblock.rs:156-202). */
typedef struct {
  uint8_t code;           /* mb_type code 0..24 */
  int mode;               /* MODE_* from mb_type_intra, slice/macroblock.rs:682-716 */
  int intra16x16_pred_mode;
  isz qpy, qp1y, qp1c, qpc;
  uint8_t prev_intra4x4_pred_mode_flag[16], rem_intra4x4_pred_mode[16];
  uint8_t prev_intra8x8_pred_mode_flag[4], rem_intra8x8_pred_mode[4];
  uint8_t intra_chroma_pred_mode;
  isz luma_pred_samples[16][4][4];   /* [blk][x][y] */
  isz luma16x16_pred_samples[16][16];/* [x][y] */
  isz luma8x8_pred_samples[4][8][8]; /* [blk][x][y] */
  isz chroma_pred_samples[8][16];    /* [x][y] */
  isz block_luma_dc[16];
  isz block_luma_ac[16][15];
  isz block_luma_4x4[16][16];
  isz block_luma_8x8[4][64];
  isz block_chroma_dc[2][8];
  isz block_chroma_ac[2][8][15];
} CurMb;

/* Slice + Frame state (slice/mod.rs:111-174, frame/mod.rs:16-46). */
typedef struct {
  isz pic_width_in_mbs, pic_height_in_mbs;
  isz width_l, height_l, width_c, height_c;
  isz curr_mb_addr;
  isz chroma_qp_index_offset, second_chroma_qp_index_offset;
  isz scaling_list4x4[16], scaling_list8x8[64]; /* list 0 */
  NbMb* macroblocks; /* n_mb, filled in as MBs are decoded */
  CurMb mb;          /* slice.mb() */
  uint8_t *luma, *cb, *cr;
  isz level_scale4x4[6][4][4];
  isz level_scale8x8[6][8][8];
} Ctx;

#define LUMA(c, x, y) ((c)->luma[(size_t)(y) * (size_t)(c)->width_l + (size_t)(x)])
#define CB(c, x, y) ((c)->cb[(size_t)(y) * (size_t)(c)->width_c + (size_t)(x)])
#define CR(c, x, y) ((c)->cr[(size_t)(y) * (size_t)(c)->width_c + (size_t)(x)])

/* ---- slice/mod.rs:576-622: mb_nb_p + mb_available (no MBAFF, one slice group,
 *      first_mb_in_slice = 0). Returns the neighbour's address or -1 (Macroblock::unavailable). */
static isz mb_nb_p(const Ctx* s, int position) {
  isz mbaddr = s->curr_mb_addr;
  isz w = s->pic_width_in_mbs;
  switch (position) {
    case POS_THIS: return s->curr_mb_addr;
    case POS_A:
      if ((mbaddr % w) == 0) return -1;
      mbaddr -= 1;
      break;
    case POS_B: mbaddr -= w; break;
    case POS_C:
      if (((mbaddr + 1) % w) == 0) return -1;
      mbaddr -= w - 1;
      break;
    case POS_D:
      if ((mbaddr % w) == 0) return -1;
      mbaddr -= w + 1;
      break;
    default: return -1;
  }
  /* mb_available: first_mb_in_slice <= mbaddr <= curr_mb_addr (slice/mod.rs:615-622) */
  if (mbaddr < 0 || mbaddr > s->curr_mb_addr) return -1;
  return mbaddr;
}

/* ---- frame/mod.rs:185-209 -------------------------------------------------------------------- */
static void inverse_scanner4x4(const isz v[16], isz c[4][4]) {
  c[0][0] = v[0];  c[0][1] = v[1];  c[1][0] = v[2];  c[2][0] = v[3];
  c[1][1] = v[4];  c[0][2] = v[5];  c[0][3] = v[6];  c[1][2] = v[7];
  c[2][1] = v[8];  c[3][0] = v[9];  c[3][1] = v[10]; c[2][2] = v[11];
  c[1][3] = v[12]; c[2][3] = v[13]; c[3][2] = v[14]; c[3][3] = v[15];
}

/* ---- frame/mod.rs:212-284: 8x8 frame zig-zag as (row, col) per coefficient index ------------- */
static const uint8_t ZZ8[64][2] = {
    {0, 0}, {0, 1}, {1, 0}, {2, 0}, {1, 1}, {0, 2}, {0, 3}, {1, 2}, {2, 1}, {3, 0}, {4, 0},
    {3, 1}, {2, 2}, {1, 3}, {0, 4}, {0, 5}, {1, 4}, {2, 3}, {3, 2}, {4, 1}, {5, 0}, {6, 0},
    {5, 1}, {4, 2}, {3, 3}, {2, 4}, {1, 5}, {0, 6}, {0, 7}, {1, 6}, {2, 5}, {3, 4}, {4, 3},
    {5, 2}, {6, 1}, {7, 0}, {7, 1}, {6, 2}, {5, 3}, {4, 4}, {3, 5}, {2, 6}, {1, 7}, {2, 7},
    {3, 6}, {4, 5}, {5, 4}, {6, 3}, {7, 2}, {7, 3}, {6, 4}, {5, 5}, {4, 6}, {3, 7}, {4, 7},
    {5, 6}, {6, 5}, {7, 4}, {7, 5}, {6, 6}, {5, 7}, {6, 7}, {7, 6}, {7, 7}};

static void inverse_scanner_8x8(const isz v[64], isz c[8][8]) {
  for (int k = 0; k < 64; k++) c[ZZ8[k][0]][ZZ8[k][1]] = v[k];
}

/* ---- frame/transform.rs:8-78 (8.5.9). Only ever invoked with is_luma = true on intra MBs, so the
 *      list index is 0 for both tables; recomputed for every macroblock like the reference. ---- */
static void scaling(Ctx* s) {
  static const isz V4X4[6][3] = {{10, 16, 13}, {11, 18, 14}, {13, 20, 16},
                                 {14, 23, 18}, {16, 25, 20}, {18, 29, 23}};
  static const isz V8X8[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26},
                                 {26, 23, 42, 24, 33, 31}, {28, 25, 45, 26, 35, 33},
                                 {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
  isz w4[4][4], w8[8][8];
  inverse_scanner4x4(s->scaling_list4x4, w4);
  for (int m = 0; m < 6; m++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        int cls = (i % 2 == 0 && j % 2 == 0) ? 0 : ((i % 2 == 1 && j % 2 == 1) ? 1 : 2);
        s->level_scale4x4[m][i][j] = w4[i][j] * V4X4[m][cls];
      }
  inverse_scanner_8x8(s->scaling_list8x8, w8);
  for (int m = 0; m < 6; m++)
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) {
        int cls;
        if (i % 4 == 0 && j % 4 == 0) cls = 0;
        else if (i % 2 == 1 && j % 2 == 1) cls = 1;
        else if (i % 4 == 2 && j % 4 == 2) cls = 2;
        else if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) cls = 3;
        else if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) cls = 4;
        else cls = 5;
        s->level_scale8x8[m][i][j] = w8[i][j] * V8X8[m][cls];
      }
}

/* ---- frame/transform.rs:194-226 (8.5.8): QPc for Cb / Cr; QpBdOffsetC = 0 (8-bit) ------------- */
static isz get_qpc(const Ctx* s, isz qpy, int is_chroma_cb) {
  static const isz QPCS[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36,
                               36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
  isz off = is_chroma_cb ? s->chroma_qp_index_offset : s->second_chroma_qp_index_offset;
  isz qpi = clampz(qpy + off, 0, 51);
  return qpi < 30 ? qpi : QPCS[qpi - 30];
}
static void chroma_quantization_parameters(Ctx* s, int is_chroma_cb) {
  s->mb.qpc = get_qpc(s, s->mb.qpy, is_chroma_cb);
  s->mb.qp1c = s->mb.qpc + 0;
}

/* ---- frame/transform.rs:116-191 (8.5.12) ------------------------------------------------------ */
static void scaling_and_transform4x4(Ctx* s, isz c[4][4], int is_luma, int is_chroma_cb,
                                     isz r[4][4]) {
  chroma_quantization_parameters(s, is_chroma_cb);
  isz q_p = is_luma ? s->mb.qp1y : s->mb.qp1c;
  isz d[4][4], f[4][4], h[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      if ((s->mb.mode == MODE_I16x16 || !is_luma) && j == 0 && i == 0) {
        d[0][0] = c[0][0];
      } else if (q_p >= 24) {
        d[i][j] = (c[i][j] * s->level_scale4x4[q_p % 6][i][j]) << (q_p / 6 - 4);
      } else {
        d[i][j] = (c[i][j] * s->level_scale4x4[q_p % 6][i][j] + ((isz)1 << (3 - q_p / 6))) >>
                  (4 - q_p / 6);
      }
    }
  for (int i = 0; i < 4; i++) {
    isz e0 = d[i][0] + d[i][2];
    isz e1 = d[i][0] - d[i][2];
    isz e2 = (d[i][1] >> 1) - d[i][3];
    isz e3 = d[i][1] + (d[i][3] >> 1);
    f[i][0] = e0 + e3;
    f[i][1] = e1 + e2;
    f[i][2] = e1 - e2;
    f[i][3] = e0 - e3;
  }
  for (int j = 0; j < 4; j++) {
    isz g0 = f[0][j] + f[2][j];
    isz g1 = f[0][j] - f[2][j];
    isz g2 = (f[1][j] >> 1) - f[3][j];
    isz g3 = f[1][j] + (f[3][j] >> 1);
    h[0][j] = g0 + g3;
    h[1][j] = g1 + g2;
    h[2][j] = g1 - g2;
    h[3][j] = g0 - g3;
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r[i][j] = (h[i][j] + 32) >> 6;
}

/* ---- frame/pred8x8.rs:51-150 (8.5.13) --------------------------------------------------------- */
static void idct8_1d(const isz d[8], isz g[8]) {
  isz e0 = d[0] + d[4];
  isz e1 = -d[3] + d[5] - d[7] - (d[7] >> 1);
  isz e2 = d[0] - d[4];
  isz e3 = d[1] + d[7] - d[3] - (d[3] >> 1);
  isz e4 = (d[2] >> 1) - d[6];
  isz e5 = -d[1] + d[7] + d[5] + (d[5] >> 1);
  isz e6 = d[2] + (d[6] >> 1);
  isz e7 = d[3] + d[5] + d[1] + (d[1] >> 1);
  isz f0 = e0 + e6;
  isz f1 = e1 + (e7 >> 2);
  isz f2 = e2 + e4;
  isz f3 = e3 + (e5 >> 2);
  isz f4 = e2 - e4;
  isz f5 = (e3 >> 2) - e5;
  isz f6 = e0 - e6;
  isz f7 = e7 - (e1 >> 2);
  g[0] = f0 + f7;
  g[1] = f2 + f5;
  g[2] = f4 + f3;
  g[3] = f6 + f1;
  g[4] = f6 - f1;
  g[5] = f4 - f3;
  g[6] = f2 - f5;
  g[7] = f0 - f7;
}
static void scaling_and_transform8x8(Ctx* s, isz c[8][8], isz r[8][8]) {
  chroma_quantization_parameters(s, 0);
  isz q_p = s->mb.qp1y; /* is_luma is always true on this path */
  isz d[8][8], g[8][8], m[8][8];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) {
      if (q_p >= 36)
        d[i][j] = (c[i][j] * s->level_scale8x8[q_p % 6][i][j]) << (q_p / 6 - 6);
      else
        d[i][j] = (c[i][j] * s->level_scale8x8[q_p % 6][i][j] + ((isz)1 << (5 - q_p / 6))) >>
                  (6 - q_p / 6);
    }
  for (int i = 0; i < 8; i++) idct8_1d(d[i], g[i]);
  for (int j = 0; j < 8; j++) {
    isz col[8], out[8];
    for (int i = 0; i < 8; i++) col[i] = g[i][j];
    idct8_1d(col, out);
    for (int i = 0; i < 8; i++) m[i][j] = out[i];
  }
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) r[i][j] = (m[i][j] + 32) >> 6;
}

/* ---- frame/pred16x16.rs:428-482 (8.5.10) ------------------------------------------------------ */
static void transform_intra16x16_dc(Ctx* s, isz c[4][4], isz dc_y[4][4]) {
  static const isz A[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
  isz q_p = s->mb.qp1y;
  isz g[4][4] = {{0}}, f[4][4] = {{0}};
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++) g[i][j] += A[i][k] * c[k][j];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      for (int k = 0; k < 4; k++) f[i][j] += g[i][k] * A[k][j];
  isz ls = s->level_scale4x4[q_p % 6][0][0];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      if (q_p >= 36) dc_y[i][j] = (f[i][j] * ls) << (q_p / 6 - 6);
      else dc_y[i][j] = (f[i][j] * ls + ((isz)1 << (5 - q_p / 6))) >> (6 - q_p / 6);
    }
}

/* ---- frame/trans_chroma.rs:369-456 (8.5.11.1), ChromaArrayType 1 ------------------------------ */
static void transform_chroma_dc(Ctx* s, isz c[2][2], int is_chroma_cb, isz dc_c[2][2]) {
  static const isz a[2][2] = {{1, 1}, {1, -1}};
  chroma_quantization_parameters(s, is_chroma_cb);
  isz q_p = s->mb.qp1c;
  isz g[2][2] = {{0}}, f[2][2] = {{0}};
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) g[i][j] += a[i][k] * c[k][j];
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      for (int k = 0; k < 2; k++) f[i][j] += g[i][k] * a[k][j];
  /* Q1: the luma LevelScale4x4 left behind by the luma driver is used (no chroma scaling() call) */
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++)
      dc_c[i][j] = ((f[i][j] * s->level_scale4x4[q_p % 6][0][0]) << (q_p / 6)) >> 5;
}

/* ---- frame/mod.rs:93-165 (8.5.14) -------------------------------------------------------------- */
enum { B16x16, B8x8, B4x4 };
static void picture_construction(Ctx* s, const isz* u, int blk_type, isz blk_idx, int is_luma,
                                 int is_chroma_cb) {
  isz x_p = inverse_raster_scan(s->curr_mb_addr, 16, 16, s->width_l, 0);
  isz y_p = inverse_raster_scan(s->curr_mb_addr, 16, 16, s->width_l, 1);
  isz x_o = 0, y_o = 0;
  if (is_luma) {
    isz n_e;
    if (blk_type == B16x16) {
      n_e = 16;
    } else if (blk_type == B4x4) {
      x_o = inverse_raster_scan(blk_idx / 4, 8, 8, 16, 0) + inverse_raster_scan(blk_idx % 4, 4, 4, 8, 0);
      y_o = inverse_raster_scan(blk_idx / 4, 8, 8, 16, 1) + inverse_raster_scan(blk_idx % 4, 4, 4, 8, 1);
      n_e = 4;
    } else {
      x_o = inverse_raster_scan(blk_idx, 8, 8, 16, 0);
      y_o = inverse_raster_scan(blk_idx, 8, 8, 16, 1);
      n_e = 8;
    }
    for (isz i = 0; i < n_e; i++)
      for (isz j = 0; j < n_e; j++) LUMA(s, x_p + x_o + j, y_p + y_o + i) = (uint8_t)u[i * n_e + j];
  } else {
    for (isz i = 0; i < 8; i++)
      for (isz j = 0; j < 8; j++) {
        isz x = x_p / 2 + x_o + j, y = y_p / 2 + y_o + i;
        if (is_chroma_cb) CB(s, x, y) = (uint8_t)u[i * 8 + j];
        else CR(s, x, y) = (uint8_t)u[i * 8 + j];
      }
  }
}

/* Neighbour luma sample fetch shared by the three luma predictors: from_coords -> mb_nb_p ->
 * coords -> plane read, or -1 (pred4x4.rs:33-63, pred8x8.rs:183-199, pred16x16.rs:104-133). */
static isz fetch_luma(const Ctx* s, isz x_n, isz y_n) {
  int pos = from_coords(x_n, y_n, 16, 16);
  isz mbaddr_n = pos == POS_NONE ? -1 : mb_nb_p(s, pos);
  if (mbaddr_n < 0) return -1;
  isz xw, yw;
  pos_coords(x_n, y_n, 16, 16, &xw, &yw);
  isz x_m = inverse_raster_scan(mbaddr_n, 16, 16, s->width_l, 0);
  isz y_m = inverse_raster_scan(mbaddr_n, 16, 16, s->width_l, 1);
  return LUMA(s, x_m + xw, y_m + yw);
}

/* ---- frame/pred4x4.rs:363-427 (8.3.1.1) -------------------------------------------------------- */
static void intra4x4_pred_mode(Ctx* s, int blk) {
  isz x = inverse_raster_scan(blk / 4, 8, 8, 16, 0) + inverse_raster_scan(blk % 4, 4, 4, 8, 0);
  isz y = inverse_raster_scan(blk / 4, 8, 8, 16, 1) + inverse_raster_scan(blk % 4, 4, 4, 8, 1);
  int pa = from_coords(x - 1, y, 16, 16), pb = from_coords(x, y - 1, 16, 16);
  isz addr_a = pa == POS_NONE ? -1 : mb_nb_p(s, pa);
  isz addr_b = pb == POS_NONE ? -1 : mb_nb_p(s, pb);
  isz idx_a = pos_blk_idx4x4(x - 1, y, 16, 16), idx_b = pos_blk_idx4x4(x, y - 1, 16, 16);
  int dc_pred_mode_predicted_flag = addr_a < 0 || addr_b < 0;
  isz mode_a, mode_b;
  /* the current MB's own entry is read through s->mb (slice.mb()), others through macroblocks[] */
  const NbMb* mb_a = addr_a < 0 ? NULL : &s->macroblocks[addr_a];
  const NbMb* mb_b = addr_b < 0 ? NULL : &s->macroblocks[addr_b];
  if (dc_pred_mode_predicted_flag || (mb_a->mode != MODE_I4x4 && mb_a->mode != MODE_I8x8)) mode_a = 2;
  else if (mb_a->mode == MODE_I4x4) mode_a = mb_a->intra4x4_pred_mode[idx_a];
  else mode_a = mb_a->intra8x8_pred_mode[idx_a >> 2];
  if (dc_pred_mode_predicted_flag || (mb_b->mode != MODE_I4x4 && mb_b->mode != MODE_I8x8)) mode_b = 2;
  else if (mb_b->mode == MODE_I4x4) mode_b = mb_b->intra4x4_pred_mode[idx_b];
  else mode_b = mb_b->intra8x8_pred_mode[idx_b >> 2];
  isz pred = mode_a < mode_b ? mode_a : mode_b;
  NbMb* me = &s->macroblocks[s->curr_mb_addr];
  if (s->mb.prev_intra4x4_pred_mode_flag[blk] != 0) me->intra4x4_pred_mode[blk] = pred;
  else if ((isz)s->mb.rem_intra4x4_pred_mode[blk] < pred) me->intra4x4_pred_mode[blk] = s->mb.rem_intra4x4_pred_mode[blk];
  else me->intra4x4_pred_mode[blk] = (isz)s->mb.rem_intra4x4_pred_mode[blk] + 1;
}

/* ---- frame/pred4x4.rs:10-360 (8.3.1.2) --------------------------------------------------------- */
#define P4(x, y) smp[((y) + 1) * 9 + ((x) + 1)]
static void intra4x4_prediction(Ctx* s, int blk) {
  static const isz RX[13] = {-1, -1, -1, -1, -1, 0, 1, 2, 3, 4, 5, 6, 7};
  static const isz RY[13] = {-1, 0, 1, 2, 3, -1, -1, -1, -1, -1, -1, -1, -1};
  isz x_o = inverse_raster_scan(blk / 4, 8, 8, 16, 0) + inverse_raster_scan(blk % 4, 4, 4, 8, 0);
  isz y_o = inverse_raster_scan(blk / 4, 8, 8, 16, 1) + inverse_raster_scan(blk % 4, 4, 4, 8, 1);
  isz smp[45];
  for (int i = 0; i < 45; i++) smp[i] = -1;
  for (int i = 0; i < 13; i++) {
    isz x = RX[i], y = RY[i];
    isz v = fetch_luma(s, x_o + x, y_o + y);
    if ((x > 3) && (blk == 3 || blk == 11)) v = -1; /* pred4x4.rs:42 */
    P4(x, y) = v;
  }
  if (P4(4, -1) < 0 && P4(5, -1) < 0 && P4(6, -1) < 0 && P4(7, -1) < 0 && P4(3, -1) >= 0) {
    P4(4, -1) = P4(3, -1);
    P4(5, -1) = P4(3, -1);
    P4(6, -1) = P4(3, -1);
    P4(7, -1) = P4(3, -1);
  }
  intra4x4_pred_mode(s, blk);
  isz mode = s->macroblocks[s->curr_mb_addr].intra4x4_pred_mode[blk];
  isz(*out)[4] = s->mb.luma_pred_samples[blk]; /* [x][y] */
  int top = P4(0, -1) >= 0 && P4(1, -1) >= 0 && P4(2, -1) >= 0 && P4(3, -1) >= 0;
  int left = P4(-1, 0) >= 0 && P4(-1, 1) >= 0 && P4(-1, 2) >= 0 && P4(-1, 3) >= 0;
  int topright = P4(4, -1) >= 0 && P4(5, -1) >= 0 && P4(6, -1) >= 0 && P4(7, -1) >= 0;
  int corner = P4(-1, -1) >= 0;
  if (mode == 0) { /* Vertical */
    if (top)
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) out[x][y] = P4(x, -1);
  } else if (mode == 1) { /* Horizontal */
    if (left)
      for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++) out[x][y] = P4(-1, y);
  } else if (mode == 2) { /* DC */
    isz val;
    if (top && left)
      val = (P4(0, -1) + P4(1, -1) + P4(2, -1) + P4(3, -1) + P4(-1, 0) + P4(-1, 1) + P4(-1, 2) + P4(-1, 3) + 4) >> 3;
    else if (!top && left) val = (P4(-1, 0) + P4(-1, 1) + P4(-1, 2) + P4(-1, 3) + 2) >> 2;
    else if (top && !left) val = (P4(0, -1) + P4(1, -1) + P4(2, -1) + P4(3, -1) + 2) >> 2;
    else val = 128;
    for (int x = 0; x < 4; x++)
      for (int y = 0; y < 4; y++) out[x][y] = val;
  } else if (mode == 3) { /* Diagonal down left */
    if (top && topright)
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) {
          if (x == 3 && y == 3) out[x][y] = (P4(6, -1) + 3 * P4(7, -1) + 2) >> 2;
          else out[x][y] = (P4(x + y, -1) + 2 * P4(x + y + 1, -1) + P4(x + y + 2, -1) + 2) >> 2;
        }
  } else if (mode == 4) { /* Diagonal down right */
    if (top && corner && left)
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) {
          if (x > y) out[x][y] = (P4(x - y - 2, -1) + 2 * P4(x - y - 1, -1) + P4(x - y, -1) + 2) >> 2;
          else if (x < y) out[x][y] = (P4(-1, y - x - 2) + 2 * P4(-1, y - x - 1) + P4(-1, y - x) + 2) >> 2;
          else out[x][y] = (P4(0, -1) + 2 * P4(-1, -1) + P4(-1, 0) + 2) >> 2;
        }
  } else if (mode == 5) { /* Vertical right */
    if (top && corner && left)
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) {
          isz z = 2 * x - y;
          if (z == 0 || z == 2 || z == 4 || z == 6)
            out[x][y] = (P4(x - (y >> 1) - 1, -1) + P4(x - (y >> 1), -1) + 1) >> 1;
          else if (z == 1 || z == 3 || z == 5)
            out[x][y] = (P4(x - (y >> 1) - 2, -1) + 2 * P4(x - (y >> 1) - 1, -1) + P4(x - (y >> 1), -1) + 2) >> 2;
          else if (z == -1) out[x][y] = (P4(-1, 0) + 2 * P4(-1, -1) + P4(0, -1) + 2) >> 2;
          else out[x][y] = (P4(-1, y - 1) + 2 * P4(-1, y - 2) + P4(-1, y - 3) + 2) >> 2;
        }
  } else if (mode == 6) { /* Horizontal down */
    if (top && corner && left)
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) {
          isz z = 2 * y - x;
          if (z == 0 || z == 2 || z == 4 || z == 6)
            out[x][y] = (P4(-1, y - (x >> 1) - 1) + P4(-1, y - (x >> 1)) + 1) >> 1;
          else if (z == 1 || z == 3 || z == 5)
            out[x][y] = (P4(-1, y - (x >> 1) - 2) + 2 * P4(-1, y - (x >> 1) - 1) + P4(-1, y - (x >> 1)) + 2) >> 2;
          else if (z == -1) out[x][y] = (P4(-1, 0) + 2 * P4(-1, -1) + P4(0, -1) + 2) >> 2;
          else out[x][y] = (P4(x - 1, -1) + 2 * P4(x - 2, -1) + P4(x - 3, -1) + 2) >> 2;
        }
  } else if (mode == 7) { /* Vertical left */
    if (top && topright)
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) {
          if (y == 0 || y == 2) out[x][y] = (P4(x + (y >> 1), -1) + P4(x + (y >> 1) + 1, -1) + 1) >> 1;
          else out[x][y] = (P4(x + (y >> 1), -1) + 2 * P4(x + (y >> 1) + 1, -1) + P4(x + (y >> 1) + 2, -1) + 2) >> 2;
        }
  } else if (mode == 8 && left) { /* Horizontal up */
    for (isz y = 0; y < 4; y++)
      for (isz x = 0; x < 4; x++) {
        isz z = x + 2 * y;
        if (z == 0 || z == 2 || z == 4) out[x][y] = (P4(-1, y + (x >> 1)) + P4(-1, y + (x >> 1) + 1) + 1) >> 1;
        else if (z == 1 || z == 3)
          out[x][y] = (P4(-1, y + (x >> 1)) + 2 * P4(-1, y + (x >> 1) + 1) + P4(-1, y + (x >> 1) + 2) + 2) >> 2;
        else if (z == 5) out[x][y] = (P4(-1, 2) + 3 * P4(-1, 3) + 2) >> 2;
        else out[x][y] = P4(-1, 3);
      }
  }
  /* any other case: nothing written, prediction stays 0 (quirk Q4) */
}
#undef P4

/* ---- frame/transform.rs:81-113 (8.5.1) --------------------------------------------------------- */
static void transform_for_4x4_luma_residual_blocks(Ctx* s) {
  scaling(s);
  for (int blk = 0; blk < 16; blk++) {
    isz c[4][4], r[4][4], u[16];
    inverse_scanner4x4(s->mb.block_luma_4x4[blk], c);
    scaling_and_transform4x4(s, c, 1, 0, r);
    intra4x4_prediction(s, blk);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) u[i * 4 + j] = clampz(s->mb.luma_pred_samples[blk][j][i] + r[i][j], 0, 255);
    picture_construction(s, u, B4x4, blk, 1, 0);
  }
}

/* ---- frame/pred8x8.rs:698-764 (8.3.2.1) -------------------------------------------------------- */
static void intra8x8_pred_mode(Ctx* s, int blk) {
  isz x = (blk % 2) * 8, y = (blk / 2) * 8;
  int pa = from_coords(x - 1, y, 16, 16), pb = from_coords(x, y - 1, 16, 16);
  isz addr_a = pa == POS_NONE ? -1 : mb_nb_p(s, pa);
  isz addr_b = pb == POS_NONE ? -1 : mb_nb_p(s, pb);
  isz idx_a = pos_blk_idx8x8(x - 1, y, 16, 16), idx_b = pos_blk_idx8x8(x, y - 1, 16, 16);
  int dc_flag = addr_a < 0 || addr_b < 0;
  const NbMb* mb_a = addr_a < 0 ? NULL : &s->macroblocks[addr_a];
  const NbMb* mb_b = addr_b < 0 ? NULL : &s->macroblocks[addr_b];
  isz mode_a, mode_b;
  if (dc_flag || (mb_a->mode != MODE_I4x4 && mb_a->mode != MODE_I8x8)) mode_a = 2;
  else if (mb_a->mode == MODE_I8x8) mode_a = mb_a->intra8x8_pred_mode[idx_a];
  else mode_a = mb_a->intra4x4_pred_mode[idx_a * 4 + 1];
  if (dc_flag || (mb_b->mode != MODE_I4x4 && mb_b->mode != MODE_I8x8)) mode_b = 2;
  else if (mb_b->mode == MODE_I8x8) mode_b = mb_b->intra8x8_pred_mode[idx_b];
  else mode_b = mb_b->intra4x4_pred_mode[idx_b * 4 + 2];
  isz pred = mode_a < mode_b ? mode_a : mode_b;
  NbMb* me = &s->macroblocks[s->curr_mb_addr];
  if (s->mb.prev_intra8x8_pred_mode_flag[blk] != 0) me->intra8x8_pred_mode[blk] = pred;
  else if ((isz)s->mb.rem_intra8x8_pred_mode[blk] < pred) me->intra8x8_pred_mode[blk] = s->mb.rem_intra8x8_pred_mode[blk];
  else me->intra8x8_pred_mode[blk] = (isz)s->mb.rem_intra8x8_pred_mode[blk] + 1;
}

/* ---- frame/pred8x8.rs:152-696 (8.3.2.2) -------------------------------------------------------- */
#define P8(a, x, y) (a)[((y) + 1) * 17 + ((x) + 1)]
static int all_ge0_row(const isz* a, int x0, int x1) {
  for (int x = x0; x <= x1; x++)
    if (P8(a, x, -1) < 0) return 0;
  return 1;
}
static int all_ge0_col(const isz* a, int y0, int y1) {
  for (int y = y0; y <= y1; y++)
    if (P8(a, -1, y) < 0) return 0;
  return 1;
}
static int intra8x8_prediction(Ctx* s, int blk) {
  isz p[9 * 17], p1[9 * 17];
  for (int i = 0; i < 9 * 17; i++) p[i] = p1[i] = -1;
  isz x_o = inverse_raster_scan(blk, 8, 8, 16, 0), y_o = inverse_raster_scan(blk, 8, 8, 16, 1);
  for (int i = 0; i < 25; i++) {
    isz x = i < 9 ? -1 : i - 9, y = i < 9 ? i - 1 : -1;
    P8(p, x, y) = fetch_luma(s, x_o + x, y_o + y);
  }
  /* top-right replication, pred8x8.rs:202-220 */
  {
    int none = 1;
    for (int x = 8; x < 16; x++)
      if (P8(p, x, -1) >= 0) none = 0;
    if (none && P8(p, 7, -1) >= 0)
      for (int x = 8; x < 16; x++) P8(p, x, -1) = P8(p, 7, -1);
  }
  /* reference sample filtering, pred8x8.rs:222-288 */
  if (all_ge0_row(p, 0, 15)) {
    if (P8(p, -1, -1) >= 0) P8(p1, 0, -1) = (P8(p, -1, -1) + 2 * P8(p, 0, -1) + P8(p, 1, -1) + 2) >> 2;
    else P8(p1, 0, -1) = (3 * P8(p, 0, -1) + P8(p, 1, -1) + 2) >> 2;
    /* Q2: the loop starts at x = 0 and overwrites the line above with the raw corner (may be -1) */
    for (int x = 0; x < 15; x++) P8(p1, x, -1) = (P8(p, x - 1, -1) + 2 * P8(p, x, -1) + P8(p, x + 1, -1) + 2) >> 2;
    P8(p1, 15, -1) = (P8(p, 14, -1) + 3 * P8(p, 15, -1) + 2) >> 2;
  }
  if (P8(p, -1, -1) >= 0) {
    if (P8(p, 0, -1) < 0 || P8(p, -1, 0) < 0) {
      if (P8(p, 0, -1) >= 0) P8(p1, -1, -1) = (3 * P8(p, -1, -1) + P8(p, 0, -1) + 2) >> 2;
      else if (P8(p, 0, -1) < 0 && P8(p, -1, 0) >= 0) P8(p1, -1, -1) = (3 * P8(p, -1, -1) + P8(p, -1, 0) + 2) >> 2;
      else P8(p1, -1, -1) = P8(p, -1, -1);
    } else {
      P8(p1, -1, -1) = (P8(p, 0, -1) + 2 * P8(p, -1, -1) + P8(p, -1, 0) + 2) >> 2;
    }
  }
  if (all_ge0_col(p, 0, 7)) {
    if (P8(p, -1, -1) >= 0) P8(p1, -1, 0) = (P8(p, -1, -1) + 2 * P8(p, -1, 0) + P8(p, -1, 1) + 2) >> 2;
    else P8(p1, -1, 0) = (3 * P8(p, -1, 0) + P8(p, -1, 1) + 2) >> 2;
    for (int y = 1; y < 7; y++) P8(p1, -1, y) = (P8(p, -1, y - 1) + 2 * P8(p, -1, y) + P8(p, -1, y + 1) + 2) >> 2;
    P8(p1, -1, 7) = (P8(p, -1, 6) + 3 * P8(p, -1, 7) + 2) >> 2;
  }
  memcpy(p, p1, sizeof p);

  intra8x8_pred_mode(s, blk);
  isz mode = s->macroblocks[s->curr_mb_addr].intra8x8_pred_mode[blk];
  isz(*out)[8] = s->mb.luma8x8_pred_samples[blk]; /* [x][y] */
  int top = all_ge0_row(p, 0, 7), topright = all_ge0_row(p, 8, 15), left = all_ge0_col(p, 0, 7);
  int corner = P8(p, -1, -1) >= 0;
  if (mode == 0) {
    if (top)
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) out[x][y] = P8(p, x, -1);
  } else if (mode == 1) {
    if (left)
      for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++) out[x][y] = P8(p, -1, y);
  } else if (mode == 2) {
    isz st = 0, sl = 0, val;
    for (int k = 0; k < 8; k++) {
      st += P8(p, k, -1);
      sl += P8(p, -1, k);
    }
    if (top && left) val = (st + sl + 8) >> 4;
    else if (!top && left) val = (sl + 4) >> 3;
    else if (top && !left) val = (st + 4) >> 3;
    else val = 128;
    for (int y = 0; y < 8; y++)
      for (int x = 0; x < 8; x++) out[x][y] = val;
  } else if (mode == 3) {
    if (top && topright)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          if (x == 7 && y == 7) out[x][y] = (P8(p, 14, -1) + 3 * P8(p, 15, -1) + 2) >> 2;
          else out[x][y] = (P8(p, x + y, -1) + 2 * P8(p, x + y + 1, -1) + P8(p, x + y + 2, -1) + 2) >> 2;
        }
  } else if (mode == 4) {
    if (top && corner && left)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          if (x > y) out[x][y] = (P8(p, x - y - 2, -1) + 2 * P8(p, x - y - 1, -1) + P8(p, x - y, -1) + 2) >> 2;
          else if (x < y) out[x][y] = (P8(p, -1, y - x - 2) + 2 * P8(p, -1, y - x - 1) + P8(p, -1, y - x) + 2) >> 2;
          else out[x][y] = (P8(p, 0, -1) + 2 * P8(p, -1, -1) + P8(p, -1, 0) + 2) >> 2;
        }
  } else if (mode == 5) {
    if (top && corner && left)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          isz z = 2 * x - y;
          if (z >= 0 && z <= 14 && (z % 2) == 0)
            out[x][y] = (P8(p, x - (y >> 1) - 1, -1) + P8(p, x - (y >> 1), -1) + 1) >> 1;
          else if (z >= 1 && z <= 13)
            out[x][y] = (P8(p, x - (y >> 1) - 2, -1) + 2 * P8(p, x - (y >> 1) - 1, -1) + P8(p, x - (y >> 1), -1) + 2) >> 2;
          else if (z == -1) out[x][y] = (P8(p, -1, 0) + 2 * P8(p, -1, -1) + P8(p, 0, -1) + 2) >> 2;
          else out[x][y] = (P8(p, -1, y - 2 * x - 1) + 2 * P8(p, -1, y - 2 * x - 2) + P8(p, -1, y - 2 * x - 3) + 2) >> 2;
        }
  } else if (mode == 6) {
    if (top && corner && left)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          isz z = 2 * y - x;
          if (z >= 0 && z <= 14 && (z % 2) == 0)
            out[x][y] = (P8(p, -1, y - (x >> 1) - 1) + P8(p, -1, y - (x >> 1)) + 1) >> 1;
          else if (z >= 1 && z <= 13)
            out[x][y] = (P8(p, -1, y - (x >> 1) - 2) + 2 * P8(p, -1, y - (x >> 1) - 1) + P8(p, -1, y - (x >> 1)) + 2) >> 2;
          else if (z == -1) out[x][y] = (P8(p, -1, 0) + 2 * P8(p, -1, -1) + P8(p, 0, -1) + 2) >> 2;
          else out[x][y] = (P8(p, x - 2 * y - 1, -1) + 2 * P8(p, x - 2 * y - 2, -1) + P8(p, x - 2 * y - 3, -1) + 2) >> 2;
        }
  } else if (mode == 7) {
    if (top && topright)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          if ((y % 2) == 0) out[x][y] = (P8(p, x + (y >> 1), -1) + P8(p, x + (y >> 1) + 1, -1) + 1) >> 1;
          else out[x][y] = (P8(p, x + (y >> 1), -1) + 2 * P8(p, x + (y >> 1) + 1, -1) + P8(p, x + (y >> 1) + 2, -1) + 2) >> 2;
        }
  } else if (mode == 8) {
    if (left)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) {
          isz z = x + 2 * y;
          if (z <= 12 && (z % 2) == 0) out[x][y] = (P8(p, -1, y + (x >> 1)) + P8(p, -1, y + (x >> 1) + 1) + 1) >> 1;
          else if (z <= 11)
            out[x][y] = (P8(p, -1, y + (x >> 1)) + 2 * P8(p, -1, y + (x >> 1) + 1) + P8(p, -1, y + (x >> 1) + 2) + 2) >> 2;
          else if (z == 13) out[x][y] = (P8(p, -1, 6) + 3 * P8(p, -1, 7) + 2) >> 2;
          else out[x][y] = P8(p, -1, 7);
        }
  } else {
    return -1; /* pred8x8.rs:693-695 panics */
  }
  return 0;
}

/* ---- frame/pred8x8.rs:17-48 -------------------------------------------------------------------- */
static int transform_for_8x8_luma_residual_blocks(Ctx* s) {
  scaling(s);
  for (int blk = 0; blk < 4; blk++) {
    isz c[8][8], r[8][8], u[64];
    inverse_scanner_8x8(s->mb.block_luma_8x8[blk], c);
    scaling_and_transform8x8(s, c, r);
    if (intra8x8_prediction(s, blk) != 0) return -1;
    for (int i = 0; i < 8; i++)
      for (int j = 0; j < 8; j++) u[i * 8 + j] = clampz(s->mb.luma8x8_pred_samples[blk][j][i] + r[i][j], 0, 255);
    picture_construction(s, u, B8x8, blk, 1, 0);
  }
  return 0;
}

/* ---- frame/pred16x16.rs:79-425 (8.3.3) --------------------------------------------------------- */
static void intra16x16_prediction(Ctx* s) {
  isz p[17 * 17];
  for (int i = 0; i < 17 * 17; i++) p[i] = -1;
  for (int i = 0; i < 33; i++) {
    isz x = i < 17 ? -1 : i - 17, y = i < 17 ? i - 1 : -1;
    P8(p, x, y) = fetch_luma(s, x, y);
  }
  int top = all_ge0_row(p, 0, 15), left = all_ge0_col(p, 0, 15);
  isz(*out)[16] = s->mb.luma16x16_pred_samples; /* [x][y] */
  int mode = s->mb.intra16x16_pred_mode;
  if (mode == 0) {
    if (top)
      for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) out[x][y] = P8(p, x, -1);
  } else if (mode == 1) {
    if (left)
      for (int y = 0; y < 16; y++)
        for (int x = 0; x < 16; x++) out[x][y] = P8(p, -1, y);
  } else if (mode == 2) {
    isz st = 0, sl = 0, val;
    for (int k = 0; k < 16; k++) {
      st += P8(p, k, -1);
      sl += P8(p, -1, k);
    }
    if (top && left) val = (st + sl + 16) >> 5;
    else if (!top && left) val = (sl + 8) >> 4;
    else if (top && !left) val = (st + 8) >> 4;
    else val = 128;
    for (int x = 0; x < 16; x++)
      for (int y = 0; y < 16; y++) out[x][y] = val;
  } else if (mode == 3 && top && left) {
    /* Q5: the corner p[-1,-1] is read (x = 7 below) without having been tested */
    isz h = 0, v = 0;
    for (isz x = 0; x <= 7; x++) h += (x + 1) * (P8(p, 8 + x, -1) - P8(p, 6 - x, -1));
    for (isz y = 0; y <= 7; y++) v += (y + 1) * (P8(p, -1, 8 + y) - P8(p, -1, 6 - y));
    isz a = 16 * (P8(p, -1, 15) + P8(p, 15, -1));
    isz b = (5 * h + 32) >> 6;
    isz c = (5 * v + 32) >> 6;
    for (isz y = 0; y < 16; y++)
      for (isz x = 0; x < 16; x++) out[x][y] = clampz((a + b * (x - 7) + c * (y - 7) + 16) >> 5, 0, 255);
  }
}
#undef P8

/* ---- frame/pred16x16.rs:13-76 (8.5.2) ---------------------------------------------------------- */
static void transform_for_16x16_luma_residual_blocks(Ctx* s) {
  scaling(s);
  isz c[4][4], dc_y[4][4];
  inverse_scanner4x4(s->mb.block_luma_dc, c);
  transform_intra16x16_dc(s, c, dc_y);
  isz r_mb[16][16]; /* [x][y] */
  const isz dc_y_to_luma[16] = {dc_y[0][0], dc_y[0][1], dc_y[1][0], dc_y[1][1], dc_y[0][2], dc_y[0][3],
                                dc_y[1][2], dc_y[1][3], dc_y[2][0], dc_y[2][1], dc_y[3][0], dc_y[3][1],
                                dc_y[2][2], dc_y[2][3], dc_y[3][2], dc_y[3][3]};
  for (int blk = 0; blk < 16; blk++) {
    isz luma_list[16], cc[4][4], r[4][4];
    luma_list[0] = dc_y_to_luma[blk];
    for (int k = 0; k < 15; k++) luma_list[1 + k] = s->mb.block_luma_ac[blk][k];
    inverse_scanner4x4(luma_list, cc);
    scaling_and_transform4x4(s, cc, 1, 0, r);
    isz x_o = inverse_raster_scan(blk / 4, 8, 8, 16, 0) + inverse_raster_scan(blk % 4, 4, 4, 8, 0);
    isz y_o = inverse_raster_scan(blk / 4, 8, 8, 16, 1) + inverse_raster_scan(blk % 4, 4, 4, 8, 1);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) r_mb[x_o + j][y_o + i] = r[i][j];
  }
  intra16x16_prediction(s);
  isz u[256];
  for (int i = 0; i < 16; i++)
    for (int j = 0; j < 16; j++) u[i * 16 + j] = clampz(s->mb.luma16x16_pred_samples[j][i] + r_mb[j][i], 0, 255);
  picture_construction(s, u, B16x16, 0, 1, 0);
}

/* ---- frame/trans_chroma.rs:96-366 (8.3.4), ChromaArrayType 1 ----------------------------------- */
#define PC(x, y) smp[((y) + 1) * 9 + ((x) + 1)]
static void intra_chroma_prediction(Ctx* s, int is_chroma_cb) {
  isz smp[81];
  for (int i = 0; i < 81; i++) smp[i] = -1;
  for (int i = 0; i < 17; i++) {
    isz x = i < 9 ? -1 : i - 9, y = i < 9 ? i - 1 : -1;
    int pos = from_coords(x, y, 8, 8);
    isz mbaddr_n = pos == POS_NONE ? -1 : mb_nb_p(s, pos);
    if (mbaddr_n < 0) {
      PC(x, y) = -1;
    } else {
      isz xw, yw;
      pos_coords(x, y, 8, 8, &xw, &yw);
      isz x_l = inverse_raster_scan(mbaddr_n, 16, 16, s->width_l, 0);
      isz y_l = inverse_raster_scan(mbaddr_n, 16, 16, s->width_l, 1);
      isz x_m = (x_l >> 4) * 8;
      isz y_m = ((y_l >> 4) * 8) + (y_l % 2);
      PC(x, y) = is_chroma_cb ? CB(s, x_m + xw, y_m + yw) : CR(s, x_m + xw, y_m + yw);
    }
  }
  isz(*out)[16] = s->mb.chroma_pred_samples; /* [x][y] */
  int mode = s->mb.intra_chroma_pred_mode;
  if (mode == 0) {
    for (isz blk = 0; blk < 4; blk++) {
      isz x_o = inverse_raster_scan(blk, 4, 4, 8, 0), y_o = inverse_raster_scan(blk, 4, 4, 8, 1);
      isz t0 = PC(x_o, -1), t1 = PC(1 + x_o, -1), t2 = PC(2 + x_o, -1), t3 = PC(3 + x_o, -1);
      isz l0 = PC(-1, y_o), l1 = PC(-1, 1 + y_o), l2 = PC(-1, 2 + y_o), l3 = PC(-1, 3 + y_o);
      isz val = 0;
      if ((x_o == 0 && y_o == 0) || (x_o > 0 && y_o > 0)) {
        int t_ge = t0 >= 0 && t1 >= 0 && t2 >= 0 && t3 >= 0;
        int l_ge = l0 >= 0 && l1 >= 0 && l2 >= 0 && l3 >= 0;
        int t_gt = t0 > 0 && t1 > 0 && t2 > 0 && t3 > 0; /* Q3 */
        int l_gt = l0 > 0 && l1 > 0 && l2 > 0 && l3 > 0; /* Q3 */
        if (t_ge && l_ge) val = (t0 + t1 + t2 + t3 + l0 + l1 + l2 + l3 + 4) >> 3;
        else if (!t_ge && l_ge) val = (l0 + l1 + l2 + l3 + 2) >> 2;
        else if (t_gt && !l_gt) val = (t0 + t1 + t2 + t3 + 2) >> 2;
        else val = 128;
      } else if (x_o > 0 && y_o == 0) {
        if (t0 >= 0 && t1 >= 0 && t2 >= 0 && t3 >= 0) val = (t0 + t1 + t2 + t3 + 2) >> 2;
        else if (l0 >= 0 && l1 >= 0 && l2 >= 0 && l3 > 0) val = (l0 + l1 + l2 + l3 + 2) >> 2; /* Q3 */
        else val = 128;
      } else if (x_o == 0 && y_o > 0) {
        if (l0 >= 0 && l1 >= 0 && l2 >= 0 && l3 > 0) val = (l0 + l1 + l2 + l3 + 2) >> 2;      /* Q3 */
        else if (t0 >= 0 && t1 >= 0 && t2 >= 0 && t3 > 0) val = (t0 + t1 + t2 + t3 + 2) >> 2; /* Q3 */
        else val = 128;
      }
      for (isz y = 0; y < 4; y++)
        for (isz x = 0; x < 4; x++) out[x + x_o][y + y_o] = val;
    }
  } else if (mode == 1) {
    int flag = 1;
    for (isz y = 0; y < 8; y++)
      if (PC(-1, y) < 0) { flag = 0; break; }
    if (flag)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) out[x][y] = PC(-1, y);
  } else if (mode == 2) {
    int flag = 1;
    for (isz x = 0; x < 8; x++)
      if (PC(x, -1) < 0) { flag = 0; break; }
    if (flag)
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) out[x][y] = PC(x, -1);
  } else if (mode == 3) {
    int flag = 1;
    for (isz x = 0; x < 8; x++)
      if (PC(x, -1) < 0) { flag = 0; break; }
    for (isz y = -1; y < 8; y++)
      if (PC(-1, y) < 0) { flag = 0; break; }
    if (flag) {
      isz h = 0, v = 0;
      for (isz x1 = 0; x1 <= 3; x1++) h += (x1 + 1) * (PC(4 + x1, -1) - PC(2 - x1, -1));
      for (isz y1 = 0; y1 <= 3; y1++) v += (y1 + 1) * (PC(-1, 4 + y1) - PC(-1, 2 - y1));
      isz a = 16 * (PC(-1, 7) + PC(7, -1));
      isz b = (34 * h + 32) >> 6;
      isz c = (34 * v + 32) >> 6;
      for (isz y = 0; y < 8; y++)
        for (isz x = 0; x < 8; x++) out[x][y] = clampz((a + b * (x - 3) + c * (y - 3) + 16) >> 5, 0, 255);
    }
  }
}
#undef PC

/* ---- frame/trans_chroma.rs:14-94 (8.5.4), ChromaArrayType 1 ------------------------------------ */
static void transform_chroma_samples(Ctx* s, int is_chroma_cb) {
  int i_cb_cr = is_chroma_cb ? 0 : 1;
  isz c[2][2], dc_c[2][2];
  c[0][0] = s->mb.block_chroma_dc[i_cb_cr][0];
  c[0][1] = s->mb.block_chroma_dc[i_cb_cr][1];
  c[1][0] = s->mb.block_chroma_dc[i_cb_cr][2];
  c[1][1] = s->mb.block_chroma_dc[i_cb_cr][3];
  transform_chroma_dc(s, c, is_chroma_cb, dc_c);
  const isz dc_cto_chroma[4] = {dc_c[0][0], dc_c[0][1], dc_c[1][0], dc_c[1][1]};
  isz r_mb[8][16]; /* [x][y] */
  for (int blk = 0; blk < 4; blk++) {
    isz chroma_list[16], cc[4][4], r[4][4];
    chroma_list[0] = dc_cto_chroma[blk];
    for (int k = 0; k < 15; k++) chroma_list[1 + k] = s->mb.block_chroma_ac[i_cb_cr][blk][k];
    inverse_scanner4x4(chroma_list, cc);
    scaling_and_transform4x4(s, cc, 0, is_chroma_cb, r);
    isz x_o = inverse_raster_scan(blk, 4, 4, 8, 0), y_o = inverse_raster_scan(blk, 4, 4, 8, 1);
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) r_mb[x_o + j][y_o + i] = r[i][j];
  }
  intra_chroma_prediction(s, is_chroma_cb);
  isz u[64];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) u[i * 8 + j] = clampz(s->mb.chroma_pred_samples[j][i] + r_mb[j][i], 0, 255);
  picture_construction(s, u, B4x4, 0, 0, is_chroma_cb);
}

/* ---- frame/mod.rs:72-90: Frame::decode ---------------------------------------------------------- */
static int frame_decode(Ctx* s) {
  if (s->mb.mode == MODE_I4x4) {
    transform_for_4x4_luma_residual_blocks(s);
  } else if (s->mb.mode == MODE_I8x8) {
    if (transform_for_8x8_luma_residual_blocks(s) != 0) return DRYV_ORACLE_ERR_UNSUPPORTED;
  } else {
    transform_for_16x16_luma_residual_blocks(s);
  }
  transform_chroma_samples(s, 1);
  transform_chroma_samples(s, 0);
  return 0;
}

/* What CABAC would have left in slice.mb() before calling frame.decode (cabac/mod.rs:89-208),
 * taken from the SoA record of macroblock `mbaddr` of one picture. */
static int load_mb(Ctx* s, const dryv_mb_soa* soa, size_t idx) {
  CurMb* mb = &s->mb;
  memset(mb, 0, sizeof *mb); /* Macroblock::empty(), slice/macroblock.rs:156-202 */
  uint8_t code = soa->mb_type[idx];
  if (code > 24) return DRYV_ORACLE_ERR_UNSUPPORTED; /* I_PCM / inter: todo!() frame/mod.rs:85-88 */
  uint8_t t8 = soa->transform_size_8x8_flag[idx];
  mb->code = code;
  if (code == 0) {
    mb->mode = t8 ? MODE_I8x8 : MODE_I4x4; /* slice/macroblock.rs:683-688 */
  } else {
    mb->mode = MODE_I16x16;
    mb->intra16x16_pred_mode = (code - 1) % 4;
  }
  mb->qp1y = soa->qp[idx];
  if (mb->qp1y > 51) return DRYV_ORACLE_ERR_ARG;
  mb->qpy = mb->qp1y; /* QpBdOffsetY = 0 */
  mb->intra_chroma_pred_mode = soa->intra_chroma_pred_mode[idx];
  const uint8_t* ps = soa->pred_syntax + idx * 16;
  for (int k = 0; k < 16; k++) {
    mb->prev_intra4x4_pred_mode_flag[k] = (ps[k] >> 3) & 1;
    mb->rem_intra4x4_pred_mode[k] = ps[k] & 7;
  }
  for (int k = 0; k < 4; k++) {
    mb->prev_intra8x8_pred_mode_flag[k] = (ps[k] >> 3) & 1;
    mb->rem_intra8x8_pred_mode[k] = ps[k] & 7;
  }
  const int16_t* cf = soa->coeff + idx * DRYV_COEFFS_PER_MB;
  if (mb->mode == MODE_I4x4) {
    for (int b = 0; b < 16; b++)
      for (int k = 0; k < 16; k++) mb->block_luma_4x4[b][k] = cf[b * 16 + k];
  } else if (mb->mode == MODE_I8x8) {
    for (int b = 0; b < 4; b++)
      for (int k = 0; k < 64; k++) mb->block_luma_8x8[b][k] = cf[b * 64 + k];
  } else {
    for (int b = 0; b < 16; b++) {
      mb->block_luma_dc[b] = cf[b * 16];
      for (int k = 0; k < 15; k++) mb->block_luma_ac[b][k] = cf[b * 16 + 1 + k];
    }
  }
  for (int pl = 0; pl < 2; pl++)
    for (int b = 0; b < 4; b++) {
      const int16_t* blk = cf + 256 + pl * 64 + b * 16;
      mb->block_chroma_dc[pl][b] = blk[0];
      for (int k = 0; k < 15; k++) mb->block_chroma_ac[pl][b][k] = blk[1 + k];
    }
  NbMb* me = &s->macroblocks[s->curr_mb_addr];
  memset(me, 0, sizeof *me);
  me->mode = mb->mode;
  return 0;
}

static int ctx_init(Ctx* s, const dryv_pic_params* pp, uint8_t* frame_out, NbMb* nb) {
  if (!pp || pp->pic_width_in_mbs == 0 || pp->pic_height_in_mbs == 0) return DRYV_ORACLE_ERR_ARG;
  memset(s, 0, sizeof *s);
  s->pic_width_in_mbs = pp->pic_width_in_mbs;
  s->pic_height_in_mbs = pp->pic_height_in_mbs;
  s->width_l = s->pic_width_in_mbs * 16;
  s->height_l = s->pic_height_in_mbs * 16;
  s->width_c = s->width_l / 2;
  s->height_c = s->height_l / 2;
  s->chroma_qp_index_offset = pp->chroma_qp_index_offset;
  s->second_chroma_qp_index_offset = pp->second_chroma_qp_index_offset;
  for (int k = 0; k < 16; k++) s->scaling_list4x4[k] = pp->scaling_list4x4[k];
  for (int k = 0; k < 64; k++) s->scaling_list8x8[k] = pp->scaling_list8x8[k];
  s->macroblocks = nb;
  s->luma = frame_out;
  s->cb = frame_out + (size_t)s->width_l * s->height_l;
  s->cr = s->cb + (size_t)s->width_c * s->height_c;
  return 0;
}

size_t dryv_oracle_frame_bytes(const dryv_pic_params* pp) {
  return (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs * 384;
}

/* One picture: Frame::new (zeroed planes, frame/mod.rs:29-46) + the MB loop of Slice::data
 * (slice/mod.rs:199-254) calling Frame::decode per MB. */
static int reconstruct_one(const dryv_pic_params* pp, const dryv_mb_soa* soa, size_t frame,
                           uint8_t* out) {
  size_t n_mb = (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  NbMb* nb = (NbMb*)calloc(n_mb, sizeof(NbMb));
  Ctx* s = (Ctx*)malloc(sizeof(Ctx));
  if (!nb || !s) { free(nb); free(s); return DRYV_ORACLE_ERR_ARG; }
  int rc = ctx_init(s, pp, out, nb);
  if (rc == 0) {
    memset(out, 0, n_mb * 384);
    for (size_t a = 0; a < n_mb && rc == 0; a++) {
      s->curr_mb_addr = (isz)a;
      rc = load_mb(s, soa, frame * n_mb + a);
      if (rc == 0) rc = frame_decode(s);
    }
  }
  free(nb);
  free(s);
  return rc;
}

int dryv_oracle_reconstruct(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                            uint8_t* out_yuv) {
  if (!pp || !soa || !out_yuv) return DRYV_ORACLE_ERR_ARG;
  size_t fb = dryv_oracle_frame_bytes(pp);
  for (uint32_t f = 0; f < n_frames; f++) {
    int rc = reconstruct_one(pp, soa, f, out_yuv + (size_t)f * fb);
    if (rc != 0) return rc;
  }
  return 0;
}

typedef struct {
  const dryv_pic_params* pp;
  const dryv_mb_soa* soa;
  uint8_t* out;
  uint32_t n_frames, stride, first;
  int rc;
} Job;
static void* job_main(void* arg) {
  Job* j = (Job*)arg;
  size_t fb = dryv_oracle_frame_bytes(j->pp);
  for (uint32_t f = j->first; f < j->n_frames; f += j->stride) {
    int rc = reconstruct_one(j->pp, j->soa, f, j->out + (size_t)f * fb);
    if (rc != 0) { j->rc = rc; break; }
  }
  return NULL;
}
/* One picture per thread at a time (pictures are independent); n_threads host threads. */
int dryv_oracle_reconstruct_mt(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                               uint8_t* out_yuv, uint32_t n_threads) {
  if (!pp || !soa || !out_yuv || n_threads == 0) return DRYV_ORACLE_ERR_ARG;
  if (n_threads > 256) n_threads = 256;
  if (n_threads > n_frames) n_threads = n_frames ? n_frames : 1;
  pthread_t th[256];
  Job jobs[256];
  for (uint32_t t = 0; t < n_threads; t++) {
    jobs[t] = (Job){pp, soa, out_yuv, n_frames, n_threads, t, 0};
    if (pthread_create(&th[t], NULL, job_main, &jobs[t]) != 0) return DRYV_ORACLE_ERR_ARG;
  }
  int rc = 0;
  for (uint32_t t = 0; t < n_threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc != 0) rc = jobs[t].rc;
  }
  return rc;
}

/* BASELINE config 2: residual path only. out = clip(pred + r) with `pred_yuv` a caller-supplied
 * prediction picture; same scaling/transform functions and picture construction as above. */
int dryv_oracle_residual_add(const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                             const uint8_t* pred_yuv, uint8_t* out_yuv) {
  if (!pp || !soa || !out_yuv || !pred_yuv) return DRYV_ORACLE_ERR_ARG;
  size_t n_mb = (size_t)pp->pic_width_in_mbs * pp->pic_height_in_mbs;
  size_t fb = dryv_oracle_frame_bytes(pp);
  NbMb* nb = (NbMb*)calloc(n_mb, sizeof(NbMb));
  Ctx* s = (Ctx*)malloc(sizeof(Ctx));
  if (!nb || !s) { free(nb); free(s); return DRYV_ORACLE_ERR_ARG; }
  int rc = 0;
  for (uint32_t f = 0; f < n_frames && rc == 0; f++) {
    rc = ctx_init(s, pp, out_yuv + f * fb, nb);
    const uint8_t* pl = pred_yuv + f * fb;
    const uint8_t* pcb = pl + (size_t)s->width_l * s->height_l;
    const uint8_t* pcr = pcb + (size_t)s->width_c * s->height_c;
    for (size_t a = 0; a < n_mb && rc == 0; a++) {
      s->curr_mb_addr = (isz)a;
      rc = load_mb(s, soa, f * n_mb + a);
      if (rc != 0) break;
      isz x_p = inverse_raster_scan((isz)a, 16, 16, s->width_l, 0);
      isz y_p = inverse_raster_scan((isz)a, 16, 16, s->width_l, 1);
      scaling(s);
      isz r_l[16][16]; /* [y][x] */
      if (s->mb.mode == MODE_I8x8) {
        for (int blk = 0; blk < 4; blk++) {
          isz c[8][8], r[8][8];
          inverse_scanner_8x8(s->mb.block_luma_8x8[blk], c);
          scaling_and_transform8x8(s, c, r);
          for (int i = 0; i < 8; i++)
            for (int j = 0; j < 8; j++) r_l[(blk / 2) * 8 + i][(blk % 2) * 8 + j] = r[i][j];
        }
      } else {
        isz dcs[16];
        if (s->mb.mode == MODE_I16x16) {
          isz c[4][4], dc_y[4][4];
          inverse_scanner4x4(s->mb.block_luma_dc, c);
          transform_intra16x16_dc(s, c, dc_y);
          const isz m[16] = {dc_y[0][0], dc_y[0][1], dc_y[1][0], dc_y[1][1], dc_y[0][2], dc_y[0][3],
                             dc_y[1][2], dc_y[1][3], dc_y[2][0], dc_y[2][1], dc_y[3][0], dc_y[3][1],
                             dc_y[2][2], dc_y[2][3], dc_y[3][2], dc_y[3][3]};
          memcpy(dcs, m, sizeof m);
        }
        for (int blk = 0; blk < 16; blk++) {
          isz list[16], c[4][4], r[4][4];
          if (s->mb.mode == MODE_I16x16) {
            list[0] = dcs[blk];
            for (int k = 0; k < 15; k++) list[1 + k] = s->mb.block_luma_ac[blk][k];
          } else {
            memcpy(list, s->mb.block_luma_4x4[blk], sizeof list);
          }
          inverse_scanner4x4(list, c);
          scaling_and_transform4x4(s, c, 1, 0, r);
          isz x_o = inverse_raster_scan(blk / 4, 8, 8, 16, 0) + inverse_raster_scan(blk % 4, 4, 4, 8, 0);
          isz y_o = inverse_raster_scan(blk / 4, 8, 8, 16, 1) + inverse_raster_scan(blk % 4, 4, 4, 8, 1);
          for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) r_l[y_o + i][x_o + j] = r[i][j];
        }
      }
      for (isz y = 0; y < 16; y++)
        for (isz x = 0; x < 16; x++) {
          size_t o = (size_t)(y_p + y) * s->width_l + (size_t)(x_p + x);
          s->luma[o] = (uint8_t)clampz((isz)pl[o] + r_l[y][x], 0, 255);
        }
      for (int cbf = 1; cbf >= 0; cbf--) {
        int pli = cbf ? 0 : 1;
        isz c[2][2], dc_c[2][2];
        c[0][0] = s->mb.block_chroma_dc[pli][0];
        c[0][1] = s->mb.block_chroma_dc[pli][1];
        c[1][0] = s->mb.block_chroma_dc[pli][2];
        c[1][1] = s->mb.block_chroma_dc[pli][3];
        transform_chroma_dc(s, c, cbf, dc_c);
        const isz dcc[4] = {dc_c[0][0], dc_c[0][1], dc_c[1][0], dc_c[1][1]};
        for (int blk = 0; blk < 4; blk++) {
          isz list[16], cc[4][4], r[4][4];
          list[0] = dcc[blk];
          for (int k = 0; k < 15; k++) list[1 + k] = s->mb.block_chroma_ac[pli][blk][k];
          inverse_scanner4x4(list, cc);
          scaling_and_transform4x4(s, cc, 0, cbf, r);
          for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
              size_t o = (size_t)(y_p / 2 + (blk / 2) * 4 + i) * s->width_c + (size_t)(x_p / 2 + (blk % 2) * 4 + j);
              const uint8_t* pp_ = cbf ? pcb : pcr;
              uint8_t* op = cbf ? s->cb : s->cr;
              op[o] = (uint8_t)clampz((isz)pp_[o] + r[i][j], 0, 255);
            }
        }
      }
    }
  }
  free(nb);
  free(s);
  return rc;
}

/* Stage-level entry points for unit tests (flat scaling or the lists in pp). `mode`: 0 luma of an
 * I4x4 MB, 1 luma of an I16x16 MB (DC passthrough), 2 chroma Cb, 3 chroma Cr. */
int dryv_oracle_block4x4(const dryv_pic_params* pp, int qp1y, int mode, const int16_t coeff_zz[16],
                         int32_t r_out[16]) {
  NbMb nb;
  Ctx* s = (Ctx*)malloc(sizeof(Ctx));
  uint8_t dummy[384];
  dryv_pic_params p1 = *pp;
  p1.pic_width_in_mbs = p1.pic_height_in_mbs = 1;
  if (!s || ctx_init(s, &p1, dummy, &nb) != 0) { free(s); return DRYV_ORACLE_ERR_ARG; }
  memset(&s->mb, 0, sizeof s->mb);
  s->mb.mode = mode == 1 ? MODE_I16x16 : MODE_I4x4;
  s->mb.qp1y = s->mb.qpy = qp1y;
  scaling(s);
  isz list[16], c[4][4], r[4][4];
  for (int k = 0; k < 16; k++) list[k] = coeff_zz[k];
  inverse_scanner4x4(list, c);
  scaling_and_transform4x4(s, c, mode < 2, mode == 2, r);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r_out[i * 4 + j] = (int32_t)r[i][j];
  free(s);
  return 0;
}

int dryv_oracle_block8x8(const dryv_pic_params* pp, int qp1y, const int16_t coeff_zz[64],
                         int32_t r_out[64]) {
  NbMb nb;
  Ctx* s = (Ctx*)malloc(sizeof(Ctx));
  uint8_t dummy[384];
  dryv_pic_params p1 = *pp;
  p1.pic_width_in_mbs = p1.pic_height_in_mbs = 1;
  if (!s || ctx_init(s, &p1, dummy, &nb) != 0) { free(s); return DRYV_ORACLE_ERR_ARG; }
  memset(&s->mb, 0, sizeof s->mb);
  s->mb.mode = MODE_I8x8;
  s->mb.qp1y = s->mb.qpy = qp1y;
  scaling(s);
  isz list[64], c[8][8], r[8][8];
  for (int k = 0; k < 64; k++) list[k] = coeff_zz[k];
  inverse_scanner_8x8(list, c);
  scaling_and_transform8x8(s, c, r);
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) r_out[i * 8 + j] = (int32_t)r[i][j];
  free(s);
  return 0;
}

/* Frame::write_to_yuv_file, frame/mod.rs:48-70: Y rows, Cb rows, Cr rows, no header. */
int dryv_oracle_write_yuv_file(const uint8_t* frame_yuv, size_t bytes, const char* path) {
  FILE* f = fopen(path, "wb");
  if (!f) return DRYV_ORACLE_ERR_ARG;
  size_t w = fwrite(frame_yuv, 1, bytes, f);
  fclose(f);
  return w == bytes ? 0 : DRYV_ORACLE_ERR_ARG;
}
