/*
 * synth.c — seeded generator of spec-legal macroblock syntax buffers (workload generator).
 *
 * There is no H.264 encoder, MP4 asset or Rust toolchain in the build image, so the buffers CABAC
 * would emit (include/dryv_recon.h: dryv_mb_soa) are synthesised: per macroblock a type, legal
 * prediction modes expressed as prev/rem syntax (the forward direction of pred4x4.rs:363-427 /
 * pred8x8.rs:698-764 of the reference), a QP, and coefficient levels obtained by forward-transforming
 * and quantising a bounded random spatial residual with the standard H.264 forward 4x4/8x8 integer
 * transforms (so every dequant/IDCT intermediate stays inside the range a conforming stream allows).
 *
 * PRNG: SplitMix64, one stream per picture, state = seed (documented so buffers can be regenerated).
 * This is neither the oracle nor the product path: bench.py, tests/ and smoke() use it for inputs.
 */
#include "synth.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } Rng;
static inline uint64_t rng_next(Rng* r) {
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline uint32_t rng_below(Rng* r, uint32_t n) { return (uint32_t)((rng_next(r) >> 32) * (uint64_t)n >> 32); }

/* 16-bit uniform -> rounded, clipped Laplace(0,b) sample, via inverse CDF tables built once. */
static int16_t g_lap_normal[65536];
static int16_t g_lap_stress[65536];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void build_table(int16_t* t, double b, int clip) {
  for (int i = 0; i < 65536; i++) {
    double u = (i + 0.5) / 65536.0 - 0.5;
    double a = fabs(u);
    double x = -b * log(1.0 - 2.0 * a);
    long v = lround(x);
    if (v > clip) v = clip;
    t[i] = (int16_t)(u < 0 ? -v : v);
  }
}
static void build_tables(void) {
  build_table(g_lap_normal, 6.0, 64);
  build_table(g_lap_stress, 40.0, 255);
}

static const uint8_t ZZ4[16][2] = {{0, 0}, {0, 1}, {1, 0}, {2, 0}, {1, 1}, {0, 2}, {0, 3}, {1, 2},
                                   {2, 1}, {3, 0}, {3, 1}, {2, 2}, {1, 3}, {2, 3}, {3, 2}, {3, 3}};
static const uint8_t ZZ8[64][2] = {
    {0, 0}, {0, 1}, {1, 0}, {2, 0}, {1, 1}, {0, 2}, {0, 3}, {1, 2}, {2, 1}, {3, 0}, {4, 0},
    {3, 1}, {2, 2}, {1, 3}, {0, 4}, {0, 5}, {1, 4}, {2, 3}, {3, 2}, {4, 1}, {5, 0}, {6, 0},
    {5, 1}, {4, 2}, {3, 3}, {2, 4}, {1, 5}, {0, 6}, {0, 7}, {1, 6}, {2, 5}, {3, 4}, {4, 3},
    {5, 2}, {6, 1}, {7, 0}, {7, 1}, {6, 2}, {5, 3}, {4, 4}, {3, 5}, {2, 6}, {1, 7}, {2, 7},
    {3, 6}, {4, 5}, {5, 4}, {6, 3}, {7, 2}, {7, 3}, {6, 4}, {5, 5}, {4, 6}, {3, 7}, {4, 7},
    {5, 6}, {6, 5}, {7, 4}, {7, 5}, {6, 6}, {5, 7}, {6, 7}, {7, 6}, {7, 7}};

/* Standard encoder-side multiplication factors (inverse of normAdjust, 2^15 / 2^16 fixed point). */
static const int MF4[6][3] = {{13107, 5243, 8066}, {11916, 4660, 7490}, {10082, 4194, 6554},
                              {9362, 3647, 5825},  {8192, 3355, 5243},  {7282, 2893, 4559}};
static const int MF8[6][6] = {{13107, 11428, 20972, 12222, 16777, 15481}, {11916, 10826, 19174, 11058, 14980, 14290},
                              {10082, 8943, 15978, 9675, 12710, 11985},   {9362, 8228, 14913, 8931, 11984, 11259},
                              {8192, 7346, 13159, 7740, 10486, 9777},     {7282, 6428, 11570, 6830, 9118, 8640}};
static int cls4(int i, int j) { return (i % 2 == 0 && j % 2 == 0) ? 0 : ((i % 2 == 1 && j % 2 == 1) ? 1 : 2); }
static int cls8(int i, int j) {
  if (i % 4 == 0 && j % 4 == 0) return 0;
  if (i % 2 == 1 && j % 2 == 1) return 1;
  if (i % 4 == 2 && j % 4 == 2) return 2;
  if ((i % 4 == 0 && j % 2 == 1) || (i % 2 == 1 && j % 4 == 0)) return 3;
  if ((i % 4 == 0 && j % 4 == 2) || (i % 4 == 2 && j % 4 == 0)) return 4;
  return 5;
}
static const int QPC_TAB[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
static int qpc_of(int qpy, int off) {
  int q = qpy + off;
  q = q < 0 ? 0 : (q > 51 ? 51 : q);
  return q < 30 ? q : QPC_TAB[q - 30];
}

static inline int clip16(long v) { return (int)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v)); }
static inline int quant(long w, long mf, int qbits, long f) {
  long a = w < 0 ? -w : w;
  long l = (a * mf + f) >> qbits;
  return clip16(w < 0 ? -l : l);
}

static void fwd4x4(const int x[4][4], long y[4][4]) {
  long t[4][4];
  for (int i = 0; i < 4; i++) {
    long a0 = x[i][0] + x[i][3], a1 = x[i][1] + x[i][2], a2 = x[i][1] - x[i][2], a3 = x[i][0] - x[i][3];
    t[i][0] = a0 + a1; t[i][1] = 2 * a3 + a2; t[i][2] = a0 - a1; t[i][3] = a3 - 2 * a2;
  }
  for (int j = 0; j < 4; j++) {
    long a0 = t[0][j] + t[3][j], a1 = t[1][j] + t[2][j], a2 = t[1][j] - t[2][j], a3 = t[0][j] - t[3][j];
    y[0][j] = a0 + a1; y[1][j] = 2 * a3 + a2; y[2][j] = a0 - a1; y[3][j] = a3 - 2 * a2;
  }
}
static void fwd8_1d(const long p[8], long o[8]) {
  long a0 = p[0] + p[7], a1 = p[1] + p[6], a2 = p[2] + p[5], a3 = p[3] + p[4];
  long b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
  a0 = p[0] - p[7]; a1 = p[1] - p[6]; a2 = p[2] - p[5]; a3 = p[3] - p[4];
  long b4 = a1 + a2 + ((a0 >> 1) + a0);
  long b5 = a0 - a3 - ((a2 >> 1) + a2);
  long b6 = a0 + a3 - ((a1 >> 1) + a1);
  long b7 = a1 - a2 + ((a3 >> 1) + a3);
  o[0] = b0 + b1; o[2] = b2 + (b3 >> 1); o[4] = b0 - b1; o[6] = (b2 >> 1) - b3;
  o[1] = b4 + (b7 >> 2); o[3] = b5 + (b6 >> 2); o[5] = b6 - (b5 >> 2); o[7] = -b7 + (b4 >> 2);
}
static void fwd8x8(const int x[8][8], long y[8][8]) {
  long t[8][8];
  for (int i = 0; i < 8; i++) {
    long p[8], o[8];
    for (int j = 0; j < 8; j++) p[j] = x[i][j];
    fwd8_1d(p, o);
    for (int j = 0; j < 8; j++) t[i][j] = o[j];
  }
  for (int j = 0; j < 8; j++) {
    long p[8], o[8];
    for (int i = 0; i < 8; i++) p[i] = t[i][j];
    fwd8_1d(p, o);
    for (int i = 0; i < 8; i++) y[i][j] = o[i];
  }
}

typedef struct {
  uint8_t cls;    /* 0 I4x4, 1 I8x8, 2 I16x16 */
  uint8_t m4[16]; /* resolved Intra4x4PredMode per block */
  uint8_t m8[4];  /* resolved Intra8x8PredMode per block */
} MbModes;

static const uint8_t BLK4_X[16] = {0, 4, 0, 4, 8, 12, 8, 12, 0, 4, 0, 4, 8, 12, 8, 12};
static const uint8_t BLK4_Y[16] = {0, 0, 4, 4, 0, 0, 4, 4, 8, 8, 12, 12, 8, 8, 12, 12};
static int blk4_of(int x, int y) { return 8 * (y / 8) + 4 * (x / 8) + 2 * ((y % 8) / 4) + ((x % 8) / 4); }

/* neighbour mode as an Intra4x4 block sees it (pred4x4.rs:394-412); nb NULL = unavailable */
static int nb_mode_for4(const MbModes* nb, int blk4) {
  if (nb->cls == 0) return nb->m4[blk4];
  if (nb->cls == 1) return nb->m8[blk4 >> 2];
  return 2;
}
/* neighbour mode as an Intra8x8 block sees it (pred8x8.rs:731-751); n = 1 for A, 2 for B */
static int nb_mode_for8(const MbModes* nb, int blk8, int n) {
  if (nb->cls == 1) return nb->m8[blk8];
  if (nb->cls == 0) return nb->m4[blk8 * 4 + n];
  return 2;
}

static int pick_legal(Rng* r, int top, int left, int corner) {
  int legal[9], n = 0;
  legal[n++] = 2;
  if (top) { legal[n++] = 0; legal[n++] = 3; legal[n++] = 7; }
  if (left) { legal[n++] = 1; legal[n++] = 8; }
  if (top && left && corner) { legal[n++] = 4; legal[n++] = 5; legal[n++] = 6; }
  return legal[rng_below(r, (uint32_t)n)];
}
static int mode_is_legal(int m, int top, int left, int corner) {
  switch (m) {
    case 2: return 1;
    case 0: case 3: case 7: return top;
    case 1: case 8: return left;
    default: return top && left && corner;
  }
}
static uint8_t encode_mode(Rng* r, int pred, int top, int left, int corner, int* chosen) {
  int m;
  if (rng_below(r, 100) < 30 && mode_is_legal(pred, top, left, corner)) m = pred;
  else m = pick_legal(r, top, left, corner);
  *chosen = m;
  if (m == pred) return (uint8_t)(8 | rng_below(r, 8)); /* prev flag set; rem bits are don't-care */
  return (uint8_t)(m < pred ? m : m - 1);
}

int dryv_synth_frame(const dryv_pic_params* pp, const dryv_synth_cfg* cfg, uint64_t seed, uint8_t* mb_type,
                     uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp_out, uint8_t* pred_syntax, int16_t* coeff) {
  if (!pp || !cfg || !mb_type || !t8x8 || !chroma_mode || !qp_out || !pred_syntax || !coeff) return -1;
  pthread_once(&g_once, build_tables);
  const int W = pp->pic_width_in_mbs, H = pp->pic_height_in_mbs;
  if (W <= 0 || H <= 0) return -1;
  MbModes* modes = (MbModes*)calloc((size_t)W * H, sizeof(MbModes));
  if (!modes) return -1;
  Rng rng = {seed * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL};
  int w4[4][4], w8[8][8];
  for (int k = 0; k < 16; k++) w4[ZZ4[k][0]][ZZ4[k][1]] = pp->scaling_list4x4[k] ? pp->scaling_list4x4[k] : 16;
  for (int k = 0; k < 64; k++) w8[ZZ8[k][0]][ZZ8[k][1]] = pp->scaling_list8x8[k] ? pp->scaling_list8x8[k] : 16;

  for (int my = 0; my < H; my++)
    for (int mx = 0; mx < W; mx++) {
      const size_t a = (size_t)my * W + mx;
      MbModes* me = &modes[a];
      const int availA = mx > 0, availB = my > 0, availD = mx > 0 && my > 0;
      const MbModes* nbA = availA ? &modes[a - 1] : NULL;
      const MbModes* nbB = availB ? &modes[a - W] : NULL;
      uint8_t* ps = pred_syntax + a * 16;
      int16_t* cf = coeff + a * DRYV_COEFFS_PER_MB;
      memset(ps, 0, 16);
      memset(cf, 0, DRYV_COEFFS_PER_MB * sizeof(int16_t));

      uint32_t tsel = rng_below(&rng, 100);
      int cls = tsel < (uint32_t)cfg->pct_i4x4 ? 0 : (tsel < (uint32_t)(cfg->pct_i4x4 + cfg->pct_i8x8) ? 1 : 2);
      if (cfg->standard_only && cls == 1 && mx == 0) cls = 0;  /* see dryv_synth_cfg::standard_only */
      me->cls = (uint8_t)cls;
      int qp = cfg->qp_base;
      if (cfg->qp_jitter > 0) qp += (int)rng_below(&rng, 2 * (uint32_t)cfg->qp_jitter + 1) - cfg->qp_jitter;
      qp = qp < 0 ? 0 : (qp > 51 ? 51 : qp);
      qp_out[a] = (uint8_t)qp;

      /* chroma mode: 0 DC, 1 H (needs A), 2 V (needs B), 3 Plane (A and B) */
      {
        int legal[4], n = 0;
        legal[n++] = 0;
        if (availA) legal[n++] = 1;
        if (availB) legal[n++] = 2;
        if (availA && availB) legal[n++] = 3;
        chroma_mode[a] = (uint8_t)legal[rng_below(&rng, (uint32_t)n)];
      }
      int pred16 = 0;
      if (cls == 0) {
        t8x8[a] = 0;
        for (int b = 0; b < 16; b++) {
          int x = BLK4_X[b], y = BLK4_Y[b];
          int top = y > 0 || availB, left = x > 0 || availA;
          int corner = (x > 0 && y > 0) ? 1 : (x > 0 ? availB : (y > 0 ? availA : availD));
          int pred = 2;
          int haveA = x > 0 || availA, haveB = y > 0 || availB;
          if (haveA && haveB) {
            int ma = x > 0 ? me->m4[blk4_of(x - 1, y)] : nb_mode_for4(nbA, blk4_of(15, y));
            int mb = y > 0 ? me->m4[blk4_of(x, y - 1)] : nb_mode_for4(nbB, blk4_of(x, 15));
            pred = ma < mb ? ma : mb;
          }
          int m;
          ps[b] = encode_mode(&rng, pred, top, left, corner, &m);
          me->m4[b] = (uint8_t)m;
        }
      } else if (cls == 1) {
        t8x8[a] = 1;
        for (int b = 0; b < 4; b++) {
          int x = (b % 2) * 8, y = (b / 2) * 8;
          int top = y > 0 || availB, left = x > 0 || availA;
          int corner = (x > 0 && y > 0) ? 1 : (x > 0 ? availB : (y > 0 ? availA : availD));
          int pred = 2;
          if (left && top) {
            int ma = x > 0 ? me->m8[b - 1] : nb_mode_for8(nbA, 2 * (y / 8) + 1, 1);
            int mb = y > 0 ? me->m8[b - 2] : nb_mode_for8(nbB, 2 + (x / 8), 2);
            pred = ma < mb ? ma : mb;
          }
          int m;
          ps[b] = encode_mode(&rng, pred, top, left, corner, &m);
          me->m8[b] = (uint8_t)m;
        }
      } else {
        t8x8[a] = 0;
        int legal[4], n = 0;
        legal[n++] = 2;
        if (availB) legal[n++] = 0;
        if (availA) legal[n++] = 1;
        if (availA && availB) legal[n++] = 3;
        pred16 = legal[rng_below(&rng, (uint32_t)n)];
      }

      /* spatial residual: 256 luma + 64 Cb + 64 Cr */
      int res_l[16][16], res_c[2][8][8];
      {
        int stress = rng_below(&rng, 100) < (uint32_t)cfg->stress_pct;
        int flat = 0; /* 1: all -255 (forces pixel 0), 2: all +255 */
        if (stress) {
          uint32_t k = rng_below(&rng, 4);
          flat = k == 0 ? 1 : (k == 1 ? 2 : 0);
        }
        const int16_t* tab = stress ? g_lap_stress : g_lap_normal;
        if (cfg->zero_residual) {
          memset(res_l, 0, sizeof res_l);
          memset(res_c, 0, sizeof res_c);
        } else if (flat) {
          int v = flat == 1 ? -255 : 255;
          for (int i = 0; i < 16; i++) for (int j = 0; j < 16; j++) res_l[i][j] = v;
          for (int p = 0; p < 2; p++) for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) res_c[p][i][j] = v;
        } else {
          for (int i = 0; i < 16; i++)
            for (int j = 0; j < 16; j += 4) {
              uint64_t u = rng_next(&rng);
              res_l[i][j] = tab[u & 0xffff]; res_l[i][j + 1] = tab[(u >> 16) & 0xffff];
              res_l[i][j + 2] = tab[(u >> 32) & 0xffff]; res_l[i][j + 3] = tab[(u >> 48) & 0xffff];
            }
          for (int p = 0; p < 2; p++)
            for (int i = 0; i < 8; i++)
              for (int j = 0; j < 8; j += 4) {
                uint64_t u = rng_next(&rng);
                res_c[p][i][j] = tab[u & 0xffff]; res_c[p][i][j + 1] = tab[(u >> 16) & 0xffff];
                res_c[p][i][j + 2] = tab[(u >> 32) & 0xffff]; res_c[p][i][j + 3] = tab[(u >> 48) & 0xffff];
              }
        }
      }

      int any_luma_ac = 0;
      if (cls == 1) {
        int qbits = 16 + qp / 6;
        long f = (1L << qbits) / 3;
        for (int b = 0; b < 4; b++) {
          int x[8][8];
          long y[8][8];
          for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) x[i][j] = res_l[(b / 2) * 8 + i][(b % 2) * 8 + j];
          fwd8x8(x, y);
          for (int k = 0; k < 64; k++) {
            int i = ZZ8[k][0], j = ZZ8[k][1];
            long mf = (long)MF8[qp % 6][cls8(i, j)] * 16 / w8[i][j];
            cf[b * 64 + k] = (int16_t)quant(y[i][j], mf, qbits, f);
          }
        }
      } else {
        int qbits = 15 + qp / 6;
        long f = (1L << qbits) / 3;
        long dc[4][4];
        for (int b = 0; b < 16; b++) {
          int x[4][4];
          long y[4][4];
          for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) x[i][j] = res_l[BLK4_Y[b] + i][BLK4_X[b] + j];
          fwd4x4(x, y);
          for (int k = (cls == 2 ? 1 : 0); k < 16; k++) {
            int i = ZZ4[k][0], j = ZZ4[k][1];
            long mf = (long)MF4[qp % 6][cls4(i, j)] * 16 / w4[i][j];
            int l = quant(y[i][j], mf, qbits, f);
            cf[b * 16 + k] = (int16_t)l;
            if (k > 0 && l != 0) any_luma_ac = 1;
          }
          dc[BLK4_Y[b] / 4][BLK4_X[b] / 4] = y[0][0];
        }
        if (cls == 2) {
          /* forward 4x4 Hadamard of the 16 DCs, /2, DC quantiser (qbits + 1, 2f) */
          static const int A[4][4] = {{1, 1, 1, 1}, {1, 1, -1, -1}, {1, -1, -1, 1}, {1, -1, 1, -1}};
          long g[4][4] = {{0}}, h[4][4] = {{0}};
          for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) g[i][j] += A[i][k] * dc[k][j];
          for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) h[i][j] += g[i][k] * A[k][j];
          long mf = (long)MF4[qp % 6][0] * 16 / w4[0][0];
          for (int k = 0; k < 16; k++) {
            int i = ZZ4[k][0], j = ZZ4[k][1];
            cf[k * 16] = (int16_t)quant(h[i][j] / 2, mf, qbits + 1, 2 * f);
          }
        }
      }
      int any_c_dc = 0, any_c_ac = 0;
      for (int p = 0; p < 2; p++) {
        int qpc = qpc_of(qp, p == 0 ? pp->chroma_qp_index_offset : pp->second_chroma_qp_index_offset);
        int qbits = 15 + qpc / 6;
        long f = (1L << qbits) / 3;
        long dc[2][2];
        int16_t* cc = cf + 256 + p * 64;
        for (int b = 0; b < 4; b++) {
          int x[4][4];
          long y[4][4];
          for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) x[i][j] = res_c[p][(b / 2) * 4 + i][(b % 2) * 4 + j];
          fwd4x4(x, y);
          for (int k = 1; k < 16; k++) {
            int i = ZZ4[k][0], j = ZZ4[k][1];
            /* Q1 of the reference: chroma is dequantised with the luma list, so quantise against it */
            long mf = (long)MF4[qpc % 6][cls4(i, j)] * 16 / w4[i][j];
            int l = quant(y[i][j], mf, qbits, f);
            cc[b * 16 + k] = (int16_t)l;
            if (l != 0) any_c_ac = 1;
          }
          dc[b / 2][b % 2] = y[0][0];
        }
        long hd[2][2] = {{dc[0][0] + dc[0][1] + dc[1][0] + dc[1][1], dc[0][0] - dc[0][1] + dc[1][0] - dc[1][1]},
                         {dc[0][0] + dc[0][1] - dc[1][0] - dc[1][1], dc[0][0] - dc[0][1] - dc[1][0] + dc[1][1]}};
        long mf = (long)MF4[qpc % 6][0] * 16 / w4[0][0];
        for (int b = 0; b < 4; b++) {
          int l = quant(hd[b / 2][b % 2], mf, qbits + 1, 2 * f);
          cc[b * 16] = (int16_t)l;
          if (l != 0) any_c_dc = 1;
        }
      }
      if (cls == 2) {
        int cbp_c = any_c_ac ? 2 : (any_c_dc ? 1 : 0);
        mb_type[a] = (uint8_t)(1 + pred16 + 4 * cbp_c + (any_luma_ac ? 12 : 0));
      } else {
        mb_type[a] = 0;
      }
    }
  free(modes);
  return 0;
}

typedef struct {
  const dryv_pic_params* pp;
  const dryv_synth_cfg* cfg;
  uint64_t seed0;
  uint32_t n_frames, first, stride;
  uint8_t *mb_type, *t8x8, *chroma_mode, *qp, *pred_syntax;
  int16_t* coeff;
  int rc;
} Job;
static void* job_main(void* arg) {
  Job* j = (Job*)arg;
  size_t n_mb = (size_t)j->pp->pic_width_in_mbs * j->pp->pic_height_in_mbs;
  for (uint32_t f = j->first; f < j->n_frames; f += j->stride) {
    size_t o = (size_t)f * n_mb;
    dryv_synth_cfg c = *j->cfg;
    if (c.qp_step_per_frame) {
      c.qp_base += (int)f * c.qp_step_per_frame;
    }
    int rc = dryv_synth_frame(j->pp, &c, j->seed0 + f, j->mb_type + o, j->t8x8 + o, j->chroma_mode + o, j->qp + o,
                              j->pred_syntax + o * 16, j->coeff + o * DRYV_COEFFS_PER_MB);
    if (rc != 0) { j->rc = rc; break; }
  }
  return NULL;
}

int dryv_synth_batch(const dryv_pic_params* pp, const dryv_synth_cfg* cfg, uint64_t seed0, uint32_t n_frames,
                     uint8_t* mb_type, uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp, uint8_t* pred_syntax,
                     int16_t* coeff, uint32_t n_threads) {
  if (n_threads == 0) n_threads = 1;
  if (n_threads > 128) n_threads = 128;
  if (n_threads > n_frames) n_threads = n_frames ? n_frames : 1;
  pthread_t th[128];
  Job jobs[128];
  for (uint32_t t = 0; t < n_threads; t++) {
    jobs[t] = (Job){pp, cfg, seed0, n_frames, t, n_threads, mb_type, t8x8, chroma_mode, qp, pred_syntax, coeff, 0};
    if (pthread_create(&th[t], NULL, job_main, &jobs[t]) != 0) return -1;
  }
  int rc = 0;
  for (uint32_t t = 0; t < n_threads; t++) {
    pthread_join(th[t], NULL);
    if (jobs[t].rc != 0) rc = jobs[t].rc;
  }
  return rc;
}
