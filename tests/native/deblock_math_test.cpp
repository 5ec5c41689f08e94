// CPU unit test of the packed two-lines-per-register edge filters (dryv_b200/csrc/deblock_packed.cuh compiled for the host)
// against a scalar statement of H.264 8.7.2.3 (bS < 4) and 8.7.2.4 (bS 4), luma and chroma, every table index.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../dryv_b200/csrc/deblock_packed.cuh"

static const int kAlpha[52] = DRYV_DB_ALPHA;
static const int kBeta[52] = DRYV_DB_BETA;
static const int kTc0[52] = DRYV_DB_TC0_BS3;

static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// s[0..3] = p3..p0, s[4..7] = q0..q3, in place
static void scalar_edge(int* s, bool strong, bool chroma, int alpha, int beta, int tc0) {
  const int p3 = s[0], p2 = s[1], p1 = s[2], p0 = s[3], q0 = s[4], q1 = s[5], q2 = s[6], q3 = s[7];
  if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
  const bool ap = !chroma && abs(p2 - p0) < beta, aq = !chroma && abs(q2 - q0) < beta;
  if (strong) {
    const bool small = abs(p0 - q0) < ((alpha >> 2) + 2);
    if (ap && small) {
      s[3] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3;
      s[2] = (p2 + p1 + p0 + q0 + 2) >> 2;
      s[1] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
    } else {
      s[3] = (2 * p1 + p0 + q1 + 2) >> 2;
    }
    if (aq && small) {
      s[4] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3;
      s[5] = (p0 + q0 + q1 + q2 + 2) >> 2;
      s[6] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
    } else {
      s[4] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
    return;
  }
  const int tc = chroma ? tc0 + 1 : tc0 + (ap ? 1 : 0) + (aq ? 1 : 0);
  const int d = clip3(-tc, tc, (((q0 - p0) * 4) + (p1 - q1) + 4) >> 3);
  s[3] = clip3(0, 255, p0 + d);
  s[4] = clip3(0, 255, q0 - d);
  if (ap) s[2] = p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1);
  if (aq) s[5] = q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1);
}

static unsigned rng_state = 12345u;
static unsigned rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

int main() {
  long cases = 0, changed = 0;
  for (int ia = 0; ia < 52; ia++) {
    for (int ib = 0; ib < 52; ib += (ia % 3 == 0 ? 1 : 3)) {
      for (int rep = 0; rep < 400; rep++) {
        int line[2][8];
        for (int h = 0; h < 2; h++) {
          const int kind = rnd() % 4;   // flat-ish, step across the edge, noisy, extremes
          const int base = rnd() % 256, step = (int)(rnd() % 41) - 20, noise = kind == 2 ? 64 : (kind == 0 ? 3 : 8);
          for (int i = 0; i < 8; i++) {
            int v = base + (i >= 4 ? step : 0) + (int)(rnd() % (2 * noise + 1)) - noise;
            if (kind == 3) v = (rnd() & 1) ? 255 - (int)(rnd() % 4) : (int)(rnd() % 4);
            line[h][i] = clip3(0, 255, v);
          }
        }
        for (int mode = 0; mode < 8; mode++) {
          const bool strong = mode & 1, chroma = mode & 2;
          const unsigned off = (mode & 4) ? ((rnd() & 1) ? 0xffffu : 0xffff0000u) : 0u;
          int want[2][8];
          memcpy(want, line, sizeof want);
          for (int h = 0; h < 2; h++)
            if (!((off >> (16 * h)) & 1u)) scalar_edge(want[h], strong, chroma, kAlpha[ia], kBeta[ib], kTc0[ia]);
          uint32_t r[8];
          for (int i = 0; i < 8; i++) r[i] = (uint32_t)line[0][i] | ((uint32_t)line[1][i] << 16);
          const dryv::EdgeConst k = dryv::make_edge_const(kAlpha[ia], kBeta[ib], kTc0[ia], chroma);
          const uint32_t cm = chroma ? 0xffffffffu : 0u;
          if (strong) dryv::filter_edge_strong(r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], k, off, cm);
          else dryv::filter_edge_normal(r[1], r[2], r[3], r[4], r[5], r[6], k, off, cm);
          for (int h = 0; h < 2; h++)
            for (int i = 0; i < 8; i++) {
              const int got = (int)((r[i] >> (16 * h)) & 0xffffu);
              if (got != want[h][i]) {
                printf("MISMATCH ia %d ib %d mode %d half %d sample %d: got %d want %d (in:", ia, ib, mode, h, i, got, want[h][i]);
                for (int j = 0; j < 8; j++) printf(" %d", line[h][j]);
                printf(")\n");
                return 1;
              }
              if (want[h][i] != line[h][i]) changed++;
            }
          cases++;
        }
      }
    }
  }
  printf("ok: %ld packed edge filters equal the scalar ones (%ld samples changed)\n", cases, changed);
  return 0;
}
