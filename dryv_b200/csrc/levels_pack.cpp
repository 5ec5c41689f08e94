// levels_pack.cpp — host side of the compact level stream (include/dryv_recon.h, dryv_mb_levels_compact).
//
// dryv_recon_pack_levels converts the dense per-macroblock level arrays (what the reference's
// residual_cabac leaves in Macroblock::block_*, cabac/mod.rs:563-675) into the significance-map + level
// records a CABAC host would append directly; dryv_recon_unpack_levels is the inverse, kept for the CPU
// tests of the format. Neither runs on the reconstruction path: the GPU expands the stream itself
// (expand_levels_kernel, recon.cu).
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/dryv_recon.h"
#include "levels_record.h"

namespace {

using namespace dryv_levels;

template <class F>
void parallel_for(size_t n, int threads, F f) {
  if (threads <= 1 || n < 4096) {
    f(0, n);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (n + (size_t)threads - 1) / (size_t)threads;
  for (int t = 0; t < threads; t++) {
    const size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
    if (lo >= hi) break;
    pool.emplace_back([=] { f(lo, hi); });
  }
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int dryv_recon_pack_levels(const int16_t* coeff, size_t n_mbs, uint32_t* offset, uint8_t* stream, size_t stream_cap,
                           int threads) {
  if (!coeff || !offset || !stream || n_mbs == 0) return DRYV_ERR_ARG;
  // pass 1: record sizes (offset[i + 1] holds the size of record i), then an exclusive prefix sum
  parallel_for(n_mbs, threads, [=](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) offset[i + 1] = record_size(coeff + i * DRYV_COEFFS_PER_MB);
  });
  uint64_t acc = 0;
  offset[0] = 0;
  for (size_t i = 0; i < n_mbs; i++) {
    acc += offset[i + 1];
    if (acc > 0xffffffffull || acc > stream_cap) return DRYV_ERR_ARG;
    offset[i + 1] = (uint32_t)acc;
  }
  // pass 2: the records
  parallel_for(n_mbs, threads, [=](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++)
      write_record(coeff + i * DRYV_COEFFS_PER_MB, stream + offset[i], offset[i + 1] - offset[i]);
  });
  return DRYV_OK;
}

int dryv_recon_unpack_levels(const dryv_mb_levels_compact* lv, size_t n_mbs, int16_t* coeff) {
  if (!lv || !lv->offset || !lv->stream || !coeff || n_mbs == 0) return DRYV_ERR_ARG;
  for (size_t i = 0; i < n_mbs; i++) {
    const uint32_t o = lv->offset[i], e = lv->offset[i + 1];
    if (e < o || (o & 3u) || e - o < 4 || e - o > DRYV_COMPACT_MAX_RECORD) return DRYV_ERR_ARG;
    const uint8_t* rec = lv->stream + o;
    const uint8_t* end = lv->stream + e;
    uint32_t hdr;
    memcpy(&hdr, rec, 4);
    if (hdr & 0x3f000000u) return DRYV_ERR_ARG;
    const int mode = (int)(hdr >> 30);
    if (mode > kModeInt16) return DRYV_ERR_ARG;
    int ncoded = 0;
    for (int b = 0; b < kSlots; b++) ncoded += (hdr >> b) & 1u;
    const uint8_t* mp = rec + 4;
    const uint8_t* p = mp + 2 * ncoded;
    if (p > end) return DRYV_ERR_ARG;
    int16_t* c = coeff + i * DRYV_COEFFS_PER_MB;
    memset(c, 0, DRYV_COEFFS_PER_MB * sizeof(int16_t));
    uint32_t nnz = 0;
    for (int b = 0; b < ncoded; b++) {
      uint16_t m;
      memcpy(&m, mp + 2 * b, 2);
      if (!m) return DRYV_ERR_ARG;  // a coded slot holds at least one level
      nnz += (uint32_t)__builtin_popcount(m);
    }
    const uint8_t* esc = p + (((nnz + 1) / 2 + 1) & ~1u);  // nibble coding only
    if (mode == kModeNibble && esc > end) return DRYV_ERR_ARG;
    uint32_t j = 0;
    for (int b = 0; b < kSlots; b++) {
      if (!((hdr >> b) & 1u)) continue;
      uint16_t m;
      memcpy(&m, mp, 2);
      mp += 2;
      for (int k = 0; k < 16; k++) {
        if (!((m >> k) & 1u)) continue;
        int16_t v;
        if (mode == kModeNibble) {
          const uint32_t code = (p[j >> 1] >> (4 * (j & 1))) & 15u;
          j++;
          if (code & 7u) {
            v = (int16_t)((code & 8u) ? -(int)(code & 7u) : (int)(code & 7u));
          } else {
            if (esc + 2 > end) return DRYV_ERR_ARG;
            memcpy(&v, esc, 2);
            esc += 2;
          }
        } else if (mode == kModeInt16) {
          if (p + 2 > end) return DRYV_ERR_ARG;
          memcpy(&v, p, 2);
          p += 2;
        } else {
          if (p + 1 > end) return DRYV_ERR_ARG;
          v = (int16_t)(int8_t)*p++;
        }
        c[b * 16 + k] = v;
      }
    }
  }
  return DRYV_OK;
}

}  // extern "C"
