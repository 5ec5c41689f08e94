// levels_record.h — one macroblock's record of the compact level stream (include/dryv_recon.h, dryv_mb_levels_compact):
// size and writer, shared by the dense -> compact converter (levels_pack.cpp) and the CABAC host, which emits records
// directly (cabac_host.cpp). Host code only.
#pragma once
#include <stdint.h>
#include <string.h>

#include "../../include/dryv_recon.h"

namespace dryv_levels {

constexpr int kSlots = DRYV_COEFFS_PER_MB / 16;

// Level codings (header bits 30..31): 0 = int8 per level, 1 = 4-bit code per level + int16 escapes, 2 = int16 per level.
enum { kModeInt8 = 0, kModeNibble = 1, kModeInt16 = 2 };

struct MbStats {
  uint32_t ncoded = 0, nnz = 0, nesc = 0;  // coded slots, non-zero levels, levels outside -7..7
  bool wide = false;                       // a level outside int8
  uint16_t mask[kSlots];
};

inline MbStats scan(const int16_t* c) {
  MbStats st;
  uint32_t wide = 0;
  for (int b = 0; b < kSlots; b++) {
    const int16_t* s = c + b * 16;
    uint64_t w[4];
    memcpy(w, s, 32);
    if (!(w[0] | w[1] | w[2] | w[3])) {  // most slots of a quantised picture hold nothing
      st.mask[b] = 0;
      continue;
    }
    uint32_t m = 0, esc = 0;
    for (int k = 0; k < 16; k++) {  // no data-dependent branch: a zero level counts for nothing below
      const int v = s[k];
      m |= (uint32_t)(v != 0) << k;
      esc += (uint32_t)(v + 7) > 14u;
      wide |= (uint32_t)(v + 128) > 255u;
    }
    st.mask[b] = (uint16_t)m;
    st.nnz += (uint32_t)__builtin_popcount(m);
    st.nesc += esc;
    st.ncoded++;
  }
  st.wide = wide != 0;
  return st;
}

// bytes of the level part in each coding, and the smallest legal one
inline uint32_t level_bytes(const MbStats& st, int mode) {
  if (mode == kModeInt8) return st.nnz;
  if (mode == kModeInt16) return 2 * st.nnz;
  return (((st.nnz + 1) / 2 + 1) & ~1u) + 2 * st.nesc;  // nibbles padded to 2 bytes, then the int16 escapes
}
inline int pick_mode(const MbStats& st) {
  int best = st.wide ? kModeInt16 : kModeInt8;
  if (level_bytes(st, kModeNibble) < level_bytes(st, best)) best = kModeNibble;
  return best;
}

inline uint32_t record_size(const MbStats& st, int mode) {
  const uint32_t bytes = 4 + 2 * st.ncoded + level_bytes(st, mode);
  return (bytes + 3u) & ~3u;
}
// size in bytes of macroblock `c`'s record
inline uint32_t record_size(const int16_t* c) {
  const MbStats st = scan(c);
  return record_size(st, pick_mode(st));
}

// writes the record of macroblock `c` (whose scan() is `st`) in coding `mode`, `size` = record_size(st, mode) bytes
inline void write_record(const int16_t* c, const MbStats& st, int mode, uint8_t* rec, uint32_t size) {
  uint32_t hdr = (uint32_t)mode << 30;
  for (int b = 0; b < kSlots; b++)
    if (st.mask[b]) hdr |= 1u << b;
  memcpy(rec, &hdr, 4);
  uint8_t* p = rec + 4;
  for (int b = 0; b < kSlots; b++)
    if (st.mask[b]) {
      memcpy(p, &st.mask[b], 2);
      p += 2;
    }
  if (mode == kModeNibble) {
    const uint32_t nib_bytes = ((st.nnz + 1) / 2 + 1) & ~1u;
    memset(p, 0, nib_bytes);
    uint8_t* esc = p + nib_bytes;
    uint32_t j = 0;
    for (int b = 0; b < kSlots; b++) {
      for (uint32_t m = st.mask[b]; m; m &= m - 1) {  // set bits in ascending order
        const int16_t v = c[b * 16 + __builtin_ctz(m)];
        uint32_t code = 0;  // 0 = escape
        if (v >= -7 && v <= 7) code = v < 0 ? (8u | (uint32_t)-v) : (uint32_t)v;
        else {
          memcpy(esc, &v, 2);
          esc += 2;
        }
        p[j >> 1] |= (uint8_t)(code << (4 * (j & 1)));
        j++;
      }
    }
    p = esc;
  } else {
    for (int b = 0; b < kSlots; b++) {
      for (uint32_t m = st.mask[b]; m; m &= m - 1) {
        const int16_t v = c[b * 16 + __builtin_ctz(m)];
        if (mode == kModeInt16) {
          memcpy(p, &v, 2);
          p += 2;
        } else {
          *p++ = (uint8_t)(int8_t)v;
        }
      }
    }
  }
  while (p < rec + size) *p++ = 0;
}
inline void write_record(const int16_t* c, uint8_t* rec, uint32_t size) {
  const MbStats st = scan(c);
  write_record(c, st, pick_mode(st), rec, size);
}

}  // namespace dryv_levels
