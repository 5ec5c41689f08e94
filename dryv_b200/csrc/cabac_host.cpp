// cabac_host.cpp — CPU host side of the reconstruction path (include/dryv_cabac_host.h): Annex-B bytes -> SoA syntax.
//
// Restates, for IDR I-slice pictures, what the reference does before it reconstructs a macroblock: NAL split and
// emulation prevention (src/video/sample/nal.rs:230-253, src/byte/bit.rs:144-149), SPS/PPS
// (src/video/atom/avcc/sps.rs:42-121, pps.rs:30-58), slice header (src/video/slice/header.rs:145-315), the macroblock
// loop (src/video/slice/mod.rs:184-317), macroblock_layer / mb_pred / residual (src/video/cabac/mod.rs:89-210, :212-343,
// :433-675), the syntax-element context selection (:677-1111) and the arithmetic decoding engine (:1207-1308).
// Written from the text of ITU-T H.264 (7.3, 9.3) as the inverse of the stream writer in tests/avc/stream.py, which
// libavcodec accepts; organised around a per-macroblock neighbour record instead of the reference's Macroblock structs.
// Runs on the CPU by design (CABAC is serial inside a slice); pictures are independent and go to separate threads.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/dryv_cabac_host.h"
#include "levels_record.h"

namespace {

#include "cabac_tables.inc"

// ---- bit reader over an RBSP (emulation prevention bytes already removed) ---------------------------------------
struct Bits {
  const uint8_t* p;
  size_t n, pos = 0;  // pos in bits
  bool bad = false;
  Bits(const uint8_t* d, size_t len) : p(d), n(len * 8) {}
  int bit() {
    if (pos >= n) {
      bad = true;
      return 0;
    }
    const int b = (p[pos >> 3] >> (7 - (pos & 7))) & 1;
    pos++;
    return b;
  }
  uint32_t u(int k) {
    uint32_t v = 0;
    while (k-- > 0) v = (v << 1) | (uint32_t)bit();
    return v;
  }
  uint32_t ue() {
    int z = 0;
    while (!bit() && !bad && z < 32) z++;
    if (z >= 32) {  // not a valid Exp-Golomb code of this syntax (and 1u << 32 is undefined)
      bad = true;
      return 0;
    }
    return z == 0 ? 0 : ((1u << z) - 1 + u(z));
  }
  int32_t se() {
    const uint32_t k = ue();
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
  }
  // 7.2 more_rbsp_data(): anything before the last set bit (the stop bit) is data
  bool more_rbsp_data() const {
    size_t last = n;
    while (last > 0 && !((p[(last - 1) >> 3] >> (7 - ((last - 1) & 7))) & 1)) last--;
    return last > 0 && pos < last - 1;
  }
};

struct Nal {
  int ref_idc, type;
  std::vector<uint8_t> rbsp;
};

void push_rbsp(int header_byte, const uint8_t* d, size_t n, std::vector<Nal>& out);

// Annex B: NAL units separated by 00 00 01 start codes; 00 00 03 -> 00 00 inside a unit
void split_nals(const uint8_t* d, size_t len, std::vector<Nal>& out) {
  size_t i = 0;
  auto start_at = [&](size_t k) { return k + 2 < len && d[k] == 0 && d[k + 1] == 0 && d[k + 2] == 1; };
  while (i + 2 < len && !start_at(i)) i++;
  while (i + 3 < len) {
    i += 3;  // past the start code
    size_t e = i;
    while (e < len && !start_at(e)) e++;
    size_t end = e;
    while (end > i && d[end - 1] == 0) end--;  // trailing zero bytes belong to the next start code
    if (end > i) push_rbsp(d[i], d + i + 1, end - i - 1, out);
    i = e;
  }
}

// one NAL unit (header byte + payload with emulation prevention) -> Nal with the RBSP
void push_rbsp(int header_byte, const uint8_t* d, size_t n, std::vector<Nal>& out) {
  Nal nal;
  nal.ref_idc = (header_byte >> 5) & 3;
  nal.type = header_byte & 31;
  nal.rbsp.reserve(n);
  int zeros = 0;
  for (size_t k = 0; k < n; k++) {
    if (zeros >= 2 && d[k] == 3) {
      zeros = 0;
      continue;
    }
    nal.rbsp.push_back(d[k]);
    zeros = d[k] == 0 ? zeros + 1 : 0;
  }
  out.push_back(std::move(nal));
}

// ---- MP4 / QuickTime container (ISO/IEC 14496-12, avcC per 14496-15): the video track's samples as NAL units -------
// What the reference does in src/video/atom/** (atom tree), src/video/atom/avcc/mod.rs:26-46 (avcC) and
// src/video/sample/mod.rs:74-110 (stco / stsc / stsz -> sample bytes), src/video/sample/nal.rs:230-253 (length-prefixed
// NAL units). Only what the path needs: the first 'vide' track with an avc1 sample entry, every sample's NAL units.
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint64_t be64(const uint8_t* p) { return ((uint64_t)be32(p) << 32) | be32(p + 4); }

struct BoxIter {  // children of a container: [p, end)
  const uint8_t* p;
  const uint8_t* end;
  bool next(uint32_t& type, const uint8_t*& body, const uint8_t*& body_end) {
    if (end - p < 8) return false;
    uint64_t size = be32(p);
    type = be32(p + 4);
    size_t hdr = 8;
    if (size == 1) {
      if (end - p < 16) return false;
      size = be64(p + 8);
      hdr = 16;
    } else if (size == 0) {
      size = (uint64_t)(end - p);
    }
    if (size < hdr || size > (uint64_t)(end - p)) return false;
    body = p + hdr;
    body_end = p + size;
    p += size;
    return true;
  }
};
constexpr uint32_t fourcc(char a, char b, char c, char d) {
  return ((uint32_t)(uint8_t)a << 24) | ((uint32_t)(uint8_t)b << 16) | ((uint32_t)(uint8_t)c << 8) | (uint32_t)(uint8_t)d;
}
bool find_box(const uint8_t* p, const uint8_t* end, uint32_t want, const uint8_t*& body, const uint8_t*& body_end) {
  BoxIter it{p, end};
  uint32_t t;
  while (it.next(t, body, body_end))
    if (t == want) return true;
  return false;
}

bool looks_like_mp4(const uint8_t* d, size_t len) { return len >= 12 && be32(d + 4) == fourcc('f', 't', 'y', 'p'); }

// -> DRYV_OK and the track's NAL units (SPS, PPS from avcC, then every sample's units), or an error code
int collect_nals_mp4(const uint8_t* d, size_t len, std::vector<Nal>& out) {
  const uint8_t *moov, *moov_end;
  if (!find_box(d, d + len, fourcc('m', 'o', 'o', 'v'), moov, moov_end)) return DRYV_ERR_ARG;
  BoxIter traks{moov, moov_end};
  uint32_t t;
  const uint8_t *b, *e;
  while (traks.next(t, b, e)) {
    if (t != fourcc('t', 'r', 'a', 'k')) continue;
    const uint8_t *mdia, *mdia_e, *hdlr, *hdlr_e, *minf, *minf_e, *stbl, *stbl_e, *stsd, *stsd_e;
    if (!find_box(b, e, fourcc('m', 'd', 'i', 'a'), mdia, mdia_e)) continue;
    if (!find_box(mdia, mdia_e, fourcc('h', 'd', 'l', 'r'), hdlr, hdlr_e) || hdlr_e - hdlr < 12 ||
        be32(hdlr + 8) != fourcc('v', 'i', 'd', 'e'))
      continue;
    if (!find_box(mdia, mdia_e, fourcc('m', 'i', 'n', 'f'), minf, minf_e) ||
        !find_box(minf, minf_e, fourcc('s', 't', 'b', 'l'), stbl, stbl_e) ||
        !find_box(stbl, stbl_e, fourcc('s', 't', 's', 'd'), stsd, stsd_e) || stsd_e - stsd < 8)
      return DRYV_ERR_ARG;
    // stsd: version/flags, entry_count, then sample entries; avc1 = 78 bytes of VisualSampleEntry fields, then boxes
    const uint8_t *avc1, *avc1_e, *avcc, *avcc_e;
    if (!find_box(stsd + 8, stsd_e, fourcc('a', 'v', 'c', '1'), avc1, avc1_e)) return DRYV_ERR_UNSUPPORTED;
    if (avc1_e - avc1 < 78 || !find_box(avc1 + 78, avc1_e, fourcc('a', 'v', 'c', 'C'), avcc, avcc_e) || avcc_e - avcc < 7)
      return DRYV_ERR_ARG;
    const int nal_len_size = (avcc[4] & 3) + 1;
    const uint8_t* q = avcc + 5;
    for (int pass = 0; pass < 2; pass++) {  // SPS list (count in the low 5 bits), then PPS list
      if (q >= avcc_e) return DRYV_ERR_ARG;
      int cnt = pass == 0 ? (*q & 31) : *q;
      q++;
      for (int i = 0; i < cnt; i++) {
        if (avcc_e - q < 2) return DRYV_ERR_ARG;
        const size_t n = ((size_t)q[0] << 8) | q[1];
        q += 2;
        if ((size_t)(avcc_e - q) < n || n < 1) return DRYV_ERR_ARG;
        push_rbsp(q[0], q + 1, n - 1, out);
        q += n;
      }
    }
    // sample table: sizes (stsz), chunk offsets (stco / co64), samples per chunk (stsc)
    const uint8_t *stsz, *stsz_e, *stsc, *stsc_e, *stco, *stco_e;
    if (!find_box(stbl, stbl_e, fourcc('s', 't', 's', 'z'), stsz, stsz_e) || stsz_e - stsz < 12 ||
        !find_box(stbl, stbl_e, fourcc('s', 't', 's', 'c'), stsc, stsc_e) || stsc_e - stsc < 8)
      return DRYV_ERR_ARG;
    bool co64 = false;
    if (!find_box(stbl, stbl_e, fourcc('s', 't', 'c', 'o'), stco, stco_e)) {
      if (!find_box(stbl, stbl_e, fourcc('c', 'o', '6', '4'), stco, stco_e)) return DRYV_ERR_ARG;
      co64 = true;
    }
    if (stco_e - stco < 8) return DRYV_ERR_ARG;
    const uint32_t fixed_size = be32(stsz + 4), n_samples = be32(stsz + 8);
    if (!fixed_size && (uint64_t)(stsz_e - stsz - 12) < 4ull * n_samples) return DRYV_ERR_ARG;
    const uint32_t n_chunks = be32(stco + 4), n_stsc = be32(stsc + 4);
    if ((uint64_t)(stco_e - stco - 8) < (co64 ? 8ull : 4ull) * n_chunks || (uint64_t)(stsc_e - stsc - 8) < 12ull * n_stsc)
      return DRYV_ERR_ARG;
    uint32_t sample = 0, run = 0;
    for (uint32_t chunk = 1; chunk <= n_chunks && sample < n_samples; chunk++) {
      while (run + 1 < n_stsc && be32(stsc + 8 + 12 * (run + 1)) <= chunk) run++;  // stsc runs: first_chunk is 1-based
      const uint32_t per_chunk = n_stsc ? be32(stsc + 8 + 12 * run + 4) : 0;
      uint64_t off = co64 ? be64(stco + 8 + 8ull * (chunk - 1)) : be32(stco + 8 + 4ull * (chunk - 1));
      for (uint32_t k = 0; k < per_chunk && sample < n_samples; k++, sample++) {
        const uint64_t sz = fixed_size ? fixed_size : be32(stsz + 12 + 4ull * sample);
        if (off > len || sz > len - off) return DRYV_ERR_ARG;
        const uint8_t* sp = d + off;
        const uint8_t* se = sp + sz;
        while (se - sp > nal_len_size) {  // length-prefixed NAL units
          uint64_t n = 0;
          for (int i = 0; i < nal_len_size; i++) n = (n << 8) | sp[i];
          sp += nal_len_size;
          if (n < 1 || n > (uint64_t)(se - sp)) return DRYV_ERR_ARG;
          push_rbsp(sp[0], sp + 1, (size_t)n - 1, out);
          sp += n;
        }
        off += sz;
      }
    }
    return DRYV_OK;
  }
  return DRYV_ERR_ARG;  // no video track
}

// Annex-B byte stream or MP4 file -> NAL units
int collect_nals(const uint8_t* d, size_t len, std::vector<Nal>& out) {
  if (looks_like_mp4(d, len)) return collect_nals_mp4(d, len, out);
  split_nals(d, len, out);
  return DRYV_OK;
}

// Scaling matrix of an SPS / PPS as the reference builds it (ScalingLists::new, src/video/atom/avcc/sps.rs:207-248):
// a list that is not present, or that signals useDefaultScalingMatrixFlag, becomes the Default table of its kind
// directly (the reference has no fall-back rule A / B: SURVEY.md quirk Q6). Values in zig-zag order.
struct Matrix {
  bool present = false;
  int n8 = 0;  // 8x8 lists parsed: 2 (4:2:0), or 0 for a PPS with transform_8x8_mode_flag = 0
  uint8_t l4[6][16];
  uint8_t l8[6][64];
};
const uint8_t kDefault4Intra[16] = {6, 13, 13, 20, 20, 20, 28, 28, 28, 28, 32, 32, 32, 37, 37, 42};
const uint8_t kDefault4Inter[16] = {10, 14, 14, 20, 20, 20, 24, 24, 24, 24, 27, 27, 27, 30, 30, 34};
const uint8_t kDefault8Intra[64] = {6,  10, 10, 13, 11, 13, 16, 16, 16, 16, 18, 18, 18, 18, 18, 23, 23, 23, 23, 23, 23, 25,
                                    25, 25, 25, 25, 25, 25, 27, 27, 27, 27, 27, 27, 27, 27, 29, 29, 29, 29, 29, 29, 29, 31,
                                    31, 31, 31, 31, 31, 33, 33, 33, 33, 33, 36, 36, 36, 36, 38, 38, 38, 40, 40, 42};
const uint8_t kDefault8Inter[64] = {9,  13, 13, 15, 13, 15, 17, 17, 17, 17, 19, 19, 19, 19, 19, 21, 21, 21, 21, 21, 21, 22,
                                    22, 22, 22, 22, 22, 22, 24, 24, 24, 24, 24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 27,
                                    27, 27, 27, 27, 27, 28, 28, 28, 28, 28, 30, 30, 30, 30, 32, 32, 32, 33, 33, 35};
// scaling_list(), sps.rs:179-198 (7.3.2.1.1.1); returns useDefaultScalingMatrixFlag
bool parse_scaling_list(Bits& b, uint8_t* out, int size) {
  bool use_default = false;
  int last = 8, next = 8;
  for (int j = 0; j < size; j++) {
    if (next != 0) {
      const int delta = b.se();
      next = (last + delta + 256) % 256;
      if (next < 0) next += 256;  // malformed delta
      use_default = j == 0 && next == 0;
    }
    out[j] = (uint8_t)(next == 0 ? last : next);
    last = out[j];
  }
  return use_default;
}
void parse_matrix(Bits& b, int n_lists, Matrix& m) {
  m.present = true;
  m.n8 = n_lists - 6;
  for (int i = 0; i < n_lists && !b.bad; i++) {
    const bool intra = i < 6 ? i < 3 : ((i - 6) & 1) == 0;
    uint8_t* dst = i < 6 ? m.l4[i] : m.l8[i - 6];
    const int size = i < 6 ? 16 : 64;
    const uint8_t* def = i < 6 ? (intra ? kDefault4Intra : kDefault4Inter) : (intra ? kDefault8Intra : kDefault8Inter);
    bool use_default = true;
    if (b.u(1)) use_default = parse_scaling_list(b, dst, size);  // scaling_list_present_flag
    if (use_default) memcpy(dst, def, (size_t)size);
  }
}

struct Sps {
  bool ok = false;
  int id = 0;
  int w_mbs = 0, h_mbs = 0, log2_max_frame_num = 4, poc_type = 0, log2_max_poc_lsb = 4, delta_pic_order_always_zero = 0;
  uint32_t crop_l = 0, crop_r = 0, crop_t = 0, crop_b = 0;  // frame_crop_*_offset, in units of two luma samples (4:2:0 frames)
  Matrix matrix;  // seq_scaling_matrix
};
struct Pps {
  bool ok = false;
  int id = 0, sps_id = 0;
  int bottom_field_pic_order = 0, pic_init_qp = 26, cb_off = 0, cr_off = 0, deblocking_control = 0, redundant_pic_cnt = 0,
      transform_8x8_mode = 0;
  Matrix matrix;  // pic_scaling_matrix
};

int parse_sps(const Nal& nal, Sps& s) {
  Bits b(nal.rbsp.data(), nal.rbsp.size());
  const int profile = (int)b.u(8);
  b.u(8);
  b.u(8);
  s.id = (int)b.ue();
  if (s.id > 31) return DRYV_ERR_ARG;
  if (profile == 100 || profile == 110 || profile == 122 || profile == 244 || profile == 44 || profile == 83 || profile == 86 ||
      profile == 118 || profile == 128 || profile == 138 || profile == 139 || profile == 134 || profile == 135) {
    if (b.ue() != 1) return DRYV_ERR_UNSUPPORTED;                  // chroma_format_idc: 4:2:0 only
    if (b.ue() != 0 || b.ue() != 0) return DRYV_ERR_UNSUPPORTED;   // 8-bit only
    b.u(1);                                                        // qpprime_y_zero_transform_bypass_flag
    if (b.u(1)) parse_matrix(b, 8, s.matrix);                      // seq_scaling_matrix_present_flag, sps.rs:89-93
  }
  s.log2_max_frame_num = (int)b.ue() + 4;
  s.poc_type = (int)b.ue();
  if (s.poc_type == 0) {
    s.log2_max_poc_lsb = (int)b.ue() + 4;
  } else if (s.poc_type == 1) {
    s.delta_pic_order_always_zero = (int)b.u(1);
    b.se();
    b.se();
    const uint32_t cyc = b.ue();
    for (uint32_t i = 0; i < cyc && !b.bad; i++) b.se();
  }
  b.ue();  // max_num_ref_frames
  b.u(1);  // gaps_in_frame_num_value_allowed_flag
  s.w_mbs = (int)b.ue() + 1;
  s.h_mbs = (int)b.ue() + 1;
  if (!b.u(1)) return DRYV_ERR_UNSUPPORTED;  // frame_mbs_only_flag
  if (b.bad || s.w_mbs > 1024 || s.h_mbs > 1024) return DRYV_ERR_ARG;
  b.u(1);  // direct_8x8_inference_flag
  if (b.u(1)) {  // frame_cropping_flag, atom/avcc/sps.rs:252-267: reconstruction ignores it (dryv never crops), dryv_cabac_surface reports it
    s.crop_l = b.ue();
    s.crop_r = b.ue();
    s.crop_t = b.ue();
    s.crop_b = b.ue();
  }
  if (b.bad) return DRYV_ERR_ARG;
  s.ok = true;  // the VUI is not needed
  return DRYV_OK;
}

int parse_pps(const Nal& nal, Pps& p) {
  Bits b(nal.rbsp.data(), nal.rbsp.size());
  p.id = (int)b.ue();
  p.sps_id = (int)b.ue();
  if (p.id > 255 || p.sps_id > 31) return DRYV_ERR_ARG;
  if (!b.u(1)) return DRYV_ERR_UNSUPPORTED;  // entropy_coding_mode_flag: CABAC only (the reference has no CAVLC)
  p.bottom_field_pic_order = (int)b.u(1);
  if (b.ue() != 0) return DRYV_ERR_UNSUPPORTED;  // slice groups
  b.ue();
  b.ue();
  b.u(1);
  b.u(2);
  p.pic_init_qp = 26 + b.se();
  b.se();
  p.cb_off = p.cr_off = b.se();
  p.deblocking_control = (int)b.u(1);
  b.u(1);  // constrained_intra_pred_flag: irrelevant inside an I picture
  p.redundant_pic_cnt = (int)b.u(1);
  p.transform_8x8_mode = 0;
  if (b.more_rbsp_data()) {  // ExtraRbspData, pps.rs:68-88
    p.transform_8x8_mode = (int)b.u(1);
    if (b.u(1)) parse_matrix(b, 6 + 2 * p.transform_8x8_mode, p.matrix);  // pic_scaling_matrix_present_flag
    p.cr_off = b.se();
  }
  if (b.bad) return DRYV_ERR_ARG;
  p.ok = true;
  return DRYV_OK;
}

// ---- 9.3.3.2 arithmetic decoding engine + 9.3.1.1 context initialisation -----------------------------------------
// codIOffset is kept scaled: `value` = codIOffset * 2^look + the next `look` bits of the stream, so "offset >= range"
// is value >= range << look, renormalising by s bits is look -= s, and the stream is read a byte at a time instead of a
// bit at a time (the arithmetic is that of 9.3.3.2.2 - 9.3.3.2.4 exactly; src/video/cabac/mod.rs:1207-1278 reads bits).
// The arithmetic decoder's scalars live apart from the context states: a loop that keeps a local copy of the core has
// them in registers across the byte stores to the state table (which may alias anything else).
struct CabacCore {
  const uint8_t* p = nullptr;    // next byte of the slice data
  const uint8_t* end = nullptr;
  size_t overrun = 0;            // zero bytes fed past the end of the data
  uint32_t range = 510;
  uint64_t value = 0;
  int look = 0;
  inline void refill() {  // keeps 8 <= look <= 23: one operation consumes at most 6 bits
    if (look < 8) {
      uint32_t two = 0;
      if (end - p >= 2) {
        two = ((uint32_t)p[0] << 8) | p[1];
        p += 2;
      } else {
        for (int i = 0; i < 2; i++) {
          two <<= 8;
          if (p < end) two |= *p++;
          else overrun++;
        }
      }
      value = (value << 16) | two;
      look += 16;
    }
  }
  // 9.3.3.2.1 without a data-dependent branch: the LPS / MPS outcome is a mask, the renormalisation shift a bit count.
  // `state`: (pStateIdx << 1) | valMPS per context; `next`: [0..127] state after an MPS, [128..255] after an LPS.
  inline int decision(uint8_t* state, const uint8_t* next, int ctx) {
    const uint32_t st = state[ctx];
    const uint32_t lps = kRangeTabLps[(st >> 1) * 4 + ((range >> 6) & 3)];
    range -= lps;
    const uint64_t scaled = (uint64_t)range << look;
    const uint32_t is_lps = value >= scaled;
    const uint64_t mask = 0ull - (uint64_t)is_lps;
    value -= scaled & mask;
    range = is_lps ? lps : range;
    state[ctx] = next[st + (is_lps << 7)];
    const int sh = __builtin_clz(range) - 23;  // 0 when range is already in [256, 511]
    range <<= sh;
    look -= sh;
    refill();
    return (int)((st & 1u) ^ is_lps);
  }
  inline int bypass() {
    look -= 1;
    const uint64_t scaled = (uint64_t)range << look;
    const uint32_t bin = value >= scaled;
    value -= scaled & (0ull - (uint64_t)bin);
    refill();
    return (int)bin;
  }
  inline int terminate() {
    range -= 2;
    if (value >= ((uint64_t)range << look)) return 1;
    const int sh = __builtin_clz(range) - 23;
    range <<= sh;
    look -= sh;
    refill();
    return 0;
  }
};

struct Cabac {
  CabacCore k;
  uint8_t state[1024];  // (pStateIdx << 1) | valMPS
  uint8_t next[256];
  // bits consumed beyond the end of the data (0 for a well-formed slice)
  bool overran() const { return k.overrun * 8 > (size_t)k.look; }
  void init(const uint8_t* data, const uint8_t* data_end, int slice_qp) {
    k = CabacCore();
    k.p = data;
    k.end = data_end;
    const int q = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
    for (int i = 0; i < 1024; i++) {
      int pre = ((kCtxInitM[i] * q) >> 4) + kCtxInitN[i];
      pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
      state[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);
    }
    for (int st = 0; st < 128; st++) {
      const int ps = st >> 1, m = st & 1;
      next[st] = (uint8_t)((kTransIdxMps[ps] << 1) | m);
      next[128 + st] = (uint8_t)((kTransIdxLps[ps] << 1) | (ps == 0 ? !m : m));
    }
    k.look = -9;  // the first nine bits are codIOffset itself
    k.refill();
    k.refill();
  }
  int decision(int ctx) { return k.decision(state, next, ctx); }
  int bypass() { return k.bypass(); }
  int terminate() { return k.terminate(); }
};

// What the context selection of later macroblocks needs to know about a parsed macroblock
struct MbCtx {
  uint8_t avail = 0, i16 = 0, t8 = 0, cbp_luma = 0, cbp_chroma = 0, chroma_mode = 0, cbf_dc = 0;
  uint8_t cbf_cdc[2] = {0, 0};
  uint8_t cbf_cac[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  uint8_t cbf_luma[16] = {0};
};

const int kCbfBase[5] = {85, 89, 93, 97, 101};
const int kSigBase[6] = {105, 120, 134, 149, 152, 402};
const int kLastBase[6] = {166, 181, 195, 210, 213, 417};
const int kAbsBase[6] = {227, 237, 247, 257, 266, 426};
const int kBlkX[16] = {0, 4, 0, 4, 8, 12, 8, 12, 0, 4, 0, 4, 8, 12, 8, 12};
const int kBlkY[16] = {0, 0, 4, 4, 0, 0, 4, 4, 8, 8, 12, 12, 8, 8, 12, 12};
inline int blk4_of(int x, int y) { return 8 * (y / 8) + 4 * (x / 8) + 2 * ((y % 8) / 4) + ((x % 8) / 4); }

struct SliceParser {
  Cabac c;
  int W, H;
  std::vector<MbCtx> info;
  int qp_prev = 26;
  bool prev_delta_nonzero = false;
  bool unsupported = false;

  // 7.3.5.3.3 residual_block_cabac: `n` levels in coding order into out[0..n) (already zeroed); returns coded_block_flag
  template <int cat>
  int residual_block(int n, int16_t* out, int stride_unused, int cbf_inc, bool code_cbf) {
    (void)stride_unused;
    if (code_cbf && !c.decision(kCbfBase[cat] + cbf_inc)) return 0;
    CabacCore eng = c.k;  // registers for the length of the block
    uint8_t* const state = c.state;
    const uint8_t* const next = c.next;
    // positions of the significant coefficients, in coding order (the level loop below walks them backwards)
    uint8_t pos[64];
    int count = 0;
    bool open_end = true;  // no last_significant_coeff_flag seen: the final position is significant by inference
    for (int i = 0; i < n - 1; i++) {
      const int si = cat == 5 ? kSig8x8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
      pos[count] = (uint8_t)i;
      if (eng.decision(state, next, kSigBase[cat] + si)) {
        count++;
        const int li = cat == 5 ? kLast8x8[i] : (cat == 3 ? (i < 2 ? i : 2) : i);
        if (eng.decision(state, next, kLastBase[cat] + li)) {
          open_end = false;
          break;
        }
      }
    }
    if (open_end) pos[count++] = (uint8_t)(n - 1);
    int eq1 = 0, gt1 = 0;
    const int lim = 4 - (cat == 3 ? 1 : 0);
    for (int j = count - 1; j >= 0; j--) {
      const int ctx0 = kAbsBase[cat] + (gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4));
      int a = 0;
      if (eng.decision(state, next, ctx0)) {
        const int ctxn = kAbsBase[cat] + 5 + (gt1 < lim ? gt1 : lim);
        a = 1;
        while (a < 14 && eng.decision(state, next, ctxn)) a++;
        if (a == 14) {  // 0-th order Exp-Golomb suffix, bypass
          int k = 0;
          while (eng.bypass() && k < 24) {
            a += 1 << k;
            k++;
          }
          while (k-- > 0) a += eng.bypass() << k;
        }
        gt1++;
      } else {
        eq1++;
      }
      const int sign = eng.bypass();
      const int v = (a + 1 ^ -sign) + sign;  // sign ? -(a + 1) : a + 1
      out[pos[j]] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
    }
    c.k = eng;
    return 1;
  }

  static int cond(const MbCtx* nb, int flag_if_available) { return nb ? flag_if_available : 1; }  // intra MB: unavailable -> 1

  // one macroblock: 7.3.5 macroblock_layer for I slices
  void macroblock(int addr, uint8_t* mb_type, uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp, uint8_t* pred_syntax,
                  int16_t* coeff) {
    const int x = addr % W, y = addr / W;
    const MbCtx* A = x > 0 ? &info[addr - 1] : nullptr;
    const MbCtx* B = y > 0 ? &info[addr - W] : nullptr;
    MbCtx me;
    me.avail = 1;
    memset(pred_syntax, 0, 16);
    memset(coeff, 0, DRYV_COEFFS_PER_MB * sizeof(int16_t));
    // mb_type (9.3.2.5, Table 9-36, I slices)
    int code = 0, pred16 = 0, cbp_l = 0, cbp_c = 0;
    if (c.decision(3 + (A && A->i16 ? 1 : 0) + (B && B->i16 ? 1 : 0))) {
      if (c.terminate()) {  // I_PCM
        unsupported = true;
        return;
      }
      me.i16 = 1;
      cbp_l = c.decision(3 + 3) ? 15 : 0;
      if (c.decision(3 + 4)) cbp_c = c.decision(3 + 5) ? 2 : 1;
      pred16 = c.decision(3 + 6) << 1;
      pred16 |= c.decision(3 + 7);
      code = 1 + pred16 + 4 * cbp_c + (cbp_l ? 12 : 0);
    }
    *mb_type = (uint8_t)code;
    if (!me.i16) {
      me.t8 = transform8 ? (uint8_t)c.decision(399 + (A && A->t8 ? 1 : 0) + (B && B->t8 ? 1 : 0)) : 0;
      for (int k = 0; k < (me.t8 ? 4 : 16); k++) {
        int syn = c.decision(68) << 3;
        if (!syn) {
          syn |= c.decision(69);
          syn |= c.decision(69) << 1;
          syn |= c.decision(69) << 2;
        }
        pred_syntax[k] = (uint8_t)syn;
      }
    }
    *t8x8 = me.t8;
    // intra_chroma_pred_mode: TU, cMax 3
    {
      int cm = 0;
      if (c.decision(64 + (A && A->chroma_mode ? 1 : 0) + (B && B->chroma_mode ? 1 : 0))) {
        cm = 1;
        if (c.decision(64 + 3)) {
          cm = 2;
          if (c.decision(64 + 3)) cm = 3;
        }
      }
      me.chroma_mode = (uint8_t)cm;
      *chroma_mode = (uint8_t)cm;
    }
    if (!me.i16) {  // coded_block_pattern (9.3.3.1.1.4)
      for (int b8 = 0; b8 < 4; b8++) {
        const int x8 = b8 & 1, y8 = b8 >> 1;
        const int ca = x8 ? !((cbp_l >> (b8 - 1)) & 1) : (A ? !((A->cbp_luma >> (b8 + 1)) & 1) : 0);
        const int cb = y8 ? !((cbp_l >> (b8 - 2)) & 1) : (B ? !((B->cbp_luma >> (b8 + 2)) & 1) : 0);
        cbp_l |= c.decision(73 + ca + 2 * cb) << b8;
      }
      if (c.decision(77 + (A && A->cbp_chroma ? 1 : 0) + 2 * (B && B->cbp_chroma ? 1 : 0)))
        cbp_c = c.decision(77 + 4 + (A && A->cbp_chroma == 2 ? 1 : 0) + 2 * (B && B->cbp_chroma == 2 ? 1 : 0)) ? 2 : 1;
    }
    me.cbp_luma = (uint8_t)cbp_l;
    me.cbp_chroma = (uint8_t)cbp_c;
    if (me.i16 || cbp_l || cbp_c) {
      // mb_qp_delta (9.3.2.7 / 9.3.3.1.1.5): unary, mapped
      int v = 0, ctx = 60 + (prev_delta_nonzero ? 1 : 0);
      while (c.decision(ctx) && v < 104) {
        v++;
        ctx = v == 1 ? 60 + 2 : 60 + 3;
      }
      const int delta = (v & 1) ? (v + 1) / 2 : -(v / 2);
      prev_delta_nonzero = delta != 0;
      qp_prev = ((qp_prev + delta + 52) % 52);  // cabac/mod.rs:186-191 with QpBdOffsetY = 0
      // residual (7.3.5.3): luma DC, luma blocks, chroma DC, chroma AC
      if (me.i16) {
        int16_t dc[16] = {0};
        const int inc = cond(A, A && A->i16 ? A->cbf_dc : 0) + 2 * cond(B, B && B->i16 ? B->cbf_dc : 0);
        me.cbf_dc = (uint8_t)residual_block<0>(16, dc, 0, inc, true);
        for (int b = 0; b < 16; b++) coeff[b * 16] = dc[b];
      }
      for (int b8 = 0; b8 < 4; b8++) {
        if (!((cbp_l >> b8) & 1)) continue;
        if (me.t8) {
          residual_block<5>(64, coeff + b8 * 64, 0, 0, false);
          for (int k = 0; k < 4; k++) me.cbf_luma[4 * b8 + k] = 1;
        } else {
          for (int k = 0; k < 4; k++) {
            const int blk = 4 * b8 + k, bx = kBlkX[blk], by = kBlkY[blk];
            const int fa = bx > 0 ? me.cbf_luma[blk4_of(bx - 4, by)] : cond(A, A ? A->cbf_luma[blk4_of(12, by)] : 0);
            const int fb = by > 0 ? me.cbf_luma[blk4_of(bx, by - 4)] : cond(B, B ? B->cbf_luma[blk4_of(bx, 12)] : 0);
            me.cbf_luma[blk] = me.i16 ? (uint8_t)residual_block<1>(15, coeff + blk * 16 + 1, 0, fa + 2 * fb, true)
                                      : (uint8_t)residual_block<2>(16, coeff + blk * 16, 0, fa + 2 * fb, true);
          }
        }
      }
      if (cbp_c) {
        for (int pl = 0; pl < 2; pl++) {
          int16_t dc[4] = {0, 0, 0, 0};
          const int inc = cond(A, A ? A->cbf_cdc[pl] : 0) + 2 * cond(B, B ? B->cbf_cdc[pl] : 0);
          me.cbf_cdc[pl] = (uint8_t)residual_block<3>(4, dc, 0, inc, true);
          for (int k = 0; k < 4; k++) coeff[(16 + 4 * pl + k) * 16] = dc[k];
        }
      }
      if (cbp_c == 2) {
        for (int pl = 0; pl < 2; pl++)
          for (int k = 0; k < 4; k++) {
            const int fa = (k & 1) ? me.cbf_cac[pl][k - 1] : cond(A, A ? A->cbf_cac[pl][k + 1] : 0);
            const int fb = (k >> 1) ? me.cbf_cac[pl][k - 2] : cond(B, B ? B->cbf_cac[pl][k + 2] : 0);
            me.cbf_cac[pl][k] = (uint8_t)residual_block<4>(15, coeff + (16 + 4 * pl + k) * 16 + 1, 0, fa + 2 * fb, true);
          }
      }
    } else {
      prev_delta_nonzero = false;
    }
    *qp = (uint8_t)qp_prev;
    info[addr] = me;
  }

  int transform8 = 0;
};

// One IDR picture with the parameter sets in force when its slice arrived (copies: a set re-sent later does not reach back)
struct Picture {
  const Nal* nal;
  Sps sps;
  Pps pps;
};
struct Stream {
  std::vector<Picture> pic;
};

// Parameter sets are kept by id and activated through the slice header's pic_parameter_set_id -> seq_parameter_set_id
// (7.4.1.2.1), also when they arrive between pictures. (The reference reads the sets of the avcC box only and decodes the
// first sample only, src/video/decoder.rs:86-150: for that picture this is the same thing.) Every picture of a stream
// must have the geometry of the first.
int analyse(const std::vector<Nal>& nals, Stream& st) {
  std::vector<Sps> sps(32);
  std::vector<Pps> pps(256);
  for (const Nal& nal : nals) {
    if (nal.type == 7) {
      Sps s;
      const int rc = parse_sps(nal, s);
      if (rc != DRYV_OK) return rc;
      sps[(size_t)s.id] = s;
    } else if (nal.type == 8) {
      Pps p;
      const int rc = parse_pps(nal, p);
      if (rc != DRYV_OK) return rc;
      pps[(size_t)p.id] = p;
    } else if (nal.type == 5) {
      Bits b(nal.rbsp.data(), nal.rbsp.size());
      b.ue();  // first_mb_in_slice
      b.ue();  // slice_type
      const uint32_t pid = b.ue();
      if (b.bad || pid > 255 || !pps[pid].ok || !sps[(size_t)pps[pid].sps_id].ok) return DRYV_ERR_ARG;
      Picture pc{&nal, sps[(size_t)pps[pid].sps_id], pps[pid]};
      if (!st.pic.empty() && (pc.sps.w_mbs != st.pic[0].sps.w_mbs || pc.sps.h_mbs != st.pic[0].sps.h_mbs))
        return DRYV_ERR_UNSUPPORTED;  // a batch has one geometry
      st.pic.push_back(pc);
    } else if (nal.type == 1) {
      return DRYV_ERR_UNSUPPORTED;  // non-IDR pictures: the path reconstructs IDR pictures only
    }
  }
  if (st.pic.empty()) return DRYV_ERR_ARG;
  return DRYV_OK;
}

// dryv_pic_params of one picture. Scaling lists as the reference picks them (SliceHeader::scaling_lists,
// src/video/slice/header.rs:317-332): the SPS matrix if the SPS has one, else the PPS matrix, else Flat_16; the kernels
// take list 0 of each size (Intra Y: Frame::scaling idx 0 for luma, and dryv dequantises chroma with the luma list, quirk
// Q1). A PPS matrix without 8x8 lists (transform_8x8_mode_flag = 0) makes the reference index an empty list
// (frame/transform.rs:48): unsupported.
int picture_params(const Picture& pc, dryv_pic_params* pp) {
  memset(pp, 0, sizeof *pp);
  pp->pic_width_in_mbs = (uint16_t)pc.sps.w_mbs;
  pp->pic_height_in_mbs = (uint16_t)pc.sps.h_mbs;
  pp->chroma_qp_index_offset = (int8_t)pc.pps.cb_off;
  pp->second_chroma_qp_index_offset = (int8_t)pc.pps.cr_off;
  const Matrix* m = pc.sps.matrix.present ? &pc.sps.matrix : (pc.pps.matrix.present ? &pc.pps.matrix : nullptr);
  if (!m) {
    memset(pp->scaling_list4x4, 16, sizeof pp->scaling_list4x4);  // Flat_4x4_16 / Flat_8x8_16
    memset(pp->scaling_list8x8, 16, sizeof pp->scaling_list8x8);
    return DRYV_OK;
  }
  if (m->n8 < 1) return DRYV_ERR_UNSUPPORTED;
  memcpy(pp->scaling_list4x4, m->l4[0], 16);
  memcpy(pp->scaling_list8x8, m->l8[0], 64);
  return DRYV_OK;
}

// Compact sink of one picture: the records of its macroblocks back to back, and their sizes
struct CompactPicture {
  std::vector<uint8_t> stream;
  std::vector<uint32_t> size;
};

struct SliceInfo {
  int slice_qp = 26, disable_deblocking_filter_idc = 0, alpha_div2 = 0, beta_div2 = 0, pps_id = 0;
};

// slice_header (7.3.3) of an IDR picture; leaves `b` at the first byte of slice_data
int parse_slice_header(const Picture& st, Bits& b, SliceInfo& si) {
  const Nal& nal = *st.nal;
  if (b.ue() != 0) return DRYV_ERR_UNSUPPORTED;  // first_mb_in_slice: one slice per picture
  const uint32_t slice_type = b.ue();
  if (slice_type != 2 && slice_type != 7) return DRYV_ERR_UNSUPPORTED;
  si.pps_id = (int)b.ue();            // pic_parameter_set_id (resolved by analyse)
  b.u(st.sps.log2_max_frame_num);     // frame_num
  b.ue();                             // idr_pic_id
  if (st.sps.poc_type == 0) {
    b.u(st.sps.log2_max_poc_lsb);
    if (st.pps.bottom_field_pic_order) b.se();
  } else if (st.sps.poc_type == 1 && !st.sps.delta_pic_order_always_zero) {
    b.se();
    if (st.pps.bottom_field_pic_order) b.se();
  }
  if (st.pps.redundant_pic_cnt) b.ue();
  if (nal.ref_idc != 0) b.u(2);       // dec_ref_pic_marking of an IDR picture
  si.slice_qp = st.pps.pic_init_qp + b.se();
  if (st.pps.deblocking_control) {
    // dryv has no deblocking filter (README.md:15; the fields are parsed in slice/header.rs:609-640 and ignored): a stream
    // that asks for it reconstructs to the unfiltered pictures. dryv_cabac_slice_info reports what the stream asked for.
    si.disable_deblocking_filter_idc = (int)b.ue();
    if (si.disable_deblocking_filter_idc != 1) {
      si.alpha_div2 = b.se();
      si.beta_div2 = b.se();
    }
  }
  while (b.pos & 7) b.bit();          // cabac_alignment_one_bit
  if (b.bad || si.slice_qp < 0 || si.slice_qp > 51 || si.disable_deblocking_filter_idc > 2 || si.alpha_div2 < -6 ||
      si.alpha_div2 > 6 || si.beta_div2 < -6 || si.beta_div2 > 6)
    return DRYV_ERR_ARG;
  return DRYV_OK;
}

// `coeff` (dense, 384 int16 per macroblock) or `compact` (records, include/dryv_recon.h) receives the levels
int parse_picture(const Picture& st, uint8_t* mb_type, uint8_t* t8x8, uint8_t* chroma_mode, uint8_t* qp,
                  uint8_t* pred_syntax, int16_t* coeff, CompactPicture* compact = nullptr) {
  const Nal& nal = *st.nal;
  Bits b(nal.rbsp.data(), nal.rbsp.size());
  SliceInfo si;
  const int hrc = parse_slice_header(st, b, si);
  if (hrc != DRYV_OK) return hrc;
  const int slice_qp = si.slice_qp;
  SliceParser sp;
  sp.W = st.sps.w_mbs;
  sp.H = st.sps.h_mbs;
  sp.info.assign((size_t)sp.W * sp.H, MbCtx());
  sp.transform8 = st.pps.transform_8x8_mode;
  sp.qp_prev = slice_qp;
  sp.c.init(nal.rbsp.data() + (b.pos >> 3), nal.rbsp.data() + nal.rbsp.size(), slice_qp);
  const int n = sp.W * sp.H;
  int16_t one_mb[DRYV_COEFFS_PER_MB];
  if (compact) {
    compact->stream.reserve((size_t)n * 256);
    compact->size.reserve((size_t)n);
  }
  for (int addr = 0; addr < n; addr++) {
    sp.macroblock(addr, mb_type + addr, t8x8 + addr, chroma_mode + addr, qp + addr, pred_syntax + (size_t)addr * 16,
                  compact ? one_mb : coeff + (size_t)addr * DRYV_COEFFS_PER_MB);
    if (sp.unsupported) return DRYV_ERR_UNSUPPORTED;
    if (compact) {  // the macroblock's record, straight from what residual_block decoded
      const dryv_levels::MbStats ms = dryv_levels::scan(one_mb);
      const int mode = dryv_levels::pick_mode(ms);
      const uint32_t sz = dryv_levels::record_size(ms, mode);
      const size_t at = compact->stream.size();
      compact->stream.resize(at + sz);
      dryv_levels::write_record(one_mb, ms, mode, compact->stream.data() + at, sz);
      compact->size.push_back(sz);
    }
    const int end = sp.c.terminate();  // end_of_slice_flag
    if (sp.c.overran()) return DRYV_ERR_ARG;
    if (end != (addr == n - 1)) return DRYV_ERR_ARG;  // slice ends early / runs past the picture
  }
  return DRYV_OK;
}

}  // namespace

extern "C" {

int dryv_cabac_scan(const uint8_t* annexb, size_t len, dryv_pic_params* pp, uint32_t* n_pictures) {
  if (!annexb || !pp || !n_pictures || len < 8) return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  rc = picture_params(st.pic[0], pp);
  if (rc != DRYV_OK) return rc;
  *n_pictures = (uint32_t)st.pic.size();
  return DRYV_OK;
}

int dryv_cabac_picture_params(const uint8_t* annexb, size_t len, uint32_t picture, dryv_pic_params* pp) {
  if (!annexb || !pp || len < 8) return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  if (picture >= st.pic.size()) return DRYV_ERR_ARG;
  return picture_params(st.pic[picture], pp);
}

int dryv_cabac_slice_info(const uint8_t* annexb, size_t len, uint32_t picture, dryv_slice_info* out) {
  if (!annexb || !out || len < 8) return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  if (picture >= st.pic.size()) return DRYV_ERR_ARG;
  const Picture& pc = st.pic[picture];
  Bits b(pc.nal->rbsp.data(), pc.nal->rbsp.size());
  SliceInfo si;
  rc = parse_slice_header(pc, b, si);
  if (rc != DRYV_OK) return rc;
  memset(out, 0, sizeof *out);
  out->pic_parameter_set_id = (uint8_t)si.pps_id;
  out->seq_parameter_set_id = (uint8_t)pc.sps.id;
  out->slice_qp = (uint8_t)si.slice_qp;
  out->disable_deblocking_filter_idc = (uint8_t)si.disable_deblocking_filter_idc;
  out->slice_alpha_c0_offset_div2 = (int8_t)si.alpha_div2;
  out->slice_beta_offset_div2 = (int8_t)si.beta_div2;
  out->scaling_matrix_source = pc.sps.matrix.present ? 1 : (pc.pps.matrix.present ? 2 : 0);
  return DRYV_OK;
}

int dryv_cabac_surface(const uint8_t* annexb, size_t len, dryv_surface* out) {
  if (!annexb || !out || len < 8) return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  const Sps& sps = st.pic[0].sps;
  const uint64_t W = 16ull * (uint64_t)sps.w_mbs, H = 16ull * (uint64_t)sps.h_mbs;
  const uint64_t l = 2ull * sps.crop_l, r = 2ull * sps.crop_r, t = 2ull * sps.crop_t, b = 2ull * sps.crop_b;
  if (l + r >= W || t + b >= H) return DRYV_ERR_ARG;
  out->format = DRYV_SURFACE_I420;
  out->crop_left = (uint32_t)l;
  out->crop_top = (uint32_t)t;
  out->width = (uint32_t)(W - l - r);
  out->height = (uint32_t)(H - t - b);
  return DRYV_OK;
}

int dryv_cabac_parse(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t n_pictures, uint8_t* mb_type,
                     uint8_t* transform_size_8x8_flag, uint8_t* intra_chroma_pred_mode, uint8_t* qp, uint8_t* pred_syntax,
                     int16_t* coeff, int threads) {
  return dryv_cabac_parse_range(annexb, len, pp, 0, n_pictures, 1, mb_type, transform_size_8x8_flag, intra_chroma_pred_mode, qp,
                                pred_syntax, coeff, threads);
}

int dryv_cabac_parse_range(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t first_picture,
                           uint32_t n_pictures, int must_be_all, uint8_t* mb_type, uint8_t* transform_size_8x8_flag,
                           uint8_t* intra_chroma_pred_mode, uint8_t* qp, uint8_t* pred_syntax, int16_t* coeff, int threads) {
  if (!annexb || !pp || !mb_type || !transform_size_8x8_flag || !intra_chroma_pred_mode || !qp || !pred_syntax || !coeff ||
      n_pictures == 0)
    return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  if (st.pic[0].sps.w_mbs != pp->pic_width_in_mbs || st.pic[0].sps.h_mbs != pp->pic_height_in_mbs) return DRYV_ERR_ARG;
  if ((uint64_t)first_picture + n_pictures > st.pic.size() || (must_be_all && st.pic.size() != n_pictures)) return DRYV_ERR_ARG;
  const size_t n_mb = (size_t)st.pic[0].sps.w_mbs * st.pic[0].sps.h_mbs;
  std::atomic<uint32_t> next(0);
  std::atomic<int> status(DRYV_OK);
  auto work = [&]() {
    for (;;) {
      const uint32_t f = next.fetch_add(1);
      if (f >= n_pictures) break;
      const size_t o = (size_t)f * n_mb;
      const int r = parse_picture(st.pic[first_picture + f], mb_type + o, transform_size_8x8_flag + o, intra_chroma_pred_mode + o, qp + o,
                                  pred_syntax + o * 16, coeff + o * DRYV_COEFFS_PER_MB);
      if (r != DRYV_OK) {
        int expect = DRYV_OK;
        status.compare_exchange_strong(expect, r);
      }
    }
  };
  if (threads <= 1 || n_pictures == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    const uint32_t nt = (uint32_t)threads < n_pictures ? (uint32_t)threads : n_pictures;
    for (uint32_t t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  return status.load();
}

int dryv_cabac_parse_compact(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t first_picture,
                             uint32_t n_pictures, uint8_t* mb_type, uint8_t* transform_size_8x8_flag,
                             uint8_t* intra_chroma_pred_mode, uint8_t* qp, uint8_t* pred_syntax, uint32_t* offset,
                             uint8_t* stream, size_t stream_cap, int threads) {
  if (!annexb || !pp || !mb_type || !transform_size_8x8_flag || !intra_chroma_pred_mode || !qp || !pred_syntax || !offset ||
      !stream || n_pictures == 0)
    return DRYV_ERR_ARG;
  std::vector<Nal> nals;
  int rc = collect_nals(annexb, len, nals);
  if (rc != DRYV_OK) return rc;
  Stream st;
  rc = analyse(nals, st);
  if (rc != DRYV_OK) return rc;
  if (st.pic[0].sps.w_mbs != pp->pic_width_in_mbs || st.pic[0].sps.h_mbs != pp->pic_height_in_mbs) return DRYV_ERR_ARG;
  if ((uint64_t)first_picture + n_pictures > st.pic.size()) return DRYV_ERR_ARG;
  const size_t n_mb = (size_t)st.pic[0].sps.w_mbs * st.pic[0].sps.h_mbs;
  std::vector<CompactPicture> pics(n_pictures);
  std::atomic<uint32_t> next(0);
  std::atomic<int> status(DRYV_OK);
  auto work = [&]() {
    for (;;) {
      const uint32_t f = next.fetch_add(1);
      if (f >= n_pictures) break;
      const size_t o = (size_t)f * n_mb;
      const int r = parse_picture(st.pic[first_picture + f], mb_type + o, transform_size_8x8_flag + o,
                                  intra_chroma_pred_mode + o, qp + o, pred_syntax + o * 16, nullptr, &pics[f]);
      if (r != DRYV_OK) {
        int expect = DRYV_OK;
        status.compare_exchange_strong(expect, r);
      }
    }
  };
  const uint32_t nt = threads <= 1 ? 1u : ((uint32_t)threads < n_pictures ? (uint32_t)threads : n_pictures);
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nt; t++) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  if (status.load() != DRYV_OK) return status.load();
  // offsets: a prefix sum over the record sizes; then every picture's records are copied to their place
  uint64_t acc = 0;
  std::vector<uint64_t> pic_at(n_pictures);
  offset[0] = 0;
  for (uint32_t f = 0; f < n_pictures; f++) {
    pic_at[f] = acc;
    if (pics[f].size.size() != n_mb) return DRYV_ERR_ARG;
    for (size_t i = 0; i < n_mb; i++) {
      acc += pics[f].size[i];
      if (acc > 0xffffffffull || acc > stream_cap) return DRYV_ERR_ARG;
      offset[(size_t)f * n_mb + i + 1] = (uint32_t)acc;
    }
  }
  next.store(0);
  auto copy = [&]() {
    for (;;) {
      const uint32_t f = next.fetch_add(1);
      if (f >= n_pictures) break;
      memcpy(stream + pic_at[f], pics[f].stream.data(), pics[f].stream.size());
    }
  };
  if (nt == 1) {
    copy();
  } else {
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < nt; t++) pool.emplace_back(copy);
    for (auto& th : pool) th.join();
  }
  return DRYV_OK;
}

}  // extern "C"
