/*
 * dryv_cabac_host.h — CPU host side of the reconstruction path: Annex-B H.264 bytes -> the syntax buffers of
 * include/dryv_recon.h.
 *
 * This is the CALLER of the hot path, restated in C++ for IDR I-slice pictures: what the
 * reference does in
 *   NAL split / emulation prevention     src/video/sample/nal.rs:230-253, src/byte/bit.rs:144-149
 *   SPS / PPS                            src/video/atom/avcc/sps.rs:42-121, src/video/atom/avcc/pps.rs:30-58
 *   slice header                         src/video/slice/header.rs:145-315
 *   macroblock loop                      src/video/slice/mod.rs:184-317
 *   CabacContext::macroblock_layer       src/video/cabac/mod.rs:89-210  (up to, and instead of, frame.decode at :208)
 *   residual / residual_cabac            src/video/cabac/mod.rs:433-675
 *   arithmetic decoding engine           src/video/cabac/mod.rs:1207-1308
 * with the macroblock's parsed syntax appended to structure-of-arrays buffers instead of being reconstructed on the
 * spot. It stays on the CPU, as in the reference (CABAC is serial per slice); independent IDR pictures are parsed on
 * separate host threads.
 *
 * Input (`annexb`, `len` in both calls) is either an Annex-B byte stream (start-code separated NAL units) or a whole MP4 /
 * QuickTime file, recognised by its leading ftyp box: the file `dryv <file>` opens. Of the container only what the path
 * needs is restated — the first video track's avc1/avcC sample entry (SPS, PPS, NAL length size) and its sample table
 * (stsz, stco/co64, stsc) with length-prefixed NAL units — i.e. src/video/atom/root.rs:17-52, atom/stbl.rs:367-420,
 * atom/avcc/mod.rs:26-46, src/video/sample/mod.rs:74-110 and sample/nal.rs:230-253. Unlike the reference, which decodes the
 * first sample only (src/video/decoder.rs:88), every IDR picture of the track is parsed.
 *
 * Supported, like the reference's reconstruction: 8-bit 4:2:0, frame macroblocks, CABAC, one slice per picture
 * (first_mb_in_slice == 0), I slices with I_NxN (4x4 / 8x8 transform) and I_16x16 macroblocks, SPS / PPS scaling
 * matrices with the reference's own selection and fall-back (atom/avcc/sps.rs:207-248, slice/header.rs:317-332:
 * SURVEY.md quirk Q6). Parameter sets are kept by id and activated through the slice header, also between pictures;
 * every picture of a stream must have the geometry of the first.
 * Anything else (I_PCM, inter slices, CAVLC, slice groups, ...) returns DRYV_ERR_UNSUPPORTED.
 */
#ifndef DRYV_CABAC_HOST_H
#define DRYV_CABAC_HOST_H

#include "dryv_recon.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Walks the NAL units, fills `pp` with the parameters of the FIRST IDR picture (geometry, chroma QP offsets, the
 * scaling lists the reference would use for it) and counts the IDR pictures. Returns DRYV_OK, DRYV_ERR_ARG (no
 * SPS/PPS/IDR, truncated data) or DRYV_ERR_UNSUPPORTED. */
int dryv_cabac_scan(const uint8_t* annexb, size_t len, dryv_pic_params* pp, uint32_t* n_pictures);

/* The same parameters for picture `picture` (0-based): pictures of one stream may name different picture parameter sets
 * (chroma QP offsets, scaling matrix), and a dryv_recon_submit batch shares one dryv_pic_params, so a host groups the
 * pictures by equal parameters. */
int dryv_cabac_picture_params(const uint8_t* annexb, size_t len, uint32_t picture, dryv_pic_params* pp);

/* What the slice header of picture `picture` says about things this path does not do itself. The reference has no
 * deblocking filter (README.md:15; the fields are parsed in slice/header.rs:609-640 and ignored): a consumer that wants
 * what a conformant decoder outputs calls dryv_recon_deblock_device with these offsets when
 * disable_deblocking_filter_idc != 1 (idc 2, "not across slice edges", equals 0 for one slice per picture). */
typedef struct dryv_slice_info {
  uint8_t pic_parameter_set_id, seq_parameter_set_id;
  uint8_t slice_qp;                       /* SliceQPY */
  uint8_t disable_deblocking_filter_idc;  /* 0 when the PPS has no deblocking_filter_control_present_flag */
  int8_t slice_alpha_c0_offset_div2, slice_beta_offset_div2;
  uint8_t scaling_matrix_source;          /* 0 flat, 1 SPS matrix, 2 PPS matrix (slice/header.rs:317-332) */
  uint8_t reserved;
} dryv_slice_info;
int dryv_cabac_slice_info(const uint8_t* annexb, size_t len, uint32_t picture, dryv_slice_info* out);

/* The display rectangle the stream's SPS asks for (frame_cropping_flag and frame_crop_*_offset, 7.4.2.1.1; the fields the
 * reference parses in atom/avcc/sps.rs:252-267 and never applies), as a DRYV_SURFACE_I420 surface for
 * dryv_recon_export_device / dryv_recon_set_surface: the whole coded picture when the SPS does not crop. Returns
 * DRYV_ERR_ARG if the offsets leave no picture. */
int dryv_cabac_surface(const uint8_t* annexb, size_t len, dryv_surface* out);

/* Parses every IDR picture of the stream (in stream order) into the caller's structure-of-arrays buffers, laid out as
 * dryv_mb_soa describes for `n_pictures` pictures of `pp` geometry: mb_type / transform_size_8x8_flag /
 * intra_chroma_pred_mode / qp: n_pictures * n_mb bytes each; pred_syntax: 16 bytes per macroblock; coeff: 384 int16 per
 * macroblock (dense; convert with dryv_recon_pack_levels for dryv_recon_submit_compact). `threads` pictures are parsed
 * concurrently (<= 1: on the calling thread). Returns DRYV_OK, DRYV_ERR_ARG (stream does not match pp / n_pictures,
 * bitstream ends early) or DRYV_ERR_UNSUPPORTED (mb_type I_PCM, non-I slice, ...). Needs no GPU. */
int dryv_cabac_parse(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t n_pictures,
                     uint8_t* mb_type, uint8_t* transform_size_8x8_flag, uint8_t* intra_chroma_pred_mode, uint8_t* qp,
                     uint8_t* pred_syntax, int16_t* coeff, int threads);

/* The same for pictures [first_picture, first_picture + n_pictures) of the stream only (buffers sized for n_pictures):
 * what each rank of a multi-GPU host calls for its share of the pictures (dryv_b200/shard.py), nothing is exchanged.
 * `must_be_all` != 0 additionally requires the range to be the whole stream. */
int dryv_cabac_parse_range(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t first_picture,
                           uint32_t n_pictures, int must_be_all, uint8_t* mb_type, uint8_t* transform_size_8x8_flag,
                           uint8_t* intra_chroma_pred_mode, uint8_t* qp, uint8_t* pred_syntax, int16_t* coeff, int threads);

/* The same with the levels emitted as the compact stream dryv_recon_submit_compact takes (include/dryv_recon.h,
 * dryv_mb_levels_compact) instead of dense arrays: the parser has the significance map and the levels in hand, so the
 * records are written as the blocks are decoded and no dense array is ever filled. `offset` must hold
 * n_pictures * n_mb + 1 entries, `stream` stream_cap bytes (n_pictures * n_mb * DRYV_COMPACT_MAX_RECORD always suffices);
 * offset[n_pictures * n_mb] is the number of stream bytes written. */
int dryv_cabac_parse_compact(const uint8_t* annexb, size_t len, const dryv_pic_params* pp, uint32_t first_picture,
                             uint32_t n_pictures, uint8_t* mb_type, uint8_t* transform_size_8x8_flag,
                             uint8_t* intra_chroma_pred_mode, uint8_t* qp, uint8_t* pred_syntax, uint32_t* offset,
                             uint8_t* stream, size_t stream_cap, int threads);

#ifdef __cplusplus
}
#endif
#endif /* DRYV_CABAC_HOST_H */
