/*
 * dryv_recon.h — C ABI of the B200 (sm_100a) AVC intra macroblock reconstruction path.
 *
 * This library replaces ONE path of the dryv H.264 decoder (reference = Stuff7/dryv, Rust):
 * everything under src/video/frame/ — inverse quantisation, the 4x4/8x8 inverse integer transforms,
 * Intra4x4/8x8/16x16/chroma prediction, residual add + clip, picture construction and the planar YUV
 * frame the decoder writes to ./temp/yuv_frame.
 *
 * The reference has no FFI; the seam is three Rust call sites (file:line relative to the reference):
 *   Frame::new(&slice)                       src/video/decoder.rs:124   (frame/mod.rs:29-46)
 *   frame.decode(slice)   once per MB        src/video/cabac/mod.rs:208 (frame/mod.rs:72-90)
 *   frame.write_to_yuv_file("temp/yuv_frame") src/video/decoder.rs:141-143 (frame/mod.rs:48-70)
 * A host that keeps CABAC/slice/atom parsing on the CPU appends each macroblock's parsed syntax to the
 * structure-of-arrays buffers below (instead of calling Frame::decode), then calls dryv_recon_submit
 * once per batch of independent IDR pictures. INTEGRATION.md shows the Rust binding.
 *
 * No CPU fallback exists behind this ABI: every entry point that reconstructs runs CUDA kernels and
 * fails with DRYV_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef DRYV_RECON_H
#define DRYV_RECON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRYV_RECON_ABI_VERSION 1

/* Return codes (the reference's hot path returns () and panics/todo!()s on unsupported input:
 * frame/mod.rs:85-88, pred8x8.rs:693-695; here nothing aborts across the ABI). */
enum {
  DRYV_OK = 0,
  DRYV_ERR_ARG = -1,         /* null pointer, zero size, geometry out of range */
  DRYV_ERR_UNSUPPORTED = -2, /* mb_type > 24 (I_PCM / inter), chroma mode > 3 ... */
  DRYV_ERR_CUDA = -3,        /* CUDA runtime error, no device */
  DRYV_ERR_WATCHDOG = -4     /* wavefront spin exceeded its bound (never expected) */
};

/* Coefficients per macroblock: 16 luma 4x4 blocks (or 4 luma 8x8 blocks) + 4 Cb + 4 Cr blocks of 16. */
#define DRYV_COEFFS_PER_MB 384
#define DRYV_MB_TYPE_I_NXN 0  /* slice/consts.rs:8  */
#define DRYV_MB_TYPE_I_PCM 25 /* rejected */

/* Per-batch picture parameters: every picture of a batch shares them.
 * Sources in the reference: pic_width_in_mbs / pic_height_in_mbs  slice/mod.rs:125-138;
 * chroma_qp_index_offset / second_chroma_qp_index_offset  atom/avcc/pps.rs:22,65 (the second one
 * equals the first when the PPS has no extension, frame/transform.rs:194-204);
 * scaling_list4x4[0] / scaling_list8x8[0]  slice/header.rs:317-332 (flat 16 when no matrix).
 * Only list 0 (Intra-Y) is consumed: chroma re-uses the luma LevelScale4x4 in the reference
 * (trans_chroma.rs never calls Frame::scaling) and so does this library.
 * 8-bit 4:2:0, frame macroblocks, one slice per picture (first_mb_in_slice == 0). */
typedef struct dryv_pic_params {
  uint16_t pic_width_in_mbs;             /* 1..1024 */
  uint16_t pic_height_in_mbs;            /* 1..1024 */
  int8_t chroma_qp_index_offset;         /* -12..12 (Cb) */
  int8_t second_chroma_qp_index_offset;  /* -12..12 (Cr) */
  uint8_t reserved0[2];
  uint32_t flags;                        /* must be 0 */
  uint8_t scaling_list4x4[16];           /* zig-zag order, list 0 */
  uint8_t scaling_list8x8[64];           /* zig-zag order, list 0 */
} dryv_pic_params;

/* Per-macroblock syntax, structure of arrays, macroblock raster order (= mbaddr) inside a picture,
 * picture f at element offset f * n_mb where n_mb = pic_width_in_mbs * pic_height_in_mbs.
 * Field sources: struct Macroblock, slice/macroblock.rs:21-129.
 *
 *  mb_type      dryv/H.264 I-slice code: 0 = I_NxN, 1..24 = I_16x16_<pred>_<cbpC>_<cbpL>
 *               (slice/consts.rs:8-113; pred16 = (code-1)%4, slice/macroblock.rs:682-716)
 *  transform_size_8x8_flag   0/1; with mb_type 0 selects Intra4x4 vs Intra8x8
 *  intra_chroma_pred_mode    0 DC, 1 Horizontal, 2 Vertical, 3 Plane
 *  qp           QP'Y = qp1y (cabac/mod.rs:186-191), 0..51
 *  pred_syntax  [16] per MB: bit 3 = prev_intra{4x4,8x8}_pred_mode_flag, bits 0..2 =
 *               rem_intra{4x4,8x8}_pred_mode; Intra8x8 uses entries 0..3; ignored for Intra16x16
 *  coeff        [24][16] int16 per MB, every block in coefficient (zig-zag) order as CABAC stores it:
 *                 Intra4x4 : block b (0..15, spec 4x4 block order) = block_luma_4x4[0][b][0..16]
 *                 Intra8x8 : the 256 luma slots = block_luma_8x8[0][b8][0..64], b8 = 0..3
 *                 Intra16x16: slot [b][0] = block_luma_dc[0][b] (the b-th entry of the DC list),
 *                             slots [b][1..16] = block_luma_ac[0][b][0..15]
 *                 chroma   : block 16+b (Cb) / 20+b (Cr), b = 0..3: slot [0] = block_chroma_dc[iCbCr][b],
 *                             slots [1..16] = block_chroma_ac[iCbCr][b][0..15]
 *               768 bytes per MB, 16-byte aligned.
 *               Arithmetic range: the kernels compute in int32 where the reference computes in isize. Every level a
 *               conforming stream can carry is exact (8.5.12.1: dequantised coefficients and transform intermediates fit
 *               16 + bit-depth bits); results are identical to the reference's for any input with
 *               |level| * LevelScale(qP, i, j) << max(qP/6 - 4, 0) below 2^26 (flat lists: every int16 level up to qP 41,
 *               |level| < 9000 at qP 51; tested at +-2047). Beyond that (large levels against large custom scaling-list entries at
 *               high qP) int32 wraps where isize does not and the output is unspecified.
 * Reconstruction never reads coded_block_pattern: absent blocks are all-zero arrays (cabac/mod.rs:669-673).
 */
typedef struct dryv_mb_soa {
  const uint8_t* mb_type;
  const uint8_t* transform_size_8x8_flag;
  const uint8_t* intra_chroma_pred_mode;
  const uint8_t* qp;
  const uint8_t* pred_syntax; /* 16 bytes per MB */
  const int16_t* coeff;       /* DRYV_COEFFS_PER_MB int16 per MB */
} dryv_mb_soa;

/* Bytes of one reconstructed picture: Y (W*H) then Cb (W/2*H/2) then Cr, row-major, macroblock
 * aligned, no cropping — byte-identical to what Frame::write_to_yuv_file emits (frame/mod.rs:48-70). */
size_t dryv_recon_frame_bytes(const dryv_pic_params* pp);

typedef struct dryv_recon_ctx dryv_recon_ctx;

int dryv_recon_abi_version(void);

/* One context per CUDA device (owns device buffers, streams, wavefront progress counters).
 * Not thread-safe: one submitting thread per context; distinct contexts are independent. */
int dryv_recon_create(int device, dryv_recon_ctx** out);
void dryv_recon_destroy(dryv_recon_ctx* ctx);
const char* dryv_recon_last_error(dryv_recon_ctx* ctx);

/* Pinned (page-locked) host memory for the SoA buffers and the output frames. */
int dryv_recon_alloc_pinned(size_t bytes, void** out);
void dryv_recon_free_pinned(void* p);

/* Replaces Frame::new + n_mb x Frame::decode (+ the planes write_to_yuv_file serialises) for
 * n_frames independent IDR pictures. HOST pointers in `soa` / `out_yuv` (pinned for full speed).
 * Asynchronous on the context's streams: H2D copy, kernels, D2H copy are pipelined in chunks of
 * pictures; call dryv_recon_wait before reading out_yuv or reusing the inputs. */
int dryv_recon_submit(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                      uint32_t n_frames, uint8_t* out_yuv);
int dryv_recon_wait(dryv_recon_ctx* ctx);

/* Streaming use: dryv_recon_submit / dryv_recon_submit_compact may be called again before the previous batch
 * has been waited for (up to 4 outstanding; a fifth call blocks on the oldest). The batches run in order and
 * the next batch's H2D copies overlap the previous batch's D2H copies. dryv_recon_wait_oldest returns when the
 * oldest outstanding batch's pictures are complete in its out_yuv (which, like its inputs, must stay valid and
 * untouched until then); it returns DRYV_OK at once when nothing is outstanding. Device-side status
 * (DRYV_ERR_UNSUPPORTED, DRYV_ERR_WATCHDOG) is collected by dryv_recon_wait, which waits for everything.
 * dryv_recon_last_submit_ms then spans from the first H2D of the first batch queued since the previous
 * dryv_recon_wait to the last D2H of the last one. */
int dryv_recon_wait_oldest(dryv_recon_ctx* ctx);

/* Same computation with DEVICE pointers (inputs already resident in HBM, output left in HBM),
 * enqueued on `cuda_stream` (a cudaStream_t; NULL = the context's own stream). Completion and
 * device-side status are collected by dryv_recon_wait (which waits for every launch of the context).
 * Independent batches enqueued on DIFFERENT streams overlap on the device (up to four in flight per
 * context): the start-up stagger of one batch's wavefront fills the tail of the previous one, which is
 * worth 1.2 x on back-to-back 64-picture 1080p batches and 2 - 3 x on 8- to 16-picture ones. */
int dryv_recon_reconstruct_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp,
                                  const dryv_mb_soa* d_soa, uint32_t n_frames, uint8_t* d_out_yuv,
                                  void* cuda_stream);

/* BASELINE config "dequant + 4x4/8x8 IDCT + residual add only": no intra prediction, no wavefront.
 * out = clip(pred + residual) where `d_pred_yuv` is a caller-supplied prediction picture in the output
 * layout. Covers frame/transform.rs:116-191, pred8x8.rs:51-150, pred16x16.rs:428-482,
 * trans_chroma.rs:369-456 and the clip/picture-construction of frame/mod.rs:93-165. DEVICE pointers. */
int dryv_recon_residual_add_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp,
                                   const dryv_mb_soa* d_soa, uint32_t n_frames,
                                   const uint8_t* d_pred_yuv, uint8_t* d_out_yuv, void* cuda_stream);

/* ---------------------------------------------------------------------------------------------------
 * Compact level stream: the wire format between the CABAC host and the GPU.
 *
 * CABAC hands the reference a significance map plus the non-zero levels of each block and the reference
 * scatters them into dense zero-filled arrays (residual_cabac, cabac/mod.rs:563-675: significant_coeff_flag map :632-646, levels :653-667; absent blocks
 * stay all-zero, :669-673). Shipping those dense arrays over PCIe moves mostly zeros (768 B per MB), and the
 * host-buffer path is PCIe bound, so dryv_recon_submit_compact takes the levels the way CABAC produced
 * them and a small kernel re-creates the dense `coeff` layout of dryv_mb_soa in HBM.
 *
 * One record per macroblock, records in macroblock order, each starting on a 4-byte boundary:
 *   uint32 header   bits 0..23: slot b (the b-th group of 16 int16 of the MB's `coeff` array, b = 0..23)
 *                               holds at least one non-zero level;  bits 30..31: level coding (below);
 *                               bits 24..29 must be 0
 *   uint16 mask[n]  one per coded slot, ascending b: bit k = coefficient k of the slot is non-zero
 *   levels          the non-zero levels of all coded slots, ascending (b, k), in one of three codings:
 *                     0  int8 per level
 *                     1  one 4-bit code per level, two per byte, low nibble first, padded to a multiple of
 *                        2 bytes (bit 3 = sign, bits 0..2 = |level| 1..7; 0 = escape), followed by the
 *                        escaped levels as int16 little endian in the same order
 *                     2  int16 little endian per level
 *   zero padding to the next multiple of 4
 * (dryv_recon_pack_levels picks the shortest coding per macroblock: at QP 26, where 98 % of the non-zero
 * levels are within +-7, a macroblock's levels take ~120 bytes instead of 768.)
 * offset[i] is the byte offset of macroblock i's record inside `stream`; offset[n_mbs] is the stream size.
 * A record is at most DRYV_COMPACT_MAX_RECORD bytes; a stream is at most 4 GiB - 1.
 */
#define DRYV_COMPACT_MAX_RECORD (4 + 24 * 2 + DRYV_COEFFS_PER_MB * 2)

typedef struct dryv_mb_levels_compact {
  const uint32_t* offset; /* [n_mbs + 1] */
  const uint8_t* stream;
} dryv_mb_levels_compact;

/* Host-side converter dense -> compact for callers that already hold dense arrays (tests, bench); a CABAC
 * host appends records directly. `offset` must hold n_mbs + 1 entries, `stream` stream_cap bytes
 * (n_mbs * DRYV_COMPACT_MAX_RECORD always suffices). `threads` <= 1 runs on the calling thread.
 * Returns DRYV_OK, or DRYV_ERR_ARG when the stream does not fit. Needs no GPU. */
int dryv_recon_pack_levels(const int16_t* coeff, size_t n_mbs, uint32_t* offset, uint8_t* stream,
                           size_t stream_cap, int threads);

/* Host-side inverse (diagnostic, used by the CPU tests of the format): compact -> dense int16 [n_mbs][384].
 * Returns DRYV_ERR_ARG on a malformed stream. Needs no GPU; nothing on the reconstruction path calls it. */
int dryv_recon_unpack_levels(const dryv_mb_levels_compact* levels, size_t n_mbs, int16_t* coeff);

/* dryv_recon_submit with the levels in the compact format: `soa->coeff` is ignored (may be NULL), every
 * other field of `soa` is read as in dryv_recon_submit. HOST pointers. */
int dryv_recon_submit_compact(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                              const dryv_mb_levels_compact* levels, uint32_t n_frames, uint8_t* out_yuv);

/* The expansion kernel alone, DEVICE pointers: d_levels->offset / stream are device arrays; writes the dense
 * int16 [n_mbs][384] array at d_coeff (16-byte aligned). Malformed records are reported by dryv_recon_wait
 * as DRYV_ERR_UNSUPPORTED. */
int dryv_recon_expand_levels_device(dryv_recon_ctx* ctx, const dryv_mb_levels_compact* d_levels, size_t n_mbs,
                                    int16_t* d_coeff, void* cuda_stream);

/* ---- Output surface ---------------------------------------------------------------------
 * The reference writes the coded picture, macroblock aligned and planar (frame/mod.rs:48-70); "frame cropping" is an
 * open item of its roadmap (README.md:13) although the crop fields are already parsed (atom/avcc/sps.rs:252-267).
 * A surface is the rectangle of the coded picture a consumer wants and the layout it wants it in:
 *   DRYV_SURFACE_I420: Y (width x height), Cb (width/2 x height/2), Cr — planar, rows packed;
 *   DRYV_SURFACE_NV12: Y (width x height), then height/2 rows of width bytes: Cb, Cr pairs interleaved.
 * crop_left / crop_top / width / height are in luma samples and even (4:2:0); the rectangle must lie inside the coded
 * picture. The SPS crop rectangle of a stream (7.4.2.1.1: frame_crop_*_offset x CropUnit, CropUnitX = CropUnitY = 2
 * for 4:2:0 frame pictures) is returned by dryv_cabac_surface (include/dryv_cabac_host.h). */
#define DRYV_SURFACE_I420 0u
#define DRYV_SURFACE_NV12 1u
typedef struct dryv_surface {
  uint32_t format;
  uint32_t crop_left, crop_top;
  uint32_t width, height;
} dryv_surface;

/* Bytes of one exported picture (width * height * 3 / 2); 0 if `s` is NULL or malformed (odd / zero fields, unknown format). */
size_t dryv_recon_surface_bytes(const dryv_surface* s);

/* Device buffers: d_yuv holds n_frames reconstructed pictures of `pp` geometry (dryv_recon_reconstruct_device's output),
 * d_out receives n_frames surfaces of dryv_recon_surface_bytes each. Asynchronous on `cuda_stream` (0: the context's own),
 * stream-ordered after the reconstruction when that ran on the same stream. One streaming pass, no host staging. */
int dryv_recon_export_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const uint8_t* d_yuv, uint32_t n_frames,
                             const dryv_surface* s, uint8_t* d_out, void* cuda_stream);

/* Selects what dryv_recon_submit / dryv_recon_submit_compact hand back from now on: with a surface set, `out_yuv`
 * receives n_frames * dryv_recon_surface_bytes(s) bytes (the export runs on the GPU behind each stage's kernels, so the
 * D2H copy moves the cropped pictures only); s == NULL restores the coded pictures (the default, the reference's layout).
 * Returns DRYV_ERR_ARG for a malformed surface; that it fits the pictures is checked by the submit that uses it. */
int dryv_recon_set_surface(dryv_recon_ctx* ctx, const dryv_surface* s);

/* ---- Deblocking post-pass — NOT part of dryv parity --------------------------------------
 * The reference has no in-loop filter (README.md:15; it parses disable_deblocking_filter_idc and the two offsets in
 * slice/header.rs:609-640 and ignores them), so its output — and everything above in this header — is the unfiltered
 * reconstruction. For streams that ask for the filter this applies H.264 8.7 to n_frames reconstructed intra pictures IN
 * PLACE (device memory, 16-byte aligned, the dryv_recon_reconstruct_device layout): bS 4 on macroblock edges, 3 on transform edges, every
 * edge filtered (disable_deblocking_filter_idc 0; with one slice per picture 2 is the same). d_soa: the batch's device
 * syntax buffers, of which qp and transform_size_8x8_flag are read. The offsets are the slice header's
 * slice_alpha_c0_offset_div2 / slice_beta_offset_div2 (-6..6). Asynchronous on `cuda_stream` (0: the context's own);
 * errors (watchdog) are reported by dryv_recon_wait. */
int dryv_recon_deblock_device(dryv_recon_ctx* ctx, const dryv_pic_params* pp, const dryv_mb_soa* d_soa, uint32_t n_frames,
                              int slice_alpha_c0_offset_div2, int slice_beta_offset_div2, uint8_t* d_yuv, void* cuda_stream);

/* The same post-pass as a mode of the host submit paths: with enable != 0 every dryv_recon_reconstruct /
 * dryv_recon_reconstruct_compact filters the pictures of a slot behind their reconstruction, before the surface export
 * and the copy back (the dryv_recon_multi_* forms own their contexts and always reconstruct unfiltered). Off by default and
 * off on the dryv-parity path — the reference decodes every stream unfiltered. DRYV_ERR_ARG for offsets outside -6..6. */
int dryv_recon_set_deblock(dryv_recon_ctx* ctx, int enable, int slice_alpha_c0_offset_div2, int slice_beta_offset_div2);

/* Frame::write_to_yuv_file (frame/mod.rs:48-70): writes one reconstructed picture (host memory, the
 * layout above) to `path`, creating the parent directory of "temp/yuv_frame"-style paths if needed. */
int dryv_recon_write_yuv_file(const uint8_t* frame_yuv, size_t bytes, const char* path);

/* Device-measured duration (CUDA events: first H2D enqueue -> last D2H complete) of the most recent
 * dryv_recon_submit, in milliseconds; valid after dryv_recon_wait. Returns < 0 if unavailable. */
double dryv_recon_last_submit_ms(dryv_recon_ctx* ctx);

/* Diagnostic: the host-built lookup tables the kernels consume for `pp` (LevelScale, zig-zag, tap and
 * schedule tables; layout = struct dryv::DeviceTables, dryv_b200/csrc/recon_tables.h). Copies up to `cap`
 * bytes into `out` and returns the table size in bytes. Needs no GPU. */
size_t dryv_recon_device_tables(const dryv_pic_params* pp, void* out, size_t cap);

/* Number of kernel launches issued by this context so far (bench bookkeeping). */
uint64_t dryv_recon_launch_count(dryv_recon_ctx* ctx);

/* Device-measured durations (CUDA events on the launching stream), in milliseconds, of the most recent
 * wavefront-kernel launches of this context, newest first; at most 64 are kept. Returns how many were written
 * (<= cap) or a negative error code. Valid after dryv_recon_wait. Bench bookkeeping: the roofline figure is
 * quoted on this kernel alone. */
int dryv_recon_wavefront_times(dryv_recon_ctx* ctx, float* out_ms, int cap);

/* ---- several GPUs in one process --------------------------------------------------------------
 * Independent IDR pictures never reference each other on this path (the reference starts every picture from zeroed
 * planes: Frame::new, src/video/frame/mod.rs:29-46; one Frame per sample, src/video/decoder.rs:124), so several GPUs are
 * used by dealing pictures to them: one dryv_recon_ctx and one host thread per device, device d of N reconstructs the
 * d-th contiguous block of pictures (the first n_frames % N blocks hold one picture more) and writes it into its slice
 * of the caller's buffer. No data is exchanged between devices: no collective, no NCCL.
 * `devices` lists the CUDA devices to use (a device may be named more than once: each entry gets a context of its own);
 * NULL = devices 0 .. n_devices-1, and every visible device when n_devices <= 0. */
typedef struct dryv_recon_multi dryv_recon_multi;
int dryv_recon_multi_create(const int* devices, int n_devices, dryv_recon_multi** out);
void dryv_recon_multi_destroy(dryv_recon_multi* m);
int dryv_recon_multi_device_count(const dryv_recon_multi* m);
const char* dryv_recon_multi_last_error(dryv_recon_multi* m);
/* dryv_recon_submit + dryv_recon_wait (dryv_recon_submit_compact + dryv_recon_wait) of every device's share, concurrently;
 * HOST pointers, blocking, the result is byte-identical to a single device's. Returns the first device's error code
 * that is not DRYV_OK (dryv_recon_multi_last_error names the device). */
int dryv_recon_multi_reconstruct(dryv_recon_multi* m, const dryv_pic_params* pp, const dryv_mb_soa* soa, uint32_t n_frames,
                                 uint8_t* out_yuv);
int dryv_recon_multi_reconstruct_compact(dryv_recon_multi* m, const dryv_pic_params* pp, const dryv_mb_soa* soa,
                                         const dryv_mb_levels_compact* levels, uint32_t n_frames, uint8_t* out_yuv);

#ifdef __cplusplus
}
#endif
#endif /* DRYV_RECON_H */
