//! recon_ffi.rs — `extern "C"` declarations of include/dryv_recon.h plus the structure-of-arrays
//! collector that replaces `frame.decode(slice)` (src/video/cabac/mod.rs:208 of the reference).
//!
//! UNVERIFIED TEXT: no Rust toolchain exists in the build image, so this module has not been compiled.
//! The C ABI it binds is exercised from Python (ctypes) and C by this repository's tests; field order,
//! widths and return codes below are transcribed from include/dryv_recon.h.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const DRYV_OK: c_int = 0;
pub const DRYV_ERR_ARG: c_int = -1;
pub const DRYV_ERR_UNSUPPORTED: c_int = -2;
pub const DRYV_ERR_CUDA: c_int = -3;
pub const DRYV_ERR_WATCHDOG: c_int = -4;
pub const DRYV_COEFFS_PER_MB: usize = 384;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dryv_pic_params {
  pub pic_width_in_mbs: u16,
  pub pic_height_in_mbs: u16,
  pub chroma_qp_index_offset: i8,
  pub second_chroma_qp_index_offset: i8,
  pub reserved0: [u8; 2],
  pub flags: u32,
  pub scaling_list4x4: [u8; 16],
  pub scaling_list8x8: [u8; 64],
}

#[repr(C)]
pub struct dryv_mb_soa {
  pub mb_type: *const u8,
  pub transform_size_8x8_flag: *const u8,
  pub intra_chroma_pred_mode: *const u8,
  pub qp: *const u8,
  pub pred_syntax: *const u8,
  pub coeff: *const i16,
}

/// include/dryv_recon.h `dryv_mb_levels_compact`: the compact level stream (the wire format of the host-buffer path).
#[repr(C)]
pub struct dryv_mb_levels_compact {
  pub offset: *const u32, // [n_mbs + 1] byte offsets into `stream`, multiples of 4
  pub stream: *const u8,
}
pub const DRYV_COMPACT_MAX_RECORD: usize = 4 + 24 * 2 + DRYV_COEFFS_PER_MB * 2;

#[repr(C)]
pub struct dryv_recon_ctx {
  _private: [u8; 0],
}

extern "C" {
  pub fn dryv_recon_abi_version() -> c_int;
  pub fn dryv_recon_frame_bytes(pp: *const dryv_pic_params) -> usize;
  pub fn dryv_recon_create(device: c_int, out: *mut *mut dryv_recon_ctx) -> c_int;
  pub fn dryv_recon_destroy(ctx: *mut dryv_recon_ctx);
  pub fn dryv_recon_last_error(ctx: *mut dryv_recon_ctx) -> *const c_char;
  pub fn dryv_recon_alloc_pinned(bytes: usize, out: *mut *mut c_void) -> c_int;
  pub fn dryv_recon_free_pinned(p: *mut c_void);
  pub fn dryv_recon_submit(
    ctx: *mut dryv_recon_ctx,
    pp: *const dryv_pic_params,
    soa: *const dryv_mb_soa,
    n_frames: u32,
    out_yuv: *mut u8,
  ) -> c_int;
  pub fn dryv_recon_submit_compact(
    ctx: *mut dryv_recon_ctx,
    pp: *const dryv_pic_params,
    soa: *const dryv_mb_soa, // `coeff` may be null
    levels: *const dryv_mb_levels_compact,
    n_frames: u32,
    out_yuv: *mut u8,
  ) -> c_int;
  pub fn dryv_recon_pack_levels(
    coeff: *const i16,
    n_mbs: usize,
    offset: *mut u32,
    stream: *mut u8,
    stream_cap: usize,
    threads: c_int,
  ) -> c_int;
  pub fn dryv_recon_wait(ctx: *mut dryv_recon_ctx) -> c_int;
  /// streaming use: up to four submits may be outstanding; returns when the oldest one's pictures are complete
  pub fn dryv_recon_wait_oldest(ctx: *mut dryv_recon_ctx) -> c_int;
  pub fn dryv_recon_write_yuv_file(frame_yuv: *const u8, bytes: usize, path: *const c_char) -> c_int;
  /// output surface (crop rectangle, I420 / NV12): with one set, the submit calls hand back
  /// `n_frames * dryv_recon_surface_bytes(s)` bytes; null restores the coded pictures (what the reference writes)
  pub fn dryv_recon_surface_bytes(s: *const dryv_surface) -> usize;
  pub fn dryv_recon_set_surface(ctx: *mut dryv_recon_ctx, s: *const dryv_surface) -> c_int;
}

/// `dryv_surface` of include/dryv_recon.h. For the SPS display rectangle fill it from the fields the reference already
/// parses (src/video/atom/avcc/sps.rs:252-267): crop_left = 2 * frame_crop_left_offset, crop_top = 2 * frame_crop_top_offset,
/// width = 16 * pic_width_in_mbs - 2 * (left + right), height = 16 * pic_height_in_mbs - 2 * (top + bottom).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct dryv_surface {
  pub format: u32, // 0 = I420 planes, 1 = NV12
  pub crop_left: u32,
  pub crop_top: u32,
  pub width: u32,
  pub height: u32,
}

/// Pinned structure-of-arrays buffers for `n_frames` pictures of `n_mb` macroblocks each.
/// One instance replaces the `Frame` the decoder creates per slice NAL (src/video/decoder.rs:124).
pub struct SoaBatch {
  pub pp: dryv_pic_params,
  pub n_mb: usize,
  pub n_frames: usize,
  base: *mut u8, // one pinned allocation: coeff | pred_syntax | mb_type | t8x8 | chroma_mode | qp
  pub out: *mut u8, // pinned: n_frames pictures, Y | Cb | Cr each
}

impl SoaBatch {
  pub fn new(pp: dryv_pic_params, n_frames: usize) -> Option<Self> {
    let n_mb = pp.pic_width_in_mbs as usize * pp.pic_height_in_mbs as usize;
    let total = n_mb * n_frames;
    let (mut base, mut out) = (std::ptr::null_mut::<c_void>(), std::ptr::null_mut::<c_void>());
    unsafe {
      if dryv_recon_alloc_pinned(total * (768 + 16 + 4), &mut base) != DRYV_OK {
        return None;
      }
      if dryv_recon_alloc_pinned(total * 384, &mut out) != DRYV_OK {
        dryv_recon_free_pinned(base);
        return None;
      }
      std::ptr::write_bytes(base as *mut u8, 0, total * (768 + 16 + 4));
    }
    Some(Self { pp, n_mb, n_frames, base: base as *mut u8, out: out as *mut u8 })
  }
  fn total(&self) -> usize {
    self.n_mb * self.n_frames
  }
  fn coeff(&self) -> *mut i16 {
    self.base as *mut i16
  }
  fn pred_syntax(&self) -> *mut u8 {
    unsafe { self.base.add(self.total() * 768) }
  }
  fn bytes(&self, k: usize) -> *mut u8 {
    unsafe { self.base.add(self.total() * (768 + 16 + k)) }
  }
  pub fn soa(&self) -> dryv_mb_soa {
    dryv_mb_soa {
      mb_type: self.bytes(0),
      transform_size_8x8_flag: self.bytes(1),
      intra_chroma_pred_mode: self.bytes(2),
      qp: self.bytes(3),
      pred_syntax: self.pred_syntax(),
      coeff: self.coeff(),
    }
  }

  /// Replaces `frame.decode(slice)` at src/video/cabac/mod.rs:208: copies the macroblock CABAC has just
  /// parsed into the SoA slot (frame `f`, address `slice.curr_mb_addr`). Field sources:
  /// struct Macroblock, src/video/slice/macroblock.rs:21-129.
  ///
  /// `mb_code` = `*mb.mb_type` (0 = I_NxN, 1..24 = I_16x16_*; 25 = I_PCM and inter codes make
  /// dryv_recon_wait return DRYV_ERR_UNSUPPORTED where the reference hits todo!() at frame/mod.rs:85-88).
  #[allow(clippy::too_many_arguments)]
  pub fn push_macroblock(
    &mut self,
    f: usize,
    mb_addr: usize,
    mb_code: u8,
    transform_size_8x8_flag: u8,
    intra_chroma_pred_mode: u8,
    qp1y: isize,
    prev_flag4: &[u8; 16],
    rem4: &[u8; 16],
    prev_flag8: &[u8; 4],
    rem8: &[u8; 4],
    block_luma_4x4: &[[isize; 16]; 16],
    block_luma_8x8: &[[isize; 64]; 4],
    block_luma_dc: &[isize; 16],
    block_luma_ac: &[[isize; 15]; 16],
    block_chroma_dc: &[[isize; 8]; 2],
    block_chroma_ac: &[[[isize; 15]; 8]; 2],
  ) {
    let i = f * self.n_mb + mb_addr;
    unsafe {
      *self.bytes(0).add(i) = mb_code;
      *self.bytes(1).add(i) = transform_size_8x8_flag;
      *self.bytes(2).add(i) = intra_chroma_pred_mode;
      *self.bytes(3).add(i) = qp1y as u8;
      let syn = self.pred_syntax().add(i * 16);
      let c = self.coeff().add(i * DRYV_COEFFS_PER_MB);
      if mb_code == 0 && transform_size_8x8_flag == 0 {
        for b in 0..16 {
          *syn.add(b) = (prev_flag4[b] << 3) | (rem4[b] & 7);
          for k in 0..16 {
            *c.add(b * 16 + k) = block_luma_4x4[b][k] as i16;
          }
        }
      } else if mb_code == 0 {
        for b in 0..4 {
          *syn.add(b) = (prev_flag8[b] << 3) | (rem8[b] & 7);
          for k in 0..64 {
            *c.add(b * 64 + k) = block_luma_8x8[b][k] as i16;
          }
        }
      } else {
        for b in 0..16 {
          *c.add(b * 16) = block_luma_dc[b] as i16;
          for k in 0..15 {
            *c.add(b * 16 + 1 + k) = block_luma_ac[b][k] as i16;
          }
        }
      }
      for pl in 0..2 {
        for b in 0..4 {
          let dst = c.add(256 + pl * 64 + b * 16);
          *dst = block_chroma_dc[pl][b] as i16;
          for k in 0..15 {
            *dst.add(1 + k) = block_chroma_ac[pl][b][k] as i16;
          }
        }
      }
    }
  }

  /// Reconstructs every picture of the batch on `ctx`'s GPU and returns picture `f` as the byte slice
  /// `Frame::write_to_yuv_file` would have written (src/video/frame/mod.rs:48-70).
  pub fn reconstruct(&mut self, ctx: *mut dryv_recon_ctx) -> Result<(), c_int> {
    let soa = self.soa();
    unsafe {
      let rc = dryv_recon_submit(ctx, &self.pp, &soa, self.n_frames as u32, self.out);
      if rc != DRYV_OK {
        return Err(rc);
      }
      let rc = dryv_recon_wait(ctx);
      if rc != DRYV_OK {
        return Err(rc);
      }
    }
    Ok(())
  }
  pub fn picture(&self, f: usize) -> &[u8] {
    let n = self.n_mb * 384;
    unsafe { std::slice::from_raw_parts(self.out.add(f * n), n) }
  }
}

impl Drop for SoaBatch {
  fn drop(&mut self) {
    unsafe {
      dryv_recon_free_pinned(self.base as *mut c_void);
      dryv_recon_free_pinned(self.out as *mut c_void);
    }
  }
}

/// Appends compact level records (include/dryv_recon.h, `dryv_mb_levels_compact`) in macroblock order, the way
/// `residual_cabac` (src/video/cabac/mod.rs:563-675) produces them: per block a significance map and the non-zero
/// levels. Call `begin_macroblock`, then `push_level(slot, k, level)` for every non-zero level in ascending
/// (slot, k) order — slot = index of the 16-coefficient group inside the macroblock's `coeff` layout (0..23, see
/// dryv_mb_soa), k = coefficient index inside it — then `end_macroblock`. Pictures of a batch are appended one
/// after the other; `levels()` is what `dryv_recon_submit_compact` takes. (For full PCIe speed keep `stream` and
/// `offset` in memory from dryv_recon_alloc_pinned; plain Vecs are shown for brevity.)
pub struct CompactLevels {
  pub offset: Vec<u32>,
  pub stream: Vec<u8>,
  masks: [u16; 24],
  vals: Vec<i16>,
}

impl CompactLevels {
  pub fn new() -> Self {
    Self { offset: vec![0], stream: Vec::new(), masks: [0; 24], vals: Vec::with_capacity(DRYV_COEFFS_PER_MB) }
  }
  pub fn begin_macroblock(&mut self) {
    self.masks = [0; 24];
    self.vals.clear();
  }
  pub fn push_level(&mut self, slot: usize, k: usize, level: isize) {
    if level != 0 {
      self.masks[slot] |= 1 << k;
      self.vals.push(level as i16);
    }
  }
  /// Writes the record with level coding 0 (int8) or 2 (int16); coding 1 (4-bit codes + int16 escapes, see
  /// include/dryv_recon.h) is what dryv_recon_pack_levels picks when it is shorter and is left out here for brevity.
  pub fn end_macroblock(&mut self) {
    let wide = self.vals.iter().any(|&v| v < -128 || v > 127);
    let mut hdr: u32 = if wide { 2 << 30 } else { 0 };
    for b in 0..24 {
      if self.masks[b] != 0 {
        hdr |= 1 << b;
      }
    }
    self.stream.extend_from_slice(&hdr.to_le_bytes());
    for b in 0..24 {
      if self.masks[b] != 0 {
        self.stream.extend_from_slice(&self.masks[b].to_le_bytes());
      }
    }
    for &v in &self.vals {
      if wide {
        self.stream.extend_from_slice(&v.to_le_bytes());
      } else {
        self.stream.push(v as i8 as u8);
      }
    }
    while self.stream.len() % 4 != 0 {
      self.stream.push(0);
    }
    self.offset.push(self.stream.len() as u32);
  }
  pub fn levels(&self) -> dryv_mb_levels_compact {
    dryv_mb_levels_compact { offset: self.offset.as_ptr(), stream: self.stream.as_ptr() }
  }
}
