"""Host-side mirror of the reference's reconstruction interface (src/video/frame/mod.rs):

    Frame::new(&slice)                 -> Frame.new(slice)                 frame/mod.rs:29-46
    frame.decode(&mut slice)           -> frame.decode(slice)              frame/mod.rs:72-90 (once per MB)
    frame.write_to_yuv_file(path)      -> frame.write_to_yuv_file(path)    frame/mod.rs:48-70

The reference reconstructs inside decode(); here decode() only appends the current macroblock's parsed
syntax (the fields CABAC left in slice.mb(), struct Macroblock, slice/macroblock.rs:21-129) to the
structure-of-arrays buffers, and the picture is reconstructed on the GPU in one submit when the planes are
first needed (write_to_yuv_file / planes()). Same names, argument meaning and error behaviour:
I_PCM and inter macroblocks raise NotImplementedError where the reference hits todo!() (frame/mod.rs:85-88).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .abi import PicParams, SyntaxBatch

MB_TYPE_I_NXN = 0
MB_TYPE_I_PCM = 25


@dataclass
class Macroblock:
    """Recon-relevant subset of struct Macroblock (slice/macroblock.rs:21-129), reference field names."""
    mb_type: int = 0
    transform_size_8x8_flag: int = 0
    qp1y: int = 0
    intra_chroma_pred_mode: int = 0
    prev_intra4x4_pred_mode_flag: list = field(default_factory=lambda: [0] * 16)
    rem_intra4x4_pred_mode: list = field(default_factory=lambda: [0] * 16)
    prev_intra8x8_pred_mode_flag: list = field(default_factory=lambda: [0] * 4)
    rem_intra8x8_pred_mode: list = field(default_factory=lambda: [0] * 4)
    block_luma_dc: np.ndarray = field(default_factory=lambda: np.zeros(16, np.int64))
    block_luma_ac: np.ndarray = field(default_factory=lambda: np.zeros((16, 15), np.int64))
    block_luma_4x4: np.ndarray = field(default_factory=lambda: np.zeros((16, 16), np.int64))
    block_luma_8x8: np.ndarray = field(default_factory=lambda: np.zeros((4, 64), np.int64))
    block_chroma_dc: np.ndarray = field(default_factory=lambda: np.zeros((2, 8), np.int64))
    block_chroma_ac: np.ndarray = field(default_factory=lambda: np.zeros((2, 8, 15), np.int64))


@dataclass
class Slice:
    """What Frame reads from struct Slice / SliceHeader / PPS (slice/mod.rs:111-174, slice/header.rs:317-332)."""
    pic_width_in_mbs: int
    pic_height_in_mbs: int
    chroma_qp_index_offset: int = 0
    second_chroma_qp_index_offset: int | None = None
    scaling_list4x4: list | None = None  # list 0, zig-zag order; None = flat 16
    scaling_list8x8: list | None = None
    curr_mb_addr: int = 0
    macroblock: Macroblock = field(default_factory=Macroblock)

    def mb(self) -> Macroblock:
        return self.macroblock


def pack_macroblock(mb: Macroblock, batch: SyntaxBatch, idx: int) -> None:
    """One macroblock -> the SoA record `idx` (layout: include/dryv_recon.h, dryv_mb_soa)."""
    code = int(mb.mb_type)
    if code == MB_TYPE_I_PCM:
        raise NotImplementedError("Sample construction process for I PCM macroblocks")  # frame/mod.rs:85-86
    if code > MB_TYPE_I_PCM:
        raise NotImplementedError("Inter prediction")  # frame/mod.rs:87-88
    batch.mb_type[idx] = code
    batch.transform_size_8x8_flag[idx] = 1 if mb.transform_size_8x8_flag else 0
    batch.intra_chroma_pred_mode[idx] = mb.intra_chroma_pred_mode
    batch.qp[idx] = mb.qp1y
    ps = batch.pred_syntax[idx]
    ps[:] = 0
    cf = batch.coeff[idx]
    if code == MB_TYPE_I_NXN and not mb.transform_size_8x8_flag:
        for k in range(16):
            ps[k] = ((1 if mb.prev_intra4x4_pred_mode_flag[k] else 0) << 3) | (int(mb.rem_intra4x4_pred_mode[k]) & 7)
        cf[:256] = np.asarray(mb.block_luma_4x4).reshape(256)
    elif code == MB_TYPE_I_NXN:
        for k in range(4):
            ps[k] = ((1 if mb.prev_intra8x8_pred_mode_flag[k] else 0) << 3) | (int(mb.rem_intra8x8_pred_mode[k]) & 7)
        cf[:256] = np.asarray(mb.block_luma_8x8).reshape(256)
    else:
        luma = cf[:256].reshape(16, 16)
        luma[:, 0] = np.asarray(mb.block_luma_dc)[:16]
        luma[:, 1:] = np.asarray(mb.block_luma_ac)[:16, :15]
    for pl in range(2):
        c = cf[256 + pl * 64:256 + (pl + 1) * 64].reshape(4, 16)
        c[:, 0] = np.asarray(mb.block_chroma_dc)[pl, :4]
        c[:, 1:] = np.asarray(mb.block_chroma_ac)[pl, :4, :15]


class Frame:
    """One picture being reconstructed. `ctx` is a dryv_b200.recon.ReconContext (GPU); there is no CPU path."""

    def __init__(self, pp: PicParams, ctx=None):
        self.pp = pp
        self.ctx = ctx
        self.width_l = pp.pic_width_in_mbs * 16
        self.height_l = pp.pic_height_in_mbs * 16
        self.width_c = self.width_l // 2
        self.height_c = self.height_l // 2
        self.batch = SyntaxBatch.empty(pp, 1)
        self.decoded = np.zeros(pp.n_mb, bool)
        self._yuv = None

    @classmethod
    def new(cls, slice: Slice, ctx=None) -> "Frame":
        pp = PicParams.make(slice.pic_width_in_mbs, slice.pic_height_in_mbs, slice.chroma_qp_index_offset,
                            slice.second_chroma_qp_index_offset, slice.scaling_list4x4, slice.scaling_list8x8)
        return cls(pp, ctx)

    def decode(self, slice: Slice) -> None:
        addr = int(slice.curr_mb_addr)
        if not 0 <= addr < self.pp.n_mb:
            raise IndexError(f"curr_mb_addr {addr} outside the picture")
        pack_macroblock(slice.mb(), self.batch, addr)
        self.decoded[addr] = True
        self._yuv = None

    def reconstruct(self) -> np.ndarray:
        """Runs the GPU path for this picture (all macroblocks must have been decode()d)."""
        if self._yuv is None:
            if not self.decoded.all():
                raise RuntimeError("picture incomplete: decode() every macroblock before reading the planes")
            if self.ctx is None:
                from .recon import ReconContext
                self.ctx = ReconContext(0)
            self._yuv = self.ctx.reconstruct(self.batch)[0]
        return self._yuv

    def planes(self):
        yuv = self.reconstruct()
        nl = self.width_l * self.height_l
        nc = self.width_c * self.height_c
        return (yuv[:nl].reshape(self.height_l, self.width_l),
                yuv[nl:nl + nc].reshape(self.height_c, self.width_c),
                yuv[nl + nc:].reshape(self.height_c, self.width_c))

    def write_to_yuv_file(self, file_path: str) -> None:
        from .recon import write_yuv_file
        write_yuv_file(self.reconstruct(), file_path)
