"""ctypes mirror of include/dryv_recon.h (the C ABI data contract).

`PicParams` = dryv_pic_params, `MbSoa` = dryv_mb_soa; `SyntaxBatch` owns the numpy arrays a host would
fill while CABAC-parsing (reference: struct Macroblock, src/video/slice/macroblock.rs:21-129) and hands
out the pointer struct. No computation lives here.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

COEFFS_PER_MB = 384

ZIGZAG4 = [(0, 0), (0, 1), (1, 0), (2, 0), (1, 1), (0, 2), (0, 3), (1, 2),
           (2, 1), (3, 0), (3, 1), (2, 2), (1, 3), (2, 3), (3, 2), (3, 3)]


class PicParams(C.Structure):
    _fields_ = [
        ("pic_width_in_mbs", C.c_uint16),
        ("pic_height_in_mbs", C.c_uint16),
        ("chroma_qp_index_offset", C.c_int8),
        ("second_chroma_qp_index_offset", C.c_int8),
        ("reserved0", C.c_uint8 * 2),
        ("flags", C.c_uint32),
        ("scaling_list4x4", C.c_uint8 * 16),
        ("scaling_list8x8", C.c_uint8 * 64),
    ]

    @classmethod
    def make(cls, w_mbs: int, h_mbs: int, cb_off: int = 0, cr_off: int | None = None,
             list4x4=None, list8x8=None) -> "PicParams":
        pp = cls()
        pp.pic_width_in_mbs = w_mbs
        pp.pic_height_in_mbs = h_mbs
        pp.chroma_qp_index_offset = cb_off
        pp.second_chroma_qp_index_offset = cb_off if cr_off is None else cr_off
        l4 = [16] * 16 if list4x4 is None else list(list4x4)
        l8 = [16] * 64 if list8x8 is None else list(list8x8)
        assert len(l4) == 16 and len(l8) == 64
        pp.scaling_list4x4[:] = l4
        pp.scaling_list8x8[:] = l8
        return pp

    @property
    def n_mb(self) -> int:
        return int(self.pic_width_in_mbs) * int(self.pic_height_in_mbs)

    @property
    def frame_bytes(self) -> int:
        return self.n_mb * 384

    @property
    def luma_pixels(self) -> int:
        return self.n_mb * 256


class MbSoa(C.Structure):
    _fields_ = [
        ("mb_type", C.c_void_p),
        ("transform_size_8x8_flag", C.c_void_p),
        ("intra_chroma_pred_mode", C.c_void_p),
        ("qp", C.c_void_p),
        ("pred_syntax", C.c_void_p),
        ("coeff", C.c_void_p),
    ]


class LevelsCompact(C.Structure):
    """dryv_mb_levels_compact: the compact level stream (include/dryv_recon.h)."""
    _fields_ = [("offset", C.c_void_p), ("stream", C.c_void_p)]


COMPACT_MAX_RECORD = 4 + 24 * 2 + COEFFS_PER_MB * 2

SURFACE_I420, SURFACE_NV12 = 0, 1


class Surface(C.Structure):
    """dryv_surface: the rectangle of the coded picture to hand out and its layout (include/dryv_recon.h)."""
    _fields_ = [("format", C.c_uint32), ("crop_left", C.c_uint32), ("crop_top", C.c_uint32), ("width", C.c_uint32),
                ("height", C.c_uint32)]

    @classmethod
    def make(cls, width, height, crop_left=0, crop_top=0, fmt=SURFACE_I420):
        return cls(fmt, crop_left, crop_top, width, height)

    @property
    def nbytes(self) -> int:
        return self.width * self.height * 3 // 2


FIELDS = ("mb_type", "transform_size_8x8_flag", "intra_chroma_pred_mode", "qp", "pred_syntax", "coeff")


@dataclass
class SyntaxBatch:
    """Host-side SoA syntax buffers for `n_frames` pictures of `pp` geometry."""

    pp: PicParams
    n_frames: int
    mb_type: np.ndarray                  # u8 [n_frames * n_mb]
    transform_size_8x8_flag: np.ndarray  # u8 [n_frames * n_mb]
    intra_chroma_pred_mode: np.ndarray   # u8 [n_frames * n_mb]
    qp: np.ndarray                       # u8 [n_frames * n_mb]
    pred_syntax: np.ndarray              # u8 [n_frames * n_mb, 16]
    coeff: np.ndarray                    # i16 [n_frames * n_mb, 384]

    @classmethod
    def empty(cls, pp: PicParams, n_frames: int) -> "SyntaxBatch":
        n = pp.n_mb * n_frames
        return cls(pp, n_frames,
                   np.zeros(n, np.uint8), np.zeros(n, np.uint8), np.zeros(n, np.uint8),
                   np.zeros(n, np.uint8), np.zeros((n, 16), np.uint8),
                   np.zeros((n, COEFFS_PER_MB), np.int16))

    def arrays(self):
        return [getattr(self, f) for f in FIELDS]

    def as_soa(self) -> MbSoa:
        soa = MbSoa()
        for f in FIELDS:
            a = getattr(self, f)
            assert a.flags["C_CONTIGUOUS"]
            setattr(soa, f, a.ctypes.data)
        return soa

    def frames(self, lo: int, hi: int) -> "SyntaxBatch":
        """View of pictures [lo, hi) (no copy)."""
        n = self.pp.n_mb
        return SyntaxBatch(self.pp, hi - lo, *[a[lo * n:hi * n] for a in self.arrays()])

    def copy(self) -> "SyntaxBatch":
        return SyntaxBatch(self.pp, self.n_frames, *[a.copy() for a in self.arrays()])

    @property
    def input_bytes(self) -> int:
        return sum(a.nbytes for a in self.arrays())
