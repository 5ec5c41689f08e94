"""CPU host side of the path (include/dryv_cabac_host.h): Annex-B H.264 bytes -> SyntaxBatch.

The caller of the reconstruction path in the reference is its CABAC parser (src/video/cabac/mod.rs:89-210); this binds the
C++ restatement of it for IDR I-slice pictures (dryv_b200/csrc/cabac_host.cpp). CPU code by design: entropy decoding stays
on the host, the GPU gets the syntax buffers."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import PicParams, Surface, SyntaxBatch
from .abi import COMPACT_MAX_RECORD
from .recon import OK, CompactLevels, PinnedArray, ReconError, load_library


def scan(stream: bytes):
    """-> (PicParams, number of IDR pictures) of an Annex-B stream."""
    lib = load_library()
    buf = np.frombuffer(stream, np.uint8)
    pp, n = PicParams(), C.c_uint32()
    rc = lib.dryv_cabac_scan(buf.ctypes.data, buf.size, C.byref(pp), C.byref(n))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_scan: not a supported H.264 stream")
    return pp, int(n.value)


def picture_params(stream: bytes, picture: int) -> PicParams:
    """dryv_cabac_picture_params: the parameters (geometry, chroma QP offsets, the scaling lists the reference would pick)
    of picture `picture`; pictures of one stream may name different parameter sets."""
    lib = load_library()
    buf = np.frombuffer(stream, np.uint8)
    pp = PicParams()
    rc = lib.dryv_cabac_picture_params(buf.ctypes.data, buf.size, picture, C.byref(pp))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_picture_params")
    return pp


class SliceInfo(C.Structure):
    """include/dryv_cabac_host.h dryv_slice_info"""
    _fields_ = [("pic_parameter_set_id", C.c_uint8), ("seq_parameter_set_id", C.c_uint8), ("slice_qp", C.c_uint8),
                ("disable_deblocking_filter_idc", C.c_uint8), ("slice_alpha_c0_offset_div2", C.c_int8),
                ("slice_beta_offset_div2", C.c_int8), ("scaling_matrix_source", C.c_uint8), ("reserved", C.c_uint8)]


def slice_info(stream: bytes, picture: int) -> SliceInfo:
    """dryv_cabac_slice_info: what picture `picture`'s slice header asks for (deblocking filter, parameter sets)."""
    lib = load_library()
    buf = np.frombuffer(stream, np.uint8)
    si = SliceInfo()
    rc = lib.dryv_cabac_slice_info(buf.ctypes.data, buf.size, picture, C.byref(si))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_slice_info")
    return si


def surface(stream: bytes) -> Surface:
    """The display rectangle the stream's SPS asks for (frame_crop_*_offset) as an I420 Surface; the whole coded picture
    when the SPS does not crop."""
    lib = load_library()
    buf = np.frombuffer(stream, np.uint8)
    sf = Surface()
    rc = lib.dryv_cabac_surface(buf.ctypes.data, buf.size, C.byref(sf))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_surface")
    return sf


def parse(stream: bytes, threads: int = 0, out: SyntaxBatch | None = None, first: int = 0,
          count: int | None = None) -> SyntaxBatch:
    """CABAC-parses the IDR pictures [first, first + count) of the stream (default: all) into syntax buffers (dense
    levels). `stream` is an MP4 file or an Annex-B byte stream."""
    lib = load_library()
    pp, n = scan(stream)
    whole = count is None and first == 0
    count = n - first if count is None else count
    b = out if out is not None else SyntaxBatch.empty(pp, count)
    buf = np.frombuffer(stream, np.uint8)
    rc = lib.dryv_cabac_parse_range(buf.ctypes.data, buf.size, C.byref(pp), first, count, 1 if whole else 0,
                                    b.mb_type.ctypes.data, b.transform_size_8x8_flag.ctypes.data,
                                    b.intra_chroma_pred_mode.ctypes.data, b.qp.ctypes.data, b.pred_syntax.ctypes.data,
                                    b.coeff.ctypes.data, threads or (os.cpu_count() or 1))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_parse_range")
    return b


def parse_compact(stream: bytes, threads: int = 0, first: int = 0, count: int | None = None, pinned: bool = False):
    """CABAC-parses pictures [first, first + count) straight into the compact level stream: -> (SyntaxBatch whose
    `coeff` is empty, CompactLevels), the pair ReconContext.submit_compact takes."""
    lib = load_library()
    pp, n = scan(stream)
    count = n - first if count is None else count
    n_mbs = pp.n_mb * count
    b = SyntaxBatch(pp, count, np.zeros(n_mbs, np.uint8), np.zeros(n_mbs, np.uint8), np.zeros(n_mbs, np.uint8),
                    np.zeros(n_mbs, np.uint8), np.zeros((n_mbs, 16), np.uint8), np.zeros((0, 384), np.int16))
    cap = n_mbs * COMPACT_MAX_RECORD
    scratch = np.empty(cap, np.uint8)
    off = np.empty(n_mbs + 1, np.uint32)
    buf = np.frombuffer(stream, np.uint8)
    rc = lib.dryv_cabac_parse_compact(buf.ctypes.data, buf.size, C.byref(pp), first, count, b.mb_type.ctypes.data,
                                      b.transform_size_8x8_flag.ctypes.data, b.intra_chroma_pred_mode.ctypes.data,
                                      b.qp.ctypes.data, b.pred_syntax.ctypes.data, off.ctypes.data, scratch.ctypes.data, cap,
                                      threads or (os.cpu_count() or 1))
    if rc != OK:
        raise ReconError(rc, "dryv_cabac_parse_compact")
    used = int(off[-1])
    if pinned:
        po, ps = PinnedArray(n_mbs + 1, np.uint32), PinnedArray(max(used, 4), np.uint8)
        po.array[:] = off
        ps.array[:used] = scratch[:used]
        return b, CompactLevels(po.array, ps.array, (po, ps))
    return b, CompactLevels(off, scratch[:max(used, 4)].copy())
