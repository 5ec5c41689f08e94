"""ctypes wrapper over oracle/libdryv_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dryv_b200.abi import MbSoa, PicParams, SyntaxBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdryv_oracle.so")

__all__ = ["build", "reconstruct", "residual_add", "block4x4", "block8x8", "write_yuv_file"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "dryv_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdryv_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.dryv_oracle_reconstruct_mt.restype = C.c_int
        _lib.dryv_oracle_reconstruct_mt.argtypes = [C.POINTER(PicParams), C.POINTER(MbSoa), C.c_uint32, C.c_void_p,
                                                    C.c_uint32]
        _lib.dryv_oracle_residual_add.restype = C.c_int
        _lib.dryv_oracle_residual_add.argtypes = [C.POINTER(PicParams), C.POINTER(MbSoa), C.c_uint32, C.c_void_p,
                                                  C.c_void_p]
        _lib.dryv_oracle_block4x4.restype = C.c_int
        _lib.dryv_oracle_block4x4.argtypes = [C.POINTER(PicParams), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib.dryv_oracle_block8x8.restype = C.c_int
        _lib.dryv_oracle_block8x8.argtypes = [C.POINTER(PicParams), C.c_int, C.c_void_p, C.c_void_p]
        _lib.dryv_oracle_write_yuv_file.restype = C.c_int
        _lib.dryv_oracle_write_yuv_file.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
    return _lib


def reconstruct(batch: SyntaxBatch, threads: int = 1) -> np.ndarray:
    """Frame::new + Frame::decode per MB for every picture; returns u8 [n_frames, frame_bytes]."""
    lib = _load()
    out = np.empty((batch.n_frames, batch.pp.frame_bytes), np.uint8)
    soa = batch.as_soa()
    rc = lib.dryv_oracle_reconstruct_mt(C.byref(batch.pp), C.byref(soa), batch.n_frames, out.ctypes.data,
                                        max(1, threads))
    if rc != 0:
        raise ValueError(f"oracle rejected the input: {rc}")
    return out


def residual_add(batch: SyntaxBatch, pred: np.ndarray) -> np.ndarray:
    lib = _load()
    pred = np.ascontiguousarray(pred, np.uint8).reshape(batch.n_frames, batch.pp.frame_bytes)
    out = np.empty_like(pred)
    soa = batch.as_soa()
    rc = lib.dryv_oracle_residual_add(C.byref(batch.pp), C.byref(soa), batch.n_frames, pred.ctypes.data,
                                      out.ctypes.data)
    if rc != 0:
        raise ValueError(f"oracle rejected the input: {rc}")
    return out


def block4x4(pp: PicParams, qp: int, mode: int, coeff_zz) -> np.ndarray:
    """mode: 0 luma (I4x4 MB), 1 luma of an I16x16 MB (DC passthrough), 2 Cb, 3 Cr. Returns r[4,4]."""
    lib = _load()
    c = np.ascontiguousarray(coeff_zz, np.int16)
    r = np.zeros(16, np.int32)
    rc = lib.dryv_oracle_block4x4(C.byref(pp), qp, mode, c.ctypes.data, r.ctypes.data)
    assert rc == 0
    return r.reshape(4, 4)


def block8x8(pp: PicParams, qp: int, coeff_zz) -> np.ndarray:
    lib = _load()
    c = np.ascontiguousarray(coeff_zz, np.int16)
    r = np.zeros(64, np.int32)
    rc = lib.dryv_oracle_block8x8(C.byref(pp), qp, c.ctypes.data, r.ctypes.data)
    assert rc == 0
    return r.reshape(8, 8)


def write_yuv_file(frame: np.ndarray, path: str) -> None:
    lib = _load()
    frame = np.ascontiguousarray(frame, np.uint8)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    rc = lib.dryv_oracle_write_yuv_file(frame.ctypes.data, frame.nbytes, path.encode())
    if rc != 0:
        raise OSError(f"cannot write {path}")
