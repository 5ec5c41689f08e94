"""Seeded spec-legal syntax-buffer generator (ctypes wrapper over synth.c).

Workload generator for tests/, bench.py and smoke(): there is no H.264 encoder or sample media in the
build image, so the buffers dryv's CABAC stage would emit are synthesised (see synth.c header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from ..abi import PicParams, SyntaxBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdryv_synth.so")


class SynthCfg(C.Structure):
    _fields_ = [
        ("qp_base", C.c_int32),
        ("qp_jitter", C.c_int32),
        ("pct_i4x4", C.c_int32),
        ("pct_i8x8", C.c_int32),
        ("stress_pct", C.c_int32),
        ("zero_residual", C.c_int32),
        ("qp_step_per_frame", C.c_int32),
        ("standard_only", C.c_int32),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", _LIB_PATH, src,
                               "-lpthread", "-lm"])
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.dryv_synth_batch.restype = C.c_int
        _lib.dryv_synth_batch.argtypes = [C.POINTER(PicParams), C.POINTER(SynthCfg), C.c_uint64, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_uint32]
    return _lib


def generate(pp: PicParams, n_frames: int, seed0: int, *, qp_base: int = 26, qp_jitter: int = 2,
             pct_i4x4: int = 40, pct_i8x8: int = 25, stress_pct: int = 10, zero_residual: bool = False,
             qp_step_per_frame: int = 0, standard_only: bool = False, threads: int | None = None,
             out: SyntaxBatch | None = None) -> SyntaxBatch:
    """Pictures f = 0..n_frames-1 are seeded seed0 + f (SplitMix64), MB mix and QP per SURVEY.md §8(d)."""
    lib = _load()
    cfg = SynthCfg(qp_base, qp_jitter, pct_i4x4, pct_i8x8, stress_pct, int(zero_residual), qp_step_per_frame,
                   int(standard_only))
    b = out if out is not None else SyntaxBatch.empty(pp, n_frames)
    if threads is None:
        threads = min(os.cpu_count() or 1, 64)
    rc = lib.dryv_synth_batch(C.byref(pp), C.byref(cfg), seed0, n_frames,
                              b.mb_type.ctypes.data, b.transform_size_8x8_flag.ctypes.data,
                              b.intra_chroma_pred_mode.ctypes.data, b.qp.ctypes.data,
                              b.pred_syntax.ctypes.data, b.coeff.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f"dryv_synth_batch failed: {rc}")
    return b
