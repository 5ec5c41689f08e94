"""ctypes binding of libdryv_recon.so — the C ABI declared in include/dryv_recon.h.

The shared library holds the hand-written sm_100a CUDA kernels (dryv_b200/csrc). There is no CPU
fallback: if the library is missing or no B200 is present, construction raises.
PyTorch is used only as the owner of device memory / streams for the device-pointer entry points.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .abi import COMPACT_MAX_RECORD, FIELDS, LevelsCompact, MbSoa, PicParams, Surface, SyntaxBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("DRYV_RECON_LIB") or os.path.join(CSRC, "libdryv_recon.so")  # override: dev variants only

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WATCHDOG = 0, -1, -2, -3, -4

EXPORTS = [
    "dryv_recon_abi_version", "dryv_recon_frame_bytes", "dryv_recon_create", "dryv_recon_destroy",
    "dryv_recon_last_error", "dryv_recon_alloc_pinned", "dryv_recon_free_pinned", "dryv_recon_submit",
    "dryv_recon_wait", "dryv_recon_reconstruct_device", "dryv_recon_residual_add_device",
    "dryv_recon_write_yuv_file", "dryv_recon_launch_count", "dryv_recon_last_submit_ms", "dryv_recon_device_tables",
    "dryv_recon_wavefront_times", "dryv_recon_pack_levels", "dryv_recon_unpack_levels", "dryv_recon_submit_compact",
    "dryv_recon_expand_levels_device", "dryv_recon_wait_oldest", "dryv_recon_surface_bytes", "dryv_recon_export_device",
    "dryv_recon_set_surface", "dryv_recon_deblock_device", "dryv_recon_set_deblock",
    "dryv_recon_multi_create", "dryv_recon_multi_destroy", "dryv_recon_multi_device_count", "dryv_recon_multi_last_error",
    "dryv_recon_multi_reconstruct", "dryv_recon_multi_reconstruct_compact",
]
HOST_EXPORTS = ["dryv_cabac_scan", "dryv_cabac_parse", "dryv_cabac_parse_range", "dryv_cabac_parse_compact",
                "dryv_cabac_surface", "dryv_cabac_picture_params", "dryv_cabac_slice_info"]  # include/dryv_cabac_host.h

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("recon.cu", "recon_tables.cpp", "levels_pack.cpp", "cabac_host.cpp", "multi.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("recon_kernels.cuh", "residual_stage.cuh", "deblock_kernel.cuh", "deblock_packed.cuh", "recon_tables.h", "cabac_tables.inc", "levels_record.h")] + [
        os.path.join(_HERE, "..", "include", "dryv_recon.h"), os.path.join(_HERE, "..", "include", "dryv_cabac_host.h")]
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB_PATH]
    subprocess.check_call(cmd)
    return LIB_PATH


class ReconError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dryv_recon error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """Load libdryv_recon.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ReconError(ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u32, sz = C.c_void_p, C.c_uint32, C.c_size_t
    lib.dryv_recon_abi_version.restype = C.c_int
    lib.dryv_recon_frame_bytes.restype = sz
    lib.dryv_recon_frame_bytes.argtypes = [C.POINTER(PicParams)]
    lib.dryv_recon_create.restype = C.c_int
    lib.dryv_recon_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.dryv_recon_destroy.restype = None
    lib.dryv_recon_destroy.argtypes = [vp]
    lib.dryv_recon_last_error.restype = C.c_char_p
    lib.dryv_recon_last_error.argtypes = [vp]
    lib.dryv_recon_alloc_pinned.restype = C.c_int
    lib.dryv_recon_alloc_pinned.argtypes = [sz, C.POINTER(vp)]
    lib.dryv_recon_free_pinned.restype = None
    lib.dryv_recon_free_pinned.argtypes = [vp]
    lib.dryv_recon_submit.restype = C.c_int
    lib.dryv_recon_submit.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), u32, vp]
    lib.dryv_recon_wait.restype = C.c_int
    lib.dryv_recon_wait.argtypes = [vp]
    lib.dryv_recon_wait_oldest.restype = C.c_int
    lib.dryv_recon_wait_oldest.argtypes = [vp]
    lib.dryv_recon_reconstruct_device.restype = C.c_int
    lib.dryv_recon_reconstruct_device.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), u32, vp, vp]
    lib.dryv_recon_residual_add_device.restype = C.c_int
    lib.dryv_recon_residual_add_device.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), u32, vp, vp, vp]
    lib.dryv_recon_write_yuv_file.restype = C.c_int
    lib.dryv_recon_write_yuv_file.argtypes = [vp, sz, C.c_char_p]
    lib.dryv_recon_last_submit_ms.restype = C.c_double
    lib.dryv_recon_last_submit_ms.argtypes = [vp]
    lib.dryv_recon_device_tables.restype = sz
    lib.dryv_recon_device_tables.argtypes = [C.POINTER(PicParams), vp, sz]
    lib.dryv_recon_launch_count.restype = C.c_uint64
    lib.dryv_recon_launch_count.argtypes = [vp]
    lib.dryv_recon_wavefront_times.restype = C.c_int
    lib.dryv_recon_wavefront_times.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.dryv_cabac_scan.restype = C.c_int
    lib.dryv_cabac_scan.argtypes = [vp, sz, C.POINTER(PicParams), C.POINTER(u32)]
    lib.dryv_cabac_parse.restype = C.c_int
    lib.dryv_cabac_parse.argtypes = [vp, sz, C.POINTER(PicParams), u32, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.dryv_cabac_parse_compact.restype = C.c_int
    lib.dryv_cabac_parse_compact.argtypes = [vp, sz, C.POINTER(PicParams), u32, u32, vp, vp, vp, vp, vp, vp, vp, sz, C.c_int]
    lib.dryv_cabac_parse_range.restype = C.c_int
    lib.dryv_cabac_parse_range.argtypes = [vp, sz, C.POINTER(PicParams), u32, u32, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
    lib.dryv_recon_pack_levels.restype = C.c_int
    lib.dryv_recon_pack_levels.argtypes = [vp, sz, vp, vp, sz, C.c_int]
    lib.dryv_recon_unpack_levels.restype = C.c_int
    lib.dryv_recon_unpack_levels.argtypes = [C.POINTER(LevelsCompact), sz, vp]
    lib.dryv_recon_submit_compact.restype = C.c_int
    lib.dryv_recon_submit_compact.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), C.POINTER(LevelsCompact), u32, vp]
    lib.dryv_recon_expand_levels_device.restype = C.c_int
    lib.dryv_recon_expand_levels_device.argtypes = [vp, C.POINTER(LevelsCompact), sz, vp, vp]
    lib.dryv_recon_deblock_device.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), u32, C.c_int, C.c_int, vp, vp]
    lib.dryv_recon_deblock_device.restype = C.c_int
    lib.dryv_recon_set_deblock.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.dryv_recon_set_deblock.restype = C.c_int
    lib.dryv_recon_surface_bytes.argtypes = [C.POINTER(Surface)]
    lib.dryv_recon_surface_bytes.restype = sz
    lib.dryv_recon_export_device.argtypes = [vp, C.POINTER(PicParams), vp, u32, C.POINTER(Surface), vp, vp]
    lib.dryv_recon_export_device.restype = C.c_int
    lib.dryv_recon_set_surface.argtypes = [vp, C.POINTER(Surface)]
    lib.dryv_recon_set_surface.restype = C.c_int
    lib.dryv_cabac_surface.argtypes = [vp, sz, C.POINTER(Surface)]
    lib.dryv_cabac_surface.restype = C.c_int
    lib.dryv_cabac_picture_params.argtypes = [vp, sz, u32, C.POINTER(PicParams)]
    lib.dryv_cabac_picture_params.restype = C.c_int
    lib.dryv_cabac_slice_info.argtypes = [vp, sz, u32, vp]
    lib.dryv_cabac_slice_info.restype = C.c_int
    lib.dryv_recon_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.dryv_recon_multi_create.restype = C.c_int
    lib.dryv_recon_multi_destroy.argtypes = [vp]
    lib.dryv_recon_multi_destroy.restype = None
    lib.dryv_recon_multi_device_count.argtypes = [vp]
    lib.dryv_recon_multi_device_count.restype = C.c_int
    lib.dryv_recon_multi_last_error.argtypes = [vp]
    lib.dryv_recon_multi_last_error.restype = C.c_char_p
    lib.dryv_recon_multi_reconstruct.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), u32, vp]
    lib.dryv_recon_multi_reconstruct.restype = C.c_int
    lib.dryv_recon_multi_reconstruct_compact.argtypes = [vp, C.POINTER(PicParams), C.POINTER(MbSoa), C.POINTER(LevelsCompact), u32, vp]
    lib.dryv_recon_multi_reconstruct_compact.restype = C.c_int
    _lib = lib
    return lib


class PinnedArray:
    """numpy view over page-locked host memory from dryv_recon_alloc_pinned."""

    def __init__(self, shape, dtype):
        lib = load_library()
        self.shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        self.dtype = np.dtype(dtype)
        nbytes = max(1, int(np.prod(self.shape)) * self.dtype.itemsize)
        p = C.c_void_p()
        rc = lib.dryv_recon_alloc_pinned(nbytes, C.byref(p))
        if rc != OK:
            raise ReconError(rc, "pinned allocation failed")
        self._ptr = p
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            load_library().dryv_recon_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):  # pragma: no cover
        try:
            self.free()
        except Exception:
            pass


def pinned_batch(pp: PicParams, n_frames: int):
    """A SyntaxBatch whose arrays live in pinned memory (+ the PinnedArray owners to keep alive)."""
    n = pp.n_mb * n_frames
    owners = [PinnedArray(n, np.uint8), PinnedArray(n, np.uint8), PinnedArray(n, np.uint8), PinnedArray(n, np.uint8),
              PinnedArray((n, 16), np.uint8), PinnedArray((n, 384), np.int16)]
    b = SyntaxBatch(pp, n_frames, *[o.array for o in owners])
    return b, owners


class CompactLevels:
    """The compact level stream of a batch (dryv_mb_levels_compact): `offset` u32 [n_mbs + 1], `stream` u8."""

    def __init__(self, offset: np.ndarray, stream: np.ndarray, owners=()):
        self.offset, self.stream, self._owners = offset, stream, owners

    @property
    def n_mbs(self) -> int:
        return self.offset.size - 1

    @property
    def nbytes(self) -> int:
        """Bytes that travel: the offsets and the used part of the stream."""
        return int(self.offset.nbytes + int(self.offset[-1]))

    def as_struct(self) -> LevelsCompact:
        lv = LevelsCompact()
        lv.offset = self.offset.ctypes.data
        lv.stream = self.stream.ctypes.data
        return lv

    def unpack(self) -> np.ndarray:
        """Host-side inverse (dryv_recon_unpack_levels, diagnostic): dense int16 [n_mbs, 384]."""
        out = np.empty((self.n_mbs, 384), np.int16)
        lv = self.as_struct()
        rc = load_library().dryv_recon_unpack_levels(C.byref(lv), self.n_mbs, out.ctypes.data)
        if rc != OK:
            raise ReconError(rc, "malformed compact level stream")
        return out


def pack_levels(coeff: np.ndarray, threads: int = 0, pinned: bool = False) -> CompactLevels:
    """dryv_recon_pack_levels: dense int16 [n_mbs, 384] -> compact stream (host-side converter, no GPU needed)."""
    lib = load_library()
    coeff = np.ascontiguousarray(coeff, np.int16).reshape(-1, 384)
    n = coeff.shape[0]
    cap = n * COMPACT_MAX_RECORD
    scratch = np.empty(cap, np.uint8)
    off = np.empty(n + 1, np.uint32)
    rc = lib.dryv_recon_pack_levels(coeff.ctypes.data, n, off.ctypes.data, scratch.ctypes.data, cap,
                                    threads or (os.cpu_count() or 1))
    if rc != OK:
        raise ReconError(rc, "dryv_recon_pack_levels")
    used = int(off[-1])
    if pinned:
        po, ps = PinnedArray(n + 1, np.uint32), PinnedArray(max(used, 4), np.uint8)
        po.array[:] = off
        ps.array[:used] = scratch[:used]
        return CompactLevels(po.array, ps.array, (po, ps))
    return CompactLevels(off, scratch[:max(used, 4)].copy())


class DeviceSoa:
    """SoA syntax buffers resident in HBM (torch owns the memory)."""

    def __init__(self, batch: SyntaxBatch, device="cuda:0"):
        import torch
        self.pp = batch.pp
        self.n_frames = batch.n_frames
        self.tensors = {f: torch.from_numpy(np.ascontiguousarray(getattr(batch, f))).to(device) for f in FIELDS}

    def as_soa(self) -> MbSoa:
        soa = MbSoa()
        for f in FIELDS:
            setattr(soa, f, self.tensors[f].data_ptr())
        return soa

    def frames(self, lo: int, hi: int) -> "DeviceSoa":
        """View of pictures [lo, hi) (no copy)."""
        n = self.pp.n_mb
        v = object.__new__(DeviceSoa)
        v.pp, v.n_frames = self.pp, hi - lo
        v.tensors = {f: t[lo * n:hi * n] for f, t in self.tensors.items()}
        return v

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors.values())


class ReconContext:
    """One dryv_recon_ctx (one CUDA device). Not thread-safe; create one per device."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.dryv_recon_create(device, C.byref(h))
        if rc != OK:
            raise ReconError(rc, f"dryv_recon_create(device={device}) failed: no usable sm_100 GPU "
                                 "(this path has no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.dryv_recon_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise ReconError(rc, (self.lib.dryv_recon_last_error(self.h) or b"").decode())

    # -- output surface (crop rectangle, I420 / NV12): include/dryv_recon.h, SURVEY.md §8(f) next-3 ---
    def set_surface(self, surface: "Surface | None"):
        """What submit / submit_compact hand back from now on: the given surface, or (None) the coded pictures."""
        self._check(self.lib.dryv_recon_set_surface(self.h, C.byref(surface) if surface is not None else None))
        self._surface = surface

    def _out_bytes(self, pp: PicParams) -> int:
        sf = getattr(self, "_surface", None)
        return sf.nbytes if sf is not None else pp.frame_bytes

    def set_deblock(self, enable: bool, alpha_div2: int = 0, beta_div2: int = 0):
        """dryv_recon_set_deblock: the host submit paths (reconstruct, reconstruct_compact) filter every picture behind the
        reconstruction — off by default, the reference has no filter."""
        self._check(self.lib.dryv_recon_set_deblock(self.h, 1 if enable else 0, alpha_div2, beta_div2))

    def deblock_device(self, dsoa: "DeviceSoa", d_yuv, alpha_div2: int = 0, beta_div2: int = 0, stream_ptr: int = 0):
        """dryv_recon_deblock_device: the optional H.264 in-loop filter over reconstructed pictures, in place (not dryv parity)."""
        soa = dsoa.as_soa()
        self._check(self.lib.dryv_recon_deblock_device(self.h, C.byref(dsoa.pp), C.byref(soa), dsoa.n_frames, alpha_div2,
                                                       beta_div2, d_yuv.data_ptr(), stream_ptr or None))

    def export_device(self, pp: PicParams, d_yuv, n_frames: int, surface: "Surface", d_out, stream_ptr: int = 0):
        """dryv_recon_export_device on torch device tensors: coded pictures -> surfaces, asynchronous on the stream."""
        assert d_out.numel() >= n_frames * surface.nbytes and d_yuv.numel() >= n_frames * pp.frame_bytes
        self._check(self.lib.dryv_recon_export_device(self.h, C.byref(pp), d_yuv.data_ptr(), n_frames, C.byref(surface),
                                                      d_out.data_ptr(), stream_ptr or None))

    # -- host-buffer path (what a decoder host calls): H2D + kernels + D2H inside --------------------
    def submit(self, batch: SyntaxBatch, out: np.ndarray):
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size >= batch.n_frames * self._out_bytes(batch.pp)
        soa = batch.as_soa()
        self._keep = (batch, soa, out)
        self.__dict__.setdefault("_keep_all", []).append(self._keep)
        self._check(self.lib.dryv_recon_submit(self.h, C.byref(batch.pp), C.byref(soa), batch.n_frames,
                                               out.ctypes.data))

    def submit_compact(self, batch: SyntaxBatch, levels: CompactLevels, out: np.ndarray):
        """dryv_recon_submit_compact: `batch.coeff` is not read, the levels come from the compact stream."""
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size >= batch.n_frames * self._out_bytes(batch.pp)
        assert levels.n_mbs == batch.n_frames * batch.pp.n_mb
        soa = batch.as_soa()
        soa.coeff = None
        lv = levels.as_struct()
        self._keep = (batch, soa, levels, lv, out)
        self.__dict__.setdefault("_keep_all", []).append(self._keep)
        self._check(self.lib.dryv_recon_submit_compact(self.h, C.byref(batch.pp), C.byref(soa), C.byref(lv),
                                                       batch.n_frames, out.ctypes.data))

    def reconstruct_compact(self, batch: SyntaxBatch, levels: CompactLevels, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((batch.n_frames, batch.pp.frame_bytes), np.uint8)
        self.submit_compact(batch, levels, out)
        self.wait()
        return out

    def expand_levels_device(self, d_offset, d_stream, n_mbs: int, d_coeff, stream_ptr: int = 0):
        """The expansion kernel alone on torch device tensors (tests)."""
        lv = LevelsCompact()
        lv.offset, lv.stream = d_offset.data_ptr(), d_stream.data_ptr()
        self._check(self.lib.dryv_recon_expand_levels_device(self.h, C.byref(lv), n_mbs, d_coeff.data_ptr(),
                                                             stream_ptr or None))

    def wait(self):
        self._check(self.lib.dryv_recon_wait(self.h))
        self._keep_all = []

    def wait_oldest(self):
        """Streaming use: returns when the oldest outstanding submit's pictures are complete."""
        self._check(self.lib.dryv_recon_wait_oldest(self.h))

    def reconstruct(self, batch: SyntaxBatch, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((batch.n_frames, batch.pp.frame_bytes), np.uint8)
        self.submit(batch, out)
        self.wait()
        return out

    # -- device-pointer path ------------------------------------------------------------------------
    def reconstruct_device(self, dsoa: DeviceSoa, d_out, stream_ptr: int = 0):
        soa = dsoa.as_soa()
        self._check(self.lib.dryv_recon_reconstruct_device(self.h, C.byref(dsoa.pp), C.byref(soa), dsoa.n_frames,
                                                           d_out.data_ptr(), stream_ptr or None))

    def residual_add_device(self, dsoa: DeviceSoa, d_pred, d_out, stream_ptr: int = 0):
        soa = dsoa.as_soa()
        self._check(self.lib.dryv_recon_residual_add_device(self.h, C.byref(dsoa.pp), C.byref(soa), dsoa.n_frames,
                                                            d_pred.data_ptr(), d_out.data_ptr(), stream_ptr or None))

    @property
    def last_submit_ms(self) -> float:
        return float(self.lib.dryv_recon_last_submit_ms(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.dryv_recon_launch_count(self.h))

    def wavefront_times_ms(self, n: int = 64) -> list:
        """CUDA-event durations of the most recent wavefront-kernel launches, newest first (call after wait())."""
        buf = (C.c_float * n)()
        got = self.lib.dryv_recon_wavefront_times(self.h, buf, n)
        if got < 0:
            raise ReconError(got, "dryv_recon_wavefront_times")
        return [float(buf[i]) for i in range(got)]


class MultiDeviceContext:
    """dryv_recon_multi: one dryv_recon_ctx and one host thread per device inside this process; pictures are dealt in
    contiguous blocks (shard.frames_for_rank) and land in disjoint slices of one output array. `devices`: list of CUDA
    device indices (None: every visible device); an index may repeat (each entry gets its own context)."""

    def __init__(self, devices=None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.dryv_recon_multi_create(None, 0, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.dryv_recon_multi_create(arr, len(devices), C.byref(h))
        if rc != OK:
            raise ReconError(rc, "dryv_recon_multi_create failed (no usable sm_100 GPU / bad device list; no CPU fallback)")
        self.h = h

    @property
    def n_devices(self) -> int:
        return int(self.lib.dryv_recon_multi_device_count(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.dryv_recon_multi_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != OK:
            raise ReconError(rc, (self.lib.dryv_recon_multi_last_error(self.h) or b"").decode())

    def reconstruct(self, batch: SyntaxBatch, out: np.ndarray | None = None, levels: "CompactLevels | None" = None) -> np.ndarray:
        if out is None:
            out = np.empty((batch.n_frames, batch.pp.frame_bytes), np.uint8)
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.size >= batch.n_frames * batch.pp.frame_bytes
        soa = batch.as_soa()
        if levels is None:
            self._check(self.lib.dryv_recon_multi_reconstruct(self.h, C.byref(batch.pp), C.byref(soa), batch.n_frames,
                                                              out.ctypes.data))
        else:
            soa.coeff = None
            lv = levels.as_struct()
            self._check(self.lib.dryv_recon_multi_reconstruct_compact(self.h, C.byref(batch.pp), C.byref(soa), C.byref(lv),
                                                                      batch.n_frames, out.ctypes.data))
        return out


def write_yuv_file(frame: np.ndarray, path: str):
    """Frame::write_to_yuv_file (reference src/video/frame/mod.rs:48-70) for one reconstructed picture."""
    lib = load_library()
    frame = np.ascontiguousarray(frame, np.uint8)
    rc = lib.dryv_recon_write_yuv_file(frame.ctypes.data, frame.nbytes, path.encode())
    if rc != OK:
        raise OSError(f"cannot write {path}")
