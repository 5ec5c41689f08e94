"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, agrees
with the ctypes mirror on struct layout, builds its host-side tables correctly and fails loudly (never
falls back) without a GPU. No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dryv_b200 import recon
from dryv_b200.abi import MbSoa, PicParams
from oracle import spec_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "dryv_recon.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dryv_recon_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(recon_lib):
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(recon_lib, n), f"{n} declared in include/dryv_recon.h but not exported"
    assert sorted(recon.EXPORTS) == names


def test_struct_layout_matches_header():
    assert C.sizeof(PicParams) == 2 + 2 + 1 + 1 + 2 + 4 + 16 + 64
    assert PicParams.scaling_list4x4.offset == 12 and PicParams.scaling_list8x8.offset == 28
    assert C.sizeof(MbSoa) == 6 * C.sizeof(C.c_void_p)


def test_abi_version_and_frame_bytes(recon_lib):
    assert recon_lib.dryv_recon_abi_version() == 1
    pp = PicParams.make(120, 68)
    assert recon_lib.dryv_recon_frame_bytes(C.byref(pp)) == 1920 * 1088 * 3 // 2 == pp.frame_bytes


def test_no_gpu_means_error_not_fallback(recon_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert recon_lib.dryv_recon_create(0, C.byref(h)) == recon.ERR_CUDA
    assert not h.value
    with pytest.raises(recon.ReconError):
        recon.ReconContext(0)


def test_null_arguments_are_rejected(recon_lib):
    assert recon_lib.dryv_recon_create(0, None) == recon.ERR_ARG
    assert recon_lib.dryv_recon_wait(None) == recon.ERR_ARG
    assert recon_lib.dryv_recon_submit(None, None, None, 1, None) == recon.ERR_ARG
    assert recon_lib.dryv_recon_write_yuv_file(None, 0, b"x") == recon.ERR_ARG


def test_write_yuv_file_creates_temp_dir(recon_lib, tmp_path):
    frame = np.arange(384, dtype=np.uint8)
    path = tmp_path / "temp" / "yuv_frame"
    recon.write_yuv_file(frame, str(path))
    assert np.array_equal(np.fromfile(path, np.uint8), frame)


class I4Step(C.Structure):
    _fields_ = [("org", C.c_uint32), ("res2", C.c_uint32)]


class Tables(C.Structure):
    # mirrors dryv_b200/csrc/recon_tables.h: residual part | prediction part | t4 (global memory only)
    _fields_ = [("ls8", (C.c_uint16 * 64) * 6), ("t4b", (C.c_uint32 * 8) * 52), ("ls00", C.c_int32 * 8),
                ("t4b_e", C.c_uint8 * 52), ("qpc", C.c_uint8 * 52), ("zz8inv", (C.c_uint8 * 8) * 8),
                ("setbits4", C.c_uint16 * 16), ("pad0", C.c_uint8 * 8),
                ("i4row", (C.c_uint16 * 16) * 16), ("tap4", ((C.c_uint32 * 4) * 16) * 22),
                ("tap8", ((C.c_uint8 * 8) * 32) * 9), ("i4tab", (I4Step * 2) * 11), ("i4sched", (C.c_uint8 * 2) * 10),
                ("pad1", C.c_uint8 * 12), ("t4", (C.c_int32 * 16) * 52)]


TILE_STRIDE = 48
TAP4_BIAS = 256


def get_tables(lib, pp):
    t = Tables()
    n = lib.dryv_recon_device_tables(C.byref(pp), C.byref(t), C.sizeof(t))
    assert n == C.sizeof(t)
    return t


@pytest.mark.parametrize("custom", [False, True])
def test_level_scale_tables(recon_lib, custom):
    l4 = list(range(6, 38, 2)) if custom else [16] * 16
    l8 = [8 + (k * 3) % 40 for k in range(64)] if custom else [16] * 64
    t = get_tables(recon_lib, PicParams.make(2, 2, 0, 0, l4, l8))
    for qp in range(52):
        ls = spec_model.level_scale4(l4, qp % 6)
        for k, (i, j) in enumerate(spec_model.ZZ4):
            assert t.t4[qp][k] == int(ls[i, j]) << max(qp // 6 - 4, 0)
    for m in range(6):
        assert np.array_equal(np.array(t.ls8[m]).reshape(8, 8), spec_model.level_scale8(l8, m))
    assert [int(v) for v in t.qpc] == spec_model.QPC
    assert [int(v) for v in t.ls00[:6]] == [int(spec_model.level_scale4(l4, m)[0, 0]) for m in range(6)]
    # byte-scale form (residual_stage.cuh): where it exists, level * factor * 2^e is the standard's dequantisation exactly
    for qp in range(52):
        e = t.t4b_e[qp]
        if custom and e == 0xff:
            continue
        assert e != 0xff, "flat lists have the form at every qP"
        shr = max(4 - qp // 6, 0)
        for k in range(16):
            f = (t.t4b[qp][k >> 1] >> (24 * (k & 1))) & 0xff
            assert (f << e) << shr == t.t4[qp][k] and (e == 0 or f % 4 == 0)
    for m in range(16):
        bits = [b for b in range(4) if m & (1 << b)]
        assert [(t.setbits4[m] >> (4 * i)) & 15 for i in range(4)] == bits + [15] * (4 - len(bits))


def test_zigzag_and_tap_tables(recon_lib):
    t = get_tables(recon_lib, PicParams.make(1, 1))
    for k, (i, j) in enumerate(spec_model.ZZ8):
        assert t.zz8inv[i][j] == k
    # the reference's own table starts 0,1 / 1,0 / 2,0 / 1,1 (frame/mod.rs:215-219) and ends ... 7,6 / 7,7
    assert spec_model.ZZ8[:5] == [(0, 0), (0, 1), (1, 0), (2, 0), (1, 1)] and spec_model.ZZ8[-2:] == [(7, 6), (7, 7)]

    # every tap triple reproduces the closed-form predictor on random edges (modes other than DC, which the
    # kernels compute from the block's top word and left column directly)
    rng = np.random.default_rng(1)
    for n in (4, 8):
        T = rng.integers(0, 256, 2 * n).tolist()
        L = rng.integers(0, 256, n).tolist()
        TL = int(rng.integers(0, 256))
        E = T + L + [TL]
        for mode in (0, 1, 3, 4, 5, 6, 7, 8):
            want = spec_model.pred_nxn(n, mode, T, L, TL)
            if n == 8:
                for lane in range(32):
                    y, x = lane >> 2, (lane & 3) * 2
                    for q in range(2):
                        i0, i1, i2 = (t.tap8[mode][lane][q * 3 + k] for k in range(3))
                        assert (E[i0] + 2 * E[i1] + E[i2] + 2) >> 2 == want[y, x + q], (n, mode, x + q, y)
                continue
            # 4x4: the table holds tile byte offsets; place the edge samples in a scratch tile and gather
            for variant in range(2):
                Tv = T[:4] + [T[3]] * 4 if variant else T
                wantv = spec_model.pred_nxn(4, mode, Tv, L, TL)
                for p in range(16):
                    org = 1024  # origin of the block inside a big scratch tile
                    tile = {}
                    for i in range(8):
                        tile[org - TILE_STRIDE + i] = T[i]
                    for k in range(4):
                        tile[org + TILE_STRIDE * k - 1] = L[k]
                    tile[org - TILE_STRIDE - 1] = TL
                    e = [tile[org - TAP4_BIAS + t.tap4[9 * variant + mode][p][k]] for k in range(3)]
                    assert t.tap4[9 * variant + mode][p][3] == 1
                    assert (e[0] + 2 * e[1] + e[2] + 2) >> 2 == wantv[p >> 2, p & 3], (variant, mode, p)


def test_intra4x4_schedule_respects_decode_order(recon_lib):
    """A block may only be scheduled after every neighbour it can read (left, top, top-left, and top-right
    unless the reference treats it as unavailable) — the availability rules of pred4x4.rs:39-43 — and the
    block of half-warp B always sits 8 px right / 4 px up of half-warp A's."""
    t = get_tables(recon_lib, PicParams.make(1, 1))
    pos = lambda b: (((b >> 2) & 1) * 2 + (b & 1), (b >> 3) * 2 + ((b >> 1) & 1))  # noqa: E731
    step_of, blk_at = {}, {}
    for s in range(10):
        a, b = t.i4sched[s][0], t.i4sched[s][1]
        step_of[pos(a)] = s
        blk_at[pos(a)] = a
        if b != 0xff:
            step_of[pos(b)] = s
            blk_at[pos(b)] = b
            assert (pos(b)[0] - pos(a)[0], pos(b)[1] - pos(a)[1]) == (2, -1)
    assert len(step_of) == 16
    # the step table and the per-availability row info agree with the schedule and the availability rules
    for s in range(10):
        for h in range(2):
            e, b = t.i4tab[s][h], t.i4sched[s][h]
            if b == 0xff:   # no block for this half-warp in this step: a dummy block in the tile's padding columns
                assert e.org == (4 + 1) * TILE_STRIDE + 16 + 20 and e.res2 == 0
                continue
            gx, gy = pos(b)
            assert e.org == (4 * gy + 1) * TILE_STRIDE + 16 + 4 * gx and e.res2 == 2 * (4 * gy * 20 + 4 * gx)
            for av in range(16):
                A, B, Cc, D = bool(av & 1), bool(av & 2), bool(av & 4), bool(av & 8)
                aT, aL = gy > 0 or B, gx > 0 or A
                aTL = True if (gx > 0 and gy > 0) else (B if gx > 0 else (A if gy > 0 else D))
                aTR = False if b in (3, 7, 11, 13, 15) else (Cc if b == 5 else (B if b in (0, 1, 4) else True))
                legal = 0x004 | (0x089 if aT else 0) | (0x102 if aL else 0) | (0x070 if (aT and aL and aTL) else 0)
                w = t.i4row[av][8 + s if h else s]
                assert (w & 0x1ff) == legal and (w >> 9) == (0 if aTR else 1)
    for h in range(2):   # look-ahead entry of the last step
        assert (t.i4tab[10][h].org, t.i4tab[10][h].res2) == (t.i4tab[9][h].org, t.i4tab[9][h].res2)
    # row kinds: 0 illegal, 1 three-tap gather, 2..5 the DC flavours
    kinds = [t.tap4[r][0][3] for r in range(22)]
    assert kinds == [1, 1, 2, 1, 1, 1, 1, 1, 1] * 2 + [0, 3, 4, 5]
    for (gx, gy), s in step_of.items():
        tr = blk_at[(gx, gy)] not in (3, 7, 11, 13, 15)
        for dx, dy, needed in ((-1, 0, True), (0, -1, True), (-1, -1, True), (1, -1, tr)):
            nx, ny = gx + dx, gy + dy
            if needed and 0 <= nx < 4 and 0 <= ny < 4:
                assert step_of[(nx, ny)] < s
