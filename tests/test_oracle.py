"""CPU tests of the oracle (test infrastructure): hand-computed known answers, agreement with the
independent spec model, and the committed golden fixtures. No GPU."""
import numpy as np
import pytest

import oracle
from oracle import spec_model
from dryv_b200 import synth
from dryv_b200.abi import PicParams, SyntaxBatch
from helpers import golden_cases, load_golden


def flat():
    return PicParams.make(1, 1)


# ---- hand-computed known answers (derived on paper from transform.rs:143-187 / pred8x8.rs:71-145) ----
def test_kat_dc_only_4x4_qp24():
    # c00 = 4, qP = 24: d00 = (4 * 16*10) << 0 = 640; a lone DC spreads unchanged through both butterflies,
    # r = (640 + 32) >> 6 = 10 everywhere
    c = np.zeros(16, np.int16)
    c[0] = 4
    assert (oracle.block4x4(flat(), 24, 0, c) == 10).all()


def test_kat_dc_only_4x4_qp_below_24_rounds():
    # qP = 5: LevelScale(5 % 6, 0, 0) = 16 * 18 = 288; d00 = (3 * 288 + 2^3) >> 4 = 54; r = (54 + 32) >> 6 = 1
    c = np.zeros(16, np.int16)
    c[0] = 3
    assert (oracle.block4x4(flat(), 5, 0, c) == 1).all()


def test_kat_intra16x16_and_chroma_dc_passthrough():
    # Intra16x16 luma and chroma blocks take c00 as an already-scaled DC (transform.rs:145-146): r = (c00 + 32) >> 6
    c = np.zeros(16, np.int16)
    c[0] = 640
    for mode in (1, 2, 3):
        assert (oracle.block4x4(flat(), 30, mode, c) == 10).all()


def test_kat_single_ac_coefficient_4x4():
    # zig-zag index 1 -> c[0][1], qP = 24: d01 = 1 * 16*13 = 208.
    # row 0: e2 = (208 >> 1) = 104, e3 = 208 -> f0 = 208, f1 = 104, f2 = -104, f3 = -208; columns copy row 0
    # r = (f + 32) >> 6 -> 3, 2, -2, -3 (arithmetic shift: (-104 + 32) >> 6 = -2, (-208 + 32) >> 6 = -3)
    c = np.zeros(16, np.int16)
    c[1] = 1
    r = oracle.block4x4(flat(), 24, 0, c)
    assert (r == np.array([3, 2, -2, -3])[None, :]).all()


def test_kat_dc_only_8x8_qp36():
    # qP = 36: d00 = (2 * 16*20) << 0 = 640 -> r = (640 + 32) >> 6 = 10 everywhere
    c = np.zeros(64, np.int16)
    c[0] = 2
    assert (oracle.block8x8(flat(), 36, c) == 10).all()


def test_kat_dc_only_8x8_qp_below_36_rounds():
    # qP = 7: LevelScale8x8(1, 0, 0) = 16 * 22 = 352; d00 = (5 * 352 + 2^4) >> 5 = 55; r = (55 + 32) >> 6 = 1
    c = np.zeros(64, np.int16)
    c[0] = 5
    assert (oracle.block8x8(flat(), 7, c) == 1).all()


def one_mb_batch(mb_type=0, t8=0, cm=0, qp=24):
    b = SyntaxBatch.empty(PicParams.make(1, 1), 1)
    b.mb_type[0], b.transform_size_8x8_flag[0], b.intra_chroma_pred_mode[0], b.qp[0] = mb_type, t8, cm, qp
    return b


def test_kat_picture_without_neighbours_is_128():
    # no neighbours, DC everywhere, zero residual: every sample is 1 << (bitDepth - 1)
    for mb_type, t8 in ((3, 0), (0, 0)):  # I16x16 DC (code 1 + 2), I4x4 with prev flags -> DC
        b = one_mb_batch(mb_type, t8)
        b.pred_syntax[0, :] = 8
        assert (oracle.reconstruct(b) == 128).all()


def test_kat_intra8x8_first_column_quirk_q2():
    # Same picture as Intra8x8: NOT flat in dryv. Block 2 (MB column 0) has a top row (block 0 = 128) but no
    # corner; the reference filter loop overwrites p'[0,-1] with (-1 + 2*128 + 128 + 2) >> 2 = 96
    # (pred8x8.rs:245-247), so DC = (96 + 7*128 + 4) >> 3 = 124. Block 3 then sees left 124 / top 128, all
    # filtered samples unchanged: DC = (8*128 + 8*124 + 8) >> 4 = 126.
    b = one_mb_batch(0, 1)
    b.pred_syntax[0, :] = 8
    y = oracle.reconstruct(b)[0][:256].reshape(16, 16)
    assert (y[:8] == 128).all() and (y[8:, :8] == 124).all() and (y[8:, 8:] == 126).all()


def test_kat_intra4x4_dc_propagates():
    # block 0 gets DC coefficient 4 at qP 24 -> 128 + 10 = 138; every later block is DC-predicted from
    # neighbours that are all 138, with zero residual
    b = one_mb_batch(0, 0)
    b.pred_syntax[0, :] = 8
    b.coeff[0, 0] = 4
    out = oracle.reconstruct(b)[0]
    assert (out[:256] == 138).all() and (out[256:] == 128).all()


def test_kat_horizontal_copy_and_chroma_dc_from_left():
    # two MBs: the second is Intra16x16 horizontal (code 2), chroma DC: everything copies 128 + r of MB 0
    pp = PicParams.make(2, 1)
    b = SyntaxBatch.empty(pp, 1)
    b.mb_type[:] = (3, 2)
    b.qp[:] = 24
    b.coeff[0, 256] = 8  # Cb DC of block 0 in MB 0: f = (8, 8, 8, 8), dcC = ((8*160) << 4) >> 5 = 640 -> +10
    out = oracle.reconstruct(b)[0]
    y = out[:512].reshape(16, 32)
    cb = out[512:640].reshape(8, 16)
    assert (y == 128).all()
    assert (cb[:, :8] == 138).all()
    # MB 1 chroma DC: blocks 0/2 see left = 138 only -> 138; block 1 (x>0, y=0) has no top, left p[-1,3] > 0 -> 138
    assert (cb[:, 8:] == 138).all()


def test_unsupported_mb_type_is_rejected():
    b = one_mb_batch(25)
    with pytest.raises(ValueError):
        oracle.reconstruct(b)


# ---- independent restatement of the standard (+ the reference's deviations) ---------------------------
@pytest.mark.parametrize("w,h,seed,kw", [
    (6, 4, 1, {}), (5, 5, 2, dict(qp_base=10)), (7, 3, 3, dict(qp_base=40, stress_pct=40)),
    (4, 4, 4, dict(pct_i4x4=100, pct_i8x8=0)), (4, 4, 5, dict(pct_i4x4=0, pct_i8x8=100)),
    (4, 4, 6, dict(pct_i4x4=0, pct_i8x8=0)), (1, 1, 7, {}), (1, 5, 8, {}), (6, 1, 9, {}), (3, 3, 10, dict(qp_base=51)),
    (3, 3, 11, dict(qp_base=0, qp_jitter=0)),
])
@pytest.mark.parametrize("custom", [False, True])
def test_oracle_matches_spec_model(w, h, seed, kw, custom):
    l4 = list(range(6, 38, 2)) if custom else None
    l8 = [8 + (k * 3) % 40 for k in range(64)] if custom else None
    pp = PicParams.make(w, h, 3 if custom else 0, -4 if custom else 0, l4, l8)
    b = synth.generate(pp, 2, seed, **kw)
    assert np.array_equal(oracle.reconstruct(b), spec_model.reconstruct(b))


def test_reference_deviations_fire_in_stress_data():
    # the quirks must actually be exercised by the generator, otherwise parity on them is vacuous
    pp = PicParams.make(8, 6)
    b = synth.generate(pp, 3, 77, stress_pct=50)
    with_q = spec_model.reconstruct(b, quirks=True)
    without_q = spec_model.reconstruct(b, quirks=False)
    assert np.array_equal(oracle.reconstruct(b), with_q)
    assert (with_q != without_q).sum() > 100


def test_q2_intra8x8_in_first_column():
    # an Intra8x8 MB in MB column 0 below another MB: block 0 has top but no corner -> filtered p'[0,-1] uses -1
    pp = PicParams.make(1, 2)
    b = synth.generate(pp, 4, 5, pct_i4x4=0, pct_i8x8=100, stress_pct=0)
    assert np.array_equal(oracle.reconstruct(b), spec_model.reconstruct(b, quirks=True))
    assert not np.array_equal(spec_model.reconstruct(b, quirks=True), spec_model.reconstruct(b, quirks=False))


# ---- golden fixtures --------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases())
def test_oracle_reproduces_golden(name):
    b, expected = load_golden(name)
    assert np.array_equal(oracle.reconstruct(b), expected)


def test_multithreaded_oracle_is_identical():
    pp = PicParams.make(10, 6)
    b = synth.generate(pp, 6, 42)
    assert np.array_equal(oracle.reconstruct(b, threads=1), oracle.reconstruct(b, threads=4))


def test_residual_add_consistency():
    # zero levels: residual add is the identity on the prediction picture
    pp = PicParams.make(3, 2)
    b = synth.generate(pp, 1, 3, zero_residual=True)
    pred = np.random.default_rng(0).integers(0, 256, (1, pp.frame_bytes), dtype=np.uint8)
    assert np.array_equal(oracle.residual_add(b, pred), pred)


def test_write_yuv_file_layout(tmp_path):
    pp = PicParams.make(2, 1)
    b = synth.generate(pp, 1, 9)
    out = oracle.reconstruct(b)[0]
    path = tmp_path / "temp" / "yuv_frame"
    oracle.write_yuv_file(out, str(path))
    data = np.fromfile(path, np.uint8)
    assert data.size == 32 * 16 * 3 // 2 and np.array_equal(data, out)
