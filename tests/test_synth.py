"""CPU tests of the seeded syntax-buffer generator (workload generator)."""
import numpy as np

import oracle
from dryv_b200 import synth
from dryv_b200.abi import PicParams


def test_deterministic_and_seed_sensitive():
    pp = PicParams.make(5, 4)
    a = synth.generate(pp, 3, 100, threads=1)
    b = synth.generate(pp, 3, 100, threads=3)
    c = synth.generate(pp, 3, 101)
    for x, y in zip(a.arrays(), b.arrays()):
        assert np.array_equal(x, y)
    assert not np.array_equal(a.coeff, c.coeff)
    # picture f of a batch seeded s equals picture 0 of a batch seeded s + f
    d = synth.generate(pp, 1, 102)
    e = synth.generate(pp, 3, 100)
    assert np.array_equal(e.coeff[2 * pp.n_mb:], d.coeff)


def test_syntax_ranges_and_mix():
    pp = PicParams.make(30, 20)
    b = synth.generate(pp, 2, 7)
    assert b.mb_type.max() <= 24 and b.intra_chroma_pred_mode.max() <= 3 and b.qp.max() <= 51
    nxn = b.mb_type == 0
    frac4 = (nxn & (b.transform_size_8x8_flag == 0)).mean()
    frac8 = (nxn & (b.transform_size_8x8_flag == 1)).mean()
    assert 0.33 < frac4 < 0.47 and 0.19 < frac8 < 0.31
    assert (b.transform_size_8x8_flag[~nxn] == 0).all()
    # Intra16x16 mb_type is consistent with the levels present (cbp bits folded into the code, cabac/mod.rs:170-174)
    i16 = np.nonzero(~nxn)[0]
    luma = b.coeff[i16, :256].reshape(-1, 16, 16)
    has_ac = (luma[:, :, 1:] != 0).any(axis=(1, 2))
    assert np.array_equal(b.mb_type[i16] >= 13, has_ac)


def test_first_row_and_column_modes_are_legal():
    # top-left MB can only use DC-type prediction; the oracle reconstructs it without touching missing neighbours
    pp = PicParams.make(4, 4)
    b = synth.generate(pp, 4, 9, zero_residual=True)
    n = pp.n_mb
    for f in range(4):
        assert b.intra_chroma_pred_mode[f * n] == 0
        if b.mb_type[f * n] != 0:
            assert (b.mb_type[f * n] - 1) % 4 == 2
    out = oracle.reconstruct(b)
    # with zero residual and legal modes the first MB is flat 128
    y = out[0, :n * 256].reshape(64, 64)
    assert (y[:16, :16] == 128).all()


def test_forward_transform_round_trip_at_low_qp():
    # the generator quantises a bounded spatial residual; at QP 4 reconstruction error must be tiny:
    # residual_add on a mid-grey picture gives back grey + residual within the quantiser step
    pp = PicParams.make(6, 4)
    lo = synth.generate(pp, 1, 5, qp_base=4, qp_jitter=0, stress_pct=0)
    hi = synth.generate(pp, 1, 5, qp_base=40, qp_jitter=0, stress_pct=0)
    pred = np.full((1, pp.frame_bytes), 128, np.uint8)
    r_lo = oracle.residual_add(lo, pred).astype(np.int32) - 128
    r_hi = oracle.residual_add(hi, pred).astype(np.int32) - 128
    # same seed -> same spatial residual; the coarse quantiser must be a blurred version of the fine one
    assert np.abs(r_lo).mean() > 2.0                      # Laplace(b = 6) residual survives
    assert np.abs(r_lo - r_hi).mean() > 1.0               # QP 40 throws detail away
    assert np.abs(r_lo).max() <= 64 + 4                   # clip(+-64) plus at most the QP-4 step
    assert (lo.coeff != 0).mean() > 5 * (hi.coeff != 0).mean()


def test_qp_sweep():
    pp = PicParams.make(4, 3)
    b = synth.generate(pp, 6, 50, qp_base=10, qp_jitter=0, qp_step_per_frame=7)
    q = b.qp.reshape(6, -1)
    assert [int(q[f, 0]) for f in range(6)] == [10, 17, 24, 31, 38, 45]
