"""Parity tests proper (B200 only): the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, against the committed golden fixtures, and — at BASELINE sizes — through
size-independent properties. The bar is bit-exact."""
import ctypes as C

import numpy as np
import pytest

import oracle
from dryv_b200 import recon, synth
from dryv_b200.abi import PicParams, SyntaxBatch
from helpers import first_difference, golden_cases, load_golden

pytestmark = pytest.mark.gpu


def device_reconstruct(ctx, batch):
    import torch
    ds = recon.DeviceSoa(batch)
    d_out = torch.zeros((batch.n_frames, batch.pp.frame_bytes), dtype=torch.uint8, device="cuda")
    ctx.reconstruct_device(ds, d_out)
    ctx.wait()
    return d_out.cpu().numpy()


def check(ctx, batch, host=True, device=True):
    ref = oracle.reconstruct(batch, threads=4)
    if host:
        got = ctx.reconstruct(batch)
        assert np.array_equal(ref, got), "host path: " + first_difference(batch.pp, ref, got)
    if device:
        got = device_reconstruct(ctx, batch)
        assert np.array_equal(ref, got), "device path: " + first_difference(batch.pp, ref, got)


# ---- oracle snapshots (tests/golden/*.npz are written by make_golden.py FROM THE ORACLE: they hold the CUDA path to the
# oracle's past output without executing it, they are not independent evidence; the libavcodec-derived fixtures are in
# tests/golden/avc/, the dryv pinning kit in tests/golden/pin/) ------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases())
def test_golden(gpu_ctx, name):
    b, expected = load_golden(name)
    got = gpu_ctx.reconstruct(b)
    assert np.array_equal(expected, got), first_difference(b.pp, expected, got)


# ---- oracle parity on seeded inputs -----------------------------------------------------------------------
@pytest.mark.parametrize("p4,p8", [(100, 0), (0, 100), (0, 0), (40, 25)])
@pytest.mark.parametrize("qp", [0, 4, 23, 24, 26, 35, 36, 45, 51])
def test_classes_and_qp(gpu_ctx, p4, p8, qp):
    pp = PicParams.make(9, 6, cb_off=2, cr_off=-3)
    check(gpu_ctx, synth.generate(pp, 2, 1000 + qp, qp_base=qp, pct_i4x4=p4, pct_i8x8=p8), host=False)


@pytest.mark.parametrize("w,h", [(1, 1), (1, 2), (2, 1), (1, 9), (9, 1), (2, 2), (3, 7), (17, 5), (40, 23)])
def test_ragged_geometries(gpu_ctx, w, h):
    check(gpu_ctx, synth.generate(PicParams.make(w, h), 3, 50 + w * 31 + h))


@pytest.mark.parametrize("cb,cr", [(-12, 12), (12, -12), (5, 5), (0, 7)])
def test_chroma_qp_offsets(gpu_ctx, cb, cr):
    pp = PicParams.make(6, 5, cb, cr)
    for qp in (2, 28, 49):
        check(gpu_ctx, synth.generate(pp, 1, 70 + qp, qp_base=qp), host=False)


def test_non_flat_scaling_lists_q1(gpu_ctx):
    # chroma must be dequantised with the luma list (reference quirk Q1)
    l4 = list(range(6, 38, 2))
    l8 = [8 + (k * 3) % 40 for k in range(64)]
    pp = PicParams.make(8, 5, 1, -2, l4, l8)
    for qp in (12, 30, 44):
        check(gpu_ctx, synth.generate(pp, 2, 300 + qp, qp_base=qp))
    # switching lists on the same context rebuilds the device tables
    check(gpu_ctx, synth.generate(PicParams.make(8, 5), 1, 301))


def test_stress_clamping_and_zero_borders(gpu_ctx):
    # 60 % wide-residual MBs: pixels clamp to 0 / 255 and the "> 0" tests of quirk Q3 fire
    pp = PicParams.make(24, 14)
    b = synth.generate(pp, 4, 777, stress_pct=60)
    ref = oracle.reconstruct(b, threads=4)
    assert (ref == 0).mean() > 0.02 and (ref == 255).mean() > 0.01
    got = gpu_ctx.reconstruct(b)
    assert np.array_equal(ref, got), first_difference(pp, ref, got)


def test_intra8x8_first_column_q2(gpu_ctx):
    check(gpu_ctx, synth.generate(PicParams.make(1, 6), 4, 5, pct_i4x4=0, pct_i8x8=100))


def test_zero_residual_prediction_only(gpu_ctx):
    check(gpu_ctx, synth.generate(PicParams.make(15, 9), 3, 8, zero_residual=True))


def test_extreme_levels_within_int32(gpu_ctx):
    # random (non-conformant but bounded) levels: int32 arithmetic still equals the oracle's int64
    pp = PicParams.make(6, 4)
    b = synth.generate(pp, 2, 4242)
    rng = np.random.default_rng(0)
    b.coeff[:] = rng.integers(-2047, 2048, b.coeff.shape, dtype=np.int16)
    check(gpu_ctx, b)


def test_illegal_modes_predict_zero_q4(gpu_ctx):
    # modes whose neighbours are missing: the reference writes no prediction (stays 0); no error is raised
    pp = PicParams.make(4, 3)
    b = synth.generate(pp, 2, 99)
    rng = np.random.default_rng(1)
    b.pred_syntax[:] = rng.integers(0, 16, b.pred_syntax.shape, dtype=np.uint8)
    b.intra_chroma_pred_mode[:] = rng.integers(0, 4, b.intra_chroma_pred_mode.shape, dtype=np.uint8)
    i16 = b.mb_type != 0
    b.mb_type[i16] = rng.integers(1, 25, int(i16.sum()), dtype=np.uint8)
    check(gpu_ctx, b)


def test_residual_add_kernel(gpu_ctx):
    import torch
    pp = PicParams.make(13, 7, 1, -1)
    for qp, p4, p8 in ((3, 40, 30), (26, 100, 0), (26, 0, 100), (47, 0, 0)):
        b = synth.generate(pp, 2, 600 + qp, qp_base=qp, pct_i4x4=p4, pct_i8x8=p8)
        pred = np.random.default_rng(qp).integers(0, 256, (2, pp.frame_bytes), dtype=np.uint8)
        ref = oracle.residual_add(b, pred)
        ds = recon.DeviceSoa(b)
        d_pred = torch.from_numpy(pred).cuda()
        d_out = torch.zeros_like(d_pred)
        gpu_ctx.residual_add_device(ds, d_pred, d_out)
        gpu_ctx.wait()
        got = d_out.cpu().numpy()
        assert np.array_equal(ref, got), first_difference(pp, ref, got)


# ---- error behaviour ---------------------------------------------------------------------------------------
def test_unsupported_syntax_returns_error_code(gpu_ctx):
    pp = PicParams.make(3, 2)
    for field, value in (("mb_type", 25), ("mb_type", 40), ("intra_chroma_pred_mode", 4), ("qp", 52)):
        b = synth.generate(pp, 1, 5)
        getattr(b, field)[3] = value
        with pytest.raises(recon.ReconError) as e:
            gpu_ctx.reconstruct(b)
        assert e.value.code == recon.ERR_UNSUPPORTED
    # the context stays usable afterwards
    check(gpu_ctx, synth.generate(pp, 1, 6), device=False)


def test_bad_arguments(gpu_ctx, recon_lib):
    pp = PicParams.make(0, 4)
    b = SyntaxBatch.empty(PicParams.make(1, 1), 1)
    soa = b.as_soa()
    out = np.zeros(384, np.uint8)
    assert recon_lib.dryv_recon_submit(gpu_ctx.h, C.byref(pp), C.byref(soa), 1, out.ctypes.data) == recon.ERR_ARG
    ok = PicParams.make(1, 1)
    assert recon_lib.dryv_recon_submit(gpu_ctx.h, C.byref(ok), C.byref(soa), 0, out.ctypes.data) == recon.ERR_ARG
    assert recon_lib.dryv_recon_submit(gpu_ctx.h, C.byref(ok), C.byref(soa), 1, None) == recon.ERR_ARG


# ---- BASELINE sizes: oracle on a few pictures + size-independent properties on the whole batch -------------
def test_1080p_batch(gpu_ctx):
    pp = PicParams.make(120, 68)
    b = synth.generate(pp, 24, 3000)
    got = device_reconstruct(gpu_ctx, b)
    ref = oracle.reconstruct(b.frames(0, 6), threads=6)
    assert np.array_equal(ref, got[:6]), first_difference(pp, ref, got[:6])
    # determinism: a second run over the same buffers is byte-identical
    assert np.array_equal(got, device_reconstruct(gpu_ctx, b))
    # independence: every picture equals its reconstruction in a batch of its own (no cross-picture state)
    for f in (7, 15, 23):
        assert np.array_equal(got[f], device_reconstruct(gpu_ctx, b.frames(f, f + 1))[0])
    # host path (pinned-less numpy buffers, chunked H2D / kernel / D2H pipeline) gives the same bytes
    assert np.array_equal(got, gpu_ctx.reconstruct(b))


def test_2160p_pictures(gpu_ctx):
    pp = PicParams.make(240, 135)
    b = synth.generate(pp, 3, 4000)
    got = device_reconstruct(gpu_ctx, b)
    ref = oracle.reconstruct(b, threads=3)
    assert np.array_equal(ref, got), first_difference(pp, ref, got)


def test_qp_sweep_streams(gpu_ctx):
    # BASELINE configs[4] in miniature: QP 10..45 sweep, dense -> sparse levels
    pp = PicParams.make(30, 17)
    b = synth.generate(pp, 36, 5000, qp_base=10, qp_jitter=0, qp_step_per_frame=1)
    got = device_reconstruct(gpu_ctx, b)
    ref = oracle.reconstruct(b, threads=8)
    assert np.array_equal(ref, got), first_difference(pp, ref, got)


def test_pinned_buffers_and_repeated_submits(gpu_ctx):
    pp = PicParams.make(20, 12)
    hb, owners = recon.pinned_batch(pp, 5)
    synth.generate(pp, 5, 31, out=hb)
    out = recon.PinnedArray((5, pp.frame_bytes), np.uint8)
    ref = oracle.reconstruct(hb)
    for _ in range(3):
        out.array[:] = 0
        gpu_ctx.submit(hb, out.array)
        gpu_ctx.wait()
        assert np.array_equal(ref, out.array)
    assert gpu_ctx.last_submit_ms > 0


def test_smoke_entry():
    import __graft_entry__
    __graft_entry__.smoke()


def test_independent_batches_overlap_on_different_streams(gpu_ctx):
    # up to four launches of a context may be in flight (round-robin control blocks guarded by completion events):
    # six batches over three streams, different inputs and outputs, every result must be its own
    import torch
    pp = PicParams.make(24, 14)
    streams = [torch.cuda.Stream() for _ in range(3)]
    batches = [synth.generate(pp, 3 + (k % 2), 6100 + k) for k in range(6)]
    dsoas = [recon.DeviceSoa(b) for b in batches]
    outs = [torch.zeros((b.n_frames, pp.frame_bytes), dtype=torch.uint8, device="cuda") for b in batches]
    torch.cuda.synchronize()
    for rep in range(2):
        for k in range(6):
            gpu_ctx.reconstruct_device(dsoas[k], outs[k], streams[k % 3].cuda_stream)
    gpu_ctx.wait()      # waits for every launch of the context, whatever stream it went to
    for k in range(6):
        assert np.array_equal(outs[k].cpu().numpy(), oracle.reconstruct(batches[k], threads=4)), k


def test_wait_covers_side_kernels_on_several_caller_streams(gpu_ctx):
    # residual-add and deblocking launches on two different caller streams, then only dryv_recon_wait — no torch
    # synchronisation — before the results are read back over a third stream-ordered copy: every launch of the context is
    # waited for, whatever stream it went to (one completion event per caller stream).
    import torch
    from oracle import deblock as dbl
    pp = PicParams.make(40, 30)
    b = synth.generate(pp, 6, 6400, qp_base=30)
    ds = recon.DeviceSoa(b)
    g = torch.Generator(device="cpu").manual_seed(5)
    pred = torch.randint(0, 256, (6, pp.frame_bytes), dtype=torch.uint8, generator=g)
    d_pred = pred.cuda()
    d_res = torch.zeros_like(d_pred)
    d_rec = torch.zeros_like(d_pred)
    gpu_ctx.reconstruct_device(ds, d_rec)
    gpu_ctx.wait()
    rec = d_rec.cpu().numpy()
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for rep in range(3):
        gpu_ctx.residual_add_device(ds, d_pred, d_res, sa.cuda_stream)
    gpu_ctx.deblock_device(ds, d_rec, 0, 0, sb.cuda_stream)
    gpu_ctx.wait()
    h_res = torch.empty_like(d_res, device="cpu").pin_memory()
    h_db = torch.empty_like(d_rec, device="cpu").pin_memory()
    sc = torch.cuda.Stream()   # a stream that is ordered behind nothing but the wait above
    with torch.cuda.stream(sc):
        h_res.copy_(d_res, non_blocking=True)
        h_db.copy_(d_rec, non_blocking=True)
    sc.synchronize()
    assert np.array_equal(h_res.numpy(), oracle.residual_add(b, pred.numpy()))
    for f in range(6):
        sl = slice(f * pp.n_mb, (f + 1) * pp.n_mb)
        want = dbl.deblock(rec[f], 40, 30, b.qp[sl], b.transform_size_8x8_flag[sl], 0, 0, 0, 0)
        assert np.array_equal(h_db[f].numpy(), want), f


# ---- the split formulation (DRYV_SPLIT=1: residual-fields kernel + one-warp-per-row walkers, split_kernels.cuh) --------
def test_split_path_is_bit_exact(monkeypatch):
    """Off by default (measured slower, profiles/r02_split_path.txt), but it is the same arithmetic through a second
    schedule of the same dependencies, so it is held to the same bar: classes, ragged sizes, stress macroblocks, host and
    device entry points, and the unsupported-syntax status."""
    monkeypatch.setenv("DRYV_SPLIT", "1")
    ctx = recon.ReconContext(0)  # the switch is read when the context is created
    try:
        for w, h, n, kw in [(1, 1, 2, {}), (2, 1, 2, {}), (1, 9, 2, {}), (17, 5, 3, {}), (40, 23, 3, dict(stress_pct=60)),
                            (9, 6, 2, dict(pct_i4x4=100, pct_i8x8=0)), (9, 6, 2, dict(pct_i4x4=0, pct_i8x8=100)),
                            (9, 6, 2, dict(pct_i4x4=0, pct_i8x8=0, qp_base=45)), (120, 68, 6, {})]:
            check(ctx, synth.generate(PicParams.make(w, h, cb_off=1, cr_off=-2), n, 7000 + 13 * w + h, **kw), host=(w < 100))
        b = synth.generate(PicParams.make(6, 4), 1, 5)
        b.mb_type[7] = 25  # I_PCM
        with pytest.raises(recon.ReconError) as e:
            ctx.reconstruct(b)
        assert e.value.code == recon.ERR_UNSUPPORTED
        check(ctx, synth.generate(PicParams.make(6, 4), 1, 6), device=False)  # the context stays usable
    finally:
        ctx.close()
