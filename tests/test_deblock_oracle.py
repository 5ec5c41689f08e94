"""Groundwork for SURVEY.md §8(f) next-4: the deblocking-filter oracle (oracle/deblock.py) against libavcodec on streams
that enable the in-loop filter. There is no deblocking kernel yet and the reference has none, so the product path keeps the
filter off; this pins the oracle a later kernel will be checked against."""
import numpy as np
import pytest

import oracle
from avc import decode, stream
from dryv_b200 import synth
from dryv_b200.abi import PicParams
from oracle import deblock as dbl

needs_libavcodec = pytest.mark.skipif(not decode.available(), reason="cv2 with the FFmpeg backend is not available")

CASES = [
    dict(w=5, h=4, n=2, seed=1, qp_base=30),
    dict(w=6, h=4, n=2, seed=2, qp_base=38, pct_i4x4=100, pct_i8x8=0),
    dict(w=6, h=4, n=1, seed=3, qp_base=34, pct_i4x4=0, pct_i8x8=100),
    dict(w=5, h=3, n=1, seed=4, qp_base=24, pct_i4x4=0, pct_i8x8=0),
    dict(w=7, h=5, n=2, seed=5, qp_base=44, stress_pct=30),
    dict(w=4, h=3, n=1, seed=6, qp_base=36, offs=(3, -2)),
    dict(w=4, h=3, n=1, seed=7, qp_base=28, offs=(-3, 4)),
    dict(w=1, h=1, n=1, seed=8, qp_base=40), dict(w=1, h=4, n=1, seed=9, qp_base=40), dict(w=5, h=1, n=1, seed=10, qp_base=40),
]


@needs_libavcodec
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_deblock_oracle_equals_libavcodec_luma(case):
    c = dict(case)
    w, h, n = c.pop("w"), c.pop("h"), c.pop("n")
    offs = c.pop("offs", (0, 0))
    pp = PicParams.make(w, h)
    b = synth.generate(pp, n, 9000 + c.pop("seed"), standard_only=True, **c)
    s = stream.encode_stream(b, deblock=offs)     # canonicalises b (qp of macroblocks without residual, ...)
    theirs = decode.decode_luma(s, n, 16 * w, 16 * h)
    frames = oracle.reconstruct(b)
    changed = 0
    for f in range(n):
        sl = slice(f * pp.n_mb, (f + 1) * pp.n_mb)
        out = dbl.deblock(frames[f], w, h, b.qp[sl], b.transform_size_8x8_flag[sl], alpha_div2=offs[0], beta_div2=offs[1])
        ours = out[:256 * pp.n_mb].reshape(16 * h, 16 * w)
        assert np.array_equal(ours, theirs[f]), f
        changed += int((out != frames[f]).sum())
    if w * h > 1:
        assert changed > 0    # the filter did something


def test_deblock_oracle_leaves_flat_pictures_alone():
    frame = np.full(2 * 2 * 384, 77, np.uint8)
    out = dbl.deblock(frame, 2, 2, np.full(4, 40), np.zeros(4, np.uint8))
    assert np.array_equal(out, frame)


# ---- the CUDA post-pass (dryv_recon_deblock_device) against the oracle above, and end to end against libavcodec ----------
GPU_CASES = [
    dict(w=5, h=4, n=3, seed=21, qp_base=30),
    dict(w=9, h=7, n=2, seed=22, qp_base=38, pct_i4x4=100, pct_i8x8=0),
    dict(w=6, h=5, n=2, seed=23, qp_base=34, pct_i4x4=0, pct_i8x8=100, cb=4, cr=-5),
    dict(w=8, h=6, n=2, seed=24, qp_base=44, stress_pct=30, offs=(3, -2)),
    dict(w=7, h=3, n=2, seed=25, qp_base=26, offs=(-4, 5)),
    dict(w=1, h=1, n=2, seed=26, qp_base=40), dict(w=1, h=6, n=1, seed=27, qp_base=40), dict(w=7, h=1, n=1, seed=28, qp_base=40),
    dict(w=2, h=2, n=5, seed=29, qp_base=36),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_deblock_kernel_matches_oracle(case):
    import torch
    from dryv_b200 import recon
    c = dict(case)
    w, h, n = c.pop("w"), c.pop("h"), c.pop("n")
    offs = c.pop("offs", (0, 0))
    pp = PicParams.make(w, h, c.pop("cb", 0), c.pop("cr", 0))
    b = synth.generate(pp, n, 9100 + c.pop("seed"), **c)
    frames = oracle.reconstruct(b)
    ctx = recon.ReconContext(0)
    ds = recon.DeviceSoa(b)
    d = torch.zeros((n, pp.frame_bytes), dtype=torch.uint8, device="cuda")
    ctx.reconstruct_device(ds, d)
    ctx.deblock_device(ds, d, offs[0], offs[1])      # same (context) stream: ordered behind the reconstruction
    ctx.wait()
    got = d.cpu().numpy()
    for f in range(n):
        sl = slice(f * pp.n_mb, (f + 1) * pp.n_mb)
        want = dbl.deblock(frames[f], w, h, b.qp[sl], b.transform_size_8x8_flag[sl], int(pp.chroma_qp_index_offset),
                           int(pp.second_chroma_qp_index_offset), offs[0], offs[1])
        assert np.array_equal(got[f], want), (f, int((got[f] != want).sum()))
    # a second pass over fresh pictures through the same context (counters are re-armed per launch)
    ctx.reconstruct_device(ds, d)
    ctx.deblock_device(ds, d, offs[0], offs[1])
    ctx.wait()
    assert np.array_equal(d.cpu().numpy(), got)
    ctx.close()


@pytest.mark.gpu
def test_deblock_as_a_mode_of_the_host_submit_path(monkeypatch):
    """dryv_recon_set_deblock: reconstruct() from host buffers returns filtered pictures, slot by slot (three pipeline
    stages of an odd number of pictures each), and unfiltered ones again once the mode is off."""
    from dryv_b200 import recon
    monkeypatch.setenv("DRYV_CHUNK_MB", "1")
    pp = PicParams.make(6, 5, 2, -1)
    n = 101
    b = synth.generate(pp, n, 9177, qp_base=33)
    frames = oracle.reconstruct(b)
    want = np.stack([dbl.deblock(frames[f], 6, 5, b.qp[f * pp.n_mb:(f + 1) * pp.n_mb],
                                 b.transform_size_8x8_flag[f * pp.n_mb:(f + 1) * pp.n_mb], 2, -1, 2, -3) for f in range(n)])
    ctx = recon.ReconContext(0)
    ctx.set_deblock(True, 2, -3)
    assert np.array_equal(ctx.reconstruct(b), want)
    assert np.array_equal(ctx.reconstruct(b), want)
    ctx.set_deblock(False)
    assert np.array_equal(ctx.reconstruct(b), frames)
    with pytest.raises(Exception):
        ctx.set_deblock(True, 7, 0)
    ctx.close()


@pytest.mark.gpu
def test_deblock_kernel_on_a_batch_larger_than_the_resident_rows():
    """More macroblock rows than the launch has warps: rows are dealt by ticket, later rows start as earlier ones finish.
    Checked by a property the domain offers: every picture of a batch of identical pictures comes out identical to the
    first, and the first equals the oracle."""
    import torch
    from dryv_b200 import recon
    pp = PicParams.make(4, 40)
    one = synth.generate(pp, 1, 777, qp_base=36)
    n = 160                                  # 6400 rows > 148 SMs x 8 CTAs x 4 warps
    from dryv_b200.abi import FIELDS, SyntaxBatch
    b = SyntaxBatch(pp, n, *[np.ascontiguousarray(np.concatenate([getattr(one, f)] * n)) for f in FIELDS])
    ctx = recon.ReconContext(0)
    ds = recon.DeviceSoa(b)
    d = torch.zeros((n, pp.frame_bytes), dtype=torch.uint8, device="cuda")
    ctx.reconstruct_device(ds, d)
    ctx.deblock_device(ds, d)
    ctx.wait()
    want = dbl.deblock(oracle.reconstruct(one)[0], 4, 40, one.qp, one.transform_size_8x8_flag)
    assert np.array_equal(d[0].cpu().numpy(), want)
    assert bool((d == d[0:1]).all())
    ctx.close()


@pytest.mark.gpu
@needs_libavcodec
def test_bytes_to_deblocked_pictures_equal_libavcodec():
    """Stream with the filter enabled -> CABAC host -> reconstruction -> deblocking kernel == libavcodec's luma."""
    import torch
    from dryv_b200 import host, recon
    pp = PicParams.make(8, 6)
    b = synth.generate(pp, 3, 4242, standard_only=True, qp_base=34)
    s = stream.encode_stream(b, deblock=(1, -1))
    parsed = host.parse(s)
    ctx = recon.ReconContext(0)
    ds = recon.DeviceSoa(parsed)
    d = torch.zeros((3, pp.frame_bytes), dtype=torch.uint8, device="cuda")
    ctx.reconstruct_device(ds, d)
    ctx.deblock_device(ds, d, 1, -1)
    ctx.wait()
    theirs = decode.decode_luma(s, 3, 128, 96)
    assert np.array_equal(d.cpu().numpy()[:, :128 * 96].reshape(3, 96, 128), theirs)
    ctx.close()


@needs_libavcodec
@pytest.mark.parametrize("case", [
    dict(w=5, h=4, n=2, seed=31, qp_base=30),
    dict(w=6, h=4, n=2, seed=32, qp_base=38, cb=4, cr=-5),
    dict(w=6, h=4, n=1, seed=33, qp_base=34, pct_i4x4=0, pct_i8x8=100, cb=-6, cr=7),
    dict(w=7, h=5, n=2, seed=34, qp_base=44, stress_pct=30, offs=(2, -3)),
    dict(w=4, h=3, n=1, seed=35, qp_base=24, offs=(-2, 5), cb=12, cr=-12),
], ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_deblock_oracle_equals_libavcodec_in_chroma_too(case):
    """Chroma through the BGR pictures (avc/decode.py): the filtered planes, written as YUV4MPEG2, must convert to the BGR
    pictures libavcodec's decode of the filter-enabled stream converts to. The unfiltered input is the standard model's
    reconstruction (dryv's chroma quirk Q3 is not libavcodec's behaviour)."""
    from oracle import spec_model
    c = dict(case)
    w, h, n = c.pop("w"), c.pop("h"), c.pop("n")
    offs = c.pop("offs", (0, 0))
    pp = PicParams.make(w, h, c.pop("cb", 0), c.pop("cr", 0))
    b = synth.generate(pp, n, 9200 + c.pop("seed"), standard_only=True, **c)
    s = stream.encode_stream(b, deblock=offs)
    frames = spec_model.reconstruct(b, quirks=False)
    out = np.stack([dbl.deblock(frames[f], w, h, b.qp[f * pp.n_mb:(f + 1) * pp.n_mb],
                                b.transform_size_8x8_flag[f * pp.n_mb:(f + 1) * pp.n_mb], int(pp.chroma_qp_index_offset),
                                int(pp.second_chroma_qp_index_offset), offs[0], offs[1]) for f in range(n)])
    theirs = decode.decode_bgr(s, n)
    assert np.array_equal(theirs, decode.bgr_of_pictures(out, 16 * w, 16 * h))
    assert not np.array_equal(theirs, decode.bgr_of_pictures(frames, 16 * w, 16 * h))   # the filter is visible in BGR
