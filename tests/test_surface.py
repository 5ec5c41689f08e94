"""Output surface (include/dryv_recon.h dryv_surface): crop rectangle + I420 / NV12 layout.

CPU: validation, the SPS crop rectangle out of the host parser, and the crop semantics against libavcodec. GPU (marked):
the export kernels and the submit paths with a surface set, against oracle/surface.py applied to the oracle's pictures."""
import ctypes as C

import numpy as np
import pytest

import oracle as ref
from avc import decode, mp4, stream
from dryv_b200 import host, recon, synth
from dryv_b200.abi import SURFACE_I420, SURFACE_NV12, PicParams, Surface
from oracle import surface as osurf

needs_libavcodec = pytest.mark.skipif(not decode.available(), reason="cv2 with the FFmpeg backend is not available")


def test_surface_bytes_and_validation():
    lib = recon.load_library()
    assert lib.dryv_recon_surface_bytes(C.byref(Surface.make(1920, 1080))) == 1920 * 1080 * 3 // 2
    assert lib.dryv_recon_surface_bytes(C.byref(Surface.make(2, 2, 14, 14, SURFACE_NV12))) == 6
    for bad in (Surface.make(0, 16), Surface.make(16, 0), Surface.make(15, 16), Surface.make(16, 16, 1, 0),
                Surface.make(16, 16, 0, 3), Surface.make(16, 16, fmt=7), Surface.make(1 << 16, 16)):
        assert lib.dryv_recon_surface_bytes(C.byref(bad)) == 0
    assert lib.dryv_recon_surface_bytes(None) == 0


@pytest.mark.parametrize("crop", [None, (0, 0, 0, 4), (1, 2, 3, 4), (0, 7, 0, 0), (3, 0, 5, 0)])
def test_sps_crop_rectangle_from_the_host_parser(crop):
    pp = PicParams.make(6, 5)
    b = synth.generate(pp, 1, 42)
    s = stream.encode_stream(b, crop=crop)
    for data in (s, mp4.mux(s, 96, 80)):
        sf = host.surface(data)
        assert (sf.crop_left, sf.crop_top, sf.width, sf.height) == osurf.sps_rectangle(6, 5, crop)
        assert sf.format == SURFACE_I420
    # the crop fields do not disturb the parse
    got = host.parse(s)
    assert np.array_equal(got.coeff, b.coeff) and np.array_equal(got.mb_type, b.mb_type)


def test_sps_crop_that_leaves_nothing_is_rejected():
    pp = PicParams.make(2, 2)
    s = stream.encode_stream(synth.generate(pp, 1, 1), crop=(8, 8, 0, 0))
    with pytest.raises(recon.ReconError) as e:
        host.surface(s)
    assert e.value.code == recon.ERR_ARG


@needs_libavcodec
@pytest.mark.parametrize("crop", [(0, 0, 0, 4), (0, 3, 0, 0), (0, 5, 2, 3)])
def test_crop_semantics_match_libavcodec(crop):
    """libavcodec applies the SPS rectangle itself: its luma == oracle/surface.py applied to the reconstructed picture
    (right / bottom / top offsets; a left offset makes libavcodec keep extra columns unless it runs with unaligned output)."""
    pp = PicParams.make(7, 5)
    b = synth.generate(pp, 2, 77, standard_only=True)
    s = stream.encode_stream(b, crop=crop)
    l, t, w, h = osurf.sps_rectangle(7, 5, crop)
    theirs = decode.decode_luma(s, 2, w, h)
    frames = ref.reconstruct(b)
    for f in range(2):
        ours = osurf.export(frames[f], 7, 5, l, t, w, h)[:w * h].reshape(h, w)
        assert np.array_equal(ours, theirs[f])


SURFACES = [
    # (w_mbs, h_mbs, crop_left, crop_top, width, height)
    (120, 68, 0, 0, 1920, 1080),     # 1080p: the usual bottom crop, 16-byte vectors
    (120, 68, 0, 0, 1920, 1088),     # nothing cropped: a repack
    (8, 6, 16, 16, 96, 64),          # 16-byte aligned rectangle
    (8, 6, 8, 2, 104, 70),           # 8-byte luma vectors
    (8, 6, 4, 6, 100, 58),           # 4-byte
    (8, 6, 2, 4, 90, 50),            # 2-byte luma, byte-wise I420 chroma
    (8, 6, 126, 94, 2, 2),           # the last 2x2 samples
    (1, 1, 0, 0, 16, 16), (1, 1, 6, 10, 6, 2), (40, 23, 0, 0, 640, 360),
    (5, 4, 0, 0, 64, 64), (5, 3, 16, 0, 64, 32),   # odd macroblock count: coded chroma rows are only 8-byte aligned
]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [SURFACE_I420, SURFACE_NV12], ids=["i420", "nv12"])
@pytest.mark.parametrize("geo", SURFACES, ids=lambda g: "x".join(map(str, g)))
def test_export_device_matches_oracle(geo, fmt):
    import torch
    wm, hm, l, t, w, h = geo
    pp = PicParams.make(wm, hm)
    n = 3
    rng = np.random.default_rng(sum(geo) + fmt)
    frames = rng.integers(0, 256, (n, pp.frame_bytes), dtype=np.uint8)
    sf = Surface.make(w, h, l, t, fmt)
    d_yuv = torch.from_numpy(frames).cuda()
    d_out = torch.full((n, sf.nbytes), 0xA5, dtype=torch.uint8, device="cuda")
    ctx = recon.ReconContext(0)
    ctx.export_device(pp, d_yuv, n, sf, d_out)
    ctx.wait()
    got = d_out.cpu().numpy()
    for f in range(n):
        assert np.array_equal(got[f], osurf.export(frames[f], wm, hm, l, t, w, h, fmt)), f
    ctx.close()


@pytest.mark.gpu
def test_export_rejects_a_rectangle_outside_the_picture():
    import torch
    pp = PicParams.make(4, 4)
    ctx = recon.ReconContext(0)
    d_yuv = torch.zeros(pp.frame_bytes, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(2 * pp.frame_bytes, dtype=torch.uint8, device="cuda")
    for sf in (Surface.make(64, 64, 2, 0), Surface.make(66, 64), Surface.make(16, 16, 0, 50)):
        with pytest.raises(recon.ReconError) as e:
            ctx.export_device(pp, d_yuv, 1, sf, d_out)
        assert e.value.code == recon.ERR_ARG
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", [SURFACE_I420, SURFACE_NV12], ids=["i420", "nv12"])
def test_submit_hands_back_the_surface(fmt):
    pp = PicParams.make(12, 9)
    b = synth.generate(pp, 5, 2024)
    want = ref.reconstruct(b)
    sf = Surface.make(180, 136, 4, 2, fmt)
    ctx = recon.ReconContext(0)
    ctx.set_surface(sf)
    out = np.zeros((5, sf.nbytes), np.uint8)
    ctx.submit(b, out)
    ctx.wait()
    out_c = np.zeros_like(out)
    ctx.submit_compact(b, recon.pack_levels(b.coeff), out_c)
    ctx.wait()
    for f in range(5):
        exp = osurf.export(want[f], 12, 9, 4, 2, 180, 136, fmt)
        assert np.array_equal(out[f], exp) and np.array_equal(out_c[f], exp)
    # a surface that does not fit the pictures is refused by the submit, and the default comes back with None
    ctx.set_surface(Surface.make(400, 16))
    with pytest.raises(recon.ReconError):
        ctx.submit(b, np.zeros((5, 400 * 16 * 3 // 2), np.uint8))
    ctx.set_surface(None)
    assert np.array_equal(ctx.reconstruct(b), want)
    ctx.close()


@pytest.mark.gpu
def test_bytes_to_cropped_surface_through_the_whole_path():
    """MP4 bytes -> CABAC host -> GPU reconstruction -> SPS display rectangle, as 1080p streams carry it (bottom crop)."""
    pp = PicParams.make(8, 5)
    b = synth.generate(pp, 3, 99, standard_only=True)
    data = mp4.mux(stream.encode_stream(b, crop=(0, 0, 0, 4)), 128, 72)
    sf = host.surface(data)
    assert (sf.width, sf.height) == (128, 72)
    parsed = host.parse(data)
    ctx = recon.ReconContext(0)
    ctx.set_surface(sf)
    out = np.zeros((3, sf.nbytes), np.uint8)
    ctx.submit(parsed, out)
    ctx.wait()
    want = ref.reconstruct(b)
    for f in range(3):
        assert np.array_equal(out[f], osurf.export(want[f], 8, 5, 0, 0, 128, 72))
    if decode.available():
        theirs = decode.decode_luma(stream.encode_stream(b, crop=(0, 0, 0, 4)), 3, 128, 72)
        assert np.array_equal(out[:, :128 * 72].reshape(3, 72, 128), theirs)
    ctx.close()
