"""Cross-check against an independent, conformant H.264 decoder (libavcodec through cv2).

The reference ships no test vectors and cannot be built here, so nothing of its own pins the oracle. What can be
pinned is everything the reference does BY THE STANDARD: tests/avc/stream.py writes real CABAC High-profile I-slice
streams from the syntax buffers, libavcodec decodes them, and
  * the standard-conformant numpy model (oracle/spec_model.py, quirks=False) must equal libavcodec's luma on every stream;
  * the C oracle (= dryv's behaviour) must equal libavcodec's luma wherever dryv's one luma deviation (SURVEY quirk Q2:
    Intra8x8 in macroblock column 0) cannot fire, and where it can, the first differing macroblock must be such a one;
  * a committed fixture (stream + libavcodec's luma, tests/golden/avc/) holds the oracle — and, with -m gpu, the CUDA
    path — to libavcodec's output even where cv2 is missing.
cv2 exposes the decoder's luma plane only (tests/avc/decode.py); chroma is pinned through the BGR pictures instead: our planar
pictures, written as a YUV4MPEG2 file, go through the same libavformat + swscale conversion as the decoded stream, and
  * the standard model must give the same BGR pictures as libavcodec on every stream (one flipped chroma LSB is seen);
  * the C oracle must do so wherever dryv's one chroma deviation (the "> 0" availability tests of the
    chroma DC predictor) cannot fire, i.e. while no reconstructed chroma sample is 0; where it fires, luma is untouched and
    the first differing chroma macroblock is DC-predicted."""
import os

import numpy as np
import pytest

import oracle
from avc import decode, stream
from dryv_b200 import synth
from dryv_b200.abi import FIELDS, PicParams, SyntaxBatch
from oracle import spec_model

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "avc", "crosscheck.npz")
needs_libavcodec = pytest.mark.skipif(not decode.available(), reason="cv2 with the FFmpeg backend is not available")


def luma_of(frames, pp):
    n = pp.n_mb * 256
    return frames[:, :n].reshape(frames.shape[0], pp.pic_height_in_mbs * 16, pp.pic_width_in_mbs * 16)


def load_fixture():
    z = np.load(FIXTURE)
    pp = PicParams.make(int(z["w_mbs"]), int(z["h_mbs"]), int(z["cb_off"]), int(z["cr_off"]))
    b = SyntaxBatch(pp, int(z["n_frames"]), *[np.ascontiguousarray(z[f]) for f in FIELDS])
    return b, z["stream"].tobytes(), z["libavcodec_luma"]


CASES = [
    dict(w=4, h=3, n=1, seed=1, pct_i4x4=0, pct_i8x8=0, zero_residual=True),      # Intra16x16 prediction only
    dict(w=5, h=4, n=2, seed=3, pct_i4x4=0, pct_i8x8=0),                          # Intra16x16 + DC/AC residual
    dict(w=5, h=4, n=2, seed=4, pct_i4x4=100, pct_i8x8=0),                        # Intra4x4
    dict(w=5, h=4, n=2, seed=5, pct_i4x4=0, pct_i8x8=100),                        # Intra8x8 (Q2 fires in column 0)
    dict(w=7, h=5, n=3, seed=6),                                                  # the bench mix
    dict(w=6, h=4, n=2, seed=7, cb=3, cr=-4, qp_base=18, stress_pct=40),          # chroma QP offsets, large levels
    dict(w=6, h=4, n=2, seed=8, qp_base=40),                                      # sparse levels
    dict(w=1, h=1, n=2, seed=9), dict(w=1, h=5, n=1, seed=10), dict(w=6, h=1, n=1, seed=11),
]


def make(case, **extra):
    c = dict(case, **extra)
    pp = PicParams.make(c.pop("w"), c.pop("h"), c.pop("cb", 0), c.pop("cr", 0))
    b = synth.generate(pp, c.pop("n"), 7000 + c.pop("seed"), **c)
    return pp, b, stream.encode_stream(b)


@needs_libavcodec
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_libavcodec_equals_standard_model(case):
    pp, b, data = make(case)
    got = decode.decode_luma(data, b.n_frames, pp.pic_width_in_mbs * 16, pp.pic_height_in_mbs * 16)
    assert np.array_equal(got, luma_of(spec_model.reconstruct(b, quirks=False), pp))


@needs_libavcodec
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_oracle_luma_equals_libavcodec_where_dryv_follows_the_standard(case):
    pp, b, data = make(case, standard_only=True)
    got = decode.decode_luma(data, b.n_frames, pp.pic_width_in_mbs * 16, pp.pic_height_in_mbs * 16)
    assert np.array_equal(got, luma_of(oracle.reconstruct(b), pp))


@needs_libavcodec
def test_q2_is_the_only_luma_difference():
    pp, b, data = make(dict(w=6, h=5, n=3, seed=12, pct_i4x4=20, pct_i8x8=60))
    W, H = pp.pic_width_in_mbs, pp.pic_height_in_mbs
    got = decode.decode_luma(data, b.n_frames, W * 16, H * 16)
    ours = luma_of(oracle.reconstruct(b), pp)
    assert np.array_equal(ours, luma_of(spec_model.reconstruct(b, quirks=True), pp))
    fired = 0
    for f in range(b.n_frames):
        bad = (got[f] != ours[f]).reshape(H, 16, W, 16).any(axis=(1, 3))
        if not bad.any():
            continue
        y, x = np.argwhere(bad)[0]          # first differing macroblock in decoding order
        idx = f * pp.n_mb + y * W + x
        assert x == 0 and b.mb_type[idx] == 0 and b.transform_size_8x8_flag[idx] == 1
        fired += 1
    assert fired > 0


def test_stream_writer_is_deterministic_and_matches_fixture():
    b, data, _ = load_fixture()
    assert stream.encode_stream(b) == data          # the fixture's syntax is already canonical
    assert data[:5] == b"\x00\x00\x00\x01\x67" and b"\x00\x00\x00\x01\x68" in data and b"\x00\x00\x00\x01\x65" in data


def test_oracle_equals_libavcodec_fixture():
    b, _, luma = load_fixture()
    assert np.array_equal(luma_of(oracle.reconstruct(b), b.pp), luma)
    assert np.array_equal(luma_of(spec_model.reconstruct(b, quirks=False), b.pp), luma)


@needs_libavcodec
def test_fixture_still_decodes_to_the_same_luma():
    b, data, luma = load_fixture()
    assert np.array_equal(decode.decode_luma(data, b.n_frames, luma.shape[2], luma.shape[1]), luma)


# ---- per-class fixtures: libavcodec's luma AND BGR pictures (chroma) of Intra4x4-only / Intra8x8-only / Intra16x16-only streams
# with chroma QP offsets, decoded in the build container (tests/golden/avc/make_avc_golden.py); the reference follows the
# H.264 text on every sample of these pictures (no Intra8x8 in column 0, no chroma sample equal to 0)
CLASS_FIXTURE = os.path.join(os.path.dirname(FIXTURE), "classes.npz")


def load_class_fixture(name):
    z = np.load(CLASS_FIXTURE)
    pp = PicParams.make(int(z["w_mbs"]), int(z["h_mbs"]), int(z["cb_off"]), int(z["cr_off"]))
    b = SyntaxBatch(pp, int(z["n_frames"]), *[np.ascontiguousarray(z[name + "_" + f]) for f in FIELDS])
    return b, z[name + "_stream"].tobytes(), z[name + "_luma"], z[name + "_bgr"]


def check_against_class_fixture(frames, b, luma, bgr):
    pp = b.pp
    assert np.array_equal(luma_of(frames, pp), luma)
    if decode.available():   # the chroma planes, through the same swscale conversion libavcodec's pictures went through
        assert np.array_equal(decode.bgr_of_pictures(frames, pp.pic_width_in_mbs * 16, pp.pic_height_in_mbs * 16), bgr)


@pytest.mark.parametrize("name", ["i4x4", "i8x8", "i16x16"])
def test_oracle_equals_libavcodec_class_fixture(name):
    b, data, luma, bgr = load_class_fixture(name)
    assert stream.encode_stream(b) == data
    check_against_class_fixture(oracle.reconstruct(b), b, luma, bgr)
    assert np.array_equal(spec_model.reconstruct(b, quirks=False), oracle.reconstruct(b))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["i4x4", "i8x8", "i16x16"])
def test_cuda_path_equals_libavcodec_class_fixture(gpu_ctx, name):
    b, _, luma, bgr = load_class_fixture(name)
    check_against_class_fixture(gpu_ctx.reconstruct(b), b, luma, bgr)


@pytest.mark.gpu
def test_cuda_path_equals_libavcodec_fixture(gpu_ctx):
    b, _, luma = load_fixture()
    assert np.array_equal(luma_of(gpu_ctx.reconstruct(b), b.pp), luma)


@pytest.mark.gpu
@needs_libavcodec
def test_cuda_path_equals_libavcodec_live(gpu_ctx):
    pp, b, data = make(dict(w=20, h=12, n=4, seed=13), standard_only=True)
    got = decode.decode_luma(data, b.n_frames, pp.pic_width_in_mbs * 16, pp.pic_height_in_mbs * 16)
    assert np.array_equal(got, luma_of(gpu_ctx.reconstruct(b), pp))


# ---- chroma: the planes themselves are not reachable through cv2, the BGR pictures swscale makes of them are (avc/decode.py) --
def bgr_equal(frames, pp, data):
    """libavcodec's decode of `data`, converted to BGR, equals the same conversion of our planar pictures."""
    W, H = pp.pic_width_in_mbs * 16, pp.pic_height_in_mbs * 16
    theirs = decode.decode_bgr(data, len(frames))
    ours = decode.bgr_of_pictures(frames, W, H)
    return theirs.shape == ours.shape and bool(np.array_equal(theirs, ours))


@needs_libavcodec
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_libavcodec_equals_standard_model_in_chroma_too(case):
    pp, b, data = make(case)
    assert bgr_equal(spec_model.reconstruct(b, quirks=False), pp, data)


@needs_libavcodec
def test_bgr_comparison_sees_single_chroma_lsbs():
    """Sensitivity of the comparison above: flipping the lowest bit of one chroma sample changes the BGR picture."""
    pp, b, data = make(dict(w=5, h=4, n=1, seed=3, pct_i4x4=0, pct_i8x8=0))
    good = spec_model.reconstruct(b, quirks=False)
    assert bgr_equal(good, pp, data)
    rng = np.random.default_rng(1)
    n_luma = pp.n_mb * 256
    seen = 0
    for pos in rng.integers(n_luma, pp.frame_bytes, 12):
        bad = good.copy()
        bad[0, pos] ^= 1
        seen += not bgr_equal(bad, pp, data)
    assert seen >= 11     # a sample whose whole neighbourhood is clipped in B / R may hide


@needs_libavcodec
def test_q3_is_the_only_chroma_difference():
    """dryv's chroma deviates from the standard only through Q3 (the "> 0" availability tests of the chroma DC predictor,
    trans_chroma.rs:209-268): the C oracle equals libavcodec in every plane on pictures where no reconstructed chroma
    sample is 0 next to a DC-predicted block, and where it differs the first difference is a DC-predicted chroma block."""
    # mid-range content, no stress macroblocks: no chroma sample reaches 0, Q3 cannot fire, Q2 is excluded by standard_only
    pp, b, data = make(dict(w=7, h=5, n=3, seed=6, stress_pct=0), standard_only=True)
    ours = oracle.reconstruct(b)
    n_luma = pp.n_mb * 256
    assert ours[:, n_luma:].min() > 0
    assert bgr_equal(ours, pp, data)
    # with stress macroblocks chroma clamps to 0 and Q3 fires: the oracle follows the quirks model, not the standard
    pp, b, data = make(dict(w=6, h=4, n=3, seed=5, stress_pct=60), standard_only=True)
    ours = oracle.reconstruct(b)
    assert np.array_equal(ours, spec_model.reconstruct(b, quirks=True))
    std = spec_model.reconstruct(b, quirks=False)
    assert bgr_equal(std, pp, data)
    diff = ours != std
    assert not diff[:, :pp.n_mb * 256].any()         # luma untouched
    assert diff.any()                                # Q3 fired; the first differing chroma macroblock uses DC prediction
    if True:
        W, H = pp.pic_width_in_mbs, pp.pic_height_in_mbs
        f = int(np.argwhere(diff.any(axis=1))[0][0])
        cb = diff[f, pp.n_mb * 256:pp.n_mb * 320].reshape(H, 8, W, 8).any(axis=(1, 3))
        cr = diff[f, pp.n_mb * 320:].reshape(H, 8, W, 8).any(axis=(1, 3))
        y, x = np.argwhere(cb | cr)[0]
        assert b.intra_chroma_pred_mode[f * pp.n_mb + y * W + x] == 0


@pytest.mark.gpu
@needs_libavcodec
def test_cuda_path_equals_libavcodec_live_in_chroma_too(gpu_ctx):
    pp, b, data = make(dict(w=7, h=5, n=3, seed=6, stress_pct=0), standard_only=True)   # Q2 / Q3 cannot fire (see above)
    assert bgr_equal(gpu_ctx.reconstruct(b), pp, data)
