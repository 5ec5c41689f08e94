"""The CPU host side of the path (include/dryv_cabac_host.h, dryv_b200/csrc/cabac_host.cpp): Annex-B H.264 bytes ->
syntax buffers. Checked as the inverse of the stream writer libavcodec accepts (tests/avc/stream.py), against the
committed libavcodec fixture, and — with -m gpu — end to end: bytes -> CABAC parse -> compact level stream -> CUDA
reconstruction -> the YUV frame dryv writes to ./temp/yuv_frame."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
from avc import stream
from dryv_b200 import host, recon, synth
from dryv_b200.abi import FIELDS, PicParams
from test_libavcodec_crosscheck import load_fixture, luma_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def coded_pred_syntax(b):
    """pred_syntax with its don't-care bits cleared: rem bits of flagged blocks, entries the class does not code."""
    out = np.zeros_like(b.pred_syntax)
    nxn = b.mb_type == 0
    k = np.where(b.transform_size_8x8_flag != 0, 4, 16)
    used = nxn[:, None] & (np.arange(16)[None, :] < k[:, None])
    ps = b.pred_syntax & 15
    out[used] = np.where(ps & 8, 8, ps)[used]
    return out


def assert_same_syntax(a, b):
    for f in FIELDS:
        if f != "pred_syntax":
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    assert np.array_equal(coded_pred_syntax(a), coded_pred_syntax(b))


def test_header_symbols_are_exported(recon_lib):
    text = open(os.path.join(ROOT, "include", "dryv_cabac_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"\b(dryv_cabac_\w+)\s*\(", text)))
    assert names == sorted(recon.HOST_EXPORTS)
    for n in names:
        assert hasattr(recon_lib, n)


@pytest.mark.parametrize("w,h,n,kw", [
    (4, 3, 1, dict(pct_i4x4=0, pct_i8x8=0, zero_residual=True)),
    (5, 4, 2, dict(pct_i4x4=0, pct_i8x8=0)),
    (5, 4, 2, dict(pct_i4x4=100, pct_i8x8=0)),
    (5, 4, 2, dict(pct_i4x4=0, pct_i8x8=100)),
    (7, 5, 3, dict()),
    (6, 4, 2, dict(qp_base=18, stress_pct=60)),
    (6, 4, 2, dict(qp_base=2, qp_jitter=2)),
    (6, 4, 2, dict(qp_base=49, qp_jitter=2)),
    (1, 1, 2, dict()), (1, 6, 1, dict()), (9, 1, 1, dict()),
    (40, 23, 2, dict()),          # 640x368: the size of BASELINE configs[0]
])
def test_parse_inverts_the_stream_writer(recon_lib, w, h, n, kw):
    pp = PicParams.make(w, h, 3, -2)
    b = synth.generate(pp, n, 100 + w * 7 + h, **kw)
    data = stream.encode_stream(b)          # canonicalises b
    pp2, n2 = host.scan(data)
    assert (pp2.pic_width_in_mbs, pp2.pic_height_in_mbs, pp2.chroma_qp_index_offset,
            pp2.second_chroma_qp_index_offset, n2) == (w, h, 3, -2, n)
    assert list(pp2.scaling_list4x4) == [16] * 16 and list(pp2.scaling_list8x8) == [16] * 64
    got = host.parse(data, threads=1)
    assert_same_syntax(b, got)
    assert np.array_equal(oracle.reconstruct(got), oracle.reconstruct(b))
    if n > 1:
        assert_same_syntax(host.parse(data, threads=4), got)
        tail = host.parse(data, threads=1, first=1, count=n - 1)     # a rank's share of the pictures
        assert_same_syntax(tail, got.frames(1, n))


def test_parse_compact_emits_the_records_pack_levels_would(recon_lib):
    pp = PicParams.make(9, 7, 1, 0)
    b = synth.generate(pp, 4, 5150, stress_pct=30)
    data = stream.encode_stream(b)
    dense = host.parse(data, threads=1)
    for threads, first, count in ((1, 0, None), (3, 0, None), (2, 1, 2)):
        syn, lv = host.parse_compact(data, threads=threads, first=first, count=count)
        ref = dense.frames(first, first + syn.n_frames)
        want = recon.pack_levels(ref.coeff, threads=1)
        assert np.array_equal(lv.offset, want.offset)
        assert np.array_equal(lv.stream[:int(lv.offset[-1])], want.stream[:int(want.offset[-1])])
        assert np.array_equal(lv.unpack(), ref.coeff)
        for f in ("mb_type", "transform_size_8x8_flag", "intra_chroma_pred_mode", "qp", "pred_syntax"):
            assert np.array_equal(getattr(syn, f), getattr(ref, f)), f
        assert syn.coeff.size == 0


def test_parse_of_the_libavcodec_fixture(recon_lib):
    b, data, luma = load_fixture()
    got = host.parse(data)
    assert_same_syntax(b, got)
    assert np.array_equal(luma_of(oracle.reconstruct(got), got.pp), luma)     # bytes -> parse -> oracle == libavcodec


def test_malformed_and_unsupported_streams(recon_lib):
    b = synth.generate(PicParams.make(4, 3), 2, 77)
    data = stream.encode_stream(b)
    pp, n = host.scan(data)
    buf = lambda d: np.frombuffer(d, np.uint8)  # noqa: E731
    out = synth.generate(pp, n, 0)

    def parse(d, pp=pp, n=n):
        a = buf(d)
        return recon_lib.dryv_cabac_parse(a.ctypes.data, a.size, C.byref(pp), n, out.mb_type.ctypes.data,
                                          out.transform_size_8x8_flag.ctypes.data, out.intra_chroma_pred_mode.ctypes.data,
                                          out.qp.ctypes.data, out.pred_syntax.ctypes.data, out.coeff.ctypes.data, 1)

    assert parse(data) == recon.OK
    assert parse(data[:len(data) * 2 // 3]) == recon.ERR_ARG                  # second picture cut short
    assert parse(data, n=3) == recon.ERR_ARG                                  # wrong picture count
    assert parse(data, pp=PicParams.make(5, 3)) == recon.ERR_ARG              # wrong geometry
    assert parse(b"\x00" * 64) == recon.ERR_ARG                               # no NAL units at all
    cavlc = bytearray(data)                                                   # PPS with entropy_coding_mode_flag = 0
    i = data.index(b"\x00\x00\x00\x01\x68") + 5
    assert cavlc[i] & 0x20
    cavlc[i] &= ~0x20
    assert parse(bytes(cavlc)) == recon.ERR_UNSUPPORTED
    non_idr = bytearray(data)                                                 # turn the first IDR slice into a non-IDR one
    j = data.index(b"\x00\x00\x00\x01\x65") + 4
    non_idr[j] = 0x61
    assert parse(bytes(non_idr)) == recon.ERR_UNSUPPORTED
    with pytest.raises(recon.ReconError):
        host.scan(b"\x00\x00\x00\x01\x09\x10")
    assert recon_lib.dryv_cabac_scan(None, 0, None, None) == recon.ERR_ARG


@pytest.mark.gpu
def test_bytes_to_yuv_frame_through_the_whole_path(gpu_ctx, tmp_path):
    # BASELINE configs[0] on a self-made file: 640x368 8-bit 4:2:0 CABAC High-profile MP4 -> ./temp/yuv_frame
    pp = PicParams.make(40, 23)
    b = synth.generate(pp, 3, 360, standard_only=True)
    data = stream.encode_stream(b)
    from avc import mp4
    movie = mp4.mux(data, 640, 368)          # the kind of file `dryv <file>` opens
    parsed, levels = host.parse_compact(movie)   # demux + CABAC parse straight into the compact level stream
    out = gpu_ctx.reconstruct_compact(parsed, levels)
    assert np.array_equal(out, oracle.reconstruct(b))
    path = tmp_path / "temp" / "yuv_frame"
    recon.write_yuv_file(out[0], str(path))
    assert os.path.getsize(path) == 640 * 368 * 3 // 2 == 353280
    from avc import decode
    if decode.available():
        assert np.array_equal(decode.decode_luma(data, 3, 640, 368), luma_of(out, pp))


@pytest.mark.gpu
def test_fixture_bytes_to_cuda_equals_libavcodec(gpu_ctx):
    _, data, luma = load_fixture()
    parsed = host.parse(data)
    assert np.array_equal(luma_of(gpu_ctx.reconstruct(parsed), parsed.pp), luma)


def test_corrupted_streams_never_crash(recon_lib):
    # every entry point must survive arbitrary bytes: bounded loops, bounds-checked reads, an error code at worst
    b = synth.generate(PicParams.make(5, 4), 2, 4242, stress_pct=50)
    data = bytearray(stream.encode_stream(b))
    pp, n = host.scan(bytes(data))
    out = synth.generate(pp, n, 0)
    rng = np.random.default_rng(99)
    seen = set()
    for trial in range(300):
        d = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            kind = int(rng.integers(0, 3))
            pos = int(rng.integers(0, len(d)))
            if kind == 0:
                d[pos] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                d[pos] = int(rng.integers(0, 256))
            else:
                del d[pos:pos + int(rng.integers(1, 40))]
        a = np.frombuffer(bytes(d), np.uint8)
        pp2, n2 = PicParams(), C.c_uint32()
        rc = recon_lib.dryv_cabac_scan(a.ctypes.data, a.size, C.byref(pp2), C.byref(n2))
        assert rc in (recon.OK, recon.ERR_ARG, recon.ERR_UNSUPPORTED)
        rc = recon_lib.dryv_cabac_parse(a.ctypes.data, a.size, C.byref(pp), n, out.mb_type.ctypes.data,
                                        out.transform_size_8x8_flag.ctypes.data, out.intra_chroma_pred_mode.ctypes.data,
                                        out.qp.ctypes.data, out.pred_syntax.ctypes.data, out.coeff.ctypes.data, 2)
        assert rc in (recon.OK, recon.ERR_ARG, recon.ERR_UNSUPPORTED)
        seen.add(rc)
        if rc == recon.OK:      # whatever was parsed must be syntax the reconstruction accepts
            assert out.mb_type.max() <= 24 and out.intra_chroma_pred_mode.max() <= 3 and out.qp.max() <= 51
    assert recon.ERR_ARG in seen


# ---- MP4 container front end (tests/avc/mp4.py muxes, dryv_cabac_scan / dryv_cabac_parse demux) -----------------------
def test_mp4_file_parses_like_the_annexb_stream(recon_lib):
    from avc import decode, mp4
    pp = PicParams.make(9, 6, -1, 2)
    b = synth.generate(pp, 4, 2718, standard_only=True)
    data = stream.encode_stream(b)
    movie = mp4.mux(data, 144, 96)
    assert movie[4:8] == b"ftyp" and b"avcC" in movie and b"mdat" in movie
    pp2, n2 = host.scan(movie)
    assert (pp2.pic_width_in_mbs, pp2.pic_height_in_mbs, pp2.chroma_qp_index_offset, pp2.second_chroma_qp_index_offset,
            n2) == (9, 6, -1, 2, 4)
    assert_same_syntax(host.parse(movie), host.parse(data))
    assert_same_syntax(host.parse(movie), b)
    if decode.available():      # libavformat + libavcodec agree that this is a valid MP4 with these pictures
        import cv2
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".mp4", delete=False) as f:
            f.write(movie)
        try:
            cap = cv2.VideoCapture(f.name, cv2.CAP_FFMPEG)
            assert cap.isOpened() and int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 4
            cap.set(cv2.CAP_PROP_CONVERT_RGB, 0)
            ref = luma_of(oracle.reconstruct(b), pp)
            for k in range(4):
                ok, fr = cap.read()
                assert ok and np.array_equal(np.asarray(fr).reshape(-1)[:144 * 96].reshape(96, 144), ref[k])
        finally:
            os.unlink(f.name)


def test_corrupted_mp4_files_never_crash(recon_lib):
    from avc import mp4
    b = synth.generate(PicParams.make(4, 3), 2, 31)
    movie = bytearray(mp4.mux(stream.encode_stream(b), 64, 48))
    pp, n = host.scan(bytes(movie))
    out = synth.generate(pp, n, 0)
    rng = np.random.default_rng(5)
    seen = set()
    for trial in range(300):
        d = bytearray(movie)
        for _ in range(int(rng.integers(1, 5))):
            pos = int(rng.integers(0, len(d)))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                d[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                d[pos:pos + 4] = int(rng.integers(0, 2 ** 32)).to_bytes(4, "big")
            else:
                del d[pos:]
        if len(d) < 16:
            continue
        a = np.frombuffer(bytes(d), np.uint8)
        rc = recon_lib.dryv_cabac_parse(a.ctypes.data, a.size, C.byref(pp), n, out.mb_type.ctypes.data,
                                        out.transform_size_8x8_flag.ctypes.data, out.intra_chroma_pred_mode.ctypes.data,
                                        out.qp.ctypes.data, out.pred_syntax.ctypes.data, out.coeff.ctypes.data, 1)
        assert rc in (recon.OK, recon.ERR_ARG, recon.ERR_UNSUPPORTED)
        seen.add(rc)
    assert recon.ERR_ARG in seen
