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


# ---- scaling matrices and parameter sets by id -----------------------------------------------------
DEFAULT_4X4_INTRA = [6, 13, 13, 20, 20, 20, 28, 28, 28, 28, 32, 32, 32, 37, 37, 42]   # Table 7-3 / atom/avcc/sps.rs:161-163
DEFAULT_8X8_INTRA = [6, 10, 10, 13, 11, 13, 16, 16, 16, 16, 18, 18, 18, 18, 18, 23, 23, 23, 23, 23, 23, 25, 25, 25, 25, 25,
                     25, 25, 27, 27, 27, 27, 27, 27, 27, 27, 29, 29, 29, 29, 29, 29, 29, 31, 31, 31, 31, 31, 31, 33, 33, 33,
                     33, 33, 36, 36, 36, 36, 38, 38, 38, 40, 40, 42]


def _lists(seed):
    rng = np.random.default_rng(seed)
    return [int(v) for v in rng.integers(4, 60, 16)], [int(v) for v in rng.integers(4, 60, 64)]


@pytest.mark.parametrize("case", ["sps_explicit", "sps_absent_lists", "sps_default_flag", "pps_only", "sps_wins", "wraparound"])
def test_scaling_matrix_selection_follows_the_reference(recon_lib, case):
    """SliceHeader::scaling_lists (slice/header.rs:317-332) + ScalingLists::new (atom/avcc/sps.rs:207-248): the SPS matrix
    if there is one, else the PPS matrix, else flat; a list that is absent or flags useDefault becomes the Default table."""
    l4, l8 = _lists(1)
    m4, m8 = _lists(2)
    pp = PicParams.make(3, 2)
    b = synth.generate(pp, 1, 31, zero_residual=True)
    kw, want4, want8, src = {}, None, None, 0
    if case == "sps_explicit":
        kw = dict(sps_matrix={0: l4, 6: l8, 3: m4})
        want4, want8, src = l4, l8, 1
    elif case == "sps_absent_lists":   # a matrix with only an inter list: list 0 and list 6 fall to the Default tables
        kw = dict(sps_matrix={4: m4})
        want4, want8, src = DEFAULT_4X4_INTRA, DEFAULT_8X8_INTRA, 1
    elif case == "sps_default_flag":
        kw = dict(sps_matrix={0: "default", 6: "default"})
        want4, want8, src = DEFAULT_4X4_INTRA, DEFAULT_8X8_INTRA, 1
    elif case == "pps_only":
        kw = dict(pps_matrix={0: l4, 6: l8})
        want4, want8, src = l4, l8, 2
    elif case == "sps_wins":
        kw = dict(sps_matrix={0: l4, 6: l8}, pps_matrix={0: m4, 6: m8})
        want4, want8, src = l4, l8, 1
    else:                              # delta_scale wraps modulo 256 (7.3.2.1.1.1)
        w4 = [250, 3, 250, 3] * 4
        kw = dict(sps_matrix={0: w4, 6: l8})
        want4, want8, src = w4, l8, 1
    data = stream.encode_stream(b, **kw)
    got, n = host.scan(data)
    assert n == 1 and list(got.scaling_list4x4) == want4 and list(got.scaling_list8x8) == want8
    assert host.slice_info(data, 0).scaling_matrix_source == src
    assert_same_syntax(host.parse(data), b)


def _matrix_stream(seed):
    l4, l8 = _lists(seed)
    pp = PicParams.make(6, 4, 0, 0, l4, l8)
    # stress_pct = 0: the stress macroblocks' levels, quantised against small weights, leave the 16-bit coefficient range
    # a conforming stream keeps to (8.5.12.1 note), and libavcodec stores coefficients in int16
    b = synth.generate(pp, 2, 41 + seed, standard_only=True, stress_pct=0)
    return pp, b, stream.encode_stream(b, sps_matrix={0: l4, 6: l8}), l4, l8


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_scaling_matrix_stream_reconstructs_like_libavcodec(recon_lib, seed):
    """A stream whose SPS carries Intra-Y lists only (all three macroblock classes): bytes -> parse (lists from the SPS)
    -> oracle must give a conformant decoder's luma."""
    from avc import decode
    if not decode.available():
        pytest.skip("cv2 with the FFmpeg backend is not available")
    pp, b, data, l4, l8 = _matrix_stream(seed)
    got_pp, n = host.scan(data)
    assert list(got_pp.scaling_list4x4) == l4 and list(got_pp.scaling_list8x8) == l8
    parsed = host.parse(data)
    parsed.pp = got_pp
    assert np.array_equal(luma_of(oracle.reconstruct(parsed), pp), decode.decode_luma(data, 2, 96, 64))


@pytest.mark.gpu
def test_scaling_matrix_stream_bytes_to_cuda_equals_libavcodec(recon_lib):
    from avc import decode
    pp, b, data, l4, l8 = _matrix_stream(8)
    parsed = host.parse(data)
    parsed.pp = host.scan(data)[0]
    ctx = recon.ReconContext(0)
    got = ctx.reconstruct(parsed)
    ctx.close()
    assert np.array_equal(got, oracle.reconstruct(parsed))
    if decode.available():
        assert np.array_equal(luma_of(got, pp), decode.decode_luma(data, 2, 96, 64))


def test_parameter_sets_are_activated_by_id_per_picture(recon_lib):
    """Two PPS with different pic_init_qp; pictures alternate between them: every slice's QP comes from the PPS it names
    (7.4.1.2.1), also for a set that arrives after other pictures."""
    pp = PicParams.make(4, 3, 1, -2)
    b = synth.generate(pp, 4, 51, qp_base=30, qp_jitter=3)
    data = stream.encode_stream(b, extra_pps=(35, lambda f: f % 2 == 1))
    assert_same_syntax(host.parse(data), b)
    assert [host.slice_info(data, f).pic_parameter_set_id for f in range(4)] == [0, 1, 0, 1]
    for f in range(4):
        assert host.slice_info(data, f).slice_qp == int(b.qp[f * pp.n_mb])
        got = host.picture_params(data, f)
        assert (got.chroma_qp_index_offset, got.second_chroma_qp_index_offset) == (1, -2)
    # a slice that names a PPS the stream never sent
    bad = stream.nal_unit(3, 7, stream.sps_rbsp(4, 3)) + stream.encode_picture(b, 0, stream.canonicalise(b), pps_id=3)
    with pytest.raises(recon.ReconError) as e:
        host.scan(bad)
    assert e.value.code == recon.ERR_ARG


def test_slice_info_reports_the_deblocking_request(recon_lib):
    pp = PicParams.make(3, 2)
    b = synth.generate(pp, 1, 61)
    si = host.slice_info(stream.encode_stream(b), 0)
    assert si.disable_deblocking_filter_idc == 1
    si = host.slice_info(stream.encode_stream(b, deblock=(2, -3)), 0)
    assert (si.disable_deblocking_filter_idc, si.slice_alpha_c0_offset_div2, si.slice_beta_offset_div2) == (0, 2, -3)


def test_host_parser_survives_corrupted_input_under_sanitizers(tmp_path):
    """tests/native/cabac_fuzz.cpp: seeded byte flips, truncations, insertions and deletions of a valid Annex-B stream and
    of two MP4 files through every entry point of include/dryv_cabac_host.h, built with AddressSanitizer + UBSan — the host
    reads untrusted bytes, so any status is fine and any out-of-bounds access, overflow or crash is not."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "cabac_fuzz")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe,
                           os.path.join(root, "tests/native/cabac_fuzz.cpp"), os.path.join(root, "dryv_b200/csrc/cabac_host.cpp"),
                           os.path.join(root, "dryv_b200/csrc/levels_pack.cpp"), "-lpthread"])
    pp = PicParams.make(8, 5, 1, -2)
    b = synth.generate(pp, 3, 4242, qp_base=27)
    annexb = tmp_path / "s.264"
    annexb.write_bytes(stream.encode_stream(b))
    pin = os.path.join(root, "tests/golden/pin")
    for path, seed in ((str(annexb), 11), (os.path.join(pin, "matrices_96x64.mp4"), 12), (os.path.join(pin, "i8x8_column0_96x64.mp4"), 13)):
        out = subprocess.run([exe, path, str(seed), "250"], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and out.stdout.startswith("ok:"), (path, out.stdout[-2000:], out.stderr[-4000:])
