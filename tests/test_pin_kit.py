"""The dryv pinning kit (tests/golden/pin/, tools/make_pin_kit.py, tools/pin_against_dryv.sh): the committed MP4 files,
their predicted ./temp/yuv_frame digests, the oracle and (with -m gpu) the CUDA path must agree, libavformat must open
every file, and the deviation each file is meant to exercise must really fire in it."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from avc import decode
from dryv_b200 import host, recon
from oracle import spec_model

KIT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pin")


def entries():
    out = []
    for line in open(os.path.join(KIT, "SHA256SUMS")):
        if line.strip() and not line.startswith("#"):
            h, name = line.split()[:2]
            out.append((name, h))
    return out


@pytest.mark.parametrize("name,digest", entries())
def test_oracle_reproduces_the_committed_digest(recon_lib, name, digest):
    data = open(os.path.join(KIT, name), "rb").read()
    b = host.parse(data)
    assert b.n_frames == 1
    frame = oracle.reconstruct(b)[0]
    assert hashlib.sha256(frame.tobytes()).hexdigest() == digest
    if decode.available():   # a real MP4: libavformat demuxes it and libavcodec decodes a picture of the coded size
        w, h = 16 * b.pp.pic_width_in_mbs, 16 * b.pp.pic_height_in_mbs
        luma = decode.decode_luma(data, 1, w, h)
        assert luma.shape == (1, h, w)


def test_the_kit_is_reproducible(recon_lib, tmp_path, monkeypatch):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_pin_kit", os.path.join(os.path.dirname(KIT), "..", "..", "tools", "make_pin_kit.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "OUT", str(tmp_path))
    mod.main()
    for name, _ in entries():
        assert open(os.path.join(KIT, name), "rb").read() == open(tmp_path / name, "rb").read(), name
    assert open(os.path.join(KIT, "SHA256SUMS")).read() == open(tmp_path / "SHA256SUMS").read()


def test_q2_fires_in_the_intra8x8_file(recon_lib):
    b = host.parse(open(os.path.join(KIT, "i8x8_column0_96x64.mp4"), "rb").read())
    pp = b.pp
    col0_i8 = [(a // pp.pic_width_in_mbs) for a in range(pp.n_mb)
               if a % pp.pic_width_in_mbs == 0 and b.mb_type[a] == 0 and b.transform_size_8x8_flag[a]]
    assert col0_i8, "no Intra8x8 macroblock in column 0"
    ours, std = oracle.reconstruct(b)[0], spec_model.reconstruct(b, quirks=False)[0]
    W = 16 * pp.pic_width_in_mbs
    diff = np.nonzero(ours[:pp.n_mb * 256].reshape(-1, W) != std[:pp.n_mb * 256].reshape(-1, W))
    assert diff[0].size > 0, "the file does not exercise Q2"
    assert np.array_equal(ours, spec_model.reconstruct(b, quirks=True)[0])


def test_q3_fires_in_the_chroma_file(recon_lib):
    b = host.parse(open(os.path.join(KIT, "chroma_zero_96x64.mp4"), "rb").read())
    ours, std = oracle.reconstruct(b)[0], spec_model.reconstruct(b, quirks=False)[0]
    n = b.pp.n_mb * 256
    assert (ours[n:] == 0).any()
    # Q2 may fire in luma as well (Intra8x8 in column 0); what this file adds is a chroma difference
    assert not np.array_equal(ours[n:], std[n:]), "the file does not exercise Q3"
    assert np.array_equal(ours, spec_model.reconstruct(b, quirks=True)[0])


@pytest.mark.gpu
@pytest.mark.parametrize("name,digest", entries())
def test_cuda_path_reproduces_the_committed_digest(gpu_ctx, name, digest):
    data = open(os.path.join(KIT, name), "rb").read()
    b, levels = host.parse_compact(data)
    got = gpu_ctx.reconstruct_compact(b, levels)
    assert hashlib.sha256(got[0].tobytes()).hexdigest() == digest
    assert hashlib.sha256(gpu_ctx.reconstruct(host.parse(data))[0].tobytes()).hexdigest() == digest
