"""Writes tests/golden/avc/crosscheck.npz: a seeded mixed Intra4x4/8x8/16x16 batch (no Intra8x8 macroblock in column 0,
so the reference's one luma deviation from the standard, SURVEY quirk Q2, cannot fire), the H.264 CABAC stream
tests/avc/stream.py writes for it, and the luma planes LIBAVCODEC decoded from that stream in the build container
(cv2 4.13, FFmpeg backend, avcodec 62). The fixture lets the oracle and the CUDA path be checked against an independent
conformant decoder's output on machines without cv2.

    python tests/golden/avc/make_avc_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, TESTS)
sys.path.insert(0, os.path.dirname(TESTS))

from avc import decode, stream  # noqa: E402
from dryv_b200 import synth  # noqa: E402
from dryv_b200.abi import FIELDS, PicParams  # noqa: E402

pp = PicParams.make(9, 6, 2, -3)
b = synth.generate(pp, 3, 20261018, qp_base=24, qp_jitter=3, pct_i4x4=40, pct_i8x8=30, stress_pct=20, standard_only=True)
data = stream.encode_stream(b)          # canonicalises b in place (cbp bits of I_16x16 mb_types, inherited QPs)
luma = decode.decode_luma(data, 3, 144, 96)
np.savez_compressed(os.path.join(HERE, "crosscheck.npz"), w_mbs=9, h_mbs=6, cb_off=2, cr_off=-3, n_frames=3,
                    stream=np.frombuffer(data, np.uint8), libavcodec_luma=luma, **{f: getattr(b, f) for f in FIELDS})
print("stream bytes", len(data), "luma", luma.shape, "mean", float(luma.mean()))


# ---- one fixture per macroblock class (all three classes x chroma), luma and the BGR pictures (chroma through swscale) ----
# generator option standard_only + no reconstructed chroma sample equal to 0 in these seeds: the reference follows the H.264
# text on every sample of these pictures, so libavcodec's output is what dryv must produce (tests/test_libavcodec_crosscheck.py)
cases = {"i4x4": dict(pct_i4x4=100, pct_i8x8=0), "i8x8": dict(pct_i4x4=0, pct_i8x8=100), "i16x16": dict(pct_i4x4=0, pct_i8x8=0)}
out = {}
for name, kw in cases.items():
    pp = PicParams.make(7, 5, -2, 3)
    seed = 20261100
    while True:   # a seed whose pictures keep every chroma sample above 0 (quirk Q3 cannot fire)
        b = synth.generate(pp, 2, seed, qp_base=26, qp_jitter=3, stress_pct=0, standard_only=True, **kw)
        data = stream.encode_stream(b)
        import oracle
        fr = oracle.reconstruct(b)
        if (fr[:, pp.n_mb * 256:] > 0).all():
            break
        seed += 1
    luma = decode.decode_luma(data, 2, 112, 80)
    bgr = decode.decode_bgr(data, 2)
    assert np.array_equal(luma, fr[:, :pp.n_mb * 256].reshape(2, 80, 112)), name
    assert np.array_equal(bgr, decode.bgr_of_pictures(fr, 112, 80)), name
    out[name + "_stream"] = np.frombuffer(data, np.uint8)
    out[name + "_luma"] = luma
    out[name + "_bgr"] = bgr
    out[name + "_seed"] = seed
    for f in FIELDS:
        out[name + "_" + f] = getattr(b, f)
    print(name, "seed", seed, "stream bytes", len(data))
np.savez_compressed(os.path.join(HERE, "classes.npz"), w_mbs=7, h_mbs=5, cb_off=-2, cr_off=3, n_frames=2, **out)
