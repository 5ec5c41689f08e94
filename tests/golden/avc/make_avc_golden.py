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
