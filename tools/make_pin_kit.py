#!/usr/bin/env python
"""Writes the dryv pinning kit (tests/golden/pin/): small MP4 files made by the test tooling (tests/avc/stream.py +
tests/avc/mp4.py) and the SHA-256 the reference's ./temp/yuv_frame must have for each of them according to the oracle
(oracle/dryv_oracle.c). Anyone with a Rust toolchain runs tools/pin_against_dryv.sh to pin the oracle — and through it
the CUDA path — to dryv itself; this image has no cargo, so the digests below are the oracle's claim, not yet dryv's word.

Each file is one IDR picture (dryv decodes the first sample only, src/video/decoder.rs:88, and writes its frame to
./temp/yuv_frame, decoder.rs:141-143) chosen so that a deviation of the reference from the H.264 text fires:
  mixed_640x368.mp4        BASELINE configs[0] size, the bench's macroblock mix (Q4: illegal modes cannot occur in a legal stream)
  i8x8_column0_96x64.mp4   Intra8x8 macroblocks in macroblock column 0 (quirk Q2: the p[-1,-1] sentinel enters the filter)
  chroma_zero_96x64.mp4    stress residuals drive chroma samples to 0 next to DC-predicted blocks (quirk Q3: "> 0" tests)
  matrices_96x64.mp4       SPS scaling matrix with Intra-Y lists only (quirk Q6: absent lists -> Default tables; Q1: chroma
                           dequantised with the luma list)
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from avc import mp4, stream  # noqa: E402
from dryv_b200 import host, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "pin")


def cases():
    rng = np.random.default_rng(5)
    l4 = [int(v) for v in rng.integers(4, 60, 16)]
    l8 = [int(v) for v in rng.integers(4, 60, 64)]
    yield "mixed_640x368", PicParams.make(40, 23, 1, -1), dict(seed=360, qp_base=30), {}
    yield "i8x8_column0_96x64", PicParams.make(6, 4), dict(seed=11, pct_i4x4=20, pct_i8x8=70, qp_base=28), {}
    yield "chroma_zero_96x64", PicParams.make(6, 4, -3, 4), dict(seed=12, stress_pct=70, qp_base=24), {}
    yield "matrices_96x64", PicParams.make(6, 4, 0, 0, l4, l8), dict(seed=13, stress_pct=0, qp_base=27), \
        dict(sps_matrix={0: l4, 6: l8})


def main():
    os.makedirs(OUT, exist_ok=True)
    sums = []
    for name, pp, kw, enc in cases():
        b = synth.generate(pp, 1, kw.pop("seed"), **kw)
        data = mp4.mux(stream.encode_stream(b, **enc), 16 * pp.pic_width_in_mbs, 16 * pp.pic_height_in_mbs)
        parsed = host.parse(data)                       # what the CPU host reads back from the file ...
        frame = oracle.reconstruct(parsed)[0]           # ... and what the reference would write to ./temp/yuv_frame
        assert np.array_equal(frame, oracle.reconstruct(b)[0])
        with open(os.path.join(OUT, name + ".mp4"), "wb") as f:
            f.write(data)
        sums.append((hashlib.sha256(frame.tobytes()).hexdigest(), name + ".mp4", frame.nbytes, len(data)))
        print(f"{name}.mp4: {len(data)} bytes, yuv_frame {frame.nbytes} bytes")
    with open(os.path.join(OUT, "SHA256SUMS"), "w") as f:
        f.write("# sha256 of the ./temp/yuv_frame dryv must write for each file (oracle's claim; tools/make_pin_kit.py)\n")
        for h, n, fb, _ in sums:
            f.write(f"{h}  {n}  # {fb} bytes\n")


if __name__ == "__main__":
    main()
