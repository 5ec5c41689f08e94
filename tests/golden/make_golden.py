"""Regenerates the golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

The reference (Rust) cannot be built in this image and ships no vectors of its own, so these fixtures pin the
CPU oracle's output (oracle/dryv_oracle.c) at the time it agreed bit-for-bit with the independent spec model
(oracle/spec_model.py). Inputs come from the seeded generator; both inputs and expected output are stored so
the GPU tests need neither the oracle nor the generator."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import spec_model  # noqa: E402
from dryv_b200 import synth  # noqa: E402
from dryv_b200.abi import FIELDS, PicParams  # noqa: E402

CASES = {
    # name: (w_mbs, h_mbs, n_frames, seed, cb_off, cr_off, list4x4, list8x8, generator kwargs)
    "mixed_6x4": (6, 4, 2, 11, 0, 0, None, None, {}),
    "stress_5x5_offsets": (5, 5, 2, 12, 3, -4, None, None, dict(stress_pct=50, qp_base=30)),
    "lists_4x3": (4, 3, 1, 13, 1, 2, list(range(6, 38, 2)), [8 + (k * 3) % 40 for k in range(64)], dict(qp_base=20)),
    "i4x4_only_3x3_qp4": (3, 3, 1, 14, 0, 0, None, None, dict(pct_i4x4=100, pct_i8x8=0, qp_base=4)),
    "i8x8_only_3x3_qp44": (3, 3, 1, 15, 0, 0, None, None, dict(pct_i4x4=0, pct_i8x8=100, qp_base=44)),
    "i16x16_only_4x2": (4, 2, 1, 16, -2, 5, None, None, dict(pct_i4x4=0, pct_i8x8=0)),
    "single_mb": (1, 1, 3, 17, 0, 0, None, None, {}),
    "single_row_7x1": (7, 1, 1, 18, 0, 0, None, None, {}),
    "single_col_1x6": (1, 6, 1, 19, 0, 0, None, None, {}),
}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name, (w, h, n, seed, cb, cr, l4, l8, kw) in CASES.items():
        pp = PicParams.make(w, h, cb, cr, l4, l8)
        b = synth.generate(pp, n, seed, **kw)
        ref = oracle.reconstruct(b)
        sm = spec_model.reconstruct(b)
        assert np.array_equal(ref, sm), f"{name}: oracle and spec model disagree"
        arrays = {f: getattr(b, f) for f in FIELDS}
        np.savez_compressed(os.path.join(here, name + ".npz"), w_mbs=w, h_mbs=h, n_frames=n, cb_off=cb, cr_off=cr,
                            list4x4=np.array(l4 if l4 else [16] * 16, np.uint8),
                            list8x8=np.array(l8 if l8 else [16] * 64, np.uint8), expected=ref, **arrays)
        print(name, ref.shape, "ok")


if __name__ == "__main__":
    main()
