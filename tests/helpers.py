"""Shared helpers for the test-suite."""
import glob
import os

import numpy as np

from dryv_b200.abi import FIELDS, PicParams, SyntaxBatch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    pp = PicParams.make(int(z["w_mbs"]), int(z["h_mbs"]), int(z["cb_off"]), int(z["cr_off"]),
                        z["list4x4"].tolist(), z["list8x8"].tolist())
    b = SyntaxBatch(pp, int(z["n_frames"]), *[np.ascontiguousarray(z[f]) for f in FIELDS])
    return b, z["expected"]


def first_difference(pp, ref, got):
    """Human-readable location of the first differing macroblock (for assertion messages)."""
    if np.array_equal(ref, got):
        return "identical"
    W, H = pp.pic_width_in_mbs, pp.pic_height_in_mbs
    n = pp.n_mb
    for f in range(ref.shape[0]):
        ry, gy = ref[f][:n * 256].reshape(H * 16, W * 16), got[f][:n * 256].reshape(H * 16, W * 16)
        d = (ry != gy).reshape(H, 16, W, 16).any(axis=(1, 3))
        if d.any():
            y, x = np.argwhere(d)[0]
            return f"picture {f}: luma of MB x={x} y={y} differs ({int((ry != gy).sum())} luma bytes in the picture)"
        if not np.array_equal(ref[f], got[f]):
            return f"picture {f}: chroma differs ({int((ref[f] != got[f]).sum())} bytes)"
    return "differs"
