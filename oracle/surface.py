"""TEST INFRASTRUCTURE, not product code: only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this.

CPU statement (numpy) of the output surface of SURVEY.md §8(f) next-3: the crop rectangle of a coded picture, as packed
I420 planes or as NV12. The reference writes the uncropped, macroblock-aligned planes (src/video/frame/mod.rs:48-70: Y, then
Cb, then Cr, row-major) and lists "Frame cropping" as an open roadmap item (README.md:13) while already parsing the SPS crop
fields (src/video/atom/avcc/sps.rs:252-267), so there is no reference output to match; what is restated is H.264 7.4.2.1.1
(the display rectangle is columns CropUnitX * frame_crop_left_offset .. PicWidthInSamplesL - CropUnitX * frame_crop_right_offset - 1,
rows likewise, CropUnitX = CropUnitY = 2 for 4:2:0 frame pictures) applied to that planar layout. Parity for the crop itself
is pinned against libavcodec's output size and pixels in tests/test_surface.py (parity of the pixels: oracle/oracle.py).
"""
import numpy as np

I420, NV12 = 0, 1


def sps_rectangle(w_mbs, h_mbs, crop):
    """(crop_left, crop_top, width, height) in luma samples for SPS offsets crop = (left, right, top, bottom) or None."""
    left, right, top, bottom = crop or (0, 0, 0, 0)
    return 2 * left, 2 * top, 16 * w_mbs - 2 * (left + right), 16 * h_mbs - 2 * (top + bottom)


def export(frame, w_mbs, h_mbs, crop_left, crop_top, width, height, fmt=I420):
    """frame: uint8[w_mbs*h_mbs*384], the coded picture (Y | Cb | Cr). Returns uint8[width*height*3/2]."""
    W, H = 16 * w_mbs, 16 * h_mbs
    assert crop_left % 2 == 0 and crop_top % 2 == 0 and width % 2 == 0 and height % 2 == 0
    assert 0 < width and crop_left + width <= W and 0 < height and crop_top + height <= H
    y = frame[:W * H].reshape(H, W)
    cb = frame[W * H:W * H * 5 // 4].reshape(H // 2, W // 2)
    cr = frame[W * H * 5 // 4:].reshape(H // 2, W // 2)
    ys = y[crop_top:crop_top + height, crop_left:crop_left + width]
    sl = (slice(crop_top // 2, (crop_top + height) // 2), slice(crop_left // 2, (crop_left + width) // 2))
    if fmt == I420:
        return np.concatenate([ys.ravel(), cb[sl].ravel(), cr[sl].ravel()])
    inter = np.stack([cb[sl], cr[sl]], axis=-1)   # rows of Cb, Cr pairs
    return np.concatenate([ys.ravel(), inter.ravel()])
