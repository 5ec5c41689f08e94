"""TEST INFRASTRUCTURE, not product code: groundwork for SURVEY.md §8(f) next-4 (no kernel exists for it yet).

The H.264 in-loop deblocking filter (8.7) for the pictures this path produces: intra macroblocks only (bS = 4 on macroblock
edges, 3 inside), frame pictures, 4:2:0, 8 bit, one slice, disable_deblocking_filter_idc = 0. The reference has no
deblocking filter (README.md:15 lists it as open; its slice header parses the fields, src/video/slice/header.rs:609-640),
so it must stay OFF for dryv parity; this restates the standard's text and is pinned to libavcodec's output on streams that
enable the filter (tests/test_deblock_oracle.py: luma directly, chroma through the BGR pictures of tests/avc/decode.py).
Pure Python loops over macroblocks: small pictures only.
"""
import numpy as np

ALPHA = [0] * 16 + [4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28, 32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101,
                    113, 127, 144, 162, 182, 203, 226, 255, 255]
BETA = [0] * 16 + [2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15,
                   16, 16, 17, 17, 18, 18]
TC0_BS3 = [0] * 17 + [1] * 10 + [2] * 4 + [3] * 3 + [4] * 3 + [5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 23, 25]
assert len(ALPHA) == 52 and len(BETA) == 52 and len(TC0_BS3) == 52
# Table 8-15: QPc as a function of qPI
QPC = list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]


def _filter_edge(px, bs, qp_av, off_a, off_b, chroma):
    """px: int array [n, 8] = p3 p2 p1 p0 q0 q1 q2 q3 per line (chroma uses p1 p0 q0 q1 only). Returns the filtered copy."""
    ia = min(max(qp_av + off_a, 0), 51)
    ib = min(max(qp_av + off_b, 0), 51)
    alpha, beta = ALPHA[ia], BETA[ib]
    p3, p2, p1, p0, q0, q1, q2, q3 = (px[:, i].astype(np.int64) for i in range(8))
    on = (np.abs(p0 - q0) < alpha) & (np.abs(p1 - p0) < beta) & (np.abs(q1 - q0) < beta)
    out = px.astype(np.int64).copy()
    ap, aq = np.abs(p2 - p0), np.abs(q2 - q0)
    if bs < 4:
        tc0 = TC0_BS3[ia]
        tc = tc0 + 1 if chroma else tc0 + (ap < beta) + (aq < beta)
        delta = np.clip((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc)
        np0 = np.clip(p0 + delta, 0, 255)
        nq0 = np.clip(q0 - delta, 0, 255)
        np1, nq1 = p1.copy(), q1.copy()
        if not chroma:
            np1 = np.where(ap < beta, p1 + np.clip((p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1, -tc0, tc0), p1)
            nq1 = np.where(aq < beta, q1 + np.clip((q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1, -tc0, tc0), q1)
        out[:, 2] = np.where(on, np1, p1)
        out[:, 3] = np.where(on, np0, p0)
        out[:, 4] = np.where(on, nq0, q0)
        out[:, 5] = np.where(on, nq1, q1)
    else:
        small = np.abs(p0 - q0) < ((alpha >> 2) + 2)
        sp = (ap < beta) & small & (not chroma)
        sq = (aq < beta) & small & (not chroma)
        np0 = np.where(sp, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3, (2 * p1 + p0 + q1 + 2) >> 2)
        np1 = np.where(sp, (p2 + p1 + p0 + q0 + 2) >> 2, p1)
        np2 = np.where(sp, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3, p2)
        nq0 = np.where(sq, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3, (2 * q1 + q0 + p1 + 2) >> 2)
        nq1 = np.where(sq, (p0 + q0 + q1 + q2 + 2) >> 2, q1)
        nq2 = np.where(sq, (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3, q2)
        for i, v, old in ((1, np2, p2), (2, np1, p1), (3, np0, p0), (4, nq0, q0), (5, nq1, q1), (6, nq2, q2)):
            out[:, i] = np.where(on, v, old)
    return out


def _edge(plane, x0, y0, vertical, length, bs, qp_av, off_a, off_b, chroma):
    """Filters one edge in place: the edge lies left of column x0 (vertical) or above row y0 (horizontal)."""
    reach = 2 if chroma else 4
    if vertical:
        blk = plane[y0:y0 + length, x0 - reach:x0 + reach]
    else:
        blk = plane[y0 - reach:y0 + reach, x0:x0 + length].T
    px = np.zeros((length, 8), np.int64)
    px[:, 4 - reach:4 + reach] = blk
    res = _filter_edge(px, bs, qp_av, off_a, off_b, chroma)[:, 4 - reach:4 + reach]
    if vertical:
        plane[y0:y0 + length, x0 - reach:x0 + reach] = res
    else:
        plane[y0 - reach:y0 + reach, x0:x0 + length] = res.T


def deblock(frame, w_mbs, h_mbs, qp, t8x8, cb_off=0, cr_off=0, alpha_div2=0, beta_div2=0):
    """frame: uint8[w_mbs*h_mbs*384] reconstructed picture (Y | Cb | Cr); qp, t8x8: per-macroblock arrays (raster order).
    Returns the filtered picture in the same layout. Macroblocks in raster order, vertical edges left to right, then
    horizontal edges top to bottom, luma then each chroma plane (8.7)."""
    W, H = 16 * w_mbs, 16 * h_mbs
    y = frame[:W * H].reshape(H, W).astype(np.int64)
    cb = frame[W * H:W * H * 5 // 4].reshape(H // 2, W // 2).astype(np.int64)
    cr = frame[W * H * 5 // 4:].reshape(H // 2, W // 2).astype(np.int64)
    off_a, off_b = 2 * alpha_div2, 2 * beta_div2
    qp = np.asarray(qp, np.int64).reshape(h_mbs, w_mbs)
    t8 = np.asarray(t8x8).reshape(h_mbs, w_mbs)

    def qpc(q, off):
        return QPC[min(max(int(q) + off, 0), 51)]

    for my in range(h_mbs):
        for mx in range(w_mbs):
            q = int(qp[my, mx])
            step = 8 if t8[my, mx] else 4
            for vertical in (True, False):
                nb = (int(qp[my, mx - 1]) if mx > 0 else None) if vertical else (int(qp[my - 1, mx]) if my > 0 else None)
                for e in range(0, 16, step):
                    if e == 0 and nb is None:
                        continue
                    qa = (q + nb + 1) >> 1 if e == 0 else q
                    x0, y0 = (16 * mx + e, 16 * my) if vertical else (16 * mx, 16 * my + e)
                    _edge(y, x0, y0, vertical, 16, 4 if e == 0 else 3, qa, off_a, off_b, False)
                for plane, off in ((cb, cb_off), (cr, cr_off)):
                    for e in (0, 4):
                        if e == 0 and nb is None:
                            continue
                        qa = (qpc(q, off) + qpc(nb, off) + 1) >> 1 if e == 0 else qpc(q, off)
                        x0, y0 = (8 * mx + e, 8 * my) if vertical else (8 * mx, 8 * my + e)
                        _edge(plane, x0, y0, vertical, 8, 4 if e == 0 else 3, qa, off_a, off_b, True)
    return np.concatenate([y.ravel(), cb.ravel(), cr.ravel()]).astype(np.uint8)
