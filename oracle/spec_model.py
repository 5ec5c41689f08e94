"""spec_model.py — an independent restatement of the H.264 intra reconstruction process, written from the
text of ITU-T H.264 (8.3.1-8.3.4 intra prediction, 8.5.6/8.5.9-8.5.13 scaling and transforms, 8.5.14
picture construction) in matrix / whole-block numpy form. TEST INFRASTRUCTURE ONLY.

Purpose: catch transcription errors in oracle/dryv_oracle.c. The C oracle follows the control flow of the
Rust reference line by line; this model is organised differently on purpose (matrix products for the
transforms, closed-form block predictors, explicit availability booleans instead of -1 sentinels), so the
two only agree if both implement the same arithmetic.

The reference deviates from the standard in a few places. Each deviation is a
switch here (`quirks=True` reproduces dryv, `quirks=False` is the plain standard):
  Q1  chroma is dequantised with the luma (Intra-Y) 4x4 list
  Q2  Intra8x8 filtered p'[0,-1] uses the raw corner sentinel when the corner is unavailable
  Q3  chroma DC prediction treats some zero-valued neighbours as unavailable
"""
from __future__ import annotations

import numpy as np

ZZ4 = [(0, 0), (0, 1), (1, 0), (2, 0), (1, 1), (0, 2), (0, 3), (1, 2),
       (2, 1), (3, 0), (3, 1), (2, 2), (1, 3), (2, 3), (3, 2), (3, 3)]


def zigzag8():
    """8x8 frame zig-zag: walk anti-diagonals, alternating direction (H.264 Figure 6-?/Table 8-3 order)."""
    out = []
    for s in range(15):
        cells = [(i, s - i) for i in range(8) if 0 <= s - i < 8]
        out += cells if s % 2 == 1 else cells[::-1]
    return out


ZZ8 = zigzag8()

NORM4 = np.array([[10, 16, 13], [11, 18, 14], [13, 20, 16], [14, 23, 18], [16, 25, 20], [18, 29, 23]])
NORM8 = np.array([[20, 18, 32, 19, 25, 24], [22, 19, 35, 21, 28, 26], [26, 23, 42, 24, 33, 31],
                  [28, 25, 45, 26, 35, 33], [32, 28, 51, 30, 40, 38], [36, 32, 58, 34, 46, 43]])
QPC = list(range(30)) + [29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39]

BLK4_XY = [(0, 0), (4, 0), (0, 4), (4, 4), (8, 0), (12, 0), (8, 4), (12, 4),
           (0, 8), (4, 8), (0, 12), (4, 12), (8, 8), (12, 8), (8, 12), (12, 12)]


def level_scale4(weights_zz, m):
    w = np.zeros((4, 4), np.int64)
    for k, (i, j) in enumerate(ZZ4):
        w[i, j] = weights_zz[k]
    ls = np.zeros((4, 4), np.int64)
    for i in range(4):
        for j in range(4):
            cls = 0 if (i % 2 == 0 and j % 2 == 0) else (1 if (i % 2 == 1 and j % 2 == 1) else 2)
            ls[i, j] = w[i, j] * NORM4[m, cls]
    return ls


def level_scale8(weights_zz, m):
    w = np.zeros((8, 8), np.int64)
    for k, (i, j) in enumerate(ZZ8):
        w[i, j] = weights_zz[k]
    ls = np.zeros((8, 8), np.int64)
    for i in range(8):
        for j in range(8):
            if i % 4 == 0 and j % 4 == 0:
                cls = 0
            elif i % 2 == 1 and j % 2 == 1:
                cls = 1
            elif i % 4 == 2 and j % 4 == 2:
                cls = 2
            elif (i % 4 == 0 and j % 2 == 1) or (i % 2 == 1 and j % 4 == 0):
                cls = 3
            elif (i % 4 == 0 and j % 4 == 2) or (i % 4 == 2 and j % 4 == 0):
                cls = 4
            else:
                cls = 5
            ls[i, j] = w[i, j] * NORM8[m, cls]
    return ls


def idct4_1d(v):
    """8.5.12.2 one-dimensional 4-point transform applied along the last axis."""
    d0, d1, d2, d3 = v[..., 0], v[..., 1], v[..., 2], v[..., 3]
    e0, e1, e2, e3 = d0 + d2, d0 - d2, (d1 >> 1) - d3, d1 + (d3 >> 1)
    return np.stack([e0 + e3, e1 + e2, e1 - e2, e0 - e3], axis=-1)


def residual4x4(c, qp, ls, dc_given):
    """c: 4x4 coefficient matrix (with c[0,0] already the final DC when dc_given). Returns r (8.5.12)."""
    c = c.astype(np.int64)
    if qp >= 24:
        d = (c * ls) << (qp // 6 - 4)
    else:
        d = (c * ls + (1 << (3 - qp // 6))) >> (4 - qp // 6)
    if dc_given:
        d[0, 0] = c[0, 0]
    f = idct4_1d(d)          # rows
    h = idct4_1d(f.T).T      # columns
    return (h + 32) >> 6


def idct8_1d(v):
    a = [v[..., k] for k in range(8)]
    e0 = a[0] + a[4]
    e1 = -a[3] + a[5] - a[7] - (a[7] >> 1)
    e2 = a[0] - a[4]
    e3 = a[1] + a[7] - a[3] - (a[3] >> 1)
    e4 = (a[2] >> 1) - a[6]
    e5 = -a[1] + a[7] + a[5] + (a[5] >> 1)
    e6 = a[2] + (a[6] >> 1)
    e7 = a[3] + a[5] + a[1] + (a[1] >> 1)
    f0, f1, f2, f3 = e0 + e6, e1 + (e7 >> 2), e2 + e4, e3 + (e5 >> 2)
    f4, f5, f6, f7 = e2 - e4, (e3 >> 2) - e5, e0 - e6, e7 - (e1 >> 2)
    return np.stack([f0 + f7, f2 + f5, f4 + f3, f6 + f1, f6 - f1, f4 - f3, f2 - f5, f0 - f7], axis=-1)


def residual8x8(c, qp, ls8):
    c = c.astype(np.int64)
    if qp >= 36:
        d = (c * ls8) << (qp // 6 - 6)
    else:
        d = (c * ls8 + (1 << (5 - qp // 6))) >> (6 - qp // 6)
    g = idct8_1d(d)
    m = idct8_1d(g.T).T
    return (m + 32) >> 6


H4 = np.array([[1, 1, 1, 1], [1, 1, -1, -1], [1, -1, -1, 1], [1, -1, 1, -1]], np.int64)
H2 = np.array([[1, 1], [1, -1]], np.int64)


def luma_dc16(c, qp, ls00):
    f = H4 @ c.astype(np.int64) @ H4
    if qp >= 36:
        return (f * ls00) << (qp // 6 - 6)
    return (f * ls00 + (1 << (5 - qp // 6))) >> (6 - qp // 6)


def chroma_dc(c, qpc, ls00):
    f = H2 @ c.astype(np.int64) @ H2
    return ((f * ls00) << (qpc // 6)) >> 5


# --------------------------------------------------------------------------------------------------
# prediction (p is a dict-like accessor over already-reconstructed samples; None = unavailable)
# --------------------------------------------------------------------------------------------------
def pred_nxn(n, mode, T, L, TL):
    """Intra4x4 (n=4) / Intra8x8 (n=8) sample prediction from top T[0..2n-1], left L[0..n-1], corner TL.
    For n=8 the inputs are the filtered reference samples. Unavailable groups are None."""
    out = np.zeros((n, n), np.int64)  # [y, x]
    zmax = 2 * n - 2

    def t(k):
        return TL if k < 0 else T[k]

    def l(k):
        return TL if k < 0 else L[k]

    if mode == 0:
        for y in range(n):
            for x in range(n):
                out[y, x] = T[x]
    elif mode == 1:
        for y in range(n):
            for x in range(n):
                out[y, x] = L[y]
    elif mode == 2:
        if T is not None and L is not None:
            v = (sum(T[:n]) + sum(L[:n]) + n) >> (2 if n == 4 else 3) >> 1
        elif L is not None:
            v = (sum(L[:n]) + n // 2) >> (2 if n == 4 else 3)
        elif T is not None:
            v = (sum(T[:n]) + n // 2) >> (2 if n == 4 else 3)
        else:
            v = 128
        out[:] = v
    elif mode == 3:
        for y in range(n):
            for x in range(n):
                if x == n - 1 and y == n - 1:
                    out[y, x] = (T[2 * n - 2] + 3 * T[2 * n - 1] + 2) >> 2
                else:
                    out[y, x] = (T[x + y] + 2 * T[x + y + 1] + T[x + y + 2] + 2) >> 2
    elif mode == 4:
        for y in range(n):
            for x in range(n):
                if x > y:
                    out[y, x] = (t(x - y - 2) + 2 * t(x - y - 1) + t(x - y) + 2) >> 2
                elif x < y:
                    out[y, x] = (l(y - x - 2) + 2 * l(y - x - 1) + l(y - x) + 2) >> 2
                else:
                    out[y, x] = (T[0] + 2 * TL + L[0] + 2) >> 2
    elif mode == 5:
        for y in range(n):
            for x in range(n):
                z = 2 * x - y
                if z >= 0 and z % 2 == 0:
                    out[y, x] = (t(x - (y >> 1) - 1) + t(x - (y >> 1)) + 1) >> 1
                elif z >= 0:
                    out[y, x] = (t(x - (y >> 1) - 2) + 2 * t(x - (y >> 1) - 1) + t(x - (y >> 1)) + 2) >> 2
                elif z == -1:
                    out[y, x] = (L[0] + 2 * TL + T[0] + 2) >> 2
                else:
                    out[y, x] = (l(y - 2 * x - 1) + 2 * l(y - 2 * x - 2) + l(y - 2 * x - 3) + 2) >> 2
    elif mode == 6:
        for y in range(n):
            for x in range(n):
                z = 2 * y - x
                if z >= 0 and z % 2 == 0:
                    out[y, x] = (l(y - (x >> 1) - 1) + l(y - (x >> 1)) + 1) >> 1
                elif z >= 0:
                    out[y, x] = (l(y - (x >> 1) - 2) + 2 * l(y - (x >> 1) - 1) + l(y - (x >> 1)) + 2) >> 2
                elif z == -1:
                    out[y, x] = (L[0] + 2 * TL + T[0] + 2) >> 2
                else:
                    out[y, x] = (t(x - 2 * y - 1) + 2 * t(x - 2 * y - 2) + t(x - 2 * y - 3) + 2) >> 2
    elif mode == 7:
        for y in range(n):
            for x in range(n):
                if y % 2 == 0:
                    out[y, x] = (T[x + (y >> 1)] + T[x + (y >> 1) + 1] + 1) >> 1
                else:
                    out[y, x] = (T[x + (y >> 1)] + 2 * T[x + (y >> 1) + 1] + T[x + (y >> 1) + 2] + 2) >> 2
    elif mode == 8:
        zl = 2 * n - 3
        for y in range(n):
            for x in range(n):
                z = x + 2 * y
                if z < zl and z % 2 == 0:
                    out[y, x] = (L[y + (x >> 1)] + L[y + (x >> 1) + 1] + 1) >> 1
                elif z < zl:
                    out[y, x] = (L[y + (x >> 1)] + 2 * L[y + (x >> 1) + 1] + L[y + (x >> 1) + 2] + 2) >> 2
                elif z == zl:
                    out[y, x] = (L[n - 2] + 3 * L[n - 1] + 2) >> 2
                else:
                    out[y, x] = L[n - 1]
    del zmax
    return out


def mode_needs(mode):
    """(top, left, corner) requirements of an NxN mode."""
    return {0: (1, 0, 0), 1: (0, 1, 0), 2: (0, 0, 0), 3: (1, 0, 0), 4: (1, 1, 1), 5: (1, 1, 1), 6: (1, 1, 1),
            7: (1, 0, 0), 8: (0, 1, 0)}[mode]


def filter8(T, L, TL, quirks):
    """8.3.2.2.1 reference sample filtering for Intra8x8 (T has 16 entries or is None)."""
    Tf = Lf = TLf = None
    if T is not None:
        Tf = [0] * 16
        if TL is not None:
            Tf[0] = (TL + 2 * T[0] + T[1] + 2) >> 2
        elif quirks:
            Tf[0] = (-1 + 2 * T[0] + T[1] + 2) >> 2  # Q2
        else:
            Tf[0] = (3 * T[0] + T[1] + 2) >> 2
        for x in range(1, 15):
            Tf[x] = (T[x - 1] + 2 * T[x] + T[x + 1] + 2) >> 2
        Tf[15] = (T[14] + 3 * T[15] + 2) >> 2
    if TL is not None:
        if T is None and L is None:
            TLf = TL
        elif T is None:
            TLf = (3 * TL + L[0] + 2) >> 2
        elif L is None:
            TLf = (3 * TL + T[0] + 2) >> 2
        else:
            TLf = (T[0] + 2 * TL + L[0] + 2) >> 2
    if L is not None:
        Lf = [0] * 8
        Lf[0] = ((TL + 2 * L[0] + L[1] + 2) >> 2) if TL is not None else ((3 * L[0] + L[1] + 2) >> 2)
        for y in range(1, 7):
            Lf[y] = (L[y - 1] + 2 * L[y] + L[y + 1] + 2) >> 2
        Lf[7] = (L[6] + 3 * L[7] + 2) >> 2
    return Tf, Lf, TLf


def pred16(mode, T, L, TL):
    out = np.zeros((16, 16), np.int64)
    if mode == 0:
        out[:] = np.array(T)[None, :]
    elif mode == 1:
        out[:] = np.array(L)[:, None]
    elif mode == 2:
        if T is not None and L is not None:
            v = (sum(T) + sum(L) + 16) >> 5
        elif L is not None:
            v = (sum(L) + 8) >> 4
        elif T is not None:
            v = (sum(T) + 8) >> 4
        else:
            v = 128
        out[:] = v
    else:
        def t(k):
            return TL if k < 0 else T[k]

        def l(k):
            return TL if k < 0 else L[k]
        Hh = sum((x + 1) * (t(8 + x) - t(6 - x)) for x in range(8))
        Vv = sum((y + 1) * (l(8 + y) - l(6 - y)) for y in range(8))
        a = 16 * (L[15] + T[15])
        b = (5 * Hh + 32) >> 6
        c = (5 * Vv + 32) >> 6
        for y in range(16):
            for x in range(16):
                out[y, x] = min(max((a + b * (x - 7) + c * (y - 7) + 16) >> 5, 0), 255)
    return out


def pred_chroma(mode, T, L, TL, quirks):
    out = np.zeros((8, 8), np.int64)
    if mode == 0:
        for blk in range(4):
            xo, yo = (blk % 2) * 4, (blk // 2) * 4
            t = T[xo:xo + 4] if T is not None else None
            l = L[yo:yo + 4] if L is not None else None
            if quirks:
                # Q3: the reference's "> 0" tests (trans_chroma.rs:209-216, 242, 257, 268)
                t_gt = t is not None and all(v > 0 for v in t)
                l_gt = l is not None and all(v > 0 for v in l)
                t3_gt = t is not None and t[3] > 0
                l3_gt = l is not None and l[3] > 0
            else:
                t_gt = t3_gt = t is not None
                l_gt = l3_gt = l is not None
            if (xo, yo) in ((0, 0), (4, 4)):
                if t is not None and l is not None:
                    v = (sum(t) + sum(l) + 4) >> 3
                elif l is not None:
                    v = (sum(l) + 2) >> 2
                elif t_gt and not l_gt:
                    v = (sum(t) + 2) >> 2
                else:
                    v = 128
            elif xo > 0 and yo == 0:
                if t is not None:
                    v = (sum(t) + 2) >> 2
                elif l3_gt:
                    v = (sum(l) + 2) >> 2
                else:
                    v = 128
            else:
                if l3_gt:
                    v = (sum(l) + 2) >> 2
                elif t3_gt:
                    v = (sum(t) + 2) >> 2
                else:
                    v = 128
            out[yo:yo + 4, xo:xo + 4] = v
    elif mode == 1:
        out[:] = np.array(L)[:, None]
    elif mode == 2:
        out[:] = np.array(T)[None, :]
    else:
        def t(k):
            return TL if k < 0 else T[k]

        def l(k):
            return TL if k < 0 else L[k]
        Hh = sum((x + 1) * (t(4 + x) - t(2 - x)) for x in range(4))
        Vv = sum((y + 1) * (l(4 + y) - l(2 - y)) for y in range(4))
        a = 16 * (L[7] + T[7])
        b = (34 * Hh + 32) >> 6
        c = (34 * Vv + 32) >> 6
        for y in range(8):
            for x in range(8):
                out[y, x] = min(max((a + b * (x - 3) + c * (y - 3) + 16) >> 5, 0), 255)
    return out


# --------------------------------------------------------------------------------------------------
# whole-picture driver
# --------------------------------------------------------------------------------------------------
def reconstruct(batch, quirks=True):
    """Reconstruct every picture of a SyntaxBatch; returns u8 [n_frames, frame_bytes]. Legal streams only
    (modes whose neighbours are missing are not modelled: the generator never emits them)."""
    pp = batch.pp
    W, H = int(pp.pic_width_in_mbs), int(pp.pic_height_in_mbs)
    n_mb = W * H
    l4 = list(pp.scaling_list4x4)
    l8 = list(pp.scaling_list8x8)
    LS4 = [level_scale4(l4, m) for m in range(6)]
    LS8 = [level_scale8(l8, m) for m in range(6)]
    if not quirks and (any(v != 16 for v in l4) or any(v != 16 for v in l8)):
        # Q1 only matters with non-flat lists, and the plain standard would need the Cb/Cr lists,
        # which are not part of the dryv_pic_params contract
        raise NotImplementedError("quirks=False is only defined for flat scaling lists")
    out = np.zeros((batch.n_frames, n_mb * 384), np.uint8)
    for f in range(batch.n_frames):
        Y = np.zeros((H * 16, W * 16), np.int64)
        C = [np.zeros((H * 8, W * 8), np.int64) for _ in range(2)]
        kind = np.zeros(n_mb, np.int64)        # 0 I4x4, 1 I8x8, 2 I16x16
        m4 = np.full((n_mb, 4, 4), 2, np.int64)  # resolved modes on the 4x4 grid [gy, gx]
        for a in range(n_mb):
            i = f * n_mb + a
            mx, my = a % W, a // W
            A, B = mx > 0, my > 0
            Cc, D = my > 0 and mx < W - 1, mx > 0 and my > 0
            code, t8 = int(batch.mb_type[i]), int(batch.transform_size_8x8_flag[i])
            qp = int(batch.qp[i])
            cf = batch.coeff[i].astype(np.int64)
            ps = batch.pred_syntax[i]
            cls = 2 if code else (1 if t8 else 0)
            kind[a] = cls
            X0, Y0 = mx * 16, my * 16

            def luma(x, y):  # sample at MB-relative (x, y), None if unavailable
                if x >= 0 and y >= 0:
                    return int(Y[Y0 + y, X0 + x]) if x < 16 else None
                if y < 0 and x < 0:
                    ok = D
                elif y < 0:
                    ok = B if x < 16 else Cc
                else:
                    ok = A
                return int(Y[Y0 + y, X0 + x]) if ok else None

            def nb_mode(gx, gy):
                """mode of the 4x4-grid cell (gx, gy) relative to this MB, or None if that MB is unavailable"""
                if gx >= 0 and gy >= 0:
                    return int(m4[a, gy, gx])
                if gx < 0:
                    return int(m4[a - 1, gy, 3]) if A else None
                return int(m4[a - W, 3, gx]) if B else None

            def derive(gx, gy, syn):
                ma, mb = nb_mode(gx - 1, gy), nb_mode(gx, gy - 1)
                pred = 2 if (ma is None or mb is None) else min(ma, mb)
                prev, rem = (syn >> 3) & 1, syn & 7
                return pred if prev else (rem if rem < pred else rem + 1)

            if cls == 0:
                for b in range(16):
                    xo, yo = BLK4_XY[b]
                    gx, gy = xo // 4, yo // 4
                    mode = derive(gx, gy, int(ps[b]))
                    m4[a, gy, gx] = mode
                    T = [luma(xo + k, yo - 1) for k in range(8)]
                    if b in (3, 11):
                        T[4:] = [None] * 4
                    if T[0] is None:
                        T = None
                    elif T[4] is None:
                        T[4:] = [T[3]] * 4
                    L = [luma(xo - 1, yo + k) for k in range(4)]
                    L = None if L[0] is None else L
                    TL = luma(xo - 1, yo - 1)
                    c = np.zeros((4, 4), np.int64)
                    for k, (ii, jj) in enumerate(ZZ4):
                        c[ii, jj] = cf[b * 16 + k]
                    r = residual4x4(c, qp, LS4[qp % 6], False)
                    p = pred_nxn(4, mode, T, L, TL)
                    Y[Y0 + yo:Y0 + yo + 4, X0 + xo:X0 + xo + 4] = np.clip(p + r, 0, 255)
            elif cls == 1:
                for b in range(4):
                    xo, yo = (b % 2) * 8, (b // 2) * 8
                    gx, gy = xo // 4, yo // 4
                    mode = derive(gx, gy, int(ps[b]))
                    m4[a, gy:gy + 2, gx:gx + 2] = mode
                    T = [luma(xo + k, yo - 1) for k in range(16)]
                    if T[0] is None:
                        T = None
                    elif T[8] is None:
                        T[8:] = [T[7]] * 8
                    L = [luma(xo - 1, yo + k) for k in range(8)]
                    L = None if L[0] is None else L
                    TL = luma(xo - 1, yo - 1)
                    Tf, Lf, TLf = filter8(T, L, TL, quirks)
                    c = np.zeros((8, 8), np.int64)
                    for k, (ii, jj) in enumerate(ZZ8):
                        c[ii, jj] = cf[b * 64 + k]
                    r = residual8x8(c, qp, LS8[qp % 6])
                    p = pred_nxn(8, mode, Tf, Lf, TLf)
                    Y[Y0 + yo:Y0 + yo + 8, X0 + xo:X0 + xo + 8] = np.clip(p + r, 0, 255)
            else:
                m4[a] = 2
                c = np.zeros((4, 4), np.int64)
                for k, (ii, jj) in enumerate(ZZ4):
                    c[ii, jj] = cf[k * 16]
                dcY = luma_dc16(c, qp, int(LS4[qp % 6][0, 0]))
                R = np.zeros((16, 16), np.int64)
                for b in range(16):
                    xo, yo = BLK4_XY[b]
                    cc = np.zeros((4, 4), np.int64)
                    for k, (ii, jj) in enumerate(ZZ4):
                        cc[ii, jj] = cf[b * 16 + k]
                    cc[0, 0] = dcY[yo // 4, xo // 4]
                    R[yo:yo + 4, xo:xo + 4] = residual4x4(cc, qp, LS4[qp % 6], True)
                T = [luma(k, -1) for k in range(16)]
                T = None if T[0] is None else T
                L = [luma(-1, k) for k in range(16)]
                L = None if L[0] is None else L
                p = pred16((code - 1) % 4, T, L, luma(-1, -1))
                Y[Y0:Y0 + 16, X0:X0 + 16] = np.clip(p + R, 0, 255)

            cm = int(batch.intra_chroma_pred_mode[i])
            for pl in range(2):
                off = int(pp.chroma_qp_index_offset if pl == 0 else pp.second_chroma_qp_index_offset)
                qpc = QPC[min(max(qp + off, 0), 51)]
                ls = LS4[qpc % 6]  # Q1: the luma list
                base = 256 + pl * 64
                cdc = np.array([[cf[base + 0], cf[base + 16]], [cf[base + 32], cf[base + 48]]], np.int64)
                dcC = chroma_dc(cdc, qpc, int(ls[0, 0]))
                R = np.zeros((8, 8), np.int64)
                for b in range(4):
                    cc = np.zeros((4, 4), np.int64)
                    for k, (ii, jj) in enumerate(ZZ4):
                        cc[ii, jj] = cf[base + b * 16 + k]
                    cc[0, 0] = dcC[b // 2, b % 2]
                    R[(b // 2) * 4:(b // 2) * 4 + 4, (b % 2) * 4:(b % 2) * 4 + 4] = residual4x4(cc, qpc, ls, True)
                P = C[pl]
                cx, cy = mx * 8, my * 8
                T = [int(P[cy - 1, cx + k]) for k in range(8)] if B else None
                L = [int(P[cy + k, cx - 1]) for k in range(8)] if A else None
                TL = int(P[cy - 1, cx - 1]) if D else None
                p = pred_chroma(cm, T, L, TL, quirks)
                P[cy:cy + 8, cx:cx + 8] = np.clip(p + R, 0, 255)
        out[f] = np.concatenate([Y.reshape(-1), C[0].reshape(-1), C[1].reshape(-1)]).astype(np.uint8)
    return out
