"""Minimal H.264 High-profile CABAC I-slice WRITER: dryv_mb_soa syntax buffers -> Annex-B byte stream.

TEST TOOLING ONLY. It exists so that the CPU oracle can be cross-checked against an
independent, conformant decoder (libavcodec through cv2, tests/test_libavcodec_crosscheck.py): the reference ships no
test vectors and cannot be built here, and no H.264 encoder or sample stream exists in the image. Written from the
text of ITU-T H.264 (7.3 syntax, 9.3 CABAC); the constant tables of 9.3 are loaded from dryv_b200/csrc/cabac_tables.json, the product's own copy
(see tools/make_cabac_tables.py). The inverse of what the reference parses in src/video/cabac/mod.rs:89-675,
src/video/slice/header.rs:145-315, src/video/atom/avcc/{sps,pps}.rs.

One IDR picture per access unit, one slice per picture, 8-bit 4:2:0, frame macroblocks, flat scaling lists,
deblocking disabled (disable_deblocking_filter_idc = 1), transform_8x8_mode_flag = 1, I_NxN and I_16x16 only.
"""
from __future__ import annotations

import json
import os

import numpy as np

_T = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "dryv_b200", "csrc", "cabac_tables.json")))
CTX_INIT_I = _T["ctx_init_i"]
RANGE_LPS = _T["range_tab_lps"]
TRANS_LPS = _T["trans_idx_lps"]
TRANS_MPS = _T["trans_idx_mps"]
SIG8 = _T["sig8x8_frame"]
LAST8 = _T["last8x8"]

# ctxIdx bases (Table 9-34), frame coded blocks, ctxBlockCat 0..5
CBF_BASE = [85, 89, 93, 97, 101]
SIG_BASE = [105, 120, 134, 149, 152, 402]
LAST_BASE = [166, 181, 195, 210, 213, 417]
ABS_BASE = [227, 237, 247, 257, 266, 426]

BLK4_XY = [(0, 0), (4, 0), (0, 4), (4, 4), (8, 0), (12, 0), (8, 4), (12, 4),
           (0, 8), (4, 8), (0, 12), (4, 12), (8, 8), (12, 8), (8, 12), (12, 12)]


def blk4_of(x, y):
    return 8 * (y // 8) + 4 * (x // 8) + 2 * ((y % 8) // 4) + ((x % 8) // 4)


class BitWriter:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        for i in range(n - 1, -1, -1):
            self.bits.append((v >> i) & 1)

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def trailing(self):
        self.bits.append(1)
        while len(self.bits) % 8:
            self.bits.append(0)

    def tobytes(self):
        assert len(self.bits) % 8 == 0
        return bytes(np.packbits(np.array(self.bits, np.uint8)).tolist())


def nal_unit(ref_idc, nal_type, rbsp: bytes) -> bytes:
    out = bytearray(b"\x00\x00\x00\x01")
    out.append((ref_idc << 5) | nal_type)
    zeros = 0
    for b in rbsp:
        if zeros >= 2 and b <= 3:
            out.append(3)  # emulation_prevention_three_byte
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def write_scaling_matrix(w, matrix, n_lists):
    """7.3.2.1.1.1 scaling_list() for lists 0..n_lists-1 (0..5: 4x4, 6..: 8x8). `matrix`: {list index: values in zig-zag
    order (16 or 64 of them, 1..255) | "default" (the list is present and signals useDefaultScalingMatrixFlag)};
    lists that are not named are written as absent (scaling_list_present_flag = 0)."""
    for i in range(n_lists):
        v = matrix.get(i)
        w.u(1, 0 if v is None else 1)
        if v is None:
            continue
        if isinstance(v, str):
            w.se(-8)          # delta_scale: nextScale = 0 at j = 0 -> useDefaultScalingMatrixFlag
            continue
        assert len(v) == (16 if i < 6 else 64) and all(1 <= x <= 255 for x in v)
        last = 8
        for x in v:
            d = (x - last + 128) % 256 - 128
            w.se(d)
            last = x


def sps_rbsp(w_mbs, h_mbs, crop=None, matrix=None, sps_id=0):
    """crop: None or (left, right, top, bottom) frame_crop_*_offset values (units of two luma samples for 4:2:0 frames).
    matrix: None or the seq_scaling_matrix (see write_scaling_matrix; 8 lists)."""
    w = BitWriter()
    w.u(8, 100)   # profile_idc: High
    w.u(8, 0)     # constraint flags + reserved
    w.u(8, 51)    # level_idc
    w.ue(sps_id)  # seq_parameter_set_id
    w.ue(1)       # chroma_format_idc 4:2:0
    w.ue(0)       # bit_depth_luma_minus8
    w.ue(0)       # bit_depth_chroma_minus8
    w.u(1, 0)     # qpprime_y_zero_transform_bypass_flag
    w.u(1, 0 if matrix is None else 1)     # seq_scaling_matrix_present_flag
    if matrix is not None:
        write_scaling_matrix(w, matrix, 8)
    w.ue(0)       # log2_max_frame_num_minus4
    w.ue(2)       # pic_order_cnt_type
    w.ue(1)       # max_num_ref_frames
    w.u(1, 0)     # gaps_in_frame_num_value_allowed_flag
    w.ue(w_mbs - 1)
    w.ue(h_mbs - 1)
    w.u(1, 1)     # frame_mbs_only_flag
    w.u(1, 1)     # direct_8x8_inference_flag
    w.u(1, 1 if crop else 0)     # frame_cropping_flag
    if crop:
        for v in crop:
            w.ue(v)   # frame_crop_left / right / top / bottom_offset
    w.u(1, 0)     # vui_parameters_present_flag
    w.trailing()
    return w.tobytes()


def pps_rbsp(cb_off, cr_off, matrix=None, pps_id=0, sps_id=0, pic_init_qp=26):
    w = BitWriter()
    w.ue(pps_id)  # pic_parameter_set_id
    w.ue(sps_id)  # seq_parameter_set_id
    w.u(1, 1)     # entropy_coding_mode_flag: CABAC
    w.u(1, 0)     # bottom_field_pic_order_in_frame_present_flag
    w.ue(0)       # num_slice_groups_minus1
    w.ue(0)       # num_ref_idx_l0_default_active_minus1
    w.ue(0)       # num_ref_idx_l1_default_active_minus1
    w.u(1, 0)     # weighted_pred_flag
    w.u(2, 0)     # weighted_bipred_idc
    w.se(pic_init_qp - 26)   # pic_init_qp_minus26
    w.se(0)       # pic_init_qs_minus26
    w.se(cb_off)  # chroma_qp_index_offset
    w.u(1, 1)     # deblocking_filter_control_present_flag
    w.u(1, 0)     # constrained_intra_pred_flag
    w.u(1, 0)     # redundant_pic_cnt_present_flag
    w.u(1, 1)     # transform_8x8_mode_flag
    w.u(1, 0 if matrix is None else 1)     # pic_scaling_matrix_present_flag
    if matrix is not None:
        write_scaling_matrix(w, matrix, 8)   # 6 + 2 * transform_8x8_mode_flag lists (4:2:0)
    w.se(cr_off)  # second_chroma_qp_index_offset
    w.trailing()
    return w.tobytes()


class Cabac:
    """9.3.4.2 arithmetic encoder + 9.3.1.1 context initialisation (I slice)."""

    def __init__(self, slice_qp, out_bits):
        self.low, self.range = 0, 510
        self.first, self.outstanding = True, 0
        self.bits = out_bits
        self.state, self.mps = [], []
        q = min(max(slice_qp, 0), 51)
        for m, n in CTX_INIT_I:
            pre = min(max(((m * q) >> 4) + n, 1), 126)
            if pre <= 63:
                self.state.append(63 - pre)
                self.mps.append(0)
            else:
                self.state.append(pre - 64)
                self.mps.append(1)

    def _put(self, b):
        if self.first:
            self.first = False
        else:
            self.bits.append(b)
        while self.outstanding > 0:
            self.bits.append(1 - b)
            self.outstanding -= 1

    def _renorm(self):
        while self.range < 256:
            if self.low < 256:
                self._put(0)
            elif self.low >= 512:
                self.low -= 512
                self._put(1)
            else:
                self.low -= 256
                self.outstanding += 1
            self.range <<= 1
            self.low <<= 1

    def decision(self, ctx, b):
        s = self.state[ctx]
        lps = RANGE_LPS[s][(self.range >> 6) & 3]
        self.range -= lps
        if b != self.mps[ctx]:
            self.low += self.range
            self.range = lps
            if s == 0:
                self.mps[ctx] = 1 - self.mps[ctx]
            self.state[ctx] = TRANS_LPS[s]
        else:
            self.state[ctx] = TRANS_MPS[s]
        self._renorm()

    def bypass(self, b):
        self.low <<= 1
        if b:
            self.low += self.range
        if self.low >= 1024:
            self._put(1)
            self.low -= 1024
        elif self.low < 512:
            self._put(0)
        else:
            self.low -= 512
            self.outstanding += 1

    def terminate(self, b):
        self.range -= 2
        if b:
            self.low += self.range
            self.range = 2
            self._renorm()
            self._put((self.low >> 9) & 1)
            self.bits.append((self.low >> 8) & 1)
            self.bits.append(1)  # doubles as rbsp_stop_one_bit
        else:
            self._renorm()


class MbInfo:
    """What later macroblocks' context selection needs to know about an encoded macroblock."""
    __slots__ = ("i16", "t8", "cbp_luma", "cbp_chroma", "chroma_mode", "cbf_luma", "cbf_dc", "cbf_cdc", "cbf_cac")

    def __init__(self):
        self.i16 = False
        self.t8 = 0
        self.cbp_luma = 0
        self.cbp_chroma = 0
        self.chroma_mode = 0
        self.cbf_luma = [0] * 16      # per 4x4 luma block (an 8x8-transform block sets its four)
        self.cbf_dc = 0               # Intra16x16 luma DC block
        self.cbf_cdc = [0, 0]         # chroma DC per plane
        self.cbf_cac = [[0] * 4, [0] * 4]


def mb_fields(batch, idx):
    """(i16, pred16, t8, levels int[24][16]) of macroblock idx."""
    code = int(batch.mb_type[idx])
    i16 = code != 0
    return i16, (code - 1) % 4 if i16 else 0, int(batch.transform_size_8x8_flag[idx]) if not i16 else 0, \
        batch.coeff[idx].astype(np.int64).reshape(24, 16)


def canonicalise(batch):
    """Make the syntax buffers expressible as a bitstream without changing what they reconstruct to: the cbp fields
    folded into an I_16x16 mb_type must agree with the non-zero blocks, and a macroblock that sends no mb_qp_delta
    (I_NxN with coded_block_pattern 0) inherits the previous QP. Returns the per-MB (cbp_luma, cbp_chroma)."""
    n = batch.mb_type.size
    per = batch.pp.n_mb
    cbps = []
    for idx in range(n):
        i16, pred16, t8, lv = mb_fields(batch, idx)
        nz = lv != 0
        cdc = nz[16:24, 0].any()
        cac = nz[16:24, 1:].any()
        cbp_c = 2 if cac else (1 if cdc else 0)
        if i16:
            cbp_l = 15 if nz[:16, 1:].any() else 0
            batch.mb_type[idx] = 1 + pred16 + 4 * cbp_c + (12 if cbp_l else 0)
        else:
            cbp_l = sum(1 << b8 for b8 in range(4) if nz[4 * b8:4 * b8 + 4].any())
            if cbp_l == 0 and cbp_c == 0:
                first = idx % per == 0
                batch.qp[idx] = batch.qp[idx] if first else batch.qp[idx - 1]
        cbps.append((cbp_l, cbp_c))
    return cbps


class SliceWriter:
    def __init__(self, batch, frame, cbps, slice_qp):
        self.b, self.f, self.cbps = batch, frame, cbps
        self.W, self.H = batch.pp.pic_width_in_mbs, batch.pp.pic_height_in_mbs
        self.base = frame * batch.pp.n_mb
        self.info = [None] * (self.W * self.H)
        self.bits = []
        self.c = Cabac(slice_qp, self.bits)
        self.qp_prev = slice_qp
        self.prev_delta_nonzero = False

    # -- neighbours --------------------------------------------------------------------------------
    def nb(self, addr):
        x, y = addr % self.W, addr // self.W
        a = self.info[addr - 1] if x > 0 else None
        b = self.info[addr - self.W] if y > 0 else None
        return a, b

    # -- syntax elements -----------------------------------------------------------------------------
    def mb_type_i(self, a, b, i16, pred16, cbp_l, cbp_c):
        inc = (1 if (a is not None and a.i16) else 0) + (1 if (b is not None and b.i16) else 0)
        c = self.c
        if not i16:
            c.decision(3 + inc, 0)
            return
        c.decision(3 + inc, 1)
        c.terminate(0)                       # not I_PCM
        c.decision(3 + 3, 1 if cbp_l else 0)
        if cbp_c == 0:
            c.decision(3 + 4, 0)
            c.decision(3 + 6, pred16 >> 1)
            c.decision(3 + 7, pred16 & 1)
        else:
            c.decision(3 + 4, 1)
            c.decision(3 + 5, 1 if cbp_c == 2 else 0)
            c.decision(3 + 6, pred16 >> 1)
            c.decision(3 + 7, pred16 & 1)

    def coded_block_pattern(self, a, b, cbp_l, cbp_c):
        c = self.c
        for b8 in range(4):
            x8, y8 = b8 & 1, b8 >> 1
            # condTermFlagN = 0 if N unavailable / its 8x8 block is coded, else 1 (9.3.3.1.1.4)
            if x8 > 0:
                ca = 0 if (cbp_l >> (b8 - 1)) & 1 else 1
            else:
                ca = 0 if a is None else (0 if (a.cbp_luma >> (b8 + 1)) & 1 else 1)
            if y8 > 0:
                cb = 0 if (cbp_l >> (b8 - 2)) & 1 else 1
            else:
                cb = 0 if b is None else (0 if (b.cbp_luma >> (b8 + 2)) & 1 else 1)
            c.decision(73 + ca + 2 * cb, (cbp_l >> b8) & 1)
        ca = 1 if (a is not None and a.cbp_chroma != 0) else 0
        cb = 1 if (b is not None and b.cbp_chroma != 0) else 0
        c.decision(77 + ca + 2 * cb, 1 if cbp_c else 0)
        if cbp_c:
            ca = 1 if (a is not None and a.cbp_chroma == 2) else 0
            cb = 1 if (b is not None and b.cbp_chroma == 2) else 0
            c.decision(77 + 4 + ca + 2 * cb, 1 if cbp_c == 2 else 0)

    def mb_qp_delta(self, delta):
        c = self.c
        v = 2 * delta - 1 if delta > 0 else -2 * delta
        ctx = 60 + (1 if self.prev_delta_nonzero else 0)
        k = 0
        while True:
            c.decision(ctx, 1 if k < v else 0)
            if k >= v:
                break
            k += 1
            ctx = 60 + 2 if k == 1 else 60 + 3
        self.prev_delta_nonzero = delta != 0

    def residual_block(self, cat, coeffs, cbf_inc, code_cbf=True):
        """coeffs: the block's levels in coding order (16, 15, 4 or 64 of them). Returns coded_block_flag."""
        c = self.c
        nz = [i for i, v in enumerate(coeffs) if v != 0]
        coded = 1 if nz else 0
        if code_cbf:
            c.decision(CBF_BASE[cat] + cbf_inc, coded)
        if not coded:
            return 0
        n = len(coeffs)
        last = nz[-1]
        for i in range(n - 1):
            if cat == 5:
                si, li = SIG8[i], LAST8[i]
            elif cat == 3:
                si = li = min(i, 2)
            else:
                si = li = i
            sig = 1 if coeffs[i] != 0 else 0
            c.decision(SIG_BASE[cat] + si, sig)
            if sig:
                c.decision(LAST_BASE[cat] + li, 1 if i == last else 0)
                if i == last:
                    break
        eq1 = gt1 = 0
        for i in reversed(nz):
            v = int(coeffs[i])
            a = abs(v) - 1
            ctx0 = ABS_BASE[cat] + (0 if gt1 else min(4, 1 + eq1))
            ctxn = ABS_BASE[cat] + 5 + min(4 - (1 if cat == 3 else 0), gt1)
            # prefix: truncated unary, cMax 14
            pre = min(a, 14)
            for k in range(pre):
                c.decision(ctx0 if k == 0 else ctxn, 1)
            if pre < 14:
                c.decision(ctx0 if pre == 0 else ctxn, 0)
            else:  # suffix: 0-th order Exp-Golomb, bypass
                s, k = a - 14, 0
                while s >= (1 << k):
                    c.bypass(1)
                    s -= 1 << k
                    k += 1
                c.bypass(0)
                for j in range(k - 1, -1, -1):
                    c.bypass((s >> j) & 1)
            c.bypass(1 if v < 0 else 0)
            if a == 0:
                eq1 += 1
            else:
                gt1 += 1
        return 1

    # -- coded_block_flag context increments (9.3.3.1.1.9), intra macroblocks only ------------------
    @staticmethod
    def _cond(nbinfo, flag_of):
        if nbinfo is None:
            return 1          # unavailable neighbour, current MB intra
        return flag_of(nbinfo)

    def luma_blk_nb(self, me, a, b, blk):
        """cbf of the 4x4 luma blocks left of / above block blk (None info = unavailable MB)."""
        x, y = BLK4_XY[blk]
        fa = me.cbf_luma[blk4_of(x - 4, y)] if x > 0 else self._cond(a, lambda m: m.cbf_luma[blk4_of(12, y)])
        fb = me.cbf_luma[blk4_of(x, y - 4)] if y > 0 else self._cond(b, lambda m: m.cbf_luma[blk4_of(x, 12)])
        return fa + 2 * fb

    # -- one macroblock -------------------------------------------------------------------------------
    def macroblock(self, addr):
        bt, c = self.b, self.c
        idx = self.base + addr
        a, b = self.nb(addr)
        i16, pred16, t8, lv = mb_fields(bt, idx)
        cbp_l, cbp_c = self.cbps[idx]
        me = MbInfo()
        me.i16, me.t8, me.cbp_luma, me.cbp_chroma = i16, t8, cbp_l, cbp_c
        me.chroma_mode = int(bt.intra_chroma_pred_mode[idx])

        self.mb_type_i(a, b, i16, pred16, cbp_l, cbp_c)
        if not i16:
            inc = (1 if (a is not None and a.t8) else 0) + (1 if (b is not None and b.t8) else 0)
            c.decision(399 + inc, t8)
            for k in range(4 if t8 else 16):
                syn = int(bt.pred_syntax[idx, k])
                c.decision(68, (syn >> 3) & 1)
                if not (syn >> 3) & 1:
                    for bit in range(3):
                        c.decision(69, (syn >> bit) & 1)
        # intra_chroma_pred_mode: TU, cMax 3
        inc = (1 if (a is not None and a.chroma_mode != 0) else 0) + (1 if (b is not None and b.chroma_mode != 0) else 0)
        cm = me.chroma_mode
        c.decision(64 + inc, 1 if cm > 0 else 0)
        if cm > 0:
            c.decision(64 + 3, 1 if cm > 1 else 0)
            if cm > 1:
                c.decision(64 + 3, 1 if cm > 2 else 0)
        if not i16:
            self.coded_block_pattern(a, b, cbp_l, cbp_c)
        qp = int(bt.qp[idx])
        if i16 or cbp_l or cbp_c:
            self.mb_qp_delta(qp - self.qp_prev)
            self.qp_prev = qp
            # ---- residual ------------------------------------------------------------------------
            if i16:
                inc = self._cond(a, lambda m: m.cbf_dc if m.i16 else 0) + 2 * self._cond(b, lambda m: m.cbf_dc if m.i16 else 0)
                me.cbf_dc = self.residual_block(0, lv[:16, 0], inc)
            for b8 in range(4):
                if not (cbp_l >> b8) & 1:
                    continue
                if t8:
                    self.residual_block(5, lv[4 * b8:4 * b8 + 4].reshape(64), 0, code_cbf=False)
                    for k in range(4):
                        me.cbf_luma[4 * b8 + k] = 1
                else:
                    for k in range(4):
                        blk = 4 * b8 + k
                        inc = self.luma_blk_nb(me, a, b, blk)
                        me.cbf_luma[blk] = self.residual_block(1, lv[blk, 1:], inc) if i16 else \
                            self.residual_block(2, lv[blk], inc)
            if cbp_c:
                for pl in range(2):
                    inc = self._cond(a, lambda m: m.cbf_cdc[pl]) + 2 * self._cond(b, lambda m: m.cbf_cdc[pl])
                    me.cbf_cdc[pl] = self.residual_block(3, lv[16 + 4 * pl:20 + 4 * pl, 0], inc)
            if cbp_c == 2:
                for pl in range(2):
                    for k in range(4):
                        kx, ky = k & 1, k >> 1
                        fa = me.cbf_cac[pl][k - 1] if kx else self._cond(a, lambda m: m.cbf_cac[pl][k + 1])
                        fb = me.cbf_cac[pl][k - 2] if ky else self._cond(b, lambda m: m.cbf_cac[pl][k + 2])
                        me.cbf_cac[pl][k] = self.residual_block(4, lv[16 + 4 * pl + k, 1:], fa + 2 * fb)
        else:
            self.prev_delta_nonzero = False
        self.info[addr] = me

    def slice_data(self):
        n = self.W * self.H
        for addr in range(n):
            self.macroblock(addr)
            self.c.terminate(1 if addr == n - 1 else 0)   # end_of_slice_flag
        return self.bits


def encode_picture(batch, frame, cbps, idr_pic_id=0, deblock=None, pps_id=0, pic_init_qp=26) -> bytes:
    """One IDR access unit (slice NAL only) of picture `frame`. deblock: None (filter disabled, what dryv decodes) or
    (slice_alpha_c0_offset_div2, slice_beta_offset_div2) for a stream that asks for the in-loop filter."""
    base = frame * batch.pp.n_mb
    slice_qp = int(batch.qp[base])
    w = BitWriter()
    w.ue(0)                 # first_mb_in_slice
    w.ue(7)                 # slice_type: I (all slices of the picture)
    w.ue(pps_id)            # pic_parameter_set_id
    w.u(4, 0)               # frame_num
    w.ue(idr_pic_id)        # idr_pic_id
    w.u(1, 0)               # no_output_of_prior_pics_flag
    w.u(1, 0)               # long_term_reference_flag
    w.se(slice_qp - pic_init_qp)     # slice_qp_delta
    if deblock is None:
        w.ue(1)             # disable_deblocking_filter_idc: no deblocking (dryv has none)
    else:
        w.ue(0)             # filter every edge
        w.se(deblock[0])    # slice_alpha_c0_offset_div2
        w.se(deblock[1])    # slice_beta_offset_div2
    while len(w.bits) % 8:
        w.bits.append(1)    # cabac_alignment_one_bit
    sw = SliceWriter(batch, frame, cbps, slice_qp)
    w.bits += sw.slice_data()   # ends with the terminate bin's stop bit
    while len(w.bits) % 8:
        w.bits.append(0)
    return nal_unit(3, 5, w.tobytes())


def encode_stream(batch, crop=None, deblock=None, sps_matrix=None, pps_matrix=None, extra_pps=None) -> bytes:
    """Annex-B stream: SPS, PPS, then one IDR picture per frame of `batch` (canonicalises `batch` in place). `crop`: the
    SPS frame_crop_{left,right,top,bottom}_offset values, or None for no cropping. `sps_matrix` / `pps_matrix`: scaling
    matrices (write_scaling_matrix); the levels of `batch` are written as they are, so the caller generates them against
    the lists the decoder will end up with. `extra_pps`: None or (pic_init_qp, which) - a second PPS (id 1, that
    pic_init_qp, same offsets) is written as well and the pictures f with which(f) true name it."""
    pp = batch.pp
    if sps_matrix is None and pps_matrix is None:
        assert all(v == 16 for v in pp.scaling_list4x4) and all(v == 16 for v in pp.scaling_list8x8), "flat lists only"
    cbps = canonicalise(batch)
    cb, cr = int(pp.chroma_qp_index_offset), int(pp.second_chroma_qp_index_offset)
    out = nal_unit(3, 7, sps_rbsp(pp.pic_width_in_mbs, pp.pic_height_in_mbs, crop, sps_matrix))
    out += nal_unit(3, 8, pps_rbsp(cb, cr, pps_matrix))
    if extra_pps:
        out += nal_unit(3, 8, pps_rbsp(cb, cr, pps_matrix, pps_id=1, pic_init_qp=extra_pps[0]))
    for f in range(batch.n_frames):
        second = bool(extra_pps and extra_pps[1](f))
        out += encode_picture(batch, f, cbps, idr_pic_id=f & 1, deblock=deblock, pps_id=1 if second else 0,
                              pic_init_qp=extra_pps[0] if second else 26)
    return out
