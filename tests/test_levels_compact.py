"""The compact level stream (include/dryv_recon.h, dryv_mb_levels_compact): host-side format tests (no GPU) and,
marked gpu, the expansion kernel and dryv_recon_submit_compact against the dense path and the oracle."""
import ctypes as C

import numpy as np
import pytest

from dryv_b200 import recon, synth
from dryv_b200.abi import COMPACT_MAX_RECORD, LevelsCompact, PicParams


def reference_pack(coeff):
    """Independent pure-numpy statement of the record layout documented in the header."""
    offs, out = [0], bytearray()
    for mb in np.asarray(coeff, np.int16).reshape(-1, 24, 16):
        nzm = mb != 0
        coded = nzm.any(axis=1)
        lv = mb[nzm].astype(np.int32)
        wide = bool(((lv < -128) | (lv > 127)).any())
        small = np.abs(lv) <= 7
        int8 = lv.astype(np.int8).tobytes()
        int16 = lv.astype("<i2").tobytes()
        nib = bytearray(((len(lv) + 1) // 2 + 1) & ~1)
        for j, v in enumerate(lv):
            code = (abs(int(v)) | (8 if v < 0 else 0)) if small[j] else 0
            nib[j >> 1] |= code << (4 * (j & 1))
        nibble = bytes(nib) + lv[~small].astype("<i2").tobytes()
        mode, levels = (2, int16) if wide else (0, int8)
        if len(nibble) < len(levels):
            mode, levels = 1, nibble
        hdr = sum(1 << b for b in range(24) if coded[b]) | (mode << 30)
        rec = bytearray(np.uint32(hdr).tobytes())
        for b in range(24):
            if coded[b]:
                rec += np.uint16(sum(1 << k for k in range(16) if nzm[b, k])).tobytes()
        rec += levels
        rec += bytes(-len(rec) % 4)
        out += rec
        offs.append(len(out))
    return np.array(offs, np.uint32), np.frombuffer(bytes(out), np.uint8)


def test_pack_matches_documented_layout(recon_lib):
    pp = PicParams.make(5, 4)
    for seed, qp, stress in ((91, 26, 30), (92, 4, 100), (93, 40, 0)):
        b = synth.generate(pp, 2, seed, qp_base=qp, stress_pct=stress)
        lv = recon.pack_levels(b.coeff, threads=1)
        off, stream = reference_pack(b.coeff)
        assert np.array_equal(lv.offset, off)
        assert np.array_equal(lv.stream[:off[-1]], stream)
        modes = set(int(lv.stream[o + 3]) >> 6 for o in lv.offset[:-1])
        assert modes <= {0, 1, 2}
    assert (lv.offset % 4 == 0).all() and np.diff(lv.offset.astype(np.int64)).max() <= COMPACT_MAX_RECORD


@pytest.mark.parametrize("qp", [0, 26, 51])
@pytest.mark.parametrize("threads", [1, 5])
def test_pack_unpack_roundtrip(recon_lib, qp, threads):
    pp = PicParams.make(40, 23)
    b = synth.generate(pp, 5, 360 + qp, qp_base=qp)
    lv = recon.pack_levels(b.coeff, threads=threads)
    assert np.array_equal(lv.unpack(), b.coeff)
    if qp >= 26:
        assert lv.nbytes < b.coeff.nbytes / 2


def test_edge_records(recon_lib):
    c = np.zeros((10, 384), np.int16)
    c[1, :] = 1                      # every level set, 4-bit codes
    c[2, :] = -32768                 # every level set, int16: the largest record
    c[3, 383] = 127                  # int8 boundary values
    c[4, 0] = -128
    c[5, 17] = 128                   # first int16 value
    c[6, :] = -7                     # 4-bit boundary values
    c[7, ::3] = 7
    c[8, :] = 8                      # all escapes would cost more than int8
    c[9, :5] = [1, -1, 300, 2, -2]   # 4-bit codes with one escape
    lv = recon.pack_levels(c, threads=1)
    sizes = np.diff(lv.offset.astype(np.int64))
    assert sizes.tolist() == [4, 4 + 48 + 192, COMPACT_MAX_RECORD, 8, 8, 8, 4 + 48 + 192, 4 + 48 + 64, 4 + 48 + 384,
                              4 + 2 + 4 + 2]
    modes = [int(lv.stream[o + 3]) >> 6 for o in lv.offset[:-1]]
    assert modes == [0, 1, 2, 0, 0, 2, 1, 1, 0, 1]
    assert np.array_equal(lv.unpack(), c)
    off, stream = reference_pack(c)
    assert np.array_equal(lv.offset, off) and np.array_equal(lv.stream[:off[-1]], stream)


def test_malformed_streams_are_rejected(recon_lib):
    c = np.zeros((3, 384), np.int16)
    c[1, 5] = 9
    lv = recon.pack_levels(c, threads=1)
    out = np.empty((3, 384), np.int16)

    def unpack(offset, stream):
        s = LevelsCompact()
        s.offset, s.stream = offset.ctypes.data, stream.ctypes.data
        return recon_lib.dryv_recon_unpack_levels(C.byref(s), 3, out.ctypes.data)

    assert unpack(lv.offset, lv.stream) == recon.OK
    bad = lv.offset.copy()
    bad[2] = bad[1]                  # record shorter than its header says
    assert unpack(bad, lv.stream) == recon.ERR_ARG
    bad = lv.offset.copy()
    bad[1] += 2                      # not 4-byte aligned
    assert unpack(bad, lv.stream) == recon.ERR_ARG
    st = lv.stream.copy()
    st[lv.offset[1] + 3] |= 0x20     # reserved header bit
    assert unpack(lv.offset, st) == recon.ERR_ARG
    st = lv.stream.copy()
    st[lv.offset[1] + 3] |= 0xc0     # level coding 3 does not exist
    assert unpack(lv.offset, st) == recon.ERR_ARG
    st = lv.stream.copy()
    st[lv.offset[1] + 4:lv.offset[1] + 6] = 0   # coded slot with an empty mask
    assert unpack(lv.offset, st) == recon.ERR_ARG
    assert recon_lib.dryv_recon_pack_levels(None, 1, None, None, 0, 1) == recon.ERR_ARG
    tiny = np.empty(4, np.uint8)
    assert recon_lib.dryv_recon_pack_levels(c.ctypes.data, 3, bad.ctypes.data, tiny.ctypes.data, 4, 1) == recon.ERR_ARG
    assert recon_lib.dryv_recon_submit_compact(None, None, None, None, 1, None) == recon.ERR_ARG


# ---- GPU ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("qp,stress", [(0, 10), (26, 10), (26, 70), (45, 0), (51, 100)])
def test_expand_kernel_matches_dense(gpu_ctx, qp, stress):
    import torch
    pp = PicParams.make(17, 9)
    b = synth.generate(pp, 3, 4100 + qp + stress, qp_base=qp, stress_pct=stress)
    lv = recon.pack_levels(b.coeff)
    d_off = torch.from_numpy(lv.offset.view(np.int32)).cuda()
    d_str = torch.from_numpy(lv.stream).cuda()
    d_coeff = torch.full((lv.n_mbs, 384), 0x5555, dtype=torch.int16, device="cuda")
    gpu_ctx.expand_levels_device(d_off, d_str, lv.n_mbs, d_coeff)
    gpu_ctx.wait()
    assert np.array_equal(d_coeff.cpu().numpy(), b.coeff)


@pytest.mark.gpu
def test_expand_kernel_edge_records(gpu_ctx):
    import torch
    c = np.zeros((12, 384), np.int16)
    c[1, :] = -1
    c[2, :] = -32768
    c[3, 383] = 127
    c[4, 0] = -128
    c[5, 17] = 128
    c[6, ::2] = 32767
    c[7, :] = -7
    c[8, ::3] = 7
    c[9, :] = 8
    c[10, :5] = [1, -1, 300, 2, -2]
    c[11, :] = np.where(np.arange(384) % 5 == 0, -3000, np.arange(384) % 7 + 1)   # 4-bit codes, escapes in every slot
    lv = recon.pack_levels(c, threads=1)
    assert sorted(set(int(lv.stream[o + 3]) >> 6 for o in lv.offset[:-1])) == [0, 1, 2]
    d_off = torch.from_numpy(lv.offset.view(np.int32)).cuda()
    d_str = torch.from_numpy(lv.stream).cuda()
    d_coeff = torch.zeros((12, 384), dtype=torch.int16, device="cuda")
    gpu_ctx.expand_levels_device(d_off, d_str, 12, d_coeff)
    gpu_ctx.wait()
    assert np.array_equal(d_coeff.cpu().numpy(), c)


@pytest.mark.gpu
def test_malformed_stream_is_reported_not_read_out_of_bounds(gpu_ctx):
    import torch
    c = np.zeros((4, 384), np.int16)
    c[:, 3] = 5
    lv = recon.pack_levels(c, threads=1)
    off = lv.offset.copy()
    off[2] = off[1]                  # record 1 truncated to nothing
    d_off = torch.from_numpy(off.view(np.int32)).cuda()
    d_str = torch.from_numpy(lv.stream).cuda()
    d_coeff = torch.zeros((4, 384), dtype=torch.int16, device="cuda")
    gpu_ctx.expand_levels_device(d_off, d_str, 4, d_coeff)
    with pytest.raises(recon.ReconError) as e:
        gpu_ctx.wait()
    assert e.value.code == recon.ERR_UNSUPPORTED
    # the context stays usable
    d_off = torch.from_numpy(lv.offset.view(np.int32)).cuda()
    gpu_ctx.expand_levels_device(d_off, d_str, 4, d_coeff)
    gpu_ctx.wait()
    assert np.array_equal(d_coeff.cpu().numpy(), c)
    # a misaligned stream pointer / first offset is an argument error, not a misaligned-address fault
    d_pad = torch.zeros(lv.stream.size + 8, dtype=torch.uint8, device="cuda")
    d_pad[2:2 + lv.stream.size] = d_str
    with pytest.raises(recon.ReconError) as e:
        gpu_ctx.expand_levels_device(d_off, d_pad[2:], 4, d_coeff)
    assert e.value.code == recon.ERR_ARG
    off2 = lv.offset.copy() + 2
    with pytest.raises(recon.ReconError) as e:
        gpu_ctx.expand_levels_device(torch.from_numpy(off2.view(np.int32)).cuda(), d_pad, 4, d_coeff)
    assert e.value.code == recon.ERR_ARG
    gpu_ctx.expand_levels_device(d_off, d_str, 4, d_coeff)   # and the context is still fine
    gpu_ctx.wait()
    assert np.array_equal(d_coeff.cpu().numpy(), c)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n", [(1, 1, 1), (7, 3, 5), (40, 23, 9), (120, 68, 11)])
def test_submit_compact_equals_dense_and_oracle(gpu_ctx, w, h, n):
    import oracle
    pp = PicParams.make(w, h, 1, -1)
    b = synth.generate(pp, n, 8800 + w)
    lv = recon.pack_levels(b.coeff)
    got = gpu_ctx.reconstruct_compact(b, lv)
    assert np.array_equal(got, gpu_ctx.reconstruct(b))
    k = min(n, 3)
    assert np.array_equal(got[:k], oracle.reconstruct(b.frames(0, k), threads=4))


@pytest.mark.gpu
def test_submit_compact_many_chunks_pinned(gpu_ctx, monkeypatch):
    # small stages: every staging slot and both control blocks get reused several times
    monkeypatch.setenv("DRYV_CHUNK_OUT_MB", "1")
    pp = PicParams.make(40, 23)
    hb, owners = recon.pinned_batch(pp, 24)
    synth.generate(pp, 24, 31337, out=hb)
    lv = recon.pack_levels(hb.coeff, pinned=True)
    out = recon.PinnedArray((24, pp.frame_bytes), np.uint8)
    ref = None
    for _ in range(3):
        out.array[:] = 0
        gpu_ctx.submit_compact(hb, lv, out.array)
        gpu_ctx.wait()
        if ref is None:
            monkeypatch.delenv("DRYV_CHUNK_OUT_MB")
            ref = gpu_ctx.reconstruct(hb)
            monkeypatch.setenv("DRYV_CHUNK_OUT_MB", "1")
        assert np.array_equal(out.array, ref)
    assert gpu_ctx.last_submit_ms > 0


@pytest.mark.gpu
def test_queued_submits_and_wait_oldest(gpu_ctx):
    # streaming use: several batches queued before the first wait; both wire formats; results land in order
    pp = PicParams.make(40, 23)
    batches, levels, outs, refs = [], [], [], []
    for k in range(6):
        hb, owners = recon.pinned_batch(pp, 5 + k)
        synth.generate(pp, 5 + k, 900 + k, out=hb)
        batches.append((hb, owners))
        levels.append(recon.pack_levels(hb.coeff, pinned=True))
        outs.append(recon.PinnedArray((5 + k, pp.frame_bytes), np.uint8))
    for hb, _ in batches:
        refs.append(gpu_ctx.reconstruct(hb).copy())
    for k, (hb, _) in enumerate(batches):          # six submits, the queue holds four: the fifth blocks on the oldest
        if k % 2:
            gpu_ctx.submit(hb, outs[k].array)
        else:
            gpu_ctx.submit_compact(hb, levels[k], outs[k].array)
    gpu_ctx.wait_oldest()
    assert np.array_equal(outs[0].array, refs[0]) and np.array_equal(outs[1].array, refs[1])
    gpu_ctx.wait_oldest()
    assert np.array_equal(outs[2].array, refs[2])
    gpu_ctx.wait()
    for k in range(6):
        assert np.array_equal(outs[k].array, refs[k]), k
    assert gpu_ctx.last_submit_ms > 0
    gpu_ctx.wait_oldest()                           # nothing outstanding: returns at once
