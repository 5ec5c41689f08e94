"""The host-side mirror of the reference interface (Frame::new / decode / write_to_yuv_file)."""
import numpy as np
import pytest

import oracle
from dryv_b200 import frame as F
from dryv_b200 import synth
from dryv_b200.abi import PicParams


def mbs_from_batch(batch, idx):
    """Rebuild the reference's per-MB record (struct Macroblock) from SoA record idx."""
    mb = F.Macroblock()
    mb.mb_type = int(batch.mb_type[idx])
    mb.transform_size_8x8_flag = int(batch.transform_size_8x8_flag[idx])
    mb.qp1y = int(batch.qp[idx])
    mb.intra_chroma_pred_mode = int(batch.intra_chroma_pred_mode[idx])
    ps = batch.pred_syntax[idx]
    cf = batch.coeff[idx].astype(np.int64)
    if mb.mb_type == 0 and not mb.transform_size_8x8_flag:
        mb.prev_intra4x4_pred_mode_flag = [(int(v) >> 3) & 1 for v in ps]
        mb.rem_intra4x4_pred_mode = [int(v) & 7 for v in ps]
        mb.block_luma_4x4 = cf[:256].reshape(16, 16)
    elif mb.mb_type == 0:
        mb.prev_intra8x8_pred_mode_flag = [(int(v) >> 3) & 1 for v in ps[:4]]
        mb.rem_intra8x8_pred_mode = [int(v) & 7 for v in ps[:4]]
        mb.block_luma_8x8 = cf[:256].reshape(4, 64)
    else:
        l = cf[:256].reshape(16, 16)
        mb.block_luma_dc = l[:, 0].copy()
        mb.block_luma_ac = l[:, 1:].copy()
    for pl in range(2):
        c = cf[256 + 64 * pl:320 + 64 * pl].reshape(4, 16)
        mb.block_chroma_dc[pl, :4] = c[:, 0]
        mb.block_chroma_ac[pl, :4, :] = c[:, 1:]
    return mb


def drive(batch, ctx=None):
    pp = batch.pp
    sl = F.Slice(pp.pic_width_in_mbs, pp.pic_height_in_mbs, pp.chroma_qp_index_offset,
                 pp.second_chroma_qp_index_offset)
    fr = F.Frame.new(sl, ctx)
    for a in range(pp.n_mb):
        sl.curr_mb_addr = a
        sl.macroblock = mbs_from_batch(batch, a)
        fr.decode(sl)
    return fr


def test_decode_packs_the_soa_record():
    pp = PicParams.make(5, 3, 2, -1)
    b = synth.generate(pp, 1, 21)
    fr = drive(b)
    for name in ("mb_type", "transform_size_8x8_flag", "intra_chroma_pred_mode", "qp", "coeff"):
        assert np.array_equal(getattr(fr.batch, name), getattr(b, name)), name
    # prev/rem bits: don't-care rem bits of flagged blocks are preserved too
    used = np.where((b.mb_type == 0)[:, None], np.where((b.transform_size_8x8_flag == 1)[:, None], np.arange(16) < 4, True), False)
    assert np.array_equal(fr.batch.pred_syntax[used], b.pred_syntax[used])
    assert fr.decoded.all() and fr.pp.chroma_qp_index_offset == 2 and fr.pp.second_chroma_qp_index_offset == -1


def test_unsupported_macroblocks_raise_like_the_reference():
    sl = F.Slice(1, 1)
    fr = F.Frame.new(sl)
    sl.macroblock = F.Macroblock(mb_type=25)
    with pytest.raises(NotImplementedError, match="I PCM"):
        fr.decode(sl)
    sl.macroblock = F.Macroblock(mb_type=30)
    with pytest.raises(NotImplementedError, match="Inter"):
        fr.decode(sl)


def test_incomplete_picture_cannot_be_read():
    sl = F.Slice(2, 1)
    fr = F.Frame.new(sl)
    sl.macroblock = F.Macroblock(mb_type=3)
    fr.decode(sl)
    with pytest.raises(RuntimeError):
        fr.reconstruct()


@pytest.mark.gpu
def test_frame_api_end_to_end(gpu_ctx, tmp_path):
    pp = PicParams.make(40, 23)  # the reference's own 640x360 case (coded 640x368), synthetic syntax
    b = synth.generate(pp, 1, 360)
    fr = drive(b, gpu_ctx)
    path = tmp_path / "temp" / "yuv_frame"
    fr.write_to_yuv_file(str(path))
    data = np.fromfile(path, np.uint8)
    assert data.size == 353280
    assert np.array_equal(data, oracle.reconstruct(b)[0])
    y, cb, cr = fr.planes()
    assert y.shape == (368, 640) and cb.shape == (184, 320) and cr.shape == (184, 320)
