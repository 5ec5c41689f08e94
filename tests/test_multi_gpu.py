"""In-process multi-GPU dispatch (include/dryv_recon.h: dryv_recon_multi_*): one context and one host
thread per device, pictures dealt in contiguous blocks, outputs in disjoint slices of one buffer. The result must be
byte-identical to the single-device result and to the oracle. On a one-GPU box the dispatcher is still exercised with
two contexts on the same device; with two or more GPUs every visible device takes a share."""
import numpy as np
import pytest

import oracle
from dryv_b200 import recon, shard, synth
from dryv_b200.abi import PicParams

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_frames", [1, 5, 12])
def test_two_contexts_on_one_device_equal_one_context(n_frames):
    pp = PicParams.make(11, 6, 2, -2)
    b = synth.generate(pp, n_frames, 7100)
    ref = oracle.reconstruct(b)
    m = recon.MultiDeviceContext([0, 0, 0])
    assert m.n_devices == 3
    got = m.reconstruct(b)
    assert np.array_equal(got, ref)
    got_c = m.reconstruct(b, levels=recon.pack_levels(b.coeff))
    assert np.array_equal(got_c, ref)
    m.close()


def test_every_visible_device_takes_its_block():
    n_dev = _device_count()
    if n_dev < 2:
        pytest.skip("needs two or more GPUs")
    pp = PicParams.make(120, 68)
    n_frames = 4 * n_dev + 1   # uneven shares
    b = synth.generate(pp, n_frames, 7200)
    one = recon.ReconContext(0)
    ref = one.reconstruct(b)
    one.close()
    m = recon.MultiDeviceContext()
    assert m.n_devices == n_dev
    got = m.reconstruct(b)
    assert np.array_equal(got, ref)
    # every block is where shard.frames_for_rank puts it, first and last picture checked against the oracle
    for d in range(n_dev):
        r = shard.frames_for_rank(n_frames, d, n_dev)
        for f in (r.start, r.stop - 1):
            assert np.array_equal(got[f], oracle.reconstruct(b.frames(f, f + 1))[0]), (d, f)
    m.close()


def test_error_codes():
    lib = recon.load_library()
    assert lib.dryv_recon_multi_create(None, 0, None) == recon.ERR_ARG
    assert lib.dryv_recon_multi_device_count(None) == 0
    m = recon.MultiDeviceContext([0])
    pp = PicParams.make(2, 2)
    b = synth.generate(pp, 1, 1)
    b.mb_type[0] = 30   # not an intra macroblock type: reported by the device, named by the dispatcher
    with pytest.raises(recon.ReconError) as e:
        m.reconstruct(b)
    assert e.value.code == recon.ERR_UNSUPPORTED and "device 0" in str(e.value)
    m.close()
