"""GPU-vs-oracle diagnostic sweep (development tool): prints where the CUDA path and the CPU oracle
first differ, per macroblock class and stage. Run on a GPU box: python tools/gpu_diag.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oracle  # noqa: E402
from dryv_b200 import synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402
from dryv_b200.recon import DeviceSoa, ReconContext  # noqa: E402


def describe(pp, batch, ref, got, label):
    W, H = pp.pic_width_in_mbs, pp.pic_height_in_mbs
    nmb = pp.n_mb
    bad_total = 0
    first = None
    per_cls = {0: 0, 1: 0, 2: 0}
    for f in range(batch.n_frames):
        r, g = ref[f], got[f]
        ry, gy = r[:nmb * 256].reshape(H * 16, W * 16), g[:nmb * 256].reshape(H * 16, W * 16)
        rc, gc = r[nmb * 256:].reshape(2, H * 8, W * 8), g[nmb * 256:].reshape(2, H * 8, W * 8)
        dy = (ry != gy).reshape(H, 16, W, 16).any(axis=(1, 3))
        dc = (rc != gc).reshape(2, H, 8, W, 8).any(axis=(2, 4))
        bad = dy | dc[0] | dc[1]
        bad_total += int(bad.sum())
        if bad.any():
            ys, xs = np.nonzero(bad)
            for y, x in zip(ys, xs):
                a = f * nmb + y * W + x
                cls = 2 if batch.mb_type[a] else int(batch.transform_size_8x8_flag[a])
                per_cls[cls] += 1
            if first is None:
                y, x = ys[0], xs[0]
                a = f * nmb + y * W + x
                first = (f, x, y, int(batch.mb_type[a]), int(batch.transform_size_8x8_flag[a]),
                         int(batch.intra_chroma_pred_mode[a]), int(batch.qp[a]), bool(dy[y, x]), bool(dc[0][y, x]),
                         bool(dc[1][y, x]))
                print(f"  [{label}] first bad MB: frame {f} x {x} y {y} mb_type {first[3]} t8x8 {first[4]} "
                      f"chroma_mode {first[5]} qp {first[6]} luma_bad {first[7]} cb_bad {first[8]} cr_bad {first[9]}")
                print("   syntax:", batch.pred_syntax[a].tolist())
                if dy[y, x]:
                    print("   ref luma:\n", ry[y * 16:y * 16 + 16, x * 16:x * 16 + 16])
                    print("   got luma:\n", gy[y * 16:y * 16 + 16, x * 16:x * 16 + 16])
                for p in range(2):
                    if dc[p][y, x]:
                        print(f"   ref chroma{p}:\n", rc[p][y * 8:y * 8 + 8, x * 8:x * 8 + 8])
                        print(f"   got chroma{p}:\n", gc[p][y * 8:y * 8 + 8, x * 8:x * 8 + 8])
    total = batch.n_frames * nmb
    print(f"[{label}] bad MBs {bad_total}/{total}  by class I4x4/I8x8/I16x16 = {per_cls[0]}/{per_cls[1]}/{per_cls[2]}")
    return bad_total


def main():
    ctx = ReconContext(0)
    fails = 0
    # stage 1: residual-only kernel, per class
    for name, p4, p8 in (("res I4x4", 100, 0), ("res I8x8", 0, 100), ("res I16x16", 0, 0), ("res mixed", 40, 30)):
        for qp in (4, 26, 45):
            pp = PicParams.make(8, 5, cb_off=2, cr_off=-3)
            b = synth.generate(pp, 2, 100 + qp, qp_base=qp, pct_i4x4=p4, pct_i8x8=p8)
            rng = np.random.default_rng(qp)
            pred = rng.integers(0, 256, (2, pp.frame_bytes), dtype=np.uint8)
            ref = oracle.residual_add(b, pred)
            ds = DeviceSoa(b)
            d_pred = torch.from_numpy(pred).cuda()
            d_out = torch.zeros_like(d_pred)
            ctx.residual_add_device(ds, d_pred, d_out)
            ctx.wait()
            fails += describe(pp, b, ref, d_out.cpu().numpy(), f"{name} qp{qp}")
    # stage 2: prediction only (zero residual), per class, then with residual
    for zero in (True, False):
        for name, p4, p8 in (("I16x16", 0, 0), ("I4x4", 100, 0), ("I8x8", 0, 100), ("mixed", 40, 25)):
            pp = PicParams.make(11, 7, cb_off=1)
            b = synth.generate(pp, 3, 7, pct_i4x4=p4, pct_i8x8=p8, zero_residual=zero)
            ref = oracle.reconstruct(b)
            ds = DeviceSoa(b)
            d_out = torch.zeros((3, pp.frame_bytes), dtype=torch.uint8, device="cuda")
            ctx.reconstruct_device(ds, d_out)
            ctx.wait()
            fails += describe(pp, b, ref, d_out.cpu().numpy(), f"recon {name} zero_res={zero}")
    # stage 3: host path, bigger
    pp = PicParams.make(40, 23)
    b = synth.generate(pp, 5, 360)
    ref = oracle.reconstruct(b, threads=4)
    t = time.time()
    got = ctx.reconstruct(b)
    print("host path 5x(40x23) took", time.time() - t)
    fails += describe(pp, b, ref, got, "host 640x368 x5")
    pp = PicParams.make(120, 68)
    b = synth.generate(pp, 16, 3000)
    ref = oracle.reconstruct(b, threads=8)
    ds = DeviceSoa(b)
    d_out = torch.zeros((16, pp.frame_bytes), dtype=torch.uint8, device="cuda")
    for it in range(3):
        torch.cuda.synchronize()
        t = time.time()
        ctx.reconstruct_device(ds, d_out)
        ctx.wait()
        dt = time.time() - t
        print(f"1080p x16 device path: {dt*1e3:.3f} ms  -> {16*pp.luma_pixels/dt/1e6:.0f} Mpx/s")
    fails += describe(pp, b, ref, d_out.cpu().numpy(), "1080p x16")
    print("TOTAL BAD", fails)
    return 0 if fails == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
