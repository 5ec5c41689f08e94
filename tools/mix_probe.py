"""Development: wavefront-kernel time for pure and mixed macroblock-class workloads (is the mixed case slower than
the weighted mean of the pure cases, i.e. does service-time variance between coupled rows cost throughput?).
usage: python tools/mix_probe.py [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dryv_b200 import recon, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pp = PicParams.make(120, 68)
ctx = recon.ReconContext(0)
res = {}
for name, p4, p8 in (("I4x4 only", 100, 0), ("I8x8 only", 0, 100), ("I16x16 only", 0, 0), ("40/25/35 mix", 40, 25)):
    b = synth.generate(pp, frames, 3000, pct_i4x4=p4, pct_i8x8=p8)
    ds = recon.DeviceSoa(b)
    d_out = torch.zeros((frames, pp.frame_bytes), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.reconstruct_device(ds, d_out)
    ctx.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(10):
        ctx.reconstruct_device(ds, d_out, s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    ctx.wait()
    res[name] = e0.elapsed_time(e1) / 10
    print(f"{name:14s} {res[name]:.4f} ms per step ({frames} x 1080p)")
w = 0.40 * res["I4x4 only"] + 0.25 * res["I8x8 only"] + 0.35 * res["I16x16 only"]
print(f"weighted mean of the pure cases {w:.4f} ms; mixed / weighted = {res['40/25/35 mix'] / w:.3f}")
