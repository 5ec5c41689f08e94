#!/usr/bin/env python
"""Development probe: deblock_wavefront_kernel on the bench workload (64 x 1080p, QP 26) — every timed launch filters a fresh
copy of the reconstructed pictures (the bench leg filters in place over and over), parity of the first and last picture
against oracle/deblock.py, ms per launch (CUDA events), fraction of the HBM roofline. DRYV_RECON_LIB picks a variant."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import oracle
from dryv_b200 import recon, synth
from dryv_b200.abi import PicParams
from oracle import deblock as dbl

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
qp = int(sys.argv[2]) if len(sys.argv) > 2 else 26
steps = 10
pp = PicParams.make(120, 68)
b = synth.generate(pp, frames, 3000, qp_base=qp)
ctx = recon.ReconContext(0)
ds = recon.DeviceSoa(b)
stream = torch.cuda.Stream()
sp = stream.cuda_stream
d_rec = torch.zeros((frames, pp.frame_bytes), dtype=torch.uint8, device="cuda")
ctx.reconstruct_device(ds, d_rec, sp)
ctx.wait()
torch.cuda.synchronize()
d = d_rec.clone()
ms = []
with torch.cuda.stream(stream):
    for i in range(steps + 2):
        d.copy_(d_rec)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.deblock_device(ds, d, 0, 0, sp)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1))
ctx.wait()
got = d.cpu().numpy()
rec = d_rec.cpu().numpy()
ok = True
for f in (0, frames - 1):
    sl = slice(f * pp.n_mb, (f + 1) * pp.n_mb)
    want = dbl.deblock(rec[f], pp.pic_width_in_mbs, pp.pic_height_in_mbs, b.qp[sl], b.transform_size_8x8_flag[sl], 0, 0, 0, 0)
    ok = ok and bool(np.array_equal(got[f], want))
t = float(np.median(ms))
gbs = frames * pp.n_mb * 770 / (t * 1e-3) / 1e9
print(json.dumps({"lib": os.environ.get("DRYV_RECON_LIB", "default"), "frames": frames, "qp": qp, "ms": round(t, 4),
                  "ms_min": round(min(ms), 4), "GB/s": round(gbs, 1), "frac": round(gbs / 6467.1, 4), "parity_first_last": ok}))
