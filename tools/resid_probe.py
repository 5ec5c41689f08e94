#!/usr/bin/env python
"""Development probe: recon_residual_add_kernel (BASELINE configs[1]) on the bench workload — parity of the first and last
picture against the oracle, ms per launch (CUDA events), fraction of the HBM roofline. DRYV_RECON_LIB picks a variant."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import oracle
from dryv_b200 import recon, synth
from dryv_b200.abi import PicParams

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
qp = int(sys.argv[2]) if len(sys.argv) > 2 else 26
steps = 20
pp = PicParams.make(120, 68)
b = synth.generate(pp, frames, 3000, qp_base=qp)
ctx = recon.ReconContext(0)
ds = recon.DeviceSoa(b)
g = torch.Generator(device="cpu").manual_seed(1080)
pred = torch.randint(0, 256, (frames, pp.frame_bytes), dtype=torch.uint8, generator=g)
d_pred = pred.cuda()
d_out = torch.zeros_like(d_pred)
stream = torch.cuda.Stream()
torch.cuda.synchronize()
for _ in range(3):
    ctx.residual_add_device(ds, d_pred, d_out, stream.cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record(stream)
for _ in range(steps):
    ctx.residual_add_device(ds, d_pred, d_out, stream.cuda_stream)
e1.record(stream)
torch.cuda.synchronize()
ctx.wait()
ms = e0.elapsed_time(e1) / steps
got = d_out.cpu().numpy()
ok = True
for f in (0, frames - 1):
    ref = oracle.residual_add(b.frames(f, f + 1), pred[f:f + 1].numpy())
    ok = ok and bool(np.array_equal(ref, got[f:f + 1]))
n_mb = frames * pp.n_mb
gbs = n_mb * 1540 / (ms * 1e-3) / 1e9
print(json.dumps({"lib": os.environ.get("DRYV_RECON_LIB", "default"), "frames": frames, "qp": qp, "ms": round(ms, 4), "GB/s": round(gbs, 1),
                  "frac": round(gbs / 6467.1, 4), "parity_first_last": ok}))
