"""Development: do back-to-back batches overlap usefully when they alternate over two contexts / streams (the tail of one
wavefront launch filled by the start-up stagger of the next)? usage: python tools/overlap_probe.py [frames] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dryv_b200 import recon, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pp = PicParams.make(120, 68)
b = synth.generate(pp, frames, 3000)
ds = recon.DeviceSoa(b)
NS = 4
ctxs = [recon.ReconContext(0)] * NS  # one context: the library rotates its control blocks over the launches
outs = [torch.zeros((frames, pp.frame_bytes), dtype=torch.uint8, device="cuda") for _ in range(NS)]
streams = [torch.cuda.Stream() for _ in range(NS)]


def run(n_streams):
    for k in range(2 * NS):
        ctxs[k % n_streams].reconstruct_device(ds, outs[k % n_streams], streams[k % n_streams].cuda_stream)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_streams)]
    e0.record(torch.cuda.current_stream())
    for s in streams[:n_streams]:
        s.wait_event(e0)
    for k in range(steps):
        ctxs[k % n_streams].reconstruct_device(ds, outs[k % n_streams], streams[k % n_streams].cuda_stream)
    for s, e in zip(streams[:n_streams], ends):
        e.record(s)
    torch.cuda.synchronize()
    for c in ctxs:
        c.wait()
    return max(e0.elapsed_time(e) for e in ends) / steps


res = [run(k) for k in range(1, NS + 1)]
print(f"{frames} x 1080p, {steps} steps, ms per step with 1..{NS} alternating streams: " + "  ".join(f"{r:.4f}" for r in res) +
      f"; outputs equal: {all(bool(torch.equal(outs[0], o)) for o in outs[1:])}")
