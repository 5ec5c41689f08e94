"""Per-chunk stage timeline of dryv_recon_submit_compact / dryv_recon_submit (development aid).
usage: DRYV_SUBMIT_TRACE=1 python tools/submit_trace.py [frames] [compact|dense]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from dryv_b200 import recon, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "compact"
pp = PicParams.make(120, 68)
hb, owners = recon.pinned_batch(pp, n)
synth.generate(pp, n, 3000, qp_base=26, out=hb)
lv = recon.pack_levels(hb.coeff, pinned=True)
out = recon.PinnedArray((n, pp.frame_bytes), np.uint8)
ctx = recon.ReconContext(0)
os.environ.pop("DRYV_SUBMIT_TRACE", None)
for _ in range(3):
    ctx.submit_compact(hb, lv, out.array) if mode == "compact" else ctx.submit(hb, out.array)
    ctx.wait()
os.environ["DRYV_SUBMIT_TRACE"] = "1"
ctx.submit_compact(hb, lv, out.array) if mode == "compact" else ctx.submit(hb, out.array)
ctx.wait()
print("total ms", ctx.last_submit_ms)
