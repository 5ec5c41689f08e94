"""Development tool: build the CUDA library with -DDRYV_TRACE and print the wavefront timeline of picture 0:
per-row start lag, per-macroblock wait / work times. Run on a GPU box: python tools/trace_wave.py [frames]"""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dryv_b200 import recon, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
p4, p8 = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40, 25)
extra = sys.argv[4:]
so = os.path.join(recon.CSRC, "libdryv_recon_trace.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + recon.NVCC_FLAGS + ["-DDRYV_TRACE"] + extra +
                      [os.path.join(recon.CSRC, "recon.cu"), os.path.join(recon.CSRC, "recon_tables.cpp"), "-o", so])
recon.LIB_PATH = so
ctx = recon.ReconContext(0)
pp = PicParams.make(120, 68)
W, H = 120, 68
b = synth.generate(pp, frames, 3000, pct_i4x4=p4, pct_i8x8=p8)
ds = recon.DeviceSoa(b)
d_out = torch.zeros((frames, pp.frame_bytes), dtype=torch.uint8, device="cuda")
ctx.lib.dryv_recon_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
for _ in range(3):
    ctx.reconstruct_device(ds, d_out)
ctx.wait()
assert ctx.lib.dryv_recon_debug_trace(ctx.h, None, W * H + 1) == 0
ctx.reconstruct_device(ds, d_out)
ctx.wait()
trx = np.zeros((H * W + 1, 4), np.uint32)
assert ctx.lib.dryv_recon_debug_trace(ctx.h, trx.ctypes.data, W * H + 1) == 0
tr = trx[:H * W].reshape(H, W, 4)
m0, m1 = int(trx[H * W, 0]), int(trx[H * W, 1])
t = tr.astype(np.int64)
t0 = t[0, 0, 0]
t = (t - t0) & 0xffffffff
cls = np.where(b.mb_type[:W * H] != 0, 2, b.transform_size_8x8_flag[:W * H]).reshape(H, W)
print(f"mix I4x4 {p4}% I8x8 {p8}%")
print(f"mode pre-pass of picture 0: starts {((m0 - int(t0)) & 0xffffffff) - (1 << 32 if ((m0 - int(t0)) & 0xffffffff) > 1 << 31 else 0)} ns, "
      f"ends {((m1 - int(t0)) & 0xffffffff) - (1 << 32 if ((m1 - int(t0)) & 0xffffffff) > 1 << 31 else 0)} ns relative to the first macroblock of the wavefront")
print(f"frames {frames}: picture 0 spans {t[..., 3].max() / 1e3:.1f} us (first MB ready -> last MB done)")
start = t[:, 0, 0]
lag = np.diff(start)
print("row start lag (ns): mean %.0f  min %d  max %d" % (lag.mean(), lag.min(), lag.max()))
wait_lines = t[..., 1] - t[..., 0]
pred = t[..., 2] - t[..., 1]
tail = t[..., 3] - t[..., 2]
gap = np.zeros_like(wait_lines)
gap[:, 1:] = t[:, 1:, 0] - t[:, :-1, 3]
print("per MB (ns), mean: wait-lines %.0f  predict %.0f  store+publish %.0f  gap-to-next (slot wait + loop) %.0f  total %.0f"
      % (wait_lines.mean(), pred.mean(), tail.mean(), gap[:, 1:].mean(), (t[:, -1, 3] - t[:, 0, 0]).mean() / W))
for k, name in enumerate(("I4x4", "I8x8", "I16x16")):
    m = cls == k
    if m.any():
        print(f"  {name}: predict {pred[m].mean():.0f} ns  wait-lines {wait_lines[m].mean():.0f} ns")
print("fraction of MBs with wait-lines > 300 ns: %.2f ; > 1000 ns: %.2f" % ((wait_lines > 300).mean(), (wait_lines > 1000).mean()))
rows = [1, 2, 10, 30, 60]
for r in rows:
    print(f"row {r}: start {start[r]/1e3:.1f} us, end {t[r,-1,3]/1e3:.1f} us, duration {(t[r,-1,3]-start[r])/1e3:.1f} us, "
          f"wait-lines sum {wait_lines[r].sum()/1e3:.1f} us, lag behind row above at x=60: {(t[r,60,0]-t[r-1,60,0])/1e3:.2f} us")
# how far ahead is the row above when a MB starts: index of the last MB the row above has finished
ahead = np.zeros((H, W), np.int64)
for r in range(1, H):
    done_above = t[r - 1, :, 3]
    ahead[r] = np.searchsorted(done_above, t[r, :, 1], side="right") - np.arange(W)
print("row above is ahead by (MBs finished beyond x) at prediction start: mean %.2f, p10 %.0f, p50 %.0f, p90 %.0f"
      % (ahead[1:].mean(), np.percentile(ahead[1:], 10), np.percentile(ahead[1:], 50), np.percentile(ahead[1:], 90)))
