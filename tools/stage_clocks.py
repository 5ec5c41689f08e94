"""Development tool: build the CUDA library with -DDRYV_STAGE_CLOCKS into a scratch .so and print the
average cycles each warp role spends per macroblock in each stage. Run on a GPU box."""
import ctypes as C
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dryv_b200 import recon, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 64
so = os.path.join(recon.CSRC, "libdryv_recon_clocks.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + recon.NVCC_FLAGS + ["-DDRYV_STAGE_CLOCKS",
                      os.path.join(recon.CSRC, "recon.cu"), os.path.join(recon.CSRC, "recon_tables.cpp"),
                      os.path.join(recon.CSRC, "levels_pack.cpp"), os.path.join(recon.CSRC, "cabac_host.cpp"), os.path.join(recon.CSRC, "multi.cpp"), "-o", so])
recon.LIB_PATH = so
ctx = recon.ReconContext(0)
pp = PicParams.make(120, 68)
b = synth.generate(pp, frames, 3000)
ds = recon.DeviceSoa(b)
d_out = torch.zeros((frames, pp.frame_bytes), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.reconstruct_device(ds, d_out)
ctx.wait()
clk = (C.c_ulonglong * 16)()
ctx.lib.dryv_recon_debug_clocks.argtypes = [C.c_void_p, C.c_void_p]
ctx.lib.dryv_recon_debug_clocks(ctx.h, clk)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
ctx.reconstruct_device(ds, d_out)
ctx.wait()
e1.record()
torch.cuda.synchronize()
ctx.lib.dryv_recon_debug_clocks(ctx.h, clk)
nmb = frames * pp.n_mb
cls = np.where(b.mb_type != 0, 2, b.transform_size_8x8_flag)
n4, n8, n16 = [(cls == k).sum() for k in (0, 1, 2)]
print(f"frames {frames}, {nmb} MBs ({n4} I4x4, {n8} I8x8, {n16} I16x16)")
fn = ["headers (shuffles, masks)", "wait free slot (bar.sync, deferred)", "wait levels + residual", "wait chroma line (above)",
      "mode record + tap rows + hand-off", "chroma pred + store + carry", "issue level fetch (bulk copy)", "issue next headers/modes loads"]
ln = ["wait filled slot", "row start", "wait luma line (above)", "I4x4 pred (per I4x4 MB)", "I8x8 pred (per I8x8 MB)",
      "I16x16 pred (per I16 MB)", "-", "publish + store + carry"]
tot_f = sum(clk[:8]) / nmb
tot_l = sum(clk[8:]) / nmb
print("front warp: cycles per MB")
for i in range(8):
    print(f"  {fn[i]:32s} {clk[i] / nmb:9.1f}")
print(f"  total {tot_f:9.1f}")
print("pixel (luma) warp: cycles per MB")
for i in range(8):
    d = {3: n4, 4: n8, 5: n16}.get(i, nmb)
    print(f"  {ln[i]:32s} {clk[8 + i] / nmb:9.1f}   (per own MB: {clk[8 + i] / max(d, 1):9.1f})")
print(f"  total {tot_l:9.1f}")
