"""Development: CPU CABAC parse rate of dryv_cabac_parse (pictures per second, Mpixels/s, Mbit/s), one thread and all
threads, on a synthetic 640x368 stream. usage: python tools/host_parse_rate.py [pictures] [qp]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from avc import stream  # noqa: E402
from dryv_b200 import host, synth  # noqa: E402
from dryv_b200.abi import PicParams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
qp = int(sys.argv[2]) if len(sys.argv) > 2 else 26
pp = PicParams.make(40, 23)
b = synth.generate(pp, n, 360, qp_base=qp)
data = stream.encode_stream(b)
for name, fn in (("dense levels", host.parse), ("compact stream", host.parse_compact)):
    for threads in (1, os.cpu_count() or 1):
        best = 1e9
        for _ in range(5):
            t = time.perf_counter()
            fn(data, threads=threads)
            best = min(best, time.perf_counter() - t)
        print(f"{name:14s} {threads:3d} thread(s): {n / best:8.1f} pictures/s  {n * pp.luma_pixels / best / 1e6:8.1f} Mpixels/s  "
              f"{len(data) * 8 / best / 1e6:8.1f} Mbit/s  ({len(data) / n / 1024:.1f} KiB per picture at QP {qp})")
