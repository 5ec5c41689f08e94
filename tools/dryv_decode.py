#!/usr/bin/env python
"""The whole path on one MP4 (or Annex-B H.264) file, the way `dryv <file>` works (src/main.rs:34-51,
src/video/decoder.rs:87-150): demux the video track and CABAC-parse its IDR pictures on the CPU
(dryv_b200/csrc/cabac_host.cpp), reconstruct them on the GPU through the compact level stream, and write the first picture
to ./temp/yuv_frame in the reference's byte layout (src/video/frame/mod.rs:48-70).

    python tools/dryv_decode.py movie.mp4 [out_path]          # needs a B200
    python tools/dryv_decode.py --display [--nv12] movie.mp4 [out_path]   # the SPS display rectangle instead of the coded
                                                              # picture (the reference's open "frame cropping" item)
    python tools/dryv_decode.py --make-sample sample.mp4      # writes a 640x368 synthetic CABAC High-profile MP4 (no GPU)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--make-sample":
        from avc import stream
        from dryv_b200 import synth
        from dryv_b200.abi import PicParams
        b = synth.generate(PicParams.make(40, 23), 4, 360, standard_only=True)
        data = stream.encode_stream(b)
        if sys.argv[2].endswith(".mp4"):
            from avc import mp4
            data = mp4.mux(data, 640, 368)
        open(sys.argv[2], "wb").write(data)
        print(f"{sys.argv[2]}: {len(data)} bytes, 4 IDR pictures of 640x368")
        return 0
    from dryv_b200 import host, recon
    from dryv_b200.abi import SURFACE_NV12
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    display, nv12 = "--display" in sys.argv, "--nv12" in sys.argv
    data = open(args[0], "rb").read()
    out_path = args[1] if len(args) > 1 else "temp/yuv_frame"
    t0 = time.perf_counter()
    batch, levels = host.parse_compact(data)   # demux + CABAC parse straight into the compact level stream
    t1 = time.perf_counter()
    ctx = recon.ReconContext(0)
    if display or nv12:
        import numpy as np
        sf = host.surface(data)                # the rectangle the SPS asks for (the whole picture if it does not crop)
        if not display:
            sf.crop_left = sf.crop_top = 0
            sf.width, sf.height = 16 * batch.pp.pic_width_in_mbs, 16 * batch.pp.pic_height_in_mbs
        if nv12:
            sf.format = SURFACE_NV12
        ctx.set_surface(sf)
        frames = ctx.reconstruct_compact(batch, levels, np.empty((batch.n_frames, sf.nbytes), np.uint8))
    else:
        frames = ctx.reconstruct_compact(batch, levels)
    t2 = time.perf_counter()
    recon.write_yuv_file(frames[0], out_path)
    pp = batch.pp
    print(f"{batch.n_frames} IDR pictures of {pp.pic_width_in_mbs * 16}x{pp.pic_height_in_mbs * 16}: CABAC parse "
          f"{(t1 - t0) * 1e3:.1f} ms (CPU), reconstruction {(t2 - t1) * 1e3:.1f} ms incl. context creation (GPU); "
          f"first picture -> {out_path} ({frames[0].nbytes} bytes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
