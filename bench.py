#!/usr/bin/env python
"""bench.py — IDR reconstruct Mpixels/s on N B200s (one process per GPU, frame-sharded, no collective).

Workload (BASELINE.json configs[2], the configuration the metric is quoted on for one GPU): 1080p
(120x68 MB) full intra reconstruction, mixed Intra4x4/8x8/16x16 + chroma, batch of 64 IDR pictures per
GPU, seeded synthetic spec-legal syntax buffers (dryv_b200/synth). Weak scaling: every rank gets its
own 64 pictures.

  python bench.py --gpus 1 --steps 10 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # CPU path (oracle port of the Rust reference) on host cores

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import numpy as np  # noqa: E402

METRIC = "idr_reconstruct_mpixels_per_s"
UNIT = "Mpixels/s"
BYTES_PER_MB_FULL = 1172   # 768 levels + 20 syntax + 384 pixels out (SURVEY.md §8d / BASELINE.md §3)
BYTES_PER_MB_RESID = 1540  # 768 + 4 + 384 prediction in + 384 out
# ncu counters of one recon_wavefront_kernel launch on the default workload (64 x 1080p), read from the newest committed
# capture summary (profiles/rNN_wavefront_counters.json: dram bytes, warp instructions); None when there is none
def profile_counters():
    import glob
    best = None
    for p in sorted(glob.glob(os.path.join(_ROOT, "profiles", "r*_wavefront_counters.json"))):
        try:
            with open(p) as f:
                best = dict(json.load(f), source=os.path.relpath(p, _ROOT))
        except Exception:
            pass
    return best


def issue_peak():
    """Measured issue rate of a balanced ALU + FMA-pipe integer stream (tools/micro/int_issue.cu, committed output):
    T warp-instructions/s for the whole chip, the roof the kernels' instruction counts are read against."""
    try:
        with open(os.path.join(_ROOT, "profiles", "r02_int_issue.txt")) as f:
            vals = [float(l.split("=>")[1].split("T")[0]) for l in f if l.startswith("mix IADD+IMAD")]
        return max(vals), "profiles/r02_int_issue.txt (mix IADD+IMAD)"
    except Exception:
        return None, None
# independent batches kept in flight by the device-resident timed region: consecutive steps rotate over this many CUDA
# streams and output buffers (the library's four wavefront control blocks allow up to four)
N_FLIGHT = 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="IDR pictures per GPU per step")
    ap.add_argument("--width-mbs", type=int, default=120)
    ap.add_argument("--height-mbs", type=int, default=68)
    ap.add_argument("--qp", type=int, default=26)
    ap.add_argument("--seed", type=int, default=3000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the side measurements (configs[1], [3], [4], export, deblock)")
    ap.add_argument("--config4-frames", type=int, default=256, help="BASELINE configs[3]: 2160p pictures in total, over all GPUs")
    ap.add_argument("--config5-streams", type=int, default=32, help="BASELINE configs[4]: independent 1080p streams (QP 10..45)")
    ap.add_argument("--config5-frames", type=int, default=8, help="IDR pictures per stream")
    return ap.parse_args()


def measured_peak_gbs():
    p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def workload_name(args):
    return (f"{args.width_mbs * 16}x{args.height_mbs * 16} full intra reconstruct (mixed Intra4x4/8x8/16x16 + chroma, "
            f"MB wavefront), batch of {args.frames} IDR frames per GPU (BASELINE.json configs[2])")


def config_dict(args, n_gpus):
    return {
        "workload": workload_name(args),
        "pic_width_in_mbs": args.width_mbs, "pic_height_in_mbs": args.height_mbs,
        "frames_per_gpu": args.frames, "qp_base": args.qp, "seed": args.seed,
        "mb_mix": "40% Intra4x4 / 25% Intra8x8 / 35% Intra16x16, 10% stress MBs",
        "sharding": f"frames x{n_gpus} (independent IDR pictures per GPU, no collective)",
        "l2": "per-step working set (levels+syntax in, pictures out) is larger than the 126 MB L2, no flush needed",
        "step_overlap": f"GPU arm, device-resident value: consecutive steps rotate over {N_FLIGHT} CUDA streams and output "
                        "buffers, so independent batches overlap; single_stream is one batch at a time",
    }


def cpu_oracle_run(batch, threads):
    import oracle
    t0 = time.perf_counter()
    oracle.reconstruct(batch, threads=threads)
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path. dryv is Rust and cannot be built in
    this image (no cargo/rustc), so this times the C oracle port (oracle/) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from dryv_b200 import synth
    from dryv_b200.abi import PicParams
    cores = os.cpu_count() or 1
    pp = PicParams.make(args.width_mbs, args.height_mbs)
    n = args.frames if cores >= 8 else max(1, min(args.frames, 2 * cores))
    batch = synth.generate(pp, n, args.seed, qp_base=args.qp)
    for _ in range(args.warmup):
        cpu_oracle_run(batch, cores)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_oracle_run(batch, cores)
    per_step = t / max(1, args.steps)
    value = n * pp.luma_pixels / per_step / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "i64 (oracle port, isize in the reference)",
        "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} of {args.frames} pictures per step, one picture per host thread, "
                                   f"{cores} threads (C oracle port of dryv's Rust frame/ path; the Rust reference "
                                   "itself cannot be built here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0



def _device_soa_of(pp, seeds, dev, qps=None, chunk=32):
    """Generate pictures seeded `seeds` (one synth call each, so the seeds need not be consecutive) straight into
    device-resident SoA buffers, `chunk` pictures of host memory at a time. Returns (DeviceSoa, first SyntaxBatch picture,
    last SyntaxBatch picture) - the two host copies are kept for the parity check."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from dryv_b200 import recon, synth
    from dryv_b200.abi import FIELDS, SyntaxBatch
    n = len(seeds)
    tens = None
    first = last = None
    pool = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1))
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        hb = SyntaxBatch.empty(pp, c1 - c0)
        jobs = [pool.submit(synth.generate, pp, 1, seeds[i], qp_base=(qps[i] if qps else 26), threads=1,
                            out=hb.frames(i - c0, i - c0 + 1)) for i in range(c0, c1)]
        for j in jobs:
            j.result()
        if tens is None:
            tens = {f: torch.empty((n * pp.n_mb,) + getattr(hb, f).shape[1:], dtype=torch.from_numpy(getattr(hb, f)).dtype,
                                   device=dev) for f in FIELDS}
        for f in FIELDS:
            tens[f][c0 * pp.n_mb:c1 * pp.n_mb].copy_(torch.from_numpy(getattr(hb, f)))
        if c0 == 0:
            first = hb.frames(0, 1).copy()
        if c1 == n:
            last = hb.frames(c1 - c0 - 1, c1 - c0).copy()
    pool.shutdown()
    d = object.__new__(recon.DeviceSoa)
    d.pp, d.n_frames, d.tensors = pp, n, tens
    return d, first, last


def run_config4(args, ctx, dev, rank, world, barrier, streams):
    """BASELINE configs[3]: 2160p intra reconstruct, `--config4-frames` (256) IDR pictures in total, picture f on GPU
    f mod N (seed 4000 + f): strong scaling, no collective. Returns this rank's (ms per pass, pictures, parity)."""
    import torch
    import oracle
    from dryv_b200.abi import PicParams
    pp = PicParams.make(240, 135)
    mine = list(range(rank, args.config4_frames, world))
    if not mine:
        return 0.0, 0, True
    dsoa, first, last = _device_soa_of(pp, [4000 + f for f in mine], dev, chunk=16)
    d_out = torch.zeros((len(mine), pp.frame_bytes), dtype=torch.uint8, device=dev)
    sub = 32   # pictures per launch; launches rotate over the streams so that consecutive ones overlap

    def one_pass():
        for i, lo in enumerate(range(0, len(mine), sub)):
            hi = min(len(mine), lo + sub)
            ctx.reconstruct_device(dsoa.frames(lo, hi), d_out[lo:hi], streams[i % len(streams)].cuda_stream)

    one_pass()
    ctx.wait()
    reps = max(2, min(args.steps, 5))
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
    barrier()
    e0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_event(e0)
    for _ in range(reps):
        one_pass()
    for e, s_ in zip(ends, streams):
        e.record(s_)
    barrier()
    ctx.wait()
    ms = max(e0.elapsed_time(e) for e in ends) / reps
    ok = bool(np.array_equal(d_out[0].cpu().numpy(), oracle.reconstruct(first)[0])) and \
        bool(np.array_equal(d_out[len(mine) - 1].cpu().numpy(), oracle.reconstruct(last)[0]))
    return ms, len(mine), ok


def config5_qp(s, n_streams):
    return 10 + (35 * s) // max(1, n_streams - 1)


def run_config5(args, ctx, dev, rank, world, barrier, streams):
    """BASELINE configs[4]: `--config5-streams` (32) independent 1080p IDR streams with QP 10..45 (dense -> sparse levels),
    stream s on GPU s mod N, `--config5-frames` pictures each (seed 5000 + 64 s + k); every stream is reconstructed as a
    batch of its own, launches of different streams overlap. Returns (ms per pass over this rank's streams, pictures,
    parity, {qp: isolated ms of that stream's batch} for QP 10 / 26-ish / 45 when the rank owns them)."""
    import torch
    import oracle
    from dryv_b200.abi import PicParams
    pp = PicParams.make(120, 68)
    ns, k = args.config5_streams, args.config5_frames
    mine = list(range(rank, ns, world))
    if not mine:
        return 0.0, 0, True, {}, None
    seeds, qps = [], []
    for s_ in mine:
        for j in range(k):
            seeds.append(5000 + 64 * s_ + j)
            qps.append(config5_qp(s_, ns))
    dsoa, first, last = _device_soa_of(pp, seeds, dev, qps=qps, chunk=64)
    d_out = torch.zeros((len(seeds), pp.frame_bytes), dtype=torch.uint8, device=dev)

    def launch(i, stream):
        ctx.reconstruct_device(dsoa.frames(i * k, (i + 1) * k), d_out[i * k:(i + 1) * k], stream.cuda_stream)

    def one_pass():
        for i in range(len(mine)):
            launch(i, streams[i % len(streams)])

    one_pass()
    ctx.wait()
    reps = max(2, min(args.steps, 5))
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
    barrier()
    e0.record(streams[0])
    for s_ in streams[1:]:
        s_.wait_event(e0)
    for _ in range(reps):
        one_pass()
    for e, s_ in zip(ends, streams):
        e.record(s_)
    barrier()
    ctx.wait()
    ms = max(e0.elapsed_time(e) for e in ends) / reps
    ok = bool(np.array_equal(d_out[0].cpu().numpy(), oracle.reconstruct(first)[0])) and \
        bool(np.array_equal(d_out[len(seeds) - 1].cpu().numpy(), oracle.reconstruct(last)[0]))
    per_qp = {}
    want = {config5_qp(0, ns), config5_qp(ns - 1, ns), min((config5_qp(s_, ns) for s_ in range(ns)), key=lambda q: abs(q - 26))}
    for i, s_ in enumerate(mine):
        q = config5_qp(s_, ns)
        if q in want and q not in per_qp:
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            a0.record(streams[0])
            for _ in range(5):
                launch(i, streams[0])
            a1.record(streams[0])
            torch.cuda.synchronize(dev)
            ctx.wait()
            per_qp[q] = a0.elapsed_time(a1) / 5
    # the same pictures as ONE batch (streams are independent IDR pictures of one geometry, so a caller that holds several
    # streams can hand them over together): what the launch-per-stream figure leaves on the table
    merged = None
    if len(mine) > 1:
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.reconstruct_device(dsoa, d_out, streams[0].cuda_stream)
        torch.cuda.synchronize(dev)
        m0.record(streams[0])
        for _ in range(reps):
            ctx.reconstruct_device(dsoa, d_out, streams[0].cuda_stream)
        m1.record(streams[0])
        torch.cuda.synchronize(dev)
        ctx.wait()
        merged = m0.elapsed_time(m1) / reps
        ok = ok and bool(np.array_equal(d_out[len(seeds) - 1].cpu().numpy(), oracle.reconstruct(last)[0]))
    return ms, len(seeds), ok, per_qp, merged


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dryv_b200 import recon, synth
    from dryv_b200.abi import PicParams

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the reconstruction path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    pp = PicParams.make(args.width_mbs, args.height_mbs)
    n_frames = args.frames
    # pinned host buffers (the e2e path copies from these every step)
    hbatch, owners = recon.pinned_batch(pp, n_frames)
    synth.generate(pp, n_frames, args.seed + rank * n_frames, qp_base=args.qp, out=hbatch)
    hout = recon.PinnedArray((n_frames, pp.frame_bytes), np.uint8)

    ctx = recon.ReconContext(local_rank)
    dsoa = recon.DeviceSoa(hbatch, device=dev)
    d_out = torch.zeros((n_frames, pp.frame_bytes), dtype=torch.uint8, device=dev)
    d_outs = [d_out] + [torch.zeros_like(d_out) for _ in range(N_FLIGHT - 1)]
    # Dedicated (non-default) torch streams: kernels are launched on them through the C ABI and the torch.cuda.Event
    # objects below are recorded on the same streams. Consecutive steps rotate over the N_FLIGHT streams and as many output
    # buffers: the batches are independent, so the library lets a step start while the previous one drains (the start-up
    # stagger of one wavefront fills the tail of the other). Every step still does all of its work.
    streams = [torch.cuda.Stream(dev) for _ in range(N_FLIGHT)]
    stream = streams[0]
    sptr = stream.cuda_stream
    assert all(s.cuda_stream != 0 for s in streams)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step(k=0, n_streams=N_FLIGHT):
        i = k % n_streams
        ctx.reconstruct_device(dsoa, d_outs[i], streams[i].cuda_stream)

    for k in range(max(3, args.warmup)):
        step(k)
    ctx.wait()

    # one batch at a time on one stream: the isolated duration of the dominant kernel (CUDA events recorded around each
    # launch on the launching stream) and the serial step time, reported beside the headline
    iso_steps = max(3, min(args.steps, 10))
    e0s, e1s = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0s.record(stream)
    for _ in range(iso_steps):
        step(0, 1)
    e1s.record(stream)
    barrier()
    ctx.wait()
    ms_serial_step = e0s.elapsed_time(e1s) / iso_steps
    wave_ms = ctx.wavefront_times_ms(min(64, iso_steps))
    wave_ms_isolated = sum(wave_ms) / max(1, len(wave_ms))

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = ctx.launch_count
    ev0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(N_FLIGHT)]
    barrier()
    ev0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(ev0)
    for k in range(args.steps):
        step(k)
    for e, s in zip(ends, streams):
        e.record(s)
    barrier()
    ctx.wait()
    ms_total = max(ev0.elapsed_time(e) for e in ends)
    launches = ctx.launch_count - launches0
    # effective time of the dominant kernel per launch inside the timed region (two launches share the GPU at a time, so
    # the event pair around one launch spans more than its share): the timed region is nothing but these launches
    wave_ms_avg = ms_total / args.steps

    # ---- e2e through the host-buffer C ABI calls: pinned H2D + kernels + D2H inside the timed region.
    # Two wire formats for the levels: the compact stream (significance masks + non-zero levels, the form CABAC
    # produces them in; dryv_recon_submit_compact) is the headline, the dense int16 arrays (dryv_recon_submit) are
    # reported beside it. Both are packed/generated once, outside the timed region, into pinned memory.
    e2e_ms = e2e_dense_ms = None
    levels = None
    if not args.no_e2e:
        levels = recon.pack_levels(hbatch.coeff, pinned=True)

        hout2 = recon.PinnedArray((n_frames, pp.frame_bytes), np.uint8)
        outs = [hout.array, hout2.array]

        def timed(fn):
            """K steps queued back to back (streaming use of the ABI: the next step's H2D runs under this step's D2H,
            dryv_recon_wait_oldest hands each step's pictures to the host as they complete); every step copies its
            inputs from pinned host memory and its pictures back. Returns (ms per step, ms of one isolated step)."""
            for k in range(2):
                fn(outs[k & 1])
                ctx.wait()
            single = ctx.last_submit_ms
            barrier()
            for k in range(args.steps):
                fn(outs[k & 1])
                if k >= 1:
                    ctx.wait_oldest()   # step k-1 is complete in outs[(k-1) & 1] before step k+1 reuses that buffer
            ctx.wait()
            per_step = ctx.last_submit_ms / args.steps   # CUDA events: first H2D of the first step .. last D2H of the last
            barrier()
            return per_step, single

        e2e_ms, e2e_single_ms = timed(lambda o: ctx.submit_compact(hbatch, levels, o))
        e2e_out0 = outs[(args.steps - 1) & 1][0].copy()
        e2e_dense_ms, e2e_dense_single_ms = timed(lambda o: ctx.submit(hbatch, o))

        def copy_floor(h2d_bytes, d2h_bytes):
            """The same bytes as one e2e step, copied by plain cudaMemcpyAsync in both directions at once (pinned host
            memory, two streams, all ranks together): what the link alone allows, measured in this run."""
            src_pin = recon.PinnedArray(h2d_bytes, np.uint8)
            src = torch.from_numpy(src_pin.array)
            dst_h = torch.from_numpy(hout.array.reshape(-1))[:d2h_bytes]
            d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
            d_src = d_out.reshape(-1)[:d2h_bytes]
            sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            f0, fa, fb = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            reps = max(3, min(args.steps, 10))
            barrier()
            f0.record(sa)
            sb.wait_event(f0)
            for _ in range(reps):
                with torch.cuda.stream(sa):
                    d_in.copy_(src, non_blocking=True)
                with torch.cuda.stream(sb):
                    dst_h.copy_(d_src, non_blocking=True)
            fa.record(sa)
            fb.record(sb)
            barrier()
            return max(f0.elapsed_time(fa), f0.elapsed_time(fb)) / reps

        floor_ms = copy_floor(int(levels.nbytes + (hbatch.input_bytes - hbatch.coeff.nbytes)), int(hout.array.nbytes))
        floor_dense_ms = copy_floor(int(hbatch.input_bytes), int(hout.array.nbytes))
    clocks = sampler.stop()

    # parity on every rank, outside the timed region: the rank's first and last picture against the oracle, and the
    # output buffers of the batches in flight against each other; AND-reduced over the ranks
    import oracle
    last_buf = d_outs[(args.steps - 1) % N_FLIGHT]
    got0, got_last = last_buf[0].cpu().numpy(), last_buf[n_frames - 1].cpu().numpy()
    ref0 = oracle.reconstruct(hbatch.frames(0, 1))[0]
    ref_last = oracle.reconstruct(hbatch.frames(n_frames - 1, n_frames))[0]
    parity = bool(np.array_equal(got0, ref0)) and bool(np.array_equal(got_last, ref_last)) and \
        all(bool(torch.equal(d_outs[0], o)) for o in d_outs[1:])

    # BASELINE configs[3] and configs[4] (all ranks take part; rank 0 reports)
    c4 = c5 = None
    if not args.no_extra:
        c4 = run_config4(args, ctx, dev, rank, world, barrier, streams)
        c5 = run_config5(args, ctx, dev, rank, world, barrier, streams)

    t = torch.tensor([ms_total, e2e_ms if e2e_ms is not None else 0.0, wave_ms_avg,
                      e2e_dense_ms if e2e_dense_ms is not None else 0.0,
                      c4[0] if c4 else 0.0, c5[0] if c5 else 0.0,
                      floor_ms if e2e_ms is not None else 0.0, floor_dense_ms if e2e_ms is not None else 0.0,
                      (c5[4] or 0.0) if c5 else 0.0],
                     dtype=torch.float64, device=dev)
    flags = torch.tensor([1 if parity else 0, 1 if (c4 is None or c4[2]) else 0, 1 if (c5 is None or c5[2]) else 0],
                         dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ms_total, e2e_ms_max, wave_ms_avg, e2e_dense_ms_max = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    c4_ms, c5_ms, floor_ms_max, floor_dense_ms_max = float(t[4]), float(t[5]), float(t[6]), float(t[7])
    c5_merged_ms = float(t[8])
    parity_all, c4_ok, c5_ok = bool(flags[0]), bool(flags[1]), bool(flags[2])
    ms_per_step = ms_total / args.steps
    total_px = world * n_frames * pp.luma_pixels
    value = total_px / (ms_per_step * 1e-3) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_gbs()
    n_mb_step = n_frames * pp.n_mb
    # a step = ticket memset + resolve_modes_kernel (prediction-mode pre-pass) + recon_wavefront_kernel. The wavefront
    # kernel is a programmatic dependent of the pre-pass and overlaps it; with N_FLIGHT batches in flight the GPU time one
    # launch costs is the timed region divided by the launches in it (kernel_ms); kernel_ms_isolated is the CUDA-event
    # duration of a launch that has the GPU to itself, quoted against the wavefront kernel's algorithmic bytes as well
    kernel_s = wave_ms_avg * 1e-3
    achieved = n_mb_step * BYTES_PER_MB_FULL / kernel_s / 1e9
    counters = profile_counters()
    default_workload = (n_frames, args.width_mbs, args.height_mbs, args.qp) == (64, 120, 68, 26)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32 (int16 levels -> int32 arithmetic -> u8 pixels)", "data": "synthetic",
        "config": config_dict(args, world),
        "clocks": clocks,
        "gpu_launches": int(launches),
        "batches_in_flight": N_FLIGHT,
        "single_stream": {"ms_per_step": ms_serial_step,
                          "value": n_frames * pp.luma_pixels / (ms_serial_step * 1e-3) / 1e6,
                          "note": "rank 0, one batch at a time on one stream (no overlap between consecutive batches)"},
        "parity_vs_oracle_first_picture": parity,
        "parity_all_ranks_first_and_last_picture": parity_all,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (counters or {}).get("dram_bytes") if default_workload else None,
                     "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                       f"({(counters or {}).get('source')}); algorithmic bytes per launch = "
                                       f"{n_mb_step * BYTES_PER_MB_FULL}",
                     "peak_source": peak_src, "kernel": "dryv::recon_wavefront_kernel",
                     "kernel_ms": wave_ms_avg, "kernel_share_of_step": wave_ms_avg / ms_per_step,
                     "kernel_ms_isolated": wave_ms_isolated,
                     "frac_isolated": n_mb_step * BYTES_PER_MB_FULL / (wave_ms_isolated * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_mb": BYTES_PER_MB_FULL, "mbs_per_launch": n_mb_step},
    }
    # the wavefront occupancy bound: a picture is a chain of W + 2(H - 1) dependent macroblock steps, and
    # the kernel runs min(rows, SMs x 10 resident row teams) rows at a time
    teams = min(n_frames * args.height_mbs, torch.cuda.get_device_properties(dev).multi_processor_count * 10)
    steps = args.width_mbs + 2 * (args.height_mbs - 1)
    mb_period_us = wave_ms_isolated * 1e3 * teams / n_mb_step   # from a launch that has the GPU to itself
    line["wavefront"] = {
        "dependent_steps_per_picture": steps, "row_teams": teams,
        "rows_per_team": n_frames * args.height_mbs / teams,
        "mb_period_us_per_team": mb_period_us,
        "critical_path_ms_at_that_period": steps * mb_period_us * 1e-3,
        "note": "isolated kernel time = rows_per_team x pic_width_in_mbs x mb_period (+ start-up stagger and tail, which "
                "back-to-back batches on separate streams hide); the period is the serial instruction stream of a row team's "
                "slower warp (DESIGN.md §5), ~25x what the HBM roofline would allow",
    }
    # The kernel is bound by instruction issue, not by HBM: its warp-instruction count (ncu) against the measured issue
    # rate of a balanced integer stream (tools/micro/int_issue.cu)
    ipeak, ipeak_src = issue_peak()
    if default_workload and counters and counters.get("warp_instructions") and ipeak:
        wi = float(counters["warp_instructions"])
        line["roofline_alu"] = {
            "bound": "issue", "warp_instructions_per_launch": wi, "warp_instructions_per_mb": wi / n_mb_step,
            "achieved": wi / kernel_s / 1e12, "peak": ipeak, "unit": "T warp-instr/s", "frac": wi / kernel_s / 1e12 / ipeak,
            "frac_isolated": wi / (wave_ms_isolated * 1e-3) / 1e12 / ipeak,
            "counter_source": counters.get("source"), "peak_source": ipeak_src,
            "warp_instructions_per_mb_at_hbm_target": ipeak * 1e12 / (0.6 * peak * 1e9 / BYTES_PER_MB_FULL),
            "note": "what the instruction stream would have to shrink to for the 60 % HBM target at a 100 % issue rate"}
    if c4 is not None:
        pp4_px = 240 * 135 * 256
        line["config4_2160p"] = {
            "workload": f"3840x2160 full intra reconstruct, {args.config4_frames} IDR pictures in total, picture f on GPU f mod N "
                        "(seed 4000 + f), 32 pictures per launch (BASELINE.json configs[3])",
            "scaling": "strong", "n_gpus": world, "pictures_total": args.config4_frames, "ms_per_pass": c4_ms,
            "value": args.config4_frames * pp4_px / (c4_ms * 1e-3) / 1e6 if c4_ms > 0 else None, "unit": UNIT,
            "per_gpu_value": args.config4_frames * pp4_px / (c4_ms * 1e-3) / 1e6 / world if c4_ms > 0 else None,
            "roofline_frac_per_gpu": (args.config4_frames * 240 * 135 * BYTES_PER_MB_FULL / world / (c4_ms * 1e-3) / 1e9 / peak) if c4_ms > 0 else None,
            "parity_all_ranks_first_and_last_picture": c4_ok}
    if c5 is not None:
        ns5, k5 = args.config5_streams, args.config5_frames
        px5 = ns5 * k5 * 120 * 68 * 256
        line["config5_qp_sweep"] = {
            "workload": f"{ns5} independent 1920x1088 IDR streams, QP {config5_qp(0, ns5)}..{config5_qp(ns5 - 1, ns5)} (dense -> "
                        f"sparse levels), stream s on GPU s mod N, {k5} pictures per stream (seed 5000 + 64 s + k), one launch per "
                        "stream, launches of different streams overlap (BASELINE.json configs[4])",
            "n_gpus": world, "ms_per_pass": c5_ms, "value": px5 / (c5_ms * 1e-3) / 1e6 if c5_ms > 0 else None, "unit": UNIT,
            "rank0_isolated_stream": {f"qp{q}": {"ms_per_batch": ms_, "value": k5 * 120 * 68 * 256 / (ms_ * 1e-3) / 1e6}
                                      for q, ms_ in sorted(c5[3].items())},
            "merged_batch": ({"ms_per_pass": c5_merged_ms, "value": px5 / (c5_merged_ms * 1e-3) / 1e6, "unit": UNIT,
                              "note": "each rank's streams handed over as one batch (one launch per rank) instead of one "
                                      "launch per stream; max over ranks"} if c5_merged_ms > 0 else None),
            "parity_all_ranks_first_and_last_picture": c5_ok}
    if e2e_ms is not None:
        syntax_bytes = int(hbatch.input_bytes - hbatch.coeff.nbytes)
        line["e2e"] = {"value": total_px / (e2e_ms_max * 1e-3) / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": int(levels.nbytes + syntax_bytes),
                       "d2h_bytes_per_step": int(hout.array.nbytes), "ms_per_step": e2e_ms_max,
                       "single_step_ms": e2e_single_ms,
                       "pcie_floor_ms": floor_ms_max, "fraction_of_pcie_floor": floor_ms_max / e2e_ms_max,
                       "pcie_floor_how": "the step's H2D and D2H bytes moved by plain cudaMemcpyAsync on two streams at once, "
                                         "pinned host memory, all ranks together, same run (max over ranks)",
                       "levels_wire_format": "compact (per MB: coded-slot mask, 16-bit significance masks, non-zero "
                                             "levels as int8/int16; include/dryv_recon.h dryv_mb_levels_compact)",
                       "parity_vs_oracle_first_picture": bool(np.array_equal(e2e_out0, ref0)),
                       "how": "K x dryv_recon_submit_compact queued back to back on pinned host buffers (two output buffers "
                              "alternate, dryv_recon_wait_oldest per step, dryv_recon_wait at the end), CUDA-event timed from "
                              "the first H2D to the last D2H, divided by K; single_step_ms is one isolated submit + wait; the "
                              "stream is packed once outside the timed region, like the dense arrays are generated"}
        line["e2e_dense"] = {"value": total_px / (e2e_dense_ms_max * 1e-3) / 1e6, "unit": UNIT,
                             "h2d_bytes_per_step": int(hbatch.input_bytes),
                             "d2h_bytes_per_step": int(hout.array.nbytes), "ms_per_step": e2e_dense_ms_max,
                             "single_step_ms": e2e_dense_single_ms,
                             "pcie_floor_ms": floor_dense_ms_max, "fraction_of_pcie_floor": floor_dense_ms_max / e2e_dense_ms_max,
                             "how": "same with dryv_recon_submit (dense int16 levels)"}

    if not args.no_extra:
        # side measurement, BASELINE.json configs[1]: dequant + IDCT + residual add only
        g = torch.Generator(device="cpu").manual_seed(1080)
        d_pred = torch.randint(0, 256, d_out.shape, dtype=torch.uint8, generator=g).to(dev)
        for _ in range(3):
            ctx.residual_add_device(dsoa, d_pred, d_out, sptr)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a0.record(stream)
        for _ in range(args.steps):
            ctx.residual_add_device(dsoa, d_pred, d_out, sptr)
        a1.record(stream)
        torch.cuda.synchronize(dev)
        ctx.wait()
        rms = a0.elapsed_time(a1) / args.steps
        # configs[1] as written is ONE 1080p picture: its single-launch latency, rotating over distinct pictures of the
        # batch so that every launch streams its 12.6 MB from HBM rather than from L2
        rot = min(32, n_frames)
        views = [(dsoa.frames(f, f + 1), d_pred[f:f + 1], d_out[f:f + 1]) for f in range(rot)]
        for v, p_, o_ in views[:3]:
            ctx.residual_add_device(v, p_, o_, sptr)
        torch.cuda.synchronize(dev)
        a0.record(stream)
        for v, p_, o_ in views:
            ctx.residual_add_device(v, p_, o_, sptr)
        a1.record(stream)
        torch.cuda.synchronize(dev)
        ctx.wait()
        one_us = a0.elapsed_time(a1) / rot * 1e3
        rach = n_mb_step * BYTES_PER_MB_RESID / (rms * 1e-3) / 1e9
        line["residual_only"] = {"workload": "dequant + 4x4/8x8 IDCT + residual add only (BASELINE.json configs[1] "
                                             f"kernel, same {n_frames}-picture buffers so it streams from HBM)",
                                 "value": n_frames * pp.luma_pixels / (rms * 1e-3) / 1e6, "unit": UNIT,
                                 "ms_per_step": rms,
                                 "single_picture_launch_us": one_us,
                                 "single_picture_mpixels_per_s": pp.luma_pixels / one_us,
                                 "roofline": {"bound": "hbm", "achieved": rach, "peak": peak, "unit": "GB/s",
                                              "frac": rach / peak, "algorithmic_bytes_per_mb": BYTES_PER_MB_RESID}}

        # side measurement, SURVEY.md §8(f) next-3: the display rectangle of every picture of the batch (the SPS crop of a
        # 1080p / 2160p stream: the coded height rounded down to a multiple of 8 lines below it) as I420 planes and as NV12;
        # a streaming pass, checked against oracle/surface.py on the first picture
        from dryv_b200.abi import SURFACE_I420, SURFACE_NV12, Surface
        from oracle import surface as osurf
        W_, H_ = 16 * pp.pic_width_in_mbs, 16 * pp.pic_height_in_mbs
        h_disp = H_ - 8 if H_ > 16 else H_
        exp = {}
        for name, fmt in (("i420", SURFACE_I420), ("nv12", SURFACE_NV12)):
            sf = Surface.make(W_, h_disp, 0, 0, fmt)
            d_sf = torch.empty((n_frames, sf.nbytes), dtype=torch.uint8, device=dev)
            for _ in range(3):
                ctx.export_device(pp, d_out, n_frames, sf, d_sf, sptr)
            torch.cuda.synchronize(dev)
            a0.record(stream)
            for _ in range(args.steps):
                ctx.export_device(pp, d_out, n_frames, sf, d_sf, sptr)
            a1.record(stream)
            torch.cuda.synchronize(dev)
            ctx.wait()
            ems = a0.elapsed_time(a1) / args.steps
            ok = bool(np.array_equal(d_sf[0].cpu().numpy(), osurf.export(d_out[0].cpu().numpy(), pp.pic_width_in_mbs,
                                                                         pp.pic_height_in_mbs, 0, 0, W_, h_disp, fmt)))
            eb = 2.0 * n_frames * sf.nbytes   # every surface byte is read once and written once
            exp[name] = {"ms_per_step": ems, "value": n_frames * W_ * h_disp / (ems * 1e-3) / 1e6, "unit": UNIT,
                         "parity_vs_oracle_first_picture": ok,
                         "roofline": {"bound": "hbm", "achieved": eb / (ems * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                      "frac": eb / (ems * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_pixel": 3.0}}
            del d_sf
        line["surface_export"] = {"workload": f"{W_}x{h_disp} display rectangle of the {n_frames} coded pictures "
                                              "(dryv_recon_export_device), device-resident", **exp}

        # side measurement, SURVEY.md §8(f) next-4: the optional in-loop deblocking post-pass over the batch's pictures, in
        # place (not part of dryv parity: the reference has no filter). Parity of the kernel: tests/test_deblock_oracle.py.
        # Every timed launch filters a FRESH copy of the reconstruction (filtering the same buffer again and again would time
        # ever smoother pictures, on which the filter switches itself off); the copy is outside the event pair.
        d_db = d_outs[0].clone()
        dts = []
        with torch.cuda.stream(stream):
            for i in range(2 + args.steps):
                d_db.copy_(d_outs[0])
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record(stream)
                ctx.deblock_device(dsoa, d_db, 0, 0, sptr)
                b1.record(stream)
                torch.cuda.synchronize(dev)
                if i >= 2:
                    dts.append(b0.elapsed_time(b1))
        ctx.wait()
        dms = sum(dts) / len(dts)
        dbytes = n_mb_step * (384 + 384 + 2)   # every sample read and written once, qp + transform flag per macroblock
        line["deblock"] = {"workload": f"H.264 8.7 deblocking of the {n_frames} reconstructed pictures, in place "
                                       "(dryv_recon_deblock_device; intra: bS 4 / 3 on every edge)",
                           "ms_per_step": dms, "value": n_frames * pp.luma_pixels / (dms * 1e-3) / 1e6, "unit": UNIT,
                           "roofline": {"bound": "hbm", "achieved": dbytes / (dms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": dbytes / (dms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_mb": 770}}
        del d_db

    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = n_frames if cores >= 8 else max(1, min(n_frames, 2 * cores))
        sample = hbatch.frames(0, n)
        cpu_oracle_run(sample.frames(0, min(n, cores)), cores)  # warm
        reps = 3
        best = min(cpu_oracle_run(sample, cores) for _ in range(reps))
        one = cpu_oracle_run(sample.frames(0, min(4, n)), 1)
        line["cpu_baseline"] = {
            "value": n * pp.luma_pixels / best / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
            "single_thread_value": min(4, n) * pp.luma_pixels / one / 1e6,
            "sample": f"{n} of the {n_frames} pictures of rank 0's batch, best of {reps}, one picture per host thread "
                      f"({cores} threads); C oracle port of dryv's Rust frame/ path (the Rust binary cannot be built here)"}

    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
