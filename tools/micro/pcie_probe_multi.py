"""How much host<->device bandwidth N GPUs of one box get when they copy at the same time (one process per GPU,
launched with torch.distributed.run; gloo only for the barrier). Prints, per rank, H2D / D2H GB/s alone and together."""
import os
import time

import torch
import torch.distributed as dist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
n_in, n_out = 105 << 20, 200 << 20
h = torch.empty(n_in, dtype=torch.uint8).pin_memory()
d = torch.empty(n_in, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_in, do_out, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best


a, b, c = run(True, False), run(False, True), run(True, True)
out = torch.tensor([a, b, c], dtype=torch.float64)
gathered = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
dist.all_gather(gathered, out)
if rank == 0:
    for r, g in enumerate(gathered):
        a, b, c = g.tolist()
        print(f"rank {r}: H2D {n_in / a / 1e9:5.1f} GB/s  D2H {n_out / b / 1e9:5.1f} GB/s  both {(n_in + n_out) / c / 1e9:5.1f} GB/s total ({c * 1e3:.2f} ms)")
    worst = max(g[2].item() for g in gathered)
    print(f"world {world}: aggregate both-direction {(n_in + n_out) * world / worst / 1e9:.1f} GB/s")
dist.destroy_process_group()
