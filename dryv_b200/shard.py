"""Frame sharding across GPUs: independent IDR pictures are dealt to ranks, nothing is exchanged.

The reference is single-process and has no notion of this; pictures never reference each other on this path
(Frame::new starts from zeroed planes, src/video/frame/mod.rs:29-46), so the only multi-GPU logic is which
rank reconstructs which picture and where its output lands."""
from __future__ import annotations


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """Contiguous block of pictures owned by `rank` (the first n_frames % world ranks get one extra)."""
    if world <= 0 or not 0 <= rank < world or n_frames < 0:
        raise ValueError("bad sharding arguments")
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def owner_of(frame: int, n_frames: int, world: int) -> int:
    for r in range(world):
        if frame in frames_for_rank(n_frames, r, world):
            return r
    raise ValueError("frame out of range")
