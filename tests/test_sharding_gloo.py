"""Multi-GPU host logic on CPU: two gloo ranks shard independent pictures (no data-path collective),
each reconstructs its share with the oracle standing in for the device, and the gathered result equals
the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dryv_b200 import shard


def test_frames_for_rank_partitions_everything():
    for n in (0, 1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += list(shard.frames_for_rank(n, r, world))
            assert seen == list(range(n))
            sizes = [len(shard.frames_for_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert shard.owner_of(5, 8, 2) == 1
    with pytest.raises(ValueError):
        shard.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, tmp):
    import oracle
    from dryv_b200 import synth
    from dryv_b200.abi import PicParams
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pp = PicParams.make(6, 4)
    n_total = 5
    mine = shard.frames_for_rank(n_total, rank, world)
    # every rank generates only its own pictures (picture f is seeded base + f, like bench.py)
    out = np.zeros((n_total, pp.frame_bytes), np.uint8)
    if len(mine):
        b = synth.generate(pp, len(mine), 900 + mine.start)
        out[mine.start:mine.stop] = oracle.reconstruct(b)
    t = torch.from_numpy(out.astype(np.int32))
    dist.all_reduce(t)  # test-only gather (disjoint rows, zeros elsewhere); the product path exchanges nothing
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), t.numpy().astype(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import oracle
    from dryv_b200 import synth
    from dryv_b200.abi import PicParams
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npy")
    pp = PicParams.make(6, 4)
    ref = oracle.reconstruct(synth.generate(pp, 5, 900))
    assert np.array_equal(got, ref)


def _stream_worker(rank, world, port, tmp):
    import oracle
    from dryv_b200 import host
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    movie = open(os.path.join(tmp, "movie.mp4"), "rb").read()
    pp, n_total = host.scan(movie)
    mine = shard.frames_for_rank(n_total, rank, world)
    out = np.zeros((n_total, pp.frame_bytes), np.uint8)
    if len(mine):   # every rank demuxes the file but CABAC-parses and reconstructs only its own pictures
        b = host.parse(movie, threads=1, first=mine.start, count=len(mine))
        out[mine.start:mine.stop] = oracle.reconstruct(b)
    t = torch.from_numpy(out.astype(np.int32))
    dist.all_reduce(t)  # test-only gather; the product path exchanges nothing
    if rank == 0:
        np.save(os.path.join(tmp, "gathered_stream.npy"), t.numpy().astype(np.uint8))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_one_mp4_file(tmp_path):
    """The real multi-GPU host flow on CPU: one MP4 file, each rank parses its share of the IDR pictures."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle
    from avc import mp4, stream
    from dryv_b200 import synth
    from dryv_b200.abi import PicParams
    pp = PicParams.make(6, 4, 1, -1)
    b = synth.generate(pp, 5, 1234)
    (tmp_path / "movie.mp4").write_bytes(mp4.mux(stream.encode_stream(b), 96, 64))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stream_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "gathered_stream.npy"), oracle.reconstruct(b))
