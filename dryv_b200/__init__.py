"""dryv_b200 — B200 (sm_100a) implementation of dryv's AVC intra macroblock reconstruction path.

  abi     ctypes mirror of include/dryv_recon.h
  recon   binding of the CUDA library (csrc/), contexts, pinned buffers   -- no CPU fallback
  frame   host-side mirror of the reference's Frame::new / decode / write_to_yuv_file
  shard   picture -> GPU assignment (no collective: pictures are independent)
  synth   seeded spec-legal syntax-buffer generator (workload generator for tests and bench)
"""
from . import abi  # noqa: F401

__all__ = ["abi", "recon", "frame", "shard", "synth"]
