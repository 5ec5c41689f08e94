"""CPU oracle of dryv's reconstruction path — TEST INFRASTRUCTURE ONLY (see dryv_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED by the reference (no tests/fixtures there; Rust toolchain absent here); the standard-conformant
part of its behaviour is pinned to libavcodec's output (tests/test_libavcodec_crosscheck.py).
"""
from .oracle import *  # noqa: F401,F403
