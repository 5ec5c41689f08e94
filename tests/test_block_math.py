"""CPU unit test of the lane-local arithmetic of the packed residual stage (dryv_b200/csrc/residual_stage.cuh compiled
for the host) against the oracle's 4x4 scaling + transform: see tests/native/block_math_test.cpp."""
import os
import subprocess

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_block_math_matches_the_oracle(tmp_path):
    oracle.build()
    exe = str(tmp_path / "block_math_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/native/block_math_test.cpp"),
                           os.path.join(ROOT, "dryv_b200/csrc/recon_tables.cpp"),
                           os.path.join(ROOT, "oracle/libdryv_oracle.so"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok:"), out.stdout + out.stderr
