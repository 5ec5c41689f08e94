"""CPU unit test of the lane-local arithmetic of the packed residual stage (dryv_b200/csrc/residual_stage.cuh compiled
for the host) against the oracle's 4x4 scaling + transform: see tests/native/block_math_test.cpp."""
import os
import subprocess

import pytest

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


def test_packed_deblock_filters_match_the_scalar_statement(tmp_path):
    """dryv_b200/csrc/deblock_packed.cuh (two lines per register, branch-free) against a scalar statement of H.264
    8.7.2.3 / 8.7.2.4 — tests/native/deblock_math_test.cpp."""
    exe = str(tmp_path / "deblock_math_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/native/deblock_math_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok:"), out.stdout + out.stderr


@pytest.mark.gpu
def test_packed_deblock_filters_device_equals_host(tmp_path):
    """The same functions compiled for sm_100a (real PRMT / VIMNMX / VIADDMNMX instructions) against their host build."""
    exe = str(tmp_path / "deblock_math_gpu_test")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-o", exe,
                           os.path.join(ROOT, "tests/native/deblock_math_gpu_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("ok:"), out.stdout + out.stderr
