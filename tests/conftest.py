import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _build_host_libs():
    """gcc-built pieces (oracle = test infrastructure, synth = workload generator) are built on demand."""
    import oracle
    from dryv_b200 import synth
    oracle.build()
    synth.build()


@pytest.fixture(scope="session")
def recon_lib():
    """The CUDA library, built in-tree by __graft_entry__.build() (nvcc cross-compiles without a GPU)."""
    from dryv_b200 import recon
    if not os.path.exists(recon.LIB_PATH):
        recon.build()
    return recon.load_library()


@pytest.fixture(scope="session")
def gpu_ctx(recon_lib):
    from dryv_b200 import recon
    ctx = recon.ReconContext(0)
    yield ctx
    ctx.close()
