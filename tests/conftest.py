import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native pieces are build artefacts (git-ignored): build them when a fresh checkout runs the tests.  nvcc
    # cross-compiles sm_100a without a GPU; on the GPU box the prebuilt files travel with the snapshot.
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "kcftools_b200", "libkcfgpu.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "kcftools_b200", "csrc"), "-j4"], stdout=subprocess.DEVNULL)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libkcforacle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def ctx():
    """one libkcfgpu context on cuda:0 (GPU tests only)"""
    from kcftools_b200.api import Context
    c = Context(0)
    yield c
    c.close()
