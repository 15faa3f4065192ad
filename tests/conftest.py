import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle.pyoracle import Reference
    if not Reference.available() and not Path("/root/reference/src").is_dir():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Reference()


@pytest.fixture(scope="session")
def srtlib():
    """The product's C-ABI library; builds it if the .so is missing (nvcc cross-compiles without a GPU)."""
    from simpleraytracing_b200 import capi
    if not capi.LIB_PATH.exists():
        import subprocess
        subprocess.run(["make", "-C", str(ROOT / "simpleraytracing_b200" / "csrc")], check=True)
    return capi.Library()


@pytest.fixture(scope="session")
def dragon():
    from simpleraytracing_b200 import scene
    return scene.dragon()


@pytest.fixture(scope="session")
def dragon_cam(oracle, dragon):
    return oracle.camera(dragon)


@pytest.fixture(scope="session")
def gpu_ctx(srtlib):
    """A context on cuda:0. The GPU tests must run the CUDA path: no device -> hard failure, never a skip
    to some fallback."""
    from simpleraytracing_b200 import Context
    ctx = Context(0, srtlib)
    yield ctx
    ctx.close()


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def load_golden(name):
    g = np.load(GOLDEN / name)
    return dict(L=g["L"], dist=g["dist"], ssum=g["ssum"].astype(np.int32), hits=g["hits"].astype(np.uint32))
