import os
import sys

import numpy as np
import pytest

# The one-device emulation of a decomposed run (test_gpu_dist.py) keeps up to 8 ranks x 3 streams busy with kernels
# that wait for each other's flags: with the default 8 hardware work queues, streams share a queue and a waiting
# kernel can sit in front of the one it waits for.  Must be set before the CUDA context exists.  (A real multi-GPU
# run has one process and 3 such streams per device and does not need it.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a bare `pytest tests/` on a GPU-less box skips the gpu tier instead of failing
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build everything once (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True


@pytest.fixture(scope="session")
def oracle(built):
    """The checker: the reference's own compiled code where oracle/_ref is present (it ships to
    the GPU box), the pinned C port otherwise / for the window operators."""
    from oracle import Checker
    chk = Checker()
    print(f"[conftest] parity checker = {chk.kind}")
    return chk


@pytest.fixture(scope="session")
def port(built):
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref(built):
    from oracle import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libfluid_ref.so not built (no /root/reference here)")
    r = Ref()
    if not r.saturates():
        pytest.skip("host lacks AVX-512F: reference float->uint32 wraps instead of saturating")
    return r


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def ctx(built):
    import esp32_fluid_simulation_b200 as fb
    return fb.Context(0)


def bits(a):
    """Bit pattern view for exact float comparison (distinguishes -0.0, NaNs)."""
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    return a


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ba, bb = bits(a), bits(b)
    if not np.array_equal(ba, bb):
        bad = np.argwhere(ba != bb)
        first = tuple(bad[0])
        raise AssertionError(
            f"{what}: {len(bad)} of {ba.size} words differ; first at {first}: "
            f"{a[first]!r} vs {b[first]!r}")
