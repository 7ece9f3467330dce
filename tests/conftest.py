import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import rlcontrol_b200 as rb
    return rb.Engine(0)


@pytest.fixture(autouse=True)
def _no_stale_kernel_error_flag(request):
    """After every GPU test: the session handle's device error flag (bounded mbarrier waits, operand-range checks) must be
    clear -- a test that provokes it reads (= clears) it itself, so a flag left behind is a failure of THAT test and never
    leaks into the next one."""
    yield
    if "eng" in request.fixturenames:
        e = request.getfixturevalue("eng")
        code = e.umma_error()
        assert code == 0, f"kernel error flag {code} left set by {request.node.name}"


def rel_err(x, ref):
    """|x-ref| / max(|ref|, rms(ref) per state) -- the error metric of SURVEY 7 (element-wise
    relative error is undefined where Q crosses 0)."""
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    if ref.ndim == 2:
        rms = np.sqrt(np.mean(ref * ref, axis=1, keepdims=True))
    else:
        rms = np.sqrt(np.mean(ref * ref))
    den = np.maximum(np.abs(ref), np.maximum(rms, 1e-30))
    return np.abs(x - ref) / den
