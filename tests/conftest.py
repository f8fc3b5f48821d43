import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a CUDA device skips the gpu-marked tests instead of failing them."""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def params():
    from orcai_b200 import runtime

    return runtime.bundled_parameters()


@pytest.fixture(scope="session")
def ctx(params):
    """liborcai_b200 context on cuda:0 with seeded synthetic weights loaded."""
    if not _have_gpu():
        pytest.skip("no CUDA device")
    from orcai_b200 import runtime
    from orcai_b200.weights import synthetic_weights

    P, S = params
    c = runtime.get_context(P, S, 0)
    c.load_weights(synthetic_weights(P, S, seed=1234))
    return c


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(autouse=True)
def _reference_precision_by_default(request):
    """GPU tests start on the fp32 network path; tests of the fast path select it explicitly."""
    if "gpu" in request.keywords and _have_gpu():
        from orcai_b200 import runtime

        P, S = runtime.bundled_parameters()
        c = runtime.get_context(P, S, 0)
        c.set_option("net_path", 0)
        c.set_option("tail_path", 1)
        c.owner = None          # options were changed behind the back of whichever model had bound the context
    yield
