import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no CUDA device is visible so a plain
    `pytest tests/` on the CPU container stays green."""
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
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _fresh_philox_state(request):
    """The Philox site / step counters and the deterministic-mode switch are process-global (like torch's default
    generator): every GPU test starts from step 0 / site 0 / atomic mode, so no test depends on what ran before it."""
    if "gpu" not in request.keywords:
        yield
        return
    import torch
    if not torch.cuda.is_available():
        yield
        return
    import vae_gan_b200 as v
    dev = torch.device("cuda", torch.cuda.current_device())
    v.rng.reset_sites()
    v.rng.step_tensor(dev).zero_()
    yield
    if v.is_deterministic():
        v.set_deterministic(False)
