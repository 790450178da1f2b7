import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    try:        # property tests: the same examples on every run (a suite that is green stays green); HYPOTHESIS_PROFILE=explore
        from hypothesis import settings          # draws fresh ones to go hunting

        settings.register_profile("repeatable", derandomize=True, database=None, deadline=None)
        settings.register_profile("explore", deadline=None)
        settings.load_profile(os.environ.get("HYPOTHESIS_PROFILE", "repeatable"))
    except ImportError:
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def v6home(tmp_path, monkeypatch):
    """Isolated config/data/log root for CLI + runtime tests."""
    monkeypatch.setenv("V6B200_HOME", str(tmp_path / "v6home"))
    return tmp_path / "v6home"
