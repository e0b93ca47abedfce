import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture
def mas_env(monkeypatch):
    """Set MAS_* tuning knobs for one test.  The library reads them once, so it is told to re-read them now
    and again after the environment has been restored."""
    from torch_tts_b200 import _lib

    def set_env(**kv):
        for k, v in kv.items():
            monkeypatch.setenv(k, str(v))
        _lib.reload_config()

    yield set_env
    monkeypatch.undo()
    _lib.reload_config()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
