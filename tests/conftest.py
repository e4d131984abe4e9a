import os
import socket
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs at least one CUDA device (run with -m gpu on a B200 box)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices")


def pytest_collection_modifyitems(config, items):
    import torch

    ngpu = torch.cuda.device_count() if torch.cuda.is_available() else 0
    for item in items:
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))
        elif "gpu" in item.keywords and ngpu < 1:
            item.add_marker(pytest.mark.skip(reason="needs a GPU"))


def free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.fixture
def clean_cgx_env(monkeypatch):
    for k in list(os.environ):
        if k.startswith("CGX_"):
            monkeypatch.delenv(k, raising=False)
    import torch_cgx_b200

    torch_cgx_b200.reset_layers()
    yield
    torch_cgx_b200.reset_layers()
