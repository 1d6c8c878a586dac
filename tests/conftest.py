import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tf_faster_rcnn_b200 import paths as _paths  # noqa: E402
_paths.add_lib_path()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tf_faster_rcnn_b200 import _native
    _native.check(_native.lib().frcnn_check_device(0), "check_device")
    return torch.device("cuda:0")
