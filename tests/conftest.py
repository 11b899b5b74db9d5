import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def cuda_lib():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # the PyTorch fp32 reference ops must not depend on cuDNN/TF32: loading cuDNN on a fresh box can take
    # minutes (observed) and TF32 would blur the comparison
    torch.backends.cudnn.enabled = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from b200sd import lib

    lib.load()  # raises loudly if libb200sd.so is missing -- never fall back
    return lib
