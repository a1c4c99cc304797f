import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Builds the C restatement if needed."""
    from oracle import arfe_oracle, build_oracle
    build_oracle.build_c_oracle()
    return arfe_oracle


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    # The module tests compare cuDNN / cuBLAS convolutions and FCs (PyTorch, outside the
    # path) with the fp32 CPU oracle: keep them in true fp32 (cuDNN's TF32 default gives
    # ~1e-3 relative error, e.g. once channels-last inputs select its tensor-core kernels)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")
