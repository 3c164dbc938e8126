import logging
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_SRC = "/root/reference/src"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    logging.getLogger("rho_tts").setLevel(logging.ERROR)


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "rho_tts"))


@pytest.fixture(scope="session")
def reference_basetts():
    """The reference's BaseTTS, when /root/reference is present (authoring container only)."""
    if not have_reference():
        pytest.skip("/root/reference not present on this box")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    from rho_tts.base_tts import BaseTTS
    logging.getLogger("rho_tts.base_tts").setLevel(logging.ERROR)
    return BaseTTS


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch.device("cuda", 0)
