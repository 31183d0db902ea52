import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def torch_stage_backend(monkeypatch):
    """TEST-ONLY: run the host-side composition (fusion_gcn_b200.functional) on top of the plain-torch
    stage oracle so that it can be checked on a CPU box.  The package itself has no such path."""
    from oracle import stages
    import fusion_gcn_b200.functional as FN
    monkeypatch.setattr(FN, "K", stages)
    return stages
