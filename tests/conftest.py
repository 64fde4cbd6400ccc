import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def align_golden():
    return dict(np.load(os.path.join(GOLDEN, "align.npz")))


def load_blocks(name):
    import json
    from types import SimpleNamespace

    import torch
    z = np.load(os.path.join(GOLDEN, f"blocks_{name}.npz"))
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd::")}
    cfg = SimpleNamespace(**json.loads(bytes(z["cfg_json"]).decode()))
    return z, sd, cfg


@pytest.fixture(scope="session", params=["trf", "con"])
def blocks_golden(request):
    return (request.param,) + load_blocks(request.param)
