"""CPU: the C-ABI library builds, loads, and exports every symbol include/cassnat_b200.h declares
(no compute calls: there is no GPU here)."""
import os
import re

from cassnat_asr_public_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cassnat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header():
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert lib.cn_version() >= 100
    assert _lib.status_string(0) == "ok" and "workspace" in _lib.status_string(-4)


def test_workspace_queries():
    lib = _lib.load()
    assert lib.cn_bestpath_workspace_bytes(256, 1000) >= 3 * 256 * 1000 * 4
    assert lib.cn_viterbi_workspace_bytes(256, 1000, 200) == 256 * 1000 * 32 * 4
    assert lib.cn_viterbi_workspace_bytes(2, 10, 600) < 0          # U > 511 unsupported


def test_no_fallback_without_cuda():
    import pytest
    import torch

    from cassnat_asr_public_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        ops.best_path_align(torch.zeros(1, 4, 5), torch.ones(1, 1, 4, dtype=torch.bool), torch.tensor([4]), 0)
