import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    # a fresh checkout has no built library (it is git-ignored): build it once instead of failing every test that
    # loads the C-ABI.  nvcc cross-compiles sm_100a without a device.
    lib = ROOT / "odeintr_b200" / "libodeintr_b200.so"
    if not lib.exists():
        import subprocess

        subprocess.run(["make", "-C", str(ROOT / "odeintr_b200" / "csrc")], check=True, capture_output=True)


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # -m gpu tests need a device; everywhere else they are skipped (never silently passed)
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
