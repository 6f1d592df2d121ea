import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the shared library is a build artefact (git-ignored): build it once if a fresh checkout lacks it
    lib = os.path.join(ROOT, "ultrasphere_b200", "lib", "libultrasphere_b200.so")
    if not os.path.isfile(lib):
        import subprocess

        subprocess.run(["make", "-C", os.path.join(ROOT, "ultrasphere_b200", "csrc")], check=True)


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
