"""pytest configuration: the `gpu` marker and shared helpers.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors, host logic, C-ABI
exports, gloo world_size-2 sharding).  `-m gpu` runs on a B200 and calls the kernels through the
C ABI; those tests compare against the CPU oracle in oracle/ (test infrastructure only).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist before any test touches the package (build is incremental)."""
    from physics_llm_inference_b200 import build
    if not os.path.exists(build.LIB):
        build.build()
    return build.LIB
