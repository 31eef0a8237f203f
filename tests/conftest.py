import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C-ABI library is a build artefact (git-ignored): build it once if this checkout has none yet
    # (nvcc cross-compiles without a GPU; ~45 s).  A failing build surfaces in test_abi.py, not here.
    from wsi_hgnn_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        try:
            from wsi_hgnn_b200 import build
            build.build(verbose=False)
        except Exception as e:                       # noqa: BLE001
            sys.stderr.write(f"[conftest] building libwsi_hgnn.so failed: {e}\n")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
