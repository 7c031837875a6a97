import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


EMU = os.environ.get("TKB_EMU", "0") == "1"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    if EMU:
        # TKB_EMU=1: the gpu-marked tests run on the CPU emulator (tests/emulate): the package's host layer on CPU tensors, the
        # library's CUDA sources compiled against cuda_emu.h. Test infrastructure only; see tests/emulate/emu_torch.py.
        sys.path.insert(0, os.path.join(ROOT, "tests", "emulate"))
        import emu_torch
        emu_torch.install()


def pytest_terminal_summary(terminalreporter):
    if EMU:
        import emu_lib
        terminalreporter.write_line("emulator: %s" % emu_lib.stats())


def pytest_collection_modifyitems(config, items):
    # GPU tests skip themselves cleanly when collected on a machine without a device.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu or EMU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz")) for name in ("scan", "lut", "ivf", "encode")}


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """The CUDA library and the C oracle are built in-tree before any test runs."""
    from tinyknn_b200 import build as tkb_build
    from oracle import build_oracle
    if tkb_build.needs_build():
        tkb_build.build()
    build_oracle.build()
