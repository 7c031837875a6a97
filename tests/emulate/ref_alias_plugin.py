"""pytest plugin (tests only): `import tinyknn` resolves to tinyknn_b200 running on the CPU emulator, so that the reference's
OWN test files can be collected from where they lie and run unmodified against this package
(`python -m pytest -p ref_alias_plugin /root/reference/tests`, see tests/test_emulated_kernels.py)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import emu_torch                                                   # noqa: E402

emu_torch.install()
import tinyknn_b200                                                # noqa: E402

sys.modules["tinyknn"] = tinyknn_b200
for _sub in ("_fast_pq", "_fast_pq_avx", "_transform", "utils", "fast_pq", "ivf"):
    sys.modules["tinyknn." + _sub] = __import__("tinyknn_b200." + _sub, fromlist=["_"])
