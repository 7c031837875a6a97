"""Dry run of bench.py on the CPU emulator (TEST INFRASTRUCTURE ONLY; never a measurement): the same main(), with torch.cuda
replaced by CPU stand-ins and the library by libtkb_emu.so, on the `tiny` workload. It exists so that a Python-level mistake in
bench.py (a wrong key, a formatting error in the JSON line) is caught in the build container instead of costing the round its
bench line on the GPU box. The printed numbers mean nothing.

    python tests/emulate/emu_bench.py --workload tiny --queries 128 --steps 2 --warmup 3 --cpu-seconds 1
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch                                                       # noqa: E402

import emu_torch                                                   # noqa: E402

emu_torch.install()
fake = emu_torch._Cuda()
for name in ("is_available", "current_device", "device_count", "set_device", "current_stream", "Stream", "stream", "synchronize",
             "mem_get_info", "empty_cache", "Event", "CUDAGraph", "graph"):
    setattr(torch.cuda, name, getattr(fake, name))


class _Profiler:
    @staticmethod
    def start():
        pass

    @staticmethod
    def stop():
        pass


torch.cuda.profiler = _Profiler
torch.Tensor.cuda = lambda self, *a, **k: self.clone()
_tensor = torch.tensor
torch.tensor = lambda *a, **k: _tensor(*a, **{x: y for x, y in k.items() if x != "device"})

import torch.distributed as _dist                                   # noqa: E402

_init = _dist.init_process_group


def _init_gloo(backend=None, **kw):                                 # N > 1 dry runs: gloo stands in for NCCL
    kw.pop("device_id", None)
    return _init("gloo", **kw)


_dist.init_process_group = _init_gloo

import __graft_entry__                                             # noqa: E402

__graft_entry__.build = lambda: None                               # the real build ran already; nothing to compile here

if __name__ == "__main__":
    sys.argv = [os.path.join(ROOT, "bench.py")] + sys.argv[1:]
    runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
