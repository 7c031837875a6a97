"""Runs the WHOLE Python package on the CPU emulator (TEST INFRASTRUCTURE ONLY; installed by tests/conftest.py when
TKB_EMU=1, never by the package): every module's `lib` (the ctypes handle of libtinyknn_b200.so) is replaced by
libtkb_emu.so -- the same CUDA sources compiled against cuda_emu.h -- and `_device`'s torch handle by a proxy whose tensors
live on the CPU ("device" memory = host memory) and whose torch.cuda is a stand-in (one stream, no-op synchronisation).
With that the GPU parity tests (-m gpu) execute the product's host layer and the product's kernel sources, and compare them
with the oracle, on a machine without a GPU. Nothing here is reachable from the package itself."""
import contextlib
import sys
import time


class _Stream:
    cuda_stream = 0

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def synchronize(self):
        pass

    def record_event(self, ev=None):
        ev = ev or _Event()
        ev.record()
        return ev


class _Event:
    def __init__(self, enable_timing=False, **kw):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _Graph:
    """torch.cuda.CUDAGraph on the emulator: the library records the launches (with copies of their arguments), memsets and
    copies issued while the capture is open, `replay()` runs them again."""
    capturing = False
    current = None

    def __init__(self):
        self.id = -1
        self.nodes = 0
        self.pool = []                 # tensors allocated during the capture: the graph's private memory pool keeps them alive

    def replay(self):
        import emu_lib
        if emu_lib.load().emu_graph_replay(self.id) != 0:
            raise RuntimeError("emulated graph replay failed: %s" % emu_lib.load().emu_last_error().decode())


class _Capture:
    def __init__(self, graph):
        self.graph = graph

    def __enter__(self):
        import emu_lib
        self.graph.id = emu_lib.load().emu_graph_begin()
        _Graph.capturing, _Graph.current = True, self.graph

    def __exit__(self, *exc):
        import emu_lib
        _Graph.capturing, _Graph.current = False, None
        self.graph.nodes = emu_lib.load().emu_graph_end()
        return False


def _no_sync_in_capture(what):
    if _Graph.capturing:
        raise RuntimeError("%s while a CUDA graph is being captured (a synchronising call: illegal on hardware)" % what)


class _Cuda:
    Event = _Event
    CUDAGraph = _Graph
    _cur = _Stream()

    @staticmethod
    def graph(g, *a, **kw):
        return _Capture(g)

    @staticmethod
    def is_available():
        return True

    @staticmethod
    def current_device():
        return 0

    @staticmethod
    def device_count():
        return 1

    @staticmethod
    def set_device(dev):
        pass

    @classmethod
    def current_stream(cls, device=None):
        return cls._cur

    @staticmethod
    def Stream(*a, **kw):
        return _Stream()

    @staticmethod
    def stream(s):
        return contextlib.nullcontext()

    @staticmethod
    def synchronize(device=None):
        _no_sync_in_capture("torch.cuda.synchronize()")

    @staticmethod
    def mem_get_info(device=None):
        return (8 << 30, 16 << 30)

    @staticmethod
    def empty_cache():
        pass


class _Torch:
    """torch with every device being the CPU."""

    def __init__(self, real):
        self._real = real
        self.cuda = _Cuda()

    def device(self, *a, **kw):
        return self._real.device("cpu")

    def empty(self, *a, **kw):
        kw.pop("pin_memory", None)
        x = self._real.empty(*a, **kw)
        if _Graph.capturing:
            _Graph.current.pool.append(x)
        return x

    def __getattr__(self, name):
        return getattr(self._real, name)


_installed = False


def install():
    global _installed
    if _installed:
        return
    import torch
    import tinyknn_b200                                          # noqa: F401  (the real library loads fine without a GPU)
    from tinyknn_b200 import _lib, _device as D
    sys.path.insert(0, __file__.rsplit("/", 1)[0])
    import emu_lib
    emu = emu_lib.load()
    real = _lib.lib
    for name, mod in list(sys.modules.items()):
        if mod is not None and (name == "tinyknn_b200" or name.startswith("tinyknn_b200.")):
            for attr, val in list(vars(mod).items()):
                if val is real:
                    setattr(mod, attr, emu)
    D._torch = _Torch(torch)
    # a device-to-host copy is a COPY: on the CPU `.cpu()` would alias the "device" buffer, and a later launch that reuses the
    # buffer would silently change what a test had read back
    def _cpu(self, *a, **k):
        _no_sync_in_capture("Tensor.cpu()")
        return self.clone()

    item = torch.Tensor.item

    def _item(self):
        _no_sync_in_capture("Tensor.item()")
        return item(self)

    torch.Tensor.cpu = _cpu
    torch.Tensor.item = _item
    torch.Tensor.pin_memory = lambda self, *a, **k: self            # no driver: pageable memory stands in for pinned memory
    upload = D.upload

    def upload_copy(arr, non_blocking=False):
        # on the CPU `from_numpy(...).to(device)` aliases the host array; a device copy never does
        return upload(arr, non_blocking).clone()

    D.upload = upload_copy
    for name, mod in list(sys.modules.items()):                  # modules that bound `upload` by name
        if mod is not None and name.startswith("tinyknn_b200.") and getattr(mod, "upload", None) is upload:
            mod.upload = upload_copy
    _installed = True
