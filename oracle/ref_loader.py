"""Loaders for the compiled reference (test infrastructure only -- never imported by tinyknn_b200/).

`load_ref_kernels()`  -> (sse_module, avx_module): the reference's own Cython kernels compiled
                         by oracle/build_ref.py into oracle/_ref/. Travels to the GPU box.
`load_ref_package()`  -> the reference's full Python package, imported from /root/reference with
                         its compiled modules resolved from oracle/_ref/. Only works in the build
                         container (the GPU box has no /root/reference); used by
                         tests/golden/make_golden.py to generate fixtures and by CPU-side tests
                         that validate the numpy restatement in oracle/restate.py.
"""
import importlib.machinery
import importlib.util
import os
import sys

from . import build_ref

_REF_ROOT = os.environ.get("TINYKNN_REFERENCE", "/root/reference")


def have_ref_kernels():
    return build_ref.have_ref()


def have_ref_package():
    return have_ref_kernels() and os.path.isfile(os.path.join(_REF_ROOT, "tinyknn", "__init__.py"))


def _load_ext(mod):
    name = "_tkb_oracle_ref." + mod
    if name in sys.modules:
        return sys.modules[name]
    path = build_ref.ref_so_path(mod)
    loader = importlib.machinery.ExtensionFileLoader(name, path)
    spec = importlib.util.spec_from_file_location(name, path, loader=loader)
    m = importlib.util.module_from_spec(spec)
    loader.exec_module(m)
    sys.modules[name] = m
    return m


def load_ref_kernels():
    if not have_ref_kernels():
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py in the build container)")
    return _load_ext("_fast_pq"), _load_ext("_fast_pq_avx")


def load_ref_package():
    """Import the reference package under the name `tinyknn` (numpy>=2 alias applied first)."""
    if "tinyknn" in sys.modules and getattr(sys.modules["tinyknn"], "_tkb_is_reference", False):
        return sys.modules["tinyknn"]
    if not have_ref_package():
        raise RuntimeError("reference package unavailable (needs /root/reference and oracle/_ref)")
    import numpy._core._methods as _m
    sys.modules.setdefault("numpy.core._methods", _m)          # fast_pq.py:15 imports the numpy<2 path
    pkg_dir = os.path.join(_REF_ROOT, "tinyknn")
    spec = importlib.util.spec_from_file_location(
        "tinyknn", os.path.join(pkg_dir, "__init__.py"),
        submodule_search_locations=[pkg_dir, build_ref.OUT])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["tinyknn"] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop("tinyknn", None)
        raise
    mod._tkb_is_reference = True
    return mod
