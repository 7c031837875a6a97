#!/usr/bin/env python3
"""Build the UNMODIFIED reference kernels into oracle/_ref/ (test infrastructure only).

This is the recipe the task calls `oracle/_ref`: the two Cython sources of the
reference are translated and compiled *from where they lie* under
/root/reference (read-only); nothing is copied into the repo and only build
outputs land in oracle/_ref/ (git-ignored, but shipped to the GPU box).

  /root/reference/tinyknn/_fast_pq.pyx      -> oracle/_ref/_fast_pq.<abi>.so      (SSE order)
  /root/reference/tinyknn/_fast_pq_256.pyx  -> oracle/_ref/_fast_pq_avx.<abi>.so  (AVX order, default)

Differences from the reference's own setup.py (setup.py:17-47), neither of
which touches the kernel arithmetic:
  * Cython 3.x needs `legacy_implicit_noexcept=True` to accept the
    `cdef inline __m128i ...` helpers.
  * `-march=x86-64-v3` (AVX2/BMI2/FMA) instead of `-march=native`, so the .so
    built in this container cannot SIGILL on the GPU box's host CPU.

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may load what this produces.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TINYKNN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")

MODULES = [
    # (pyx, module name, extra flags)
    ("tinyknn/_fast_pq.pyx", "_fast_pq", []),
    ("tinyknn/_fast_pq_256.pyx", "_fast_pq_avx", ["-mavx"]),
]

CFLAGS = ["-O3", "-march=x86-64-v3", "-ffast-math", "-Wno-unused-function",
          "-fprefetch-loop-arrays", "-fPIC", "-shared", "-std=c++17",
          "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION"]


def ref_so_path(mod):
    return os.path.join(OUT, mod + sysconfig.get_config_var("EXT_SUFFIX"))


def have_ref():
    return all(os.path.exists(ref_so_path(m)) for _, m, _ in MODULES)


def build(force=False):
    """Returns True when oracle/_ref holds both compiled reference modules."""
    if have_ref() and not force:
        return True
    if not os.path.isdir(os.path.join(REF, "tinyknn")):
        return have_ref()
    import numpy as np
    os.makedirs(os.path.join(OUT, "_gen"), exist_ok=True)
    inc = ["-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()]
    for pyx, mod, extra in MODULES:
        cpp = os.path.join(OUT, "_gen", mod + ".cpp")
        subprocess.check_call([
            sys.executable, "-m", "cython", "--cplus", "-3",
            "-X", "legacy_implicit_noexcept=True",
            "-X", "boundscheck=False", "-X", "wraparound=False",
            "-X", "cdivision=True", "-X", "initializedcheck=False",
            "-X", "nonecheck=False", "-X", "overflowcheck=False",
            "--module-name", "tinyknn." + mod,
            "-o", cpp, os.path.join(REF, pyx)])
        subprocess.check_call(["g++"] + CFLAGS + extra + inc + [cpp, "-o", ref_so_path(mod)])
    # the generated C++ is derived from reference sources: do not keep it around
    import shutil
    shutil.rmtree(os.path.join(OUT, "_gen"), ignore_errors=True)
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "unavailable (no /root/reference)")
