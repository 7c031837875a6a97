"""Builds libtinyknn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m tinyknn_b200.build [--force]

The .so is git-ignored but ships to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtinyknn_b200.so")
SOURCES = ["tkb_api.cu", "tkb_scan.cu", "tkb_scan_fast.cu", "tkb_scan_tc.cu", "tkb_heap.cu", "tkb_lut.cu", "tkb_rescore.cu", "tkb_plan.cu", "tkb_fused.cu", "tkb_encode.cu", "tkb_assign.cu", "tkb_coarse.cu", "tkb_kmeans.cu"]
HEADERS = [os.path.join(CSRC, "tkb_common.cuh"), os.path.join(CSRC, "tkb_scan_core.cuh"), os.path.join(CSRC, "tkb_rescore_core.cuh"),
           os.path.join(os.path.dirname(HERE), "include", "tinyknn_b200.h")]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libtinyknn_b200.so")


STAMP = LIB + ".srchash"


def source_hash():
    """Content hash of every source that goes into the library (mtimes do not survive the trip to the GPU box)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose and out:
            print(out)
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
                          + ["-Xcompiler", "-fvisibility=hidden"])
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
