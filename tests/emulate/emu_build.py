"""Builds tests/emulate/_build/libtkb_emu.so: the SOURCES of libtinyknn_b200.so (tinyknn_b200/csrc/*.cu) translated to plain
C++ and compiled against cuda_emu.h, so that the C ABI of include/tinyknn_b200.h runs on the CPU with "device" pointers that
are host pointers. TEST INFRASTRUCTURE ONLY -- the package never loads this library.

The translation is textual and small:
  * `kernel<targs><<<grid, block, smem, stream>>>(args)`  ->  `::emu::launch(dim3(grid), dim3(block), smem, [=]() { kernel<targs>(args); })`
    (the closure owns copies of the arguments, like a kernel node of a CUDA graph: cuda_emu.cpp can record and replay it)
  * `extern __shared__ [__align__(n)] T name[];`         ->  `T *name = reinterpret_cast<T *>(::emu::dyn_smem());`
  * `__noinline__`                                       ->  `__attribute__((noinline))`
everything else (qualifiers, intrinsics, the runtime API) is supplied by cuda_emu.h / shim/cuda_runtime.h; the four inline-PTX
helpers of the library carry an `#ifdef TKB_EMULATE` alternative next to the asm statement.
"""
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "tinyknn_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libtkb_emu.so")
CXXFLAGS = ["-std=c++17", "-O1", "-g", "-ffp-contract=off", "-mfma", "-fPIC", "-DTKB_EMULATE", "-w",
            "-I", os.path.join(HERE, "shim"), "-I", HERE, "-I", CSRC]

# TKB_EMU_UBSAN=1: misaligned accesses through the vector types (uint4, uint2, float4, double2: what the GPU reports as
# "misaligned address") and out-of-bounds indexing of fixed-size arrays abort with a message instead of passing silently on x86
if os.environ.get("TKB_EMU_UBSAN", "0") == "1":
    CXXFLAGS = CXXFLAGS + ["-fsanitize=alignment,bounds", "-fno-sanitize-recover=all"]
    OUT = OUT + "_ubsan"
    LIB = os.path.join(OUT, "libtkb_emu.so")

_KERNEL_EXPR = re.compile(r"([A-Za-z_]\w*(?:\s*<[^<>;(){}]*>)?)\s*$")
_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\s*\d+\s*\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\s*\[\s*\]\s*;")


def _balanced(text, start, open_ch="(", close_ch=")"):
    """Index just past the bracket that closes text[start] (which must be open_ch)."""
    assert text[start] == open_ch, text[start:start + 20]
    depth = 0
    for i in range(start, len(text)):
        c = text[i]
        if c == open_ch:
            depth += 1
        elif c == close_ch:
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced bracket")


def translate(src):
    out, pos, n_launch = [], 0, 0
    while True:
        at = src.find("<<<", pos)
        if at < 0:
            out.append(src[pos:])
            break
        m = _KERNEL_EXPR.search(src, pos, at)
        if not m:
            raise ValueError("cannot find the kernel expression before <<< at offset %d" % at)
        end_cfg = src.index(">>>", at)
        cfg = src[at + 3:end_cfg]
        parts, depth, cur = [], 0, ""
        for c in cfg:                                            # split the launch configuration at top-level commas
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += c
        parts.append(cur)
        while len(parts) < 4:
            parts.append("0")
        grid, block, smem, _stream = (p.strip() for p in parts[:4])
        a0 = end_cfg + 3
        while src[a0] in " \t\\\n":
            a0 += 1
        a1 = _balanced(src, a0)
        out.append(src[pos:m.start(1)])
        out.append("::emu::launch(dim3(%s), dim3(%s), (size_t)(%s), [=]() { %s%s; })" % (grid, block, smem, m.group(1), src[a0:a1]))
        pos = a1
        n_launch += 1
    text = "".join(out)
    text, n_sm = _EXTERN_SHARED.subn(lambda m: "%s *%s = reinterpret_cast<%s *>(::emu::dyn_smem());" % (m.group(1), m.group(2), m.group(1)), text)
    text = re.sub(r"\b__noinline__\b", "__attribute__((noinline))", text)      # libstdc++ spells the attribute __noinline__ itself
    return text, n_launch, n_sm


def sources():
    sys.path.insert(0, ROOT)
    from tinyknn_b200 import build as B
    return [os.path.join(CSRC, s) for s in B.SOURCES], B.HEADERS


def source_hash():
    h = hashlib.sha256(" ".join(CXXFLAGS).encode())
    cu, hdr = sources()
    for p in cu + hdr + [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "emu_build.py", "shim/cuda_runtime.h", "selftest.cu")]:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = LIB + ".srchash"
    digest = source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    cu, hdr = sources()
    procs, objs = [], []
    total_launch = 0
    for path in hdr:                                             # the headers next to the translated sources (found first by #include "")
        if path.endswith(".cuh"):
            with open(os.path.join(OUT, "src", os.path.basename(path)), "w") as f:
                f.write('#line 1 "%s"\n' % path)
                f.write(translate(open(path).read())[0])
    for path in cu + [os.path.join(HERE, "selftest.cu")]:
        text, n_launch, n_sm = translate(open(path).read())
        total_launch += n_launch
        if verbose:
            print("%s: %d launches, %d dynamic shared-memory declarations" % (os.path.basename(path), n_launch, n_sm))
        cpp = os.path.join(OUT, "src", os.path.basename(path).replace(".cu", ".emu.cpp"))
        with open(cpp, "w") as f:
            f.write('#line 1 "%s"\n' % path)
            f.write(text)
        obj = cpp.replace(".cpp", ".o")
        objs.append(obj)
        procs.append((path, subprocess.Popen(["g++"] + CXXFLAGS + ["-c", cpp, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rt = os.path.join(OUT, "src", "cuda_emu.o")
    objs.append(rt)
    procs.append(("cuda_emu.cpp", subprocess.Popen(["g++"] + CXXFLAGS + ["-c", os.path.join(HERE, "cuda_emu.cpp"), "-o", rt],
                                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for path, p in procs:
        outp, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("g++ failed on %s:\n%s" % (path, outp[-6000:]))
    tmp = LIB + ".tmp%d" % os.getpid()
    subprocess.check_call(["g++", "-shared", "-o", tmp] + objs + (["-fsanitize=alignment,bounds"] if "-fno-sanitize-recover=all" in CXXFLAGS else []))
    os.replace(tmp, LIB)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
