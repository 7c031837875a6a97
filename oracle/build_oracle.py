#!/usr/bin/env python3
"""Compile oracle/pq_oracle.c -> oracle/_build/libpq_oracle.so (test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pq_oracle.c")
OUT = os.path.join(HERE, "_build", "libpq_oracle.so")


def build(force=False):
    if (not force and os.path.exists(OUT)
            and (not os.path.exists(SRC) or os.path.getmtime(OUT) >= os.path.getmtime(SRC))):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + ".tmp%d" % os.getpid()
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", SRC, "-o", tmp])
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
