#!/usr/bin/env python3
"""Compile oracle/pq_oracle.c -> oracle/_build/libpq_oracle.so (test infrastructure only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pq_oracle.c")
OUT = os.path.join(HERE, "_build", "libpq_oracle.so")


def build(force=False):
    build_chain(force)
    if (not force and os.path.exists(OUT)
            and (not os.path.exists(SRC) or os.path.getmtime(OUT) >= os.path.getmtime(SRC))):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + ".tmp%d" % os.getpid()
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", SRC, "-o", tmp])
    os.replace(tmp, OUT)
    return OUT


CHAIN_SRC = os.path.join(HERE, "chain_oracle.c")
CHAIN_OUT = os.path.join(HERE, "_build", "libchain_oracle.so")


def build_chain(force=False):
    """oracle/chain_oracle.c -> oracle/_build/libchain_oracle.so (explicit roundings: no FP contraction)."""
    if not force and os.path.exists(CHAIN_OUT) and os.path.getmtime(CHAIN_OUT) >= os.path.getmtime(CHAIN_SRC):
        return CHAIN_OUT
    os.makedirs(os.path.dirname(CHAIN_OUT), exist_ok=True)
    tmp = CHAIN_OUT + ".tmp%d" % os.getpid()
    base = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-ffp-contract=off", CHAIN_SRC, "-o", tmp, "-lm"]
    try:                                                    # hardware FMA when the host has it; libm's fma() is exact either way
        subprocess.check_call(base[:1] + ["-mfma"] + base[1:], stderr=subprocess.DEVNULL)
    except subprocess.CalledProcessError:
        subprocess.check_call(base)
    os.replace(tmp, CHAIN_OUT)
    return CHAIN_OUT


if __name__ == "__main__":
    print(build_chain(force=True))
    print(build(force=True))
