"""Compile the oracle's C restatements into ``oracle/_build/`` (TEST INFRASTRUCTURE ONLY).

The reference is pure Python (no native sources), so there is no ``oracle/_ref`` binary to build;
see DESIGN.md "Oracle".
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(HERE, "knn_canonical.c")
    out = os.path.join(out_dir, "libknn_canonical.so")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


if __name__ == "__main__":
    print(build(force=True))
