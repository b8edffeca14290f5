"""In-tree build of ``hpcs_b200/lib/libhpcs_b200.so`` (sm_100a only) with plain nvcc.

The built library is git-ignored but travels to the GPU box with the working-tree snapshot.
``python -m hpcs_b200.build [--force] [--verbose]``
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhpcs_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

SOURCES = ["abi.cu", "knn.cu", "knn_d3.cu", "knn_tc.cu", "edge_feat.cu", "edge_bwd.cu", "hyp_loss.cu", "sampler.cu", "hyp_ops.cu", "linkage.cu", "cut.cu", "input.cu", "peaks.cu", "edgeconv.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; hpcs_b200 has no prebuilt or CPU fallback")
    return cand


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "hpcs_b200.h"))
    return hs


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    built = os.path.getmtime(target)
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile each .cu to an object (in parallel, only the stale ones) and link the shared library."""
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = _headers()
    jobs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _newer(obj, [path, *headers]):
            cmd = [_nvcc(), *NVCC_FLAGS, "-c", path, "-o", obj]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]

    def run(cmd):
        return cmd, subprocess.run(cmd, capture_output=True, text=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for cmd, proc in pool.map(run, jobs):
                if verbose or proc.returncode != 0:
                    sys.stderr.write(" ".join(cmd[-3:]) + "\n" + proc.stdout + proc.stderr)
                if proc.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {cmd[-3]}")
    if jobs or not os.path.exists(LIB_PATH):
        proc = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", *objs, "-o", LIB_PATH],
                              capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
            raise RuntimeError("link of libhpcs_b200.so failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
