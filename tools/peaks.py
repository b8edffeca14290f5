#!/usr/bin/env python
"""Run the pipe / cache peak probes (csrc/peaks.cu) on the current GPU and print one JSON object.
bench.py imports :func:`measure` and calls it in-process; stand-alone use: ``python tools/peaks.py``."""
from __future__ import annotations

import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (probe id, loop trips, unit scale, unit)
PROBES = {
    "fp32_ffma2_tflops": (0, 4096, 1e12, "TFLOP/s"),
    "fp32_ffma_tflops": (1, 4096, 1e12, "TFLOP/s"),
    "alu_minmax_tops": (2, 4096, 1e12, "Top/s"),
    "fp64_dfma_tflops": (3, 2048, 1e12, "TFLOP/s"),
    "fp64_dmma_tflops": (4, 1024, 1e12, "TFLOP/s"),
    "tf32_umma_tflops": (5, 400, 1e12, "TFLOP/s"),
    "l2_gather_128B_gbs": (6, 1024, 1e9, "GB/s"),
    "l2_gather_256B_gbs": (7, 1024, 1e9, "GB/s"),
}


def measure(dev=None, only=None, reps: int = 5) -> dict:
    from hpcs_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if dev is None else dev
    out = torch.zeros(1, device=dev)
    table = torch.randn(4 << 18, device=dev)                   # 4 MB: the loss kernel's table size, L2 resident
    res = {}
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr(dev)
        for name, (which, iters, scale, unit) in PROBES.items():
            if only and name not in only:
                continue
            work = ctypes.c_double(0)

            def launch():
                _lib.check(lib.hpcs_peak_probe(which, iters, table.data_ptr(), table.numel() * 4, out.data_ptr(),
                                               ctypes.byref(work), stream), "hpcs_peak_probe")
            for _ in range(2):
                launch()
            torch.cuda.synchronize()
            best = None
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                launch()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
            res[name] = round(work.value / (best * 1e-3) / scale, 2)
            res[name + "_ms"] = round(best, 3)
    res["how"] = ("csrc/peaks.cu: dependency-free register / L2 loops, best of %d launches of 1-3 ms, CUDA events; "
                  "148 SMs x 8 CTAs x 256 threads (tcgen05 probe: one CTA per SM, M=128 N=256 K=8 from resident shared memory)" % reps)
    return res


if __name__ == "__main__":
    print(json.dumps(measure()))
