#!/usr/bin/env python
"""Complete-linkage decode: time per call and the number of parallel rounds / merges they produced (reads the workspace counters)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import hpcs_b200 as hb
from hpcs_b200 import _lib
lib = _lib.load()
for B, N in ((8, 1024), (32, 1024), (8, 4096)):
    x = bench.decode_inputs(B, N, 0).cuda()
    scale = torch.tensor([1e-3], device="cuda")
    leaves = hb.normalize_project(x, scale)
    ws = _lib.workspace(lib.hpcs_linkage_workspace_bytes(B, N, 32, 1), x.device)
    Z = torch.empty((B, N - 1, 4), dtype=torch.float64, device="cuda")
    def run():
        _lib.check(lib.hpcs_linkage_f64(leaves.data_ptr(), B, N, 32, 1, Z.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(x.device)), "linkage")
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    off = lib.hpcs_linkage_debug_counters_offset(B, N, 1)
    extra = ""
    if off:
        c = ws[off:off + 64].view(torch.int32).cpu().tolist()
        extra = f"  cloud 0: merges {c[2]} in {c[3]} rounds; us: snapshot {c[4]} rowmin {c[5]} pairs {c[6]} scratch {c[7]} update {c[8]} barriers {c[9]}"
    print(f"B={B} N={N}: {e0.elapsed_time(e1)/5:.3f} ms per call{extra}")
