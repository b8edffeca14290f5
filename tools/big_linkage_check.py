"""Large-N decode check against scipy (test infrastructure): N = 8192 and 6000, complete and single linkage, bit-equality.
  python tools/big_linkage_check.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpcs_b200 as hb
from scipy.cluster.hierarchy import linkage
gen = torch.Generator().manual_seed(4)
for N, B in ((8192, 1), (6000, 2)):
    x = torch.randn(B, N, 32, generator=gen) * 0.3
    scale = torch.tensor([1e-3]).cuda()
    for method in ("complete", "single"):
        t = time.time()
        Z, leaves = hb.decode_linkage_batch(x.cuda(), scale, method, return_leaves=True)
        torch.cuda.synchronize(); dt = time.time() - t
        Z = Z.cpu().numpy(); leaves = leaves.cpu().numpy()
        ok = all(np.array_equal(Z[b], linkage(leaves[b], method=method, metric="cosine")) for b in range(B))
        print(N, B, method, "bit-equal to scipy:", ok, f"{dt*1e3:.1f} ms")
