"""Debug harness for the tensor-core kNN path: compares method="auto" with the FFMA kernel and prints
mismatch statistics (never asserts), plus timings.   python tools/knn_tc_check.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpcs_b200 as hb  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    cases = [(2, 63, 128, 20, "randn"), (2, 63, 1024, 20, "randn"), (3, 32, 300, 10, "randn"), (1, 16, 128, 40, "randn"),
             (2, 63, 1024, 20, "clustered"), (2, 63, 512, 20, "dups"), (32, 63, 1024, 20, "randn"), (4, 63, 4096, 20, "randn"),
             (32, 63, 1024, 20, "smooth")]
    for B, D, N, k, kind in cases:
        x = torch.randn(B, D, N, generator=gen)
        if kind == "clustered":
            x = x * 1e-3 + torch.randn(B, D, 1, generator=gen) * 5
        elif kind == "dups":
            x[:, :, N // 2:] = x[:, :, :N // 2]
        elif kind == "smooth":      # features correlated with position, like real EdgeConv activations
            base = torch.randn(B, 3, N, generator=gen)
            w = torch.randn(D, 3, generator=gen)
            x = torch.tanh(torch.einsum("dc,bcn->bdn", w, base)) + 0.05 * x
        xd = x.to(dev)
        st = {}
        ia, va = hb.knn(xd, k, return_values=True, method="auto", stats=st)
        ie, ve = hb.knn(xd, k, return_values=True, method="ffma")
        torch.cuda.synchronize()
        bad_rows = (ia != ie).any(-1)
        vbad = (va.view(torch.int32) != ve.view(torch.int32)).any(-1)
        print(f"B={B} D={D} N={N} k={k} {kind:9s} idx-mismatch rows {int(bad_rows.sum())}/{B * N}  val-mismatch rows {int(vbad.sum())}"
              f"  fallback rows {st['fallback_rows']}", flush=True)
        if bad_rows.any():
            b, i = bad_rows.nonzero()[0].tolist()
            print("   first bad row", b, i, "\n   auto", ia[b, i].tolist(), "\n   ffma", ie[b, i].tolist(),
                  "\n   auto v", va[b, i].tolist(), "\n   ffma v", ve[b, i].tolist())
        if B >= 4:
            print(f"   time auto {timeit(lambda: hb.knn(xd, k)):8.1f} us   ffma {timeit(lambda: hb.knn(xd, k, method='ffma')):8.1f} us")


if __name__ == "__main__":
    main()
