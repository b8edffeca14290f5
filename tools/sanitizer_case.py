"""Small, ragged cases of every kernel added in round 2 (fused EdgeConv fwd+bwd, kNN second chance and list overflow, parallel
complete linkage with a tied cloud, large-N edge backward scatter, input rotation): the command line meant for
`compute-sanitizer --tool memcheck python tools/sanitizer_case.py`.  compute-sanitizer is CLOSED on this pool (gpurun answers
"runs under it have left GPUs needing a reset"), so this round it only ran plain; the bounds of the new kernels were reviewed by
hand."""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import hpcs_b200 as hb, bench
from hpcs_b200.edgeconv import edgeconv
from test_gpu_edgeconv import VNConv, param_list
torch.manual_seed(0)
# fused layers, small ragged shapes, train mode, fwd+bwd (C=21 two-conv, C=1 two-conv, one-conv)
for C, two, N, k, B in ((21, True, 77, 7, 2), (1, True, 130, 20, 1), (21, False, 50, 10, 3)):
    convs = [VNConv(2 * C).cuda().train()] + ([VNConv(21).cuda().train()] if two else [])
    x = torch.randn(B, C, 3, N, device="cuda", requires_grad=True)
    y = edgeconv(x, k, convs[0], convs[1] if two else None)
    torch.autograd.grad(y.sum(), [x] + param_list(convs))
# kNN second chance + overflow
x = bench.clustered_features(2, seed=5).cuda()
st = {}; hb.knn(x, 20, stats=st); print("knn", st)
gen = torch.Generator().manual_seed(9)
tight = torch.cat([torch.randn(1, 63, 1).expand(1, 63, 400) + 1e-3 * torch.randn(1, 63, 400, generator=gen), torch.randn(1, 63, 112, generator=gen)], dim=2).contiguous().cuda()
st = {}; hb.knn(tight, 20, stats=st); print("knn tight", st)
# complete linkage rounds (+ tie redo), edge bwd scatter, rotate, sampler state
e = torch.randn(3, 200, 32, device="cuda"); e[1, 5] = e[1, 9]
Z = hb.decode_linkage_batch(e, torch.tensor([0.2], device="cuda"), "complete"); print("Z", Z.shape)
xx = torch.randn(1, 2, 3, 12000, device="cuda"); idx = hb.knn(xx.view(1, 6, 12000), 4)
from hpcs_b200 import graph as hg
g = torch.randn(1, 4, 3, 12000, 4, device="cuda"); print(hg.edge_features_backward(g, xx, idx).shape)
print(hb.rotate_points(torch.randn(2, 100, 3, device="cuda"), "so3").shape)
torch.cuda.synchronize(); print("done")
