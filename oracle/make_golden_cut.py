"""Generate ``tests/golden/optimal_k.npz`` by executing the UNMODIFIED reference ``get_optimal_k``
(hpcs/utils/scores.py:141-177, index='iou': the call of base_hyp_hc.py:198) in this container.

TEST INFRASTRUCTURE ONLY.  ``python -m oracle.make_golden_cut``; the outputs are small and committed, because
``/root/reference`` does not exist on the GPU box.  Inputs: seeded Gaussian-mixture embeddings, ground-truth part labels
with gaps in the label ids (remap_labels must be exercised), dendrograms from scipy (single and complete linkage,
cosine metric, like _decode_linkage)."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    sys.path.insert(0, "/root/reference")
    from hpcs.utils.scores import get_optimal_k            # the reference, unmodified
    from scipy.cluster.hierarchy import linkage
    warnings.filterwarnings("ignore")
    rng = np.random.default_rng(11)
    rec = {}
    cases = [(60, 3, "single", 0.15), (200, 4, "complete", 0.3), (300, 6, "single", 0.5), (256, 2, "complete", 0.2),
             (150, 5, "complete", 0.9), (200, 4, "complete", 2.0), (300, 5, "single", 1.2), (180, 3, "complete", 3.0),
             (120, 4, "single", 0.8)]
    for ci, (n, parts, method, noise) in enumerate(cases):
        cen = rng.standard_normal((parts, 16))
        lab = rng.integers(0, parts, n)
        x = cen[lab] + noise * rng.standard_normal((n, 16))
        if ci == len(cases) - 1:
            x[60:90] = x[:30]                               # duplicated points: tied merge heights
        y = torch.from_numpy(lab * 3 + 2)                   # label ids with gaps: 2, 5, 8, ...
        Z = linkage(x.astype(np.float32), method=method, metric="cosine")
        pred, k, score = get_optimal_k(y, Z, "iou")
        rec[f"y{ci}"], rec[f"Z{ci}"] = y.numpy(), Z
        rec[f"pred{ci}"], rec[f"k{ci}"], rec[f"score{ci}"] = np.asarray(pred), np.int64(k), np.float64(score)
        rpred, rk, rscore = get_optimal_k(y, Z, "ri")       # adjusted Rand index variant (viz.py:489)
        rec[f"ri_pred{ci}"] = np.asarray(rpred) if rpred is not None else np.full(n, -1, dtype=np.int32)
        rec[f"ri_k{ci}"], rec[f"ri_score{ci}"] = np.int64(rk), np.float64(rscore)
        print(ci, n, parts, method, "iou: best k", k, "score", float(score), "| ri: best k", rk, "score", float(rscore))
    np.savez_compressed(os.path.join(OUT, "optimal_k.npz"), n_cases=np.int64(len(cases)), **rec)


if __name__ == "__main__":
    main()
