/* Canonical-order kNN: CPU oracle for the CUDA kNN kernels.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:4-10 (knn): for every cloud b and row i,
 *     pd[i][j] = -|x_i|^2 - (-2 x_i.x_j) - |x_j|^2          (== -|x_i - x_j|^2)
 * and the k largest entries of the row, largest first, self included.
 *
 * The reference leaves the fp32 accumulation order (cuBLAS/MKL) and the order of exact ties
 * (torch.topk) unspecified (SURVEY.md Finding 5).  This file fixes both, and the CUDA kernels
 * reproduce it bit for bit:
 *     sq_i   = fma chain over d = 0..D-1:  sq  = fmaf(x[d][i], x[d][i], sq),  sq  starts at +0
 *     dot_ij = fma chain over d = 0..D-1:  dot = fmaf(x[d][i], x[d][j], dot), dot starts at +0
 *     pd_ij  = fmaf(2, dot_ij, -sq_i) - sq_j        (two roundings, like (-xx - inner) - xx^T)
 *     order  : larger pd first; equal pd -> lower j first.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (fmaf is correctly rounded in libm/hardware).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { float v; int32_t j; } cand_t;

static int better(float va, int32_t ja, float vb, int32_t jb) {
    return (va > vb) || (va == vb && ja < jb);
}

int knn_canonical_f32(const float *x, int B, int D, int N, int k, int64_t *idx, float *val) {
    if (k > N || k <= 0) return 1;
    float *sq = (float *)malloc(sizeof(float) * (size_t)N);
    cand_t *best = (cand_t *)malloc(sizeof(cand_t) * (size_t)k);
    if (!sq || !best) return 2;
    for (int b = 0; b < B; ++b) {
        const float *xb = x + (size_t)b * D * N;
        for (int i = 0; i < N; ++i) {
            float s = 0.0f;
            for (int d = 0; d < D; ++d) s = fmaf(xb[(size_t)d * N + i], xb[(size_t)d * N + i], s);
            sq[i] = s;
        }
        for (int i = 0; i < N; ++i) {
            int filled = 0;
            for (int j = 0; j < N; ++j) {
                float dot = 0.0f;
                for (int d = 0; d < D; ++d) dot = fmaf(xb[(size_t)d * N + i], xb[(size_t)d * N + j], dot);
                float pd = fmaf(2.0f, dot, -sq[i]) - sq[j];
                /* sorted insertion, list kept best-first */
                if (filled == k && !better(pd, j, best[k - 1].v, best[k - 1].j)) continue;
                int pos = filled < k ? filled : k - 1;
                while (pos > 0 && better(pd, j, best[pos - 1].v, best[pos - 1].j)) {
                    best[pos] = best[pos - 1];
                    --pos;
                }
                best[pos].v = pd; best[pos].j = j;
                if (filled < k) ++filled;
            }
            for (int m = 0; m < k; ++m) {
                idx[((size_t)b * N + i) * k + m] = best[m].j;
                if (val) val[((size_t)b * N + i) * k + m] = best[m].v;
            }
        }
    }
    free(sq); free(best);
    return 0;
}
