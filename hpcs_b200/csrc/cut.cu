// Flat clusters from a dendrogram on the GPU: scipy.cluster.hierarchy.fcluster(Z, k, criterion='maxclust') for a
// batch of linkage matrices and a list of k (first half).
//
// Where the reference spends its test step once decoding is on the GPU: get_optimal_k (hpcs/utils/scores.py:141-177)
// calls fcluster(linkage_matrix, k, 'maxclust') for k = 1 .. n_true + 4 on every cloud, on the host, after copying Z
// back.  Here Z stays on the device and all (cloud, k) pairs come out of one launch, with scipy's exact cluster
// NUMBERING (get_optimal_k returns the label array, so the ids matter, not just the partition).
//
// scipy's algorithm (scipy/cluster/_hierarchy.pyx, restated):
//   MC[i]  = max merge height in the subtree of row i  (= Z[i,2] for the monotone Z of single / complete linkage);
//   cutoff = the smallest merge height at which cutting leaves <= k clusters (a bisection over MC; a run of equal
//            heights is never split, so ties can give fewer than k); k >= n short-cuts to "label = point index + 1";
//   labels = depth-first walk from the root, left child first: a subtree whose MC <= cutoff becomes ONE cluster,
//            numbered when the walk ENTERS it; a leaf hanging directly under a node above the cutoff becomes a singleton
//            cluster, numbered when the walk LEAVES that node (left leaf before right leaf).
// Only the ~k nodes above the cutoff need the walk.  Everything below is "fill a subtree with one id", and the leaves
// of a subtree are a contiguous interval of the dendrogram's leaf order, which does not depend on k: one top-down pass
// per cloud computes every node's interval start, then a (cloud, k) pair costs a walk over <= k nodes and N stores.
#include "common.cuh"

namespace hpcs {

constexpr int kCutMaxK = 256;          // clusters a single cut may produce (stack and cluster list live in shared memory)

// One CTA per cloud.  Z[B][N-1][4] fp64 (scipy linkage format), ks[K] device ints, labels[B][K][N] int32 (1-based).
__global__ void __launch_bounds__(512)
fcluster_maxclust_kernel(const double* __restrict__ Z_all, int N, const int* __restrict__ ks, int K, int* __restrict__ labels_all) {
    extern __shared__ int cs[];
    const int M = N - 1;
    int* left = cs;                       // [M]  child ids (< N: leaf)
    int* right = left + M;                // [M]
    int* size = right + M;                // [M]  leaves under the row
    int* lo = size + M;                   // [2N-1] first position of a node's leaves in the dendrogram's leaf order
    int* order = lo + 2 * N - 1;          // [N]  leaf at a position
    __shared__ int cl_node[kCutMaxK], cl_id[kCutMaxK], stack[kCutMaxK], state[kCutMaxK];
    __shared__ int n_cl;
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const double* Z = Z_all + (size_t)b * M * 4;
    for (int i = tid; i < M; i += nthr) {
        left[i] = (int)Z[(size_t)i * 4 + 0];
        right[i] = (int)Z[(size_t)i * 4 + 1];
        size[i] = (int)Z[(size_t)i * 4 + 3];
    }
    __syncthreads();
    // interval starts, top-down: a serial chain along the tree's depth, rows in descending order see their parent first
    if (tid == 0) {
        lo[N + M - 1] = 0;
        for (int i = M - 1; i >= 0; --i) {
            const int l = left[i], r = right[i], at = lo[N + i];
            lo[l] = at;
            lo[r] = at + (l < N ? 1 : size[l - N]);
        }
    }
    __syncthreads();
    for (int leaf = tid; leaf < N; leaf += nthr) order[lo[leaf]] = leaf;
    __syncthreads();

    for (int ki = 0; ki < K; ++ki) {
        int* out = labels_all + ((size_t)b * K + ki) * N;
        const int k = ks[ki];
        if (k >= N) {                                                  // scipy's shortcut: one cluster per point, by point index
            for (int leaf = tid; leaf < N; leaf += nthr) out[leaf] = leaf + 1;
            continue;                                                  // uniform: k is the same for the whole CTA
        }
        if (tid == 0) {
            // rows with MC <= cutoff: [0, c).  The cut merges the fewest rows that leave <= k clusters; a run of equal
            // heights cannot be split, so ties give fewer clusters than asked for.
            const int j = N - k - 1;                                  // merging rows 0..j leaves k clusters
            int c = 0;
            if (j >= 0) {
                const double hj = Z[(size_t)j * 4 + 2];
                c = j + 1;
                while (c < M && Z[(size_t)c * 4 + 2] == hj) ++c;
            }
            // walk the rows >= c from the root; state: 0 = left not tried, 1 = right not tried, 2 = leaves
            int ncl = 0, next_id = 0, sp = 0;
            auto cluster = [&](int node) { if (ncl < kCutMaxK) { cl_node[ncl] = node; cl_id[ncl] = ++next_id; ++ncl; } else ++next_id; };
            if (M - 1 < c) cluster(N + M - 1);                         // the whole tree is one cluster
            else { stack[0] = M - 1; state[0] = 0; sp = 1; }
            while (sp > 0) {
                const int row = stack[sp - 1];
                const int l = left[row], r = right[row];
                if (state[sp - 1] == 0) {
                    state[sp - 1] = 1;
                    if (l >= N) {
                        if (l - N >= c) { if (sp < kCutMaxK) { stack[sp] = l - N; state[sp] = 0; ++sp; } continue; }
                        cluster(l);                                    // entered: numbered now
                    }
                }
                if (state[sp - 1] == 1) {
                    state[sp - 1] = 2;
                    if (r >= N) {
                        if (r - N >= c) { if (sp < kCutMaxK) { stack[sp] = r - N; state[sp] = 0; ++sp; } continue; }
                        cluster(r);
                    }
                }
                if (l < N) cluster(l);                                 // singleton leaves, on the way out
                if (r < N) cluster(r);
                --sp;
            }
            n_cl = ncl;
        }
        __syncthreads();
        for (int q = 0; q < n_cl; ++q) {
            const int node = cl_node[q], id = cl_id[q];
            const int at = lo[node], cnt = node < N ? 1 : size[node - N];
            for (int p = tid; p < cnt; p += nthr) out[order[at + p]] = id;
        }
        __syncthreads();
    }
}

// ---- scoring a cut against the ground-truth parts: the 'iou' index of get_optimal_k (scores.py:152-171) -------------
// One CTA per (k, cloud).  Confusion counts C[true part][cluster] by shared-memory atomics; IoU of every pair as
// float32(C / (|part| + |cluster| - C)) like the reference's float32 score matrix filled from sklearn's float64
// jaccard_score; every true part takes its FIRST best cluster (torch.max); parts are applied in order, so a later part
// overwrites an earlier one that chose the same cluster; score = agreements / (2N - agreements), which is what the
// one-hot logical_and / logical_or ratio of the reference evaluates to.  k > n_true + extra is not scored (-1).
// mode 1 ('ri', scores.py:154-159): sklearn's adjusted_rand_score from the same counts -- pair confusion matrix in
// 64-bit integers (sum of squares, row / column weighted sums), one double division at the end, 1.0 when both
// disagreement counts are zero.
__global__ void __launch_bounds__(256)
cut_iou_score_kernel(const int* __restrict__ labels_all, const int* __restrict__ ytrue_all, const int* __restrict__ n_true,
                     const int* __restrict__ ks, int K, int N, int t_cap, int p_cap, int extra, int mode,
                     double* __restrict__ scores) {
    extern __shared__ int cm[];
    int* C = cm;                          // [t_cap][p_cap]
    int* ct = C + t_cap * p_cap;          // [t_cap] part sizes
    int* cp = ct + t_cap;                 // [p_cap] cluster sizes
    int* ind = cp + p_cap;                // [t_cap] chosen cluster of a part
    int* owner = ind + t_cap;             // [p_cap] part that ends up owning a cluster, -1 = none
    const int ki = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, nthr = blockDim.x;
    const int T = n_true[b], k = ks[ki];
    double* out = scores + (size_t)b * K + ki;
    if (k > T + extra || T > t_cap) { if (tid == 0) *out = -1.0; return; }
    const int* lab = labels_all + ((size_t)b * K + ki) * N;
    const int* yt = ytrue_all + (size_t)b * N;
    for (int i = tid; i < t_cap * p_cap + t_cap + p_cap; i += nthr) C[i] = 0;     // C, ct, cp are contiguous
    __syncthreads();
    for (int p = tid; p < N; p += nthr) {
        const int t = yt[p], c = lab[p] - 1;
        if ((unsigned)t < (unsigned)T && (unsigned)c < (unsigned)p_cap) {
            atomicAdd(&C[t * p_cap + c], 1);
            atomicAdd(&ct[t], 1);
            atomicAdd(&cp[c], 1);
        }
    }
    __syncthreads();
    if (mode == 1) {
        if (tid == 0) {
            long long sumsq = 0, wc = 0, wk = 0;
            for (int t = 0; t < T; ++t)
                for (int c = 0; c < p_cap; ++c) {
                    const long long v = C[t * p_cap + c];
                    sumsq += v * v;
                    wc += v * cp[c];
                    wk += v * ct[t];
                }
            const long long n = N, tp = sumsq - n, fp = wc - sumsq, fn = wk - sumsq, tn = n * n - fp - fn - sumsq;
            *out = (fn == 0 && fp == 0) ? 1.0 : (2.0 * (double)(tp * tn - fn * fp)) / (double)((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn));
        }
        return;
    }
    for (int t = tid; t < T; t += nthr) {
        float best = -1.f;
        int bj = 0;
        for (int c = 0; c < p_cap; ++c) {
            if (cp[c] == 0) continue;                                  // labels are 1..P: empty columns lie beyond P
            const int inter = C[t * p_cap + c], uni = ct[t] + cp[c] - inter;
            const float v = uni > 0 ? (float)((double)inter / (double)uni) : 0.f;
            if (v > best) { best = v; bj = c; }                        // strict: the first maximum wins
        }
        ind[t] = bj;
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < p_cap; ++c) owner[c] = -1;
        for (int t = 0; t < T; ++t) owner[ind[t]] = t;
        long long agree = 0;
        for (int c = 0; c < p_cap; ++c) if (owner[c] >= 0) agree += C[owner[c] * p_cap + c];
        *out = (double)agree / (double)(2LL * N - agree);
    }
}

}  // namespace hpcs

extern "C" int hpcs_fcluster_maxclust_i32(const double* Z, int B, int N, const int* ks, int K, int k_max, int32_t* labels,
                                          void* stream) {
    using namespace hpcs;
    if (!Z || !ks || !labels) return fail(HPCS_ERR_ARG, "fcluster: null pointer");
    if (B <= 0 || N < 3 || K <= 0 || B > 65535) return fail(HPCS_ERR_ARG, "fcluster: bad arguments B=%d N=%d K=%d (N >= 3)", B, N, K);
    if (k_max < 1 || k_max > kCutMaxK) return fail(HPCS_ERR_ARG, "fcluster: k must be in [1, %d]", kCutMaxK);
    const size_t smem = ((size_t)3 * (N - 1) + (2 * N - 1) + N) * sizeof(int);
    if (smem > 220 * 1024) return fail(HPCS_ERR_ARG, "fcluster: N=%d too large (max ~9300)", N);
    cudaFuncSetAttribute(fcluster_maxclust_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    fcluster_maxclust_kernel<<<B, 512, smem, as_stream(stream)>>>(Z, N, ks, K, labels);
    return check_launch("fcluster_maxclust_kernel");
}

extern "C" int hpcs_cut_scores_f64(const int32_t* labels, const int32_t* ytrue, const int32_t* n_true, const int* ks, int B,
                                   int K, int N, int t_cap, int k_max, int extra, int index, double* scores, void* stream) {
    using namespace hpcs;
    if (!labels || !ytrue || !n_true || !ks || !scores) return fail(HPCS_ERR_ARG, "cut_scores: null pointer");
    if (B <= 0 || K <= 0 || N <= 0 || t_cap <= 0 || k_max <= 0 || B > 65535 || (index != 0 && index != 1)) return fail(HPCS_ERR_ARG, "cut_scores: bad arguments");
    if (index == 1 && N > 40000) return fail(HPCS_ERR_ARG, "cut_scores: N=%d overflows the 64-bit pair counts of the Rand index", N);
    const size_t smem = ((size_t)t_cap * k_max + 2 * (size_t)t_cap + 2 * (size_t)k_max) * sizeof(int);
    if (smem > 200 * 1024) return fail(HPCS_ERR_ARG, "cut_scores: %d parts x %d clusters do not fit shared memory", t_cap, k_max);
    cudaFuncSetAttribute(cut_iou_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cut_iou_score_kernel<<<dim3(K, B), 256, smem, as_stream(stream)>>>(labels, ytrue, n_true, ks, K, N, t_cap, k_max, extra, index, scores);
    return check_launch("cut_iou_score_kernel");
}
