/* hpcs_b200 -- C ABI of the B200 (sm_100a) hot path for HPCS.
 *
 * The reference (TheCrossProduct/HPCS) is pure Python; it has no FFI.  Its plugin surface for
 * this path is the set of Python signatures listed below, and each entry point here is what a
 * binding for that signature calls (see INTEGRATION.md for the ctypes stubs).  Paths are
 * relative to the reference root.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all
 *     buffers including workspaces (query the size with the matching *_workspace_bytes);
 *   - `stream` is a cudaStream_t passed as void*; calls are stream-ordered, never synchronise the
 *     device, allocate nothing, and keep no global mutable state (re-entrant across streams);
 *   - return 0 on success, non-zero otherwise; hpcs_last_error() returns a thread-local message;
 *   - kernels exist for sm_100a only; there is no CPU fallback.
 */
#ifndef HPCS_B200_H
#define HPCS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: + hpcs_edge_rev_build, hpcs_edge_feat_bwd_prebuilt_f32, hpcs_triplet_sample_i32, hpcs_fcluster_maxclust_i32,
 *    hpcs_cut_scores_f64 (additions only; every version-1 entry point is unchanged)
 * 3: + hpcs_triplet_sample_state_i32, hpcs_rotate_points_f32, hpcs_one_hot_f32, hpcs_peak_probe,
 *    hpcs_vn_point_linear_f32, hpcs_edgeconv_* (round 2; additions only)
 */
#define HPCS_ABI_VERSION 4

enum {
    HPCS_OK = 0,
    HPCS_ERR_ARG = 1,        /* bad shape / null pointer / unsupported size */
    HPCS_ERR_WORKSPACE = 2,  /* workspace too small */
    HPCS_ERR_CUDA = 3,       /* launch or runtime error, see hpcs_last_error() */
    HPCS_ERR_DEVICE = 4      /* current device is not compute capability 10.x */
};

int hpcs_abi_version(void);
const char* hpcs_last_error(void);
/* 0 when the current CUDA device can run this library (CC 10.x), HPCS_ERR_DEVICE otherwise. */
int hpcs_device_check(void);
/* Number of kernels this library has launched in the calling process (monotonic counter). */
uint64_t hpcs_launch_count(void);

/* ---- part 1: kNN graph + edge features --------------------------------------------------
 * knn(x, k)                       hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:4-10
 *   x[B,D,N] fp32 -> idx[B,N,k] int64; k largest of -|xi-xj|^2 per row, best first, self
 *   included.  Canonical arithmetic and tie-break (lower index first): see oracle/knn_canonical.c.
 *   Optional `val` [B,N,k] fp32 receives the selected values (may be NULL). */
size_t hpcs_knn_workspace_bytes(int B, int D, int N, int k);
int hpcs_knn_f32(const float* x, int B, int D, int N, int k, int64_t* idx, float* val,
                 void* ws, size_t ws_bytes, void* stream);
/* Same contract, always the all-FFMA exact kernel.  hpcs_knn_f32 picks, for 16 <= D <= 63 and
 * 128 <= N <= 16384, the tensor-core path (tcgen05 TF32 Gram tiles -> candidate lists -> exact fp32
 * re-rank with a proven-safe test and an exact redo of the rows that fail it); both return the
 * same bits, which the parity tests check against each other and against the oracle. */
int hpcs_knn_ffma_f32(const float* x, int B, int D, int N, int k, int64_t* idx, float* val,
                      void* ws, size_t ws_bytes, void* stream);
/* Statistics: number of rows the last hpcs_knn_f32 call on this workspace redid with the exact
 * fallback (0 when the tensor-core path does not apply).  Synchronises `stream`; rows_host is a
 * HOST pointer. */
int hpcs_knn_fallback_rows(const void* ws, size_t ws_bytes, int B, int D, int N, int k, void* stream,
                           int* rows_host);
/* stats_host[2] (HOST): {rows redone by the exact kernels, rows that took the second chance}.  A row whose 32-entry candidate
 * list cannot be proven (more neighbours inside the TF32 error bound than the list has slack: tight clusters) first gets a
 * second tensor-core pass that collects every candidate above (k-th exact value)/2 - bound -- a rigorous superset of its true
 * k nearest -- evaluated exactly; only lists that overflow go on to the exact redo.  Synchronises `stream`. */
int hpcs_knn_path_stats(const void* ws, size_t ws_bytes, int B, int D, int N, int k, void* stream, int* stats_host);

/* get_graph_feature(x, k, idx)    hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:13-41
 * get_graph_feature_cross         hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:44-69   (cross != 0)
 *   x[B,C,3,N] (== [B,3C,N]) fp32, idx[B,N,k] int64 -> out[B,(2|3)C,3,N,k] fp32 contiguous:
 *   channels [0,C) x_j - x_i, [C,2C) x_i, [2C,3C) cross(x_j, x_i). */
int hpcs_edge_feat_fwd_f32(const float* x, const int64_t* idx, int B, int C, int N, int k, int cross,
                           float* out, void* stream);
/* backward of the above wrt x: gx[B,C,3,N] (overwritten).  Deterministic: the scatter is turned
 * into a gather through a reverse (target -> sources) CSR built in the workspace. */
size_t hpcs_edge_feat_bwd_workspace_bytes(int B, int N, int k);
/* 1 when hpcs_edge_feat_bwd_f32 will take the persistent TMA gather for this shape (no cross term, N <= 2048,
 * N*k <= 65535, N*k % 4 == 0, 16-byte aligned gout, at least one N*k plane fits shared memory: N*k <= ~45000); that path sums in a fixed
 * order (bitwise repeatable).  Other shapes (any N) scatter with fp32 reductions: results agree to rounding, not bit for bit
 * run to run; the cross variant keeps the reverse-CSR gather (N up to ~11000). */
int hpcs_edge_feat_bwd_is_fast(const float* gout, int N, int k, int cross);
int hpcs_edge_feat_bwd_f32(const float* gout, const float* x, const int64_t* idx, int B, int C, int N,
                           int k, int cross, float* gx, void* ws, size_t ws_bytes, void* stream);
/* The reverse graph depends on idx only, so a caller that will need the backward can build it while the forward
 * runs (hpcs_b200/graph.py launches it on a second stream next to hpcs_edge_feat_fwd_f32, where the build's
 * latency-bound 30 us hide behind an HBM-bound kernel) and hand the filled workspace to the backward:
 *   hpcs_edge_rev_build              fills ws for (idx, B, N, k); only for shapes hpcs_edge_feat_bwd_is_fast accepts
 *                                    (pass any 16-byte aligned pointer as gout there), HPCS_ERR_ARG otherwise.
 *   hpcs_edge_feat_bwd_prebuilt_f32  same result as hpcs_edge_feat_bwd_f32, skipping the build when the persistent
 *                                    gather applies; falls back to the full call (rebuilding in ws) otherwise. */
int hpcs_edge_rev_build(const int64_t* idx, int B, int N, int k, void* ws, size_t ws_bytes, void* stream);
int hpcs_edge_feat_bwd_prebuilt_f32(const float* gout, const float* x, const int64_t* idx, int B, int C, int N,
                                    int k, int cross, float* gx, void* ws, size_t ws_bytes, void* stream);

/* ---- part 2: Poincare-ball triplet objective -------------------------------------------------
 * MetricHyperbolicLoss.compute_hyp    hpcs/loss/ultrametric_loss.py:57-93 (+ :139-143)
 * RandomTripletMarginMiner.mine (filter part)  hpcs/miner/triplet_margin_miner.py:16-38
 * CosineSimilarity                    hpcs/distances/cosine.py:4-16
 * hyp_lca(.., return_coord=False) on equal-radius rows   hpcs/distances/lca.py:37-52
 *   x[n,D] fp32, triplets (a,p,ng)[T0] int64 as produced by the sampler, scale dev[1],
 *   filter_mode: 0 keep all, 1 'easy' (sim(a,p)-sim(a,n) > margin), 2 'semihard'
 *   (0 < gap <= margin), 3 'hard' (gap <= margin and gap <= 0), 4 gap <= margin.
 *   Writes loss dev[1] fp32 and kept dev[1] int64.  With need_grad != 0 the same pass also
 *   accumulates the gradient state in `ws`, consumed by hpcs_hyp_triplet_bwd_f32.
 *   The (B.N)^2 similarity matrix of the reference is never formed. */
size_t hpcs_hyp_triplet_workspace_bytes(int64_t n, int D);
int hpcs_hyp_triplet_fwd_f32(const float* x, int64_t n, int D, const int64_t* a, const int64_t* p,
                             const int64_t* ng, int64_t T0, const float* scale, float temperature,
                             int filter_mode, float margin, int need_grad, float* loss, int64_t* kept,
                             void* ws, size_t ws_bytes, void* stream);
/* same, triplet indices as int32 (the reference samples int64 on the host; int32 halves the 39 MB/step upload) */
int hpcs_hyp_triplet_fwd_i32_f32(const float* x, int64_t n, int D, const int32_t* a, const int32_t* p,
                                 const int32_t* ng, int64_t T0, const float* scale, float temperature,
                                 int filter_mode, float margin, int need_grad, float* loss, int64_t* kept,
                                 void* ws, size_t ws_bytes, void* stream);
int hpcs_hyp_triplet_bwd_f32(const float* gloss, const float* x, int64_t n, int D, const float* scale,
                             const void* ws, size_t ws_bytes, float* gx, float* gscale, void* stream);
/* the miner's filter alone: keep[T0] uint8 (1 = triplet survives). */
int hpcs_triplet_filter_f32(const float* x, int64_t n, int D, const int64_t* a, const int64_t* p,
                            const int64_t* ng, int64_t T0, int filter_mode, float margin, uint8_t* keep,
                            void* ws, size_t ws_bytes, void* stream);
int hpcs_triplet_filter_i32_f32(const float* x, int64_t n, int D, const int32_t* a, const int32_t* p,
                                const int32_t* ng, int64_t T0, int filter_mode, float margin, uint8_t* keep,
                                void* ws, size_t ws_bytes, void* stream);

/* get_balanced_random_triplet_indices on the device   hpcs/miner/loss_and_miner_utils.py:7-75   
 *   order[n] int32: stable argsort of the labels; seg[4][L] int64: per label with >= 2 members and >= 1 non-member, in
 *   ascending label order: {start in order, members m_l, k_l = int(t_per_anchor (max_count / m_l)^fraction), first
 *   triplet}; T0 = sum m_l k_l.  Writes a,p,ng[T0] int32: the anchors exactly as the reference emits them (label by
 *   label, member by member, k_l times each), positives uniform over the other members, negatives uniform over the
 *   non-members, drawn from Philox4x32-10 keyed by (seed, triplet number) -- same distribution as the reference
 *   sampler, not the same draws (the host sampler keeps RNG parity and stays the default). */
int hpcs_triplet_sample_i32(const int32_t* order, int64_t n, const int64_t* seg, int L, int64_t T0, uint64_t seed,
                            int32_t* a, int32_t* p, int32_t* ng, void* stream);

/* Same sampler with its key in device memory: state[3] uint64 = {seed, step, 0}.  The Philox key is (seed, step); the
 * launch advances `step` itself when its last block retires, so a CUDA graph that captured the call draws NEW triplets on
 * every replay (a by-value seed would be frozen into the graph).  state[2] is scratch and must start at 0. */
int hpcs_triplet_sample_state_i32(const int32_t* order, int64_t n, const int64_t* seg, int L, int64_t T0, uint64_t* state,
                                  int32_t* a, int32_t* p, int32_t* ng, void* stream);

/* hyp_lca(a, b, return_coord)         hpcs/distances/lca.py:37-52 (general, unequal norms)
 *   a,b[T,D] fp32 -> out[T,D] (return_coord) or out[T,1] = 2 artanh(|proj|); scalar chain in fp64. */
int hpcs_hyp_lca_fwd_f32(const float* a, const float* b, int64_t T, int D, int return_coord, float* out,
                         void* stream);
int hpcs_hyp_lca_bwd_f32(const float* gout, const float* a, const float* b, int64_t T, int D,
                         int return_coord, float* ga, float* gb, void* stream);

/* ExpMap.forward = expmap_1(u, 0)     hpcs/nn/hyperbolic/hyp_embed.py:6-10, hpcs/utils/poincare.py:50-54 */
int hpcs_expmap0_fwd_f32(const float* u, int64_t rows, int D, float* y, void* stream);
int hpcs_expmap0_bwd_f32(const float* gy, const float* u, int64_t rows, int D, float* gu, void* stream);

/* ---- part 3: linkage decode --------------------------------------------------------------------
 * BaseSimilarityHypHC._decode_linkage  hpcs/models/base_hyp_hc.py:81-86
 *   leaves = project(normalize(x) * clamp(scale, 1e-4, 1))  (fp32; hpcs_leaves_f32), then
 *   scipy.cluster.hierarchy.linkage(leaves, method, 'cosine'): fp64 cosine distances,
 *   method 0 = 'single' (Prim MST), 1 = 'complete' (nearest-neighbour chain), stable sort by
 *   height, union-find relabel.  Z[B,N-1,4] fp64 in scipy linkage format, one dendrogram per
 *   cloud; the B clouds are processed by one launch sequence. */
int hpcs_leaves_f32(const float* x, int64_t rows, int D, const float* scale, float* leaves, void* stream);
size_t hpcs_linkage_workspace_bytes(int B, int N, int D, int method);
/* diagnostics: byte offset, inside the workspace, of the [B][16] int counters the parallel complete-linkage rounds leave behind
 * ({-, -, merges, rounds, microseconds in: snapshot, row minima, pairing, scratch rows, row/column update, barriers}); 0 when
 * that path does not apply to the shape */
size_t hpcs_linkage_debug_counters_offset(int B, int N, int method);
int hpcs_linkage_f64(const float* leaves, int B, int N, int D, int method, double* Z,
                     void* ws, size_t ws_bytes, void* stream);

/* scipy.cluster.hierarchy.fcluster(Z, k, criterion='maxclust') for a batch   (hpcs/utils/scores.py:151)
 *   Z[B,N-1,4] fp64 in scipy linkage format with non-decreasing heights (what hpcs_linkage_f64 writes), ks[K] device
 *   ints, k_max = max(ks) <= 256 (host copy, for validation) -> labels[B,K,N] int32, 1-based, with scipy's exact cluster
 *   numbering (depth-first, left child first; subtrees numbered on entry, singleton leaves on the way out; k >= N:
 *   label = point index + 1); tied heights are never split, so ties can give fewer than k clusters, as in scipy. */
int hpcs_fcluster_maxclust_i32(const double* Z, int B, int N, const int* ks, int K, int k_max, int32_t* labels,
                               void* stream);

/* the model-selection indices of get_optimal_k   hpcs/utils/scores.py:152-171   (second half)
 *   labels[B,K,N] from hpcs_fcluster_maxclust_i32 (1-based), ytrue[B,N] ground-truth parts remapped to 0..n_true[b]-1
 *   (scores.py:126-139), ks[K] device ints -> scores[B,K] fp64; -1 where k > n_true[b] + extra (the reference stops there).
 *   index 0 = 'iou' (base_hyp_hc.py:198): float32 IoU matrix, first best cluster per part, later parts overwrite,
 *   agreements / (2N - agreements);  index 1 = 'ri': sklearn's adjusted_rand_score (pair counts in 64-bit integers).
 *   t_cap >= max n_true, k_max >= max ks (host copies, they size shared memory). */
int hpcs_cut_scores_f64(const int32_t* labels, const int32_t* ytrue, const int32_t* n_true, const int* ks, int B, int K,
                        int N, int t_cap, int k_max, int extra, int index, double* scores, void* stream);

/* ---- per-step input pipeline on the device    ------------------------------------------------
 * ShapeNetHypHC._forward / PartNetHypHC._forward   hpcs/models/shapenet_hyp_hc.py:63-73, partnet_hyp_hc.py:82-95
 *   The reference rotates the batch on the host with pytorch3d, uploads it and transposes.  Here: pts[B,N,3] fp32 (device)
 *   -> out[B,3,N] = (pts @ R_b)^T, the backbone's layout, in one pass.  mode 0: no rotation (params unused);
 *   1: params = R[B,3,3] (row-vector convention, out = p @ R);  2 ('so3'): params = o[B,4], the randn(B,4) draws of
 *   pytorch3d.transforms.random_rotations, turned into unit quaternions and matrices on the device;  3 ('z'): params =
 *   u[B], the rand(B) draws of RotateAxisAngle(angle=u*360, axis='Z', degrees=True).  rot_out (optional) receives the
 *   matrices R[B,3,3] actually applied. */
int hpcs_rotate_points_f32(const float* pts, const float* params, int mode, int B, int N, float* out, float* rot_out,
                           void* stream);
/* to_categorical(y, num_classes)   hpcs/utils/data.py:24-29: y[rows] int64 -> out[rows, num_classes] fp32 one-hot
 * (a row of zeros for a label outside [0, num_classes), where torch.eye indexing would raise). */
int hpcs_one_hot_f32(const int64_t* y, int64_t rows, int num_classes, float* out, void* stream);
/* MetricHyperbolicLoss.get_logits(embeddings, labels)   hpcs/loss/ultrametric_loss.py:95-112, called twice per step for the
 * accuracy / IoU metrics (hpcs/models/base_hyp_hc.py:88-99) next to the loss's own evaluation.  The reference assembles it from
 * pytorch-metric-learning CosFaceLoss pieces (get_cosine, get_target_mask, a boolean-mask gather that synchronises the host,
 * modify_cosine_of_target_classes, scale_logits); here one forward-only kernel:
 *   emb[n,D] fp32, W[D,classes] fp32 (the loss's weight), labels[n] int64 ->
 *   logits[n,classes] = scale * (cos(emb_i, W_c) - margin * [labels_i == c]),  rows and columns normalised like F.normalize. */
int hpcs_cosface_logits_f32(const float* emb, const float* W, const int64_t* labels, int64_t n, int D, int classes, float margin,
                            float scale, float* logits, void* stream);

/* ---- fused EdgeConv layer    -------------------------------------------------------------------------
 * x = get_graph_feature(x, k); x = convA(x); [x = convB(x);] x = mean_pool(x)
 *   hpcs/nn/dgcnn/vn_dgcnn_partseg.py:65-68,70-73,75-77; VNLinearLeakyReLU hpcs/nn/dgcnn/utils/vn_layers.py:48-77 (with its
 *   VNBatchNorm :112-132), mean_pool :152-153.  The [B,2C,3,N,k] edge tensor is never formed (see csrc/edgeconv.cu).
 *
 * hpcs_vn_point_linear_f32: the layer's first Linear split into per-point maps.  x[B,C,3,N]; W4[4][21][C] = {Wa_feat, Wa_dir,
 *   Wb_feat - Wa_feat, Wb_dir - Wa_dir} with W = [Wa | Wb] the [21, 2C] map_to_feat / map_to_dir weights of convA
 *   -> UU[B*N][128], VV[B*N][128] rows [feat(o*3+c) 63 | 0 | dir 63 | 0].
 * coef: DEVICE buffer of hpcs_edgeconv_coef_floats() floats: per stage s (0 = convA, 1 = convB) six 21-vectors at
 *   [s*126 + which*21]: a, b (BatchNorm on the norm folded to y = a r + b), mu, rstd (rhat = (r - mu) rstd), s1m = mean(gy),
 *   s2m = mean(gy rhat) (backward, training mode; 0 in eval mode); convB weights [21][24] (rows zero-padded) at float
 *   256 (map_to_feat) and 760 (map_to_dir); for a C = 1 layer convA's own weights [4][21] = {Wa_feat, Wa_dir, Wb_feat, Wb_dir}
 *   at float 1264.  Device-resident so that batch statistics never visit the host.
 * x_direct: NULL, or x[B,1,3,N] for a C = 1 layer: the first conv is then evaluated per edge from the coordinates themselves,
 *   p = Wa (x_j - x_i) + Wb x_i as the reference does (no |x| vs |x_j - x_i| cancellation, a 12-byte gather); UU / VV unused.
 * hpcs_edgeconv_fwd_f32: stages 1 | 2.  mode 0: stats[21][2] (fp64) += sum r, sum r^2 of the convA norms over all B*N*k
 *   edges; mode 1: same for the convB norms (needs convA's a, b); mode 2: out[B,21,3,N]; with ysum/yrsum != NULL also the
 *   per-point sums [B*N][63] of the coefficients that make the last stage's BatchNorm backward sums a per-point dot product.
 * hpcs_edgeconv_bwd_stage2_f32 (two-stage layers): G[B,21,3,N] = d loss / d out -> gO1 (scratch of hpcs_edgeconv_scratch_floats
 *   floats, private tile-major layout: the gradient wrt convA's output per edge), dW2[21][2][21] += (out, {feat, dir}, in), stats1[21][2] (fp64) += sum gy, sum gy rhat of convA.
 * hpcs_edgeconv_bwd_stage1_f32: gO1 (or NULL for a one-stage layer: then gO1 = G[n]/k) -> gUU[B*N][128] += (caller zeroes),
 *   gVV[B*N][128] = (floats 63 and 127 of a gVV row are not written); the caller contracts them with x and W4. */
/* BatchNorm bookkeeping of a conv on the device (no host round trip of batch statistics):
 *   hpcs_edgeconv_bn_fold_f32: stats[21][2] (fp64: sum r, sum r^2 over M norms; training) or the running buffers (eval) -> folded
 *     a, b, mu, rstd in coef for `stage`; in training the running buffers are updated like nn.BatchNorm2d (momentum < 0: cumulative
 *     average over *num_batches_tracked, which the caller has already incremented).
 *   hpcs_edgeconv_bn_sums_f32: BatchNorm-backward sums of the LAST conv from G and the forward's ysum / yrsum -> s1m, s2m in coef
 *     (training), d gamma, d beta; `sums` is fp64 scratch [21][2].  hpcs_edgeconv_bn_sums_finish_f32: the same last step for sums
 *     accumulated by hpcs_edgeconv_bwd_stage2_f32 (the first conv of a two-conv layer). */
int hpcs_edgeconv_bn_fold_f32(const double* stats, int64_t M, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, const int64_t* num_batches_tracked, float momentum, float eps, int training,
                              float* coef, int stage, void* stream);
int hpcs_edgeconv_bn_sums_f32(const float* G, const float* ysum, const float* yrsum, int B, int N, int k, int training, double* sums,
                              float* coef, int stage, float* dgamma, float* dbeta, void* stream);
int hpcs_edgeconv_bn_sums_finish_f32(const double* sums, int64_t M, int training, float* coef, int stage, float* dgamma, float* dbeta,
                                     void* stream);
int hpcs_edgeconv_coef_floats(void);
size_t hpcs_edgeconv_scratch_floats(int B, int N, int k);   /* floats of the gO1 scratch between the two backward kernels */
int hpcs_vn_point_linear_f32(const float* x, const float* W4, int B, int C, int N, float* UU, float* VV, void* stream);
/* its backward: gUU, gVV [B*N][128] -> gx[B,C,3,N] (overwritten) and dW4[4][21][C] += (the caller zeroes it) */
int hpcs_vn_point_linear_bwd_f32(const float* gUU, const float* gVV, const float* x, const float* W4, int B, int C, int N, float* gx,
                                 float* dW4, void* stream);
int hpcs_edgeconv_fwd_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k, int stages,
                          const float* coef, const float* x_direct, int mode, double* stats, float* out, float* ysum,
                          float* yrsum, void* stream);
int hpcs_edgeconv_bwd_stage2_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k, const float* coef,
                                 const float* x_direct, const float* G, float* gO1, float* dW2, double* stats1, void* stream);
int hpcs_edgeconv_bwd_stage1_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k, const float* coef,
                                 const float* x_direct, const float* gO1, const float* G, float* gUU, float* gVV, void* stream);

/* ---- measurement support: pipe / cache peak probes (bench.py; no reference counterpart) ------------------------------
 * The roofline denominators MEASURED_PEAKS.json lacks (SURVEY 8d asks for them): which = 0 packed fp32 FMA (FFMA2),
 * 1 scalar fp32 FMA, 2 ALU pipe (fp32 min/max), 3 fp64 FMA, 4 fp64 tensor cores (mma.sync m8n8k4), 5 TF32 tcgen05.mma
 * (M=128, N=256, operands resident in shared memory), 6 / 7 random 128- / 256-byte row gathers from `table` (device memory,
 * size it to stay L2-resident).  One launch of `iters` loop trips; *work_host (HOST pointer, may be NULL) receives the flops
 * (bytes for 6, 7) the launch performs, the caller times it.  `out` is one device float (never written in practice). */
int hpcs_peak_probe(int which, int iters, const void* table, size_t table_bytes, float* out, double* work_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPCS_B200_H */
