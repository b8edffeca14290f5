// Batched per-cloud kNN graph construction (exact fp32, canonical order).
//
// Replaces knn() of hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:4-10, which materialises three [B,N,N]
// temporaries plus a batched SGEMM with K = 3 / 63 and a topk.  Here nothing of size N x N ever
// exists: one warp owns QW query rows; the 32 lanes of the warp each take one candidate column of
// the current 32-wide chunk, keep its D features in registers, and evaluate the canonical
// distance against the warp's queries (query features broadcast from shared memory).  The running
// top-k of a query is a sorted list spread over the lanes of the warp (rank r lives in lane r%32,
// slot r/32); a candidate that beats the current k-th value is inserted with two warp shuffles.
//
// Canonical arithmetic (bit-exact with oracle/knn_canonical.c):
//   sq_i  = fma chain over d ascending;  dot_ij = fma chain over d ascending;
//   pd_ij = fmaf(2, dot_ij, -sq_i) - sq_j;   order: larger pd first, ties -> lower j.
// Zero padding of d up to a multiple of 4 does not change any bit of pd (fma(0,0,acc) == acc).
#include <float.h>

#include "common.cuh"
#include "knn_rows.cuh"

namespace hpcs {

// tensor-core path (knn_tc.cu)
bool knn_tc_applicable(int D, int N, int k);
size_t knn_tc_workspace_bytes(int B, int D, int N, int k);
int knn_tc_run(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, void* ws, size_t ws_bytes,
               cudaStream_t st);
int knn_tc_fallback_rows(const void* ws, int B, int D, int N, int k, cudaStream_t st, int* out_host);

// coordinate kNN, warp per row (knn_d3.cu)
bool knn_d3_applicable(int D, int N, int k);
int knn_d3_run(const float* x, int B, int N, int k, int64_t* idx, float* val, cudaStream_t st);

constexpr int kKnnWarps = 8;
constexpr int kRowQueue = 48;          // per-thread FIFO depth of the row-parallel kernel (flush when > 16 pending)

__global__ void knn_sqnorm_kernel(const float* __restrict__ x, int D, int N, float* __restrict__ sq) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const float* xb = x + (size_t)b * D * N;
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
        const float v = __ldg(xb + (size_t)d * N + j);
        s = __fmaf_rn(v, v, s);
    }
    sq[(size_t)b * N + j] = s;
}

// DREG: candidate features held in registers per d-block (multiple of 4).
// QW:   query rows per warp.   SLOTS: ceil(k/32).   STAGED: whole cloud copied to shared memory.
template <int DREG, int QW, int SLOTS, bool STAGED>
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_kernel(const float* __restrict__ x, const float* __restrict__ sq, int D, int Dp, int N, int k,
           int64_t* __restrict__ idx, float* __restrict__ val, const int* __restrict__ gate, int gate_min) {
    extern __shared__ __align__(16) float smem[];
    if (gate && *gate <= gate_min) return;                // whole-grid early exit (tensor-core path: nothing to redo)
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int QB = kKnnWarps * QW;            // queries per CTA
    const int q0 = blockIdx.x * QB;
    const float* xb = x + (size_t)b * D * N;
    const float* sqb = sq + (size_t)b * N;

    float* qs = smem;                             // [QB][Dp], zero padded
    float* xs = smem + QB * Dp;                   // STAGED: [Dp][N] candidates, then [N] norms
    float* sqs = xs + (size_t)Dp * N;

    for (int e = threadIdx.x; e < QB * Dp; e += blockDim.x) {
        const int d = e / QB, qi = e % QB;        // consecutive threads -> consecutive points
        const int i = q0 + qi;
        qs[qi * Dp + d] = (d < D && i < N) ? __ldg(xb + (size_t)d * N + i) : 0.f;
    }
    if (STAGED) {
        for (int e = threadIdx.x; e < Dp * N; e += blockDim.x) {
            const int d = e / N;
            xs[e] = d < D ? __ldg(xb + e) : 0.f;
        }
        for (int e = threadIdx.x; e < N; e += blockDim.x) sqs[e] = __ldg(sqb + e);
    }
    __syncthreads();

    // lane-distributed sorted lists, one per query of this warp
    float lv[QW][SLOTS];
    int li[QW][SLOTS];
    float tau[QW], sqq[QW];
#pragma unroll
    for (int q = 0; q < QW; ++q) {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) { lv[q][s] = -INFINITY; li[q][s] = 0x7fffffff; }
        tau[q] = -INFINITY;
        const int i = q0 + warp * QW + q;
        sqq[q] = i < N ? (STAGED ? sqs[i] : __ldg(sqb + i)) : 0.f;
    }
    const int tau_slot = (k - 1) >> 5, tau_lane = (k - 1) & 31;

    for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < N;
        float acc[QW];
#pragma unroll
        for (int q = 0; q < QW; ++q) acc[q] = 0.f;
        for (int db = 0; db < Dp; db += DREG) {
            float c[DREG];
#pragma unroll
            for (int t = 0; t < DREG; ++t) {
                const int d = db + t;
                float v = 0.f;
                if (valid && d < Dp) v = STAGED ? xs[(size_t)d * N + j] : (d < D ? __ldg(xb + (size_t)d * N + j) : 0.f);
                c[t] = v;
            }
#pragma unroll
            for (int t = 0; t < DREG; t += 4) {
                if (db + t < Dp) {                 // warp-uniform
#pragma unroll
                    for (int q = 0; q < QW; ++q) {
                        const float4 qv = *reinterpret_cast<const float4*>(qs + (warp * QW + q) * Dp + db + t);
                        float a = acc[q];
                        a = __fmaf_rn(qv.x, c[t], a);
                        a = __fmaf_rn(qv.y, c[t + 1], a);
                        a = __fmaf_rn(qv.z, c[t + 2], a);
                        a = __fmaf_rn(qv.w, c[t + 3], a);
                        acc[q] = a;
                    }
                }
            }
        }
        const float sqc = valid ? (STAGED ? sqs[j] : __ldg(sqb + j)) : 0.f;
#pragma unroll
        for (int q = 0; q < QW; ++q) {
            const float t = __fmaf_rn(2.f, acc[q], -sqq[q]);
            const float pd = valid ? __fsub_rn(t, sqc) : -INFINITY;
            unsigned bm = __ballot_sync(kFull, pd > tau[q]);
            while (bm) {
                const int src = __ffs(bm) - 1;
                bm &= bm - 1;
                const float v = __shfl_sync(kFull, pd, src);
                if (v > tau[q]) {                  // tau may have risen since the ballot
                    const int jj = j0 + src;
                    int pos = 0;
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) pos += __popc(__ballot_sync(kFull, lv[q][s] >= v));
#pragma unroll
                    for (int s = SLOTS - 1; s >= 0; --s) {
                        float pv = __shfl_up_sync(kFull, lv[q][s], 1);
                        int pi = __shfl_up_sync(kFull, li[q][s], 1);
                        if (s > 0) {
                            const float cv = __shfl_sync(kFull, lv[q][s - 1], 31);
                            const int ci = __shfl_sync(kFull, li[q][s - 1], 31);
                            if (lane == 0) { pv = cv; pi = ci; }
                        }
                        const int r = s * 32 + lane;
                        if (r > pos) { lv[q][s] = pv; li[q][s] = pi; }
                        else if (r == pos) { lv[q][s] = v; li[q][s] = jj; }
                    }
                    float tv = lv[q][0];
#pragma unroll
                    for (int s = 1; s < SLOTS; ++s) if (s == tau_slot) tv = lv[q][s];
                    tau[q] = __shfl_sync(kFull, tv, tau_lane);
                }
            }
        }
    }

#pragma unroll
    for (int q = 0; q < QW; ++q) {
        const int i = q0 + warp * QW + q;
        if (i >= N) continue;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const int r = s * 32 + lane;
            if (r < k) {
                const int jj = li[q][s];
                const size_t o = ((size_t)b * N + i) * k + r;
                idx[o] = jj < N ? jj : i;
                if (val) val[o] = lv[q][s];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row-parallel exact kernel (the fast path for D = 3 and D = 63, k <= 40)
// ------------------------------------------------------------------------------------------------
// One thread = one query row, its D features in registers.  The CTA streams candidate tiles
// [D][TC] (+ their norms) through shared memory; a warp reads 4 consecutive candidates of one
// feature with a single broadcast LDS.128, so the inner loop is 4 independent canonical fma chains.
// Selection: RowSelector (knn_rows.cuh).  Same canonical arithmetic and order as knn_kernel.
// SPLIT > 1: SPLIT threads share one query row, each scanning every SPLIT-th 32-candidate chunk of a tile (the part
// index is warp-uniform, so the broadcast loads stay broadcasts); their sorted lists are merged at the end with
// the canonical order.  For D = 3 the selection, not the arithmetic, is the cost, and one thread per row leaves
// only 7 warps per SM at B*N = 32768.  Kept as an option: each part re-fills its own top-K, and on B200 that extra
// insertion work outweighs the added warps (see launch_knn_rows), so SPLIT = 1 is what runs.
template <int D, int K, int ROWS, int TC, int SPLIT>
__global__ void __launch_bounds__(ROWS * SPLIT)
knn_rows_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int k,
                int64_t* __restrict__ idx, float* __restrict__ val, const int* __restrict__ gate, int gate_min) {
    constexpr int QCAP = kRowQueue;
    constexpr int THREADS = ROWS * SPLIT;
    if (gate && *gate <= gate_min) return;                // whole-grid early exit (tensor-core path: nothing to redo)
    extern __shared__ __align__(16) float smem[];
    float* cs = smem;                                  // [D][TC]
    float* sqc = cs + D * TC;                          // [TC]
    float* qv = sqc + TC;                              // [QCAP][THREADS]   (after the scan: [SPLIT][K][ROWS] merge lists)
    int* qj = reinterpret_cast<int*>(qv + QCAP * THREADS);
    const int b = blockIdx.y;
    const int r_in = threadIdx.x % ROWS, part = threadIdx.x / ROWS;    // ROWS is a multiple of 32: part is warp-uniform
    const int i = blockIdx.x * ROWS + r_in;            // query row (may be >= N in the last CTA)
    const float* xb = x + (size_t)b * D * N;
    const float* sqb = sq + (size_t)b * N;
    const int iq = i < N ? i : N - 1;                  // idle rows shadow the last one (keeps warps uniform)

    float q[D];
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = __ldg(xb + (size_t)d * N + iq);
    const float nsq = -__ldg(sqb + iq);

    RowSelector<K, QCAP> sel;
    sel.init(qv + threadIdx.x, qj + threadIdx.x, THREADS);

    for (int j0 = 0; j0 < N; j0 += TC) {
        __syncthreads();                               // previous tile fully consumed
        for (int e = threadIdx.x; e < D * TC; e += THREADS) {
            const int d = e / TC, c = e - d * TC;
            cs[e] = j0 + c < N ? __ldg(xb + (size_t)d * N + j0 + c) : 0.f;
        }
        for (int c = threadIdx.x; c < TC; c += THREADS) sqc[c] = j0 + c < N ? __ldg(sqb + j0 + c) : 0.f;
        __syncthreads();
        const int lim = min(TC, N - j0);
        for (int c0 = 32 * part; c0 < lim; c0 += 32 * SPLIT) {
            sel.maybe_flush(32);
#pragma unroll
            for (int c = c0; c < c0 + 32; c += 4) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    const float4 cv = *reinterpret_cast<const float4*>(cs + d * TC + c);
                    a0 = __fmaf_rn(q[d], cv.x, a0);
                    a1 = __fmaf_rn(q[d], cv.y, a1);
                    a2 = __fmaf_rn(q[d], cv.z, a2);
                    a3 = __fmaf_rn(q[d], cv.w, a3);
                }
                const float4 sv = *reinterpret_cast<const float4*>(sqc + c);
                const float p0 = __fsub_rn(__fmaf_rn(2.f, a0, nsq), sv.x);
                const float p1 = __fsub_rn(__fmaf_rn(2.f, a1, nsq), sv.y);
                const float p2 = __fsub_rn(__fmaf_rn(2.f, a2, nsq), sv.z);
                const float p3 = __fsub_rn(__fmaf_rn(2.f, a3, nsq), sv.w);
                const int j = j0 + c;
                if (j + 0 < N) sel.offer(p0, j + 0);
                if (j + 1 < N) sel.offer(p1, j + 1);
                if (j + 2 < N) sel.offer(p2, j + 2);
                if (j + 3 < N) sel.offer(p3, j + 3);
            }
        }
    }
    sel.flush();
    if (SPLIT > 1) {
        // every part publishes its sorted list; part 0 merges the SPLIT heads K times (larger value first, equal
        // values -> lower index: the canonical order, independent of which part saw a candidate)
        __syncthreads();                               // FIFOs are dead: reuse them
        float* mv = qv;                                // [SPLIT][K][ROWS]
        int* mj = qj;
#pragma unroll
        for (int m = 0; m < K; ++m) {
            mv[(part * K + m) * ROWS + r_in] = sel.top.val[m];
            mj[(part * K + m) * ROWS + r_in] = sel.top.idx[m];
        }
        __syncthreads();
        if (part != 0 || i >= N) return;
        int head[SPLIT];
        float hv[SPLIT];
        int hj[SPLIT];
#pragma unroll
        for (int s = 0; s < SPLIT; ++s) { head[s] = 0; hv[s] = mv[(s * K) * ROWS + r_in]; hj[s] = mj[(s * K) * ROWS + r_in]; }
        int64_t* oi = idx + ((size_t)b * N + i) * k;
        float* ov = val ? val + ((size_t)b * N + i) * k : nullptr;
        for (int m = 0; m < k; ++m) {
            int best = 0;
#pragma unroll
            for (int s = 1; s < SPLIT; ++s)
                if (hv[s] > hv[best] || (hv[s] == hv[best] && hj[s] < hj[best])) best = s;
            float bv = hv[0];
            int bj = hj[0];
#pragma unroll
            for (int s = 1; s < SPLIT; ++s) if (s == best) { bv = hv[s]; bj = hj[s]; }
            oi[m] = bj < N ? bj : i;
            if (ov) ov[m] = bv;
#pragma unroll
            for (int s = 0; s < SPLIT; ++s) {
                if (s == best) {
                    ++head[s];
                    const bool more = head[s] < K;
                    hv[s] = more ? mv[(s * K + head[s]) * ROWS + r_in] : -INFINITY;
                    hj[s] = more ? mj[(s * K + head[s]) * ROWS + r_in] : 0x7fffffff;
                }
            }
        }
        return;
    }
    if (i < N) {
        int64_t* oi = idx + ((size_t)b * N + i) * k;
        float* ov = val ? val + ((size_t)b * N + i) * k : nullptr;
#pragma unroll
        for (int m = 0; m < K; ++m) {
            if (m < k) {
                const int jj = sel.top.idx[m];
                oi[m] = jj < N ? jj : i;
                if (ov) ov[m] = sel.top.val[m];
            }
        }
    }
}

template <int D, int K>
static int launch_knn_rows(const float* x, const float* sq, int B, int N, int k, int64_t* idx, float* val,
                           const int* gate, int gate_min, cudaStream_t st) {
    constexpr int ROWS = 64;
    constexpr int SPLIT = 1;            // measured at B=32, N=1024, D=3: SPLIT 1 / 2 / 4 -> 68 / 81 / 116 us (the extra inserts cost more than the occupancy gains)
    constexpr int TC = D <= 4 ? 512 : 128;
    constexpr size_t fifo = 2 * (size_t)kRowQueue * ROWS * SPLIT, merge = 2 * (size_t)SPLIT * K * ROWS;
    constexpr size_t smem = ((size_t)D * TC + TC + (fifo > merge ? fifo : merge)) * sizeof(float);
    auto kern = knn_rows_kernel<D, K, ROWS, TC, SPLIT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<dim3((N + ROWS - 1) / ROWS, B), ROWS * SPLIT, smem, st>>>(x, sq, N, k, idx, val, gate, gate_min);
    return check_launch("knn_rows_kernel");
}

template <int D>
static int dispatch_rows(const float* x, const float* sq, int B, int N, int k, int64_t* idx, float* val,
                         const int* gate, int gate_min, cudaStream_t st, bool* handled) {
    *handled = true;
    if (k <= 10) return launch_knn_rows<D, 10>(x, sq, B, N, k, idx, val, gate, gate_min, st);
    if (k <= 20) return launch_knn_rows<D, 20>(x, sq, B, N, k, idx, val, gate, gate_min, st);
    if (k <= 40) return launch_knn_rows<D, 40>(x, sq, B, N, k, idx, val, gate, gate_min, st);
    *handled = false;
    return HPCS_OK;
}

template <int DREG, int QW, int SLOTS>
static int launch_knn(const float* x, const float* sq, int B, int D, int N, int k, int64_t* idx, float* val,
                      const int* gate, int gate_min, cudaStream_t st) {
    const int Dp = (D + 3) / 4 * 4;
    constexpr int QB = kKnnWarps * QW;
    dim3 grid((N + QB - 1) / QB, B), block(kKnnWarps * 32);
    const size_t smem_q = (size_t)QB * Dp * sizeof(float);
    const size_t smem_staged = smem_q + ((size_t)Dp * N + N) * sizeof(float);
    if (smem_staged <= 64 * 1024) {
        auto kern = knn_kernel<DREG, QW, SLOTS, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        kern<<<grid, block, smem_staged, st>>>(x, sq, D, Dp, N, k, idx, val, gate, gate_min);
    } else {
        if (smem_q > 200 * 1024) return fail(HPCS_ERR_ARG, "knn: D=%d too large", D);
        auto kern = knn_kernel<DREG, QW, SLOTS, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
        kern<<<grid, block, smem_q, st>>>(x, sq, D, Dp, N, k, idx, val, gate, gate_min);
    }
    return check_launch("knn_kernel");
}

template <int DREG, int QW>
static int dispatch_slots(const float* x, const float* sq, int B, int D, int N, int k, int64_t* idx, float* val,
                          const int* gate, int gate_min, cudaStream_t st) {
    if (k <= 32) return launch_knn<DREG, QW, 1>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
    if (k <= 64) return launch_knn<DREG, QW, 2>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
    if (k <= 128) return launch_knn<DREG, QW, 4>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
    return fail(HPCS_ERR_ARG, "knn: k=%d > 128 not supported", k);
}

}  // namespace hpcs

namespace hpcs {
// the exact kernels on precomputed norms; gate == nullptr: unconditional
int knn_ffma_gated(const float* x, const float* sq, int B, int D, int N, int k, int64_t* idx, float* val, const int* gate,
                   int gate_min, cudaStream_t st) {
    int rc = HPCS_OK;
    bool handled = false;
    if (D == 3) rc = dispatch_rows<3>(x, sq, B, N, k, idx, val, gate, gate_min, st, &handled);
    else if (D == 63) rc = dispatch_rows<63>(x, sq, B, N, k, idx, val, gate, gate_min, st, &handled);
    if (handled) return rc;
    if (D <= 4) return dispatch_slots<4, 4>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
    if (D <= 32) return dispatch_slots<32, 4>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
    return dispatch_slots<64, 4>(x, sq, B, D, N, k, idx, val, gate, gate_min, st);
}
}  // namespace hpcs

static int knn_ffma(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, void* ws, cudaStream_t st) {
    using namespace hpcs;
    float* sq = static_cast<float*>(ws);
    knn_sqnorm_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(x, D, N, sq);
    int rc = check_launch("knn_sqnorm_kernel");
    if (rc) return rc;
    return knn_ffma_gated(x, sq, B, D, N, k, idx, val, nullptr, 0, st);
}

static int knn_check(const float* x, int B, int D, int N, int k, int64_t* idx, void* ws, size_t ws_bytes) {
    using namespace hpcs;
    if (!x || !idx || !ws) return fail(HPCS_ERR_ARG, "knn: null pointer");
    if (B <= 0 || D <= 0 || N <= 0 || k <= 0 || k > N) return fail(HPCS_ERR_ARG, "knn: bad shape B=%d D=%d N=%d k=%d", B, D, N, k);
    if (ws_bytes < hpcs_knn_workspace_bytes(B, D, N, k)) return fail(HPCS_ERR_WORKSPACE, "knn: workspace too small");
    return HPCS_OK;
}

extern "C" {

size_t hpcs_knn_workspace_bytes(int B, int D, int N, int k) {
    size_t need = hpcs::align_up((size_t)B * N * sizeof(float), 256);
    if (hpcs::knn_tc_applicable(D, N, k)) {
        const size_t tc = hpcs::knn_tc_workspace_bytes(B, D, N, k);
        if (tc > need) need = tc;
    }
    return need;
}

int hpcs_knn_f32(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, void* ws,
                 size_t ws_bytes, void* stream) {
    using namespace hpcs;
    int rc = knn_check(x, B, D, N, k, idx, ws, ws_bytes);
    if (rc) return rc;
    if (knn_tc_applicable(D, N, k)) return knn_tc_run(x, B, D, N, k, idx, val, ws, ws_bytes, as_stream(stream));
    if (knn_d3_applicable(D, N, k)) return knn_d3_run(x, B, N, k, idx, val, as_stream(stream));
    return knn_ffma(x, B, D, N, k, idx, val, ws, as_stream(stream));
}

int hpcs_knn_fallback_rows(const void* ws, size_t ws_bytes, int B, int D, int N, int k, void* stream, int* rows_host) {
    using namespace hpcs;
    if (!ws || !rows_host) return fail(HPCS_ERR_ARG, "knn_fallback_rows: null pointer");
    *rows_host = 0;
    if (!knn_tc_applicable(D, N, k)) return HPCS_OK;
    if (ws_bytes < hpcs_knn_workspace_bytes(B, D, N, k)) return fail(HPCS_ERR_WORKSPACE, "knn_fallback_rows: workspace too small");
    int both[2] = {0, 0};
    const int rc = knn_tc_fallback_rows(ws, B, D, N, k, as_stream(stream), both);
    *rows_host = both[0];
    return rc;
}

int hpcs_knn_path_stats(const void* ws, size_t ws_bytes, int B, int D, int N, int k, void* stream, int* stats_host) {
    using namespace hpcs;
    if (!ws || !stats_host) return fail(HPCS_ERR_ARG, "knn_path_stats: null pointer");
    stats_host[0] = stats_host[1] = 0;
    if (!knn_tc_applicable(D, N, k)) return HPCS_OK;
    if (ws_bytes < hpcs_knn_workspace_bytes(B, D, N, k)) return fail(HPCS_ERR_WORKSPACE, "knn_path_stats: workspace too small");
    return knn_tc_fallback_rows(ws, B, D, N, k, as_stream(stream), stats_host);
}

int hpcs_knn_ffma_f32(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, void* ws,
                      size_t ws_bytes, void* stream) {
    using namespace hpcs;
    int rc = knn_check(x, B, D, N, k, idx, ws, ws_bytes);
    if (rc) return rc;
    return knn_ffma(x, B, D, N, k, idx, val, ws, as_stream(stream));
}

}  // extern "C"
