// Edge-feature gather for vector-neuron EdgeConv layers, forward and backward.
//
// Replaces get_graph_feature / get_graph_feature_cross of
// hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:13-41,44-69, which run index gather + repeat + cat +
// permute().contiguous() (about 1.3 GB of HBM traffic per 63-d layer for 344 MB of algorithmic
// bytes).  Here the output [B,(2|3)C,3,N,k] is written exactly once, in its final layout, with
// 128-bit streaming stores; the only reads are x (staged per vector channel in shared memory,
// 12 KB at N=1024) and idx.
//
// Backward: the scatter-add of the reference (index_put_ with accumulate, i.e. atomics) becomes a
// deterministic gather through a reverse CSR (target -> list of (n', j) sources, ascending) that
// is built once per call in the workspace and reused by all 3C channel planes.  The gradient
// plane of one channel (N*k floats, 80 KB at N=1024,k=20) is staged in shared memory so the
// random reads hit shared memory banks instead of L1/L2 sectors.
#include "common.cuh"

namespace hpcs {

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// grid: (splits, C, B); block 256.  Each CTA stages x[b, c, 0..2, :] and produces, for its share of
// the flattened (n, j) range, the 6 (or 9) output planes of vector channel c.
template <int VEC, bool CROSS>
__global__ void __launch_bounds__(256)
edge_feat_fwd_kernel(const float* __restrict__ x, const int64_t* __restrict__ idx, int C, int N, int k,
                     float* __restrict__ out) {
    extern __shared__ __align__(16) float xs[];            // [3][N]
    const int b = blockIdx.z, c = blockIdx.y;
    const float* xc = x + ((size_t)b * C + c) * 3 * N;
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) xs[e] = __ldg(xc + e);
    __syncthreads();

    const int planes = CROSS ? 3 : 2;
    const size_t plane = (size_t)N * k;                     // one (channel, component) plane
    const size_t total = plane;
    const size_t per_cta = (total / VEC + gridDim.x - 1) / gridDim.x;   // in units of VEC elements
    const size_t g_begin = (size_t)blockIdx.x * per_cta;
    size_t g_end = g_begin + per_cta;
    if (g_end > total / VEC) g_end = total / VEC;
    const int64_t* idb = idx + (size_t)b * plane;
    float* ob = out + (size_t)b * planes * C * 3 * plane;
    float* o_diff = ob + ((size_t)c * 3) * plane;
    float* o_ctr = ob + ((size_t)(C + c) * 3) * plane;
    float* o_crs = ob + ((size_t)(2 * C + c) * 3) * plane;

    for (size_t g = g_begin + threadIdx.x; g < g_end; g += blockDim.x) {
        const size_t e0 = g * VEC;
        int n = (int)(e0 / k);
        int j = (int)(e0 - (size_t)n * k);
        float ctr[3][VEC], nb[3][VEC];
        long long ids[VEC];
        if constexpr (VEC == 4) {
            const longlong2 p0 = __ldg(reinterpret_cast<const longlong2*>(idb + e0));
            const longlong2 p1 = __ldg(reinterpret_cast<const longlong2*>(idb + e0 + 2));
            ids[0] = p0.x; ids[1] = p0.y; ids[2] = p1.x; ids[3] = p1.y;
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) ids[v] = __ldg(idb + e0 + v);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            unsigned m = (unsigned)ids[v];
            m = m < (unsigned)N ? m : (unsigned)(N - 1);   // keep a bad index from faulting
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                ctr[a][v] = xs[a * N + n];
                nb[a][v] = xs[a * N + m];
            }
            if (++j == k) { j = 0; ++n; }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if constexpr (VEC == 4) {
                st_stream(reinterpret_cast<float4*>(o_diff + a * plane + e0),
                          make_float4(nb[a][0] - ctr[a][0], nb[a][1] - ctr[a][1], nb[a][2] - ctr[a][2],
                                      nb[a][3] - ctr[a][3]));
                st_stream(reinterpret_cast<float4*>(o_ctr + a * plane + e0),
                          make_float4(ctr[a][0], ctr[a][1], ctr[a][2], ctr[a][3]));
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    st_stream(o_diff + a * plane + e0 + v, nb[a][v] - ctr[a][v]);
                    st_stream(o_ctr + a * plane + e0 + v, ctr[a][v]);
                }
            }
        }
        if constexpr (CROSS) {
            // cross(nbr, ctr), torch.cross convention; products and difference rounded separately
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
                float r[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    r[v] = __fsub_rn(__fmul_rn(nb[a1][v], ctr[a2][v]), __fmul_rn(nb[a2][v], ctr[a1][v]));
                if constexpr (VEC == 4) {
                    st_stream(reinterpret_cast<float4*>(o_crs + a * plane + e0), make_float4(r[0], r[1], r[2], r[3]));
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) st_stream(o_crs + a * plane + e0 + v, r[v]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: reverse CSR
// ------------------------------------------------------------------------------------------------
// One CTA per cloud.  rev_ptr[b][N+1], rev_src[b][N*k] (source encoded as n'*k + j, ascending per
// target).  Shared memory: N counters.
__global__ void __launch_bounds__(1024)
edge_rev_csr_kernel(const int64_t* __restrict__ idx, int N, int k, int* __restrict__ rev_ptr,
                    int* __restrict__ rev_src, int* __restrict__ slot_tmp) {
    extern __shared__ int cnt[];                            // [N] counts, then running offsets
    __shared__ int warp_tot[32];
    const int b = blockIdx.x;
    const int E = N * k;
    const int64_t* idb = idx + (size_t)b * E;
    int* ptr = rev_ptr + (size_t)b * (N + 1);
    int* src = rev_src + (size_t)b * E;
    int* slot = slot_tmp + (size_t)b * E;
    for (int i = threadIdx.x; i < N; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        unsigned m = (unsigned)__ldg(idb + e);
        m = m < (unsigned)N ? m : (unsigned)(N - 1);
        slot[e] = atomicAdd(&cnt[m], 1);                    // arrival order, fixed up by the sort below
    }
    __syncthreads();
    // exclusive scan of cnt[0..N) -> ptr; block-wide, chunked by blockDim
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < N; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < N ? cnt[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nwarp ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;                             // inclusive totals
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + inc - v;
        if (i < N) { ptr[i] = before; cnt[i] = before; }
        carry += warp_tot[nwarp - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[N] = E;
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        unsigned m = (unsigned)__ldg(idb + e);
        m = m < (unsigned)N ? m : (unsigned)(N - 1);
        src[cnt[m] + slot[e]] = e;
    }
    __syncthreads();
    // make each list ascending (deterministic summation order); lists are short (mean k)
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int lo = cnt[i], hi = (i + 1 < N) ? cnt[i + 1] : E;
        for (int p = lo + 1; p < hi; ++p) {
            const int key = src[p];
            int q = p - 1;
            while (q >= lo && src[q] > key) { src[q + 1] = src[q]; --q; }
            src[q + 1] = key;
        }
    }
}

// grid: (3C, B); block 512.  One (cloud, channel*3+component) plane per CTA.
//   gx[n] = sum_j gctr[n,j] - sum_j gdiff[n,j] + sum_{(n',j): idx[n',j]=n} gdiff[n',j]   (+ cross terms)
// STAGED: the gdiff plane is copied to shared memory first.
template <bool STAGED, bool CROSS>
__global__ void __launch_bounds__(512)
edge_feat_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x, const int64_t* __restrict__ idx,
                     const int* __restrict__ rev_ptr, const int* __restrict__ rev_src, int C, int N, int k,
                     int vec, float* __restrict__ gx) {
    extern __shared__ __align__(16) float gs[];            // STAGED: [N*k] (+ 2 more planes for CROSS)
    const int b = blockIdx.y;
    const int ch = blockIdx.x;                              // c*3 + a
    const int c = ch / 3, a = ch % 3;
    const int E = N * k;
    const int planes = CROSS ? 3 : 2;
    const float* gb = gout + (size_t)b * planes * C * 3 * E;
    const float* g_diff = gb + ((size_t)c * 3 + a) * E;
    const float* g_ctr = gb + ((size_t)(C + c) * 3 + a) * E;
    const int* ptr = rev_ptr + (size_t)b * (N + 1);
    const int* src = rev_src + (size_t)b * E;

    if (STAGED) {
        if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(g_diff);
            float4* d4 = reinterpret_cast<float4*>(gs);
            for (int e = threadIdx.x; e < E / 4; e += blockDim.x) d4[e] = __ldcs(s4 + e);
        } else {
            for (int e = threadIdx.x; e < E; e += blockDim.x) gs[e] = __ldcs(g_diff + e);
        }
        __syncthreads();
    }
    const float* gd = STAGED ? gs : g_diff;

    // cross(f, x) with f = x_m (neighbour), x = x_n (centre):
    //   out_a = f_{a+1} x_{a+2} - f_{a+2} x_{a+1}
    // d/d x_n[a]: from out_{a+1} = f_{a+2} x_a - f_a x_{a+2}  -> +f_{a+2} g_{a+1}
    //             from out_{a+2} = f_a x_{a+1} - f_{a+1} x_a  -> -f_{a+1} g_{a+2}
    // d/d f[a]  : from out_{a+1}: -x_{a+2} g_{a+1};  from out_{a+2}: +x_{a+1} g_{a+2}
    const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const float* g_c1 = gb + ((size_t)(2 * C + c) * 3 + a1) * E;
    const float* g_c2 = gb + ((size_t)(2 * C + c) * 3 + a2) * E;
    const float* xc = x + ((size_t)b * C + c) * 3 * N;
    const int64_t* idb = idx + (size_t)b * E;

    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float acc = 0.f;
        const size_t row = (size_t)n * k;
        if (vec) {
            for (int j = 0; j < k; j += 4) {
                const float4 d = *reinterpret_cast<const float4*>(gd + row + j);
                const float4 t = __ldcs(reinterpret_cast<const float4*>(g_ctr + row + j));
                acc += (t.x - d.x) + (t.y - d.y) + (t.z - d.z) + (t.w - d.w);
            }
        } else {
            for (int j = 0; j < k; ++j) acc += __ldcs(g_ctr + row + j) - gd[row + j];
        }
        const int lo = __ldg(ptr + n), hi = __ldg(ptr + n + 1);
        for (int p = lo; p < hi; ++p) acc += gd[__ldg(src + p)];
        if (CROSS) {
            // centre role of n
            for (int j = 0; j < k; ++j) {
                unsigned m = (unsigned)__ldg(idb + row + j);
                m = m < (unsigned)N ? m : (unsigned)(N - 1);
                acc += __ldg(xc + a2 * N + m) * __ldg(g_c1 + row + j) - __ldg(xc + a1 * N + m) * __ldg(g_c2 + row + j);
            }
            // neighbour role of n
            for (int p = lo; p < hi; ++p) {
                const int e = __ldg(src + p);
                const int nn = e / k;
                acc += __ldg(xc + a1 * N + nn) * __ldg(g_c2 + e) - __ldg(xc + a2 * N + nn) * __ldg(g_c1 + e);
            }
        }
        gx[((size_t)b * C * 3 + ch) * N + n] = acc;
    }
}

}  // namespace hpcs

extern "C" {

int hpcs_edge_feat_fwd_f32(const float* x, const int64_t* idx, int B, int C, int N, int k, int cross,
                           float* out, void* stream) {
    using namespace hpcs;
    if (!x || !idx || !out) return fail(HPCS_ERR_ARG, "edge_feat_fwd: null pointer");
    if (B <= 0 || C <= 0 || N <= 0 || k <= 0) return fail(HPCS_ERR_ARG, "edge_feat_fwd: bad shape");
    if ((size_t)3 * N * sizeof(float) > 200 * 1024) return fail(HPCS_ERR_ARG, "edge_feat_fwd: N=%d too large (max 17066)", N);
    if (B > 65535 || C > 65535) return fail(HPCS_ERR_ARG, "edge_feat_fwd: B or C > 65535");
    cudaStream_t st = as_stream(stream);
    const size_t E = (size_t)N * k;
    const bool vec = (E % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(idx) & 15) == 0);
    // enough CTAs for ~8 per SM, but at least 2048 vector groups each
    const size_t groups = vec ? E / 4 : E;
    int splits = (8 * sm_count() + B * C - 1) / (B * C);
    const int max_splits = (int)((groups + 2047) / 2048);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    dim3 grid(splits, C, B), block(256);
    const size_t smem = (size_t)3 * N * sizeof(float);
#define HPCS_LAUNCH_FWD(V, X)                                                                         \
    {                                                                                                  \
        auto kern = edge_feat_fwd_kernel<V, X>;                                                        \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        kern<<<grid, block, smem, st>>>(x, idx, C, N, k, out);                                         \
    }
    if (vec) { if (cross) HPCS_LAUNCH_FWD(4, true) else HPCS_LAUNCH_FWD(4, false) }
    else     { if (cross) HPCS_LAUNCH_FWD(1, true) else HPCS_LAUNCH_FWD(1, false) }
#undef HPCS_LAUNCH_FWD
    return check_launch("edge_feat_fwd_kernel");
}

size_t hpcs_edge_feat_bwd_workspace_bytes(int B, int N, int k) {
    const size_t E = (size_t)N * k;
    return hpcs::align_up((size_t)B * (N + 1) * sizeof(int), 256) + 2 * hpcs::align_up((size_t)B * E * sizeof(int), 256);
}

int hpcs_edge_feat_bwd_f32(const float* gout, const float* x, const int64_t* idx, int B, int C, int N, int k,
                           int cross, float* gx, void* ws, size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!gout || !idx || !gx || !ws || (cross && !x)) return fail(HPCS_ERR_ARG, "edge_feat_bwd: null pointer");
    if (B <= 0 || C <= 0 || N <= 0 || k <= 0) return fail(HPCS_ERR_ARG, "edge_feat_bwd: bad shape");
    if (ws_bytes < hpcs_edge_feat_bwd_workspace_bytes(B, N, k)) return fail(HPCS_ERR_WORKSPACE, "edge_feat_bwd: workspace too small");
    if ((size_t)N * sizeof(int) > 200 * 1024) return fail(HPCS_ERR_ARG, "edge_feat_bwd: N=%d too large (max 51200)", N);
    if (B > 65535) return fail(HPCS_ERR_ARG, "edge_feat_bwd: B > 65535");
    cudaStream_t st = as_stream(stream);
    const size_t E = (size_t)N * k;
    char* w = static_cast<char*>(ws);
    int* rev_ptr = reinterpret_cast<int*>(w);
    w += align_up((size_t)B * (N + 1) * sizeof(int), 256);
    int* rev_src = reinterpret_cast<int*>(w);
    w += align_up((size_t)B * E * sizeof(int), 256);
    int* slot_tmp = reinterpret_cast<int*>(w);

    {
        const size_t smem = (size_t)N * sizeof(int);
        cudaFuncSetAttribute(edge_rev_csr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge_rev_csr_kernel<<<B, 1024, smem, st>>>(idx, N, k, rev_ptr, rev_src, slot_tmp);
        int rc = check_launch("edge_rev_csr_kernel");
        if (rc) return rc;
    }
    dim3 grid(3 * C, B), block(512);
    const size_t smem = E * sizeof(float);
    const bool staged = smem <= 200 * 1024;
    const int vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
#define HPCS_LAUNCH_BWD(S, X)                                                                          \
    {                                                                                                  \
        auto kern = edge_feat_bwd_kernel<S, X>;                                                        \
        if (S) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        kern<<<grid, block, S ? smem : 0, st>>>(gout, x, idx, rev_ptr, rev_src, C, N, k, vec, gx);          \
    }
    if (staged) { if (cross) HPCS_LAUNCH_BWD(true, true) else HPCS_LAUNCH_BWD(true, false) }
    else        { if (cross) HPCS_LAUNCH_BWD(false, true) else HPCS_LAUNCH_BWD(false, false) }
#undef HPCS_LAUNCH_BWD
    return check_launch("edge_feat_bwd_kernel");
}

}  // extern "C"
