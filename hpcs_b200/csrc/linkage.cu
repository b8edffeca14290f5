// Decoding leaf embeddings into a binary dendrogram (scipy linkage format) on the GPU.
//
// Replaces, per cloud, the D2H copy + scipy.cluster.hierarchy.linkage(leaves, method, 'cosine')
// of BaseSimilarityHypHC._decode_linkage (hpcs/models/base_hyp_hc.py:81-86) and the Python loop
// over clouds at :135-137: all B clouds are decoded by one launch sequence, one CTA per cloud.
//   1. pdist_cosine_kernel  -- fp64 cosine distance matrix, bit-identical to scipy's pdist
//      (two running sums over even/odd elements, separate multiply/add, see oracle);
//   2. linkage_kernel<0>    -- 'single': Prim's MST from node 0 over matrix rows (what scipy's
//      mst_single_linkage does), block-wide (value, index) arg-min per step;
//      linkage_kernel<1>    -- 'complete': nearest-neighbour chain with the max update
//      (scipy's nn_chain), same arg-min primitive, distance matrix updated in place;
//   3. in the same kernel: stable bitonic sort of the N-1 merges by height, then the union-find
//      relabel pass (smaller root id first, new id N+i, subtree size) that scipy's `label` does.
// Because all leaves share one radius, hyperbolic-LCA similarity is a decreasing function of the
// angle and cosine distance an increasing one, so single linkage over cosine distance yields the
// same merge order as single linkage over hyperbolic similarity (SURVEY.md Finding 2).
#include <float.h>

#include "common.cuh"

namespace hpcs {

constexpr int kTile = 32;

// grid: (T*(T+1)/2 upper tiles, B); block 256.  dm[b][i][j] full symmetric, zero diagonal.
__global__ void __launch_bounds__(256)
pdist_cosine_kernel(const float* __restrict__ leaves, int N, int D, double* __restrict__ dm) {
    extern __shared__ double sm[];
    double* ri = sm;                          // [32][D]   rows of the i-tile
    double* rj = ri + kTile * D;              // [D][32]   rows of the j-tile, transposed
    double* ni = rj + kTile * D;              // [32] norms
    double* nj = ni + kTile;                  // [32]
    double* tile = nj + kTile;                // [32][33] results
    const int b = blockIdx.y;
    // linear upper-triangular tile index -> (bi, bj), bi <= bj
    const int T = (N + kTile - 1) / kTile;
    int bi = 0, rem = blockIdx.x;
    while (rem >= T - bi) { rem -= T - bi; ++bi; }
    const int bj = bi + rem;
    const float* lb = leaves + (size_t)b * N * D;
    for (int e = threadIdx.x; e < kTile * D; e += blockDim.x) {
        const int r = e / D, q = e % D;
        const int gi = bi * kTile + r, gj = bj * kTile + r;
        ri[r * D + q] = gi < N ? (double)lb[(size_t)gi * D + q] : 0.0;
        rj[q * kTile + r] = gj < N ? (double)lb[(size_t)gj * D + q] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < 2 * kTile) {
        const bool second = threadIdx.x >= kTile;
        const int r = threadIdx.x % kTile;
        double even = 0.0, odd = 0.0;
        for (int q = 0; q + 1 < D; q += 2) {
            const double v0 = second ? rj[q * kTile + r] : ri[r * D + q];
            const double v1 = second ? rj[(q + 1) * kTile + r] : ri[r * D + q + 1];
            even = __dadd_rn(even, __dmul_rn(v0, v0));
            odd = __dadd_rn(odd, __dmul_rn(v1, v1));
        }
        double s = __dadd_rn(even, odd);
        if (D & 1) {
            const double v = second ? rj[(D - 1) * kTile + r] : ri[r * D + D - 1];
            s = __dadd_rn(s, __dmul_rn(v, v));
        }
        (second ? nj : ni)[r] = __dsqrt_rn(s);
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ty * 4 + r;
        double even = 0.0, odd = 0.0;
        for (int q = 0; q + 1 < D; q += 2) {
            even = __dadd_rn(even, __dmul_rn(ri[i * D + q], rj[q * kTile + tx]));
            odd = __dadd_rn(odd, __dmul_rn(ri[i * D + q + 1], rj[(q + 1) * kTile + tx]));
        }
        double s = __dadd_rn(even, odd);
        if (D & 1) s = __dadd_rn(s, __dmul_rn(ri[i * D + D - 1], rj[(D - 1) * kTile + tx]));
        double c = __ddiv_rn(s, __dmul_rn(ni[i], nj[tx]));
        if (fabs(c) > 1.0) c = copysign(1.0, c);
        const int gi = bi * kTile + i, gj = bj * kTile + tx;
        tile[i * 33 + tx] = (gi == gj) ? 0.0 : __dsub_rn(1.0, c);
    }
    __syncthreads();
    double* db = dm + (size_t)b * N * N;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = ty * 4 + r;
        const int gi = bi * kTile + i, gj = bj * kTile + tx;
        if (gi < N && gj < N) db[(size_t)gi * N + gj] = tile[i * 33 + tx];
        if (bi != bj) {                       // mirrored tile, read transposed
            const int gi2 = bj * kTile + i, gj2 = bi * kTile + tx;
            if (gi2 < N && gj2 < N) db[(size_t)gi2 * N + gj2] = tile[tx * 33 + i];
        }
    }
}

struct ArgMin {
    double v;
    int i;
};
__device__ __forceinline__ bool am_less(double v, int i, double ov, int oi) { return v < ov || (v == ov && i < oi); }

// Block-wide lexicographic (value, index) minimum; every thread gets the result.
__device__ __forceinline__ ArgMin block_argmin(double v, int i, ArgMin* scratch /*[32]*/, ArgMin* result) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFull, v, o);
        const int oi = __shfl_xor_sync(kFull, i, o);
        if (am_less(ov, oi, v, i)) { v = ov; i = oi; }
    }
    if (lane == 0) { scratch[warp].v = v; scratch[warp].i = i; }
    __syncthreads();
    if (warp == 0) {
        v = lane < nwarp ? scratch[lane].v : DBL_MAX;
        i = lane < nwarp ? scratch[lane].i : 0x7fffffff;
        if (!(lane < nwarp)) v = INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, v, o);
            const int oi = __shfl_xor_sync(kFull, i, o);
            if (am_less(ov, oi, v, i)) { v = ov; i = oi; }
        }
        if (lane == 0) { result->v = v; result->i = i; }
    }
    __syncthreads();
    return *result;
}

// METHOD 0 single, 1 complete.  One CTA per cloud.
template <int METHOD>
__global__ void __launch_bounds__(1024)
linkage_kernel(double* dm_all, int N, int NP2, int* __restrict__ recx_all, int* __restrict__ recy_all,
               double* __restrict__ rech_all, double* __restrict__ Z_all) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ ArgMin scratch[32];
    __shared__ ArgMin result;
    __shared__ int ctl[4];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int M = N - 1;
    double* dm = dm_all + (size_t)b * N * N;              // mutated by METHOD 1: plain loads only
    int* recx = recx_all + (size_t)b * N;
    int* recy = recy_all + (size_t)b * N;
    double* rech = rech_all + (size_t)b * N;
    double* Z = Z_all + (size_t)b * M * 4;

    // shared-memory regions (see header comment for the overlay plan)
    double* A = reinterpret_cast<double*>(raw);                      // [NP2]  Dmin / sort keys
    int* ordv = reinterpret_cast<int*>(raw + (size_t)8 * NP2);       // [NP2]  sort payload
    unsigned char* regC = raw + (size_t)12 * NP2;                    // 8N bytes: flags / sizes+chain / sorted (x,y)
    int* csize = reinterpret_cast<int*>(regC + (size_t)8 * N);       // [2N]
    int* parent = reinterpret_cast<int*>(raw);                       // [2N] overlays A+ordv after the sort

    if (METHOD == 0) {
        unsigned char* merged = regC;
        for (int i = tid; i < N; i += nthr) { A[i] = INFINITY; merged[i] = 0; }
        __syncthreads();
        int x = 0;
        for (int k = 0; k < M; ++k) {
            if (tid == 0) merged[x] = 1;
            __syncthreads();
            const double* row = dm + (size_t)x * N;
            double bv = INFINITY;
            int bi = 0x7fffffff;
            for (int i = tid; i < N; i += nthr) {
                if (merged[i]) continue;
                const double d = row[i];
                double cur = A[i];
                if (cur > d) { cur = d; A[i] = d; }
                if (cur < bv) { bv = cur; bi = i; }
            }
            const ArgMin r = block_argmin(bv, bi, scratch, &result);
            if (tid == 0) { recx[k] = x; recy[k] = r.i; rech[k] = r.v; }
            x = r.i;
        }
    } else {
        int* size = reinterpret_cast<int*>(regC);                     // [N]
        int* chain = size + N;                                        // [N]
        for (int i = tid; i < N; i += nthr) size[i] = 1;
        if (tid == 0) { ctl[0] = 0; /* chain length */ ctl[1] = 0; /* first active */ }
        __syncthreads();
        for (int k = 0; k < M; ++k) {
            if (tid == 0 && ctl[0] == 0) {
                int f = ctl[1];
                while (size[f] == 0) ++f;
                ctl[1] = f;
                chain[0] = f;
                ctl[0] = 1;
            }
            __syncthreads();
            int x, y;
            double cur;
            while (true) {
                const int len = ctl[0];
                x = chain[len - 1];
                const int prev = len > 1 ? chain[len - 2] : -1;
                const double* row = dm + (size_t)x * N;
                double bv = INFINITY;
                int bi = 0x7fffffff;
                for (int i = tid; i < N; i += nthr) {
                    if (size[i] == 0 || i == x) continue;
                    const double d = row[i];
                    if (d < bv) { bv = d; bi = i; }
                }
                const ArgMin r = block_argmin(bv, bi, scratch, &result);
                y = r.i; cur = r.v;
                if (prev >= 0) {
                    const double dprev = row[prev];
                    if (!(r.v < dprev)) { y = prev; cur = dprev; }
                }
                const bool done = prev >= 0 && y == prev;
                __syncthreads();                    // everyone has read ctl/chain before they change
                if (done) break;
                if (tid == 0) { chain[len] = y; ctl[0] = len + 1; }
                __syncthreads();
            }
            if (x > y) { const int t = x; x = y; y = t; }
            const int nx = size[x], ny = size[y];
            __syncthreads();
            if (tid == 0) {
                ctl[0] -= 2;
                recx[k] = x; recy[k] = y; rech[k] = cur;
                size[x] = 0; size[y] = nx + ny;
            }
            // Lance-Williams 'complete': d(i, x u y) = max(d(i,x), d(i,y)); kept symmetric
            double* rowx = dm + (size_t)x * N;
            double* rowy = dm + (size_t)y * N;
            for (int i = tid; i < N; i += nthr) {
                if (i == y || i == x || size[i] == 0) continue;
                const double v = fmax(rowx[i], rowy[i]);
                rowy[i] = v;
                dm[(size_t)i * N + y] = v;
            }
            __syncthreads();
        }
    }
    __syncthreads();

    // ---- stable sort of the merges by height ----------------------------------------------------
    for (int i = tid; i < NP2; i += nthr) {
        A[i] = i < M ? rech[i] : INFINITY;
        ordv[i] = i < M ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int kk = 2; kk <= NP2; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < NP2; i += nthr) {
                const int l = i ^ j;
                if (l > i) {
                    const double vi = A[i], vl = A[l];
                    const int oi = ordv[i], ol = ordv[l];
                    const bool up = (i & kk) == 0;
                    const bool wrong = up ? am_less(vl, ol, vi, oi) : am_less(vi, oi, vl, ol);
                    if (wrong) { A[i] = vl; A[l] = vi; ordv[i] = ol; ordv[l] = oi; }
                }
            }
            __syncthreads();
        }
    }
    int2* exy = reinterpret_cast<int2*>(regC);
    for (int i = tid; i < M; i += nthr) {
        const int o = ordv[i];
        exy[i] = make_int2(recx[o], recy[o]);
        Z[(size_t)i * 4 + 2] = A[i];
    }
    __syncthreads();
    // ---- union-find relabel (scipy `label`) -----------------------------------------------------
    for (int v = tid; v < 2 * N; v += nthr) { parent[v] = v; csize[v] = v < N ? 1 : 0; }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < M; ++i) {
            const int2 e = exy[i];
            int rx = e.x, ry = e.y;
            while (parent[rx] != rx) { const int g = parent[parent[rx]]; parent[rx] = g; rx = g; }
            while (parent[ry] != ry) { const int g = parent[parent[ry]]; parent[ry] = g; ry = g; }
            const int id = N + i;
            const int sz = csize[rx] + csize[ry];
            parent[rx] = id; parent[ry] = id; csize[id] = sz;
            Z[(size_t)i * 4 + 0] = (double)(rx < ry ? rx : ry);
            Z[(size_t)i * 4 + 1] = (double)(rx < ry ? ry : rx);
            Z[(size_t)i * 4 + 3] = (double)sz;
        }
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace hpcs

extern "C" {

size_t hpcs_linkage_workspace_bytes(int B, int N, int D, int method) {
    (void)D; (void)method;
    if (B <= 0 || N <= 1) return 0;
    using hpcs::align_up;
    return align_up((size_t)B * N * N * sizeof(double), 256) + 2 * align_up((size_t)B * N * sizeof(int), 256) +
           align_up((size_t)B * N * sizeof(double), 256);
}

int hpcs_linkage_f64(const float* leaves, int B, int N, int D, int method, double* Z, void* ws,
                     size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!leaves || !Z || !ws) return fail(HPCS_ERR_ARG, "linkage: null pointer");
    if (B <= 0 || N < 2 || D <= 0 || (method != 0 && method != 1)) return fail(HPCS_ERR_ARG, "linkage: bad arguments B=%d N=%d D=%d method=%d", B, N, D, method);
    if (B > 65535) return fail(HPCS_ERR_ARG, "linkage: B > 65535");
    if (ws_bytes < hpcs_linkage_workspace_bytes(B, N, D, method)) return fail(HPCS_ERR_WORKSPACE, "linkage: workspace too small");
    const int NP2 = next_pow2(N - 1 > 1 ? N - 1 : 2);
    const size_t smem_link = (size_t)12 * NP2 + (size_t)16 * N;
    if (smem_link > 227 * 1024) return fail(HPCS_ERR_ARG, "linkage: N=%d too large (max 8192)", N);
    const size_t smem_pd = ((size_t)2 * kTile * D + 2 * kTile + kTile * 33) * sizeof(double);
    if (smem_pd > 200 * 1024) return fail(HPCS_ERR_ARG, "linkage: D=%d too large", D);
    cudaStream_t st = as_stream(stream);
    char* w = static_cast<char*>(ws);
    double* dm = reinterpret_cast<double*>(w);  w += align_up((size_t)B * N * N * sizeof(double), 256);
    int* recx = reinterpret_cast<int*>(w);      w += align_up((size_t)B * N * sizeof(int), 256);
    int* recy = reinterpret_cast<int*>(w);      w += align_up((size_t)B * N * sizeof(int), 256);
    double* rech = reinterpret_cast<double*>(w);

    const int T = (N + kTile - 1) / kTile;
    cudaFuncSetAttribute(pdist_cosine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pd);
    pdist_cosine_kernel<<<dim3(T * (T + 1) / 2, B), 256, smem_pd, st>>>(leaves, N, D, dm);
    int rc = check_launch("pdist_cosine_kernel");
    if (rc) return rc;
    const int threads = N >= 1024 ? 1024 : (N + 31) / 32 * 32;
    if (method == 0) {
        cudaFuncSetAttribute(linkage_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_link);
        linkage_kernel<0><<<B, threads, smem_link, st>>>(dm, N, NP2, recx, recy, rech, Z);
    } else {
        cudaFuncSetAttribute(linkage_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_link);
        linkage_kernel<1><<<B, threads, smem_link, st>>>(dm, N, NP2, recx, recy, rech, Z);
    }
    return check_launch("linkage_kernel");
}

}  // extern "C"
