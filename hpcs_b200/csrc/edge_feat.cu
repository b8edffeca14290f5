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

// fast backward path (edge_bwd.cu)
bool edge_bwd_fast_applicable(const float* gout, int N, int k);
size_t edge_bwd_fast_workspace_bytes(int B, int N, int k);
int edge_bwd_fast_run(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, void* ws, cudaStream_t st);
int edge_bwd_fast_build(const int64_t* idx, int B, int N, int k, void* ws, cudaStream_t st);
int edge_bwd_fast_gather(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, const void* ws, cudaStream_t st);

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
// backward: reverse graph (target -> sources) in degree-sorted sliced-ELL form
// ------------------------------------------------------------------------------------------------
// kNN graphs in feature space are hub-heavy (in-degree up to ~N/4 at mean k), so the reverse lists are
// stored so that a warp of 32 targets of SIMILAR in-degree reads them coalesced:
//   perm[r]          target with degree rank r (descending)
//   group g          = ranks [32g, 32g+32); gslots[g] = largest in-degree in the group; goff[g] = start
//   ell[goff[g] + s*32 + l]  = s-th source (encoded n'*k + j, ascending) of target perm[32g+l], or the
//                              sentinel E (points at a zero) past the end of that target's list.
// Size: sum_g 32*gslots[g] <= E + 32*maxdeg.  The workspace holds E + 32*N entries, enough for any idx
// whose rows are duplicate-free (every kNN result); otherwise the build falls back to plain CSR
// (mode = 1) and the gather takes its thread-per-row path.
struct RevGraph {
    int* hdr;      // [8]: mode, total
    int* perm;     // [N]
    int* gslots;   // [G]
    int* goff;     // [G+1]
    int* ptr;      // [N+1]  CSR offsets (always written)
    int* ell;      // [E + 32N]  (CSR src[E] in fallback mode)
    size_t ints_per_cloud;
};

__host__ __device__ inline RevGraph rev_graph(int* base, int b, int N, int k) {
    const size_t E = (size_t)N * k, G = (N + 31) / 32;
    RevGraph r;
    size_t per = 8 + (size_t)N + G + (G + 1) + (N + 1) + (E + 32 * (size_t)N);
    per = (per + 63) / 64 * 64;
    int* p = base + (size_t)b * per;
    r.hdr = p;            p += 8;
    r.perm = p;           p += N;
    r.gslots = p;         p += G;
    r.goff = p;           p += G + 1;
    r.ptr = p;            p += N + 1;
    r.ell = p;
    r.ints_per_cloud = per;
    return r;
}

// exclusive scan of in[0..L) into out[0..L] (out[L] = total); whole block takes part; L arbitrary
__device__ inline void block_excl_scan(const int* in, int* out, int L, int* warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < L; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < L ? in[i] : 0;
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
            warp_tot[lane] = w;
        }
        __syncthreads();
        if (i < L) out[i] = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + inc - v;
        carry += warp_tot[nwarp - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[L] = carry;
    __syncthreads();
}

// One CTA (NW warps) per cloud.  A stable counting sort of the E edges by target without atomics: warp w
// owns the contiguous edge range [w*chunk, (w+1)*chunk) and a private histogram row hist[w][.]; inside a
// 32-edge step __match_any_sync ranks equal targets by lane.  The resulting lists are ascending in e, so
// the gather sums in a fixed order (deterministic gradients).
__global__ void __launch_bounds__(1024)
edge_rev_build_kernel(const int64_t* __restrict__ idx, int N, int k, int* __restrict__ ws) {
    extern __shared__ int sm_i[];
    __shared__ int warp_tot[32];
    __shared__ int s_mode;
    const int NW = blockDim.x >> 5;
    const int G = (N + 31) / 32;
    int* hist = sm_i;                         // [NW][N]
    int* deg = hist + (size_t)NW * N;         // [N]
    int* off = deg + N;                       // [N+1]
    int* bins = off + N + 1;                  // [N+2]  rows per degree key, then start offsets
    int* rank = bins + N + 2;                 // [N]
    int* gsl = rank + N;                      // [G] group slots, then [G+1] offsets in gof
    int* gof = gsl + G;                       // [G+1]
    const int b = blockIdx.x;
    const int E = N * k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t* idb = idx + (size_t)b * E;
    const RevGraph R = rev_graph(ws, b, N, k);
    const int chunk = ((E + NW - 1) / NW + 31) / 32 * 32;

    for (int i = threadIdx.x; i < NW * N; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < N + 2; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    // target of edge e, or a per-lane dummy key past N for idle lanes (so they never match a real target)
    auto target_of = [&](int e) -> unsigned {
        if (e >= E) return (unsigned)(N + lane);
        const unsigned t = (unsigned)__ldg(idb + e);
        return t < (unsigned)N ? t : (unsigned)(N - 1);
    };
    // per-warp histograms (loads issued 4 steps at a time so their L2 latency overlaps)
    for (int i0 = 0; i0 < chunk; i0 += 128) {
        unsigned tq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tq[u] = target_of(i0 + 32 * u < chunk ? warp * chunk + i0 + 32 * u + lane : E);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned t = tq[u];
            const unsigned mask = __match_any_sync(kFull, t);
            if (t < (unsigned)N && lane == __ffs(mask) - 1) hist[warp * N + t] += __popc(mask);
            __syncwarp();
        }
    }
    __syncthreads();
    // exclusive prefix over warps per target -> in-degree
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        int run = 0;
        for (int w = 0; w < NW; ++w) { const int c = hist[w * N + t]; hist[w * N + t] = run; run += c; }
        deg[t] = run;
        atomicAdd(&bins[N - min(run, N)], 1);                   // key 0 = largest degree
    }
    __syncthreads();
    block_excl_scan(deg, off, N, warp_tot);
    for (int t = threadIdx.x; t <= N; t += blockDim.x) R.ptr[t] = off[t];
    block_excl_scan(bins, bins, N + 1, warp_tot);               // bins[key] = first rank of that key
    for (int t = threadIdx.x; t < N; t += blockDim.x) {
        const int r = atomicAdd(&bins[N - min(deg[t], N)], 1);  // order inside one degree is irrelevant
        rank[t] = r;
        R.perm[r] = t;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += blockDim.x) gsl[g] = deg[R.perm[g * 32]];   // largest of the group
    __syncthreads();
    block_excl_scan(gsl, gof, G, warp_tot);
    if (threadIdx.x == 0) {
        const long long total = 32ll * gof[G];
        s_mode = total <= (long long)E + 32ll * N ? 0 : 1;
        R.hdr[0] = s_mode;
        R.hdr[1] = (int)(s_mode == 0 ? total : E);
    }
    __syncthreads();
    const int mode = s_mode;
    if (mode == 0) {
        for (int g = threadIdx.x; g < G; g += blockDim.x) { R.gslots[g] = gsl[g]; R.goff[g] = 32 * gof[g]; }
        if (threadIdx.x == 0) R.goff[G] = 32 * gof[G];
        const int total = 32 * gof[G];
        for (int i = threadIdx.x; i < total; i += blockDim.x) R.ell[i] = E;      // sentinel
        __syncthreads();
    }
    // placement, same edge order as the counting pass
    for (int i0 = 0; i0 < chunk; i0 += 128) {
        unsigned tq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) tq[u] = target_of(i0 + 32 * u < chunk ? warp * chunk + i0 + 32 * u + lane : E);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned t = tq[u];
            const int e = warp * chunk + i0 + 32 * u + lane;
            const bool valid = t < (unsigned)N;
            const unsigned mask = __match_any_sync(kFull, t);
            if (valid) {
                const int pos = hist[warp * N + t] + __popc(mask & ((1u << lane) - 1u));
                if (mode == 0) {
                    const int r = rank[t];
                    R.ell[32 * gof[r >> 5] + pos * 32 + (r & 31)] = e;
                } else {
                    R.ell[off[t] + pos] = e;
                }
            }
            __syncwarp();
            if (valid && lane == __ffs(mask) - 1) hist[warp * N + t] += __popc(mask);
            __syncwarp();
        }
    }
}

// grid: (3C, B); block 1024.  One (cloud, channel*3+component) plane per CTA.
//   gx[n] = sum_j gctr[n,j] - sum_j gdiff[n,j] + sum_{(n',j): idx[n',j]=n} gdiff[n',j]   (+ cross terms)
// STAGED: the gdiff plane is copied to shared memory (+ a zero at [E] for the ELL sentinel) and, when
// rows are float4-aligned, the per-float4 partial row sums of (gctr - gdiff) are formed during that same
// coalesced pass, so both planes are read from HBM exactly once with 128-bit loads.
// Clouds too large for the per-cloud reverse graph to be built in shared memory (N above ~11000: BASELINE configs[3] goes to
// 16384): plain scatter.  One thread per point walks its k edges in one (cloud, channel, component) plane pair; the centre
// part of the gradient is a private row sum, the neighbour part goes out as fp32 reductions on gx (L2 resident: 3C planes of
// N floats per cloud), which the caller zeroes.  Summation order is whatever the atomics give: results agree to rounding.
__global__ void __launch_bounds__(256)
edge_feat_bwd_scatter_kernel(const float* __restrict__ gout, const int64_t* __restrict__ idx, int C, int N, int k,
                             float* __restrict__ gx) {
    const int b = blockIdx.z, ch = blockIdx.y;              // ch = c*3 + a
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const size_t E = (size_t)N * k;
    const float* gb = gout + (size_t)b * 2 * C * 3 * E;
    const float* g_diff = gb + (size_t)ch * E + (size_t)n * k;
    const float* g_ctr = gb + ((size_t)C * 3 + ch) * E + (size_t)n * k;
    const int64_t* row = idx + ((size_t)b * N + n) * k;
    float* out = gx + ((size_t)b * 3 * C + ch) * N;
    float self = 0.f;
    for (int j = 0; j < k; ++j) {
        const float gd = __ldg(g_diff + j);
        self += __ldg(g_ctr + j) - gd;
        const long long m = row[j];
        if (m >= 0 && m < N) atomicAdd(out + m, gd);
    }
    atomicAdd(out + n, self);
}

template <bool STAGED, bool CROSS>
__global__ void __launch_bounds__(1024)
edge_feat_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x, const int64_t* __restrict__ idx,
                     int* __restrict__ ws, int C, int N, int k, int vec, float* __restrict__ gx) {
    extern __shared__ __align__(16) float gs[];            // STAGED: [E + 4] plane, then [E/4] row partials
    const int b = blockIdx.y;
    const int ch = blockIdx.x;                              // c*3 + a
    const int c = ch / 3, a = ch % 3;
    const int E = N * k;
    const int planes = CROSS ? 3 : 2;
    const float* gb = gout + (size_t)b * planes * C * 3 * E;
    const float* g_diff = gb + ((size_t)c * 3 + a) * E;
    const float* g_ctr = gb + ((size_t)(C + c) * 3 + a) * E;
    const RevGraph R = rev_graph(ws, b, N, k);
    float* rpart = gs + (E + 4);
    const bool fused_rows = STAGED && vec;

    if (STAGED) {
        if (vec) {
            const float4* d4 = reinterpret_cast<const float4*>(g_diff);
            const float4* c4 = reinterpret_cast<const float4*>(g_ctr);
            float4* s4 = reinterpret_cast<float4*>(gs);
            for (int f = threadIdx.x; f < E / 4; f += blockDim.x) {
                const float4 d = __ldcs(d4 + f), t = __ldcs(c4 + f);
                s4[f] = d;
                rpart[f] = (t.x - d.x) + (t.y - d.y) + (t.z - d.z) + (t.w - d.w);
            }
        } else {
            for (int e = threadIdx.x; e < E; e += blockDim.x) gs[e] = __ldcs(g_diff + e);
        }
        if (threadIdx.x == 0) gs[E] = 0.f;
        __syncthreads();
    }
    const float* gd = STAGED ? gs : g_diff;

    // cross(f, x) with f = x_m (neighbour), x = x_n (centre):  out_a = f_{a+1} x_{a+2} - f_{a+2} x_{a+1}
    // d/d x_n[a]: +f_{a+2} g_{a+1} - f_{a+1} g_{a+2};   d/d f[a]: -x_{a+2} g_{a+1} + x_{a+1} g_{a+2}
    const int a1 = (a + 1) % 3, a2 = (a + 2) % 3;
    const float* g_c1 = gb + ((size_t)(2 * C + c) * 3 + a1) * E;
    const float* g_c2 = gb + ((size_t)(2 * C + c) * 3 + a2) * E;
    const float* xc = x + ((size_t)b * C + c) * 3 * N;
    const int64_t* idb = idx + (size_t)b * E;
    float* out = gx + ((size_t)b * C * 3 + ch) * N;

    auto own_row = [&](int n) -> float {
        float own = 0.f;
        const size_t row = (size_t)n * k;
        if (fused_rows) {
            const int q = k >> 2;
            for (int i = 0; i < q; ++i) own += rpart[n * q + i];
        } else {
            for (int j = 0; j < k; ++j) own += __ldg(g_ctr + row + j) - gd[row + j];
        }
        if (CROSS) {
            for (int j = 0; j < k; ++j) {
                unsigned m = (unsigned)__ldg(idb + row + j);
                m = m < (unsigned)N ? m : (unsigned)(N - 1);
                own += __ldg(xc + a2 * N + m) * __ldg(g_c1 + row + j) - __ldg(xc + a1 * N + m) * __ldg(g_c2 + row + j);
            }
        }
        return own;
    };
    auto edge_val = [&](int e) -> float {
        if (!STAGED && e >= E) return 0.f;
        float v = gd[e];
        if (CROSS && e < E) {
            const int nn = e / k;
            v += __ldg(xc + a1 * N + nn) * __ldg(g_c2 + e) - __ldg(xc + a2 * N + nn) * __ldg(g_c1 + e);
        }
        return v;
    };

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (R.hdr[0] == 0) {
        const int G = (N + 31) / 32;
        auto ell_sum = [&](int g, int s_begin, int s_end) -> float {
            const int* col = R.ell + __ldg(R.goff + g) + lane;
            float acc = 0.f;
            int s = s_begin;
            for (; s + 16 <= s_end; s += 16) {                // 16 independent coalesced loads in flight
                int e[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) e[q] = __ldg(col + (s + q) * 32);
#pragma unroll
                for (int q = 0; q < 16; ++q) acc += edge_val(e[q]);
            }
            if (s + 8 <= s_end) {
                int e[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) e[q] = __ldg(col + (s + q) * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc += edge_val(e[q]);
                s += 8;
            }
            for (; s < s_end; ++s) acc += edge_val(__ldg(col + s * 32));
            return acc;
        };
        // groups are sorted by in-degree: the few hub groups first, all warps sharing each of them (partials
        // combined in warp order, so the sum stays deterministic), then one warp per ordinary group
        constexpr int kBig = 96;
        __shared__ float part[32][32];
        int g0 = 0;
        for (; g0 < G && __ldg(R.gslots + g0) > kBig; ++g0) {
            const int slots = __ldg(R.gslots + g0);
            const int per = (slots + nwarp - 1) / nwarp;
            part[warp][lane] = ell_sum(g0, min(warp * per, slots), min((warp + 1) * per, slots));
            __syncthreads();
            if (warp == 0) {
                float acc = 0.f;
                for (int w = 0; w < nwarp; ++w) acc += part[w][lane];
                const int r = g0 * 32 + lane;
                if (r < N) {
                    const int t = __ldg(R.perm + r);
                    out[t] = own_row(t) + acc;
                }
            }
            __syncthreads();
        }
        for (int g = g0 + warp; g < G; g += nwarp) {
            const float acc = ell_sum(g, 0, __ldg(R.gslots + g));
            const int r = g * 32 + lane;
            if (r < N) {
                const int t = __ldg(R.perm + r);
                out[t] = own_row(t) + acc;
            }
        }
    } else {                                                 // CSR fallback: thread per target
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            float acc = 0.f;
            const int lo = __ldg(R.ptr + n), hi = __ldg(R.ptr + n + 1);
            for (int p = lo; p < hi; ++p) acc += edge_val(__ldg(R.ell + p));
            out[n] = own_row(n) + acc;
        }
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
    // enough CTAs for ~8 per SM, but at least 512 vector groups each (a CTA restages its 3N coordinates; with C = 1
    // and 2048 groups per CTA only 96 CTAs existed for 148 SMs)
    const size_t groups = vec ? E / 4 : E;
    int splits = (8 * sm_count() + B * C - 1) / (B * C);
    const int max_splits = (int)((groups + 511) / 512);
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
    if (B <= 0 || N <= 0 || k <= 0) return 0;
    const size_t general = hpcs::align_up((size_t)B * hpcs::rev_graph(nullptr, 0, N, k).ints_per_cloud * sizeof(int), 256);
    const size_t fast = hpcs::align_up(hpcs::edge_bwd_fast_workspace_bytes(B, N, k), 256);
    return general > fast ? general : fast;
}

int hpcs_edge_feat_bwd_is_fast(const float* gout, int N, int k, int cross) {
    return (!cross && N > 0 && k > 0 && hpcs::edge_bwd_fast_applicable(gout, N, k)) ? 1 : 0;
}

int hpcs_edge_rev_build(const int64_t* idx, int B, int N, int k, void* ws, size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!idx || !ws) return fail(HPCS_ERR_ARG, "edge_rev_build: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || B > 65535) return fail(HPCS_ERR_ARG, "edge_rev_build: bad shape");
    if (ws_bytes < hpcs_edge_feat_bwd_workspace_bytes(B, N, k)) return fail(HPCS_ERR_WORKSPACE, "edge_rev_build: workspace too small");
    if (!edge_bwd_fast_applicable(static_cast<const float*>(nullptr), N, k)) return fail(HPCS_ERR_ARG, "edge_rev_build: shape N=%d k=%d is not on the persistent-gather path", N, k);
    return edge_bwd_fast_build(idx, B, N, k, ws, as_stream(stream));
}

int hpcs_edge_feat_bwd_prebuilt_f32(const float* gout, const float* x, const int64_t* idx, int B, int C, int N, int k,
                                    int cross, float* gx, void* ws, size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!gout || !idx || !gx || !ws) return fail(HPCS_ERR_ARG, "edge_feat_bwd_prebuilt: null pointer");
    if (B <= 0 || C <= 0 || N <= 0 || k <= 0 || B > 65535) return fail(HPCS_ERR_ARG, "edge_feat_bwd_prebuilt: bad shape");
    if (ws_bytes < hpcs_edge_feat_bwd_workspace_bytes(B, N, k)) return fail(HPCS_ERR_WORKSPACE, "edge_feat_bwd_prebuilt: workspace too small");
    if (!cross && edge_bwd_fast_applicable(gout, N, k)) return edge_bwd_fast_gather(gout, idx, B, C, N, k, gx, ws, as_stream(stream));
    return hpcs_edge_feat_bwd_f32(gout, x, idx, B, C, N, k, cross, gx, ws, ws_bytes, stream);   // e.g. a misaligned gradient: rebuild
}

int hpcs_edge_feat_bwd_f32(const float* gout, const float* x, const int64_t* idx, int B, int C, int N, int k,
                           int cross, float* gx, void* ws, size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!gout || !idx || !gx || !ws || (cross && !x)) return fail(HPCS_ERR_ARG, "edge_feat_bwd: null pointer");
    if (B <= 0 || C <= 0 || N <= 0 || k <= 0) return fail(HPCS_ERR_ARG, "edge_feat_bwd: bad shape");
    if (ws_bytes < hpcs_edge_feat_bwd_workspace_bytes(B, N, k)) return fail(HPCS_ERR_WORKSPACE, "edge_feat_bwd: workspace too small");
    if (B > 65535) return fail(HPCS_ERR_ARG, "edge_feat_bwd: B > 65535");
    cudaStream_t st = as_stream(stream);
    const size_t E = (size_t)N * k;
    if (!cross && edge_bwd_fast_applicable(gout, N, k)) return edge_bwd_fast_run(gout, idx, B, C, N, k, gx, ws, st);
    if (!cross) {
        // beyond the persistent gather (a gradient plane no longer fits shared memory: N*k > ~45000) the plain scatter with fp32
        // reductions beats the reverse-CSR kernel below (measured, C=21: B=4 N=16384 k=20 1.00 TB/s against 0.49 / 0.19 TB/s of the
        // CSR kernel at N=4096 / 8192) and has no size limit; the CSR kernel stays for the cross features, which need x.
        cudaError_t ce = cudaMemsetAsync(gx, 0, sizeof(float) * (size_t)B * 3 * C * N, st);
        if (ce != cudaSuccess) return fail(HPCS_ERR_CUDA, "edge_feat_bwd: memset: %s", cudaGetErrorString(ce));
        edge_feat_bwd_scatter_kernel<<<dim3((N + 255) / 256, 3 * C, B), 256, 0, st>>>(gout, idx, C, N, k, gx);
        return check_launch("edge_feat_bwd_scatter_kernel");
    }
    int* wsi = static_cast<int*>(ws);
    {
        // warps per CTA: as many private histogram rows as fit in ~128 KB of shared memory
        const int G = (N + 31) / 32;
        int nw = 32;
        auto smem_for = [&](int w) { return ((size_t)w * N + (size_t)N + (N + 1) + (N + 2) + N + G + (G + 1)) * sizeof(int); };
        while (nw > 1 && ((size_t)nw * N * sizeof(int) > 128 * 1024 || smem_for(nw) > 220 * 1024)) nw >>= 1;
        const size_t smem = smem_for(nw);
        if (smem > 220 * 1024) return fail(HPCS_ERR_ARG, "edge_feat_bwd: cross features with N=%d (max ~11000)", N);
        cudaFuncSetAttribute(edge_rev_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge_rev_build_kernel<<<B, nw * 32, smem, st>>>(idx, N, k, wsi);
        int rc = check_launch("edge_rev_build_kernel");
        if (rc) return rc;
    }
    dim3 grid(3 * C, B), block(1024);
    const int vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(gout) & 15) == 0);
    const size_t smem = (E + 4) * sizeof(float) + (vec ? E / 4 * sizeof(float) : 0);
    const bool staged = smem <= 200 * 1024;
#define HPCS_LAUNCH_BWD(S, X)                                                                          \
    {                                                                                                  \
        auto kern = edge_feat_bwd_kernel<S, X>;                                                        \
        if (S) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        kern<<<grid, block, S ? smem : 0, st>>>(gout, x, idx, wsi, C, N, k, vec, gx);                  \
    }
    if (staged) { if (cross) HPCS_LAUNCH_BWD(true, true) else HPCS_LAUNCH_BWD(true, false) }
    else        { if (cross) HPCS_LAUNCH_BWD(false, true) else HPCS_LAUNCH_BWD(false, false) }
#undef HPCS_LAUNCH_BWD
    return check_launch("edge_feat_bwd_kernel");
}

}  // extern "C"
