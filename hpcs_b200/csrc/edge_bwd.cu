// Edge-feature backward, fast path (no cross term): a persistent, TMA-fed gather.
//
//   gx[b,ch,n] = sum_j gctr[b,ch,n,j] - sum_j gdiff[b,ch,n,j] + sum_{(n',j): idx[b,n',j] = n} gdiff[b,ch,n',j]
//
// replaces the index_put_(accumulate=True) scatter that autograd derives for
// hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:13-41.  Two kernels per call:
//
//   edge_rev2_build_kernel   (grid S x B)  reverse graph of one cloud, S slices of <= 256 targets each.  A
//       slice CTA scans all N*k edges, records "source row n' points at target t" in a bitmap
//       bm[t][n'/32] (shared-memory atomicOr: order independent), turns the per-word popcounts into
//       prefixes, and so gets the position of every edge inside its target's list -- ascending in n',
//       i.e. a fixed summation order -- without sorting or warp-match.  Targets are ranked by in-degree
//       inside the slice and stored as sliced ELL (32 targets of similar degree per group, 16-bit edge
//       ids) so a warp reads the lists coalesced and kNN hub nodes do not unbalance it.
//       A duplicate (t, n') pair (impossible for kNN output, possible for a user-supplied idx) sets a
//       per-cloud flag; the gather then sums that cloud's planes with shared-memory atomics instead.
//
//   edge_bwd_gather_kernel   (one CTA per SM, persistent)  every CTA owns a contiguous range of
//       (cloud, channel, component) planes.  The two gradient planes of an item (N*k floats each, 80 KB at
//       N=1024, k=20) stream through a double-buffered shared-memory ring with cp.async.bulk (TMA) +
//       mbarriers: while plane q is consumed, plane q+1 is in flight, so HBM stays busy.  The centre plane
//       is reduced to row sums; the difference plane is gathered through the ELL lists, whose 16-bit
//       entries are cached in shared memory for the whole cloud.  Every input byte is read from HBM
//       exactly once, the output is written once: the kernel's roofline is HBM.
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace hpcs {

constexpr int kRevThreads = 512;
constexpr int kGatherThreads = 512;
constexpr int kSliceTargets = 256;
constexpr int kMaxChunks = 96;         // work items of one difference plane (pieces of ELL columns)

struct Rev2Dims {
    int S, NS, GS, W, G;      // slices, targets per slice, groups per slice, bitmap words per target, groups per cloud
    size_t cap;               // ELL capacity (entries) per slice
    size_t off_goff, off_perm, off_gslots, off_ell, bytes_per_cloud;
};

__host__ __device__ inline Rev2Dims rev2_dims(int N, int k) {
    Rev2Dims d;
    const size_t E = (size_t)N * k;
    d.NS = N < kSliceTargets ? (N + 31) / 32 * 32 : kSliceTargets;
    d.S = (N + d.NS - 1) / d.NS;
    d.GS = d.NS / 32;
    d.G = d.S * d.GS;
    d.W = (N + 31) / 32;
    d.cap = (E + 32 * (size_t)N + 63) / 64 * 64;
    size_t off = 64;                                        // hdr: [s] duplicate flag of slice s (S <= 8)
    d.off_goff = off;    off += (size_t)d.G * sizeof(int);
    d.off_perm = off;    off += (size_t)d.S * d.NS * sizeof(uint16_t);
    d.off_gslots = off;  off += (size_t)d.G * sizeof(uint16_t);
    off = (off + 15) / 16 * 16;
    d.off_ell = off;     off += (size_t)d.S * d.cap * sizeof(uint16_t);
    d.bytes_per_cloud = (off + 255) / 256 * 256;
    return d;
}

struct Rev2 {
    int* hdr;
    int* goff;            // [G] entry offset of the group inside the cloud's ell array
    uint16_t* perm;       // [S*NS] target with degree rank r of slice s at [s*NS + r] (>= N: padding)
    uint16_t* gslots;     // [G] list length (largest in-degree) of the group
    uint16_t* ell;        // [S*cap]
};

__host__ __device__ inline Rev2 rev2_view(void* ws, const Rev2Dims& d, int b) {
    char* p = static_cast<char*>(ws) + (size_t)b * d.bytes_per_cloud;
    Rev2 r;
    r.hdr = reinterpret_cast<int*>(p);
    r.goff = reinterpret_cast<int*>(p + d.off_goff);
    r.perm = reinterpret_cast<uint16_t*>(p + d.off_perm);
    r.gslots = reinterpret_cast<uint16_t*>(p + d.off_gslots);
    r.ell = reinterpret_cast<uint16_t*>(p + d.off_ell);
    return r;
}

// ------------------------------------------------------------------------------------------------
// reverse graph
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRevThreads)
edge_rev2_build_kernel(const int64_t* __restrict__ idx, int N, int k, unsigned k_magic, void* __restrict__ ws) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const Rev2Dims d = rev2_dims(N, k);
    const int s = blockIdx.x, b = blockIdx.y;
    const int E = N * k;
    const int t0 = s * d.NS;
    const int W = d.W, NS = d.NS, GS = d.GS;
    unsigned* bm = reinterpret_cast<unsigned*>(sm_raw);                 // [NS][W]
    uint16_t* pre = reinterpret_cast<uint16_t*>(bm + (size_t)NS * W);   // [NS][W] exclusive popcount prefix
    int* deg = reinterpret_cast<int*>(pre + (size_t)NS * W);            // [NS]
    int* rank = deg + NS;                                               // [NS]
    int* byrank = rank + NS;                                            // [NS] inverse of rank
    int* gof = byrank + NS;                                             // [GS + 1] (entries)
    const int64_t* idb = idx + (size_t)b * E;
    const Rev2 R = rev2_view(ws, d, b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

    for (int i = threadIdx.x; i < NS * W; i += blockDim.x) bm[i] = 0u;
    __syncthreads();
    // 1. bitmap of (target, source row) pairs of this slice (index loads issued 8 at a time: L2 latency overlaps)
    bool dup = false;
    for (int e0 = threadIdx.x; e0 < E; e0 += 8 * blockDim.x) {
        unsigned tq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            tq[u] = e < E ? (unsigned)__ldg(idb + e) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e >= E) break;
            const unsigned t = tq[u] < (unsigned)N ? tq[u] : (unsigned)(N - 1);
            const unsigned tt = t - (unsigned)t0;
            if (tt < (unsigned)NS) {
                const int n = (int)__umulhi((unsigned)e, k_magic);      // e / k (exact: e < 2^16)
                const unsigned bit = 1u << (n & 31);
                const unsigned old = atomicOr(&bm[tt * W + (n >> 5)], bit);
                dup |= (old & bit) != 0u;
            }
        }
    }
    const int any_dup = __syncthreads_or(dup ? 1 : 0);
    if (threadIdx.x == 0) R.hdr[s] = any_dup;
    // 2. per target: exclusive prefix of the word popcounts, in-degree
    for (int tt = warp; tt < NS; tt += nwarp) {
        int carry = 0;
        for (int w0 = 0; w0 < W; w0 += 32) {
            const int w = w0 + lane;
            const int c = w < W ? __popc(bm[tt * W + w]) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += v;
            }
            if (w < W) pre[tt * W + w] = (uint16_t)(carry + inc - c);
            carry += __shfl_sync(kFull, inc, 31);
        }
        if (lane == 0) deg[tt] = carry;
    }
    __syncthreads();
    // 3. rank by in-degree (descending, ties by target id): NS <= 256, a quadratic count is trivial
    for (int tt = threadIdx.x; tt < NS; tt += blockDim.x) {
        const int dt = deg[tt];
        int r = 0;
        for (int o = 0; o < NS; ++o) {
            const int dv = deg[o];
            r += (dv > dt) || (dv == dt && o < tt);
        }
        rank[tt] = r;
        byrank[r] = tt;
        R.perm[s * NS + r] = (uint16_t)(t0 + tt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int g = 0; g < GS; ++g) {
            const int slots = deg[byrank[g * 32]];                      // largest in the group
            gof[g] = run;
            R.gslots[s * GS + g] = (uint16_t)slots;
            R.goff[s * GS + g] = (int)(s * d.cap) + run;
            run += 32 * slots;
        }
        gof[GS] = run;
    }
    __syncthreads();
    uint16_t* ell = R.ell + (size_t)s * d.cap;
    const int total = gof[GS];
    for (int i = threadIdx.x; i < total; i += blockDim.x) ell[i] = (uint16_t)E;     // sentinel: points at a zero
    __syncthreads();
    // 4. placement: position of an edge in its target's list = number of smaller source rows with that target
    for (int e0 = threadIdx.x; e0 < E; e0 += 8 * blockDim.x) {
        unsigned tq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            tq[u] = e < E ? (unsigned)__ldg(idb + e) : 0xffffffffu;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e >= E) break;
            const unsigned t = tq[u] < (unsigned)N ? tq[u] : (unsigned)(N - 1);
            const unsigned tt = t - (unsigned)t0;
            if (tt < (unsigned)NS) {
                const int n = (int)__umulhi((unsigned)e, k_magic);      // e / k (exact: e < 2^16)
                const unsigned word = bm[tt * W + (n >> 5)];
                const int pos = pre[tt * W + (n >> 5)] + __popc(word & ((1u << (n & 31)) - 1u));
                const int r = rank[tt];
                ell[gof[r >> 5] + pos * 32 + (r & 31)] = (uint16_t)e;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// persistent gather
// ------------------------------------------------------------------------------------------------
struct GatherSmem {
    size_t off_buf0, off_buf1, off_rsum, off_part, off_perm, off_gsl, off_goff, off_coff, off_cbase, off_cdesc, off_bars, off_misc, off_cache, total;
    int cache_entries;
};

__host__ __device__ inline GatherSmem gather_smem(int N, int k, const Rev2Dims& d, size_t budget) {
    GatherSmem g;
    const size_t E = (size_t)N * k;
    size_t off = 0;
    g.off_buf0 = off;  off += (E + 4) * sizeof(float);
    g.off_buf1 = off;  off += (E + 4) * sizeof(float);
    g.off_rsum = off;  off += (size_t)N * sizeof(float);
    g.off_goff = off;  off += (size_t)d.G * sizeof(int);
    g.off_coff = off;  off += (size_t)d.G * sizeof(int);
    g.off_part = off;  off += (size_t)kMaxChunks * 32 * sizeof(float);
    g.off_cbase = off; off += (size_t)(d.G + 1) * sizeof(int);
    g.off_cdesc = off; off += (size_t)kMaxChunks * sizeof(int);
    off = (off + 15) / 16 * 16;
    g.off_bars = off;  off += 2 * sizeof(uint64_t);
    g.off_misc = off;  off += 32;
    g.off_perm = off;  off += (size_t)d.S * d.NS * sizeof(uint16_t);
    g.off_gsl = off;   off += (size_t)d.G * sizeof(uint16_t);
    off = (off + 15) / 16 * 16;
    g.off_cache = off;
    g.cache_entries = budget > off ? (int)((budget - off) / sizeof(uint16_t)) : 0;
    const size_t want = (size_t)d.S * d.cap;                           // never need more than the whole ELL array
    if ((size_t)g.cache_entries > want) g.cache_entries = (int)want;
    g.cache_entries = g.cache_entries / 32 * 32;
    g.total = off + (size_t)g.cache_entries * sizeof(uint16_t);
    return g;
}

// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     ptx::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
                 : "memory");
}

// sum of pl[e] over one ELL column (lane-strided list of `slots` 16-bit edge ids), fixed order
template <int U, typename ColPtr>
__device__ __forceinline__ float gather_batch(const float* __restrict__ pl, ColPtr col, int s, float acc) {
    int e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = col[(s + u) * 32];
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = pl[e[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
    return acc;
}

template <typename ColPtr>
__device__ __forceinline__ float gather_column(const float* __restrict__ pl, ColPtr col, int slots) {
    float acc = 0.f;
    int s = 0;
    for (; s + 16 <= slots; s += 16) acc = gather_batch<16>(pl, col, s, acc);
    if (s + 8 <= slots) { acc = gather_batch<8>(pl, col, s, acc); s += 8; }
    if (s + 4 <= slots) { acc = gather_batch<4>(pl, col, s, acc); s += 4; }
    for (; s < slots; ++s) acc += pl[col[s * 32]];
    return acc;
}

__global__ void __launch_bounds__(kGatherThreads, 1)
edge_bwd_gather_kernel(const float* __restrict__ gout, const int64_t* __restrict__ idx, const void* __restrict__ ws,
                       int B, int C, int N, int k, int per_cta, size_t smem_budget, int chunk_bytes, int dbg_mode, float* __restrict__ gx) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const Rev2Dims d = rev2_dims(N, k);
    const GatherSmem L = gather_smem(N, k, d, smem_budget);
    float* const buf0 = reinterpret_cast<float*>(sm_raw + L.off_buf0);  // two planes of E + 4 floats, back to back
    const size_t buf_stride = (size_t)N * k + 4;
    float* rsum = reinterpret_cast<float*>(sm_raw + L.off_rsum);
    int* goff_s = reinterpret_cast<int*>(sm_raw + L.off_goff);
    int* coff_s = reinterpret_cast<int*>(sm_raw + L.off_coff);        // cache offset of a group, -1 = not cached
    float* part = reinterpret_cast<float*>(sm_raw + L.off_part);      // [chunk][32] partial column sums
    int* cbase_s = reinterpret_cast<int*>(sm_raw + L.off_cbase);      // [G+1] first chunk of a group
    int* cdesc_s = reinterpret_cast<int*>(sm_raw + L.off_cdesc);      // chunk -> group | (piece << 8)
    uint64_t* full = reinterpret_cast<uint64_t*>(sm_raw + L.off_bars);
    int* misc = reinterpret_cast<int*>(sm_raw + L.off_misc);          // [0] next chunk ticket, [1] duplicate flag, [2] chunks, [3] slots per chunk
    uint16_t* perm_s = reinterpret_cast<uint16_t*>(sm_raw + L.off_perm);
    uint16_t* gsl_s = reinterpret_cast<uint16_t*>(sm_raw + L.off_gsl);
    uint16_t* cache = reinterpret_cast<uint16_t*>(sm_raw + L.off_cache);

    const int E = N * k;
    const int items = B * 3 * C;
    const bool dbg_interleave = dbg_mode & 1, dbg_skip = dbg_mode & 2;
    const int lo = dbg_interleave ? 0 : blockIdx.x * per_cta;
    const int hi = dbg_interleave ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : min(items, lo + per_cta);
    if (lo >= hi) return;
    const int nloads = 2 * (hi - lo);
    auto item_of = [&](int q) { return dbg_interleave ? (int)blockIdx.x + (q >> 1) * (int)gridDim.x : lo + (q >> 1); };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t plane_bytes = (uint32_t)E * sizeof(float);

    auto plane_ptr = [&](int q) -> const float* {
        const int item = item_of(q);
        const int b = item / (3 * C), ch = item - b * 3 * C;
        const int c = ch / 3, a = ch - 3 * c;
        const size_t pl = (q & 1) ? (size_t)c * 3 + a : (size_t)(C + c) * 3 + a;    // even q: centre plane, odd q: difference plane
        return gout + ((size_t)b * 2 * C * 3 + pl) * E;
    };
    // one plane = ceil(plane_bytes / chunk_bytes) bulk copies on the same mbarrier, issued by the lanes of warp 0
    // (a single large cp.async.bulk keeps only a few KB in flight; many medium ones fill the memory pipe)
    auto issue = [&](int q) {
        if (lane == 0) ptx::mbar_arrive_expect_tx(full + (q & 1), plane_bytes);
        __syncwarp();
        const char* src = reinterpret_cast<const char*>(plane_ptr(q));
        char* dst = reinterpret_cast<char*>(buf0 + (size_t)(q & 1) * buf_stride);
        for (uint32_t o = (uint32_t)lane * chunk_bytes; o < plane_bytes; o += 32u * chunk_bytes)
            bulk_load(dst + o, src + o, min((uint32_t)chunk_bytes, plane_bytes - o), full + (q & 1));
    };

    if (threadIdx.x == 0) {
        ptx::mbar_init(full, 1);
        ptx::mbar_init(full + 1, 1);
        ptx::fence_barrier_init();
        buf0[E] = 0.f;                                               // the ELL sentinel reads this zero
        buf0[buf_stride + E] = 0.f;
    }
    __syncthreads();
    if (warp == 0) {
        issue(0);
        issue(1);
    }

    int cur_b = -1;
    for (int q = 0; q < nloads; ++q) {
        const int item = item_of(q);
        const int b = item / (3 * C);
        const float* pl = buf0 + (size_t)(q & 1) * buf_stride;
        if (b != cur_b && !dbg_skip) {
            // ---- per-cloud reverse-graph metadata + ELL cache (overlaps the planes already in flight) ----
            cur_b = b;
            const Rev2 R = rev2_view(const_cast<void*>(ws), d, b);
            for (int i = threadIdx.x; i < d.S * d.NS; i += blockDim.x) perm_s[i] = R.perm[i];
            for (int i = threadIdx.x; i < d.G; i += blockDim.x) { gsl_s[i] = R.gslots[i]; goff_s[i] = R.goff[i]; }
            if (threadIdx.x == 0) {
                int flag = 0;
                for (int i = 0; i < d.S; ++i) flag |= R.hdr[i];
                misc[1] = flag;
            }
            __syncthreads();
            if (threadIdx.x == 32) {
                // cut every ELL column into pieces of `chs` slots (a hub group is hundreds of slots long: left
                // whole, one warp would walk it alone while fifteen wait); at most kMaxChunks pieces
                int tot = 0;
                for (int i = 0; i < d.G; ++i) tot += gsl_s[i];
                int chs = (tot + (kMaxChunks - d.G) - 1) / (kMaxChunks - d.G);
                chs = chs < 32 ? 32 : (chs + 7) / 8 * 8;
                int n = 0;
                for (int i = 0; i < d.G; ++i) {
                    cbase_s[i] = n;
                    const int pieces = ((int)gsl_s[i] + chs - 1) / chs;
                    for (int c = 0; c < pieces; ++c) cdesc_s[n++] = i | (c << 8);
                }
                cbase_s[d.G] = n;
                misc[2] = n;
                misc[3] = chs;
            }
            if (threadIdx.x == 0) {
                int run = 0;
                for (int i = 0; i < d.G; ++i) {
                    const int n_ent = 32 * (int)gsl_s[i];
                    if (run + n_ent <= L.cache_entries) { coff_s[i] = run; run += n_ent; }
                    else coff_s[i] = -1;
                }
            }
            __syncthreads();
            for (int g = warp; g < d.G; g += (int)(blockDim.x >> 5)) {
                const int co = coff_s[g];
                if (co < 0) continue;
                const int n_ent = 32 * (int)gsl_s[g];
                const uint4* src = reinterpret_cast<const uint4*>(R.ell + goff_s[g]);      // group starts are 64-byte aligned
                uint4* dst = reinterpret_cast<uint4*>(cache + co);
                for (int i = lane; i < n_ent / 8; i += 32) dst[i] = __ldg(src + i);
            }
            __syncthreads();
        }
        ptx::mbar_wait(full + (q & 1), (q >> 1) & 1);
        if (dbg_skip) {
            if (threadIdx.x == 0 && pl[7] == 123.456f) gx[0] = 1.f;
        } else if ((q & 1) == 0) {
            // ---- centre plane: row sums ----
            if ((k & 3) == 0) {
                const int q4 = k >> 2;
                for (int n = threadIdx.x; n < N; n += blockDim.x) {
                    const float4* row = reinterpret_cast<const float4*>(pl + (size_t)n * k);
                    float acc = 0.f;
                    for (int i = 0; i < q4; ++i) { const float4 v = row[i]; acc += (v.x + v.y) + (v.z + v.w); }
                    rsum[n] = acc;
                }
            } else {
                for (int n = threadIdx.x; n < N; n += blockDim.x) {
                    float acc = 0.f;
                    for (int j = 0; j < k; ++j) acc += pl[(size_t)n * k + j];
                    rsum[n] = acc;
                }
            }
            if (threadIdx.x == 0) misc[0] = 0;
        } else {
            // ---- difference plane: gather through the reverse lists, combine, write ----
            const int ch = item - b * 3 * C;
            float* out = gx + ((size_t)b * 3 * C + ch) * N;
            auto own_diff = [&](int t) -> float {
                float acc = 0.f;
                if ((k & 3) == 0) {
                    const float4* row = reinterpret_cast<const float4*>(pl + (size_t)t * k);
                    for (int i = 0; i < (k >> 2); ++i) { const float4 v = row[i]; acc += (v.x + v.y) + (v.z + v.w); }
                } else {
                    for (int j = 0; j < k; ++j) acc += pl[(size_t)t * k + j];
                }
                return acc;
            };
            if (misc[1] == 0) {
                const Rev2 R = rev2_view(const_cast<void*>(ws), d, b);
                const int nchunks = misc[2], chs = misc[3];
                for (int ticket = warp; ticket < nchunks; ticket += (int)(blockDim.x >> 5)) {   // pieces are equal-sized: static split
                    const int desc = cdesc_s[ticket];
                    const int g = desc & 255, s0 = (desc >> 8) * chs;
                    const int cnt = min(chs, (int)gsl_s[g] - s0);
                    const int co = coff_s[g];
                    part[ticket * 32 + lane] = co >= 0 ? gather_column(pl, cache + co + s0 * 32 + lane, cnt)
                                                       : gather_column(pl, R.ell + goff_s[g] + s0 * 32 + lane, cnt);
                }
                __syncthreads();
                for (int p = threadIdx.x; p < d.G * 32; p += blockDim.x) {       // pieces of a column added in order
                    const int g = p >> 5;
                    float acc = 0.f;
                    for (int c = cbase_s[g]; c < cbase_s[g + 1]; ++c) acc += part[c * 32 + (p & 31)];
                    const int t = perm_s[p];
                    if (t < N) out[t] = (rsum[t] - own_diff(t)) + acc;
                }
            } else {
                // duplicate (target, row) pairs in idx: plain scatter with shared-memory atomics (any order)
                float* acc_s = reinterpret_cast<float*>(cache);          // cache is unused for such a cloud
                const bool fits = (size_t)L.cache_entries * sizeof(uint16_t) >= (size_t)N * sizeof(float);
                const int64_t* idb = idx + (size_t)b * E;
                if (fits) {
                    for (int n = threadIdx.x; n < N; n += blockDim.x) acc_s[n] = 0.f;
                    __syncthreads();
                    for (int e = threadIdx.x; e < E; e += blockDim.x) {
                        unsigned t = (unsigned)__ldg(idb + e);
                        t = t < (unsigned)N ? t : (unsigned)(N - 1);
                        atomicAdd(&acc_s[t], pl[e]);
                    }
                    __syncthreads();
                    for (int n = threadIdx.x; n < N; n += blockDim.x) out[n] = (rsum[n] - own_diff(n)) + acc_s[n];
                } else {
                    for (int n = threadIdx.x; n < N; n += blockDim.x) out[n] = rsum[n] - own_diff(n);
                    __syncthreads();
                    for (int e = threadIdx.x; e < E; e += blockDim.x) {
                        unsigned t = (unsigned)__ldg(idb + e);
                        t = t < (unsigned)N ? t : (unsigned)(N - 1);
                        atomicAdd(out + t, pl[e]);
                    }
                }
            }
        }
        __syncthreads();                                                // plane q fully consumed by every warp
        if (warp == 0 && q + 2 < nloads) issue(q + 2);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr size_t kSmemBudget = 226 * 1024;

bool edge_bwd_fast_applicable(const float* gout, int N, int k) {
    const size_t E = (size_t)N * k;
    if (N > 2048 || E > 65535 || (E & 3) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(gout) & 15) != 0) return false;
    const Rev2Dims d = rev2_dims(N, k);
    const GatherSmem g = gather_smem(N, k, d, kSmemBudget);
    return g.cache_entries >= 1024 && g.total <= kSmemBudget;
}

size_t edge_bwd_fast_workspace_bytes(int B, int N, int k) {
    const size_t E = (size_t)N * k;
    if (N > 2048 || E > 65535) return 0;
    return (size_t)B * rev2_dims(N, k).bytes_per_cloud;
}

int edge_bwd_fast_run(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, void* ws, cudaStream_t st) {
    const Rev2Dims d = rev2_dims(N, k);
    {
        const size_t smem = (size_t)d.NS * d.W * (sizeof(unsigned) + sizeof(uint16_t)) + (size_t)(3 * d.NS + d.GS + 1) * sizeof(int);
        cudaFuncSetAttribute(edge_rev2_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        edge_rev2_build_kernel<<<dim3(d.S, B), kRevThreads, smem, st>>>(idx, N, k, 0xFFFFFFFFu / (unsigned)k + 1u, ws);
        int rc = check_launch("edge_rev2_build_kernel");
        if (rc) return rc;
    }
    const GatherSmem g = gather_smem(N, k, d, kSmemBudget);
    const int items = B * 3 * C;
    int grid = sm_count();
    if (grid > items) grid = items;
    const int per_cta = (items + grid - 1) / grid;
    grid = (items + per_cta - 1) / per_cta;
    cudaFuncSetAttribute(edge_bwd_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.total);
    int chunk = 8192;                                                  // bytes per bulk copy (multiple of 16)
    if (const char* env = getenv("HPCS_BWD_CHUNK")) chunk = atoi(env) > 0 ? atoi(env) / 16 * 16 : chunk;   // tuning knob
    int dbg = 0;
    if (const char* env = getenv("HPCS_BWD_MODE")) dbg = atoi(env);
    edge_bwd_gather_kernel<<<grid, kGatherThreads, g.total, st>>>(gout, idx, ws, B, C, N, k, per_cta, kSmemBudget, chunk, dbg, gx);
    return check_launch("edge_bwd_gather_kernel");
}

}  // namespace hpcs
