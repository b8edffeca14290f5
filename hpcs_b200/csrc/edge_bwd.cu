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
constexpr int kGatherThreads = 512;             // (1024 threads measured slower: 136 vs 130 us -- the gather is bound by the shared-memory pipe)
constexpr int kSliceTargets = 256;
constexpr int kMaxChunks = 96;         // work items of one difference plane (pieces of ELL columns)

struct Rev2Dims {
    int S, NS, GS, W, G;      // slices, targets per slice, groups per slice, bitmap words per target, groups per cloud
    size_t cap;               // ELL capacity (entries) per slice
    size_t off_goff, off_gent, off_perm, off_rpos, off_gslots, off_ell, bytes_per_cloud;
};

__host__ __device__ inline Rev2Dims rev2_dims(int N, int k) {
    Rev2Dims d;
    const size_t E = (size_t)N * k;
    d.NS = N < kSliceTargets ? (N + 31) / 32 * 32 : kSliceTargets;   // (128-target slices measured slower: 143 vs 130 us)
    d.S = (N + d.NS - 1) / d.NS;
    d.GS = d.NS / 32;
    d.G = d.S * d.GS;
    d.W = (N + 31) / 32;
    d.cap = (E + 32 * (size_t)N + 63) / 64 * 64;
    size_t off = 64;                                        // hdr: [s] duplicate flag, [8 + s] entries used, of slice s (S <= 8)
    d.off_goff = off;    off += (size_t)d.G * sizeof(int);
    d.off_gent = off;    off += (size_t)d.G * sizeof(int);
    d.off_perm = off;    off += (size_t)d.S * d.NS * sizeof(uint16_t);
    d.off_rpos = off;    off += (size_t)d.S * d.NS * sizeof(uint16_t);
    d.off_gslots = off;  off += (size_t)d.G * sizeof(uint16_t);
    off = (off + 15) / 16 * 16;
    d.off_ell = off;     off += (size_t)d.S * d.cap * sizeof(uint16_t);
    d.bytes_per_cloud = (off + 255) / 256 * 256;
    return d;
}

struct Rev2 {
    int* hdr;
    int* goff;            // [G] entry offset of the group inside the cloud's ell array
    int* gent;            // [G] entries the group occupies (multiple of 32)
    uint16_t* perm;       // [S*NS] target with degree rank r of slice s at [s*NS + r] (>= N: padding)
    uint16_t* rpos;       // [S*NS] inverse: target t -> s*NS + r
    uint16_t* gslots;     // [G] longest list of the group (bits 0-14); bit 15: stored as per-lane lists, not ELL
    uint16_t* ell;        // [S*cap]
};

__host__ __device__ inline Rev2 rev2_view(void* ws, const Rev2Dims& d, int b) {
    char* p = static_cast<char*>(ws) + (size_t)b * d.bytes_per_cloud;
    Rev2 r;
    r.hdr = reinterpret_cast<int*>(p);
    r.goff = reinterpret_cast<int*>(p + d.off_goff);
    r.gent = reinterpret_cast<int*>(p + d.off_gent);
    r.perm = reinterpret_cast<uint16_t*>(p + d.off_perm);
    r.rpos = reinterpret_cast<uint16_t*>(p + d.off_rpos);
    r.gslots = reinterpret_cast<uint16_t*>(p + d.off_gslots);
    r.ell = reinterpret_cast<uint16_t*>(p + d.off_ell);
    return r;
}

// ------------------------------------------------------------------------------------------------
// reverse graph
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRevThreads)
edge_rev2_build_kernel(const int64_t* __restrict__ idx, int N, int k, unsigned k_magic, void* __restrict__ ws) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
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
    int* gform = gof + GS + 1;                                          // [GS] 1 = per-lane lists, 0 = ELL
    int* gent = gform + GS;                                             // [GS] entries of the group
    int* lofs = gent + GS;                                              // [NS] list offset of rank r inside its group
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
        R.rpos[t0 + tt] = (uint16_t)(s * NS + r);
    }
    __syncthreads();
    // 3b. storage form per group of 32 ranks.  ELL (slot-major, padded to the longest list) reads coalesced, but a
    //     hub group (lists of 200 next to lists of 40 in feature-space kNN graphs) would be mostly padding: such a
    //     group is stored as 32 back-to-back lists behind a 64-entry header (offset and length per lane) instead.
    if (threadIdx.x < GS) {
        const int g = threadIdx.x;
        int sum = 0;
        for (int l = 0; l < 32; ++l) sum += deg[byrank[g * 32 + l]];
        const int slots = deg[byrank[g * 32]];                          // largest in the group
        const bool lists = 32 * slots > sum + (sum >> 2) + 64;          // > 25 % padding (+ header) -> per-lane lists
        gform[g] = lists ? 1 : 0;
        gent[g] = lists ? (64 + sum + 31) / 32 * 32 : 32 * slots;
        if (lists) {
            int run = 64;
            for (int l = 0; l < 32; ++l) { lofs[g * 32 + l] = run; run += deg[byrank[g * 32 + l]]; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int g = 0; g < GS; ++g) {
            gof[g] = run;
            R.gslots[s * GS + g] = (uint16_t)(deg[byrank[g * 32]] | (gform[g] << 15));
            R.goff[s * GS + g] = (int)(s * d.cap) + run;
            R.gent[s * GS + g] = gent[g];
            run += gent[g];
        }
        gof[GS] = run;
        R.hdr[8 + s] = run;
    }
    __syncthreads();
    uint16_t* ell = R.ell + (size_t)s * d.cap;
    const int total = gof[GS];
    for (int i = threadIdx.x; i < total; i += blockDim.x) ell[i] = (uint16_t)E;     // sentinel: points at a zero
    __syncthreads();
    for (int r = threadIdx.x; r < NS; r += blockDim.x) {                // headers of the list-form groups
        const int g = r >> 5;
        if (gform[g]) {
            ell[gof[g] + (r & 31)] = (uint16_t)lofs[r];
            ell[gof[g] + 32 + (r & 31)] = (uint16_t)deg[byrank[r]];
        }
    }
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
                const int g = r >> 5;
                ell[gof[g] + (gform[g] ? lofs[r] + pos : pos * 32 + (r & 31))] = (uint16_t)e;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// persistent gather
// ------------------------------------------------------------------------------------------------
struct GatherSmem {
    size_t off_buf0, off_rsum, off_part, off_meta, off_coff, off_cbase, off_cdesc, off_bars, off_misc, off_cache, total;
    int cache_entries;
};

__host__ __device__ inline GatherSmem gather_smem(int N, int k, const Rev2Dims& d, size_t budget, int nbuf) {
    GatherSmem g;
    const size_t E = (size_t)N * k;
    size_t off = 0;
    g.off_buf0 = off;  off += (size_t)nbuf * (E + 4) * sizeof(float);
    g.off_rsum = off;  off += (size_t)N * sizeof(float);
    g.off_part = off;  off += (size_t)kMaxChunks * 32 * sizeof(float);
    g.off_coff = off;  off += (size_t)d.G * sizeof(int);
    g.off_cbase = off; off += (size_t)(d.G + 1) * sizeof(int);
    g.off_cdesc = off; off += (size_t)kMaxChunks * sizeof(int);
    off = (off + 15) / 16 * 16;
    g.off_bars = off;  off += 4 * sizeof(uint64_t);
    g.off_misc = off;  off += 32;
    off = (off + 127) / 128 * 128;
    g.off_meta = off;  off += d.off_ell;                                // mirror of the head of the cloud's record
    off = (off + 127) / 128 * 128;
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

// ELL form: sum of pl[e] over one column (lane-strided list of `slots` 16-bit edge ids), fixed order
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

// list form: this lane's list starts at grp[off] and has `len` entries; entries [s0, s0 + cnt) of it
template <int U, typename ColPtr>
__device__ __forceinline__ float gather_list_batch(const float* __restrict__ pl, ColPtr lst, int s, int len, int sentinel, float acc) {
    int e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) e[u] = s + u < len ? (int)lst[s + u] : sentinel;
    float v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = pl[e[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u];
    return acc;
}

template <typename ColPtr>
__device__ __forceinline__ float gather_lists(const float* __restrict__ pl, ColPtr grp, int lane, int s0, int cnt, int sentinel) {
    const int off = grp[lane], len = grp[32 + lane];
    float acc = 0.f;
    int s = s0;
    for (; s + 16 <= s0 + cnt; s += 16) acc = gather_list_batch<16>(pl, grp + off, s, len, sentinel, acc);
    for (; s + 4 <= s0 + cnt; s += 4) acc = gather_list_batch<4>(pl, grp + off, s, len, sentinel, acc);
    for (; s < s0 + cnt; ++s) acc += pl[s < len ? (int)grp[off + s] : sentinel];
    return acc;
}

__device__ __forceinline__ float row_sum(const float* __restrict__ row, int k) {
    float acc = 0.f;
    if ((k & 3) == 0) {
        const float4* r4 = reinterpret_cast<const float4*>(row);
        for (int i = 0; i < (k >> 2); ++i) { const float4 v = r4[i]; acc += (v.x + v.y) + (v.z + v.w); }
    } else {
        for (int j = 0; j < k; ++j) acc += row[j];
    }
    return acc;
}

// (Two CTAs per SM with one plane buffer each was measured slower: the gather phase is bound by shared-memory
// wavefronts -- random 4-byte reads, ~3 bank conflicts each -- which a second resident CTA only contends for.
// Also measured slower, 144 us against 130 us: streaming only the difference planes through the two buffers and
// taking the centre plane's row sums straight from HBM into registers before the gather (coalesced LDG.128, partial
// sums redistributed by shuffles in the combine).  The L1 fills of those loads go through the same l1tex data pipe as
// the gather's shared-memory reads and slow it by more than the removed centre phase saved; TMA writes do not.)
// NBUF = 2: the two planes of an item (and the next item's) alternate between two buffers, so a load is always in flight
// while a plane is consumed.  NBUF = 1: one buffer, for shapes whose planes are too large for two (N*k up to 2^16 - 1:
// k = 40, or N = 2048): loads and compute of a CTA alternate, the other SMs fill the gaps on the HBM side.
template <int NBUF>
__global__ void __launch_bounds__(kGatherThreads, 1)
edge_bwd_gather_kernel(const float* __restrict__ gout, const int64_t* __restrict__ idx, const void* __restrict__ ws,
                       int B, int C, int N, int k, int per_cta, size_t smem_budget, long long* __restrict__ prof,
                       float* __restrict__ gx) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const Rev2Dims d = rev2_dims(N, k);
    const GatherSmem L = gather_smem(N, k, d, smem_budget, NBUF);
    float* const buf0 = reinterpret_cast<float*>(sm_raw + L.off_buf0);  // two planes of E + 4 floats, back to back
    const size_t buf_stride = (size_t)N * k + 4;
    float* rsum = reinterpret_cast<float*>(sm_raw + L.off_rsum);
    float* part = reinterpret_cast<float*>(sm_raw + L.off_part);      // [piece][32] partial column sums
    unsigned char* meta_s = sm_raw + L.off_meta;                      // same layout as the record in the workspace
    const int* hdr_s = reinterpret_cast<const int*>(meta_s);
    const int* goff_s = reinterpret_cast<const int*>(meta_s + d.off_goff);
    const uint16_t* rpos_s = reinterpret_cast<const uint16_t*>(meta_s + d.off_rpos);   // target -> rank position
    const uint16_t* gsl_s = reinterpret_cast<const uint16_t*>(meta_s + d.off_gslots);
    int* coff_s = reinterpret_cast<int*>(sm_raw + L.off_coff);        // cache offset of a group, -1 = not cached
    int* cbase_s = reinterpret_cast<int*>(sm_raw + L.off_cbase);      // [G+1] first piece of a group
    int* cdesc_s = reinterpret_cast<int*>(sm_raw + L.off_cdesc);      // piece -> group | (index in group << 8)
    uint64_t* full = reinterpret_cast<uint64_t*>(sm_raw + L.off_bars);// [0,1] planes, [2] list cache, [3] record head
    int* misc = reinterpret_cast<int*>(sm_raw + L.off_misc);          // [1] duplicate flag, [2] pieces, [3] slots per piece
    uint16_t* cache = reinterpret_cast<uint16_t*>(sm_raw + L.off_cache);

    const int E = N * k;
    const int items = B * 3 * C;
    const int lo = blockIdx.x * per_cta;
    const int hi = min(items, lo + per_cta);
    if (lo >= hi) return;
    const int nloads = 2 * (hi - lo);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t plane_bytes = (uint32_t)E * sizeof(float);

    auto plane_ptr = [&](int q) -> const float* {
        const int item = lo + (q >> 1);
        const int b = item / (3 * C), ch = item - b * 3 * C;
        const size_t pl = (q & 1) ? (size_t)ch : (size_t)(3 * C + ch);       // even q: centre plane, odd q: difference plane
        return gout + ((size_t)b * 2 * C * 3 + pl) * E;
    };
    auto issue = [&](int q) {                                                   // thread 0 only
        ptx::mbar_arrive_expect_tx(full + (q % NBUF), plane_bytes);
        bulk_load(buf0 + (size_t)(q % NBUF) * buf_stride, plane_ptr(q), plane_bytes, full + (q % NBUF));
    };

    if (threadIdx.x == 0) {
        ptx::mbar_init(full, 1);
        ptx::mbar_init(full + 1, 1);
        ptx::mbar_init(full + 2, 1);
        ptx::mbar_init(full + 3, 1);
        ptx::fence_barrier_init();
        for (int i = 0; i < NBUF; ++i) buf0[i * buf_stride + E] = 0.f;   // the sentinel entries read this zero
        for (int i = 0; i < NBUF && i < nloads; ++i) issue(i);
    }
    __syncthreads();

    // phase timing of thread 0 (cycles in [record tail, wait, centre, gather, combine, issue, record load]); compiled
    // in only with -DHPCS_BWD_PROFILE (tools/tune_bwd.py), `prof` is a device buffer of 8 int64 per CTA
#ifdef HPCS_BWD_PROFILE
    long long tprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark = clock64();
    auto lap = [&](int slot) {
        if (prof && threadIdx.x == 0) { const long long now = clock64(); tprof[slot] += now - tmark; tmark = now; }
    };
#else
    auto lap = [](int) {};
#endif
    int cur_b = -1, clouds_seen = 0, meta_loads = 0;
    bool cache_pending = false;
    for (int q = 0; q < nloads; ++q) {
        const int item = lo + (q >> 1);
        const int b = item / (3 * C);
        const float* pl = buf0 + (size_t)(q % NBUF) * buf_stride;
        if (b != cur_b) {
            // ---- reverse graph of a new cloud: small tables by plain loads, the 16-bit lists by one bulk copy per
            //      slice into the cache (waited for just before the first gather; the planes are already in flight)
            cur_b = b;
            const Rev2 R = rev2_view(const_cast<void*>(ws), d, b);
            if (threadIdx.x == 0) {
                ptx::mbar_arrive_expect_tx(full + 3, (uint32_t)d.off_ell);
                bulk_load(meta_s, R.hdr, (uint32_t)d.off_ell, full + 3);
            }
            ptx::mbar_wait(full + 3, meta_loads & 1);
            ++meta_loads;
            lap(6);
            if (threadIdx.x == 0) {
                int flag = 0, run = 0;
                uint32_t bytes = 0;
                for (int sl = 0; sl < d.S; ++sl) {
                    flag |= hdr_s[sl];
                    const int n_ent = hdr_s[8 + sl];                                 // entries used by the slice
                    const bool fits = n_ent > 0 && run + n_ent <= L.cache_entries;
                    for (int g = sl * d.GS; g < (sl + 1) * d.GS; ++g)
                        coff_s[g] = fits ? run + (goff_s[g] - (int)(sl * d.cap)) : -1;
                    if (fits) { bytes += (uint32_t)n_ent * 2u; run += n_ent; }
                }
                misc[1] = flag;
                if (bytes) {
                    ptx::mbar_arrive_expect_tx(full + 2, bytes);
                    int at = 0;
                    for (int sl = 0; sl < d.S; ++sl) {
                        const int n_ent = hdr_s[8 + sl];
                        if (coff_s[sl * d.GS] < 0) continue;
                        bulk_load(cache + at, R.ell + (size_t)sl * d.cap, (uint32_t)n_ent * 2u, full + 2);
                        at += n_ent;
                    }
                }
                misc[4] = bytes ? 1 : 0;
            } else if (warp == 1) {
                // cut every column into pieces of `chs` slots (a hub group is hundreds of slots long: left whole, one
                // warp would walk it alone while fifteen wait); at most kMaxChunks pieces.  G <= 64: two groups a lane.
                const int s_a = lane < d.G ? (int)(gsl_s[lane] & 0x7fff) : 0;
                const int s_b = lane + 32 < d.G ? (int)(gsl_s[lane + 32] & 0x7fff) : 0;
                int tot = s_a + s_b;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFull, tot, o);
                int chs = (tot + (kMaxChunks - d.G) - 1) / (kMaxChunks - d.G);
                chs = chs < 32 ? 32 : (chs + 7) / 8 * 8;
                const int p_a = (s_a + chs - 1) / chs, p_b = (s_b + chs - 1) / chs;
                int inc_a = p_a, inc_b = p_b;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int va = __shfl_up_sync(kFull, inc_a, o), vb = __shfl_up_sync(kFull, inc_b, o);
                    if (lane >= o) { inc_a += va; inc_b += vb; }
                }
                const int tot_a = __shfl_sync(kFull, inc_a, 31), tot_b = __shfl_sync(kFull, inc_b, 31);
                const int st_a = inc_a - p_a, st_b = tot_a + inc_b - p_b;
                if (lane < d.G) { cbase_s[lane] = st_a; for (int c = 0; c < p_a; ++c) cdesc_s[st_a + c] = lane | (c << 8); }
                if (lane + 32 < d.G) { cbase_s[lane + 32] = st_b; for (int c = 0; c < p_b; ++c) cdesc_s[st_b + c] = (lane + 32) | (c << 8); }
                if (lane == 0) { cbase_s[d.G] = tot_a + tot_b; misc[2] = tot_a + tot_b; misc[3] = chs; }
            }
            __syncthreads();
            cache_pending = misc[4] != 0;
        }
        lap(0);
        ptx::mbar_wait(full + (q % NBUF), (q / NBUF) & 1);
        lap(1);
        if ((q & 1) == 0) {
            // ---- centre plane: row sums ----
            for (int n = threadIdx.x; n < N; n += blockDim.x) rsum[n] = row_sum(pl + (size_t)n * k, k);
        } else {
            // ---- difference plane: gather through the reverse lists, combine, write ----
            const int ch = item - b * 3 * C;
            float* out = gx + ((size_t)b * 3 * C + ch) * N;
            if (misc[1] == 0) {
                if (cache_pending) {
                    ptx::mbar_wait(full + 2, clouds_seen & 1);
                    cache_pending = false;
                    ++clouds_seen;
                }
                const Rev2 R = rev2_view(const_cast<void*>(ws), d, b);
                const int nchunks = misc[2], chs = misc[3];
                for (int piece = warp; piece < nchunks; piece += (int)(blockDim.x >> 5)) {   // equal-sized: static split
                    const int desc = cdesc_s[piece];
                    const int g = desc & 255, s0 = (desc >> 8) * chs;
                    const int slots = gsl_s[g] & 0x7fff, lists = gsl_s[g] >> 15;
                    const int cnt = min(chs, slots - s0);
                    const int co = coff_s[g];
                    float acc;
                    if (lists) acc = co >= 0 ? gather_lists(pl, cache + co, lane, s0, cnt, E) : gather_lists(pl, R.ell + goff_s[g], lane, s0, cnt, E);
                    else acc = co >= 0 ? gather_column(pl, cache + co + s0 * 32 + lane, cnt)
                                       : gather_column(pl, R.ell + goff_s[g] + s0 * 32 + lane, cnt);
                    part[piece * 32 + lane] = acc;
                }
                __syncthreads();
                lap(3);
                for (int n = threadIdx.x; n < N; n += blockDim.x) {            // target order: coalesced, conflict-free rows
                    const int p = rpos_s[n], g = p >> 5;
                    float acc = 0.f;
                    for (int c = cbase_s[g]; c < cbase_s[g + 1]; ++c) acc += part[c * 32 + (p & 31)];   // pieces in order
                    out[n] = (rsum[n] - row_sum(pl + (size_t)n * k, k)) + acc;
                }
            } else {
                // duplicate (target, row) pairs in idx: plain scatter with shared-memory atomics (any order)
                float* acc_s = reinterpret_cast<float*>(cache);          // the cache is unused for such a cloud
                const bool fits = (size_t)L.cache_entries * sizeof(uint16_t) >= (size_t)N * sizeof(float);
                const int64_t* idb = idx + (size_t)b * E;
                if (cache_pending) {                                     // slices without duplicates may have been requested
                    ptx::mbar_wait(full + 2, clouds_seen & 1);
                    cache_pending = false;
                    ++clouds_seen;
                    __syncthreads();
                }
                if (fits) {
                    for (int n = threadIdx.x; n < N; n += blockDim.x) acc_s[n] = 0.f;
                    __syncthreads();
                    for (int e = threadIdx.x; e < E; e += blockDim.x) {
                        unsigned t = (unsigned)__ldg(idb + e);
                        t = t < (unsigned)N ? t : (unsigned)(N - 1);
                        atomicAdd(&acc_s[t], pl[e]);
                    }
                    __syncthreads();
                    for (int n = threadIdx.x; n < N; n += blockDim.x) out[n] = (rsum[n] - row_sum(pl + (size_t)n * k, k)) + acc_s[n];
                } else {
                    for (int n = threadIdx.x; n < N; n += blockDim.x) out[n] = rsum[n] - row_sum(pl + (size_t)n * k, k);
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
        lap((q & 1) ? 4 : 2);
        if (threadIdx.x == 0 && q + NBUF < nloads) issue(q + NBUF);
        lap(5);
    }
#ifdef HPCS_BWD_PROFILE
    if (prof && threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) prof[blockIdx.x * 8 + i] = tprof[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
constexpr size_t kSmemBudget = 226 * 1024;

// plane buffers the persistent gather can afford for this shape: 2, 1, or 0 (= not on this path)
static int gather_buffers(int N, int k) {
    const Rev2Dims d = rev2_dims(N, k);
    for (int nbuf = 2; nbuf >= 1; --nbuf) {
        const GatherSmem g = gather_smem(N, k, d, kSmemBudget, nbuf);
        if (g.cache_entries >= 1024 && g.total <= kSmemBudget) return nbuf;
    }
    return 0;
}

bool edge_bwd_fast_applicable(const float* gout, int N, int k) {
    const size_t E = (size_t)N * k;
    if (N > 2048 || E > 65535 || (E & 3) != 0) return false;
    if ((reinterpret_cast<uintptr_t>(gout) & 15) != 0) return false;
    return gather_buffers(N, k) != 0;
}

size_t edge_bwd_fast_workspace_bytes(int B, int N, int k) {
    const size_t E = (size_t)N * k;
    if (N > 2048 || E > 65535) return 0;
    return (size_t)B * rev2_dims(N, k).bytes_per_cloud;
}

// reverse graph of every cloud into ws (depends on idx only: may run ahead of the gradient, on another stream)
int edge_bwd_fast_build(const int64_t* idx, int B, int N, int k, void* ws, cudaStream_t st) {
    const Rev2Dims d = rev2_dims(N, k);
    const size_t smem = (size_t)d.NS * d.W * (sizeof(unsigned) + sizeof(uint16_t)) + (size_t)(4 * d.NS + 3 * d.GS + 1) * sizeof(int);
    cudaFuncSetAttribute(edge_rev2_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    edge_rev2_build_kernel<<<dim3(d.S, B), kRevThreads, smem, st>>>(idx, N, k, 0xFFFFFFFFu / (unsigned)k + 1u, ws);
    return check_launch("edge_rev2_build_kernel");
}

int edge_bwd_fast_gather(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, const void* ws, cudaStream_t st) {
    const Rev2Dims d = rev2_dims(N, k);
    const int nbuf = gather_buffers(N, k);
    const GatherSmem g = gather_smem(N, k, d, kSmemBudget, nbuf);
    const int items = B * 3 * C;
    int grid = sm_count();
    if (grid > items) grid = items;
    const int per_cta = (items + grid - 1) / grid;
    grid = (items + per_cta - 1) / per_cta;
    long long* prof = nullptr;
#ifdef HPCS_BWD_PROFILE
    if (const char* env = getenv("HPCS_BWD_PROF_PTR")) prof = reinterpret_cast<long long*>(strtoull(env, nullptr, 0));
#endif
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.total);
        kern<<<grid, kGatherThreads, g.total, st>>>(gout, idx, ws, B, C, N, k, per_cta, kSmemBudget, prof, gx);
    };
    if (nbuf == 2) launch(edge_bwd_gather_kernel<2>); else launch(edge_bwd_gather_kernel<1>);
    return check_launch("edge_bwd_gather_kernel");
}

int edge_bwd_fast_run(const float* gout, const int64_t* idx, int B, int C, int N, int k, float* gx, void* ws, cudaStream_t st) {
    const int rc = edge_bwd_fast_build(idx, B, N, k, ws, st);
    return rc ? rc : edge_bwd_fast_gather(gout, idx, B, C, N, k, gx, ws, st);
}

}  // namespace hpcs
