// Decoding leaf embeddings into a binary dendrogram (scipy linkage format) on the GPU.
//
// Replaces, per cloud, the D2H copy + scipy.cluster.hierarchy.linkage(leaves, method, 'cosine')
// of BaseSimilarityHypHC._decode_linkage (hpcs/models/base_hyp_hc.py:81-86) and the Python loop
// over clouds at :135-137: all B clouds are decoded by one launch sequence, one CTA per cloud.
//   1. pdist_norms_kernel + pdist_mma_kernel  -- fp64 cosine distance matrix, bit-identical to scipy's pdist
//      (two running sums over even/odd elements, see oracle), on the fp64 tensor cores (mma.sync m8n8k4);
//   2. linkage_kernel<0,.>  -- 'single': Prim's MST from node 0 over matrix rows (what scipy's
//      mst_single_linkage does); a thread keeps the running minima of its columns in registers, loads
//      its part of the row in one batch and the block takes a (value, index) arg-min with redux.sync
//      and ONE barrier per step -- the N-1 steps are a serial chain, so a step's latency is the cost;
//      linkage_kernel<1,.>  -- 'complete': nearest-neighbour chain with the max update
//      (scipy's nn_chain), same arg-min primitive, distance matrix updated in place;
//   3. in the same kernel: stable bitonic sort of the N-1 merges by height, then the union-find
//      relabel pass (smaller root id first, new id N+i, subtree size) that scipy's `label` does.
// Because all leaves share one radius, hyperbolic-LCA similarity is a decreasing function of the
// angle and cosine distance an increasing one, so single linkage over cosine distance yields the
// same merge order as single linkage over hyperbolic similarity (SURVEY.md Finding 2).
#include <float.h>

#include <stdlib.h>

#include "common.cuh"

namespace hpcs {

constexpr int kPT = 64;           // pdist tile edge
constexpr int kPLD = kPT + 2;     // shared-memory row pitch in doubles (keeps 16-byte alignment, spreads banks)

// ---- distance matrix ------------------------------------------------------------------------------------------------
// fp64 cosine distances, bit-identical to scipy's pdist(., 'cosine') on the same fp32 rows: dm[b][i][j] full symmetric,
// zero diagonal.  scipy sums a dot product as two running sums (even / odd elements, separate multiply and add), added
// at the end, an odd tail element last; norms the same way; then 1 - dot / (n_i n_j).  Every operand here is an fp32
// value widened to fp64, so a product has at most 48 significant bits and is exact in fp64: fma(a, b, acc) rounds the
// same real number as add(mul(a, b), acc) and gives the same bits.
__global__ void pdist_norms_kernel(const float* __restrict__ leaves, size_t rows, int D, double* __restrict__ norms) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const float* v = leaves + i * D;
    double even = 0.0, odd = 0.0;
    int q = 0;
    for (; q + 1 < D; q += 2) {
        const double v0 = (double)__ldg(v + q), v1 = (double)__ldg(v + q + 1);
        even = fma(v0, v0, even);
        odd = fma(v1, v1, odd);
    }
    double s = __dadd_rn(even, odd);
    if (D & 1) { const double t = (double)__ldg(v + D - 1); s = fma(t, t, s); }
    norms[i] = __dsqrt_rn(s);
}

// ---- the distance kernel, on the fp64 tensor cores -----------------------------------------------------------------
// mma.sync.m8n8k4.f64 adds its four products as an FMA chain in ascending k (measured on B200 with
// tools/probes/dmma_order_probe.cu: 128000 of 128000 outputs equal the chain bit for bit, 95.7 % equal a descending
// chain or a single rounding of the exact sum).  scipy's dot product is two such chains -- over the even and over the
// odd feature indices -- so feeding one accumulator the k-slices (0,2,4,6), (8,10,12,14), ... and another the slices
// (1,3,5,7), ... reproduces both running sums exactly (products of fp32-origin operands are exact in fp64, so FMA and
// multiply-then-add round the same number; zero padding adds +0 and changes nothing).  What the tensor cores buy is not
// flops (64 fp64 FMA lanes per SM per clock either way) but operand delivery: a thread loads ONE double per operand
// fragment for 8 FMAs, 2 B per FMA against 4 B at a 4x4 register tile, and issues one instruction per 256 FMAs -- the
// DFMA kernel sat at 61 % of the shared-memory pipe with the fp64 pipe 28 % busy.
// Earlier forms, measured at B=64, N=1024 / 8192: one 64x64 tile per CTA on DFMA, 4x4 outputs per thread: 333 us /
// 20.0 ms (an 8x4 micro-tile: 399 us / 23.0 ms -- 178 registers, half the warps); the same with the schedule below:
// no faster (the staging was not the limit); this kernel: 227 us / 10.6 ms.
// Schedule: CTA = (64-row block bi, cloud); it stages the A tile and its norms once and walks the column blocks
// bj = bi .. T-1 itself, the next B tile fetched from global memory into registers (PF floats per thread, 64 D <= 256 PF;
// PF = 0: any D, no prefetch) while the current one is computed; norms come from pdist_norms_kernel; the runtime
// division by D in the staging index is a multiply-high.  Warp w: rows 16 (w >> 1) .. +15, columns 32 (w & 1) .. +31 of the 64x64 tile = 2 x 4 MMA tiles, two
// accumulator fragments (even / odd chain) each.  Fragment coordinates: g = lane >> 2, t = lane & 3; A[row g][k t],
// B[k t][col g], C[row g][cols 2t, 2t+1].  Shared memory holds the tiles feature-major with the features of a slice
// step interleaved as the fragments want them; rows [2P, 8S) stay zero, the odd tail feature lives in row 8S.
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int PF>
__global__ void __launch_bounds__(256, 2)
pdist_mma_kernel(const float* __restrict__ leaves, const double* __restrict__ norms, int N, int D, unsigned d_magic,
                 double* __restrict__ dm, const int* __restrict__ gate) {
    extern __shared__ __align__(16) double sm[];
    if (gate && gate[blockIdx.y] == 0) return;           // exact redo of flagged clouds only (complete linkage, tied distances)
    const int P = D >> 1, S = (P + 3) >> 2;      // feature pairs; slice steps of 4 pairs
    const int rows = 8 * S + 1;                   // staged feature rows (+ the tail row)
    double* as = sm;                              // [rows][kPLD]
    double* bs = as + (size_t)rows * kPLD;        // [rows][kPLD]
    double* na = bs + (size_t)rows * kPLD;        // [64]
    double* nb = na + kPT;                        // [64]
    const int b = blockIdx.y, bi = blockIdx.x;
    const int T = (N + kPT - 1) / kPT;
    const float* lb = leaves + (size_t)b * N * D;
    const double* nrm = norms + (size_t)b * N;
    double* db = dm + (size_t)b * N * N;
    const int tile_elems = kPT * D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = 16 * (warp >> 1), c0 = 32 * (warp & 1);
    const bool vec_ok = (N & 1) == 0;

    auto row_of = [&](int q) { return q < 2 * P ? q : 8 * S; };     // the odd tail feature goes to the last row
    auto stage = [&](double* dst, double* ndst, int blk) {
        for (int e = threadIdx.x; e < tile_elems; e += 256) {
            const int r = D == 1 ? e : (int)__umulhi((unsigned)e, d_magic), q = e - r * D;     // (the magic number wraps to 0 for D = 1)
            const int gr = blk * kPT + r;
            dst[row_of(q) * kPLD + r] = gr < N ? (double)__ldg(lb + (size_t)gr * D + q) : 0.0;
        }
        if (threadIdx.x < kPT) { const int gr = blk * kPT + threadIdx.x; ndst[threadIdx.x] = gr < N ? nrm[gr] : 0.0; }
    };
    float pre[PF > 0 ? PF : 1];
    double pre_n = 0.0;
    auto prefetch = [&](int blk) {
        if (PF > 0) {
            const float* src = lb + (size_t)blk * kPT * D;
            const int left = (N - blk * kPT) * D;
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                const int e = threadIdx.x + 256 * i;
                pre[i] = (e < tile_elems && e < left) ? __ldg(src + e) : 0.f;
            }
            if (threadIdx.x < kPT) { const int gr = blk * kPT + threadIdx.x; pre_n = gr < N ? nrm[gr] : 0.0; }
        }
    };
    auto commit = [&]() {
        if (PF > 0) {
#pragma unroll
            for (int i = 0; i < PF; ++i) {
                const int e = threadIdx.x + 256 * i;
                if (e < tile_elems) {
                    const int r = D == 1 ? e : (int)__umulhi((unsigned)e, d_magic), q = e - r * D;     // (the magic number wraps to 0 for D = 1)
                    bs[row_of(q) * kPLD + r] = (double)pre[i];
                }
            }
            if (threadIdx.x < kPT) nb[threadIdx.x] = pre_n;
        }
    };

    for (int e = threadIdx.x; e < (8 * S - 2 * P) * kPLD; e += 256) {       // zero padding rows of both tiles, once
        as[2 * P * kPLD + e] = 0.0;
        bs[2 * P * kPLD + e] = 0.0;
    }
    if (!(D & 1)) for (int e = threadIdx.x; e < kPLD; e += 256) { as[8 * S * kPLD + e] = 0.0; bs[8 * S * kPLD + e] = 0.0; }
    stage(as, na, bi);
    if (PF > 0) { prefetch(bi); commit(); } else stage(bs, nb, bi);
    __syncthreads();

    const double* ap = as + (2 * t) * kPLD + r0 + g;                  // fragment element of slice step 0, even chain
    const double* bp = bs + (2 * t) * kPLD + c0 + g;
    for (int bj = bi; bj < T; ++bj) {
        if (bj + 1 < T) prefetch(bj + 1);
        double ev[2][4][2], od[2][4][2];
#pragma unroll
        for (int rt = 0; rt < 2; ++rt)
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) { ev[rt][ct][0] = ev[rt][ct][1] = 0.0; od[rt][ct][0] = od[rt][ct][1] = 0.0; }
        for (int s = 0; s < S; ++s) {
            double ae[2], ao[2], be[4], bo[4];
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) { ae[rt] = ap[(8 * s) * kPLD + 8 * rt]; ao[rt] = ap[(8 * s + 1) * kPLD + 8 * rt]; }
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) { be[ct] = bp[(8 * s) * kPLD + 8 * ct]; bo[ct] = bp[(8 * s + 1) * kPLD + 8 * ct]; }
#pragma unroll
            for (int rt = 0; rt < 2; ++rt)
#pragma unroll
                for (int ct = 0; ct < 4; ++ct) {
                    dmma_m8n8k4(ev[rt][ct][0], ev[rt][ct][1], ae[rt], be[ct]);
                    dmma_m8n8k4(od[rt][ct][0], od[rt][ct][1], ao[rt], bo[ct]);
                }
        }
#pragma unroll
        for (int rt = 0; rt < 2; ++rt) {
            const int i = r0 + 8 * rt + g, gi = bi * kPT + i;
            const double ni = na[i], ti = as[8 * S * kPLD + i];
#pragma unroll
            for (int ct = 0; ct < 4; ++ct) {
                const int j = c0 + 8 * ct + 2 * t, gj = bj * kPT + j;
                double res[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double sum = __dadd_rn(ev[rt][ct][e], od[rt][ct][e]);
                    if (D & 1) sum = fma(ti, bs[8 * S * kPLD + j + e], sum);
                    double cs = __ddiv_rn(sum, __dmul_rn(ni, nb[j + e]));
                    if (fabs(cs) > 1.0) cs = copysign(1.0, cs);
                    res[e] = gi == gj + e ? 0.0 : __dsub_rn(1.0, cs);
                }
                if (gi < N) {
                    double* dst = db + (size_t)gi * N + gj;
                    if (vec_ok && gj + 1 < N) *reinterpret_cast<double2*>(dst) = make_double2(res[0], res[1]);
                    else { if (gj < N) dst[0] = res[0]; if (gj + 1 < N) dst[1] = res[1]; }
                    if (bi != bj) {                                  // mirrored tile: 8 lanes (g) x 8 B contiguous per column
                        if (gj < N) db[(size_t)gj * N + gi] = res[0];
                        if (gj + 1 < N) db[(size_t)(gj + 1) * N + gi] = res[1];
                    }
                }
            }
        }
        __syncthreads();                                            // tile consumed
        if (bj + 1 < T) { if (PF > 0) commit(); else stage(bs, nb, bj + 1); }
        __syncthreads();
    }
}

__device__ __forceinline__ bool am_less(double v, int i, double ov, int oi) { return v < ov || (v == ov && i < oi); }

// ---- block-wide lexicographic (distance, index) minimum ------------------------------------------------------
// Distances are non-negative doubles (or +inf), so their bit patterns order like unsigned integers: the minimum is
// three integer warp reductions (redux.sync) -- high word, low word among the lanes that tie on the high word, index
// among the lanes that tie on both -- instead of five rounds of 64-bit shuffles.  One __syncthreads per call: the
// per-warp results go through a double-buffered scratch array and every warp reduces them again for itself.
struct LexKey {
    unsigned hi, lo;
    int idx;
};
constexpr int kNoIdx = 0x7fffffff;

__device__ __forceinline__ void warp_lexmin(LexKey& k) {
    const unsigned mh = __reduce_min_sync(kFull, k.hi);
    const unsigned l2 = k.hi == mh ? k.lo : 0xffffffffu;
    const unsigned ml = __reduce_min_sync(kFull, l2);
    const unsigned i2 = (k.hi == mh && k.lo == ml) ? (unsigned)k.idx : (unsigned)kNoIdx;
    k.idx = (int)__reduce_min_sync(kFull, i2);
    k.hi = mh;
    k.lo = ml;
}

__device__ __forceinline__ LexKey block_lexmin(double v, int idx, uint4* scratch /*[2][32]*/, unsigned& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    LexKey k{(unsigned)(bits >> 32), (unsigned)bits, idx};
    warp_lexmin(k);
    uint4* buf = scratch + (phase & 1u) * 32;
    ++phase;
    if (lane == 0) buf[warp] = make_uint4(k.hi, k.lo, (unsigned)k.idx, 0u);
    __syncthreads();
    const uint4 e = lane < nwarp ? buf[lane] : make_uint4(0xffffffffu, 0xffffffffu, (unsigned)kNoIdx, 0u);
    LexKey r{e.x, e.y, (int)e.z};
    warp_lexmin(r);
    return r;
}
__device__ __forceinline__ double key_value(const LexKey& k) {
    return __longlong_as_double((long long)(((unsigned long long)k.hi << 32) | k.lo));
}

// METHOD 0 single, 1 complete.  One CTA per cloud; thread t owns columns t, t + T, ..., t + (CPT-1) T of the
// distance matrix: its running minima (single) and its alive flags live in registers, and the CPT loads of a row are
// issued back to back before any of them is used.
// Contracted input (single linkage after Boruvka rounds, see below): the matrix is n x n with row pitch `pitch`,
// n is read from nd_all[b*8 + nd_slot], node v stands for leaf rep[v], and the first N - n merge records are already
// in recx/recy/rech.  Direct input: pitch = N, nd_all = nullptr, rep_all = nullptr.
// gate (optional): per-cloud flag; the CTA returns at once unless gate[b] != 0 (exact redo of clouds with tied heights).
template <int METHOD, int CPT>
__global__ void __launch_bounds__(1024)
linkage_kernel(double* dm_all, size_t dm_stride, int pitch, const int* __restrict__ nd_all, int nd_slot,
               const int* __restrict__ rep_all, int rep_stride, int N, int NP2, int* __restrict__ recx_all,
               int* __restrict__ recy_all, double* __restrict__ rech_all, double* __restrict__ Z_all,
               int* __restrict__ tie_flag, const int* __restrict__ gate, const int* __restrict__ alive_in,
               const int* __restrict__ rec0_in) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ uint4 scratch[64];
    __shared__ int tie_s;
    const int b = blockIdx.x;
    if (gate && gate[b] == 0) return;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int M = N - 1;
    const int n_nodes = nd_all ? nd_all[b * 8 + nd_slot] : N;       // nodes of the (contracted) matrix
    // merges already recorded: single linkage after Boruvka rounds (contracted matrix), or complete linkage after the
    // parallel reciprocal-nearest-neighbour rounds (same matrix, alive_in marks the clusters that are left)
    const int rec0 = rec0_in ? rec0_in[b] : N - n_nodes;
    const int* rep = rep_all ? rep_all + (size_t)b * rep_stride : nullptr;
    double* dm = dm_all + (size_t)b * dm_stride;          // mutated by METHOD 1: plain loads only
    int* recx = recx_all + (size_t)b * N;
    int* recy = recy_all + (size_t)b * N;
    double* rech = rech_all + (size_t)b * N;
    double* Z = Z_all + (size_t)b * M * 4;

    // shared-memory regions
    double* A = reinterpret_cast<double*>(raw);                      // [NP2]  sort keys
    int* ordv = reinterpret_cast<int*>(raw + (size_t)8 * NP2);       // [NP2]  sort payload
    const size_t sort_bytes = ((size_t)12 * (NP2 > N ? NP2 : N) + 15) / 16 * 16;   // >= 12N so that node[N] (16N bytes) ends before usize
    int* csize = reinterpret_cast<int*>(raw + sort_bytes);           // [2N] relabel scratch, directly behind the sort arrays
    unsigned char* regC = raw + sort_bytes + (size_t)8 * N;          // 8N bytes: NN chain, later the sorted (x,y)

    unsigned alive = 0u, phase = 0u;
#pragma unroll
    for (int c = 0; c < CPT; ++c)
        if (tid + c * nthr < n_nodes && (!alive_in || alive_in[(size_t)b * N + tid + c * nthr])) alive |= 1u << c;
    auto drop = [&](int col) {                                       // the owner of `col` clears its flag
        const int c = col / nthr;
        if (col - c * nthr == tid) alive &= ~(1u << c);
    };
    auto first_alive = [&]() -> int {                                // lowest column still alive (block-wide)
        int mine = kNoIdx;
#pragma unroll
        for (int c = CPT - 1; c >= 0; --c) if ((alive >> c) & 1u) mine = tid + c * nthr;
        return block_lexmin(0.0, mine, scratch, phase).idx;
    };

    if (METHOD == 0) {
        // Prim from node 0 over matrix rows, what scipy's mst_single_linkage does: (distance, index) arg-min per step
        double dmin[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) dmin[c] = INFINITY;
        int x = 0;
        for (int k = rec0; k < M; ++k) {
            drop(x);
            const double* row = dm + (size_t)x * pitch;
            double v[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = tid + c * nthr;
                v[c] = col < n_nodes ? __ldcs(row + col) : INFINITY; // each row is read once: streaming
            }
            double bv = INFINITY;
            int bi = kNoIdx;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                if ((alive >> c) & 1u) {
                    if (dmin[c] > v[c]) dmin[c] = v[c];
                    if (dmin[c] < bv) { bv = dmin[c]; bi = tid + c * nthr; }
                }
            }
            LexKey r = block_lexmin(bv, bi, scratch, phase);
            if (r.idx == kNoIdx) r.idx = first_alive();              // only NaN/inf distances left (zero-norm rows)
            if (tid == 0) {
                recx[k] = rep ? rep[x] : x;
                recy[k] = rep ? rep[r.idx] : r.idx;
                rech[k] = key_value(r);
            }
            x = r.idx;
        }
    } else {
        // nearest-neighbour chain with the 'complete' update (scipy's nn_chain), matrix updated in place
        int* chain = reinterpret_cast<int*>(regC);                    // [N]
        int len = 0, x = 0, prev = -1;
        for (int k = rec0_in ? rec0 : 0; k < M; ++k) {
            if (len == 0) {
                x = first_alive();
                prev = -1;
                if (tid == 0) chain[0] = x;
                len = 1;
            }
            int y;
            double cur;
            while (true) {
                const double* row = dm + (size_t)x * N;
                double v[CPT];
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int col = tid + c * nthr;
                    v[c] = col < N ? row[col] : INFINITY;
                }
                const double dprev = prev >= 0 ? row[prev] : INFINITY;
                double bv = INFINITY;
                int bi = kNoIdx;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int col = tid + c * nthr;
                    if (((alive >> c) & 1u) && col != x && v[c] < bv) { bv = v[c]; bi = col; }
                }
                const LexKey r = block_lexmin(bv, bi, scratch, phase);
                y = r.idx; cur = key_value(r);
                bool done = false;
                if (prev >= 0 && !(cur < dprev)) { y = prev; cur = dprev; done = true; }   // ties prefer the previous link
                if (y == kNoIdx) { y = first_alive(); }               // degenerate input (NaN rows): keep going
                if (done) break;
                if (tid == 0) chain[len] = y;
                ++len;
                prev = x;
                x = y;
            }
            if (x > y) { const int t = x; x = y; y = t; }
            if (tid == 0) { recx[k] = x; recy[k] = y; rech[k] = cur; }
            drop(x);                                                  // cluster x is dropped, y becomes the union
            // Lance-Williams 'complete': d(i, x u y) = max(d(i,x), d(i,y)); kept symmetric
            const double* rowx = dm + (size_t)x * N;
            double* rowy = dm + (size_t)y * N;
            double vx[CPT], vy[CPT];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = tid + c * nthr;
                const bool on = ((alive >> c) & 1u) && col != y;
                vx[c] = on ? rowx[col] : 0.0;
                vy[c] = on ? rowy[col] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = tid + c * nthr;
                if (((alive >> c) & 1u) && col != y) {
                    const double v = fmax(vx[c], vy[c]);
                    rowy[col] = v;
                    dm[(size_t)col * N + y] = v;
                }
            }
            __syncthreads();                       // column y (written by other threads) is read by later row scans
            len -= 2;
            if (len > 0) { x = chain[len - 1]; prev = len > 1 ? chain[len - 2] : -1; }
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
    if (tid == 0) tie_s = 0;
    __syncthreads();
    for (int i = tid; i < M; i += nthr) {
        const int o = ordv[i];
        exy[i] = make_int2(recx[o], recy[o]);
        Z[(size_t)i * 4 + 2] = A[i];
        if (i + 1 < M && A[i] == A[i + 1]) tie_s = 1;                // equal heights: merge order is not unique
    }
    __syncthreads();
    if (tie_flag && tid == 0) tie_flag[b] = (rec0_in ? tie_flag[b] : 0) | tie_s;   // (the parallel rounds may have flagged it already)
    // ---- union-find relabel (scipy `label`): ids of the two clusters a merge joins, size of the union --------------
    // The forest lives on the N leaves with union by size, one packed record per leaf {parent, size, cluster id} so
    // that a find that lands on a root has everything after ONE 16-byte load; the cluster id is the leaf id, or N + i
    // for the cluster made by sorted merge i.  Single-linkage dendrograms are chains -- one big cluster swallowing
    // points -- and union by size keeps the big cluster's root fixed: a find is one or two hops.  The loop is a serial
    // dependency chain, so it runs on one thread with as few instructions as possible; results are staged in shared
    // memory (ids over the consumed exy records) and written out by the whole CTA afterwards.
    int4* node = reinterpret_cast<int4*>(raw);                        // [N] overlays A, ordv and the head of csize
    int* usize = csize + N;                                           // [N] size of the union made by merge i (csize is [2N])
    for (int v = tid; v < N; v += nthr) node[v] = make_int4(v, 1, v, 0);
    __syncthreads();
    if (tid == 0) {
        int2 e = exy[0];
        for (int i = 0; i < M; ++i) {
            const int2 nxt = exy[i + 1 < M ? i + 1 : i];              // next record: independent of the chase below
            int rx = e.x, ry = e.y;
            int4 nx = node[rx], ny = node[ry];
            while (nx.x != rx) {                                      // path halving; stops one hop early at a root parent
                const int p = nx.x;
                const int4 np = node[p];
                if (np.x == p) { rx = p; nx = np; break; }
                node[rx].x = np.x;
                rx = np.x;
                nx = node[rx];
            }
            while (ny.x != ry) {
                const int p = ny.x;
                const int4 np = node[p];
                if (np.x == p) { ry = p; ny = np; break; }
                node[ry].x = np.x;
                ry = np.x;
                ny = node[ry];
            }
            const bool xbig = nx.y >= ny.y;
            const int big = xbig ? rx : ry, small = xbig ? ry : rx;
            node[small].x = big;
            node[big] = make_int4(big, nx.y + ny.y, N + i, 0);
            exy[i] = make_int2(min(nx.z, ny.z), max(nx.z, ny.z));
            usize[i] = nx.y + ny.y;
            e = nxt;
        }
    }
    __syncthreads();
    for (int i = tid; i < M; i += nthr) {
        const int2 ids = exy[i];
        Z[(size_t)i * 4 + 0] = (double)ids.x;
        Z[(size_t)i * 4 + 1] = (double)ids.y;
        Z[(size_t)i * 4 + 3] = (double)usize[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Boruvka rounds for single linkage
// ------------------------------------------------------------------------------------------------
// Prim's N - 1 steps are a serial chain of row fetches; a Boruvka round is fully parallel: every node takes its
// nearest other node ((distance, index) arg-min of its matrix row), these edges are MST edges, the nodes they join
// become one super-node, and the matrix is contracted to component-to-component minima.  A round leaves at most half
// the nodes (about a third on embeddings), so after R rounds Prim runs on N/3^R nodes.  The dendrogram of single
// linkage is the same for every MST as long as the N - 1 merge heights are distinct (clusters at a threshold are the
// connected components of the threshold graph); when two heights tie, scipy's order depends on its Prim visiting
// order, so such a cloud raises tie_flag and is redone by the exact Prim emulation on the untouched full matrix.
//
// Per cloud: nd[8] ints: nd[r] = nodes after r rounds (nd[0] = N).

__device__ __forceinline__ int node_count(const int* nd_all, int b, int slot, int N) {
    return slot == 0 ? N : nd_all[b * 8 + slot];
}

// Row arg-min of the n x n matrix (n = nd[slot]).  grid (rows, B), one CTA per row.
__global__ void __launch_bounds__(256)
boruvka_rowmin_kernel(const double* __restrict__ dm_all, size_t dm_stride, int pitch, const int* __restrict__ nd_all,
                      int slot, double* __restrict__ rmw_all, int* __restrict__ rmj_all, int N) {
    __shared__ uint4 scratch[64];
    const int b = blockIdx.y, i = blockIdx.x;
    const int n = node_count(nd_all, b, slot, N);
    if (i >= n) return;
    const double* row = dm_all + (size_t)b * dm_stride + (size_t)i * pitch;
    double bv = INFINITY;
    int bi = kNoIdx;
    for (int j0 = threadIdx.x; j0 < n; j0 += 4 * blockDim.x) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int j = j0 + u * blockDim.x; v[u] = j < n ? __ldcs(row + j) : INFINITY; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int j = j0 + u * blockDim.x; if (j != i && v[u] < bv) { bv = v[u]; bi = j; } }
    }
    unsigned phase = 0u;
    const LexKey r = block_lexmin(bv, bi, scratch, phase);
    if (threadIdx.x == 0) { rmw_all[(size_t)b * N + i] = key_value(r); rmj_all[(size_t)b * N + i] = r.idx; }
}

// exclusive prefix sum over vals[0..n) (ints in shared memory), in place; returns the total.  Every thread calls it.
__device__ int block_exclusive_scan(int* vals, int n, int* wsum /*[32]*/) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const int per = (n + nthr - 1) / nthr;
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += vals[i];
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += v; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < nwarp ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, winc, o); if (lane >= o) winc += v; }
        wsum[lane] = winc - w;                                   // exclusive warp offsets; lane 31's inclusive = total
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    int run = wsum[warp] + inc - sum;
    for (int i = lo; i < hi; ++i) { const int v = vals[i]; vals[i] = run; run += v; }
    const int total = wsum[32];
    __syncthreads();
    return total;
}

// Hook + label one round.  One CTA per cloud.  In: row arg-mins (rmw, rmj) of the n = nd[slot] nodes, rep_in (leaf
// standing for a node; nullptr = identity).  Out: merge records appended at rec[N - n ...], nd[slot + 1] = n',
// label-sorted member lists (moff[n' + 1], memb[n]) for the contraction, rep_out[n'].
__global__ void __launch_bounds__(1024)
boruvka_hook_kernel(const double* __restrict__ rmw_all, const int* __restrict__ rmj_all, int* __restrict__ nd_all, int slot,
                    const int* __restrict__ rep_in_all, int rep_in_stride, int* __restrict__ rep_out_all, int rep_out_stride,
                    int* __restrict__ memb_all, int* __restrict__ moff_all, int* __restrict__ recx_all,
                    int* __restrict__ recy_all, double* __restrict__ rech_all, int N) {
    extern __shared__ __align__(16) int hs[];
    __shared__ int wsum[33];
    const int b = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    int* nd = nd_all + b * 8;
    const int n = node_count(nd_all, b, slot, N);
    if (n <= 1) { if (tid == 0) nd[slot + 1] = n; return; }       // nothing left to join (moff/memb/rep_out unused)
    int* par = hs;              // [n] nearest node, then root
    int* aux = par + n;         // [n] emit flags -> record positions; then root flags -> new ids
    int* cnt = aux + n;         // [n] member counts -> offsets -> cursors
    const double* rmw = rmw_all + (size_t)b * N;
    const int* rmj = rmj_all + (size_t)b * N;
    const int* rep_in = rep_in_all ? rep_in_all + (size_t)b * rep_in_stride : nullptr;
    int* rep_out = rep_out_all + (size_t)b * rep_out_stride;
    int* memb = memb_all + (size_t)b * N;
    int* moff = moff_all + (size_t)b * (N + 1);
    int* recx = recx_all + (size_t)b * N;
    int* recy = recy_all + (size_t)b * N;
    double* rech = rech_all + (size_t)b * N;

    for (int i = tid; i < n; i += nthr) {
        const int j = rmj[i];
        par[i] = (unsigned)j < (unsigned)n ? j : (i + 1 < n ? i + 1 : 0);   // no finite distance in the row (NaN input): any other node
    }
    __syncthreads();
    // a mutual pair keeps its lower node as root; every other node emits its edge
    for (int i = tid; i < n; i += nthr) { const int j = par[i]; aux[i] = (par[j] == i && i < j) ? 0 : 1; }
    __syncthreads();
    const int rec0 = N - n;
    {
        // positions of the emitted records (scan of the emit flags), written before par is modified
        for (int i = tid; i < n; i += nthr) cnt[i] = aux[i];
        __syncthreads();
        block_exclusive_scan(cnt, n, wsum);
        for (int i = tid; i < n; i += nthr) {
            if (aux[i]) {
                const int j = par[i], pos = rec0 + cnt[i];
                recx[pos] = rep_in ? rep_in[i] : i;
                recy[pos] = rep_in ? rep_in[j] : j;
                rech[pos] = rmw[i];
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) if (!aux[i]) par[i] = i;
    __syncthreads();
    // pointer jumping to the roots (in place: a value read mid-update is still an ancestor)
    for (int it = 0; it < 14; ++it) {
        int changed = 0;
        for (int i = tid; i < n; i += nthr) {
            const int p = par[i], g = par[p];
            if (g != p) changed = 1;
            par[i] = g;
        }
        if (!__syncthreads_or(changed)) break;
    }
    // new ids of the roots, ascending
    for (int i = tid; i < n; i += nthr) { aux[i] = par[i] == i ? 1 : 0; cnt[i] = 0; }
    __syncthreads();
    const int n2 = block_exclusive_scan(aux, n, wsum);
    for (int i = tid; i < n; i += nthr) {
        const int c = aux[par[i]];
        atomicAdd(&cnt[c], 1);
        if (par[i] == i) rep_out[c] = rep_in ? rep_in[i] : i;
    }
    __syncthreads();
    block_exclusive_scan(cnt, n2, wsum);
    for (int c = tid; c < n2; c += nthr) moff[c] = cnt[c];
    if (tid == 0) { moff[n2] = n; nd[slot + 1] = n2; }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) memb[atomicAdd(&cnt[aux[par[i]]], 1)] = i;     // order inside a component is free: only minima follow
}

// Contract: out[A][C] = min over members i of A, j of C of in[i][j]; plus the row arg-min of out for the next
// round.  grid (rows of out, B), one CTA per component A; the diagonal of out is never read.
__global__ void __launch_bounds__(256)
boruvka_contract_kernel(const double* __restrict__ in_all, size_t in_stride, int in_pitch, double* __restrict__ out_all,
                        size_t out_stride, int out_pitch, const int* __restrict__ nd_all, int slot,
                        const int* __restrict__ memb_all, const int* __restrict__ moff_all, double* __restrict__ rmw_all,
                        int* __restrict__ rmj_all, int N) {
    extern __shared__ __align__(16) double colmin[];               // [n]
    __shared__ uint4 scratch[64];
    const int b = blockIdx.y, A = blockIdx.x;
    const int n = node_count(nd_all, b, slot, N), n2 = nd_all[b * 8 + slot + 1];
    if (n <= 1 || A >= n2) return;
    const double* in = in_all + (size_t)b * in_stride;
    double* out = out_all + (size_t)b * out_stride + (size_t)A * out_pitch;
    const int* memb = memb_all + (size_t)b * N;
    const int* moff = moff_all + (size_t)b * (N + 1);
    const int m0 = moff[A], m1 = moff[A + 1];
    for (int j0 = threadIdx.x; j0 < n; j0 += 2 * blockDim.x) {
        double acc[2] = {INFINITY, INFINITY};
        for (int t = m0; t < m1; t += 4) {                             // 4 member rows x 2 columns in flight per thread
            double v[4][2];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const double* row = in + (size_t)memb[min(t + w, m1 - 1)] * in_pitch;
#pragma unroll
                for (int u = 0; u < 2; ++u) { const int j = j0 + u * blockDim.x; v[w][u] = j < n ? __ldcs(row + j) : INFINITY; }
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int u = 0; u < 2; ++u) acc[u] = fmin(acc[u], v[w][u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) { const int j = j0 + u * blockDim.x; if (j < n) colmin[j] = acc[u]; }
    }
    __syncthreads();
    double bv = INFINITY;
    int bi = kNoIdx;
    // Four components per thread at a time, their list bounds loaded together, then the first four members of each
    // (components have about three): two dependent L2 latencies per batch instead of two per component.
    // (Only worth it when a thread has several components: measured 61.7 -> 69.7 us per launch at N = 1024, but
    // 4.2 -> 3.7 ms at N = 8192.)
    const bool batched = n2 > 4 * (int)blockDim.x;
    for (int C = threadIdx.x; !batched && C < n2; C += blockDim.x) {
        double m = INFINITY;
        for (int t = moff[C]; t < moff[C + 1]; ++t) m = fmin(m, colmin[memb[t]]);
        out[C] = m;
        if (C != A && m < bv) { bv = m; bi = C; }
    }
    for (int C0 = threadIdx.x; batched && C0 < n2; C0 += 4 * blockDim.x) {
        int lo[4], hi[4], mem[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int C = C0 + u * blockDim.x;
            lo[u] = C < n2 ? moff[C] : 0;
            hi[u] = C < n2 ? moff[C + 1] : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) mem[u][w] = lo[u] < hi[u] ? memb[min(lo[u] + w, hi[u] - 1)] : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int C = C0 + u * blockDim.x;
            if (C >= n2) continue;
            double m = INFINITY;
#pragma unroll
            for (int w = 0; w < 4; ++w) m = fmin(m, colmin[mem[u][w]]);
            for (int t = lo[u] + 4; t < hi[u]; ++t) m = fmin(m, colmin[memb[t]]);
            out[C] = m;
            if (C != A && m < bv) { bv = m; bi = C; }
        }
    }
    unsigned phase = 0u;
    const LexKey r = block_lexmin(bv, bi, scratch, phase);
    if (threadIdx.x == 0) { rmw_all[(size_t)b * N + A] = key_value(r); rmj_all[(size_t)b * N + A] = r.idx; }
}

// ------------------------------------------------------------------------------------------------
// Complete linkage, parallel: rounds of reciprocal nearest neighbours.
// ------------------------------------------------------------------------------------------------
// scipy's nn_chain (what linkage(method='complete') runs, and what linkage_kernel<1> walks step by step) is ~4N dependent
// row passes.  Complete linkage is reducible: if A and B are each other's nearest neighbours they are merged by the greedy
// algorithm no matter what happens elsewhere, so ALL reciprocal pairs of a round can merge at once; the Lance-Williams
// update d(C, A u B) = max(d(C,A), d(C,B)) only ever grows entries, so a cluster whose nearest neighbour was not merged
// keeps it.  When no two candidate distances of a row tie exactly the nearest neighbours are unique, the hierarchy is the
// greedy one, and after the stable sort by height + union-find relabel (the tail of linkage_kernel) Z equals scipy's bit
// for bit.  A row minimum attained twice (duplicate points) flags the cloud; flagged clouds are redone by the serial kernel
// on a recomputed matrix, in the same call.
// One thread-block CLUSTER of 8 CTAs per cloud, the whole decode of a cloud in one launch: phases are separated by cluster
// barriers (release / acquire at cluster scope), shared mutable state is read with L2 loads.  A round:
//   1. row minima (lexicographic (distance, index)) of the rows whose nearest neighbour was invalidated -- warp per row;
//   2. reciprocal pairs (x < y): merge record (x, y, d), partner[] marks both ends;
//   3. row x <- max(row x, row y) (row x is dead after the round: it is the scratch that makes phase 4 race-free);
//   4. row y and column y <- the scratch; an entry between two clusters merged in this round takes the max over the partner's
//      scratch entries; rows whose nearest neighbour was x or y are marked for phase 1; x leaves the alive set.
// Rounds run until one cluster is left or a round finds no pair (NaN rows); linkage_kernel<1> then finishes whatever is
// left from the alive mask and does the sort + relabel.  About a third of the clusters merge per round on embeddings.
constexpr int kClusterCtas = 8;
constexpr int kRoundThreads = 512;
constexpr int kRoundMaxN = 8192;          // alive / partner snapshots of a cloud live in shared memory
constexpr int kRoundMlp = 16;             // independent 64-bit loads a lane keeps in flight (the phases are L2-latency chains otherwise)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

// stop_frac: the rounds stop after a round that merged fewer than stop_frac pairs (a round costs four cluster barriers and a few L2
// latencies, ~10 us; the serial kernel ~5 us per merge)
template <int CS>                                      // CTAs per cluster: 8, or 16 (non-portable size) when few clouds are decoded
__global__ void __launch_bounds__(kRoundThreads)
complete_rounds_kernel(double* dm_all, int N, int stop_frac, int* __restrict__ alive_all, int* __restrict__ nn_all,
                       double* __restrict__ nnd_all, int* __restrict__ partner_all, int* __restrict__ rescan_all,
                       int2* __restrict__ pairs_all, int* __restrict__ cnt_all /*[B][16]: pair counters of even / odd rounds, records, rounds, 6 phase timers (us)*/,
                       int* __restrict__ recx_all, int* __restrict__ recy_all, double* __restrict__ rech_all, int* __restrict__ tie_all) {
    __shared__ unsigned char alive_s[kRoundMaxN];      // bit 0: alive, bit 1: needs a new nearest neighbour
    __shared__ int partner_s[kRoundMaxN];
    const int b = blockIdx.y;
    const int rank = (int)cluster_cta_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = CS * (kRoundThreads / 32), NT = CS * kRoundThreads;
    const int gw = rank * (kRoundThreads / 32) + warp, gt = rank * kRoundThreads + threadIdx.x;
    double* dm = dm_all + (size_t)b * N * N;
    int* alive = alive_all + (size_t)b * N;
    int* nn = nn_all + (size_t)b * N;
    double* nnd = nnd_all + (size_t)b * N;
    int* partner = partner_all + (size_t)b * N;
    int* rescan = rescan_all + (size_t)b * N;
    int2* pairs = pairs_all + (size_t)b * (N / 2 + 1);
    int* cnt = cnt_all + b * 16;
    unsigned long long t_mark = 0, t_acc[6] = {0, 0, 0, 0, 0, 0};
    auto tick = [&](int slot) {
        if (gt == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (slot >= 0) t_acc[slot] += t - t_mark;
            t_mark = t;
        }
    };
    int* recx = recx_all + (size_t)b * N;
    int* recy = recy_all + (size_t)b * N;
    double* rech = rech_all + (size_t)b * N;
    for (int i = gt; i < N; i += NT) { alive[i] = 1; rescan[i] = 1; partner[i] = -1; }
    if (gt == 0) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; tie_all[b] = 0; }
    cluster_sync_all();
    int nrec = 0, n_alive = N, round = 0;
    tick(-1);
    for (; n_alive > 1; ++round) {
        for (int i = threadIdx.x; i < N; i += kRoundThreads)                    // this round's alive set and rescan marks
            alive_s[i] = (unsigned char)((__ldcg(alive + i) ? 1 : 0) | (__ldcg(rescan + i) ? 2 : 0));
        __syncthreads();
        tick(0);
        // ---- 1. nearest neighbour of every row that needs one: warp per row, 8 independent 64-bit loads per lane in flight ----
        if (gt == 0) cnt[(round + 1) & 1] = 0;
        for (int i = gw; i < N; i += NW) {
            if (alive_s[i] != 3) continue;                                      // warp-uniform: alive and marked
            const double* row = dm + (size_t)i * N;
            double bv = INFINITY;
            int bi = kNoIdx;
            bool tied = false;
            for (int c0 = lane; c0 < N; c0 += 32 * kRoundMlp) {
                double v[kRoundMlp];
#pragma unroll
                for (int u = 0; u < kRoundMlp; ++u) { const int c = c0 + 32 * u; v[u] = c < N ? __ldcg(row + c) : INFINITY; }
#pragma unroll
                for (int u = 0; u < kRoundMlp; ++u) {
                    const int c = c0 + 32 * u;
                    if (c < N && c != i && (alive_s[c] & 1)) {
                        if (v[u] < bv) { bv = v[u]; bi = c; tied = false; }
                        else if (v[u] == bv) tied = true;
                    }
                }
            }
            const unsigned long long bits = (unsigned long long)__double_as_longlong(bv);
            LexKey k{(unsigned)(bits >> 32), (unsigned)bits, bi};
            const LexKey mine = k;
            warp_lexmin(k);
            const bool at_min = mine.hi == k.hi && mine.lo == k.lo && mine.idx != kNoIdx;
            const bool any_tie = __any_sync(kFull, at_min && (tied || mine.idx != k.idx));
            if (lane == 0) {
                nn[i] = k.idx;
                nnd[i] = key_value(k);
                rescan[i] = 0;
                if (any_tie) tie_all[b] = 1;
            }
        }
        tick(1);
        cluster_sync_all();
        tick(5);
        // ---- 2. reciprocal pairs ----------------------------------------------------------------------------------------------
        for (int i = gt; i < N; i += NT) {
            if (!(alive_s[i] & 1)) continue;
            const int j = __ldcg(nn + i);
            const bool mutual = j != kNoIdx && (alive_s[j] & 1) && __ldcg(nn + j) == i;
            partner[i] = mutual ? j : -1;
            if (mutual && i < j) {
                const int s = atomicAdd(cnt + (round & 1), 1);
                pairs[s] = make_int2(i, j);
                recx[nrec + s] = i;
                recy[nrec + s] = j;
                rech[nrec + s] = __ldcg(nnd + i);
            }
        }
        tick(2);
        cluster_sync_all();
        tick(5);
        const int np = __ldcg(cnt + (round & 1));
        if (np == 0) break;                                                     // uniform over the cluster: nothing mergeable is left
        for (int i = threadIdx.x; i < N; i += kRoundThreads) partner_s[i] = (alive_s[i] & 1) ? __ldcg(partner + i) : -2;   // -2: not alive
        __syncthreads();
        // ---- 3. scratch: row x <- max(row x, row y), every column (dead ones are never read again).  Work items are (pair, chunk
        // of 32 * kRoundMlp columns) so that late rounds with few pairs still spread over all warps of the cluster ------------------
        const int nchunk = (N + 32 * kRoundMlp - 1) / (32 * kRoundMlp);
        for (int item = gw; item < np * nchunk; item += NW) {
            const int p = item / nchunk, c0 = (item - p * nchunk) * 32 * kRoundMlp + lane;
            const int2 xy = __ldcg(pairs + p);
            double* rx = dm + (size_t)xy.x * N;
            const double* ry = dm + (size_t)xy.y * N;
            double a[kRoundMlp], bb[kRoundMlp];
#pragma unroll
            for (int u = 0; u < kRoundMlp; ++u) { const int c = c0 + 32 * u; a[u] = c < N ? __ldcg(rx + c) : 0.0; bb[u] = c < N ? __ldcg(ry + c) : 0.0; }
#pragma unroll
            for (int u = 0; u < kRoundMlp; ++u) { const int c = c0 + 32 * u; if (c < N) rx[c] = fmax(a[u], bb[u]); }
        }
        tick(3);
        cluster_sync_all();
        tick(5);
        // ---- 4. row y, column y, invalidations -----------------------------------------------------------------------------------
        for (int item = gw; item < np * nchunk; item += NW) {
            const int p = item / nchunk, q = item - p * nchunk, c0 = q * 32 * kRoundMlp + lane;
            const int2 xy = __ldcg(pairs + p);
            const double* rx = dm + (size_t)xy.x * N;
            double* ry = dm + (size_t)xy.y * N;
            double v[kRoundMlp];
            int nc[kRoundMlp];
#pragma unroll
            for (int u = 0; u < kRoundMlp; ++u) {
                const int c = c0 + 32 * u;
                v[u] = c < N ? __ldcg(rx + c) : 0.0;
                nc[u] = c < N ? __ldcg(nn + c) : -1;
            }
#pragma unroll
            for (int u = 0; u < kRoundMlp; ++u) {
                const int c = c0 + 32 * u;
                if (c >= N || c == xy.x || c == xy.y) continue;
                const int pc = partner_s[c];
                if (pc == -1) {                                                 // an unmerged cluster
                    ry[c] = v[u];
                    dm[(size_t)c * N + xy.y] = v[u];
                    if (nc[u] == xy.x || nc[u] == xy.y) rescan[c] = 1;
                } else if (pc >= 0 && c > pc) {                                 // the surviving end of another pair of this round
                    ry[c] = fmax(v[u], __ldcg(rx + pc));
                }
            }
            if (q == 0 && lane == 0) { alive[xy.x] = 0; rescan[xy.y] = 1; }
        }
        nrec += np;
        n_alive -= np;
        tick(4);
        cluster_sync_all();
        tick(5);
        if (np < stop_frac) { ++round; break; }                                 // hardly anything reciprocal is left: the serial chain finishes
    }
    if (gt == 0) {
        cnt[2] = nrec; cnt[3] = round;
        for (int q = 0; q < 6; ++q) cnt[4 + q] = (int)(t_acc[q] / 1000);       // microseconds per phase (6 = barriers), diagnostics
    }
}

__global__ void gather_rec0_kernel(const int* __restrict__ cnt, int B, int* __restrict__ rec0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) rec0[b] = cnt[16 * b + 2];
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace hpcs

namespace hpcs {

constexpr int kMaxRounds = 3;

// Workspace layout shared by hpcs_linkage_workspace_bytes and hpcs_linkage_f64 (arrays of B clouds each).
struct LinkWs {
    int rounds;                       // Boruvka rounds before Prim (single linkage, N >= 256), else 0
    int pitch[kMaxRounds + 1];        // row pitch (= node capacity) of matrix r
    size_t off_m[kMaxRounds + 1], off_recx, off_recy, off_rech, off_rmw, off_rmj, off_memb, off_moff, off_nd, off_tie, off_norm;
    size_t off_rep[kMaxRounds + 1];
    bool par_complete;                // complete linkage by parallel reciprocal-nearest-neighbour rounds (N >= 64)
    size_t off_alive, off_nn, off_nnd, off_partner, off_rescan, off_pairs, off_cnt;
    size_t total;
};

// Complete linkage: the parallel rounds pay when the GPU is not already full of independent clouds -- one serial CTA per cloud
// keeps B SMs busy, so from B ~ a quarter of the SMs up the serial kernel is as fast or faster.  Measured (ms per call, serial ->
// rounds): N = 1024: B = 8 3.08 -> 1.08, B = 32 3.66 -> 2.76, B = 64 4.02 -> 5.14; N = 4096: B = 8 24.8 -> 10.7, B = 64 35.1 -> 57.2;
// N = 8192, B = 16: 93 -> 77.  A round costs four cluster barriers (~3 us each with the skew between CTAs) and a few L2 latencies.
static bool use_parallel_complete(int B, int N) {
    if (N < 64 || N > kRoundMaxN) return false;
    if (const char* force = getenv("HPCS_COMPLETE_LINKAGE")) {                  // measurement switch: "serial" | "rounds"
        if (force[0] == 's') return false;
        if (force[0] == 'r') return true;
    }
    return B * 4 <= sm_count();
}

static LinkWs link_ws(int B, int N, int method) {
    LinkWs L{};
    L.rounds = (method == 0 && N >= 256) ? (N < 512 ? 2 : 3) : 0;
    L.pitch[0] = N;
    for (int r = 1; r <= kMaxRounds; ++r) L.pitch[r] = (L.pitch[r - 1] / 2 + 3) / 4 * 4;    // a round at least halves the nodes
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t at = off; off += align_up(bytes, 256); return at; };
    for (int r = 0; r <= L.rounds; ++r) L.off_m[r] = take((size_t)B * L.pitch[r] * L.pitch[r] * sizeof(double));
    L.off_recx = take((size_t)B * N * sizeof(int));
    L.off_recy = take((size_t)B * N * sizeof(int));
    L.off_rech = take((size_t)B * N * sizeof(double));
    L.off_norm = take((size_t)B * N * sizeof(double));
    if (L.rounds) {
        L.off_rmw = take((size_t)B * N * sizeof(double));
        L.off_rmj = take((size_t)B * N * sizeof(int));
        L.off_memb = take((size_t)B * N * sizeof(int));
        L.off_moff = take((size_t)B * (N + 1) * sizeof(int));
        L.off_nd = take((size_t)B * 8 * sizeof(int));
        L.off_tie = take((size_t)B * sizeof(int));
        for (int r = 1; r <= L.rounds; ++r) L.off_rep[r] = take((size_t)B * L.pitch[r] * sizeof(int));
    }
    L.par_complete = method == 1 && use_parallel_complete(B, N);
    if (L.par_complete) {
        L.off_alive = take((size_t)B * N * sizeof(int));
        L.off_nn = take((size_t)B * N * sizeof(int));
        L.off_nnd = take((size_t)B * N * sizeof(double));
        L.off_partner = take((size_t)B * N * sizeof(int));
        L.off_rescan = take((size_t)B * N * sizeof(int));
        L.off_pairs = take((size_t)B * (N / 2 + 1) * sizeof(int2));
        L.off_cnt = take((size_t)B * 16 * sizeof(int));
        L.off_tie = take((size_t)B * sizeof(int));
    }
    L.total = off;
    return L;
}

template <int METHOD>
static void launch_linkage(int cpt, int B, int threads, size_t smem, cudaStream_t st, double* dm, size_t dm_stride, int pitch,
                           const int* nd, int nd_slot, const int* rep, int rep_stride, int N, int NP2, int* recx, int* recy,
                           double* rech, double* Z, int* tie_flag, const int* gate, const int* alive_in = nullptr,
                           const int* rec0_in = nullptr) {
    auto go = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<B, threads, smem, st>>>(dm, dm_stride, pitch, nd, nd_slot, rep, rep_stride, N, NP2, recx, recy, rech, Z, tie_flag, gate, alive_in, rec0_in);
    };
    if (cpt == 1) go(linkage_kernel<METHOD, 1>);
    else if (cpt == 2) go(linkage_kernel<METHOD, 2>);
    else if (cpt == 4) go(linkage_kernel<METHOD, 4>);
    else go(linkage_kernel<METHOD, 8>);
}

}  // namespace hpcs

extern "C" {

/* diagnostics (tools/decode_rounds.py): byte offset of the per-cloud round counters [B][16] in the workspace, 0 if the parallel
 * complete-linkage path does not apply */
size_t hpcs_linkage_debug_counters_offset(int B, int N, int method) {
    const hpcs::LinkWs L = hpcs::link_ws(B, N, method);
    return L.par_complete ? L.off_cnt : 0;
}

size_t hpcs_linkage_workspace_bytes(int B, int N, int D, int method) {
    (void)D;
    if (B <= 0 || N <= 1) return 0;
    return hpcs::link_ws(B, N, method).total;
}

int hpcs_linkage_f64(const float* leaves, int B, int N, int D, int method, double* Z, void* ws,
                     size_t ws_bytes, void* stream) {
    using namespace hpcs;
    if (!leaves || !Z || !ws) return fail(HPCS_ERR_ARG, "linkage: null pointer");
    if (B <= 0 || N < 2 || D <= 0 || (method != 0 && method != 1)) return fail(HPCS_ERR_ARG, "linkage: bad arguments B=%d N=%d D=%d method=%d", B, N, D, method);
    if (B > 65535) return fail(HPCS_ERR_ARG, "linkage: B > 65535");
    if (ws_bytes < hpcs_linkage_workspace_bytes(B, N, D, method)) return fail(HPCS_ERR_WORKSPACE, "linkage: workspace too small");
    const int NP2 = next_pow2(N - 1 > 1 ? N - 1 : 2);
    const size_t smem_link = ((size_t)12 * (NP2 > N ? NP2 : N) + 15) / 16 * 16 + (size_t)16 * N;
    if (smem_link > 227 * 1024) return fail(HPCS_ERR_ARG, "linkage: N=%d too large (max 8192)", N);
    cudaStream_t st = as_stream(stream);
    const LinkWs L = link_ws(B, N, method);
    char* w = static_cast<char*>(ws);
    double* dm = reinterpret_cast<double*>(w + L.off_m[0]);
    int* recx = reinterpret_cast<int*>(w + L.off_recx);
    int* recy = reinterpret_cast<int*>(w + L.off_recy);
    double* rech = reinterpret_cast<double*>(w + L.off_rech);

    const int T = (N + kPT - 1) / kPT;
    double* norms = reinterpret_cast<double*>(w + L.off_norm);
    pdist_norms_kernel<<<(unsigned)(((size_t)B * N + 255) / 256), 256, 0, st>>>(leaves, (size_t)B * N, D, norms);
    int rc = check_launch("pdist_norms_kernel");
    if (rc) return rc;
    {
        const int S8 = 8 * ((D / 2 + 3) / 4) + 1;                                  // staged feature rows per tile
        const size_t smem_mma = ((size_t)2 * S8 * kPLD + 2 * kPT) * sizeof(double);
        if (smem_mma > 200 * 1024) return fail(HPCS_ERR_ARG, "linkage: D=%d too large", D);
        const unsigned d_magic = 0xFFFFFFFFu / (unsigned)D + 1u;                 // e / D == umulhi(e, magic) for e < 2^16
        const int* pdist_gate = nullptr;
        auto go = [&](auto kern) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mma);
            kern<<<dim3(T, B), 256, smem_mma, st>>>(leaves, norms, N, D, d_magic, dm, pdist_gate);
        };
        if (D <= 32) go(pdist_mma_kernel<8>);
        else if (D <= 64) go(pdist_mma_kernel<16>);
        else go(pdist_mma_kernel<0>);
        if ((rc = check_launch("pdist_mma_kernel"))) return rc;
    }
    // direct form: columns per thread 1 up to 512 points, else the smallest of 2/4/8 that covers N with <= 1024 threads
    const int cpt0 = N <= 512 ? 1 : N <= 2048 ? 2 : N <= 4096 ? 4 : 8;
    const int threads0 = ((N + cpt0 - 1) / cpt0 + 31) / 32 * 32;
    const size_t stride0 = (size_t)N * N;
    if (L.par_complete) {
        // ---- complete linkage: parallel reciprocal-nearest-neighbour rounds (one 8-CTA cluster per cloud), the serial kernel for
        // whatever is left + sort + relabel, and an exact serial redo (on a recomputed matrix) of clouds with tied distances ----
        int* alive = reinterpret_cast<int*>(w + L.off_alive);
        int* cnt = reinterpret_cast<int*>(w + L.off_cnt);
        int* tie = reinterpret_cast<int*>(w + L.off_tie);
        cudaLaunchConfig_t cfg = {};
        const int cs = kClusterCtas;                                            // (16-CTA clusters measured slower: 1.70 vs 1.24 ms at B = 8)
        cfg.gridDim = dim3(cs, B);
        cfg.blockDim = dim3(kRoundThreads);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t ce;
        int* nn_p = reinterpret_cast<int*>(w + L.off_nn);
        double* nnd_p = reinterpret_cast<double*>(w + L.off_nnd);
        int* partner_p = reinterpret_cast<int*>(w + L.off_partner);
        int* rescan_p = reinterpret_cast<int*>(w + L.off_rescan);
        int2* pairs_p = reinterpret_cast<int2*>(w + L.off_pairs);
        if (cs == 16) {
            cudaFuncSetAttribute(complete_rounds_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            ce = cudaLaunchKernelEx(&cfg, complete_rounds_kernel<16>, dm, N, 3, alive, nn_p, nnd_p, partner_p, rescan_p, pairs_p, cnt, recx, recy, rech, tie);
        } else {
            ce = cudaLaunchKernelEx(&cfg, complete_rounds_kernel<8>, dm, N, 3, alive, nn_p, nnd_p, partner_p, rescan_p, pairs_p, cnt, recx, recy, rech, tie);
        }
        if (ce != cudaSuccess) return fail(HPCS_ERR_CUDA, "complete_rounds_kernel: %s", cudaGetErrorString(ce));
        if ((rc = check_launch("complete_rounds_kernel"))) return rc;
        // rec0 of cloud b lives at cnt[4 b + 2]: hand the kernel a strided view by pointing at element 2 with stride 4 ... the
        // kernel indexes rec0_in[b], so compact it first (B ints)
        int* rec0 = reinterpret_cast<int*>(w + L.off_partner);                   // partner[] is dead after the rounds
        gather_rec0_kernel<<<(B + 255) / 256, 256, 0, st>>>(cnt, B, rec0);
        if ((rc = check_launch("gather_rec0_kernel"))) return rc;
        launch_linkage<1>(cpt0, B, threads0, smem_link, st, dm, stride0, N, nullptr, 0, nullptr, 0, N, NP2, recx, recy, rech, Z, tie, nullptr, alive, rec0);
        if ((rc = check_launch("linkage_kernel"))) return rc;
        {
            const int S8 = 8 * ((D / 2 + 3) / 4) + 1;
            const size_t smem_mma = ((size_t)2 * S8 * kPLD + 2 * kPT) * sizeof(double);
            const unsigned d_magic = 0xFFFFFFFFu / (unsigned)D + 1u;
            if (D <= 32) pdist_mma_kernel<8><<<dim3(T, B), 256, smem_mma, st>>>(leaves, norms, N, D, d_magic, dm, tie);
            else if (D <= 64) pdist_mma_kernel<16><<<dim3(T, B), 256, smem_mma, st>>>(leaves, norms, N, D, d_magic, dm, tie);
            else pdist_mma_kernel<0><<<dim3(T, B), 256, smem_mma, st>>>(leaves, norms, N, D, d_magic, dm, tie);
            if ((rc = check_launch("pdist_mma_kernel(redo)"))) return rc;
        }
        launch_linkage<1>(cpt0, B, threads0, smem_link, st, dm, stride0, N, nullptr, 0, nullptr, 0, N, NP2, recx, recy, rech, Z, nullptr, tie);
        return check_launch("linkage_kernel(redo)");
    }
    if (L.rounds == 0) {
        if (method == 0) launch_linkage<0>(cpt0, B, threads0, smem_link, st, dm, stride0, N, nullptr, 0, nullptr, 0, N, NP2, recx, recy, rech, Z, nullptr, nullptr);
        else launch_linkage<1>(cpt0, B, threads0, smem_link, st, dm, stride0, N, nullptr, 0, nullptr, 0, N, NP2, recx, recy, rech, Z, nullptr, nullptr);
        return check_launch("linkage_kernel");
    }
    // ---- single linkage: Boruvka rounds, Prim on the contracted matrix, exact redo of clouds with tied heights ----
    double* rmw = reinterpret_cast<double*>(w + L.off_rmw);
    int* rmj = reinterpret_cast<int*>(w + L.off_rmj);
    int* memb = reinterpret_cast<int*>(w + L.off_memb);
    int* moff = reinterpret_cast<int*>(w + L.off_moff);
    int* nd = reinterpret_cast<int*>(w + L.off_nd);
    int* tie = reinterpret_cast<int*>(w + L.off_tie);
    boruvka_rowmin_kernel<<<dim3(N, B), 256, 0, st>>>(dm, stride0, N, nd, 0, rmw, rmj, N);
    if ((rc = check_launch("boruvka_rowmin_kernel"))) return rc;
    for (int r = 0; r < L.rounds; ++r) {
        const int cap = L.pitch[r], cap2 = L.pitch[r + 1];
        const int* rep_in = r ? reinterpret_cast<int*>(w + L.off_rep[r]) : nullptr;
        int* rep_out = reinterpret_cast<int*>(w + L.off_rep[r + 1]);
        const size_t smem_hook = (size_t)3 * cap * sizeof(int);
        cudaFuncSetAttribute(boruvka_hook_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hook);
        boruvka_hook_kernel<<<B, cap >= 1024 ? 1024 : (cap + 31) / 32 * 32, smem_hook, st>>>(rmw, rmj, nd, r, rep_in, cap, rep_out, cap2, memb, moff,
                                                                                           recx, recy, rech, N);
        if ((rc = check_launch("boruvka_hook_kernel"))) return rc;
        const double* in = reinterpret_cast<double*>(w + L.off_m[r]);
        double* out = reinterpret_cast<double*>(w + L.off_m[r + 1]);
        const size_t smem_con = (size_t)cap * sizeof(double);
        cudaFuncSetAttribute(boruvka_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_con);
        boruvka_contract_kernel<<<dim3(cap2, B), 256, smem_con, st>>>(in, (size_t)cap * cap, cap, out, (size_t)cap2 * cap2, cap2, nd, r, memb, moff, rmw, rmj, N);
        if ((rc = check_launch("boruvka_contract_kernel"))) return rc;
    }
    const int capR = L.pitch[L.rounds];
    const int threadsR = N >= 2048 ? 1024 : 512;
    const int cptR = (capR + threadsR - 1) / threadsR <= 1 ? 1 : (capR + threadsR - 1) / threadsR <= 2 ? 2 : (capR + threadsR - 1) / threadsR <= 4 ? 4 : 8;
    launch_linkage<0>(cptR, B, threadsR, smem_link, st, reinterpret_cast<double*>(w + L.off_m[L.rounds]), (size_t)capR * capR, capR, nd, L.rounds,
                      reinterpret_cast<int*>(w + L.off_rep[L.rounds]), capR, N, NP2, recx, recy, rech, Z, tie, nullptr);
    if ((rc = check_launch("linkage_kernel"))) return rc;
    launch_linkage<0>(cpt0, B, threads0, smem_link, st, dm, stride0, N, nullptr, 0, nullptr, 0, N, NP2, recx, recy, rech, Z, nullptr, tie);
    return check_launch("linkage_kernel(redo)");
}

}  // extern "C"
