// kNN on point coordinates (D = 3): the first EdgeConv layer, knn() of
// hpcs/nn/dgcnn/utils/vn_dgcnn_util.py:4-10 called at vn_dgcnn_partseg.py:65 with x[B,3,N].
//
// With three features a distance costs four FFMA; the cost of kNN is the SELECTION of k = 20 out of N = 1024, and a
// sorted-list insertion per thread (knn_rows_kernel) spends three quarters of its instructions there, at one thread
// per query row (7 warps per SM at B*N = 32768).  Here one WARP owns a query row and the 32 lanes split the
// candidates (lane l holds candidates l, l+32, ...), so a cloud gives 32x the threads and selection becomes a
// threshold problem:
//   pass 1   every lane evaluates its N/32 canonical distances with packed FFMA2 (two candidates per instruction)
//            and keeps their maximum;
//   tau      the k-th largest of the 32 lane maxima (bitonic sort across the warp).  k lanes hold a value >= tau, so
//            the k-th best distance of the row is >= tau: everything below tau is out, exactly;
//   pass 2   the distances are evaluated again (cheaper than holding N/32 registers per row: occupancy); the survivors
//            (about 1.5 k of them, >= tau, ties included) are marked in a per-lane bit mask and appended
//            to the row's list in shared memory at offsets from a warp scan of the per-lane counts;
//   rank     each survivor counts the survivors that precede it in the canonical order (larger pd first, equal pd ->
//            lower index); rank < k is its output slot.  No sort network, no insertion.
// A row with more survivors than the lists hold (heavy exact ties: duplicated points) is redone by k rounds of
// warp arg-max over the registers -- exact for any input, just slower.
//
// Canonical arithmetic, bit-exact with oracle/knn_canonical.c:  sq = fma chain over d from 0;  dot = fma chain over d
// from 0;  pd = fmaf(2, dot, -sq_i) - sq_j.  FFMA2 rounds each half like FFMA, and t - s == t + (-s).
#include <float.h>

#include "common.cuh"

namespace hpcs {

constexpr int kD3Warps = 8;
constexpr int kD3RowsPerWarp = 4;
constexpr int kD3ListCap = 64;         // survivors per row on the fast path

struct D3Key {
    float v;
    int j;
};

__device__ __forceinline__ bool d3_before(float v, int j, float ov, int oj) { return v > ov || (v == ov && j < oj); }

// S = candidate slots per lane (N <= 32 S), a power of two >= 2
template <int S>
__global__ void __launch_bounds__(kD3Warps * 32, S <= 32 ? 5 : 3)      // 48 registers: 40 warps per SM up to N = 1024
knn_d3_kernel(const float* __restrict__ x, int N, int k, int64_t* __restrict__ idx, float* __restrict__ val) {
    constexpr int P = S / 2;                                        // candidate pairs per lane
    extern __shared__ __align__(16) unsigned char smraw[];
    float4* cxy = reinterpret_cast<float4*>(smraw);                 // [P][32] {x_a, x_b, y_a, y_b}: a = slot 2p, b = slot 2p+1
    float4* czs = cxy + P * 32;                                     // [P][32] {z_a, z_b, -sq_a, -sq_b}
    unsigned long long* dense = reinterpret_cast<unsigned long long*>(czs + P * 32);   // [warps][kD3ListCap]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* xb = x + (size_t)b * 3 * N;

    // stage the cloud: candidate j -> lane j % 32, slot j / 32; padding gets -sq = -inf so its pd is -inf
    for (int j = threadIdx.x; j < 32 * S; j += blockDim.x) {
        float cx = 0.f, cy = 0.f, cz = 0.f, nsq = -INFINITY;
        if (j < N) {
            cx = __ldg(xb + j); cy = __ldg(xb + N + j); cz = __ldg(xb + 2 * N + j);
            nsq = -__fmaf_rn(cz, cz, __fmaf_rn(cy, cy, __fmaf_rn(cx, cx, 0.f)));
        }
        const int s = j >> 5, l = j & 31;
        float* pa = reinterpret_cast<float*>(cxy + (s >> 1) * 32 + l) + (s & 1);
        float* pb = reinterpret_cast<float*>(czs + (s >> 1) * 32 + l) + (s & 1);
        pa[0] = cx; pa[2] = cy;
        pb[0] = cz; pb[2] = nsq;
    }
    __syncthreads();

    unsigned long long* mydense = dense + warp * kD3ListCap;
    const int row0 = (blockIdx.x * kD3Warps + warp) * kD3RowsPerWarp;
    for (int r = 0; r < kD3RowsPerWarp; ++r) {
        const int i = row0 + r;
        if (i >= N) break;                                          // warp-uniform
        // the query point, from the staged cloud
        const int qs = i >> 5, ql = i & 31;
        const float4 qa = cxy[(qs >> 1) * 32 + ql], qb = czs[(qs >> 1) * 32 + ql];
        const float qx = (qs & 1) ? qa.y : qa.x, qy = (qs & 1) ? qa.w : qa.z, qz = (qs & 1) ? qb.y : qb.x;
        const float nsq_i = (qs & 1) ? qb.w : qb.z;
        const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz);
        const float2 two = make_float2(2.f, 2.f), nsq2 = make_float2(nsq_i, nsq_i), zero = make_float2(0.f, 0.f);

        // ---- pass 1: lane maximum of the lane's distances.  The distances are NOT kept: pass 2 evaluates them again
        //      (5 FFMA2 per pair, same bits) -- a register array of N/32 floats per row costs more in occupancy than the
        //      second evaluation costs in issue slots.
        auto pair_dist = [&](int p) -> float2 {
            const float4 a = cxy[p * 32 + lane], c = czs[p * 32 + lane];
            float2 acc = __ffma2_rn(qx2, make_float2(a.x, a.y), zero);
            acc = __ffma2_rn(qy2, make_float2(a.z, a.w), acc);
            acc = __ffma2_rn(qz2, make_float2(c.x, c.y), acc);
            const float2 t = __ffma2_rn(two, acc, nsq2);
            return __fadd2_rn(t, make_float2(c.z, c.w));
        };
        float lmax = -INFINITY;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 d = pair_dist(p);
            lmax = fmaxf(lmax, fmaxf(d.x, d.y));
        }
        // ---- tau: k-th largest lane maximum (bitonic sort, descending across lanes) ----
        float sv = lmax;
#pragma unroll
        for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
            for (int jj = kk >> 1; jj > 0; jj >>= 1) {
                const float o = __shfl_xor_sync(kFull, sv, jj);
                const bool desc = (lane & kk) == 0;                   // this block sorts descending
                const bool lower = (lane & jj) == 0;                  // this lane keeps the first of the pair
                sv = (desc == lower) ? fmaxf(sv, o) : fminf(sv, o);
            }
        }
        const float tau = __shfl_sync(kFull, sv, k - 1);
        // ---- pass 2: a bit per surviving slot, then every lane appends its survivors (about one per lane) to the
        //      row's dense list at the offset a warp scan of the counts gives it.  A survivor's distance is evaluated
        //      again from shared memory (same fma chain, same bits) because a register array cannot be indexed by a
        //      run-time slot.  A key packs the order-preserving image of pd (high word) and ~index (low word), so
        //      "precedes in the canonical order" is one unsigned 64-bit compare.
        unsigned mask = 0u, mask_hi = 0u;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2 d = pair_dist(p);
            const unsigned bits = (d.x >= tau ? 1u : 0u) | (d.y >= tau ? 2u : 0u);
            if (2 * p < 32) mask |= bits << ((2 * p) & 31); else mask_hi |= bits << ((2 * p) & 31);
        }
        const int cnt = __popc(mask) + __popc(mask_hi);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(kFull, incl, 31);
        const bool slow = total > kD3ListCap;
        int64_t* oi = idx + ((size_t)b * N + i) * k;
        float* ov = val ? val + ((size_t)b * N + i) * k : nullptr;
        if (!slow) {
            int pos = incl - cnt;
            auto push = [&](int s) {
                const float4 a = cxy[(s >> 1) * 32 + lane], c = czs[(s >> 1) * 32 + lane];
                const bool odd = s & 1;
                float acc = __fmaf_rn(qx, odd ? a.y : a.x, 0.f);
                acc = __fmaf_rn(qy, odd ? a.w : a.z, acc);
                acc = __fmaf_rn(qz, odd ? c.y : c.x, acc);
                const float d = __fadd_rn(__fmaf_rn(2.f, acc, nsq_i), odd ? c.w : c.z);
                const unsigned u = __float_as_uint(d + 0.f);          // -0 -> +0 (equal values must compare equal)
                const unsigned ord = u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
                mydense[pos++] = ((unsigned long long)ord << 32) | (unsigned)~(s * 32 + lane);
            };
            while (mask) { const int s = __ffs(mask) - 1; mask &= mask - 1u; push(s); }
            if (S > 32) while (mask_hi) { const int s = __ffs(mask_hi) - 1; mask_hi &= mask_hi - 1u; push(s + 32); }
            __syncwarp();
            // rank by counting; total <= 64: two entries per lane
            const unsigned long long e0 = lane < total ? mydense[lane] : 0ull;
            const unsigned long long e1 = lane + 32 < total ? mydense[lane + 32] : 0ull;
            int r0 = 0, r1 = 0;
            if (total <= 32) {                                        // warp-uniform: the usual case, one entry per lane
#pragma unroll 4
                for (int f = 0; f < total; ++f) r0 += mydense[f] > e0 ? 1 : 0;        // broadcast load
            } else {
#pragma unroll 4
                for (int f = 0; f < total; ++f) {
                    const unsigned long long o = mydense[f];
                    r0 += o > e0 ? 1 : 0;
                    r1 += o > e1 ? 1 : 0;
                }
            }
            auto emit = [&](unsigned long long key, int rank) {
                const unsigned ord = (unsigned)(key >> 32);
                const unsigned u = ord ^ ((ord >> 31) ? 0x80000000u : 0xffffffffu);
                oi[rank] = (int)~(unsigned)key;
                if (ov) ov[rank] = __uint_as_float(u);
            };
            if (lane < total && r0 < k) emit(e0, r0);
            if (lane + 32 < total && r1 < k) emit(e1, r1);
            __syncwarp();                                             // lists are reused by the next row
        } else {
            // exact for any input: k rounds of "best candidate after the previous pick" over the registers
            float cv = INFINITY;
            int cj = -1;
            for (int m = 0; m < k; ++m) {
                float bv = -INFINITY;
                int bj = 0x7fffffff;
                for (int p = 0; p < P; ++p) {
                    const float2 d = pair_dist(p);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = (2 * p + h) * 32 + lane;
                        const float v = h ? d.y : d.x;
                        const bool after = j < N && d3_before(cv, cj, v, j);              // strictly after the last pick
                        if (after && d3_before(v, j, bv, bj)) { bv = v; bj = j; }
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ovv = __shfl_xor_sync(kFull, bv, o);
                    const int ojj = __shfl_xor_sync(kFull, bj, o);
                    if (d3_before(ovv, ojj, bv, bj)) { bv = ovv; bj = ojj; }
                }
                if (lane == 0) { oi[m] = bj; if (ov) ov[m] = bv; }
                cv = bv; cj = bj;
            }
        }
    }
}

bool knn_d3_applicable(int D, int N, int k) { return D == 3 && k <= 32 && N >= k && N <= 2048; }

int knn_d3_run(const float* x, int B, int N, int k, int64_t* idx, float* val, cudaStream_t st) {
    const int slots = (N + 31) / 32;
    const dim3 grid((N + kD3Warps * kD3RowsPerWarp - 1) / (kD3Warps * kD3RowsPerWarp), B), block(kD3Warps * 32);
    auto launch = [&](auto kern, int S) {
        const size_t smem = (size_t)S * 32 * sizeof(float4) + (size_t)kD3Warps * kD3ListCap * sizeof(unsigned long long);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, block, smem, st>>>(x, N, k, idx, val);
    };
    if (slots <= 4) launch(knn_d3_kernel<4>, 4);
    else if (slots <= 8) launch(knn_d3_kernel<8>, 8);
    else if (slots <= 16) launch(knn_d3_kernel<16>, 16);
    else if (slots <= 32) launch(knn_d3_kernel<32>, 32);
    else launch(knn_d3_kernel<64>, 64);
    return check_launch("knn_d3_kernel");
}

}  // namespace hpcs
