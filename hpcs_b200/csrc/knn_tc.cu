// kNN for the feature-space EdgeConv layers (16 <= D <= 63): Gram term on the 5th-generation tensor
// cores, exact fp32 re-rank, so the selected indices stay bit-exact with the canonical order.
//
//   0. knn_mean_kernel      per-cloud feature means mu[b,d] (distances are translation invariant: the tensor
//                           cores see x - mu, so a common offset costs no TF32 precision; the exact stages
//                           below always use the raw x).
//   1. knn_pack_kernel      x[B,D,N] -> two K-major row arrays of 64 fp32 per point:
//                             xc[j] = (x_j - mu, -|x_j - mu|^2/2)   tensor-core operand (A and B),
//                             xr[j] = (x_j,      -|x_j|^2/2)        raw rows for the exact re-rank,
//                           + canonical raw norms, centred norms and their per-cloud maximum.  The A tile gets
//                           a 1 patched into column 63 in shared memory, so the tensor-core product is directly
//                           the ranking key  S_ij = c_i.c_j - |c_j|^2/2  (= pd_ij/2 + |c_i|^2/2).
//   2. knn_tc_kernel        one CTA per (cloud, 128 query rows), 192 threads: warp 4 streams 64-candidate tiles
//                           of xc through a 3-stage ring with TMA (128B-swizzled boxes), warp 5 issues
//                           tcgen05.mma kind::tf32 (M=128, N=64, K=8 x 8) into a double-buffered TMEM
//                           accumulator; warps 0-3 are the epilogue (thread = query row = TMEM lane) and read
//                           the scores with tcgen05.ld.  Selection is two passes over the Gram tiles: pass 1
//                           keeps chunk maxima and a sorting network picks the row threshold, pass 2 pushes the
//                           scores above it through a shared-memory FIFO into a sorted 32-list.  A key is the
//                           shifted score pd/2 with the candidate index in its low mantissa bits.
//   3. knn_rerank_kernel    one warp per row: canonical fp32 fma-chain distance of the KL candidates
//                           (rows of xb staged coalesced through shared memory), bitonic sort by
//                           (value desc, index asc), and a safety test: the k-th exact value must beat the
//                           best value any non-candidate can have (KL-th key + TF32/quantisation error bound).
//   4. knn_fallback_kernel  rows that fail the test (near-ties denser than the error bound, duplicates)
//                           are redone exactly over all N candidates; the list lives on the device, no sync.
#include <cuda.h>
#include <float.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace hpcs {

// all-FFMA exact path (knn.cu); runs only when *gate > gate_min (device-side check at kernel entry)
int knn_ffma_gated(const float* x, const float* sq, int B, int D, int N, int k, int64_t* idx, float* val, const int* gate,
                   int gate_min, cudaStream_t st);

constexpr int kKP = 64;                 // padded row length (fp32) of xa / xb
constexpr int kTM = 128;                // query rows per CTA  (UMMA M)
constexpr int kTN = 64;                 // candidates per tile  (UMMA N)
constexpr int kTcQueue = 48;            // per-row FIFO depth in the epilogue
constexpr int kAtomBytes = kTM * 128;   // one 32-fp32 K-atom of a 128-row tile

// ---- 0. per-cloud feature means ---------------------------------------------------------------------------
// one warp per (cloud, feature) row of x; block 0 also clears the per-call counters
__global__ void __launch_bounds__(256)
knn_mean_kernel(const float* __restrict__ x, int rows, int N, float* __restrict__ mu, unsigned* __restrict__ zero_words,
                int n_zero) {
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < n_zero; i += blockDim.x) zero_words[i] = 0u;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * N;
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += __ldg(xr + j);
    s = warp_sum(s);
    if (lane == 0) mu[row] = s / (float)N;
}

// ---- 1. pack ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
knn_pack_kernel(const float* __restrict__ x, const float* __restrict__ mu, int D, int N, float* __restrict__ xc,
                float* __restrict__ xr, float* __restrict__ sq, float* __restrict__ sqc, unsigned* __restrict__ cmax_bits) {
    __shared__ float tile[kKP][33];
    __shared__ float nrm[2][32];
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
    const float* xb_in = x + (size_t)b * D * N;
    for (int d = ty; d < kKP; d += 8) tile[d][tx] = (d < D && n0 + tx < N) ? __ldg(xb_in + (size_t)d * N + n0 + tx) : 0.f;
    __syncthreads();
    if (ty < 2 && n0 + tx < N) {                                       // fma chains over d ascending (ty 0: raw = canonical)
        float s = 0.f;
        for (int d = 0; d < D; ++d) {
            const float v = ty == 0 ? tile[d][tx] : __fsub_rn(tile[d][tx], __ldg(mu + b * D + d));
            s = __fmaf_rn(v, v, s);
        }
        nrm[ty][tx] = s;
        if (ty == 0) {
            sq[(size_t)b * N + n0 + tx] = s;
            atomicMax(cmax_bits + gridDim.y + b, __float_as_uint(s));  // raw maximum lives behind the centred one
        } else {
            sqc[(size_t)b * N + n0 + tx] = s;
            atomicMax(cmax_bits + b, __float_as_uint(s));              // s >= 0: uint order == float order
        }
    }
    __syncthreads();
    // write rows: thread (ty, tx) -> rows ty, ty+8, ..; columns tx and tx+32 (coalesced 128 B per half row)
    const float m0 = tx < D ? __ldg(mu + b * D + tx) : 0.f;
    const float m1 = tx + 32 < D ? __ldg(mu + b * D + tx + 32) : 0.f;
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r;
        if (n >= N) continue;
        const size_t row = ((size_t)b * N + n) * kKP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int d = tx + 32 * h;
            float vr = tile[d][r];
            float vc = d < D ? __fsub_rn(vr, h ? m1 : m0) : 0.f;
            if (d == kKP - 1) { vr = -0.5f * nrm[0][r]; vc = -0.5f * nrm[1][r]; }   // column 63 is free (D <= 63)
            // the tensor cores drop the low 13 mantissa bits of an fp32 operand (truncation: 2^-10 relative per operand);
            // rounding to nearest TF32 here makes that a no-op and halves the error bound of the safety test (2^-11)
            unsigned vbits;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(vbits) : "f"(vc));
            vc = __uint_as_float(vbits);
            xr[row + d] = vr;
            xc[row + d] = vc;
        }
    }
}

// ---- key helpers -------------------------------------------------------------------------------------------
// key = fp32 (score - h_i) with the low `bits` mantissa bits replaced by the candidate index
__device__ __forceinline__ float make_key(float shifted, unsigned j, unsigned keep_mask) {
    return __uint_as_float((__float_as_uint(shifted) & keep_mask) | j);
}

template <int KL>
struct KeyTopK {
    float key[KL];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int m = 0; m < KL; ++m) key[m] = -INFINITY;
    }
    __device__ __forceinline__ void insert(float v) {               // v = -inf is a no-op
#pragma unroll
        for (int m = KL - 1; m >= 1; --m) key[m] = fmaxf(key[m], fminf(key[m - 1], v));
        key[0] = fmaxf(key[0], v);
    }
};

// ---- 2. tensor-core candidate kernel --------------------------------------------------------------------------
// Warp roles (192 threads): warps 0-3 epilogue (thread = query row = TMEM lane), warp 4 TMA producer, warp 5 MMA issuer.
// B tiles of kTN = 64 candidates go through a kStages-deep shared-memory ring; the accumulator (128 x 64 fp32) is
// double-buffered in TMEM, so TMA, tcgen05.mma and the selection epilogue of three different tiles overlap.
//
// TWO_PASS (lists of 32, N >= 768): the epilogue is bound by the ALU pipe (min/max of the sorted insert), and a
// running threshold admits ~5x more candidates than end up in the list.  The Gram tiles are so cheap on the tensor
// cores that they are simply computed twice: pass 1 only takes the maximum of every chunk of columns (1 op per
// score) and sets the row threshold to the 24th largest chunk maximum -- at least 24 scores reach it, ~40 on
// average -- and pass 2 inserts just those.  Scores below the threshold are bounded by it in the safety test.
constexpr int kStages = 3;
constexpr int kTcThreads = 192;
constexpr int kChunkMax = 48;            // chunk maxima per row (pass 1), == kTcQueue rows of the shared FIFO
constexpr int kThrRank = 24;             // threshold = kThrRank-th largest chunk maximum (>= k for lists of 32)
constexpr int kAtomB = kTN * 128;        // one 32-fp32 K-atom of a B tile

template <int NV>
__device__ __forceinline__ void sort_desc(float (&c)[NV]) {           // bitonic network, fully unrolled: static indices
#pragma unroll
    for (int size = 2; size <= NV; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool desc = (i & size) == 0;
                    const float hi = fmaxf(c[i], c[j]), lo = fminf(c[i], c[j]);
                    c[i] = desc ? hi : lo;
                    c[j] = desc ? lo : hi;
                }
            }
        }
    }
}

// MODE 0: one selection pass; 1: two passes (threshold from chunk maxima, then selection); 2: SECOND CHANCE for the rows whose
// 32-entry list could not be proven (knn_rerank32_kernel leaves thr2[row] = k-th exact value / 2 - error bound there, +inf
// elsewhere): one pass that appends EVERY candidate whose score reaches the row's threshold to sup[row][..] -- a rigorous
// superset of the row's true k nearest (a score below the threshold is below the k-th exact value even after the TF32 error) --
// for knn_rerank_super_kernel to evaluate exactly.  Gated on the device-side counter of unproven rows.
constexpr int kSupCap = 256;

template <int KL, int MODE>
__global__ void __launch_bounds__(kTcThreads, 2)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const float* __restrict__ sqc,
              int N, unsigned keep_mask, float* __restrict__ cand /*[B*N][KL] keys*/, float* __restrict__ tau_out /*[B*N]*/,
              const float* __restrict__ thr2, unsigned short* __restrict__ sup, int* __restrict__ supcnt, const int* __restrict__ gate) {
    constexpr bool TWO_PASS = MODE == 1;
    constexpr bool SUPER = MODE == 2;
    if (SUPER && *gate == 0) return;                                   // every row was proven: nothing to collect
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the 128B-swizzle atoms must start on 1024-byte boundaries of the shared window
    unsigned char* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                                          // 2 K-atoms x 16 KB
    unsigned char* sB = smem + 2 * kAtomBytes;                         // kStages x 2 K-atoms x 8 KB
    float* qk = reinterpret_cast<float*>(sB + kStages * 2 * kAtomB);   // [kTcQueue][128] pending keys / chunk maxima
    uint64_t* bars = reinterpret_cast<uint64_t*>(qk + kTcQueue * kTM);
    uint64_t* full_a = bars, *a_ready = bars + 1, *full_b = bars + 2, *slot_free = full_b + kStages,
              *acc_full = slot_free + kStages, *epi_done = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 2);

    const int b = blockIdx.y;
    const int m0 = blockIdx.x * kTM;                                   // first query row of this CTA (in the cloud)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int TV = (N + kTN - 1) / kTN;                                // tiles of the cloud
    const int V = TWO_PASS ? 2 * TV : TV;                              // tile visits

    if (threadIdx.x == 0) {
        ptx::mbar_init(full_a, 1);
        ptx::mbar_init(a_ready, 4);
        for (int s = 0; s < kStages; ++s) { ptx::mbar_init(full_b + s, 1); ptx::mbar_init(slot_free + s, 1); }
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(acc_full + s, 1); ptx::mbar_init(epi_done + s, 4); }
        ptx::fence_barrier_init();
    }
    if (warp == 4) ptx::tmem_alloc<2 * kTN>(tmem_slot);
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer (one elected lane) =====
        if (lane == 0) {
            ptx::tma_prefetch_desc(&map_a);
            ptx::tma_prefetch_desc(&map_b);
            const int rowA = b * N + m0;
            ptx::mbar_arrive_expect_tx(full_a, 2 * kAtomBytes);
            ptx::tma_load_2d(sA, &map_a, full_a, 0, rowA);
            ptx::tma_load_2d(sA + kAtomBytes, &map_a, full_a, 32, rowA);
            for (int v = 0; v < V; ++v) {
                const int s = v % kStages, use = v / kStages;
                if (use >= 1) ptx::mbar_wait(slot_free + s, (use - 1) & 1);                 // its previous MMA has retired
                const int rowB = b * N + (v % TV) * kTN;
                unsigned char* dst = sB + s * 2 * kAtomB;
                ptx::mbar_arrive_expect_tx(full_b + s, 2 * kAtomB);
                ptx::tma_load_2d(dst, &map_b, full_b + s, 0, rowB);
                ptx::tma_load_2d(dst + kAtomB, &map_b, full_b + s, 32, rowB);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            const uint32_t idesc = ptx::umma_idesc_tf32(kTM, kTN);
            const uint32_t a_addr = ptx::smem_u32(sA), b_addr = ptx::smem_u32(sB);
            ptx::mbar_wait(a_ready, 0);                                                     // A landed and patched
            for (int v = 0; v < V; ++v) {
                const int s = v % kStages, use = v / kStages, a = v & 1;
                ptx::mbar_wait(full_b + s, use & 1);
                if (v >= 2) ptx::mbar_wait(epi_done + a, ((v >> 1) - 1) & 1);               // accumulator drained
                ptx::tc_fence_after_sync();
                const uint32_t acc = tmem_base + a * kTN;
#pragma unroll
                for (int ka = 0; ka < 2; ++ka) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {                                       // 8 tf32 = 32 bytes per MMA
                        const uint64_t da = ptx::umma_desc_k_sw128(a_addr + ka * kAtomBytes + kk * 32);
                        const uint64_t db = ptx::umma_desc_k_sw128(b_addr + (s * 2 + ka) * kAtomB + kk * 32);
                        ptx::umma_tf32(acc, da, db, idesc, (ka | kk) != 0);
                    }
                }
                ptx::umma_commit(slot_free + s);
                ptx::umma_commit(acc_full + a);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: thread = query row = TMEM lane =====
        const int row = warp * 32 + lane;                              // row in tile
        const int i = m0 + row;                                        // row in cloud
        const float h = 0.5f * __ldg(sqc + (size_t)b * N + min(i, N - 1));
        // A = the same rows as B, with column 63 (-|c_i|^2/2 in global memory) replaced by 1: element (row, 63) of
        // K-atom 1 sits in 16-byte chunk 7 ^ (row & 7) of its 128-byte swizzled line
        ptx::mbar_wait(full_a, 0);
        *reinterpret_cast<float*>(sA + kAtomBytes + row * 128 + ((7 ^ (row & 7)) << 4) + 12) = 1.0f;
        ptx::fence_proxy_async_smem();                                 // generic-proxy write -> visible to the MMA
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_ready);
        float* myq = qk + row;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        int v0 = 0;                                                    // first tile visit of the selection pass
        float thr = -INFINITY;                                         // raw-score threshold (TWO_PASS)
        if (TWO_PASS) {
            // ---- pass 1: chunk maxima -> threshold ----
            const int cw = (N + 32 * kChunkMax - 1) / (32 * kChunkMax);   // 32-column loads per chunk: 24..48 chunks for N >= 768
            float cur = -INFINITY;
            int filled = 0, nch = 0;
            for (int v = 0; v < TV; ++v) {
                ptx::mbar_wait(acc_full + (v & 1), (v >> 1) & 1);
                ptx::tc_fence_after_sync();
#pragma unroll 1
                for (int c0 = 0; c0 < kTN; c0 += 32) {
                    const int jbase = v * kTN + c0;
                    if (jbase >= N) break;                             // warp-uniform
                    float x[32];
                    ptx::tmem_ld_32x32(trow + (v & 1) * kTN + c0, x);
                    if (jbase + 32 > N) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) if (jbase + c >= N) x[c] = -INFINITY;
                    }
                    float m = x[0];
#pragma unroll
                    for (int c = 1; c < 32; ++c) m = fmaxf(m, x[c]);
                    cur = fmaxf(cur, m);
                    if (++filled == cw || jbase + 32 >= N) { myq[nch * kTM] = cur; ++nch; cur = -INFINITY; filled = 0; }
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(epi_done + (v & 1));
            }
            if (nch <= 32) {                                           // warp-uniform (depends on N only)
                float c[32];
#pragma unroll
                for (int u = 0; u < 32; ++u) c[u] = u < nch ? myq[u * kTM] : -INFINITY;
                sort_desc<32>(c);
                thr = c[kThrRank - 1];
            } else {
                float c[64];
#pragma unroll
                for (int u = 0; u < 64; ++u) c[u] = u < nch ? myq[u * kTM] : -INFINITY;
                sort_desc<64>(c);
                thr = c[kThrRank - 1];
            }
            __syncwarp();                                              // the FIFO reuses the chunk-maxima rows
            v0 = TV;
        }
        if (SUPER) {
            // ---- second chance: every candidate whose score reaches the row's threshold goes to the row's list ----
            const float Ti = i < N ? __ldg(thr2 + (size_t)b * N + i) : INFINITY;
            unsigned short* mylist = sup + ((size_t)b * N + min(i, N - 1)) * kSupCap;
            int cnt = 0;
            for (int v = 0; v < TV; ++v) {
                ptx::mbar_wait(acc_full + (v & 1), (v >> 1) & 1);
                ptx::tc_fence_after_sync();
#pragma unroll 1
                for (int c0 = 0; c0 < kTN; c0 += 32) {
                    const int jbase = v * kTN + c0;
                    if (jbase >= N) break;                             // warp-uniform
                    float x[32];
                    ptx::tmem_ld_32x32(trow + (v & 1) * kTN + c0, x);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        if (x[c] - h >= Ti && jbase + c < N) {
                            if (cnt < kSupCap) mylist[cnt] = (unsigned short)(jbase + c);
                            ++cnt;
                        }
                    }
                }
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(epi_done + (v & 1));
            }
            if (i < N) supcnt[(size_t)b * N + i] = cnt;
        } else {
        // ---- selection pass: sorted insert of the scores that pass the threshold(s) ----
        KeyTopK<KL> top;
        top.init();
        float tau = -INFINITY;
        int qlen = 0;
        auto flush = [&]() {
            int rounds = qlen;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(kFull, rounds, o));
            for (int r = 0; r < rounds; ++r) top.insert(r < qlen ? myq[r * kTM] : -INFINITY);
            qlen = 0;
            tau = top.key[KL - 1];
        };
        for (int v = v0; v < V; ++v) {
            const int tile = v - v0;
            ptx::mbar_wait(acc_full + (v & 1), (v >> 1) & 1);
            ptx::tc_fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < kTN; c0 += 32) {
                const unsigned jbase = tile * kTN + c0;
                if (jbase >= (unsigned)N) break;                       // warp-uniform
                if (__any_sync(kFull, qlen > kTcQueue - 32)) flush();
                float x[32];
                ptx::tmem_ld_32x32(trow + (v & 1) * kTN + c0, x);
                if (TWO_PASS && jbase + 32 <= (unsigned)N) {
                    // branch-free: the key always goes to the next FIFO slot, the slot only advances for a score
                    // that reaches the row threshold (~4 % of them); no per-element divergence
                    float* slot = myq + qlen * kTM;
                    const unsigned hi_bits = jbase;                    // multiple of 32: index = jbase | c
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const unsigned bits = (__float_as_uint(x[c] - h) & keep_mask) | hi_bits;
                        *slot = __uint_as_float(bits | (unsigned)c);
                        slot += x[c] >= thr ? kTM : 0;
                    }
                    qlen = (int)(slot - myq) / kTM;
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const unsigned j = jbase + c;
                        if (x[c] >= thr && j < (unsigned)N) {
                            const float key = make_key(x[c] - h, j, keep_mask);
                            if (key > tau) { myq[qlen * kTM] = key; ++qlen; }
                        }
                    }
                }
            }
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(epi_done + (v & 1));
        }
        flush();
        if (i < N) {
            float* out = cand + ((size_t)b * N + i) * KL;
#pragma unroll
            for (int m = 0; m < KL; m += 4)
                *reinterpret_cast<float4*>(out + m) = make_float4(top.key[m], top.key[m + 1], top.key[m + 2], top.key[m + 3]);
            // bound on the key of every score that is NOT in the list: pushed out of it (<= last key) or below the
            // threshold (then shifted <= thr - h, and replacing the index bits moves a key at most to `tb`)
            float bound = top.key[KL - 1];
            if (TWO_PASS) {
                const float ts = thr - h;
                const unsigned bits = __float_as_uint(ts);
                const float tb = __uint_as_float(ts < 0.f ? (bits & keep_mask) : (bits | ~keep_mask));
                bound = fmaxf(bound, tb);
            }
            tau_out[(size_t)b * N + i] = bound;
        }
        }
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) ptx::tmem_dealloc<2 * kTN>(tmem_base);
}

// ---- 3. exact re-rank ------------------------------------------------------------------------------------------
// One warp per query row, lane c <-> candidate c (KL == 32) or candidates c, c+32 (KL == 64).
template <int KL>
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const float* __restrict__ xb /* raw rows xr */, const float* __restrict__ sq, const float* __restrict__ sqc,
                  const float* __restrict__ cand, const float* __restrict__ tau_in, const unsigned* __restrict__ cmax_bits, int D, int N, int k, unsigned keep_mask, int idx_bits,
                  int64_t* __restrict__ idx, float* __restrict__ val, int* __restrict__ fb_list, int* __restrict__ fb_count) {
    constexpr int CPL = KL / 32;                                       // candidates per lane
    constexpr int RS = kKP + 4;                                        // padded smem row stride
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rows = sm + (size_t)warp * (KL * RS + kKP);                 // [KL][RS] candidate rows
    float* qrow = rows + KL * RS;                                      // [64] query row
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const size_t gi = (size_t)b * N + i;
    const float sq_i = __ldg(sq + gi);

    float key[CPL];
    int cj[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        key[u] = __ldg(cand + gi * KL + u * 32 + lane);
        const unsigned jj = __float_as_uint(key[u]) & ~keep_mask;
        cj[u] = (key[u] == -INFINITY || jj >= (unsigned)N) ? -1 : (int)jj;
    }
    // stage: query row (coalesced) and candidate rows, two rows per warp instruction (16 lanes x float4 each)
    *reinterpret_cast<float2*>(qrow + 2 * lane) = __ldg(reinterpret_cast<const float2*>(xb + gi * kKP) + lane);
    const int half = lane >> 4, q4 = lane & 15;
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
#pragma unroll 4
        for (int r = 0; r < 32; r += 2) {
            const int src = __shfl_sync(kFull, cj[u], r + half);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)b * N + src) * kKP) + q4);
            *reinterpret_cast<float4*>(rows + (size_t)(u * 32 + r + half) * RS + 4 * q4) = v;
        }
    }
    __syncwarp();
    // canonical distance: fma chain over d ascending, pd = fmaf(2, dot, -sq_i) - sq_j
    float pd[CPL];
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        const float* cr = rows + (size_t)(u * 32 + lane) * RS;
        float acc = 0.f;
        for (int d = 0; d < D; d += 4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cr + d);
            const float4 a4 = *reinterpret_cast<const float4*>(qrow + d);
            acc = __fmaf_rn(a4.x, c4.x, acc);
            if (d + 1 < D) acc = __fmaf_rn(a4.y, c4.y, acc);
            if (d + 2 < D) acc = __fmaf_rn(a4.z, c4.z, acc);
            if (d + 3 < D) acc = __fmaf_rn(a4.w, c4.w, acc);
        }
        const float sq_j = -2.f * cr[kKP - 1];                         // column 63 holds -|x_j|^2 / 2 exactly
        pd[u] = cj[u] >= 0 ? __fsub_rn(__fmaf_rn(2.f, acc, -sq_i), sq_j) : -INFINITY;
        if (cj[u] < 0) cj[u] = 0x7fffffff;
    }
    // bitonic sort of the KL = 32*CPL elements (position e = u*32 + lane): descending value, ascending index
    // on ties.  Strides >= 32 pair two registers of the same lane, smaller strides pair lanes by shuffle.
    auto before = [](float va, int ja, float vb, int jb) { return va > vb || (va == vb && ja < jb); };
#pragma unroll
    for (int size = 2; size <= KL; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int us = stride >> 5;
#pragma unroll
                for (int u = 0; u < CPL; ++u) {
                    if ((u & us) == 0) {                               // u holds the earlier position of the pair
                        const int w = u | us;
                        const bool desc = (((u << 5) | lane) & size) == 0;
                        const bool first_ok = before(pd[u], cj[u], pd[w], cj[w]);
                        if (first_ok != desc) {
                            const float tv = pd[u]; pd[u] = pd[w]; pd[w] = tv;
                            const int tj = cj[u]; cj[u] = cj[w]; cj[w] = tj;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < CPL; ++u) {
                    const int e = (u << 5) | lane;
                    const float ov = __shfl_xor_sync(kFull, pd[u], stride);
                    const int oj = __shfl_xor_sync(kFull, cj[u], stride);
                    const bool lower = (lane & stride) == 0;           // this lane holds the earlier position
                    const bool desc = (e & size) == 0;
                    const bool mine_first = before(pd[u], cj[u], ov, oj);
                    const bool keep = (lower == desc) ? mine_first : !mine_first;
                    if (!keep) { pd[u] = ov; cj[u] = oj; }
                }
            }
        }
    }
    // safety: best possible exact value of any non-candidate < k-th exact value of the candidates
    const float kth = __shfl_sync(kFull, pd[(k - 1) >> 5], (k - 1) & 31);
    const float tau = __ldg(tau_in + gi);                              // bound on the key of every score outside the list
    const float cmax2 = __uint_as_float(__ldg(cmax_bits + b));         // max_j |x_j - mu|^2
    const float rmax2 = __uint_as_float(__ldg(cmax_bits + gridDim.y + b));   // max_j |x_j|^2
    const float ni = sqrtf(__ldg(sqc + gi)), nmax = sqrtf(cmax2);
    const float rsum = sqrtf(sq_i) + sqrtf(rmax2);
    // |tensor-core score - exact pd/2| <= TF32 rounding of both operands (round-to-nearest in the pack kernel: 2^-11 each, 2^-10
    // relative per product, summed with Cauchy-Schwarz) and of the norm column (2^-11 of |c_j|^2/2), + fp32 rounding of x - mu, of both accumulations
    // and of the shift (2^-16 and 2^-20 terms are generous), all on the CENTRED data; plus the rounding error
    // of the canonical fp32 arithmetic itself, which works on the RAW data: D-step fma chains for the dot and
    // both norms and two final roundings, |pd_canonical - pd_true| / 2 <= (D + 8) 2^-25 (|x_i| + max|x_j|)^2
    // (taken twice).  A cloud far from the origin relative to its extent fails this test honestly: its
    // canonical order is decided by fp32 rounding, which only the exact kernels reproduce.
    const float eps = 1.01f * (ni * nmax * (1.f / 1024.f + 1.f / 65536.f) + cmax2 * (1.f / 4096.f) +
                               (ni + nmax) * (ni + nmax) * (1.f / 1048576.f) +
                               (float)(D + 8) * rsum * rsum * (1.f / 16777216.f));
    const float quant = fabsf(tau) * exp2f((float)(idx_bits - 22));
    const bool safe = (tau == -INFINITY) ? (kth > -INFINITY) : (0.5f * kth > tau + quant + eps);
    if (!safe) {
        if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = (int)gi;
        return;
    }
#pragma unroll
    for (int u = 0; u < CPL; ++u) {
        const int r = u * 32 + lane;
        if (r < k) {
            idx[gi * k + r] = cj[u];
            if (val) val[gi * k + r] = pd[u];
        }
    }
}

// ---- 3b. exact re-rank, lists of 32 ---------------------------------------------------------------------------
// The list arrives sorted by approximate key, and the TF32 error is about one neighbour spacing, so the exact
// top k is almost always inside the first k + kExtra entries: only those rows are gathered (the gather is the
// cost of this kernel: 256 bytes per candidate from L2).  The entries not evaluated are then bounded by the key of
// the best one of them; if that bound is not safe the remaining entries are evaluated too, and only if the bound
// of everything outside the list still fails does the row go to the exact redo.
constexpr int kExtra = 6;

__global__ void __launch_bounds__(256)
knn_rerank32_kernel(const float* __restrict__ xr, const float* __restrict__ sq, const float* __restrict__ sqc,
                    const float* __restrict__ cand, const float* __restrict__ tau_in, const unsigned* __restrict__ cmax_bits, int D, int N,
                    int k, unsigned keep_mask, int idx_bits, int64_t* __restrict__ idx, float* __restrict__ val,
                    int* __restrict__ fb_list, int* __restrict__ fb_count, float* __restrict__ thr2) {
    constexpr int RS = kKP + 4;                                        // padded smem row stride
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rows = sm + (size_t)warp * (32 * RS + kKP);                 // [32][RS] candidate rows
    float* qrow = rows + 32 * RS;                                      // [64] query row
    const int b = blockIdx.y;
    const int i = blockIdx.x * 8 + warp;
    if (i >= N) return;
    const size_t gi = (size_t)b * N + i;
    const float sq_i = __ldg(sq + gi);
    const float key = __ldg(cand + gi * 32 + lane);
    const unsigned jj = __float_as_uint(key) & ~keep_mask;
    const int cj0 = (key == -INFINITY || jj >= (unsigned)N) ? -1 : (int)jj;
    *reinterpret_cast<float2*>(qrow + 2 * lane) = __ldg(reinterpret_cast<const float2*>(xr + gi * kKP) + lane);
    const int half = lane >> 4, q4 = lane & 15;

    float pd0 = -INFINITY;                                             // exact value of this lane's candidate, once evaluated
    const float4* xr4 = reinterpret_cast<const float4*>(xr) + (size_t)b * N * (kKP / 4) + q4;   // this lane's 16-byte column of the cloud
    float* my_dst = rows + 4 * q4;
    auto evaluate = [&](int lo, int hi) {                              // candidates [lo, hi): gather rows, canonical chains
        for (int r0 = lo; r0 < hi; r0 += 16) {                         // two rows per warp instruction (16 lanes x float4
            float4 v[8];                                               // each), eight such loads in flight
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r0 + 2 * u + half;
                const int src = __shfl_sync(kFull, cj0, rr & 31);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rr < hi && src >= 0) v[u] = __ldg(xr4 + src * (kKP / 4));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r0 + 2 * u + half;
                if (rr < hi) *reinterpret_cast<float4*>(my_dst + rr * RS) = v[u];
            }
        }
        __syncwarp();
        if (lane >= lo && lane < hi && cj0 >= 0) {
            const float* cr = rows + lane * RS;
            float acc = 0.f;
            int d = 0;
            for (; d + 4 <= D; d += 4) {                               // fma chain over d ascending
                const float4 c4 = *reinterpret_cast<const float4*>(cr + d);
                const float4 a4 = *reinterpret_cast<const float4*>(qrow + d);
                acc = __fmaf_rn(a4.x, c4.x, acc);
                acc = __fmaf_rn(a4.y, c4.y, acc);
                acc = __fmaf_rn(a4.z, c4.z, acc);
                acc = __fmaf_rn(a4.w, c4.w, acc);
            }
            for (; d < D; ++d) acc = __fmaf_rn(qrow[d], cr[d], acc);
            const float sq_j = -2.f * cr[kKP - 1];                     // column 63 holds -|x_j|^2 / 2 exactly
            pd0 = __fsub_rn(__fmaf_rn(2.f, acc, -sq_i), sq_j);
        }
        __syncwarp();
    };
    // sort by one 64-bit key: order-preserving image of the value (high word) and the complemented index (low word),
    // so "larger key first" is "larger value first, ties -> lower index"
    unsigned long long skey;
    auto sort32 = [&]() {
        const unsigned fb = __float_as_uint(pd0);
        const unsigned ord = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
        const unsigned jc = (cj0 >= 0 && pd0 > -INFINITY) ? ~(unsigned)cj0 : 0u;
        skey = ((unsigned long long)ord << 32) | jc;
#pragma unroll
        for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                const unsigned long long other = __shfl_xor_sync(kFull, skey, stride);
                const bool want_max = ((lane & stride) == 0) == ((lane & size) == 0);
                skey = want_max ? (skey > other ? skey : other) : (skey < other ? skey : other);
            }
        }
    };
    auto sorted_value = [&]() {
        const unsigned ord = (unsigned)(skey >> 32);
        return __uint_as_float((ord & 0x80000000u) ? (ord & 0x7fffffffu) : ~ord);
    };
    const float tau = __ldg(tau_in + gi);                              // bound on the key of every score outside the list
    const float cmax2 = __uint_as_float(__ldg(cmax_bits + b));         // max_j |x_j - mu|^2
    const float rmax2 = __uint_as_float(__ldg(cmax_bits + gridDim.y + b));   // max_j |x_j|^2
    const float ni = sqrtf(__ldg(sqc + gi)), nmax = sqrtf(cmax2);
    const float rsum = sqrtf(sq_i) + sqrtf(rmax2);
    // error bound: see knn_rerank_kernel
    const float eps = 1.01f * (ni * nmax * (1.f / 1024.f + 1.f / 65536.f) + cmax2 * (1.f / 4096.f) +
                               (ni + nmax) * (ni + nmax) * (1.f / 1048576.f) +
                               (float)(D + 8) * rsum * rsum * (1.f / 16777216.f));
    const float qscale = exp2f((float)(idx_bits - 22));
    auto safe_against = [&](float bound) {                             // bound: largest key of anything not evaluated
        const float kth = __shfl_sync(kFull, sorted_value(), k - 1);
        return bound == -INFINITY ? kth > -INFINITY : 0.5f * kth > bound + fabsf(bound) * qscale + eps;
    };

    const int ne = min(32, k + kExtra);
    evaluate(0, ne);
    sort32();
    bool ok = true;
    if (ne < 32) {
        const float next_key = __shfl_sync(kFull, key, ne);            // keys are sorted: the best entry not evaluated
        ok = safe_against(fmaxf(next_key, tau));
        if (!ok) {                                                     // warp-uniform
            evaluate(ne, 32);
            sort32();
            ok = safe_against(tau);
        }
    } else {
        ok = safe_against(tau);
    }
    if (!ok) {
        // second chance: anything whose tensor-core score is below (k-th exact value so far) / 2 - eps cannot be among the k nearest
        const float kth = __shfl_sync(kFull, sorted_value(), k - 1);
        if (lane == 0) {
            fb_list[atomicAdd(fb_count, 1)] = (int)gi;
            thr2[gi] = 0.5f * kth - eps;                              // -inf when fewer than k candidates were valid: collect everything
        }
        return;
    }
    if (lane == 0) thr2[gi] = INFINITY;
    if (lane < k) {
        idx[gi * k + lane] = (int64_t)(~(unsigned)skey);
        if (val) val[gi * k + lane] = sorted_value();
    }
}

// ---- 3b. second chance: exact evaluation of the superset lists ----------------------------------------------------------
// One warp per unproven row (grid-stride over fb_list).  The row's list (knn_tc_kernel MODE 2) holds every candidate that can
// still be among its k nearest; they are evaluated 32 at a time with the canonical fp32 chain (rows gathered from L2 like in
// knn_rerank32_kernel), each batch is sorted and merged into the running best 32 (larger value first, ties -> lower index), and
// the first k are the answer -- exact, because the list is a superset of the true k nearest.  A list that overflowed its
// capacity (or holds fewer than k candidates) sends the row on to the exact redo kernels.
__global__ void __launch_bounds__(256)
knn_rerank_super_kernel(const float* __restrict__ xr, const float* __restrict__ sq, const unsigned short* __restrict__ sup,
                        const int* __restrict__ supcnt, int D, int N, int k, const int* __restrict__ fb_list,
                        const int* __restrict__ fb_count, int64_t* __restrict__ idx, float* __restrict__ val, int* __restrict__ fb2_list,
                        int* __restrict__ fb2_count) {
    constexpr int RS = kKP + 4;
    extern __shared__ __align__(16) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rows = sm + (size_t)warp * (32 * RS + kKP);                 // [32][RS] candidate rows
    float* qrow = rows + 32 * RS;                                      // [64] query row
    const int count = *fb_count;
    const int half = lane >> 4, q4 = lane & 15;
    float* my_dst = rows + 4 * q4;
    for (int f = blockIdx.x * 8 + warp; f < count; f += gridDim.x * 8) {
        const size_t gi = (size_t)fb_list[f];
        const int b = (int)(gi / N);
        const int c = supcnt[gi];
        if (c > kSupCap || c < k) {                                    // warp-uniform
            if (lane == 0) fb2_list[atomicAdd(fb2_count, 1)] = (int)gi;
            continue;
        }
        const float sq_i = __ldg(sq + gi);
        __syncwarp();
        *reinterpret_cast<float2*>(qrow + 2 * lane) = __ldg(reinterpret_cast<const float2*>(xr + gi * kKP) + lane);
        const float4* xr4 = reinterpret_cast<const float4*>(xr) + (size_t)b * N * (kKP / 4) + q4;
        unsigned long long best = 0ull;                                // running best 32, sorted descending across the lanes
        for (int r0 = 0; r0 < c; r0 += 32) {
            const int cj0 = r0 + lane < c ? (int)sup[gi * kSupCap + r0 + lane] : -1;
            const int hi = min(32, c - r0);
            for (int rr0 = 0; rr0 < hi; rr0 += 16) {                   // gather: two rows per warp instruction, eight loads in flight
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int rr = rr0 + 2 * u + half;
                    const int src = __shfl_sync(kFull, cj0, rr & 31);
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rr < hi && src >= 0) v[u] = __ldg(xr4 + src * (kKP / 4));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int rr = rr0 + 2 * u + half;
                    if (rr < hi) *reinterpret_cast<float4*>(my_dst + rr * RS) = v[u];
                }
            }
            __syncwarp();
            float pd0 = -INFINITY;
            if (cj0 >= 0) {
                const float* cr = rows + lane * RS;
                float acc = 0.f;
                int d = 0;
                for (; d + 4 <= D; d += 4) {                           // fma chain over d ascending (canonical)
                    const float4 c4 = *reinterpret_cast<const float4*>(cr + d);
                    const float4 a4 = *reinterpret_cast<const float4*>(qrow + d);
                    acc = __fmaf_rn(a4.x, c4.x, acc);
                    acc = __fmaf_rn(a4.y, c4.y, acc);
                    acc = __fmaf_rn(a4.z, c4.z, acc);
                    acc = __fmaf_rn(a4.w, c4.w, acc);
                }
                for (; d < D; ++d) acc = __fmaf_rn(qrow[d], cr[d], acc);
                const float sq_j = -2.f * cr[kKP - 1];
                pd0 = __fsub_rn(__fmaf_rn(2.f, acc, -sq_i), sq_j);
            }
            __syncwarp();
            const unsigned fb = __float_as_uint(pd0);
            const unsigned ord = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
            unsigned long long skey = cj0 >= 0 ? (((unsigned long long)ord << 32) | ~(unsigned)cj0) : 0ull;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {               // sort the batch, descending
#pragma unroll
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(kFull, skey, stride);
                    const bool want_max = ((lane & stride) == 0) == ((lane & size) == 0);
                    skey = want_max ? (skey > other ? skey : other) : (skey < other ? skey : other);
                }
            }
            // best 32 of (running, batch): max(best[l], batch[31 - l]) is bitonic and holds them; one bitonic merge sorts it
            const unsigned long long rev = __shfl_sync(kFull, skey, 31 - lane);
            best = best > rev ? best : rev;
#pragma unroll
            for (int stride = 16; stride > 0; stride >>= 1) {
                const unsigned long long other = __shfl_xor_sync(kFull, best, stride);
                const bool want_max = (lane & stride) == 0;
                best = want_max ? (best > other ? best : other) : (best < other ? best : other);
            }
        }
        if (lane < k) {
            const unsigned ordv = (unsigned)(best >> 32);
            idx[gi * k + lane] = (int64_t)(~(unsigned)best);
            if (val) val[gi * k + lane] = __uint_as_float((ordv & 0x80000000u) ? (ordv & 0x7fffffffu) : ~ordv);
        }
    }
}

// ---- 4. exact redo of flagged rows ------------------------------------------------------------------------------
// One CTA per flagged row (grid-stride over the device-side list): all N canonical distances into shared memory
// (candidate features read coalesced, the fma chain unrolled so its loads overlap), then k rounds of a block-wide
// arg-max with the canonical order (larger value first, ties -> lower index).  Runs only when the list is short
// (<= gate_max rows); a cloud that fails wholesale -- far from the origin relative to its extent, so that fp32
// rounding decides the canonical order -- is instead redone by the all-FFMA kernels (knn.cu), gated the other way.
constexpr int kFbThreads = 256;

__global__ void __launch_bounds__(kFbThreads)
knn_fallback_kernel(const float* __restrict__ x, const float* __restrict__ sq, int D, int N, int k,
                    const int* __restrict__ fb_list, const int* __restrict__ fb_count, int gate_max,
                    int64_t* __restrict__ idx, float* __restrict__ val) {
    extern __shared__ __align__(16) float fsm[];
    float* qs = fsm;                       // [64] query features
    float* pd = fsm + 64;                  // [N]
    __shared__ float wv[kFbThreads / 32];
    __shared__ int wj[kFbThreads / 32];
    const int count = *fb_count;
    if (count > gate_max) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int f = blockIdx.x; f < count; f += gridDim.x) {
        const int gi = fb_list[f];
        const int b = gi / N, i = gi - b * N;
        const float* xb = x + (size_t)b * D * N;
        const float* sqb = sq + (size_t)b * N;
        __syncthreads();                                               // previous row fully written out
        if (threadIdx.x < 64) qs[threadIdx.x] = threadIdx.x < D ? __ldg(xb + (size_t)threadIdx.x * N + i) : 0.f;
        __syncthreads();
        const float nsq_i = -__ldg(sqb + i);
        for (int j = threadIdx.x; j < N; j += kFbThreads) {
            float acc = 0.f;
            int d = 0;
            for (; d + 8 <= D; d += 8) {
                float c[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) c[u] = __ldg(xb + (size_t)(d + u) * N + j);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc = __fmaf_rn(qs[d + u], c[u], acc);
            }
            for (; d < D; ++d) acc = __fmaf_rn(qs[d], __ldg(xb + (size_t)d * N + j), acc);
            pd[j] = __fsub_rn(__fmaf_rn(2.f, acc, nsq_i), __ldg(sqb + j));
        }
        __syncthreads();
        for (int r = 0; r < k; ++r) {
            float bv = -INFINITY;
            int bj = 0x7fffffff;
            for (int j = threadIdx.x; j < N; j += kFbThreads) {        // ascending j: strict '>' keeps the lower index
                const float v = pd[j];
                if (v > bv) { bv = v; bj = j; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, bv, o);
                const int oj = __shfl_xor_sync(kFull, bj, o);
                if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
            }
            if (lane == 0) { wv[warp] = bv; wj[warp] = bj; }
            __syncthreads();
            if (warp == 0) {
                bv = lane < kFbThreads / 32 ? wv[lane] : -INFINITY;
                bj = lane < kFbThreads / 32 ? wj[lane] : 0x7fffffff;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(kFull, bv, o);
                    const int oj = __shfl_xor_sync(kFull, bj, o);
                    if (ov > bv || (ov == bv && oj < bj)) { bv = ov; bj = oj; }
                }
                if (lane == 0) {
                    const int jj = bj < N ? bj : i;
                    idx[(size_t)gi * k + r] = jj;
                    if (val) val[(size_t)gi * k + r] = bv;
                    if (bj < N) pd[bj] = -INFINITY;
                }
            }
            __syncthreads();
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x 64 fp32, row-major; box = 32 fp32 (128 B) x 128 rows, 128B swizzle
static bool make_row_map(CUtensorMap* map, const float* base, size_t rows, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t gdim[2] = {kKP, rows};
    const cuuint64_t gstride[1] = {kKP * sizeof(float)};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct TcLayout {
    float *xc, *xr, *sq, *sqc, *cand, *tau, *mu;
    unsigned* cmax;
    int *fb_count, *fb_list;
    float* thr2;                   // second chance: per-row score threshold (+inf: row proven)
    unsigned short* sup;           // [rows][kSupCap] superset lists
    int *supcnt, *fb2_list, *fb2_count;
    int n_zero;
    size_t bytes;
};

static TcLayout tc_layout(void* ws, int B, int D, int N, int KL) {
    TcLayout L;
    char* p = static_cast<char*>(ws);
    size_t off = 0;
    const size_t rows = (size_t)B * N;
    L.xc = reinterpret_cast<float*>(p + off);   off += align_up(rows * kKP * sizeof(float), 1024);
    L.xr = reinterpret_cast<float*>(p + off);   off += align_up(rows * kKP * sizeof(float), 1024);
    L.sq = reinterpret_cast<float*>(p + off);   off += align_up(rows * sizeof(float), 256);
    L.sqc = reinterpret_cast<float*>(p + off);  off += align_up(rows * sizeof(float), 256);
    L.cand = reinterpret_cast<float*>(p + off); off += align_up(rows * KL * sizeof(float), 256);
    L.tau = reinterpret_cast<float*>(p + off);  off += align_up(rows * sizeof(float), 256);
    L.mu = reinterpret_cast<float*>(p + off);   off += align_up((size_t)B * D * sizeof(float), 256);
    L.cmax = reinterpret_cast<unsigned*>(p + off);  off += align_up((size_t)(2 * B + 2) * sizeof(unsigned), 256);
    L.fb_count = reinterpret_cast<int*>(L.cmax + 2 * B);              // [B] centred max, [B] raw max, two counters
    L.fb2_count = L.fb_count + 1;
    L.n_zero = 2 * B + 2;
    L.fb_list = reinterpret_cast<int*>(p + off); off += align_up(rows * sizeof(int), 256);
    L.fb2_list = L.fb_list;
    L.thr2 = nullptr; L.sup = nullptr; L.supcnt = nullptr;
    if (KL == 32) {                                                    // second-chance buffers (lists of 32 only)
        L.fb2_list = reinterpret_cast<int*>(p + off);            off += align_up(rows * sizeof(int), 256);
        L.thr2 = reinterpret_cast<float*>(p + off);              off += align_up(rows * sizeof(float), 256);
        L.supcnt = reinterpret_cast<int*>(p + off);              off += align_up(rows * sizeof(int), 256);
        L.sup = reinterpret_cast<unsigned short*>(p + off);      off += align_up(rows * kSupCap * sizeof(unsigned short), 256);
    }
    L.bytes = off;
    return L;
}

bool knn_tc_applicable(int D, int N, int k) { return D >= 16 && D <= kKP - 1 && N >= 128 && N <= 16384 && k <= 48; }

static int tc_list_len(int k) { return k <= 24 ? 32 : 64; }

size_t knn_tc_workspace_bytes(int B, int D, int N, int k) {
    return tc_layout(nullptr, B, D, N, tc_list_len(k)).bytes;
}

template <int KL>
static int run_tc(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, const TcLayout& L, cudaStream_t st) {
    const size_t rows = (size_t)B * N;
    int idx_bits = 1;
    while ((1 << idx_bits) < N) ++idx_bits;
    const unsigned keep_mask = ~((1u << idx_bits) - 1u);
    knn_mean_kernel<<<(B * D + 7) / 8, 256, 0, st>>>(x, B * D, N, L.mu, L.cmax, L.n_zero);   // also clears cmax / fb_count
    int rc = check_launch("knn_mean_kernel");
    if (rc) return rc;
    knn_pack_kernel<<<dim3((N + 31) / 32, B), 256, 0, st>>>(x, L.mu, D, N, L.xc, L.xr, L.sq, L.sqc, L.cmax);
    rc = check_launch("knn_pack_kernel");
    if (rc) return rc;
    CUtensorMap map_a, map_b;                                          // same array: 128-row boxes for A, 64-row boxes for B
    if (!make_row_map(&map_a, L.xc, rows, kTM) || !make_row_map(&map_b, L.xc, rows, kTN))
        return fail(HPCS_ERR_CUDA, "knn_tc: cuTensorMapEncodeTiled failed");
    {
        const size_t smem = 2 * (size_t)kAtomBytes + (size_t)kStages * 2 * kAtomB + (size_t)kTcQueue * kTM * sizeof(float) +
                            (6 + 2 * kStages) * sizeof(uint64_t) + 16 + 1024;
        const dim3 grid((N + kTM - 1) / kTM, B);
        if (KL == 32 && N >= 32 * kThrRank) {
            auto kern = knn_tc_kernel<KL, 1>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, kTcThreads, smem, st>>>(map_a, map_b, L.sqc, N, keep_mask, L.cand, L.tau, nullptr, nullptr, nullptr, nullptr);
        } else {
            auto kern = knn_tc_kernel<KL, 0>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, kTcThreads, smem, st>>>(map_a, map_b, L.sqc, N, keep_mask, L.cand, L.tau, nullptr, nullptr, nullptr, nullptr);
        }
        rc = check_launch("knn_tc_kernel");
        if (rc) return rc;
    }
    {
        const size_t smem = 8 * ((size_t)KL * (kKP + 4) + kKP) * sizeof(float);
        if (KL == 32) {
            cudaFuncSetAttribute(knn_rerank32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            knn_rerank32_kernel<<<dim3((N + 7) / 8, B), 256, smem, st>>>(L.xr, L.sq, L.sqc, L.cand, L.tau, L.cmax, D, N, k, keep_mask, idx_bits,
                                                                         idx, val, L.fb_list, L.fb_count, L.thr2);
        } else {
            auto kern = knn_rerank_kernel<KL>;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<dim3((N + 7) / 8, B), 256, smem, st>>>(L.xr, L.sq, L.sqc, L.cand, L.tau, L.cmax, D, N, k, keep_mask, idx_bits, idx, val,
                                                          L.fb_list, L.fb_count);
        }
        rc = check_launch("knn_rerank_kernel");
        if (rc) return rc;
    }
    // rows whose list of 32 could not be proven get a second chance (lists of 32 only): one more Gram pass that collects, per
    // row, every candidate that can still be among its k nearest, and an exact evaluation of those lists; both launches return
    // at once when every row was proven.  Tight clusters (many neighbours inside the TF32 error bound) end here, not below.
    const int* redo_list = L.fb_list;
    const int* redo_count = L.fb_count;
    if (KL == 32) {
        const size_t smem = 2 * (size_t)kAtomBytes + (size_t)kStages * 2 * kAtomB + (size_t)kTcQueue * kTM * sizeof(float) +
                            (6 + 2 * kStages) * sizeof(uint64_t) + 16 + 1024;
        auto kern = knn_tc_kernel<KL, 2>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<dim3((N + kTM - 1) / kTM, B), kTcThreads, smem, st>>>(map_a, map_b, L.sqc, N, keep_mask, nullptr, nullptr, L.thr2, L.sup, L.supcnt,
                                                                      L.fb_count);
        if ((rc = check_launch("knn_tc_kernel(second chance)"))) return rc;
        const size_t smem_r = 8 * ((size_t)32 * (kKP + 4) + kKP) * sizeof(float);
        cudaFuncSetAttribute(knn_rerank_super_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        knn_rerank_super_kernel<<<2 * sm_count(), 256, smem_r, st>>>(L.xr, L.sq, L.sup, L.supcnt, D, N, k, L.fb_list, L.fb_count, idx, val,
                                                                     L.fb2_list, L.fb2_count);
        if ((rc = check_launch("knn_rerank_super_kernel"))) return rc;
        redo_list = L.fb2_list;
        redo_count = L.fb2_count;
    }
    // rows still open: a short list is redone row by row; a long one (> 1/16 of all rows) means the canonical order of whole
    // clouds is decided by fp32 rounding, and the all-FFMA path redoes everything instead
    const int gate = (int)(rows / 16);
    {
        const size_t smem = (64 + (size_t)N) * sizeof(float);
        cudaFuncSetAttribute(knn_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        knn_fallback_kernel<<<4 * sm_count(), kFbThreads, smem, st>>>(x, L.sq, D, N, k, redo_list, redo_count, gate, idx, val);
        rc = check_launch("knn_fallback_kernel");
        if (rc) return rc;
    }
    return knn_ffma_gated(x, L.sq, B, D, N, k, idx, val, redo_count, gate, st);
}

int knn_tc_fallback_rows(const void* ws, int B, int D, int N, int k, cudaStream_t st, int* out_host) {
    const TcLayout L = tc_layout(const_cast<void*>(ws), B, D, N, tc_list_len(k));
    // out_host[0]: rows that went to the exact redo kernels; out_host[1] (lists of 32): rows that took the second chance
    int both[2] = {0, 0};
    cudaError_t e = cudaMemcpyAsync(both, L.fb_count, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
        const bool second = tc_list_len(k) == 32;
        out_host[0] = second ? both[1] : both[0];
        out_host[1] = second ? both[0] : 0;
        return HPCS_OK;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(HPCS_ERR_CUDA, "knn_tc_fallback_rows: %s", cudaGetErrorString(e));
    return HPCS_OK;
}

int knn_tc_run(const float* x, int B, int D, int N, int k, int64_t* idx, float* val, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int KL = tc_list_len(k);
    const TcLayout L = tc_layout(ws, B, D, N, KL);
    if (ws_bytes < L.bytes) return fail(HPCS_ERR_WORKSPACE, "knn: workspace too small");
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) return fail(HPCS_ERR_ARG, "knn: workspace must be 256-byte aligned");
    return KL == 32 ? run_tc<32>(x, B, D, N, k, idx, val, L, st) : run_tc<64>(x, B, D, N, k, idx, val, L, st);
}

}  // namespace hpcs
