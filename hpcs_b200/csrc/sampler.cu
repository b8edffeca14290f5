// Class-balanced random triplets on the device.
//
// Same structure as get_balanced_random_triplet_indices (hpcs/miner/loss_and_miner_utils.py:7-75): labels in
// ascending order; every member of label l is an anchor k_l times back to back (members in ascending index order);
// the positive is uniform over the OTHER members of l, the negative uniform over the non-members.  The anchor array
// is therefore identical to the reference sampler's; positives and negatives follow the same distribution but come
// from a counter-based Philox4x32-10 stream keyed by (seed, triplet pair number) instead of torch's CPU generator, so a
// common seed does not reproduce the reference's draws (the host sampler in hpcs_b200/loss.py does, and stays the
// default).  What this buys: 12 bytes per triplet (19.7 MB per step at the bench shape) never cross PCIe, and the
// per-label Python loop with its n_l x n_l masks is gone; the kernel is a pure 12 B/triplet HBM write.
//
// Inputs prepared by the host mirror (tiny): order[n] = stable argsort of labels (members of a label contiguous,
// ascending), and per valid label the segment table seg[4][L] = {start in order, members, k_l, first triplet}.
#include "common.cuh"

namespace hpcs {

struct Philox {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox{c0, c1, c2, c3};
}

constexpr int kMaxSeg = 1024;

// `state` (optional, device): {seed, step, blocks done}.  When given, the Philox key is (seed, step) read from the device and
// the last block to finish advances `step`, so a CUDA graph that captured this launch draws fresh triplets on every replay.
__global__ void __launch_bounds__(256)
triplet_sample_kernel(const int* __restrict__ order, int n, const int64_t* __restrict__ seg, int L, int64_t T0, uint64_t seed,
                      unsigned long long* __restrict__ state, int* __restrict__ a, int* __restrict__ p, int* __restrict__ ng) {
    uint32_t c2 = 0u, c3 = 0u;
    if (state) {
        seed = state[0];
        const unsigned long long step = state[1];
        c2 = (uint32_t)step;
        c3 = (uint32_t)(step >> 32);
    }
    __shared__ int64_t tstart[kMaxSeg + 1];
    __shared__ int start[kMaxSeg], members[kMaxSeg], reps[kMaxSeg];
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        start[l] = (int)seg[l];
        members[l] = (int)seg[L + l];
        reps[l] = (int)seg[2 * L + l];
        tstart[l] = seg[3 * L + l];
    }
    if (threadIdx.x == 0) tstart[L] = T0;
    __syncthreads();
    // one thread draws two consecutive triplets from one Philox block (4 words); segment offsets are 32-bit whenever a
    // label's triplets number < 2^31 (always, in practice), which keeps the division off the 64-bit slow path
    for (int64_t t0 = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x); t0 < T0; t0 += 2 * (int64_t)gridDim.x * blockDim.x) {
        const Philox r = philox4x32_10((uint32_t)t0, (uint32_t)(t0 >> 32), c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t t = t0 + h;
            if (t >= T0) break;
            int lo = 0, hi = L;                                  // segment with tstart[lo] <= t < tstart[lo + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (tstart[mid] <= t) lo = mid; else hi = mid;
            }
            const int m = members[lo], s0 = start[lo];
            const int64_t off = t - tstart[lo];
            const int slot = off < 0x7fffffffLL ? (int)((unsigned)off / (unsigned)reps[lo]) : (int)(off / reps[lo]);   // anchor's position among its label's members
            int ps = (int)__umulhi(h ? r.z : r.x, (uint32_t)(m - 1));    // uniform in [0, m-1)
            ps += ps >= slot ? 1 : 0;                                     // skip the anchor itself
            const int nd = (int)__umulhi(h ? r.w : r.y, (uint32_t)(n - m));   // uniform over the n - m non-members
            a[t] = order[s0 + slot];
            p[t] = order[s0 + ps];
            ng[t] = order[nd < s0 ? nd : nd + m];                         // order[] without the segment [s0, s0 + m)
        }
    }
    if (state) {                                                          // every block has read state[0..1] by now
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&state[2], 1ull) == (unsigned long long)gridDim.x - 1) {
                state[2] = 0ull;
                state[1] += 1ull;
            }
        }
    }
}

}  // namespace hpcs

static int triplet_sample_launch(const int* order, int64_t n, const int64_t* seg, int L, int64_t T0, uint64_t seed,
                                 unsigned long long* state, int* a, int* p, int* ng, void* stream) {
    using namespace hpcs;
    if (T0 == 0) return HPCS_OK;
    if (!order || !seg || !a || !p || !ng) return fail(HPCS_ERR_ARG, "triplet_sample: null pointer");
    if (n <= 1 || n > 0x7fffffff || L <= 0 || L > kMaxSeg || T0 < 0) return fail(HPCS_ERR_ARG, "triplet_sample: bad arguments n=%lld L=%d T0=%lld", (long long)n, L, (long long)T0);
    int64_t blocks = ((T0 + 1) / 2 + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    triplet_sample_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(order, (int)n, seg, L, T0, seed, state, a, p, ng);
    return check_launch("triplet_sample_kernel");
}

extern "C" int hpcs_triplet_sample_i32(const int* order, int64_t n, const int64_t* seg, int L, int64_t T0, uint64_t seed,
                                       int* a, int* p, int* ng, void* stream) {
    return triplet_sample_launch(order, n, seg, L, T0, seed, nullptr, a, p, ng, stream);
}

extern "C" int hpcs_triplet_sample_state_i32(const int* order, int64_t n, const int64_t* seg, int L, int64_t T0,
                                             uint64_t* state, int* a, int* p, int* ng, void* stream) {
    if (!state) return hpcs::fail(HPCS_ERR_ARG, "triplet_sample_state: null state");
    return triplet_sample_launch(order, n, seg, L, T0, 0, reinterpret_cast<unsigned long long*>(state), a, p, ng, stream);
}
