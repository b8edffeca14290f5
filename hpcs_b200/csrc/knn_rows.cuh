// Per-thread exact top-K selection used by the row-parallel kNN kernels.
//
// One thread owns one query row.  Its running top-K is a sorted register list (value descending,
// equal values in arrival order = ascending candidate index).  Candidates that beat the current
// K-th value are first appended to a small per-thread FIFO in shared memory; when some lane of the
// warp is about to run out of FIFO space the whole warp flushes: round r inserts every lane's r-th
// pending item with a branch-free, fully parallel insertion step (K independent compares, 2K
// selects), so the warp never diverges and a round costs the same whether 1 or 32 lanes have work.
#pragma once
#include "common.cuh"

namespace hpcs {

template <int K>
struct RowTopK {
    float val[K];
    int idx[K];

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int m = 0; m < K; ++m) { val[m] = -INFINITY; idx[m] = 0x7fffffff; }
    }
    __device__ __forceinline__ float tau() const { return val[K - 1]; }

    // insert (v, j) keeping order; v = -inf is a no-op.  Strict '>' keeps earlier equal values ahead.
    __device__ __forceinline__ void insert(float v, int j) {
        bool p[K];
#pragma unroll
        for (int m = 0; m < K; ++m) p[m] = v > val[m];
#pragma unroll
        for (int m = K - 1; m >= 1; --m) {
            val[m] = p[m] ? (p[m - 1] ? val[m - 1] : v) : val[m];
            idx[m] = p[m] ? (p[m - 1] ? idx[m - 1] : j) : idx[m];
        }
        val[0] = p[0] ? v : val[0];
        idx[0] = p[0] ? j : idx[0];
    }
};

// FIFO of pending candidates, column `tid` of qv/qj ([QCAP][threads] in shared memory).
template <int K, int QCAP>
struct RowSelector {
    RowTopK<K> top;
    float tau;
    int qlen;
    float* qv;
    int* qj;
    int stride;

    __device__ __forceinline__ void init(float* qv_, int* qj_, int stride_) {
        top.init();
        tau = -INFINITY;
        qlen = 0;
        qv = qv_; qj = qj_; stride = stride_;
    }
    __device__ __forceinline__ void offer(float v, int j) {
        if (v > tau) {
            qv[qlen * stride] = v;
            qj[qlen * stride] = j;
            ++qlen;
        }
    }
    // warp-collective: flush if any lane could overflow during the next `batch` offers
    __device__ __forceinline__ void maybe_flush(int batch) {
        if (__any_sync(kFull, qlen > QCAP - batch)) flush();
    }
    __device__ __forceinline__ void flush() {
        int rounds = qlen;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(kFull, rounds, o));
        for (int r = 0; r < rounds; ++r) {
            const bool has = r < qlen;
            const float v = has ? qv[r * stride] : -INFINITY;
            const int j = has ? qj[r * stride] : 0;
            top.insert(v, j);
        }
        qlen = 0;
        tau = top.tau();
    }
};

}  // namespace hpcs
