// Fused EdgeConv layer of the VN-DGCNN backbone:
//
//     get_graph_feature  ->  VNLinearLeakyReLU  [-> VNLinearLeakyReLU]  ->  mean over the k neighbours
//     hpcs/nn/dgcnn/vn_dgcnn_partseg.py:65-68,70-73,75-77; hpcs/nn/dgcnn/utils/vn_layers.py:48-77,112-132,152-153
//
// The reference materialises the edge tensor [B,2C,3,N,k] (330 MB per 63-d layer at the bench shape) and then runs
// ~30 elementwise / transpose / matmul passes over it per VN layer.  Here that tensor never exists:
//
//  * The first Linear of a layer acts on cat(x_j - x_i, x_i), so it splits into two per-POINT maps
//        P(i,j) = Wa (x_j - x_i) + Wb x_i = U[j] + V[i],   U = Wa x,  V = (Wb - Wa) x
//    (same for the direction map).  vn_point_linear_kernel writes U and V once per point as 512-byte rows
//    [feat 63 | 0 | dir 63 | 0]; an edge is then one row gather (L2 resident: 16.8 MB per array) plus 126 adds.
//  * Everything per edge -- vector norm, the BatchNorm affine on the norm, the direction-projected leaky ReLU, the
//    second VN layer's 21x21 channel mixing (weights and BatchNorm coefficients staged in shared memory from a
//    DEVICE buffer, so nothing about a launch depends on host copies of parameters: broadcast 128-bit reads, 12 FMAs each),
//    its norm / BN / ReLU -- happens in the registers of ONE thread per edge; the mean over k goes through a
//    conflict-free shared-memory transpose, and only [B,21,3,N] (8.3 MB) is written.
//  * BatchNorm in training mode needs batch statistics of the norms before the nonlinearity can be applied, so the
//    training forward is three passes of the same kernel (MODE 0: stage-1 norm sums; MODE 1: stage-2 norm sums; MODE 2:
//    everything), each re-gathering from L2 instead of storing a per-edge tensor.  Eval mode is MODE 2 alone.
//  * Backward recomputes the forward per edge.  BatchNorm's backward needs sum(gy) and sum(gy * rhat) per channel before
//    any input gradient can be formed; gy is LINEAR in the upstream gradient, which for the last stage is G[n]/k for all
//    k edges of a point, so the forward also emits the per-point sums of the linear coefficients (ysum, yrsum: 2 x 63
//    floats per point) and those two sums cost a dot product per point instead of another pass over the edges.
//    edgeconv_bwd2_kernel (two-stage layers) back-propagates stage 2 per edge: gradient of the 21x21 weights by a
//    warp-transposed shuffle reduction (no shared memory, no atomics inside the loop), gO1 = W2^T g written once as
//    [E,64] fp32 (the only per-edge tensor of the layer, 168 MB, written once and read once), and the stage-1 BatchNorm
//    sums.  edgeconv_bwd1_kernel back-propagates stage 1: the neighbour half of the gradient is scattered to gU[j]
//    with 128-bit vector reductions, the centre half is summed over k through shared memory into gV[i].
//  The per-point maps x -> (U, V) and (gU, gV) -> (gx, gW1) are tiny dense contractions over B*N points.
#include "common.cuh"

namespace hpcs {

constexpr int kVO = 21;                 // vector channels out of every VN conv in the EdgeConv layers (64 // 3)
constexpr int kVD = 3 * kVO;            // 63 floats: (channel, component) of one half row
constexpr int kRowF = 128;              // floats per point row: [0,63) feat | [63] 0 | [64,127) dir | [127] 0
constexpr float kVnEps = 1e-6f;         // EPS of vn_layers.py:10
constexpr float kOneMinusSlope = 0.8f;  // 1 - negative_slope, negative_slope = 0.2 (vn_layers.py:49)
constexpr int kRedStride = 65;          // shared-memory row stride (floats) of the k-reduction buffer

// Packed per-layer coefficients, a DEVICE buffer of kWFloats floats assembled by the host mirror with device ops
// (no host copies of parameters or batch statistics; a step stays capturable in a CUDA graph):
//   [stage s in {0,1}][which in 0..5][21]  BatchNorm on the norm, folded:  0 a, 1 b (y = a r + b), 2 mu, 3 rstd
//                                          (rhat = (r - mu) rstd), 4 s1m = mean(gy), 5 s2m = mean(gy rhat) (backward only)
//   [kOffWf + o*24 + i]  stage-2 map_to_feat weight [out o][in i], rows padded 21 -> 24 with zeros;  kOffWd: map_to_dir
//   [kOffW1 + m*21 + o]  C = 1 layers only: the first conv's weights themselves, W = [Wa | Wb] (one input channel each), so
//                        that p = Wa (x_j - x_i) + Wb x_i is evaluated from the coordinates like the reference does (the
//                        U[j] + V[i] split would cancel |x| against |x_j - x_i|, and a 12-byte gather beats a 512-byte one)
constexpr int kWPad = 24;
constexpr int kOffBn = 6 * kVO;                      // floats per stage of BatchNorm coefficients
constexpr int kOffWf = 256;
constexpr int kOffWd = kOffWf + kVO * kWPad;
constexpr int kOffW1 = kOffWd + kVO * kWPad;         // 1264: [4][21] first-conv weights of a C = 1 layer: Wa_feat, Wa_dir, Wb_feat, Wb_dir
constexpr int kWFloats = kOffW1 + 4 * kVO + 4;       // 1352
enum { BN_A = 0, BN_B = 1, BN_MU = 2, BN_RSTD = 3, BN_S1M = 4, BN_S2M = 5 };

__device__ __forceinline__ float bn_coef(const float* ws, int stage, int which, int o) { return ws[stage * kOffBn + which * kVO + o]; }

__device__ __forceinline__ void stage_weights(float* ws, const float* __restrict__ Wdev) {
    for (int i = threadIdx.x; i < kWFloats; i += blockDim.x) ws[i] = Wdev[i];
    __syncthreads();
}

// (p, d) of output channel o = rows o of the two stage-2 weight matrices times the 21 input vectors X[3i + c]
template <bool WITH_D>
__device__ __forceinline__ void mix_channel(const float* ws, int o, const float (&X)[64], float& p0, float& p1, float& p2, float& d0,
                                            float& d1, float& d2) {
    const float4* wf = reinterpret_cast<const float4*>(ws + kOffWf + o * kWPad);
    const float4* wd = reinterpret_cast<const float4*>(ws + kOffWd + o * kWPad);
    p0 = p1 = p2 = d0 = d1 = d2 = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < kWPad / 4; ++i4) {
        const float4 a = wf[i4];
        const float af[4] = {a.x, a.y, a.z, a.w};
        float bf[4] = {0.f, 0.f, 0.f, 0.f};
        if (WITH_D) {
            const float4 b = wd[i4];
            bf[0] = b.x; bf[1] = b.y; bf[2] = b.z; bf[3] = b.w;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * i4 + u;
            if (i < kVO) {
                p0 = fmaf(af[u], X[3 * i], p0); p1 = fmaf(af[u], X[3 * i + 1], p1); p2 = fmaf(af[u], X[3 * i + 2], p2);
                if (WITH_D) { d0 = fmaf(bf[u], X[3 * i], d0); d1 = fmaf(bf[u], X[3 * i + 1], d1); d2 = fmaf(bf[u], X[3 * i + 2], d2); }
            }
        }
    }
}

// acc[3i + c] += wf[o][i] gp[c] + wd[o][i] gd[c]   (transpose of mix_channel)
__device__ __forceinline__ void mix_channel_transposed(const float* ws, int o, float gp0, float gp1, float gp2, float gd0, float gd1,
                                                       float gd2, float (&acc)[64]) {
    const float4* wf = reinterpret_cast<const float4*>(ws + kOffWf + o * kWPad);
    const float4* wd = reinterpret_cast<const float4*>(ws + kOffWd + o * kWPad);
#pragma unroll
    for (int i4 = 0; i4 < kWPad / 4; ++i4) {
        const float4 a = wf[i4], b = wd[i4];
        const float af[4] = {a.x, a.y, a.z, a.w}, bf[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = 4 * i4 + u;
            if (i < kVO) {
                acc[3 * i] = fmaf(af[u], gp0, fmaf(bf[u], gd0, acc[3 * i]));
                acc[3 * i + 1] = fmaf(af[u], gp1, fmaf(bf[u], gd1, acc[3 * i + 1]));
                acc[3 * i + 2] = fmaf(af[u], gp2, fmaf(bf[u], gd2, acc[3 * i + 2]));
            }
        }
    }
}

struct ChanFwd {
    float o0, o1, o2;       // output vector
    float q0, q1, q2;       // after the BatchNorm rescale
    float nr, r, s, dot, dinv;
    bool neg;
};

// One vector channel of VNLinearLeakyReLU after the Linear maps (vn_layers.py:66-77, 124-130):
//   r = |p| + EPS;  q = p / r * bn(r);  dot = q.d;  o = q                       if dot >= 0
//                                                    o = q - 0.8 dot/(|d|^2+EPS) d  otherwise
__device__ __forceinline__ ChanFwd chan_fwd(float p0, float p1, float p2, float d0, float d1, float d2, float a, float b) {
    ChanFwd f;
    f.nr = sqrtf(fmaf(p2, p2, fmaf(p1, p1, p0 * p0)));
    f.r = f.nr + kVnEps;
    f.s = a + __fdividef(b, f.r);
    f.q0 = p0 * f.s; f.q1 = p1 * f.s; f.q2 = p2 * f.s;
    f.dot = fmaf(f.q2, d2, fmaf(f.q1, d1, f.q0 * d0));
    f.dinv = __fdividef(1.0f, fmaf(d2, d2, fmaf(d1, d1, d0 * d0)) + kVnEps);
    f.neg = f.dot < 0.f;
    const float t = f.neg ? kOneMinusSlope * f.dot * f.dinv : 0.f;
    f.o0 = fmaf(-t, d0, f.q0); f.o1 = fmaf(-t, d1, f.q1); f.o2 = fmaf(-t, d2, f.q2);
    return f;
}

// Coefficients of gy = dL/d bn(r) as a linear function of the upstream gradient gO of this channel: gy = Y . gO
__device__ __forceinline__ void chan_y(const ChanFwd& f, float p0, float p1, float p2, float d0, float d1, float d2,
                                       float& y0, float& y1, float& y2) {
    const float rinv = __fdividef(1.0f, f.r);
    const float c = f.neg ? kOneMinusSlope * f.dinv * fmaf(d2, p2, fmaf(d1, p1, d0 * p0)) : 0.f;
    y0 = fmaf(-c, d0, p0) * rinv; y1 = fmaf(-c, d1, p1) * rinv; y2 = fmaf(-c, d2, p2) * rinv;
}

// Backward of one channel: upstream gO -> gradients wrt p (through the ReLU, the rescale and BatchNorm) and wrt d.
// Also returns gy and rhat (for the BatchNorm parameter / statistics sums).
__device__ __forceinline__ void chan_bwd(const ChanFwd& f, float p0, float p1, float p2, float d0, float d1, float d2,
                                         float g0, float g1, float g2, float a, float mu, float rstd, float s1m, float s2m,
                                         float& gp0, float& gp1, float& gp2, float& gd0, float& gd1, float& gd2,
                                         float& gy, float& rhat) {
    float gq0 = g0, gq1 = g1, gq2 = g2;
    gd0 = gd1 = gd2 = 0.f;
    if (f.neg) {
        const float t = f.dot * f.dinv;
        const float hd = fmaf(g2, d2, fmaf(g1, d1, g0 * d0)) * f.dinv;
        gq0 = fmaf(-kOneMinusSlope * hd, d0, g0); gq1 = fmaf(-kOneMinusSlope * hd, d1, g1); gq2 = fmaf(-kOneMinusSlope * hd, d2, g2);
        gd0 = -kOneMinusSlope * fmaf(hd, fmaf(-2.f * t, d0, f.q0), t * g0);
        gd1 = -kOneMinusSlope * fmaf(hd, fmaf(-2.f * t, d1, f.q1), t * g1);
        gd2 = -kOneMinusSlope * fmaf(hd, fmaf(-2.f * t, d2, f.q2), t * g2);
    }
    const float gs = fmaf(gq2, p2, fmaf(gq1, p1, gq0 * p0));
    const float rinv = __fdividef(1.0f, f.r);
    gy = gs * rinv;
    rhat = (f.r - mu) * rstd;
    const float gr = fmaf(a, gy - s1m - rhat * s2m, -gs * f.s * rinv);
    const float w = f.nr > 0.f ? __fdividef(gr, f.nr) : 0.f;           // d|p|/dp = p/|p| (0 at the origin, like torch.norm)
    gp0 = fmaf(gq0, f.s, w * p0); gp1 = fmaf(gq1, f.s, w * p1); gp2 = fmaf(gq2, f.s, w * p2);
}

// 32 values per lane -> lane l receives the warp-wide sum of v[l]  (31 shuffles instead of 32 x 5)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const float keep = up ? v[j + s] : v[j];
            const float send = up ? v[j] : v[j + s];
            v[j] = keep + __shfl_xor_sync(kFull, send, s);
        }
    }
    return v[0];
}

// ---- per-point linear maps ------------------------------------------------------------------------------------------
// x[B,C,3,N], W4[4][21][C] = {Uf, Ud, Vf, Vd} -> UU[B*N][128], VV[B*N][128]
constexpr int kPlPts = 32;
__global__ void __launch_bounds__(256) vn_point_linear_kernel(const float* __restrict__ x, const float* __restrict__ W4, int C, int N,
                                                              float* __restrict__ UU, float* __restrict__ VV) {
    extern __shared__ float sm[];
    float* xs = sm;                                     // [3C][kPlPts]
    float* ws = xs + 3 * C * kPlPts;                    // [4][21][C]
    float* outs = ws + 4 * kVO * C;                     // [2][kPlPts][128]
    const int b = blockIdx.y, n0 = blockIdx.x * kPlPts;
    const int np = min(kPlPts, N - n0);
    for (int i = threadIdx.x; i < 3 * C * kPlPts; i += blockDim.x) {
        const int row = i / kPlPts, p = i % kPlPts;
        xs[i] = p < np ? x[((size_t)b * 3 * C + row) * N + n0 + p] : 0.f;
    }
    for (int i = threadIdx.x; i < 4 * kVO * C; i += blockDim.x) ws[i] = W4[i];
    for (int i = threadIdx.x; i < 2 * kPlPts * kRowF; i += blockDim.x) outs[i] = 0.f;
    __syncthreads();
    // thread -> (point p, channel o): 12 outputs
    for (int t = threadIdx.x; t < kPlPts * kVO; t += blockDim.x) {
        const int p = t % kPlPts, o = t / kPlPts;
        float acc[4][3] = {};
        for (int ci = 0; ci < C; ++ci) {
            const float x0 = xs[(ci * 3 + 0) * kPlPts + p], x1 = xs[(ci * 3 + 1) * kPlPts + p], x2 = xs[(ci * 3 + 2) * kPlPts + p];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float w = ws[(m * kVO + o) * C + ci];
                acc[m][0] = fmaf(w, x0, acc[m][0]); acc[m][1] = fmaf(w, x1, acc[m][1]); acc[m][2] = fmaf(w, x2, acc[m][2]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            outs[(0 * kPlPts + p) * kRowF + o * 3 + c] = acc[0][c];
            outs[(0 * kPlPts + p) * kRowF + 64 + o * 3 + c] = acc[1][c];
            outs[(1 * kPlPts + p) * kRowF + o * 3 + c] = acc[2][c];
            outs[(1 * kPlPts + p) * kRowF + 64 + o * 3 + c] = acc[3][c];
        }
    }
    __syncthreads();
    const size_t row0 = ((size_t)b * N + n0) * kRowF;
    for (int i = threadIdx.x; i < np * kRowF / 4; i += blockDim.x) {
        reinterpret_cast<float4*>(UU + row0)[i] = reinterpret_cast<const float4*>(outs)[i];
        reinterpret_cast<float4*>(VV + row0)[i] = reinterpret_cast<const float4*>(outs + kPlPts * kRowF)[i];
    }
}

// ---- per-edge helpers -------------------------------------------------------------------------------------------------
struct EdgeId {
    long long g;        // global point index b*N + n of the centre
    long long m;        // global point index of the neighbour
    long long b;        // cloud
    int n, mloc;        // centre / neighbour index inside the cloud
    int pl;             // point slot inside the tile
    bool valid;
};

__device__ __forceinline__ EdgeId edge_of(long long tile, int P, int k, long long BN, int N, const long long* __restrict__ idx) {
    EdgeId e;
    e.pl = threadIdx.x / k;
    const int j = threadIdx.x - e.pl * k;
    e.g = tile * P + e.pl;
    e.valid = e.pl < P && e.g < BN;
    e.m = e.b = 0;
    e.n = e.mloc = 0;
    if (e.valid) {
        e.b = e.g / N;
        e.n = (int)(e.g - e.b * N);
        e.mloc = (int)idx[e.g * k + j];
        e.m = e.b * N + e.mloc;
    }
    return e;
}

// C = 1 layers: p = Wa (x_j - x_i) + Wb x_i straight from the coordinates x[B,1,3,N]
__device__ __forceinline__ void load_pd_direct(const float* __restrict__ x, const float* ws, int N, const EdgeId& e, float (&P)[64],
                                               float (&D)[64]) {
    const float* xb = x + e.b * 3 * N;
    float xi[3], dx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xi[c] = __ldg(xb + (size_t)c * N + e.n);
        dx[c] = __ldg(xb + (size_t)c * N + e.mloc) - xi[c];
    }
#pragma unroll
    for (int o = 0; o < kVO; ++o) {
        const float waf = ws[kOffW1 + o], wad = ws[kOffW1 + kVO + o], wbf = ws[kOffW1 + 2 * kVO + o], wbd = ws[kOffW1 + 3 * kVO + o];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            P[3 * o + c] = fmaf(waf, dx[c], wbf * xi[c]);
            D[3 * o + c] = fmaf(wad, dx[c], wbd * xi[c]);
        }
    }
    P[63] = D[63] = 0.f;
}

// P = U[m] + V[g] (both halves): 32 + 32 128-bit loads, the V row is shared by the k threads of a point (L1)
__device__ __forceinline__ void load_pd(const float* __restrict__ UU, const float* __restrict__ VV, const EdgeId& e, float (&P)[64],
                                        float (&D)[64]) {
    const float4* u = reinterpret_cast<const float4*>(UU + e.m * kRowF);
    const float4* v = reinterpret_cast<const float4*>(VV + e.g * kRowF);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 a = __ldg(u + q), c = __ldg(v + q);
        P[4 * q] = a.x + c.x; P[4 * q + 1] = a.y + c.y; P[4 * q + 2] = a.z + c.z; P[4 * q + 3] = a.w + c.w;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const float4 a = __ldg(u + 16 + q), c = __ldg(v + 16 + q);
        D[4 * q] = a.x + c.x; D[4 * q + 1] = a.y + c.y; D[4 * q + 2] = a.z + c.z; D[4 * q + 3] = a.w + c.w;
    }
}

// sum over the k threads of every point of a tile, through shared memory: red[T][65] <- one 63-vector per thread
template <typename Emit>
__device__ __forceinline__ void reduce_over_k(float* red, int P, int k, bool point_major, Emit emit) {
    __syncthreads();
    for (int q = threadIdx.x; q < P * kVD; q += blockDim.x) {
        const int pl = point_major ? q / kVD : q % P;
        const int oc = point_major ? q % kVD : q / P;
        float s = 0.f;
        for (int j = 0; j < k; ++j) s += red[(pl * k + j) * kRedStride + oc];
        emit(pl, oc, s);
    }
    __syncthreads();
}

// ---- forward --------------------------------------------------------------------------------------------------------------
struct FwdArgs {
    const float* UU; const float* VV; const long long* idx; const float* Wdev; const float* x;
    long long BN; int N; int k; int P;
    double* stats;          // MODE 0/1: [21][2] sum r, sum r^2
    float* out;             // MODE 2: [B,21,3,N]
    float* ysum;            // MODE 2, optional: [B*N][63] sum_j Y
    float* yrsum;           //                   [B*N][63] sum_j Y rhat
};

// MODE 0: statistics of the stage-1 norms; MODE 1: of the stage-2 norms; MODE 2: the layer's output (+ ysum / yrsum)
template <int STAGES, int MODE, bool DIRECT>
__global__ void __launch_bounds__(256, 1) edgeconv_fwd_kernel(const FwdArgs A) {
    extern __shared__ float smem_f[];
    float* ws = smem_f;                                  // packed coefficients
    float* red = smem_f + kWFloats;                      // [T][65] k-reduction rows
    stage_weights(ws, A.Wdev);
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const int lane = threadIdx.x & 31;
    float sr[kVO], sr2[kVO];
    if (MODE != 2) {
#pragma unroll
        for (int o = 0; o < kVO; ++o) sr[o] = sr2[o] = 0.f;
    }
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        float P[64], D[64];
        if (e.valid) { if (DIRECT) load_pd_direct(A.x, ws, A.N, e, P, D); else load_pd(A.UU, A.VV, e, P, D); }
        else {
#pragma unroll
            for (int i = 0; i < 64; ++i) P[i] = D[i] = 0.f;
        }
        float* myrow = red + threadIdx.x * kRedStride;
        if (MODE == 0) {
#pragma unroll
            for (int o = 0; o < kVO; ++o) {
                const float r = sqrtf(fmaf(P[3 * o + 2], P[3 * o + 2], fmaf(P[3 * o + 1], P[3 * o + 1], P[3 * o] * P[3 * o]))) + kVnEps;
                if (e.valid) { sr[o] += r; sr2[o] = fmaf(r, r, sr2[o]); }
            }
            continue;
        }
        // ---- stage 1 in place: P <- O1 (and, for a one-stage layer in MODE 2, its Y coefficients) -------------------------
        if (STAGES == 1) {
            float Ys[kVD], Yr[kVD];
#pragma unroll
            for (int o = 0; o < kVO; ++o) {
                const ChanFwd f = chan_fwd(P[3 * o], P[3 * o + 1], P[3 * o + 2], D[3 * o], D[3 * o + 1], D[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
                if (A.ysum) {
                    chan_y(f, P[3 * o], P[3 * o + 1], P[3 * o + 2], D[3 * o], D[3 * o + 1], D[3 * o + 2], Ys[3 * o], Ys[3 * o + 1], Ys[3 * o + 2]);
                    const float rh = (f.r - bn_coef(ws, 0, BN_MU, o)) * bn_coef(ws, 0, BN_RSTD, o);
                    Yr[3 * o] = Ys[3 * o] * rh; Yr[3 * o + 1] = Ys[3 * o + 1] * rh; Yr[3 * o + 2] = Ys[3 * o + 2] * rh;
                }
                P[3 * o] = f.o0; P[3 * o + 1] = f.o1; P[3 * o + 2] = f.o2;
            }
            const float inv_k = 1.0f / A.k;
#pragma unroll
            for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? P[i] : 0.f;
            reduce_over_k(red, A.P, A.k, false, [&](int pl, int oc, float s) {
                const long long g = tile * A.P + pl;
                if (g < A.BN) A.out[((g / A.N) * kVD + oc) * A.N + g % A.N] = s * inv_k;
            });
            if (A.ysum) {
#pragma unroll
                for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? Ys[i] : 0.f;
                reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
                    const long long g = tile * A.P + pl;
                    if (g < A.BN) A.ysum[g * kVD + oc] = s;
                });
#pragma unroll
                for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? Yr[i] : 0.f;
                reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
                    const long long g = tile * A.P + pl;
                    if (g < A.BN) A.yrsum[g * kVD + oc] = s;
                });
            }
            continue;
        }
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            const ChanFwd f = chan_fwd(P[3 * o], P[3 * o + 1], P[3 * o + 2], D[3 * o], D[3 * o + 1], D[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            P[3 * o] = f.o0; P[3 * o + 1] = f.o1; P[3 * o + 2] = f.o2;
        }
        // ---- stage 2, one output channel at a time (weights are constant-bank operands) -----------------------------------
        const bool want_y = MODE == 2 && A.ysum != nullptr;
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            float p0, p1, p2, d0, d1, d2;
            if (MODE == 1) {
                mix_channel<false>(ws, o, P, p0, p1, p2, d0, d1, d2);
                const float r = sqrtf(fmaf(p2, p2, fmaf(p1, p1, p0 * p0))) + kVnEps;
                if (e.valid) { sr[o] += r; sr2[o] = fmaf(r, r, sr2[o]); }
                continue;
            }
            mix_channel<true>(ws, o, P, p0, p1, p2, d0, d1, d2);
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_B, o));
            myrow[3 * o] = e.valid ? f.o0 : 0.f; myrow[3 * o + 1] = e.valid ? f.o1 : 0.f; myrow[3 * o + 2] = e.valid ? f.o2 : 0.f;
            if (want_y) {                                  // Y and Y*rhat parked in D (its stage-1 content is dead)
                float y0, y1, y2;
                chan_y(f, p0, p1, p2, d0, d1, d2, y0, y1, y2);
                D[3 * o] = y0; D[3 * o + 1] = y1; D[3 * o + 2] = y2;
                D[63] = 0.f;
            }
        }
        if (MODE == 1) continue;
        const float inv_k = 1.0f / A.k;
        reduce_over_k(red, A.P, A.k, false, [&](int pl, int oc, float s) {
            const long long g = tile * A.P + pl;
            if (g < A.BN) A.out[((g / A.N) * kVD + oc) * A.N + g % A.N] = s * inv_k;
        });
        if (want_y) {
#pragma unroll
            for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? D[i] : 0.f;
            reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
                const long long g = tile * A.P + pl;
                if (g < A.BN) A.ysum[g * kVD + oc] = s;
            });
            // rhat of stage 2 needs the norms again: recompute them from O1 (P) -- cheaper than 21 more live registers
#pragma unroll
            for (int o = 0; o < kVO; ++o) {
                float p0, p1, p2, d0, d1, d2;
                mix_channel<false>(ws, o, P, p0, p1, p2, d0, d1, d2);
                const float r = sqrtf(fmaf(p2, p2, fmaf(p1, p1, p0 * p0))) + kVnEps;
                const float rh = (r - bn_coef(ws, 1, BN_MU, o)) * bn_coef(ws, 1, BN_RSTD, o);
                myrow[3 * o] = e.valid ? D[3 * o] * rh : 0.f; myrow[3 * o + 1] = e.valid ? D[3 * o + 1] * rh : 0.f;
                myrow[3 * o + 2] = e.valid ? D[3 * o + 2] * rh : 0.f;
            }
            reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
                const long long g = tile * A.P + pl;
                if (g < A.BN) A.yrsum[g * kVD + oc] = s;
            });
        }
    }
    if (MODE != 2) {                                     // block reduction of the 42 partial sums, fp64 from the warp level up
        __syncthreads();
        double* dsm = reinterpret_cast<double*>(red);
        const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            const double a = warp_sum((double)sr[o]), b = warp_sum((double)sr2[o]);
            if (lane == 0) { dsm[(warp * kVO + o) * 2] = a; dsm[(warp * kVO + o) * 2 + 1] = b; }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * kVO; i += blockDim.x) {
            double s = 0.;
            for (int w = 0; w < nw; ++w) s += dsm[w * 2 * kVO + i];
            atomicAdd(A.stats + i, s);
        }
    }
}

// ---- backward, stage 2 (two-stage layers) -----------------------------------------------------------------------------
struct Bwd2Args {
    const float* UU; const float* VV; const long long* idx; const float* Wdev; const float* x;
    long long BN; int N; int k; int P;
    const float* G;         // [B,21,3,N] gradient of the layer output
    float* gO1;             // [E][64] out: gradient wrt the stage-1 output of every edge
    float* dW2;             // [21][2][21] = [out o][half: feat, dir][in i] += (atomics at kernel end)
    double* stats1;         // [21][2] += sum gy1, sum gy1 rhat1
};

template <bool DIRECT>
__global__ void __launch_bounds__(256, 1) edgeconv_bwd2_kernel(const Bwd2Args A) {
    extern __shared__ float smem_f[];
    float* ws = smem_f;                                  // packed coefficients
    float* red = smem_f + kWFloats;                      // Gs[P][63]
    stage_weights(ws, A.Wdev);
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const int lane = threadIdx.x & 31;
    constexpr int kGroups = (2 * kVO * kVO + 31) / 32;   // 28 groups of 32 weight-gradient entries
    float dwacc[kGroups];
#pragma unroll
    for (int i = 0; i < kGroups; ++i) dwacc[i] = 0.f;
    float s1[kVO], s2[kVO];
#pragma unroll
    for (int o = 0; o < kVO; ++o) s1[o] = s2[o] = 0.f;
    const float inv_k = 1.0f / A.k;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        __syncthreads();
        for (int q = threadIdx.x; q < A.P * kVD; q += blockDim.x) {           // upstream gradient rows of the tile's points
            const int pl = q % A.P, oc = q / A.P;
            const long long g = tile * A.P + pl;
            red[pl * kVD + oc] = g < A.BN ? A.G[((g / A.N) * kVD + oc) * A.N + g % A.N] * inv_k : 0.f;
        }
        __syncthreads();
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        float P[64], D[64];
        if (e.valid) { if (DIRECT) load_pd_direct(A.x, ws, A.N, e, P, D); else load_pd(A.UU, A.VV, e, P, D); }
        else {
#pragma unroll
            for (int i = 0; i < 64; ++i) P[i] = D[i] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < kVO; ++o) {                                      // stage 1 forward: P <- O1
            const ChanFwd f = chan_fwd(P[3 * o], P[3 * o + 1], P[3 * o + 2], D[3 * o], D[3 * o + 1], D[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            P[3 * o] = f.o0; P[3 * o + 1] = f.o1; P[3 * o + 2] = f.o2;
        }
        float (&gO1)[64] = D;                                                 // D is dead: accumulate gO1 there
#pragma unroll
        for (int i = 0; i < 64; ++i) gO1[i] = 0.f;
        const float* grow = red + (e.pl < A.P ? e.pl : 0) * kVD;
        float stag[32];
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            float p0, p1, p2, d0, d1, d2;
            mix_channel<true>(ws, o, P, p0, p1, p2, d0, d1, d2);
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_B, o));
            const float g0 = e.valid ? grow[3 * o] : 0.f, g1 = e.valid ? grow[3 * o + 1] : 0.f, g2 = e.valid ? grow[3 * o + 2] : 0.f;
            float gp0, gp1, gp2, gd0, gd1, gd2, gy, rhat;
            chan_bwd(f, p0, p1, p2, d0, d1, d2, g0, g1, g2, bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_MU, o), bn_coef(ws, 1, BN_RSTD, o),
                     bn_coef(ws, 1, BN_S1M, o), bn_coef(ws, 1, BN_S2M, o), gp0, gp1, gp2, gd0, gd1, gd2, gy, rhat);
            if (!e.valid) { gp0 = gp1 = gp2 = gd0 = gd1 = gd2 = 0.f; }
            mix_channel_transposed(ws, o, gp0, gp1, gp2, gd0, gd1, gd2, D);       // gO1 += W2^T g
            // weight gradient: entry q = (half h, out o, in i) <- sum over edges of g_h[o] . O1[i]; 32 entries at a time go
            // through the warp transpose so that lane l ends up owning entry (group, l)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < kVO; ++i) {
                    const int q = (o * 2 + h) * kVO + i;                     // visiting order: [out o][half h][in i]
                    const float v = h == 0 ? fmaf(gp2, P[3 * i + 2], fmaf(gp1, P[3 * i + 1], gp0 * P[3 * i]))
                                           : fmaf(gd2, P[3 * i + 2], fmaf(gd1, P[3 * i + 1], gd0 * P[3 * i]));
                    stag[q & 31] = v;
                    if ((q & 31) == 31) dwacc[q >> 5] += warp_transpose_sum(stag, lane);
                }
            }
        }
        {                                                                     // the last, partial group of weight-gradient entries
#pragma unroll
            for (int j = (2 * kVO * kVO) & 31; j < 32; ++j) stag[j] = 0.f;
            dwacc[kGroups - 1] += warp_transpose_sum(stag, lane);
        }
        if (e.valid) {
            float4* dst = reinterpret_cast<float4*>(A.gO1 + (e.g * A.k + (threadIdx.x - e.pl * A.k)) * 64);
#pragma unroll
            for (int q = 0; q < 16; ++q) __stcs(dst + q, make_float4(gO1[4 * q], gO1[4 * q + 1], gO1[4 * q + 2], q == 15 ? 0.f : gO1[4 * q + 3]));
        }
        // stage-1 BatchNorm sums: gy1 = Y1 . gO1, needs the stage-1 inputs again (rows are L1 / L2 hits)
        float Q[64], Dd[64];
        if (e.valid) { if (DIRECT) load_pd_direct(A.x, ws, A.N, e, Q, Dd); else load_pd(A.UU, A.VV, e, Q, Dd); }
        else {
#pragma unroll
            for (int i = 0; i < 64; ++i) Q[i] = Dd[i] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            const ChanFwd f = chan_fwd(Q[3 * o], Q[3 * o + 1], Q[3 * o + 2], Dd[3 * o], Dd[3 * o + 1], Dd[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            float y0, y1, y2;
            chan_y(f, Q[3 * o], Q[3 * o + 1], Q[3 * o + 2], Dd[3 * o], Dd[3 * o + 1], Dd[3 * o + 2], y0, y1, y2);
            const float gy = fmaf(y2, gO1[3 * o + 2], fmaf(y1, gO1[3 * o + 1], y0 * gO1[3 * o]));
            if (e.valid) { s1[o] += gy; s2[o] = fmaf(gy, (f.r - bn_coef(ws, 0, BN_MU, o)) * bn_coef(ws, 0, BN_RSTD, o), s2[o]); }
        }
    }
    // flush: weight-gradient entries (lane l of every warp owns entry 32*group + l) and the 42 BatchNorm sums
#pragma unroll
    for (int gi = 0; gi < kGroups; ++gi) {
        const int q = gi * 32 + lane;
        if (q < 2 * kVO * kVO) atomicAdd(A.dW2 + q, dwacc[gi]);
    }
    __syncthreads();
    double* dsm = reinterpret_cast<double*>(red);
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 0; o < kVO; ++o) {
        const double a = warp_sum((double)s1[o]), b = warp_sum((double)s2[o]);
        if (lane == 0) { dsm[(warp * kVO + o) * 2] = a; dsm[(warp * kVO + o) * 2 + 1] = b; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * kVO; i += blockDim.x) {
        double s = 0.;
        for (int w = 0; w < nw; ++w) s += dsm[w * 2 * kVO + i];
        atomicAdd(A.stats1 + i, s);
    }
}

// ---- backward, stage 1 ------------------------------------------------------------------------------------------------------
struct Bwd1Args {
    const float* UU; const float* VV; const long long* idx; const float* Wdev; const float* x;
    long long BN; int N; int k; int P;
    const float* gO1;       // [E][64] (two-stage layers) or nullptr: then gO1 = G[n]/k (one-stage layers)
    const float* G;         // [B,21,3,N]
    float* gUU;             // [B*N][128] += (vector reductions; zeroed by the caller)
    float* gVV;             // [B*N][128]  = sum over the k edges of the point
};

template <bool DIRECT>
__global__ void __launch_bounds__(256, 1) edgeconv_bwd1_kernel(const Bwd1Args A) {
    extern __shared__ float smem_f[];
    float* ws = smem_f;                                  // packed coefficients
    float* red = smem_f + kWFloats;                      // [T][65] reduction rows, then Gs[P][63] behind them
    float* Gs = red + blockDim.x * kRedStride;
    stage_weights(ws, A.Wdev);
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const float inv_k = 1.0f / A.k;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (!A.gO1) {
            __syncthreads();
            for (int q = threadIdx.x; q < A.P * kVD; q += blockDim.x) {
                const int pl = q % A.P, oc = q / A.P;
                const long long g = tile * A.P + pl;
                Gs[pl * kVD + oc] = g < A.BN ? A.G[((g / A.N) * kVD + oc) * A.N + g % A.N] * inv_k : 0.f;
            }
            __syncthreads();
        }
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        float P[64], D[64];
        if (e.valid) { if (DIRECT) load_pd_direct(A.x, ws, A.N, e, P, D); else load_pd(A.UU, A.VV, e, P, D); }
        else {
#pragma unroll
            for (int i = 0; i < 64; ++i) P[i] = D[i] = 0.f;
        }
        const float4* gsrc = A.gO1 ? reinterpret_cast<const float4*>(A.gO1 + (e.g * A.k + (threadIdx.x - e.pl * A.k)) * 64) : nullptr;
        const float* grow = Gs + (e.pl < A.P ? e.pl : 0) * kVD;
        float gO[64];
        if (gsrc && e.valid) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 v = __ldcs(gsrc + q);
                gO[4 * q] = v.x; gO[4 * q + 1] = v.y; gO[4 * q + 2] = v.z; gO[4 * q + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < kVD; ++i) gO[i] = (e.valid && !gsrc) ? grow[i] : 0.f;
            gO[63] = 0.f;
        }
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            const float p0 = P[3 * o], p1 = P[3 * o + 1], p2 = P[3 * o + 2], d0 = D[3 * o], d1 = D[3 * o + 1], d2 = D[3 * o + 2];
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            float gy, rhat;
            chan_bwd(f, p0, p1, p2, d0, d1, d2, gO[3 * o], gO[3 * o + 1], gO[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_MU, o),
                     bn_coef(ws, 0, BN_RSTD, o), bn_coef(ws, 0, BN_S1M, o), bn_coef(ws, 0, BN_S2M, o), P[3 * o], P[3 * o + 1], P[3 * o + 2], D[3 * o], D[3 * o + 1], D[3 * o + 2], gy, rhat);
        }
        P[63] = D[63] = 0.f;
        if (e.valid) {                                                        // neighbour half: gU[m] += (gP | gD)
            float4* dst = reinterpret_cast<float4*>(A.gUU + e.m * kRowF);
#pragma unroll
            for (int q = 0; q < 16; ++q) atomicAdd(dst + q, make_float4(P[4 * q], P[4 * q + 1], P[4 * q + 2], P[4 * q + 3]));
#pragma unroll
            for (int q = 0; q < 16; ++q) atomicAdd(dst + 16 + q, make_float4(D[4 * q], D[4 * q + 1], D[4 * q + 2], D[4 * q + 3]));
        }
        float* myrow = red + threadIdx.x * kRedStride;                        // centre half: gV[g] = sum over its k edges
#pragma unroll
        for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? P[i] : 0.f;
        reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
            const long long g = tile * A.P + pl;
            if (g < A.BN) A.gVV[g * kRowF + oc] = s;
        });
#pragma unroll
        for (int i = 0; i < kVD; ++i) myrow[i] = e.valid ? D[i] : 0.f;
        reduce_over_k(red, A.P, A.k, true, [&](int pl, int oc, float s) {
            const long long g = tile * A.P + pl;
            if (g < A.BN) A.gVV[g * kRowF + 64 + oc] = s;
        });
    }
}

static int tile_threads(int P, int k) { return (P * k + 31) / 32 * 32; }   // whole warps: the kernels shuffle

static int tile_points(int k) {
    int P = 256 / k;
    if (P > 32) P = 32;
    return P < 1 ? 1 : P;
}

static int edge_grid(long long BN, int P) {
    const long long tiles = (BN + P - 1) / P;
    const long long cap = (long long)sm_count() * 2;
    return (int)(tiles < cap ? tiles : cap);
}

}  // namespace hpcs

using namespace hpcs;

extern "C" int hpcs_vn_point_linear_f32(const float* x, const float* W4, int B, int C, int N, float* UU, float* VV, void* stream) {
    if (!x || !W4 || !UU || !VV) return fail(HPCS_ERR_ARG, "vn_point_linear: null pointer");
    if (B <= 0 || B > 65535 || C < 1 || C > 64 || N <= 0) return fail(HPCS_ERR_ARG, "vn_point_linear: bad shape B=%d C=%d N=%d", B, C, N);
    const size_t smem = sizeof(float) * ((size_t)3 * C * kPlPts + 4 * kVO * C + 2 * kPlPts * kRowF);
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(vn_point_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    vn_point_linear_kernel<<<dim3((N + kPlPts - 1) / kPlPts, B), 256, smem, as_stream(stream)>>>(x, W4, C, N, UU, VV);
    return check_launch("vn_point_linear_kernel");
}

template <typename K>
static void opt_in_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

extern "C" int hpcs_edgeconv_coef_floats(void) { return kWFloats; }

extern "C" int hpcs_edgeconv_fwd_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k, int stages,
                                     const float* coef, const float* x_direct, int mode, double* stats, float* out, float* ysum,
                                     float* yrsum, void* stream) {
    if (!idx || !coef || (!x_direct && (!UU || !VV))) return fail(HPCS_ERR_ARG, "edgeconv_fwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || k > 256 || (stages != 1 && stages != 2) || mode < 0 || mode > 2 || (mode == 1 && stages == 1))
        return fail(HPCS_ERR_ARG, "edgeconv_fwd: bad arguments B=%d N=%d k=%d stages=%d mode=%d", B, N, k, stages, mode);
    if ((mode != 2 && !stats) || (mode == 2 && !out) || ((ysum == nullptr) != (yrsum == nullptr)))
        return fail(HPCS_ERR_ARG, "edgeconv_fwd: output pointers do not match mode %d", mode);
    FwdArgs A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), stats, out, ysum, yrsum};
    const int T = tile_threads(A.P, k);
    const size_t smem = sizeof(float) * ((size_t)kWFloats + (size_t)T * kRedStride) + 64;
    const int grid = edge_grid(A.BN, A.P);
    cudaStream_t st = as_stream(stream);
#define HPCS_LAUNCH_FWD(S, M)                                                      \
    do {                                                                           \
        if (x_direct) {                                                            \
            opt_in_smem(edgeconv_fwd_kernel<S, M, true>, smem);                    \
            edgeconv_fwd_kernel<S, M, true><<<grid, T, smem, st>>>(A);            \
        } else {                                                                   \
            opt_in_smem(edgeconv_fwd_kernel<S, M, false>, smem);                   \
            edgeconv_fwd_kernel<S, M, false><<<grid, T, smem, st>>>(A);           \
        }                                                                          \
    } while (0)
    if (stages == 1) {
        if (mode == 0) HPCS_LAUNCH_FWD(1, 0); else HPCS_LAUNCH_FWD(1, 2);
    } else {
        if (mode == 0) HPCS_LAUNCH_FWD(2, 0); else if (mode == 1) HPCS_LAUNCH_FWD(2, 1); else HPCS_LAUNCH_FWD(2, 2);
    }
#undef HPCS_LAUNCH_FWD
    return check_launch("edgeconv_fwd_kernel");
}

extern "C" int hpcs_edgeconv_bwd_stage2_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k,
                                            const float* coef, const float* x_direct, const float* G, float* gO1, float* dW2,
                                            double* stats1, void* stream) {
    if (!idx || !coef || !G || !gO1 || !dW2 || !stats1 || (!x_direct && (!UU || !VV)))
        return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage2: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || k > 256) return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage2: bad shape");
    Bwd2Args A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), G, gO1, dW2, stats1};
    const int T = tile_threads(A.P, k);
    size_t tail = sizeof(float) * (size_t)A.P * kVD;
    const size_t tail_red = sizeof(double) * 2 * kVO * ((T + 31) / 32);
    if (tail < tail_red) tail = tail_red;
    const size_t smem = sizeof(float) * kWFloats + tail;
    if (x_direct) edgeconv_bwd2_kernel<true><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    else edgeconv_bwd2_kernel<false><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    return check_launch("edgeconv_bwd2_kernel");
}

extern "C" int hpcs_edgeconv_bwd_stage1_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k,
                                            const float* coef, const float* x_direct, const float* gO1, const float* G,
                                            float* gUU, float* gVV, void* stream) {
    if (!idx || !coef || !gUU || !gVV || (!gO1 && !G) || (!x_direct && (!UU || !VV)))
        return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage1: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || k > 256) return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage1: bad shape");
    Bwd1Args A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), gO1, G, gUU, gVV};
    const int T = tile_threads(A.P, k);
    const size_t smem = sizeof(float) * ((size_t)kWFloats + (size_t)T * kRedStride + (size_t)A.P * kVD);
    if (x_direct) {
        opt_in_smem(edgeconv_bwd1_kernel<true>, smem);
        edgeconv_bwd1_kernel<true><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    } else {
        opt_in_smem(edgeconv_bwd1_kernel<false>, smem);
        edgeconv_bwd1_kernel<false><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    }
    return check_launch("edgeconv_bwd1_kernel");
}
