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

// ---- the 21x21 channel mixes of the second conv, on packed fp32 pairs (fma.rn.f32x2 = FFMA2: one issue slot for two FMAs).  The 21
// vectors of the first conv's output live as X2[c][m] = (X[2m][c], X[2m+1][c]), channel 21 = 0; the weight rows are read from shared
// memory as they lie (consecutive input channels), so neither operand needs duplicating.  Measured against scalar FMAs: stage-2
// statistics pass 116 -> 107 us, edgeconv_bwd2 835 -> 770 us (C=1 layer) and 910 -> 865 us (C=21); the full forward is unchanged
// (it waits on the barriers of its k-reduction rounds).
typedef unsigned long long f32x2;
constexpr int kXP = 11;                                   // channel pairs
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float hsum2(f32x2 v) { float lo, hi; unpack2(v, lo, hi); return lo + hi; }

// (p, d) of output channel o: two partial sums per component (even / odd input channels), 11 packed FMAs each
template <bool WITH_D>
__device__ __forceinline__ void mix_channel(const float* ws, int o, const f32x2 (&X2)[3][kXP], float& p0, float& p1, float& p2, float& d0,
                                             float& d1, float& d2) {
    const ulonglong2* wf = reinterpret_cast<const ulonglong2*>(ws + kOffWf + o * kWPad);   // (w[4q], w[4q+1]), (w[4q+2], w[4q+3])
    const ulonglong2* wd = reinterpret_cast<const ulonglong2*>(ws + kOffWd + o * kWPad);
    f32x2 ap[3] = {0ull, 0ull, 0ull}, ad[3] = {0ull, 0ull, 0ull};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const ulonglong2 a = wf[q];
        ulonglong2 b = make_ulonglong2(0ull, 0ull);
        if (WITH_D) b = wd[q];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            ap[c] = ffma2(a.x, X2[c][2 * q], ap[c]);
            if (WITH_D) ad[c] = ffma2(b.x, X2[c][2 * q], ad[c]);
            if (2 * q + 1 < kXP) {
                ap[c] = ffma2(a.y, X2[c][2 * q + 1], ap[c]);
                if (WITH_D) ad[c] = ffma2(b.y, X2[c][2 * q + 1], ad[c]);
            }
        }
    }
    p0 = hsum2(ap[0]); p1 = hsum2(ap[1]); p2 = hsum2(ap[2]);
    d0 = d1 = d2 = 0.f;
    if (WITH_D) { d0 = hsum2(ad[0]); d1 = hsum2(ad[1]); d2 = hsum2(ad[2]); }
}

// acc2[c][m] += (wf[o][2m], wf[o][2m+1]) gp[c] + (wd[o][2m], wd[o][2m+1]) gd[c]   (transpose of mix_channel)
__device__ __forceinline__ void mix_channel_transposed(const float* ws, int o, float gp0, float gp1, float gp2, float gd0, float gd1,
                                                        float gd2, f32x2 (&acc2)[3][kXP]) {
    const ulonglong2* wf = reinterpret_cast<const ulonglong2*>(ws + kOffWf + o * kWPad);
    const ulonglong2* wd = reinterpret_cast<const ulonglong2*>(ws + kOffWd + o * kWPad);
    const f32x2 gp[3] = {pack2(gp0, gp0), pack2(gp1, gp1), pack2(gp2, gp2)};
    const f32x2 gd[3] = {pack2(gd0, gd0), pack2(gd1, gd1), pack2(gd2, gd2)};
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const ulonglong2 a = wf[q], b = wd[q];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc2[c][2 * q] = ffma2(a.x, gp[c], ffma2(b.x, gd[c], acc2[c][2 * q]));
            if (2 * q + 1 < kXP) acc2[c][2 * q + 1] = ffma2(a.y, gp[c], ffma2(b.y, gd[c], acc2[c][2 * q + 1]));
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
__device__ __forceinline__ float vnorm(float p0, float p1, float p2) {
    const float r2 = fmaf(p2, p2, fmaf(p1, p1, p0 * p0));
    return r2 > 0.f ? r2 * rsqrtf(r2) : 0.f;              // MUFU.RSQ (2 ulp): no IEEE-sqrt slow path in the inner loops
}

__device__ __forceinline__ ChanFwd chan_fwd(float p0, float p1, float p2, float d0, float d1, float d2, float a, float b) {
    ChanFwd f;
    f.nr = vnorm(p0, p1, p2);
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

// ---- per-point linear maps ------------------------------------------------------------------------------------------
// x[B,C,3,N], W4[4][21][C] = {Uf, Ud, Vf, Vd} -> UU[B*N][128], VV[B*N][128]
constexpr int kPlPts = 32;
constexpr int kPlStride = kRowF + 1;                    // (a 128-float stride put a warp's 12 stores per item in one bank: 60 -> 30 us)
__global__ void __launch_bounds__(256) vn_point_linear_kernel(const float* __restrict__ x, const float* __restrict__ W4, int C, int N,
                                                              float* __restrict__ UU, float* __restrict__ VV) {
    extern __shared__ float sm[];
    float* xs = sm;                                     // [3C][kPlPts]
    float* ws = xs + 3 * C * kPlPts;                    // [4][21][C]
    float* outs = ws + 4 * kVO * C;                     // [2][kPlPts][kPlStride]: odd stride, a warp's 32 points hit 32 banks
    const int b = blockIdx.y, n0 = blockIdx.x * kPlPts;
    const int np = min(kPlPts, N - n0);
    for (int i = threadIdx.x; i < 3 * C * kPlPts; i += blockDim.x) {
        const int row = i / kPlPts, p = i % kPlPts;
        xs[i] = p < np ? x[((size_t)b * 3 * C + row) * N + n0 + p] : 0.f;
    }
    for (int i = threadIdx.x; i < 4 * kVO * C; i += blockDim.x) ws[i] = W4[i];
    for (int i = threadIdx.x; i < 2 * kPlPts * kPlStride; i += blockDim.x) outs[i] = 0.f;
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
            outs[(0 * kPlPts + p) * kPlStride + o * 3 + c] = acc[0][c];
            outs[(0 * kPlPts + p) * kPlStride + 64 + o * 3 + c] = acc[1][c];
            outs[(1 * kPlPts + p) * kPlStride + o * 3 + c] = acc[2][c];
            outs[(1 * kPlPts + p) * kPlStride + 64 + o * 3 + c] = acc[3][c];
        }
    }
    __syncthreads();
    const size_t row0 = ((size_t)b * N + n0) * kRowF;
    for (int i = threadIdx.x; i < np * kRowF; i += blockDim.x) {        // a warp writes 128 contiguous bytes of one row
        const int r = i / kRowF, col = i % kRowF;
        UU[row0 + i] = outs[r * kPlStride + col];
        VV[row0 + i] = outs[(kPlPts + r) * kPlStride + col];
    }
}

// ---- BatchNorm bookkeeping on the device (one launch instead of ~15 tiny tensor ops per conv) ------------------------------
// stats[21][2] = sum r, sum r^2 over M norms -> batch mean / biased variance; running buffers updated like nn.BatchNorm2d
// (unbiased variance, momentum; momentum < 0: cumulative average with the already incremented num_batches_tracked); folded
// coefficients a, b, mu, rstd written to the coefficient buffer of `stage`.  training == 0: the running buffers are the statistics.
__global__ void edgeconv_bn_fold_kernel(const double* __restrict__ stats, double M, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float* __restrict__ running_mean, float* __restrict__ running_var,
                                        const long long* __restrict__ num_batches_tracked, float momentum, float eps, int training,
                                        float* __restrict__ coef, int stage) {
    const int o = threadIdx.x;
    if (o >= kVO) return;
    double mean, var;
    if (training) {
        mean = stats[2 * o] / M;
        var = stats[2 * o + 1] / M - mean * mean;
        if (var < 0.) var = 0.;
        if (running_mean) {
            const double mom = momentum >= 0.f ? (double)momentum : 1.0 / (double)(num_batches_tracked ? *num_batches_tracked : 1);
            running_mean[o] = (float)((1.0 - mom) * running_mean[o] + mom * mean);
            running_var[o] = (float)((1.0 - mom) * running_var[o] + mom * var * (M / (M > 1. ? M - 1. : 1.)));
        }
    } else {
        mean = running_mean[o];
        var = running_var[o];
    }
    const float meanf = (float)mean;
    const float rstd = rsqrtf((float)var + eps);
    const float a = gamma[o] * rstd;
    float* c = coef + stage * kOffBn;
    c[BN_A * kVO + o] = a;
    c[BN_B * kVO + o] = beta[o] - meanf * a;
    c[BN_MU * kVO + o] = meanf;
    c[BN_RSTD * kVO + o] = rstd;
}

// S1[o] = sum G.ysum / k, S2[o] = sum G.yrsum / k over all points and components (the BatchNorm-backward sums of the last conv):
// one warp per (cloud, channel-component) row of G, coalesced along n; fp64 accumulation across warps.
__global__ void __launch_bounds__(256) edgeconv_bn_sums_kernel(const float* __restrict__ G, const float* __restrict__ ysum,
                                                               const float* __restrict__ yrsum, int B, int N, float inv_k,
                                                               double* __restrict__ sums /*[21][2], zeroed by the caller*/) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= B * kVD) return;
    const int b = row / kVD, oc = row - b * kVD;
    const float* g = G + (size_t)row * N;
    const float* y = ysum + (size_t)b * N * kVD + oc;
    const float* yr = yrsum + (size_t)b * N * kVD + oc;
    float s1 = 0.f, s2 = 0.f;
    for (int n = lane; n < N; n += 32) {
        const float gv = __ldg(g + n) * inv_k;
        s1 = fmaf(gv, __ldg(y + (size_t)n * kVD), s1);
        s2 = fmaf(gv, __ldg(yr + (size_t)n * kVD), s2);
    }
    const double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
    if (lane == 0) { atomicAdd(sums + 2 * (oc / 3), d1); atomicAdd(sums + 2 * (oc / 3) + 1, d2); }
}

// s1m = S1 / M, s2m = S2 / M into the coefficient buffer (training), gradients of the BatchNorm affine out
__global__ void edgeconv_bn_sums_finish_kernel(const double* __restrict__ sums, double M, int training, float* __restrict__ coef, int stage,
                                               float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int o = threadIdx.x;
    if (o >= kVO) return;
    const double s1 = sums[2 * o], s2 = sums[2 * o + 1];
    float* c = coef + stage * kOffBn;
    c[BN_S1M * kVO + o] = training ? (float)(s1 / M) : 0.f;
    c[BN_S2M * kVO + o] = training ? (float)(s2 / M) : 0.f;
    dgamma[o] = (float)s2;
    dbeta[o] = (float)s1;
}

// ---- backward of the per-point maps ------------------------------------------------------------------------------------
// gUU / gVV [B*N][128] (gradients wrt the U and V rows) -> gx[B,C,3,N] = sum_m W4[m]^T g_m  and  dW4[4][21][C] += sum over
// points and components of g_m (x) x.  64 points per block; the block's weight-gradient partials leave as one atomic each.
constexpr int kPbPts = 64;
constexpr int kCiGrp = 7;                               // input channels per thread
constexpr int kPbStride = 129;                          // odd stride: a warp's 32 points read a column without bank conflicts
constexpr int kPbXs = kPbPts + 1;                       // x rows: lanes = input channels, 3 * 65 floats apart = 21 different banks (3 * 64 floats:
                                                        // one bank, a 21-way conflict on every load of the loop -- 181 -> 86 us)
__global__ void __launch_bounds__(256) vn_point_linear_bwd_kernel(const float* __restrict__ gUU, const float* __restrict__ gVV,
                                                                  const float* __restrict__ x, const float* __restrict__ W4, int C, int N,
                                                                  int nblk, float* __restrict__ gx, float* __restrict__ dW4) {
    extern __shared__ __align__(16) float sm[];
    float* dws = sm;                                    // [4][21][C] weight-gradient sums of this CTA over all its point blocks (16-byte aligned)
    float* ws = dws + 4 * kVO * C;                      // [4][21][C]
    float* gs = ws + 4 * kVO * C;                       // [2][kPbPts][129]: rows of gUU, then of gVV
    float* xs = gs + 2 * kPbPts * kPbStride;            // [3C][kPbXs]: the weight-gradient loop reads a column of it per lane (lane = channel)
    for (int i = threadIdx.x; i < 4 * kVO * C; i += blockDim.x) { ws[i] = W4[i]; dws[i] = 0.f; }
    const int bpc = (N + kPbPts - 1) / kPbPts;          // point blocks per cloud
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int b = blk / bpc, n0 = (blk - b * bpc) * kPbPts;
    const int np = min(kPbPts, N - n0);
    const size_t row0 = ((size_t)b * N + n0) * kRowF;
    __syncthreads();                                    // the previous block's readers are done with gs / xs
    for (int i = threadIdx.x; i < 2 * kPbPts * kRowF; i += blockDim.x) {
        const int half = i / (kPbPts * kRowF), r = (i / kRowF) % kPbPts, col = i % kRowF;
        gs[(half * kPbPts + r) * kPbStride + col] = r < np ? __ldg((half ? gVV : gUU) + row0 + (size_t)r * kRowF + col) : 0.f;
    }
    for (int i = threadIdx.x; i < 3 * C * kPbPts; i += blockDim.x) {
        const int row = i / kPbPts, p = i % kPbPts;
        xs[row * kPbXs + p] = p < np ? __ldg(x + ((size_t)b * 3 * C + row) * N + n0 + p) : 0.f;
    }
    __syncthreads();
    // gx: thread -> (point p, kCiGrp input channels), all three components: the point's 4 x 21 gradient vectors are read once per
    // group of channels (m = 0,1: the gUU row (feat | dir), m = 2,3: the gVV row); weights are warp-wide broadcasts
    const int ngrp = (C + kCiGrp - 1) / kCiGrp;
    for (int t = threadIdx.x; t < kPbPts * ngrp; t += blockDim.x) {
        const int p = t % kPbPts, c0 = (t / kPbPts) * kCiGrp;
        if (p >= np) continue;
        float a[kCiGrp][3];
#pragma unroll
        for (int u = 0; u < kCiGrp; ++u) a[u][0] = a[u][1] = a[u][2] = 0.f;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const float* g = gs + ((m >> 1) * kPbPts + p) * kPbStride + (m & 1) * 64;
            for (int o = 0; o < kVO; ++o) {
                const float g0 = g[3 * o], g1 = g[3 * o + 1], g2 = g[3 * o + 2];
                const float* w = ws + (m * kVO + o) * C + c0;
#pragma unroll
                for (int u = 0; u < kCiGrp; ++u) {
                    if (c0 + u < C) {
                        const float wv = w[u];
                        a[u][0] = fmaf(wv, g0, a[u][0]); a[u][1] = fmaf(wv, g1, a[u][1]); a[u][2] = fmaf(wv, g2, a[u][2]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kCiGrp; ++u) {
            if (c0 + u < C) {
                float* dst = gx + ((size_t)b * 3 * C + (c0 + u) * 3) * N + n0 + p;
                dst[0] = a[u][0]; dst[N] = a[u][1]; dst[2 * (size_t)N] = a[u][2];
            }
        }
    }
    // dW4[m][o][ci]: thread -> (3 consecutive (m, o) rows, kCiGrp input channels, a third of the block's points): 9 + 21 loads feed
    // 63 FMAs per point (one entry per thread needed 6 loads per 3 FMAs and left the kernel bound by the shared-memory pipe);
    // the partial sums are collected in shared memory over all point blocks of the CTA
    constexpr int kMoGrp = 3, kPSplit = 3;                  // 84 rows = 28 groups of 3 (a group never straddles m: 21 = 7 x 3)
    const int pper = (kPbPts + kPSplit - 1) / kPSplit;
    for (int t = threadIdx.x; t < (4 * kVO / kMoGrp) * ngrp * kPSplit; t += blockDim.x) {
        const int ps = t % kPSplit, cg = (t / kPSplit) % ngrp, mog = t / (kPSplit * ngrp);
        const int mo0 = mog * kMoGrp, m = mo0 / kVO, o0 = mo0 - m * kVO, c0 = cg * kCiGrp;
        const float* g = gs + (m >> 1) * kPbPts * kPbStride + (m & 1) * 64 + 3 * o0;
        float acc[kMoGrp][kCiGrp];
#pragma unroll
        for (int r = 0; r < kMoGrp; ++r)
#pragma unroll
            for (int u = 0; u < kCiGrp; ++u) acc[r][u] = 0.f;
        const int p_end = min(np, (ps + 1) * pper);
        for (int pp = ps * pper; pp < p_end; ++pp) {
            float gv[kMoGrp][3];
#pragma unroll
            for (int r = 0; r < kMoGrp; ++r) { gv[r][0] = g[pp * kPbStride + 3 * r]; gv[r][1] = g[pp * kPbStride + 3 * r + 1]; gv[r][2] = g[pp * kPbStride + 3 * r + 2]; }
#pragma unroll
            for (int u = 0; u < kCiGrp; ++u) {
                if (c0 + u < C) {
                    const float* xr = xs + (c0 + u) * 3 * kPbXs + pp;
                    const float x0 = xr[0], x1 = xr[kPbXs], x2 = xr[2 * kPbXs];
#pragma unroll
                    for (int r = 0; r < kMoGrp; ++r) acc[r][u] = fmaf(gv[r][0], x0, fmaf(gv[r][1], x1, fmaf(gv[r][2], x2, acc[r][u])));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kMoGrp; ++r)
#pragma unroll
            for (int u = 0; u < kCiGrp; ++u)
                if (c0 + u < C) atomicAdd(dws + (mo0 + r) * C + c0 + u, acc[r][u]);
    }
    }
    // one 128-bit reduction per four entries and CTA (84 C is a multiple of 4; global atomics on these 7 KB were the limit of the
    // kernel: one scalar atomic per entry and 64-point block, 0.9 M per call, cost about half of its 86 us; 71 us now)
    __syncthreads();
    for (int i = threadIdx.x; i < kVO * C; i += blockDim.x)
        atomicAdd(reinterpret_cast<float4*>(dW4) + i, reinterpret_cast<const float4*>(dws)[i]);
}

// ---- per-edge machinery ---------------------------------------------------------------------------------------------
// One thread per edge; a tile is P consecutive points x their k edges (P*k <= kMaxTileThreads), two CTAs per SM.
// Shared memory of a CTA: packed coefficients | U rows of the tile's edges (cp.async, 528-byte stride: conflict-free 128-bit
// reads of a thread's own row) | V rows of the tile's points | upstream-gradient rows (backward) | the k-reduction buffer.
constexpr int kMaxTileThreads = 160;
constexpr int kRowS = 132;              // shared-memory stride (floats) of a staged 128-float row
constexpr int kGrp = 7;                 // channels per k-reduction round in the forward (21 floats per thread)
constexpr int kRedS = 25;               // stride of the reduction buffer: up to 24 floats per thread and round, odd

struct EdgeId {
    long long g;        // global point index b*N + n of the centre
    long long m;        // global point index of the neighbour
    long long b;        // cloud
    int n, mloc;        // centre / neighbour index inside the cloud
    int pl, j;          // point slot inside the tile, edge slot of the point
    bool valid;
};

__device__ __forceinline__ EdgeId edge_of(long long tile, int P, int k, long long BN, int N, const long long* __restrict__ idx) {
    EdgeId e;
    e.pl = threadIdx.x / k;
    e.j = threadIdx.x - e.pl * k;
    e.g = tile * P + e.pl;
    e.valid = e.pl < P && e.g < BN;
    e.m = e.b = 0;
    e.n = e.mloc = 0;
    if (e.valid) {
        e.b = e.g / N;
        e.n = (int)(e.g - e.b * N);
        e.mloc = (int)idx[e.g * k + e.j];
        e.m = e.b * N + e.mloc;
    } else {
        e.pl = 0;
    }
    return e;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

struct TileSmem {
    float* ws;          // [kWFloats]
    float* urows;       // [T][kRowS]   (not used by C = 1 layers)
    float* vrows;       // [P][kRowS]
    float* gs;          // [P][64] upstream gradient rows (backward)
    float* red;         // [T][kRedS]
    float* acc;         // backward accumulators
};

// urows: per-thread 528-byte slots (U rows; the backward kernels also stage their per-edge gradient rows there, so they carve
// them for C = 1 layers too); vrows only when U/V rows are used; gs for the backward; red when the kernel reduces over k
__device__ __forceinline__ TileSmem carve(float* base, int T, int P, bool urows, bool vrows, bool gs, bool red) {
    TileSmem s;
    s.ws = base;
    s.urows = s.ws + kWFloats;
    s.vrows = s.urows + (urows ? T * kRowS : 0);
    s.gs = s.vrows + (vrows ? P * kRowS : 0);
    s.red = s.gs + (gs ? P * 64 : 0);
    s.acc = s.red + (red ? T * kRedS : 0);
    return s;
}

static size_t tile_smem_bytes(int T, int P, bool urows, bool vrows, bool gs, bool red, int acc_floats) {
    size_t f = kWFloats + (urows ? (size_t)T * kRowS : 0) + (vrows ? (size_t)P * kRowS : 0) + (gs ? (size_t)P * 64 : 0) +
               (red ? (size_t)T * kRedS : 0) + acc_floats;
    return f * sizeof(float);
}

// Stage the tile's inputs: every thread copies the 512-byte U row of its neighbour, the V rows of the tile's points are copied
// cooperatively; threads without an edge zero their row so that everything downstream stays finite.
template <bool DIRECT>
__device__ __forceinline__ void stage_tile(const TileSmem& S, const float* __restrict__ UU, const float* __restrict__ VV, const EdgeId& e,
                                           long long tile, int P, long long BN) {
    if (DIRECT) return;
    // warp-cooperative: 8 lanes copy one 128-byte line of one row, so an instruction touches 4 cache lines instead of the 32 a
    // thread-per-row copy would (the L1 tag stage, not L2 bandwidth, limits a 32-line gather instruction)
    const int lane = threadIdx.x & 31;
    const long long mrow = e.valid ? e.m : -1;
    float* wbase = S.urows + (threadIdx.x & ~31) * kRowS;
#pragma unroll 8
    for (int s = 0; s < 32; ++s) {
        const int r = (lane >> 3) + 4 * (s >> 2), ch = (lane & 7) + 8 * (s & 3);
        const long long m = __shfl_sync(kFull, mrow, r);
        float* dst = wbase + r * kRowS + 4 * ch;
        if (m >= 0) cp_async16(dst, UU + m * kRowF + 4 * ch);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < P * 32; i += blockDim.x) {
        const int pl = i >> 5, q = i & 31;
        const long long g = tile * P + pl;
        if (g < BN) cp_async16(S.vrows + pl * kRowS + 4 * q, VV + g * kRowF + 4 * q);
        else reinterpret_cast<float4*>(S.vrows + pl * kRowS)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_wait_all();
}

// Calls f(o, p0, p1, p2, d0, d1, d2) for the 21 channels of the first conv's Linear outputs of this thread's edge.
template <bool DIRECT, typename F>
__device__ __forceinline__ void stage1_inputs(const TileSmem& S, const float* __restrict__ x, int N, const EdgeId& e, F&& f) {
    if (DIRECT) {                                        // p = Wa (x_j - x_i) + Wb x_i from the coordinates x[B,1,3,N]
        const float* xb = x + e.b * 3 * N;
        float xi[3], dx[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            xi[c] = e.valid ? __ldg(xb + (size_t)c * N + e.n) : 0.f;
            dx[c] = e.valid ? __ldg(xb + (size_t)c * N + e.mloc) - xi[c] : 0.f;
        }
#pragma unroll
        for (int o = 0; o < kVO; ++o) {
            const float waf = S.ws[kOffW1 + o], wad = S.ws[kOffW1 + kVO + o], wbf = S.ws[kOffW1 + 2 * kVO + o], wbd = S.ws[kOffW1 + 3 * kVO + o];
            f(o, fmaf(waf, dx[0], wbf * xi[0]), fmaf(waf, dx[1], wbf * xi[1]), fmaf(waf, dx[2], wbf * xi[2]),
              fmaf(wad, dx[0], wbd * xi[0]), fmaf(wad, dx[1], wbd * xi[1]), fmaf(wad, dx[2], wbd * xi[2]));
        }
    } else {                                             // p = U[j] + V[i]: four channels (3 x 128 bits per half row) at a time
        const float4* u = reinterpret_cast<const float4*>(S.urows + threadIdx.x * kRowS);
        const float4* v = reinterpret_cast<const float4*>(S.vrows + e.pl * kRowS);
#pragma unroll
        for (int g4 = 0; g4 < 6; ++g4) {
            float pp[12], dd[12];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (g4 < 5 || q == 0) {
                    const float4 a = u[3 * g4 + q], b = v[3 * g4 + q], c = u[16 + 3 * g4 + q], d = v[16 + 3 * g4 + q];
                    pp[4 * q] = a.x + b.x; pp[4 * q + 1] = a.y + b.y; pp[4 * q + 2] = a.z + b.z; pp[4 * q + 3] = a.w + b.w;
                    dd[4 * q] = c.x + d.x; dd[4 * q + 1] = c.y + d.y; dd[4 * q + 2] = c.z + d.z; dd[4 * q + 3] = c.w + d.w;
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int o = 4 * g4 + jj;
                if (o < kVO) f(o, pp[3 * jj], pp[3 * jj + 1], pp[3 * jj + 2], dd[3 * jj], dd[3 * jj + 1], dd[3 * jj + 2]);
            }
        }
    }
}

// Sum over the k threads of every point: each thread has put `width` floats at red[tid*kRedS ..]; emit(pl, i, sum).
template <typename Emit>
__device__ __forceinline__ void reduce_round(const TileSmem& S, int P, int k, int width, bool point_major, Emit emit) {
    __syncthreads();
    for (int q = threadIdx.x; q < P * width; q += blockDim.x) {
        const int pl = point_major ? q / width : q % P;
        const int i = point_major ? q % width : q / P;
        const float* src = S.red + pl * k * kRedS + i;
        float s = 0.f;
        for (int j = 0; j < k; ++j) s += src[j * kRedS];
        emit(pl, i, s);
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
__global__ void __launch_bounds__(kMaxTileThreads, 2) edgeconv_fwd_kernel(const FwdArgs A) {
    extern __shared__ __align__(16) float smem_f[];
    const TileSmem S = carve(smem_f, blockDim.x, A.P, !DIRECT, !DIRECT, false, true);
    stage_weights(S.ws, A.Wdev);
    const float* ws = S.ws;
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const int lane = threadIdx.x & 31;
    float sr[kVO], sr2[kVO];
    if (MODE != 2) {
#pragma unroll
        for (int o = 0; o < kVO; ++o) sr[o] = sr2[o] = 0.f;
    }
    const float inv_k = 1.0f / A.k;
    float* myred = S.red + threadIdx.x * kRedS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        __syncthreads();                                   // the previous tile's rows are no longer read
        stage_tile<DIRECT>(S, A.UU, A.VV, e, tile, A.P, A.BN);
        __syncthreads();
        if (MODE == 0) {
            stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float, float, float) {
                const float r = vnorm(p0, p1, p2) + kVnEps;
                if (e.valid) { sr[o] += r; sr2[o] = fmaf(r, r, sr2[o]); }
            });
            continue;
        }
        auto emit_out = [&](int c0) {
            return [&, c0](int pl, int i, float s) {
                const long long g = tile * A.P + pl;
                if (g < A.BN) A.out[((g / A.N) * kVD + c0 + i) * A.N + g % A.N] = s * inv_k;
            };
        };
        auto emit_rows = [&](float* dst, int c0) {
            return [&, dst, c0](int pl, int i, float s) {
                const long long g = tile * A.P + pl;
                if (g < A.BN) dst[g * kVD + c0 + i] = s;
            };
        };
        const bool want_y = MODE == 2 && A.ysum != nullptr;
        if (STAGES == 1) {
            // one conv: outputs (and the BatchNorm-backward coefficients) leave in rounds of kGrp channels (rounds of 4 -- the
            // granularity the staged rows are read in -- were 18 reductions per tile instead of 9: 198 -> 186 us)
            float ob[3 * kGrp], yb[3 * kGrp], rb[3 * kGrp];
            stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float d0, float d1, float d2) {
                const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
                const int jj = o % kGrp;
                ob[3 * jj] = f.o0; ob[3 * jj + 1] = f.o1; ob[3 * jj + 2] = f.o2;
                if (want_y) {
                    chan_y(f, p0, p1, p2, d0, d1, d2, yb[3 * jj], yb[3 * jj + 1], yb[3 * jj + 2]);
                    const float rh = (f.r - bn_coef(ws, 0, BN_MU, o)) * bn_coef(ws, 0, BN_RSTD, o);
                    rb[3 * jj] = yb[3 * jj] * rh; rb[3 * jj + 1] = yb[3 * jj + 1] * rh; rb[3 * jj + 2] = yb[3 * jj + 2] * rh;
                }
                if (jj == kGrp - 1) {
                    const int c0 = 3 * (o - jj);
#pragma unroll
                    for (int i = 0; i < 3 * kGrp; ++i) myred[i] = e.valid ? ob[i] : 0.f;
                    reduce_round(S, A.P, A.k, 3 * kGrp, false, emit_out(c0));
                    if (want_y) {
#pragma unroll
                        for (int i = 0; i < 3 * kGrp; ++i) myred[i] = e.valid ? yb[i] : 0.f;
                        reduce_round(S, A.P, A.k, 3 * kGrp, true, emit_rows(A.ysum, c0));
#pragma unroll
                        for (int i = 0; i < 3 * kGrp; ++i) myred[i] = e.valid ? rb[i] : 0.f;
                        reduce_round(S, A.P, A.k, 3 * kGrp, true, emit_rows(A.yrsum, c0));
                    }
                }
            });
            continue;
        }
        // ---- two convs: stage 1 -> O1 in registers -----------------------------------------------------------------------
        f32x2 X2[3][kXP];
        {
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
            stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float d0, float d1, float d2) {
                const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
                if (o & 1) { X2[0][o >> 1] = pack2(h0, f.o0); X2[1][o >> 1] = pack2(h1, f.o1); X2[2][o >> 1] = pack2(h2, f.o2); }
                else if (o == kVO - 1) { X2[0][o >> 1] = pack2(f.o0, 0.f); X2[1][o >> 1] = pack2(f.o1, 0.f); X2[2][o >> 1] = pack2(f.o2, 0.f); }
                else { h0 = f.o0; h1 = f.o1; h2 = f.o2; }
            });
        }
        if (MODE == 1) {
#pragma unroll
            for (int o = 0; o < kVO; ++o) {
                float p0, p1, p2, d0, d1, d2;
                mix_channel<false>(ws, o, X2, p0, p1, p2, d0, d1, d2);
                const float r = vnorm(p0, p1, p2) + kVnEps;
                if (e.valid) { sr[o] += r; sr2[o] = fmaf(r, r, sr2[o]); }
            }
            continue;
        }
        // ---- stage 2, kGrp output channels per k-reduction round (weights from shared memory, 128 bits at a time).  The rounds are
        // a real loop: unrolling all 21 channels made the kernel 130 KB of straight-line code, and per-channel rounds (a ~330-
        // instruction body) cost more in barriers than they save in instruction-cache misses (measured: 328 / 295 / 358 us)
#pragma unroll 1
        for (int g7 = 0; g7 < kVO / kGrp; ++g7) {
            float yb[3 * kGrp], rb[3 * kGrp];
#pragma unroll
            for (int jj = 0; jj < kGrp; ++jj) {
                const int o = g7 * kGrp + jj;
                float p0, p1, p2, d0, d1, d2;
                mix_channel<true>(ws, o, X2, p0, p1, p2, d0, d1, d2);
                const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_B, o));
                myred[3 * jj] = e.valid ? f.o0 : 0.f; myred[3 * jj + 1] = e.valid ? f.o1 : 0.f; myred[3 * jj + 2] = e.valid ? f.o2 : 0.f;
                if (want_y) {
                    chan_y(f, p0, p1, p2, d0, d1, d2, yb[3 * jj], yb[3 * jj + 1], yb[3 * jj + 2]);
                    const float rh = (f.r - bn_coef(ws, 1, BN_MU, o)) * bn_coef(ws, 1, BN_RSTD, o);
                    rb[3 * jj] = yb[3 * jj] * rh; rb[3 * jj + 1] = yb[3 * jj + 1] * rh; rb[3 * jj + 2] = yb[3 * jj + 2] * rh;
                }
            }
            reduce_round(S, A.P, A.k, 3 * kGrp, false, emit_out(3 * kGrp * g7));
            if (want_y) {
#pragma unroll
                for (int i = 0; i < 3 * kGrp; ++i) myred[i] = e.valid ? yb[i] : 0.f;
                reduce_round(S, A.P, A.k, 3 * kGrp, true, emit_rows(A.ysum, 3 * kGrp * g7));
#pragma unroll
                for (int i = 0; i < 3 * kGrp; ++i) myred[i] = e.valid ? rb[i] : 0.f;
                reduce_round(S, A.P, A.k, 3 * kGrp, true, emit_rows(A.yrsum, 3 * kGrp * g7));
            }
        }
    }
    if (MODE != 2) {                                     // block reduction of the 42 partial sums, fp64 from the warp level up
        __syncthreads();
        double* dsm = reinterpret_cast<double*>(S.red);
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

// 16 values per lane -> lanes l and l+16 both receive the warp-wide sum of v[l]
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const float keep = up ? v[j + s] : v[j];
            const float send = up ? v[j] : v[j + s];
            v[j] = keep + __shfl_xor_sync(kFull, send, s);
        }
    }
    return v[0] + __shfl_xor_sync(kFull, v[0], 16);
}

// ---- backward, stage 2 (two-conv layers) --------------------------------------------------------------------------------
struct Bwd2Args {
    const float* UU; const float* VV; const long long* idx; const float* Wdev; const float* x;
    long long BN; int N; int k; int P;
    const float* G;         // [B,21,3,N] gradient of the layer output
    float* gO1;             // [E][64] out: gradient wrt the first conv's output of every edge
    float* dW2;             // [21][2][21] = [out o][half: feat, dir][in i] += (atomics at kernel end)
    double* stats1;         // [21][2] += sum gy1, sum gy1 rhat1
};
constexpr int kDwEntries = 2 * kVO * kVO;                 // 882
constexpr int kDwRow = 48;                                // per output channel: [half][22 in (21 + one zero)] + 4 padding = 3 groups of 16
constexpr int kDwPad = kVO * kDwRow;                      // 1008

template <bool DIRECT>
__global__ void __launch_bounds__(kMaxTileThreads, 2) edgeconv_bwd2_kernel(const Bwd2Args A) {
    extern __shared__ __align__(16) float smem_f[];
    const TileSmem S = carve(smem_f, blockDim.x, A.P, !DIRECT, !DIRECT, true, false);
    float* dw_s = S.acc;                                 // [896] weight-gradient partial sums of this CTA
    float* st_s = S.acc + kDwPad;                        // [42] stage-1 BatchNorm sums of this CTA
    for (int i = threadIdx.x; i < kDwPad + 2 * kVO; i += blockDim.x) S.acc[i] = 0.f;
    stage_weights(S.ws, A.Wdev);
    const float* ws = S.ws;
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const int lane = threadIdx.x & 31;
    const float inv_k = 1.0f / A.k;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        __syncthreads();
        stage_tile<DIRECT>(S, A.UU, A.VV, e, tile, A.P, A.BN);
        for (int q = threadIdx.x; q < A.P * kVD; q += blockDim.x) {           // upstream gradient rows of the tile's points
            const int pl = q % A.P, oc = q / A.P;
            const long long g = tile * A.P + pl;
            S.gs[pl * 64 + oc] = g < A.BN ? A.G[((g / A.N) * kVD + oc) * A.N + g % A.N] * inv_k : 0.f;
        }
        __syncthreads();
        f32x2 X2[3][kXP];                                                      // O1: output of the first conv, channel pairs
        {
            float h0 = 0.f, h1 = 0.f, h2 = 0.f;
            stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float d0, float d1, float d2) {
                const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
                if (o & 1) { X2[0][o >> 1] = pack2(h0, f.o0); X2[1][o >> 1] = pack2(h1, f.o1); X2[2][o >> 1] = pack2(h2, f.o2); }
                else if (o == kVO - 1) { X2[0][o >> 1] = pack2(f.o0, 0.f); X2[1][o >> 1] = pack2(f.o1, 0.f); X2[2][o >> 1] = pack2(f.o2, 0.f); }
                else { h0 = f.o0; h1 = f.o1; h2 = f.o2; }
            });
        }
        f32x2 GO2[3][kXP];                                                     // gO1 = W2^T g, accumulated over output channels
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int m = 0; m < kXP; ++m) GO2[c][m] = 0ull;
        const float* grow = S.gs + e.pl * 64;
        float stag[16];
#pragma unroll 1                                                               // a real loop: the body is ~700 instructions; unrolled 21 times
        for (int o = 0; o < kVO; ++o) {                                        // it thrashes the instruction cache (1.56 ms; by 3: 1.16; 1: 0.91)
            float p0, p1, p2, d0, d1, d2;
            mix_channel<true>(ws, o, X2, p0, p1, p2, d0, d1, d2);
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_B, o));
            float gp0, gp1, gp2, gd0, gd1, gd2, gy, rhat;
            chan_bwd(f, p0, p1, p2, d0, d1, d2, grow[3 * o], grow[3 * o + 1], grow[3 * o + 2], bn_coef(ws, 1, BN_A, o), bn_coef(ws, 1, BN_MU, o),
                     bn_coef(ws, 1, BN_RSTD, o), bn_coef(ws, 1, BN_S1M, o), bn_coef(ws, 1, BN_S2M, o), gp0, gp1, gp2, gd0, gd1, gd2, gy, rhat);
            if (!e.valid) { gp0 = gp1 = gp2 = gd0 = gd1 = gd2 = 0.f; }
            mix_channel_transposed(ws, o, gp0, gp1, gp2, gd0, gd1, gd2, GO2);
            // weight gradient: entry (out o, half h, in i) <- sum over edges of g_h[o] . O1[i]; the row of this output channel is
            // [h][22] (+4 padding) and goes through the warp transpose 16 entries at a time, after which lane l (< 16) owns entry
            // 16*group + l of the row.  (Scalar FMAs on the halves of the channel pairs: packed ones here, with the duplicated
            // gradient operands and the unpacking they need, measured slower -- 954 vs 865 us for the C=21 layer.)
#pragma unroll
            for (int q2 = 0; q2 < kDwRow / 2; ++q2) {
                const int h = q2 / kXP, m = q2 - h * kXP;
                float lo = 0.f, hi = 0.f;
                if (h < 2) {
                    float x0l, x0h, x1l, x1h, x2l, x2h;
                    unpack2(X2[0][m], x0l, x0h); unpack2(X2[1][m], x1l, x1h); unpack2(X2[2][m], x2l, x2h);
                    const float a0 = h == 0 ? gp0 : gd0, a1 = h == 0 ? gp1 : gd1, a2 = h == 0 ? gp2 : gd2;
                    lo = fmaf(a2, x2l, fmaf(a1, x1l, a0 * x0l));
                    hi = fmaf(a2, x2h, fmaf(a1, x1h, a0 * x0h));
                }
                stag[(2 * q2) & 15] = lo;
                stag[(2 * q2 + 1) & 15] = hi;
                if (((2 * q2 + 1) & 15) == 15) {
                    const float tot = warp_transpose_sum16(stag, lane);
                    if (lane < 16) atomicAdd(dw_s + o * kDwRow + ((2 * q2) & ~15) + lane, tot);
                }
            }
        }
        // unpack gO1 into the flat [3 i + c] order of the scratch (register renaming, no instructions)
        float GO[64];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int m = 0; m < kXP; ++m) {
                float lo, hi;
                unpack2(GO2[c][m], lo, hi);
                GO[3 * (2 * m) + c] = lo;
                if (2 * m + 1 < kVO) GO[3 * (2 * m + 1) + c] = hi;
            }
        GO[63] = 0.f;
        // BatchNorm sums of the first conv: gy1 = Y1 . gO1 per channel (its inputs are still staged in shared memory)
        stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float d0, float d1, float d2) {
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            float y0, y1, y2;
            chan_y(f, p0, p1, p2, d0, d1, d2, y0, y1, y2);
            float gy = e.valid ? fmaf(y2, GO[3 * o + 2], fmaf(y1, GO[3 * o + 1], y0 * GO[3 * o])) : 0.f;
            float gyr = gy * (f.r - bn_coef(ws, 0, BN_MU, o)) * bn_coef(ws, 0, BN_RSTD, o);
            gy = warp_sum(gy);
            gyr = warp_sum(gyr);
            if (lane == 0) { atomicAdd(st_s + 2 * o, gy); atomicAdd(st_s + 2 * o + 1, gyr); }
        });
        // gO1 scratch layout [tile][16 chunks][T threads] of 16 bytes: thread-per-edge stores (here) and loads (stage-1 backward)
        // are both fully coalesced, 512 contiguous bytes per warp instruction
        if (e.valid) {
            float4* dst = reinterpret_cast<float4*>(A.gO1) + tile * 16 * blockDim.x + threadIdx.x;
#pragma unroll
            for (int q = 0; q < 16; ++q)
                __stcs(dst + (size_t)q * blockDim.x, make_float4(GO[4 * q], GO[4 * q + 1], GO[4 * q + 2], q == 15 ? 0.f : GO[4 * q + 3]));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kDwEntries; i += blockDim.x) {                // (o, h, in) <- row o, entry h * 22 + in
        const int o = i / (2 * kVO), r = i - o * 2 * kVO, h = r / kVO;
        atomicAdd(A.dW2 + i, dw_s[o * kDwRow + h * 2 * kXP + (r - h * kVO)]);
    }
    for (int i = threadIdx.x; i < 2 * kVO; i += blockDim.x) atomicAdd(A.stats1 + i, (double)st_s[i]);
}

// ---- backward, stage 1 ------------------------------------------------------------------------------------------------------
struct Bwd1Args {
    const float* UU; const float* VV; const long long* idx; const float* Wdev; const float* x;
    long long BN; int N; int k; int P;
    const float* gO1;       // [E][64] (two-conv layers) or nullptr: then gO1 = G[n]/k (one-conv layers)
    const float* G;         // [B,21,3,N]
    float* gUU;             // [B*N][128] += (vector reductions; zeroed by the caller)
    float* gVV;             // [B*N][128]  = sum over the k edges of the point
};

template <bool DIRECT>
__global__ void __launch_bounds__(kMaxTileThreads, 2) edgeconv_bwd1_kernel(const Bwd1Args A) {
    extern __shared__ __align__(16) float smem_f[];
    const TileSmem S = carve(smem_f, blockDim.x, A.P, true, !DIRECT, true, true);
    stage_weights(S.ws, A.Wdev);
    const float* ws = S.ws;
    const long long ntiles = (A.BN + A.P - 1) / A.P;
    const float inv_k = 1.0f / A.k;
    const int lane = threadIdx.x & 31;
    float* myred = S.red + threadIdx.x * kRedS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const EdgeId e = edge_of(tile, A.P, A.k, A.BN, A.N, A.idx);
        __syncthreads();
        stage_tile<DIRECT>(S, A.UU, A.VV, e, tile, A.P, A.BN);
        if (!A.gO1) {
            for (int q = threadIdx.x; q < A.P * kVD; q += blockDim.x) {
                const int pl = q % A.P, oc = q / A.P;
                const long long g = tile * A.P + pl;
                S.gs[pl * 64 + oc] = g < A.BN ? A.G[((g / A.N) * kVD + oc) * A.N + g % A.N] * inv_k : 0.f;
            }
        }
        __syncthreads();
        float GO[64];
        if (A.gO1) {                                                          // chunk-major scratch: coalesced (see edgeconv_bwd2_kernel)
            const float4* gsrc = reinterpret_cast<const float4*>(A.gO1) + tile * 16 * blockDim.x + threadIdx.x;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float4 v = e.valid ? __ldcs(gsrc + (size_t)q * blockDim.x) : make_float4(0.f, 0.f, 0.f, 0.f);
                GO[4 * q] = v.x; GO[4 * q + 1] = v.y; GO[4 * q + 2] = v.z; GO[4 * q + 3] = v.w;
            }
        } else {
            const float* grow = S.gs + e.pl * 64;
#pragma unroll
            for (int i = 0; i < kVD; ++i) GO[i] = e.valid ? grow[i] : 0.f;
            GO[63] = 0.f;
        }
        float4* slot = reinterpret_cast<float4*>(S.urows + threadIdx.x * kRowS);   // own row: its inputs are consumed group by group,
        float pb[12], db[12];                                                       // the gradient row (gP | gD) replaces them in place
        stage1_inputs<DIRECT>(S, A.x, A.N, e, [&](int o, float p0, float p1, float p2, float d0, float d1, float d2) {
            const ChanFwd f = chan_fwd(p0, p1, p2, d0, d1, d2, bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_B, o));
            const int jj = o & 3;
            float gy, rhat;
            chan_bwd(f, p0, p1, p2, d0, d1, d2, GO[3 * o], GO[3 * o + 1], GO[3 * o + 2], bn_coef(ws, 0, BN_A, o), bn_coef(ws, 0, BN_MU, o),
                     bn_coef(ws, 0, BN_RSTD, o), bn_coef(ws, 0, BN_S1M, o), bn_coef(ws, 0, BN_S2M, o), pb[3 * jj], pb[3 * jj + 1], pb[3 * jj + 2],
                     db[3 * jj], db[3 * jj + 1], db[3 * jj + 2], gy, rhat);
            if (jj == 3 || o == kVO - 1) {                                    // four channels = 3 x 128 bits of each half row
                const int g4 = o >> 2, nq = jj == 3 ? 3 : 1;
                if (jj != 3) { pb[3] = db[3] = 0.f; }
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    if (q < nq) {                                             // neighbour half, parked for the scatter below
                        slot[3 * g4 + q] = make_float4(pb[4 * q], pb[4 * q + 1], pb[4 * q + 2], pb[4 * q + 3]);
                        slot[16 + 3 * g4 + q] = make_float4(db[4 * q], db[4 * q + 1], db[4 * q + 2], db[4 * q + 3]);
                    }
                }
                const int width = 4 * nq;                                     // centre half: gV[g] = sum over the point's k edges
#pragma unroll
                for (int i = 0; i < 12; ++i) if (i < width) { myred[i] = e.valid ? pb[i] : 0.f; myred[12 + i] = e.valid ? db[i] : 0.f; }
                reduce_round(S, A.P, A.k, 24, true, [&](int pl, int i, float s) {
                    const long long g = tile * A.P + pl;
                    const int half = i / 12, w = i - 12 * half;
                    if (g < A.BN && w < width) A.gVV[g * kRowF + 64 * half + 12 * g4 + w] = s;
                });
            }
        });
        // neighbour half: gU[m] += (gP | gD), warp-cooperative like the gather -- 8 lanes add one 128-byte line of one row, so the
        // 128-bit reductions of an instruction land in 4 lines instead of 32
        {
            __syncwarp();
            const long long mrow = e.valid ? e.m : -1;
            const float* wbase = S.urows + (threadIdx.x & ~31) * kRowS;
#pragma unroll 8
            for (int s = 0; s < 32; ++s) {
                const int r = (lane >> 3) + 4 * (s >> 2), ch = (lane & 7) + 8 * (s & 3);
                const long long m = __shfl_sync(kFull, mrow, r);
                if (m >= 0) atomicAdd(reinterpret_cast<float4*>(A.gUU + m * kRowF) + ch, *reinterpret_cast<const float4*>(wbase + r * kRowS + 4 * ch));
            }
        }
    }
}

static int tile_points(int k) {
    int P = kMaxTileThreads / k;
    if (P > 32) P = 32;
    return P < 1 ? 1 : P;
}
static int tile_threads(int P, int k) { return (P * k + 31) / 32 * 32; }   // whole warps: the kernels shuffle

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
    const size_t smem = sizeof(float) * ((size_t)3 * C * kPlPts + 4 * kVO * C + 2 * kPlPts * kPlStride);
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(vn_point_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    vn_point_linear_kernel<<<dim3((N + kPlPts - 1) / kPlPts, B), 256, smem, as_stream(stream)>>>(x, W4, C, N, UU, VV);
    return check_launch("vn_point_linear_kernel");
}

extern "C" int hpcs_vn_point_linear_bwd_f32(const float* gUU, const float* gVV, const float* x, const float* W4, int B, int C, int N,
                                            float* gx, float* dW4, void* stream) {
    if (!gUU || !gVV || !x || !W4 || !gx || !dW4) return fail(HPCS_ERR_ARG, "vn_point_linear_bwd: null pointer");
    if (B <= 0 || B > 65535 || C < 1 || C > 64 || N <= 0) return fail(HPCS_ERR_ARG, "vn_point_linear_bwd: bad shape B=%d C=%d N=%d", B, C, N);
    const size_t smem = sizeof(float) * ((size_t)2 * kPbPts * kPbStride + (size_t)3 * C * kPbXs + 2 * 4 * kVO * C);
    if ((reinterpret_cast<uintptr_t>(dW4) & 15) != 0) return fail(HPCS_ERR_ARG, "vn_point_linear_bwd: dW4 must be 16-byte aligned");
    static thread_local size_t attr = 0;
    if (smem > 48 * 1024 && smem > attr) {
        cudaFuncSetAttribute(vn_point_linear_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr = smem;
    }
    const int nblk = B * ((N + kPbPts - 1) / kPbPts);
    const int grid = nblk < 2 * sm_count() ? nblk : 2 * sm_count();
    vn_point_linear_bwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(gUU, gVV, x, W4, C, N, nblk, gx, dW4);
    return check_launch("vn_point_linear_bwd_kernel");
}

extern "C" int hpcs_edgeconv_bn_fold_f32(const double* stats, int64_t M, const float* gamma, const float* beta, float* running_mean,
                                         float* running_var, const int64_t* num_batches_tracked, float momentum, float eps, int training,
                                         float* coef, int stage, void* stream) {
    if (!gamma || !beta || !coef || (training && !stats) || (!training && (!running_mean || !running_var)) || stage < 0 || stage > 1 || M <= 0)
        return fail(HPCS_ERR_ARG, "edgeconv_bn_fold: bad arguments");
    edgeconv_bn_fold_kernel<<<1, 32, 0, as_stream(stream)>>>(stats, (double)M, gamma, beta, running_mean, running_var,
                                                             reinterpret_cast<const long long*>(num_batches_tracked), momentum, eps, training, coef, stage);
    return check_launch("edgeconv_bn_fold_kernel");
}

extern "C" int hpcs_edgeconv_bn_sums_f32(const float* G, const float* ysum, const float* yrsum, int B, int N, int k, int training, double* sums,
                                         float* coef, int stage, float* dgamma, float* dbeta, void* stream) {
    if (!G || !ysum || !yrsum || !sums || !coef || !dgamma || !dbeta || B <= 0 || N <= 0 || k <= 0 || stage < 0 || stage > 1)
        return fail(HPCS_ERR_ARG, "edgeconv_bn_sums: bad arguments");
    cudaStream_t st = as_stream(stream);
    cudaError_t ce = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * kVO, st);
    if (ce != cudaSuccess) return fail(HPCS_ERR_CUDA, "edgeconv_bn_sums: memset: %s", cudaGetErrorString(ce));
    edgeconv_bn_sums_kernel<<<(B * kVD + 7) / 8, 256, 0, st>>>(G, ysum, yrsum, B, N, 1.0f / k, sums);
    if (int rc = check_launch("edgeconv_bn_sums_kernel")) return rc;
    edgeconv_bn_sums_finish_kernel<<<1, 32, 0, st>>>(sums, (double)B * N * k, training, coef, stage, dgamma, dbeta);
    return check_launch("edgeconv_bn_sums_finish_kernel");
}

/* same finish step for sums that a kernel already accumulated (the first conv's, from hpcs_edgeconv_bwd_stage2_f32) */
extern "C" int hpcs_edgeconv_bn_sums_finish_f32(const double* sums, int64_t M, int training, float* coef, int stage, float* dgamma, float* dbeta,
                                                void* stream) {
    if (!sums || !coef || !dgamma || !dbeta || stage < 0 || stage > 1 || M <= 0) return fail(HPCS_ERR_ARG, "edgeconv_bn_sums_finish: bad arguments");
    edgeconv_bn_sums_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(sums, (double)M, training, coef, stage, dgamma, dbeta);
    return check_launch("edgeconv_bn_sums_finish_kernel");
}

template <typename K>
static void opt_in_smem(K kernel, size_t smem) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

extern "C" int hpcs_edgeconv_coef_floats(void) { return kWFloats; }

extern "C" size_t hpcs_edgeconv_scratch_floats(int B, int N, int k) {
    if (B <= 0 || N <= 0 || k <= 0 || k > kMaxTileThreads) return 0;
    const int P = tile_points(k);
    const long long tiles = ((long long)B * N + P - 1) / P;
    return (size_t)tiles * tile_threads(P, k) * 64;
}

extern "C" int hpcs_edgeconv_fwd_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k, int stages,
                                     const float* coef, const float* x_direct, int mode, double* stats, float* out, float* ysum,
                                     float* yrsum, void* stream) {
    if (!idx || !coef || (!x_direct && (!UU || !VV))) return fail(HPCS_ERR_ARG, "edgeconv_fwd: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || k > kMaxTileThreads || (stages != 1 && stages != 2) || mode < 0 || mode > 2 || (mode == 1 && stages == 1))
        return fail(HPCS_ERR_ARG, "edgeconv_fwd: bad arguments B=%d N=%d k=%d stages=%d mode=%d", B, N, k, stages, mode);
    if ((mode != 2 && !stats) || (mode == 2 && !out) || ((ysum == nullptr) != (yrsum == nullptr)))
        return fail(HPCS_ERR_ARG, "edgeconv_fwd: output pointers do not match mode %d", mode);
    FwdArgs A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), stats, out, ysum, yrsum};
    const int T = tile_threads(A.P, k);
    const size_t smem = tile_smem_bytes(T, A.P, x_direct == nullptr, x_direct == nullptr, false, true, 0);
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
    if (B <= 0 || N <= 0 || k <= 0 || k > kMaxTileThreads) return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage2: bad shape");
    Bwd2Args A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), G, gO1, dW2, stats1};
    const int T = tile_threads(A.P, k);
    const size_t smem = tile_smem_bytes(T, A.P, x_direct == nullptr, x_direct == nullptr, true, false, kDwPad + 2 * kVO + 2);
    if (x_direct) {
        opt_in_smem(edgeconv_bwd2_kernel<true>, smem);
        edgeconv_bwd2_kernel<true><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    } else {
        opt_in_smem(edgeconv_bwd2_kernel<false>, smem);
        edgeconv_bwd2_kernel<false><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    }
    return check_launch("edgeconv_bwd2_kernel");
}

extern "C" int hpcs_edgeconv_bwd_stage1_f32(const float* UU, const float* VV, const int64_t* idx, int B, int N, int k,
                                            const float* coef, const float* x_direct, const float* gO1, const float* G,
                                            float* gUU, float* gVV, void* stream) {
    if (!idx || !coef || !gUU || !gVV || (!gO1 && !G) || (!x_direct && (!UU || !VV)))
        return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage1: null pointer");
    if (B <= 0 || N <= 0 || k <= 0 || k > kMaxTileThreads) return fail(HPCS_ERR_ARG, "edgeconv_bwd_stage1: bad shape");
    Bwd1Args A{UU, VV, reinterpret_cast<const long long*>(idx), coef, x_direct, (long long)B * N, N, k, tile_points(k), gO1, G, gUU, gVV};
    const int T = tile_threads(A.P, k);
    const size_t smem = tile_smem_bytes(T, A.P, true, x_direct == nullptr, true, true, 0);
    if (x_direct) {
        opt_in_smem(edgeconv_bwd1_kernel<true>, smem);
        edgeconv_bwd1_kernel<true><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    } else {
        opt_in_smem(edgeconv_bwd1_kernel<false>, smem);
        edgeconv_bwd1_kernel<false><<<edge_grid(A.BN, A.P), T, smem, as_stream(stream)>>>(A);
    }
    return check_launch("edgeconv_bwd1_kernel");
}
