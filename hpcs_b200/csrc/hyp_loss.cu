// Fused Poincare-ball triplet objective: similarities, easy/semihard filter, hyperbolic-LCA
// distances, softmax-weighted HypHC loss and its gradient, in one pass over the mined triplets.
//
// Replaces MetricHyperbolicLoss.compute_hyp (hpcs/loss/ultrametric_loss.py:57-93), the filter of
// RandomTripletMarginMiner.mine (hpcs/miner/triplet_margin_miner.py:16-38) and three hyp_lca calls
// (hpcs/distances/lca.py:37-52).  The reference builds the dense (B.N)^2 similarity matrix twice
// (4.29 GB each at B.N = 32768) and reads 3 scalars per triplet plus the mean from it; here the
// matrix never exists:
//   * rows are L2-normalised once into a table u[n][DP] (4 MB at n=32768, D=32: L2 resident);
//   * per triplet the three cosines are dot products of table rows;
//   * mean(mat_sim) = 0.5 (1 + |sum_i u_i|^2 / n^2) exactly (the sum runs over all ordered pairs
//     including the diagonal, like torch.mean over the full matrix).
// Work split inside a warp: LPT = DP/4 lanes cooperate on one triplet for the dot products (each
// lane holds a float4 of every row, so a warp-wide load touches whole 128-byte lines), then every
// lane does the scalar transcendental part for ONE of the 32 triplets of the block, then the
// LPT-lane groups scatter the gradient rows with 128-bit vector reductions (red.global.add.v4.f32).
#include "common.cuh"
#include "hyp_math.cuh"

namespace hpcs {

struct HypHeader {              // lives at the start of the workspace
    double S[128];              // column sums of the unit rows (padded dims are zero)
    double loss_sum;            // sum over kept triplets of `total`
    double gs_sum;              // sum over kept triplets of d total / d s
    unsigned long long kept;
    float s_used;               // clamp(scale, 1e-4, 1)
    int DP;
};

struct HypLayout {
    HypHeader* hdr;
    float* u;                   // [n][DP] unit rows
    float* invn;                // [n] 1 / max(|x_i|, eps)
    float* G;                   // [n][DP] d(sum of totals)/d u, unscaled
    size_t bytes;
};

static int padded_dim(int D) {
    int dp = 4;
    while (dp < D) dp <<= 1;
    return dp;
}

static HypLayout hyp_layout(void* ws, int64_t n, int D) {
    const int DP = padded_dim(D);
    HypLayout L;
    char* p = static_cast<char*>(ws);
    size_t off = 0;
    L.hdr = reinterpret_cast<HypHeader*>(p + off); off += align_up(sizeof(HypHeader), 256);
    L.u = reinterpret_cast<float*>(p + off);       off += align_up((size_t)n * DP * sizeof(float), 256);
    L.invn = reinterpret_cast<float*>(p + off);    off += align_up((size_t)n * sizeof(float), 256);
    L.G = reinterpret_cast<float*>(p + off);       off += align_up((size_t)n * DP * sizeof(float), 256);
    L.bytes = off;
    return L;
}

// ---- prep: unit rows, inverse norms, column sums ----------------------------------------------------
template <int LPT>
__global__ void __launch_bounds__(256)
hyp_prep_kernel(const float* __restrict__ x, int64_t n, int D, const float* __restrict__ scale,
                float* __restrict__ u, float* __restrict__ invn, HypHeader* __restrict__ hdr) {
    constexpr int DP = LPT * 4;
    constexpr int RPB = 256 / LPT;                  // rows per block iteration
    __shared__ float red[256][4];
    const int sub = threadIdx.x % LPT, grp = threadIdx.x / LPT;
    float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);
    // the trip count is uniform per block (every lane takes part in the shuffles); rows past n are masked
    for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < n; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + grp;
        const bool live = row < n;
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int d = sub * 4 + t;
            v[t] = (live && d < D) ? __ldg(x + row * D + d) : 0.f;
        }
        float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
        for (int o = LPT / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
        if (!live) continue;
        const float inv = 1.f / fmaxf(sqrtf(ss), kNormEps);
        const float4 r = make_float4(v[0] * inv, v[1] * inv, v[2] * inv, v[3] * inv);
        *reinterpret_cast<float4*>(u + row * DP + sub * 4) = r;
        if (sub == 0) invn[row] = inv;
        colsum.x += r.x; colsum.y += r.y; colsum.z += r.z; colsum.w += r.w;
    }
    red[threadIdx.x][0] = colsum.x; red[threadIdx.x][1] = colsum.y;
    red[threadIdx.x][2] = colsum.z; red[threadIdx.x][3] = colsum.w;
    __syncthreads();
    if (threadIdx.x < DP) {
        const int s = threadIdx.x / 4, t = threadIdx.x % 4;
        double acc = 0.0;
        for (int g = 0; g < RPB; ++g) acc += (double)red[g * LPT + s][t];
        atomicAdd(&hdr->S[threadIdx.x], acc);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float sc = scale ? __ldg(scale) : 1.f;
        hdr->s_used = fminf(fmaxf(sc, kScaleMin), kScaleMax);
        hdr->DP = DP;
    }
}

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ float4 axpby(float s, const float4 a, float t, const float4 b) {
    return make_float4(fmaf(s, a.x, t * b.x), fmaf(s, a.y, t * b.y), fmaf(s, a.z, t * b.z), fmaf(s, a.w, t * b.w));
}

// ---- main pass ---------------------------------------------------------------------------------------
// MODE 0: loss only, 1: loss + gradient accumulation, 2: filter flags only.
template <int LPT, int MODE, typename IdxT>
__global__ void __launch_bounds__(256, 4)            // gathers from L2 are latency-bound: 4 blocks/SM measured best (3: 142 us, 4: 128 us, 5: 134 us)
hyp_triplet_kernel(const float* __restrict__ u, float* __restrict__ G, HypHeader* __restrict__ hdr,
                   const IdxT* __restrict__ a, const IdxT* __restrict__ p, const IdxT* __restrict__ ng,
                   int64_t T0, int64_t n, float inv_temp, int filter_mode, float margin,
                   uint8_t* __restrict__ keep_out) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT, grp = lane / LPT;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float s = hdr->s_used;
    const float4* u4 = reinterpret_cast<const float4*>(u);
    float4* G4 = reinterpret_cast<float4*>(G);

    float loss_acc = 0.f, gs_acc = 0.f;
    unsigned kept_acc = 0;

    // the triplet indices of the NEXT batch are requested before the current one is processed: their (HBM) latency
    // used to sit at the top of every iteration
    auto load_idx = [&](int64_t t, unsigned& xa, unsigned& xp, unsigned& xn) {
        xa = xp = xn = 0;
        if (t < T0) {       // indices are clamped so a bad triplet cannot fault; the reference would raise instead
            xa = (unsigned)min((unsigned long long)(long long)__ldg(a + t), (unsigned long long)(n - 1));
            xp = (unsigned)min((unsigned long long)(long long)__ldg(p + t), (unsigned long long)(n - 1));
            xn = (unsigned)min((unsigned long long)(long long)__ldg(ng + t), (unsigned long long)(n - 1));
        }
    };
    unsigned na_ = 0, np_ = 0, nn_ = 0;
    load_idx(warp_global * 32 + lane, na_, np_, nn_);
    for (int64_t base = warp_global * 32; base < T0; base += warps_total * 32) {
        const int64_t t = base + lane;
        const bool valid = t < T0;
        const unsigned ia = na_, ip = np_, in_ = nn_;
        load_idx(t + warps_total * 32, na_, np_, nn_);
        float c_ap = 0.f, c_an = 0.f, c_pn = 0.f;
        // the sampler emits all triplets of an anchor back to back (t_per_anchor of them), so consecutive slots
        // usually share the anchor row: it is loaded once per run, and (below) its gradient is summed in registers
        // and sent as ONE vector reduction per run instead of one per triplet
        unsigned held_a = 0xffffffffu;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < LPT; ++r) {
            const int slot = grp * LPT + r;       // triplet grp*LPT+r: its owner lane sits in this group
            const unsigned ra = __shfl_sync(kFull, ia, slot);
            const unsigned rp = __shfl_sync(kFull, ip, slot);
            const unsigned rn = __shfl_sync(kFull, in_, slot);
            if (ra != held_a) { va = __ldg(u4 + (size_t)ra * LPT + sub); held_a = ra; }
            const float4 vp = __ldg(u4 + (size_t)rp * LPT + sub);
            const float4 vn = __ldg(u4 + (size_t)rn * LPT + sub);
            float dap = dot4(va, vp), dan = dot4(va, vn), dpn = dot4(vp, vn);
#pragma unroll
            for (int o = LPT / 2; o > 0; o >>= 1) {
                dap += __shfl_xor_sync(kFull, dap, o);
                dan += __shfl_xor_sync(kFull, dan, o);
                dpn += __shfl_xor_sync(kFull, dpn, o);
            }
            if (lane == slot) { c_ap = dap; c_an = dan; c_pn = dpn; }
        }
        if (MODE == 2) {
            const bool k2 = triplet_keep<float>(0.5f * (1.f + c_ap) - 0.5f * (1.f + c_an), filter_mode, margin);
            if (valid) keep_out[t] = k2 ? 1 : 0;
            continue;
        }
        const TripletTerms<float> tt = triplet_terms<float>(c_ap, c_an, c_pn, s, inv_temp, filter_mode, margin);
        const bool keep = valid && tt.keep;
        if (keep) { loss_acc += tt.total; gs_acc += tt.g_s; kept_acc += 1; }
        if (MODE == 1) {
            const unsigned keep_mask = __ballot_sync(kFull, keep);
            unsigned run_a = 0xffffffffu;             // anchor whose gradient is being summed in acc_a
            float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f);
            held_a = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < LPT; ++r) {
                const int slot = grp * LPT + r;       // triplet grp*LPT+r: its owner lane sits in this group
                const unsigned ra = __shfl_sync(kFull, ia, slot);
                const unsigned rp = __shfl_sync(kFull, ip, slot);
                const unsigned rn = __shfl_sync(kFull, in_, slot);
                const float g_ap = __shfl_sync(kFull, tt.g_ap, slot);
                const float g_an = __shfl_sync(kFull, tt.g_an, slot);
                const float g_pn = __shfl_sync(kFull, tt.g_pn, slot);
                if ((keep_mask >> slot) & 1u) {
                    if (ra != held_a) { va = __ldg(u4 + (size_t)ra * LPT + sub); held_a = ra; }
                    const float4 vp = __ldg(u4 + (size_t)rp * LPT + sub);
                    const float4 vn = __ldg(u4 + (size_t)rn * LPT + sub);
                    const float4 ga = axpby(g_ap, vp, g_an, vn);
                    if (ra != run_a) {
                        if (run_a != 0xffffffffu) atomicAdd(G4 + (size_t)run_a * LPT + sub, acc_a);
                        run_a = ra;
                        acc_a = ga;
                    } else {
                        acc_a.x += ga.x; acc_a.y += ga.y; acc_a.z += ga.z; acc_a.w += ga.w;
                    }
                    atomicAdd(G4 + (size_t)rp * LPT + sub, axpby(g_ap, va, g_pn, vn));
                    atomicAdd(G4 + (size_t)rn * LPT + sub, axpby(g_an, va, g_pn, vp));
                }
            }
            if (run_a != 0xffffffffu) atomicAdd(G4 + (size_t)run_a * LPT + sub, acc_a);
        }
    }
    if (MODE != 2) {
        const float ls = warp_sum(loss_acc), gs = warp_sum(gs_acc);
        const unsigned kc = __reduce_add_sync(kFull, kept_acc);
        if (lane == 0 && kc > 0) {
            atomicAdd(&hdr->loss_sum, (double)ls);
            atomicAdd(&hdr->gs_sum, (double)gs);
            atomicAdd(&hdr->kept, (unsigned long long)kc);
        }
    }
}

__global__ void hyp_finalize_kernel(const HypHeader* __restrict__ hdr, int64_t n, float* __restrict__ loss,
                                    int64_t* __restrict__ kept) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s2 = 0.0;
    for (int d = 0; d < hdr->DP; ++d) s2 += hdr->S[d] * hdr->S[d];
    const double mean_sim = 0.5 * (1.0 + s2 / ((double)n * (double)n));
    const double k = (double)hdr->kept;
    // no surviving triplet: 0/0 = NaN, like torch.mean of an empty tensor (ultrametric_loss.py:91)
    loss[0] = (float)(hdr->loss_sum / k + mean_sim);
    kept[0] = (int64_t)hdr->kept;
}

// ---- backward finish: chain through the row normalisation -------------------------------------------
//   gu_i = G_i / kept + S / n^2 ;  gx_i = gloss * (gu_i - (gu_i . u_i) u_i) / |x_i|
template <int LPT>
__global__ void __launch_bounds__(256)
hyp_bwd_kernel(const float* __restrict__ gloss, const float* __restrict__ scale, const HypHeader* __restrict__ hdr,
               const float* __restrict__ u, const float* __restrict__ invn, const float* __restrict__ G,
               int64_t n, int D, float* __restrict__ gx, float* __restrict__ gscale) {
    constexpr int DP = LPT * 4;
    constexpr int RPB = 256 / LPT;
    const int sub = threadIdx.x % LPT, grp = threadIdx.x / LPT;
    const float gl = __ldg(gloss);
    const float inv_k = (float)(1.0 / (double)hdr->kept);
    const float inv_n2 = (float)(1.0 / ((double)n * (double)n));
    float4 sv = make_float4((float)hdr->S[sub * 4] * inv_n2, (float)hdr->S[sub * 4 + 1] * inv_n2,
                            (float)hdr->S[sub * 4 + 2] * inv_n2, (float)hdr->S[sub * 4 + 3] * inv_n2);
    for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < n; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + grp;
        const bool live = row < n;
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 uu = live ? *reinterpret_cast<const float4*>(u + row * DP + sub * 4) : zero4;
        const float4 gg = live ? *reinterpret_cast<const float4*>(G + row * DP + sub * 4) : zero4;
        float4 gu = make_float4(fmaf(gg.x, inv_k, sv.x), fmaf(gg.y, inv_k, sv.y), fmaf(gg.z, inv_k, sv.z), fmaf(gg.w, inv_k, sv.w));
        float dt = dot4(gu, uu);
#pragma unroll
        for (int o = LPT / 2; o > 0; o >>= 1) dt += __shfl_xor_sync(kFull, dt, o);
        if (!live) continue;
        const float inv = invn[row];
        // |x| below eps: F.normalize divides by the constant eps, so the Jacobian is I / eps
        if (inv >= 1.f / kNormEps) dt = 0.f;
        const float f = gl * inv;
        const float o[4] = {f * (gu.x - dt * uu.x), f * (gu.y - dt * uu.y), f * (gu.z - dt * uu.z), f * (gu.w - dt * uu.w)};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int d = sub * 4 + t;
            if (d < D) gx[row * D + d] = o[t];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && gscale) {
        const float sc = __ldg(scale);
        const bool inside = sc >= kScaleMin && sc <= kScaleMax;      // torch.clamp passes grad on [min, max]
        gscale[0] = inside ? (float)((double)gl * hdr->gs_sum / (double)hdr->kept) : 0.f;
    }
}

template <int LPT>
static int run_prep(const float* x, int64_t n, int D, const float* scale, const HypLayout& L, cudaStream_t st) {
    cudaMemsetAsync(L.hdr, 0, sizeof(HypHeader), st);
    constexpr int RPB = 256 / LPT;
    int blocks = (int)((n + RPB - 1) / RPB);
    const int cap = 4 * sm_count();
    if (blocks > cap) blocks = cap;
    hyp_prep_kernel<LPT><<<blocks, 256, 0, st>>>(x, n, D, scale, L.u, L.invn, L.hdr);
    return check_launch("hyp_prep_kernel");
}

template <int LPT, typename IdxT>
static int run_triplets(int mode, const HypLayout& L, const IdxT* a, const IdxT* p, const IdxT* ng,
                        int64_t T0, int64_t n, float temperature, int filter_mode, float margin, uint8_t* keep,
                        cudaStream_t st) {
    if (T0 <= 0) return HPCS_OK;
    int64_t warps = (T0 + 31) / 32;
    int64_t blocks = (warps + 7) / 8;
    const int64_t cap = (int64_t)8 * sm_count();
    if (blocks > cap) blocks = cap;
    const float inv_temp = 1.f / temperature;
    if (mode == 0) hyp_triplet_kernel<LPT, 0, IdxT><<<(int)blocks, 256, 0, st>>>(L.u, L.G, L.hdr, a, p, ng, T0, n, inv_temp, filter_mode, margin, keep);
    else if (mode == 1) hyp_triplet_kernel<LPT, 1, IdxT><<<(int)blocks, 256, 0, st>>>(L.u, L.G, L.hdr, a, p, ng, T0, n, inv_temp, filter_mode, margin, keep);
    else hyp_triplet_kernel<LPT, 2, IdxT><<<(int)blocks, 256, 0, st>>>(L.u, L.G, L.hdr, a, p, ng, T0, n, inv_temp, filter_mode, margin, keep);
    return check_launch("hyp_triplet_kernel");
}

#define HPCS_LPT_SWITCH(DP, ...)                          \
    switch (DP) {                                         \
        case 4: { constexpr int LPT = 1; __VA_ARGS__; } break;   \
        case 8: { constexpr int LPT = 2; __VA_ARGS__; } break;   \
        case 16: { constexpr int LPT = 4; __VA_ARGS__; } break;  \
        case 32: { constexpr int LPT = 8; __VA_ARGS__; } break;  \
        case 64: { constexpr int LPT = 16; __VA_ARGS__; } break; \
        case 128: { constexpr int LPT = 32; __VA_ARGS__; } break;\
        default: rc = fail(HPCS_ERR_ARG, "embedding dim %d not supported (max 128)", D); }

}  // namespace hpcs

extern "C" {

size_t hpcs_hyp_triplet_workspace_bytes(int64_t n, int D) {
    if (n <= 0 || D <= 0 || D > 128) return 0;
    return hpcs::hyp_layout(nullptr, n, D).bytes;
}

}  // extern "C"

namespace hpcs {
template <typename IdxT>
static int triplet_fwd(const float* x, int64_t n, int D, const IdxT* a, const IdxT* p, const IdxT* ng, int64_t T0,
                       const float* scale, float temperature, int filter_mode, float margin, int need_grad, float* loss,
                       int64_t* kept, void* ws, size_t ws_bytes, void* stream) {
    if (!x || !scale || !loss || !kept || !ws || (T0 > 0 && (!a || !p || !ng))) return fail(HPCS_ERR_ARG, "hyp_triplet_fwd: null pointer");
    if (n <= 0 || D <= 0 || D > 128 || T0 < 0 || n > 0x7fffffffLL) return fail(HPCS_ERR_ARG, "hyp_triplet_fwd: bad shape n=%lld D=%d", (long long)n, D);
    if (!(temperature > 0.f)) return fail(HPCS_ERR_ARG, "hyp_triplet_fwd: temperature must be > 0");
    const HypLayout L = hyp_layout(ws, n, D);
    if (ws_bytes < L.bytes) return fail(HPCS_ERR_WORKSPACE, "hyp_triplet_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int DP = padded_dim(D);
    int rc = HPCS_OK;
    HPCS_LPT_SWITCH(DP, rc = run_prep<LPT>(x, n, D, scale, L, st));
    if (rc) return rc;
    if (need_grad) cudaMemsetAsync(L.G, 0, (size_t)n * DP * sizeof(float), st);
    HPCS_LPT_SWITCH(DP, rc = (run_triplets<LPT, IdxT>(need_grad ? 1 : 0, L, a, p, ng, T0, n, temperature, filter_mode, margin, nullptr, st)));
    if (rc) return rc;
    hyp_finalize_kernel<<<1, 32, 0, st>>>(L.hdr, n, loss, kept);
    return check_launch("hyp_finalize_kernel");
}

template <typename IdxT>
static int triplet_filter(const float* x, int64_t n, int D, const IdxT* a, const IdxT* p, const IdxT* ng, int64_t T0,
                          int filter_mode, float margin, uint8_t* keep, void* ws, size_t ws_bytes, void* stream) {
    if (!x || !ws || (T0 > 0 && (!a || !p || !ng || !keep))) return fail(HPCS_ERR_ARG, "triplet_filter: null pointer");
    if (n <= 0 || D <= 0 || D > 128 || T0 < 0 || n > 0x7fffffffLL) return fail(HPCS_ERR_ARG, "triplet_filter: bad shape");
    const HypLayout L = hyp_layout(ws, n, D);
    if (ws_bytes < L.bytes) return fail(HPCS_ERR_WORKSPACE, "triplet_filter: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int DP = padded_dim(D);
    int rc = HPCS_OK;
    HPCS_LPT_SWITCH(DP, rc = run_prep<LPT>(x, n, D, nullptr /* the filter does not depend on the scale */, L, st));
    if (rc) return rc;
    HPCS_LPT_SWITCH(DP, rc = (run_triplets<LPT, IdxT>(2, L, a, p, ng, T0, n, 1.0f, filter_mode, margin, keep, st)));
    return rc;
}
}  // namespace hpcs

extern "C" {

int hpcs_hyp_triplet_fwd_f32(const float* x, int64_t n, int D, const int64_t* a, const int64_t* p,
                             const int64_t* ng, int64_t T0, const float* scale, float temperature,
                             int filter_mode, float margin, int need_grad, float* loss, int64_t* kept,
                             void* ws, size_t ws_bytes, void* stream) {
    return hpcs::triplet_fwd<int64_t>(x, n, D, a, p, ng, T0, scale, temperature, filter_mode, margin, need_grad, loss, kept, ws,
                                      ws_bytes, stream);
}

int hpcs_hyp_triplet_fwd_i32_f32(const float* x, int64_t n, int D, const int32_t* a, const int32_t* p,
                                 const int32_t* ng, int64_t T0, const float* scale, float temperature,
                                 int filter_mode, float margin, int need_grad, float* loss, int64_t* kept,
                                 void* ws, size_t ws_bytes, void* stream) {
    return hpcs::triplet_fwd<int32_t>(x, n, D, a, p, ng, T0, scale, temperature, filter_mode, margin, need_grad, loss, kept, ws,
                                      ws_bytes, stream);
}

int hpcs_hyp_triplet_bwd_f32(const float* gloss, const float* x, int64_t n, int D, const float* scale,
                             const void* ws, size_t ws_bytes, float* gx, float* gscale, void* stream) {
    using namespace hpcs;
    (void)x;
    if (!gloss || !scale || !ws || !gx) return fail(HPCS_ERR_ARG, "hyp_triplet_bwd: null pointer");
    if (n <= 0 || D <= 0 || D > 128) return fail(HPCS_ERR_ARG, "hyp_triplet_bwd: bad shape");
    const HypLayout L = hyp_layout(const_cast<void*>(ws), n, D);
    if (ws_bytes < L.bytes) return fail(HPCS_ERR_WORKSPACE, "hyp_triplet_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    const int DP = padded_dim(D);
    int rc = HPCS_OK;
    HPCS_LPT_SWITCH(DP, {
        constexpr int RPB = 256 / LPT;
        int blocks = (int)((n + RPB - 1) / RPB);
        const int cap = 8 * sm_count();
        if (blocks > cap) blocks = cap;
        hyp_bwd_kernel<LPT><<<blocks, 256, 0, st>>>(gloss, scale, L.hdr, L.u, L.invn, L.G, n, D, gx, gscale);
        rc = check_launch("hyp_bwd_kernel");
    });
    return rc;
}

int hpcs_triplet_filter_f32(const float* x, int64_t n, int D, const int64_t* a, const int64_t* p,
                            const int64_t* ng, int64_t T0, int filter_mode, float margin, uint8_t* keep,
                            void* ws, size_t ws_bytes, void* stream) {
    return hpcs::triplet_filter<int64_t>(x, n, D, a, p, ng, T0, filter_mode, margin, keep, ws, ws_bytes, stream);
}

int hpcs_triplet_filter_i32_f32(const float* x, int64_t n, int D, const int32_t* a, const int32_t* p,
                                const int32_t* ng, int64_t T0, int filter_mode, float margin, uint8_t* keep,
                                void* ws, size_t ws_bytes, void* stream) {
    return hpcs::triplet_filter<int32_t>(x, n, D, a, p, ng, T0, filter_mode, margin, keep, ws, ws_bytes, stream);
}

}  // extern "C"
