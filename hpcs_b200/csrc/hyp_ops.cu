// Stand-alone Poincare-ball ops behind the reference's public signatures:
//   hyp_lca(a, b, return_coord)  hpcs/distances/lca.py:37-52          (general, unequal norms)
//   ExpMap.forward               hpcs/nn/hyperbolic/hyp_embed.py:6-10 -> expmap_1(u, 0)
//   normalize_embeddings+project hpcs/loss/ultrametric_loss.py:139-143, hpcs/distances/poincare.py:61-68
// These are thin, HBM-bound row kernels (one warp per row).  The general hyp_lca carries its
// scalar chain in fp64 with forward-mode duals (hyp_math.cuh), because the reference's fp32
// evaluation loses all digits at the default scale 1e-3 (SURVEY.md Finding 4).
#include "common.cuh"
#include "hyp_math.cuh"

namespace hpcs {

constexpr int kRowWarps = 8;

// ---- hyp_lca ----------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kRowWarps * 32)
hyp_lca_kernel(const float* __restrict__ gout, const float* __restrict__ a, const float* __restrict__ b,
               int64_t T, int D, int return_coord, float* __restrict__ out, float* __restrict__ ga,
               float* __restrict__ gb) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= T) return;
    const float* ar = a + row * D;
    const float* br = b + row * D;
    double A = 0.0, Bn = 0.0, ab = 0.0, ga_dot = 0.0, gb_dot = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double x = ar[d], y = br[d];
        A += x * x; Bn += y * y; ab += x * y;
        if (BWD && return_coord) {
            const double g = gout[row * D + d];
            ga_dot += g * x; gb_dot += g * y;
        }
    }
    A = warp_sum(A); Bn = warp_sum(Bn); ab = warp_sum(ab);
    const LcaGeneral r = lca_general(A, Bn, ab);
    const double lim = 1.0 - (double)kArtanhClamp;
    const double rc = r.rho.v < lim ? r.rho.v : lim;
    if (!BWD) {
        if (return_coord) {
            for (int d = lane; d < D; d += 32) out[row * D + d] = (float)(r.ca.v * ar[d] + r.cb.v * br[d]);
        } else if (lane == 0) {
            out[row] = (float)log1p(2.0 * rc / (1.0 - rc));     // 2 artanh(rc)
        }
        return;
    }
    // backward: coefficients of (a, b) in the two gradients plus, for return_coord, a copy of g
    double ka_a, ka_b, kb_a, kb_b, kg_a = 0.0, kg_b = 0.0;      // ga = kg_a*g + ka_a*a + ka_b*b ; gb likewise
    if (return_coord) {
        ga_dot = warp_sum(ga_dot); gb_dot = warp_sum(gb_dot);
        const double wA = ga_dot * r.ca.d[0] + gb_dot * r.cb.d[0];
        const double wB = ga_dot * r.ca.d[1] + gb_dot * r.cb.d[1];
        const double wab = ga_dot * r.ca.d[2] + gb_dot * r.cb.d[2];
        kg_a = r.ca.v; kg_b = r.cb.v;
        ka_a = 2.0 * wA; ka_b = wab;
        kb_b = 2.0 * wB; kb_a = wab;
    } else {
        const double g = (double)gout[row] * 2.0 / (1.0 - rc * rc);   // Artanh.backward on the clamped value
        ka_a = g * 2.0 * r.rho.d[0]; ka_b = g * r.rho.d[2];
        kb_b = g * 2.0 * r.rho.d[1]; kb_a = g * r.rho.d[2];
    }
    for (int d = lane; d < D; d += 32) {
        const double x = ar[d], y = br[d];
        const double g = return_coord ? (double)gout[row * D + d] : 0.0;
        ga[row * D + d] = (float)(kg_a * g + ka_a * x + ka_b * y);
        gb[row * D + d] = (float)(kg_b * g + kb_a * x + kb_b * y);
    }
}

// ---- expmap at the origin -----------------------------------------------------------------------------
//   y = tanh(min(|u|, 15)) u / max(|u|, 1e-15)        (hpcs/utils/math.py:81-82, poincare.py:50-54)
template <bool BWD>
__global__ void __launch_bounds__(kRowWarps * 32)
expmap0_kernel(const float* __restrict__ gy, const float* __restrict__ u, int64_t rows, int D,
               float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* ur = u + row * D;
    float ss = 0.f, gu = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float v = ur[d];
        ss = fmaf(v, v, ss);
        if (BWD) gu = fmaf(gy[row * D + d], v, gu);
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float r = fmaxf(nrm, 1e-15f);
    const float f = tanhf(fminf(r, 15.f));
    const float fr = f / r;
    if (!BWD) {
        for (int d = lane; d < D; d += 32) out[row * D + d] = fr * ur[d];
        return;
    }
    gu = warp_sum(gu);
    // d/du [ f(r) u / r ] = (f/r) I + (f'(r)/r - f/r^2) u u^T / r ; zero second term where r is clamped
    const float fp = r < 15.f ? 1.f - f * f : 0.f;
    const float k2 = nrm > 1e-15f ? (fp - fr) / (r * r) * gu : 0.f;
    for (int d = lane; d < D; d += 32) out[row * D + d] = fmaf(k2, ur[d], fr * gy[row * D + d]);
}

// ---- leaves for the decoder -----------------------------------------------------------------------------
//   e = x / max(|x|, 1e-12) * clamp(scale, 1e-4, 1);  project: if |e| > 1 - 4e-3, e <- e / |e| * (1 - 4e-3)
__global__ void __launch_bounds__(kRowWarps * 32)
leaves_kernel(const float* __restrict__ x, int64_t rows, int D, const float* __restrict__ scale,
              float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float s = fminf(fmaxf(__ldg(scale), kScaleMin), kScaleMax);
    const float* xr = x + row * D;
    float ss = 0.f;
    for (int d = lane; d < D; d += 32) ss = fmaf(xr[d], xr[d], ss);
    ss = warp_sum(ss);
    const float den = fmaxf(sqrtf(ss), kNormEps);
    float es = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float e = __fmul_rn(__fdiv_rn(xr[d], den), s);
        es = fmaf(e, e, es);
    }
    es = warp_sum(es);
    const float en = fmaxf(sqrtf(es), 1e-15f);
    const float lim = 1.f - 4e-3f;
    for (int d = lane; d < D; d += 32) {
        float e = __fmul_rn(__fdiv_rn(xr[d], den), s);
        if (en > lim) e = __fmul_rn(__fdiv_rn(e, en), lim);
        out[row * D + d] = e;
    }
}

static inline int row_blocks(int64_t rows) { return (int)((rows + kRowWarps - 1) / kRowWarps); }

}  // namespace hpcs

extern "C" {

int hpcs_hyp_lca_fwd_f32(const float* a, const float* b, int64_t T, int D, int return_coord, float* out,
                         void* stream) {
    using namespace hpcs;
    if (!a || !b || !out) return fail(HPCS_ERR_ARG, "hyp_lca_fwd: null pointer");
    if (T < 0 || D <= 0) return fail(HPCS_ERR_ARG, "hyp_lca_fwd: bad shape");
    if (T == 0) return HPCS_OK;
    hyp_lca_kernel<false><<<row_blocks(T), kRowWarps * 32, 0, as_stream(stream)>>>(nullptr, a, b, T, D, return_coord, out, nullptr, nullptr);
    return check_launch("hyp_lca_kernel<fwd>");
}

int hpcs_hyp_lca_bwd_f32(const float* gout, const float* a, const float* b, int64_t T, int D,
                         int return_coord, float* ga, float* gb, void* stream) {
    using namespace hpcs;
    if (!gout || !a || !b || !ga || !gb) return fail(HPCS_ERR_ARG, "hyp_lca_bwd: null pointer");
    if (T < 0 || D <= 0) return fail(HPCS_ERR_ARG, "hyp_lca_bwd: bad shape");
    if (T == 0) return HPCS_OK;
    hyp_lca_kernel<true><<<row_blocks(T), kRowWarps * 32, 0, as_stream(stream)>>>(gout, a, b, T, D, return_coord, nullptr, ga, gb);
    return check_launch("hyp_lca_kernel<bwd>");
}

int hpcs_expmap0_fwd_f32(const float* u, int64_t rows, int D, float* y, void* stream) {
    using namespace hpcs;
    if (!u || !y) return fail(HPCS_ERR_ARG, "expmap0_fwd: null pointer");
    if (rows < 0 || D <= 0) return fail(HPCS_ERR_ARG, "expmap0_fwd: bad shape");
    if (rows == 0) return HPCS_OK;
    expmap0_kernel<false><<<row_blocks(rows), kRowWarps * 32, 0, as_stream(stream)>>>(nullptr, u, rows, D, y);
    return check_launch("expmap0_kernel<fwd>");
}

int hpcs_expmap0_bwd_f32(const float* gy, const float* u, int64_t rows, int D, float* gu, void* stream) {
    using namespace hpcs;
    if (!gy || !u || !gu) return fail(HPCS_ERR_ARG, "expmap0_bwd: null pointer");
    if (rows < 0 || D <= 0) return fail(HPCS_ERR_ARG, "expmap0_bwd: bad shape");
    if (rows == 0) return HPCS_OK;
    expmap0_kernel<true><<<row_blocks(rows), kRowWarps * 32, 0, as_stream(stream)>>>(gy, u, rows, D, gu);
    return check_launch("expmap0_kernel<bwd>");
}

int hpcs_leaves_f32(const float* x, int64_t rows, int D, const float* scale, float* leaves, void* stream) {
    using namespace hpcs;
    if (!x || !scale || !leaves) return fail(HPCS_ERR_ARG, "leaves: null pointer");
    if (rows < 0 || D <= 0) return fail(HPCS_ERR_ARG, "leaves: bad shape");
    if (rows == 0) return HPCS_OK;
    leaves_kernel<<<row_blocks(rows), kRowWarps * 32, 0, as_stream(stream)>>>(x, rows, D, scale, leaves);
    return check_launch("leaves_kernel");
}

}  // extern "C"
