// Scalar Poincare-ball math shared by the CUDA kernels and by the host-side unit check
// (tests/csrc/hyp_math_check.cpp compiles this header with g++).  Every function is templated on
// the scalar type so the same code can be run in double on the CPU to validate the formulas.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HPCS_HD __host__ __device__ __forceinline__
#else
#define HPCS_HD inline
#endif

namespace hpcs {

constexpr float kArtanhClamp = 1e-5f;      // hpcs/utils/math.py:64
constexpr float kScaleMin = 1e-4f;         // hpcs/loss/ultrametric_loss.py:141
constexpr float kScaleMax = 1.0f;          // hpcs/loss/ultrametric_loss.py:142
constexpr float kNormEps = 1e-12f;         // F.normalize eps (hpcs/loss/ultrametric_loss.py:143)

template <typename T> HPCS_HD T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> HPCS_HD T t_min(T a, T b) { return a < b ? a : b; }
HPCS_HD float t_sqrt(float v) { return sqrtf(v); }
HPCS_HD double t_sqrt(double v) { return sqrt(v); }
HPCS_HD float t_log1p(float v) { return log1pf(v); }
HPCS_HD double t_log1p(double v) { return log1p(v); }
HPCS_HD float t_exp(float v) { return expf(v); }
HPCS_HD double t_exp(double v) { return exp(v); }

// hyp_lca(a, b, return_coord=False) for two points of equal Euclidean norm s and cos(angle) = c
// (hpcs/distances/lca.py:37-52 specialised).  The geodesic through the points is the
// circle orthogonal to the unit sphere whose centre sits at distance cc = (1+s^2) / (2 s cos(t/2))
// from the origin; its point nearest the origin has norm rho = 1 / (cc + sqrt(cc^2 - 1)), and the
// result is 2 artanh(min(rho, 1-1e-5)) (hpcs/utils/math.py:61-74; the reference's backward does not
// mask the clamp, so neither do the derivatives below).
//   cc - 1 is formed without cancellation as ((1-s)^2 + 2 s (1-h)) / (2 s h), h = cos(t/2).
template <typename T>
HPCS_HD void lca_equal_radius(T c, T s, T& d, T& dd_dc, T& dd_ds) {
    c = t_min(t_max(c, T(-1)), T(1));
    const T h2 = t_max(T(0.5) * (T(1) + c), T(1e-12));
    const T h = t_sqrt(h2);
    const T one_m_h = (T(0.5) * (T(1) - c)) / (T(1) + h);
    const T num = (T(1) - s) * (T(1) - s) + T(2) * s * one_m_h;
    const T cm1 = num / (T(2) * s * h);
    const T cc = T(1) + cm1;
    const T q = t_sqrt(t_max(cm1 * (cm1 + T(2)), T(1e-30)));
    const T rho = T(1) / (cc + q);
    const T rc = t_min(rho, T(1) - T(kArtanhClamp));
    d = t_log1p(T(2) * rc / (T(1) - rc));
    const T dd_drho = T(2) / (T(1) - rc * rc);
    const T dd_dcc = dd_drho * (-rho / q);
    dd_dc = dd_dcc * (-cc / (T(4) * h2));
    dd_ds = dd_dcc * ((s * s - T(1)) / (T(2) * s * s * h));
}

enum FilterMode { kKeepAll = 0, kEasy = 1, kSemihard = 2, kHard = 3, kBelowMargin = 4 };

// hpcs/miner/triplet_margin_miner.py:20-32 with an inverted distance: gap = sim(a,p) - sim(a,n).
template <typename T>
HPCS_HD bool triplet_keep(T gap, int mode, T margin) {
    switch (mode) {
        case kEasy: return gap > margin;
        case kSemihard: return gap <= margin && gap > T(0);
        case kHard: return gap <= margin && gap <= T(0);
        case kBelowMargin: return gap <= margin;
        default: return true;
    }
}

template <typename T>
struct TripletTerms {
    T total;            // (w_ap + w_an + w_pn) - sum_m w_m softmax_m(d / temperature)
    T g_ap, g_an, g_pn; // d total / d cos for the three pairs
    T g_s;              // d total / d s
    bool keep;
};

// One mined triplet of MetricHyperbolicLoss.compute_hyp (hpcs/loss/ultrametric_loss.py:67-89) given
// the three cosines of the unit rows: similarities w = (1+c)/2 (hpcs/distances/cosine.py:10-13),
// LCA distances, softmax weights, and the analytic derivatives.
template <typename T>
HPCS_HD TripletTerms<T> triplet_terms(T c_ap, T c_an, T c_pn, T s, T inv_temp, int mode, T margin) {
    TripletTerms<T> o;
    const T w0 = T(0.5) * (T(1) + c_ap), w1 = T(0.5) * (T(1) + c_an), w2 = T(0.5) * (T(1) + c_pn);
    o.keep = triplet_keep(w0 - w1, mode, margin);
    T d0, d1, d2, dc0, dc1, dc2, ds0, ds1, ds2;
    lca_equal_radius(c_ap, s, d0, dc0, ds0);
    lca_equal_radius(c_an, s, d1, dc1, ds1);
    lca_equal_radius(c_pn, s, d2, dc2, ds2);
    const T dm = t_max(d0, t_max(d1, d2));
    const T e0 = t_exp((d0 - dm) * inv_temp), e1 = t_exp((d1 - dm) * inv_temp), e2 = t_exp((d2 - dm) * inv_temp);
    const T inv = T(1) / (e0 + e1 + e2);
    const T p0 = e0 * inv, p1 = e1 * inv, p2 = e2 * inv;
    const T wbar = w0 * p0 + w1 * p1 + w2 * p2;
    o.total = (w0 + w1 + w2) - wbar;
    const T t0 = -p0 * inv_temp * (w0 - wbar), t1 = -p1 * inv_temp * (w1 - wbar), t2 = -p2 * inv_temp * (w2 - wbar);
    o.g_ap = T(0.5) * (T(1) - p0) + t0 * dc0;
    o.g_an = T(0.5) * (T(1) - p1) + t1 * dc1;
    o.g_pn = T(0.5) * (T(1) - p2) + t2 * dc2;
    o.g_s = t0 * ds0 + t1 * ds1 + t2 * ds2;
    return o;
}

// ---- forward-mode duals for the general (unequal norm) hyp_lca ------------------------------------
// value + partials wrt (A = |a|^2, Bn = |b|^2, ab = a.b)
struct Dual3 {
    double v, d[3];
};
HPCS_HD Dual3 mk(double v) { return Dual3{v, {0.0, 0.0, 0.0}}; }
HPCS_HD Dual3 var(double v, int i) { Dual3 r = mk(v); r.d[i] = 1.0; return r; }
HPCS_HD Dual3 operator+(Dual3 x, Dual3 y) { return Dual3{x.v + y.v, {x.d[0] + y.d[0], x.d[1] + y.d[1], x.d[2] + y.d[2]}}; }
HPCS_HD Dual3 operator-(Dual3 x, Dual3 y) { return Dual3{x.v - y.v, {x.d[0] - y.d[0], x.d[1] - y.d[1], x.d[2] - y.d[2]}}; }
HPCS_HD Dual3 operator*(Dual3 x, Dual3 y) {
    return Dual3{x.v * y.v, {x.d[0] * y.v + x.v * y.d[0], x.d[1] * y.v + x.v * y.d[1], x.d[2] * y.v + x.v * y.d[2]}};
}
HPCS_HD Dual3 operator/(Dual3 x, Dual3 y) {
    const double q = x.v / y.v, iy = 1.0 / y.v;
    return Dual3{q, {(x.d[0] - q * y.d[0]) * iy, (x.d[1] - q * y.d[1]) * iy, (x.d[2] - q * y.d[2]) * iy}};
}
HPCS_HD Dual3 dsqrt(Dual3 x) {
    const double r = sqrt(x.v), k = 0.5 / r;
    return Dual3{r, {x.d[0] * k, x.d[1] * k, x.d[2] * k}};
}

struct LcaGeneral {
    Dual3 ca, cb;     // proj = ca * a + cb * b
    Dual3 rho;        // |proj| (before the artanh clamp)
};

// hyp_lca for arbitrary a, b in the ball (hpcs/distances/lca.py:37-52).  The reflection chain of the
// reference stays inside span(a, b); carrying it out on coefficient pairs (alpha, beta) of
// alpha*a + beta*b needs only A = |a|^2, Bn = |b|^2, ab = a.b:
//   r      = a / A                                   (lca.py:15-17)
//   inv(x) = (|r|^2 - 1) / |x - r|^2 (x - r) + r     (lca.py:8-12)
//   refl   = 2 (a.y) / max(|y|^2, 1e-15) y - a       (lca.py:20-29, reflecting a across y)
//   proj   = o / (1 + sqrt(1 - |o|^2))               (lca.py:32-34)
HPCS_HD LcaGeneral lca_general(double A_, double B_, double ab_) {
    const Dual3 A = var(A_, 0), Bn = var(B_, 1), ab = var(ab_, 2);
    const Dual3 one = mk(1.0), two = mk(2.0);
    auto dot = [&](Dual3 p0, Dual3 p1, Dual3 q0, Dual3 q1) { return p0 * q0 * A + (p0 * q1 + p1 * q0) * ab + p1 * q1 * Bn; };
    const Dual3 ra = one / A;                       // r = ra * a
    const Dual3 r2m1 = one / A - one;               // |r|^2 - 1
    // b_inv = inv(b): u = b - r = (-ra, 1)
    Dual3 u0 = mk(0.0) - ra, u1 = one;
    Dual3 f = r2m1 / dot(u0, u1, u0, u1);
    const Dual3 y0 = f * u0 + ra, y1 = f * u1;      // b_inv
    // reflect a across the line through b_inv
    const Dual3 ay = dot(one, mk(0.0), y0, y1);
    Dual3 yy = dot(y0, y1, y0, y1);
    if (yy.v < 1e-15) yy = mk(1e-15);
    const Dual3 g = two * ay / yy;
    const Dual3 z0 = g * y0 - one, z1 = g * y1;     // reflected a
    // o = inv(z)
    u0 = z0 - ra; u1 = z1;
    f = r2m1 / dot(u0, u1, u0, u1);
    const Dual3 o0 = f * u0 + ra, o1 = f * u1;
    const Dual3 oo = dot(o0, o1, o0, o1);
    const Dual3 den = one + dsqrt(one - oo);
    LcaGeneral out;
    out.ca = o0 / den;
    out.cb = o1 / den;
    out.rho = dsqrt(oo) / den;
    return out;
}

}  // namespace hpcs
