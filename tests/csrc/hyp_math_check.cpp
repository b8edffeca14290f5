// Host build of hpcs_b200/csrc/hyp_math.cuh: lets the CPU test-suite check the kernel's scalar
// formulas (closed-form equal-radius LCA, per-triplet terms and derivatives, general LCA with
// forward-mode duals) against the oracle without a GPU.
#include "../../hpcs_b200/csrc/hyp_math.cuh"

using namespace hpcs;

extern "C" {

// out: d, dd_dc, dd_ds
void check_lca_equal_f32(float c, float s, float* out) { lca_equal_radius<float>(c, s, out[0], out[1], out[2]); }
void check_lca_equal_f64(double c, double s, double* out) { lca_equal_radius<double>(c, s, out[0], out[1], out[2]); }

// out: total, g_ap, g_an, g_pn, g_s, keep
void check_triplet_f32(float c_ap, float c_an, float c_pn, float s, float inv_temp, int mode, float margin, float* out) {
    TripletTerms<float> t = triplet_terms<float>(c_ap, c_an, c_pn, s, inv_temp, mode, margin);
    out[0] = t.total; out[1] = t.g_ap; out[2] = t.g_an; out[3] = t.g_pn; out[4] = t.g_s; out[5] = t.keep ? 1.f : 0.f;
}
void check_triplet_f64(double c_ap, double c_an, double c_pn, double s, double inv_temp, int mode, double margin, double* out) {
    TripletTerms<double> t = triplet_terms<double>(c_ap, c_an, c_pn, s, inv_temp, mode, margin);
    out[0] = t.total; out[1] = t.g_ap; out[2] = t.g_an; out[3] = t.g_pn; out[4] = t.g_s; out[5] = t.keep ? 1.0 : 0.0;
}

// out: ca, cb, rho, then d ca/d(A,B,ab), d cb/d(A,B,ab), d rho/d(A,B,ab)
void check_lca_general(double A, double B, double ab, double* out) {
    LcaGeneral r = lca_general(A, B, ab);
    out[0] = r.ca.v; out[1] = r.cb.v; out[2] = r.rho.v;
    for (int i = 0; i < 3; ++i) { out[3 + i] = r.ca.d[i]; out[6 + i] = r.cb.d[i]; out[9 + i] = r.rho.d[i]; }
}

}
