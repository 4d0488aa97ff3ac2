// K2 -- coefficient arrays: exp(c z), psi_k(c z) (closed form or contour mean) and the
// tableau combinations of each method.  Restates
//   rkstiff/etd.py:134-182 (psi1..3), etd4.py:87-139, etd34.py:84-149, etd5.py:115-203,
//   etd35.py:157-288, if4.py:72-83, if34.py:81-93, if45dp.py:183-238.
#pragma once
#include "common.cuh"

namespace rks {

// psi_r = r! phi_r  (etd.py:148,165,182).  z**3 follows NumPy's integer power: z*(z*z).
RKS_HD cplx psi1(cplx z) { return (cexp(z) - 1.0) / z; }
RKS_HD cplx psi2(cplx z) { return (2.0 * (cexp(z) - 1.0 - z)) / (z * z); }
RKS_HD cplx psi3(cplx z) {
    const cplx z2 = z * z;
    return (6.0 * (cexp(z) - 1.0 - z - z2 / 2.0)) / (z * z2);
}

// the nine psi values the ETD tableaux use (all already multiplied by h on output)
struct PsiSet {
    cplx p1q, p2q;      // psi1, psi2 at z/4
    cplx p1h, p2h;      // at z/2
    cplx p1t, p2t;      // at 3z/4
    cplx p1, p2, p3;    // at z
};

RKS_HD PsiSet psi_zero() {
    PsiSet p;
    p.p1q = p.p2q = p.p1h = p.p2h = p.p1t = p.p2t = p.p1 = p.p2 = p.p3 = mk(0.0, 0.0);
    return p;
}

// psi1..psi3 at w from e = exp(w) with ONE complex reciprocal (the reference divides three times; the
// results agree to a few ulp, far below the cancellation noise of these formulas near the cutoff)
RKS_HD void psi_from_exp(cplx w, cplx e, cplx& p1, cplx& p2, cplx* p3) {
    const cplx inv = mk(1.0, 0.0) / w;
    const cplx inv2 = inv * inv;
    const cplx em1 = e - 1.0;
    p1 = em1 * inv;
    p2 = (2.0 * (em1 - w)) * inv2;
    if (p3) *p3 = (6.0 * (em1 - w - (w * w) / 2.0)) * (inv2 * inv);
}

// the exponentials exp(w/4), exp(w/2), exp(3w/4), exp(w), each evaluated directly like the reference
// does (etd35.py:180-183).  Building them from one exp by squaring is 2x cheaper but quadruples the
// rounding error of e^w, which the psi numerators e^w - 1 - w - ... amplify by 1/|w|^3 just above the
// mode cutoff: measured, that moved an adaptive dt by 1.4e-9 relative (bar: 1e-9).
struct ExpSet { cplx q, h, t, f; };
template <bool FIVE>
RKS_HD ExpSet exp_set(cplx w) {
    ExpSet e;
    e.h = cexp(w * 0.5);
    e.f = cexp(w);
    if (FIVE) {
        e.q = cexp(w * 0.25);
        e.t = cexp((3.0 * w) / 4.0);
    } else {
        e.q = e.t = e.h;
    }
    return e;
}

// accumulate the psi values at one point w (w = z for the closed form, w = z + r_j on the contour)
template <bool FIVE>
RKS_HD void psi_accumulate(PsiSet& acc, cplx w, const ExpSet& e) {
    cplx a, b, c;
    if (FIVE) {
        psi_from_exp(w * 0.25, e.q, a, b, nullptr);
        acc.p1q = acc.p1q + a; acc.p2q = acc.p2q + b;
        psi_from_exp((3.0 * w) / 4.0, e.t, a, b, nullptr);
        acc.p1t = acc.p1t + a; acc.p2t = acc.p2t + b;
    }
    psi_from_exp(w * 0.5, e.h, a, b, nullptr);
    acc.p1h = acc.p1h + a; acc.p2h = acc.p2h + b;
    psi_from_exp(w, e.f, a, b, &c);
    acc.p1 = acc.p1 + a; acc.p2 = acc.p2 + b; acc.p3 = acc.p3 + c;
}
template <bool FIVE>
RKS_HD void psi_accumulate(PsiSet& acc, cplx w) { psi_accumulate<FIVE>(acc, w, exp_set<FIVE>(w)); }

RKS_HD void psi_scale(PsiSet& p, double h, double m) {
    // h * sum / M  (etd35.py:261); m == 1 for the closed form
#define RKS_SC(f) p.f = (h * p.f) / m
    RKS_SC(p1q); RKS_SC(p2q); RKS_SC(p1h); RKS_SC(p2h); RKS_SC(p1t); RKS_SC(p2t); RKS_SC(p1); RKS_SC(p2); RKS_SC(p3);
#undef RKS_SC
}

// contour node r_j = R exp(2 pi i (j + 1/2) / M), etd35.py:259
RKS_HD cplx contour_node(double radius, int j, int m) {
    const double a = (2.0 * (j + 0.5)) / m;    // angle / pi
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(a, &s, &c);
#else
    s = sin(M_PI * a); c = cos(M_PI * a);
#endif
    return mk(radius * c, radius * s);
}

// Krogstad ETD4 / ETD34 rows (etd4.py:109-116, etd34.py:119-126); out[] indexed by kro::
RKS_HD void tableau_krogstad(const PsiSet& p, cplx* out) {
    out[kro::a21] = 0.5 * p.p1h;
    out[kro::a31] = 0.5 * (p.p1h - p.p2h);
    out[kro::a32] = 0.5 * p.p2h;
    out[kro::a41] = p.p1 - p.p2;
    out[kro::a43] = p.p2;
    out[kro::a51] = p.p1 - (3.0 / 2) * p.p2 + (2.0 / 3) * p.p3;
    out[kro::a52] = p.p2 - (2.0 / 3) * p.p3;
    out[kro::a54] = -((1.0 / 2) * p.p2) + (2.0 / 3) * p.p3;
}

// ETD5 / ETD35 rows (etd5.py:149-166, etd35.py:220-237); out[] indexed by e5::
RKS_HD void tableau_etd5(const PsiSet& p, cplx* out) {
    out[e5::a21] = p.p1q / 4.0;
    out[e5::a31] = (p.p1q - p.p2q / 2.0) / 4.0;
    out[e5::a32] = p.p2q / 8.0;
    out[e5::a41] = (p.p1h - p.p2h) / 2.0;
    out[e5::a43] = p.p2h / 2.0;
    out[e5::a51] = (3.0 * (p.p1t - (3.0 * p.p2t) / 4.0)) / 4.0;
    out[e5::a52] = (-3.0 * p.p1t) / 8.0;
    out[e5::a54] = (9.0 * p.p2t) / 16.0;
    out[e5::a61] = (-77.0 * p.p1 + 59.0 * p.p2) / 42.0;
    out[e5::a62] = (8.0 * p.p1) / 7.0;
    out[e5::a63] = (111.0 * p.p1 - 87.0 * p.p2) / 28.0;
    out[e5::a65] = (-47.0 * p.p1 + 143.0 * p.p2) / 84.0;
    out[e5::a71] = (7.0 * (257.0 * p.p1 - 497.0 * p.p2 + 270.0 * p.p3)) / 2700.0;
    out[e5::a73] = (1097.0 * p.p1 - 467.0 * p.p2 - 150.0 * p.p3) / 1350.0;
    out[e5::a74] = (2.0 * (-49.0 * p.p1 + 199.0 * p.p2 - 135.0 * p.p3)) / 225.0;
    out[e5::a75] = (-313.0 * p.p1 + 883.0 * p.p2 - 90.0 * p.p3) / 1350.0;
    out[e5::a76] = (509.0 * p.p1 - 2129.0 * p.p2 + 1830.0 * p.p3) / 2700.0;
}

// IF45DP arrays (if45dp.py:204-237) for one mode; T = double (real lin_op) or cplx. out[] indexed by dp::
template <typename T>
RKS_HD void tableau_if45dp(T z, double h, int r4_fix, T* out) {
    const T E15 = cexp_t(z / 5.0), E310 = cexp_t(scale(3.0, z) / 10.0), E45 = cexp_t(scale(4.0, z) / 5.0);
    const T E89 = cexp_t(scale(8.0, z) / 9.0), E = cexp_t(z);
    const T E710 = cexp_t(scale(7.0, z) / 10.0), E19 = cexp_t(z / 9.0);
    out[dp::E15] = E15; out[dp::E310] = E310; out[dp::E45] = E45; out[dp::E89] = E89; out[dp::E] = E;
    out[dp::a21] = scale(h, E15) / 5.0;
    out[dp::a31] = scale(3.0 * h, E310) / 40.0;
    out[dp::a32] = scale(9.0 * h, cexp_t(z / 10.0)) / 40.0;
    out[dp::a41] = scale(44.0 * h, E45) / 45.0;
    out[dp::a42] = scale(-56.0 * h, cexp_t(scale(3.0, z) / 5.0)) / 15.0;
    out[dp::a43] = scale(32.0 * h, cexp_t(z / 2.0)) / 9.0;
    out[dp::a51] = scale(19372.0 * h, E89) / 6561.0;
    out[dp::a52] = scale(-25360.0 * h, cexp_t(scale(31.0, z) / 45.0)) / 2187.0;
    out[dp::a53] = scale(64448.0 * h, cexp_t(scale(53.0, z) / 90.0)) / 6561.0;
    out[dp::a54] = scale(-212.0 * h, cexp_t(scale(4.0, z) / 45.0)) / 729.0;
    out[dp::a61] = scale(9017.0 * h, E) / 3168.0;
    out[dp::a62] = scale(-355.0 * h, E45) / 33.0;
    out[dp::a63] = scale(46732.0 * h, E710) / 5247.0;
    out[dp::a64] = scale(49.0 * h, E15) / 176.0;
    out[dp::a65] = scale(-5103.0 * h, E19) / 18656.0;
    out[dp::a71] = scale(35.0 * h, E) / 384.0;
    out[dp::a73] = scale(500.0 * h, E710) / 1113.0;
    out[dp::a74] = scale(125.0 * h, E15) / 192.0;
    out[dp::a75] = scale(-2187.0 * h, E19) / 6784.0;
    out[dp::r1] = scale(h * 71.0, E) / 57600.0;
    out[dp::r3] = scale(-71.0 * h, E710) / 16695.0;
    out[dp::r4] = scale((r4_fix ? 71.0 : 17.0) * h, E15) / 1920.0;   // if45dp.py:234 ships 17
    out[dp::r5] = scale(-17253.0 * h, E19) / 339200.0;
}
// IF45DP scalars (if45dp.py:231,236,237)
// (evaluated by every thread of the last stage / the norm kernel: div_const, common.cuh, is the same correctly
// rounded quotient without the reciprocal iteration and its slow-path call)
RKS_HD double dp_a76(double h) { return div_const<84>(11.0 * h); }
RKS_HD double dp_r6(double h) { return div_const<525>(22.0 * h); }
RKS_HD double dp_r7(double h) { return div_const<40>(-h); }

}  // namespace rks
