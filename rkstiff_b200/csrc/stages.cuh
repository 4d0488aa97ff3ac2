// K1 -- per-stage linear combinations, one fused pass per RK stage.
// Restates the update_stages bodies of
//   rkstiff/if4.py:112-121, if34.py:120-130, etd4.py:167-174, etd34.py:182-190,
//   etd5.py:236-260, etd35.py:320-344, if45dp.py:140-179.
// `nv[j]` is the value of N_j at this element, `cv[slot]` the coefficient of this mode.
#pragma once
#include "common.cuh"
#include "coeffs.cuh"

namespace rks {

// which logical N buffers stage S of method M reads (bit j set => N_j)
RKS_HD constexpr unsigned stage_nl_mask(int M, int S) {
    if (M == M_IF4 || M == M_IF34)
        return S == 1 ? 0x02u : S == 2 ? 0x04u : S == 3 ? 0x08u : 0x1Eu;
    if (M == M_ETD4 || M == M_ETD34)
        return S == 1 ? 0x02u : S == 2 ? 0x06u : S == 3 ? 0x0Au : 0x1Eu;
    if (M == M_ETD5 || M == M_ETD35)
        return S == 1 ? 0x02u : S == 2 ? 0x06u : S == 3 ? 0x0Au : S == 4 ? 0x1Eu : S == 5 ? 0x3Eu : 0x7Au;
    // IF45DP
    return S == 1 ? 0x02u : S == 2 ? 0x06u : S == 3 ? 0x0Eu : S == 4 ? 0x1Eu : S == 5 ? 0x3Eu : 0x7Au;
}

// which coefficient slots stage S of method M reads
RKS_HD constexpr unsigned stage_coef_mask(int M, int S) {
#define B(x) (1u << (x))
    if (M == M_IF4 || M == M_IF34)
        return S <= 2 ? B(ifc::E2) : (B(ifc::E) | B(ifc::E2));
    if (M == M_ETD4 || M == M_ETD34)
        return S == 1 ? (B(kro::E2) | B(kro::a21)) : S == 2 ? (B(kro::E2) | B(kro::a31) | B(kro::a32))
             : S == 3 ? (B(kro::E) | B(kro::a41) | B(kro::a43))
                      : (B(kro::E) | B(kro::a51) | B(kro::a52) | B(kro::a54));
    if (M == M_ETD5 || M == M_ETD35)
        return S == 1 ? (B(e5::E14) | B(e5::a21)) : S == 2 ? (B(e5::E14) | B(e5::a31) | B(e5::a32))
             : S == 3 ? (B(e5::E12) | B(e5::a41) | B(e5::a43))
             : S == 4 ? (B(e5::E34) | B(e5::a51) | B(e5::a52) | B(e5::a54))
             : S == 5 ? (B(e5::E) | B(e5::a61) | B(e5::a62) | B(e5::a63) | B(e5::a65))
                      : (B(e5::E) | B(e5::a71) | B(e5::a73) | B(e5::a74) | B(e5::a75) | B(e5::a76));
    return S == 1 ? (B(dp::E15) | B(dp::a21)) : S == 2 ? (B(dp::E310) | B(dp::a31) | B(dp::a32))
         : S == 3 ? (B(dp::E45) | B(dp::a41) | B(dp::a42) | B(dp::a43))
         : S == 4 ? (B(dp::E89) | B(dp::a51) | B(dp::a52) | B(dp::a53) | B(dp::a54))
         : S == 5 ? (B(dp::E) | B(dp::a61) | B(dp::a62) | B(dp::a63) | B(dp::a64) | B(dp::a65))
                  : (B(dp::E) | B(dp::a71) | B(dp::a73) | B(dp::a74) | B(dp::a75));
#undef B
}

// N buffers / coefficient slots the embedded error estimate reads when it is formed inside the
// norm kernel (IF34, ETD34, IF45DP); ETD35 emits err from its last stage instead.
RKS_HD constexpr unsigned err_nl_mask(int M) {
    return (M == M_IF34 || M == M_ETD34) ? 0x30u : M == M_IF45DP ? 0xFAu : 0u;
}
RKS_HD constexpr unsigned err_coef_mask(int M) {
    return M == M_ETD34 ? (1u << kro::a54)
         : M == M_IF45DP ? ((1u << dp::r1) | (1u << dp::r3) | (1u << dp::r4) | (1u << dp::r5)) : 0u;
}


// ---------------------------------------------------------------------------------------
// CM_INDEXED: grouped coefficient records.  A grid with many modes has few DISTINCT lin_op values
// (|k|^2 on a Fourier grid), so K2 builds one record per distinct value and K1/K3 gather it through
// L2 with the mode's index.  Inside a record the slots one kernel reads are contiguous and start on a
// 32-byte sector: group g = 1..S holds the slots of stage g, group S+1 those of the embedded error
// estimate (a slot read by two stages is stored twice).
// ---------------------------------------------------------------------------------------
RKS_HD constexpr int popc32(unsigned v) { int c = 0; while (v) { v &= v - 1; ++c; } return c; }
RKS_HD constexpr unsigned group_mask(int M, int g) {
    return g <= method_stages(M) ? stage_coef_mask(M, g) : err_coef_mask(M);
}
template <typename CT> RKS_HD constexpr int group_pad(int n) {
    return (int)(((size_t)n * sizeof(CT) + 31) / 32 * 32 / sizeof(CT));
}
template <typename CT> RKS_HD constexpr int group_off(int M, int g) {      // in CT elements
    int off = 0;
    for (int q = 1; q < g; ++q) off += group_pad<CT>(popc32(group_mask(M, q)));
    return off;
}
template <typename CT> RKS_HD constexpr int record_elems(int M) { return group_off<CT>(M, method_stages(M) + 2); }
RKS_HD constexpr int group_pos(unsigned mask, int slot) { return popc32(mask & ((1u << slot) - 1u)); }

// ---------------------------------------------------------------------------------------
// CM_SEPARABLE (IF methods): every coefficient is rational * h * exp(q z) (if4.py:72-83, if45dp.py:204-237),
// and exp(q h sum_d a_d) = prod_d exp(q h a_d) when lin_op is a sum of per-axis terms.  sep_q: which of the
// distinct exponents q a slot uses; sep_scale: its rational * h factor (1 for the pure exponentials).
// IF45DP exponents: 0:1/5 1:3/10 2:4/5 3:8/9 4:1 5:1/10 6:3/5 7:1/2 8:31/45 9:53/90 10:4/45 11:7/10 12:1/9
// ---------------------------------------------------------------------------------------
RKS_HD constexpr int sep_nq(int M) { return M == M_IF45DP ? 13 : 2; }
RKS_HD constexpr int sep_q(int M, int slot) {
    if (M != M_IF45DP) return slot;                    // ifc::E -> exp(z), ifc::E2 -> exp(z/2)
    switch (slot) {
        case dp::E15: case dp::a21: case dp::a64: case dp::a74: case dp::r4: return 0;
        case dp::E310: case dp::a31: return 1;
        case dp::E45: case dp::a41: case dp::a62: return 2;
        case dp::E89: case dp::a51: return 3;
        case dp::E: case dp::a61: case dp::a71: case dp::r1: return 4;
        case dp::a32: return 5;
        case dp::a42: return 6;
        case dp::a43: return 7;
        case dp::a52: return 8;
        case dp::a53: return 9;
        case dp::a54: return 10;
        case dp::a63: case dp::a73: case dp::r3: return 11;
        default: return 12;                            // a65, a75, r5: exp(z/9)
    }
}
// exp(q z) for exponent id q, written like the reference writes the argument (if45dp.py:204-237)
template <typename T> RKS_HD T sep_exp(int M, int q, T z) {
    if (M != M_IF45DP) return q == 0 ? cexp_t(z) : cexp_t(z / 2.0);
    switch (q) {
        case 0: return cexp_t(z / 5.0);
        case 1: return cexp_t(scale(3.0, z) / 10.0);
        case 2: return cexp_t(scale(4.0, z) / 5.0);
        case 3: return cexp_t(scale(8.0, z) / 9.0);
        case 4: return cexp_t(z);
        case 5: return cexp_t(z / 10.0);
        case 6: return cexp_t(scale(3.0, z) / 5.0);
        case 7: return cexp_t(z / 2.0);
        case 8: return cexp_t(scale(31.0, z) / 45.0);
        case 9: return cexp_t(scale(53.0, z) / 90.0);
        case 10: return cexp_t(scale(4.0, z) / 45.0);
        case 11: return cexp_t(scale(7.0, z) / 10.0);
        default: return cexp_t(z / 9.0);
    }
}
RKS_HD double sep_scale(int M, int slot, double h, int r4_fix) {
    if (M != M_IF45DP) return 1.0;
    switch (slot) {
        case dp::a21: return h / 5.0;
        case dp::a31: return (3.0 * h) / 40.0;
        case dp::a32: return (9.0 * h) / 40.0;
        case dp::a41: return (44.0 * h) / 45.0;
        case dp::a42: return (-56.0 * h) / 15.0;
        case dp::a43: return (32.0 * h) / 9.0;
        case dp::a51: return (19372.0 * h) / 6561.0;
        case dp::a52: return (-25360.0 * h) / 2187.0;
        case dp::a53: return (64448.0 * h) / 6561.0;
        case dp::a54: return (-212.0 * h) / 729.0;
        case dp::a61: return (9017.0 * h) / 3168.0;
        case dp::a62: return (-355.0 * h) / 33.0;
        case dp::a63: return (46732.0 * h) / 5247.0;
        case dp::a64: return (49.0 * h) / 176.0;
        case dp::a65: return (-5103.0 * h) / 18656.0;
        case dp::a71: return (35.0 * h) / 384.0;
        case dp::a73: return (500.0 * h) / 1113.0;
        case dp::a74: return (125.0 * h) / 192.0;
        case dp::a75: return (-2187.0 * h) / 6784.0;
        case dp::r1: return (h * 71.0) / 57600.0;
        case dp::r3: return (-71.0 * h) / 16695.0;
        case dp::r4: return ((r4_fix ? 71.0 : 17.0) * h) / 1920.0;
        case dp::r5: return (-17253.0 * h) / 339200.0;
        default: return 1.0;                           // E15, E310, E45, E89, E
    }
}
// exponent ids a kernel needs: stage g of method M (g = S+1: the embedded error estimate)
RKS_HD constexpr unsigned sep_qmask(int M, int g) {
    unsigned m = 0;
    const unsigned slots = group_mask(M, g);
    for (int s = 0; s < method_ncoef(M); ++s)
        if (slots & (1u << s)) m |= 1u << sep_q(M, s);
    return m;
}

template <int M, int S, typename CT>
RKS_HD cplx stage_combine(cplx u, const cplx* nv, const CT* cv, double h) {
    if (M == M_IF4 || M == M_IF34) {
        // if4.py:112-120: coefficients are formed on the fly from h, E, E2
        if (S == 1) return cmul(cv[ifc::E2], u) + cmul(scale(h, cv[ifc::E2]), nv[1]) / 2.0;
        if (S == 2) return cmul(cv[ifc::E2], u) + (h * nv[2]) / 2.0;
        if (S == 3) return cmul(cv[ifc::E], u) + cmul(scale(h, cv[ifc::E2]), nv[3]);
        return cmul(cv[ifc::E], u)
             + h * (div_const<6>(cmul(cv[ifc::E], nv[1])) + div_const<3>(cmul(cv[ifc::E2], nv[2]))
                    + div_const<3>(cmul(cv[ifc::E2], nv[3])) + div_const<6>(nv[4]));
    }
    if (M == M_ETD4 || M == M_ETD34) {
        if (S == 1) return cmul(cv[kro::E2], u) + cmul(cv[kro::a21], nv[1]);
        if (S == 2) return cmul(cv[kro::E2], u) + cmul(cv[kro::a31], nv[1]) + cmul(cv[kro::a32], nv[2]);
        if (S == 3) return cmul(cv[kro::E], u) + cmul(cv[kro::a41], nv[1]) + cmul(cv[kro::a43], nv[3]);
        return cmul(cv[kro::E], u) + cmul(cv[kro::a51], nv[1]) + cmul(cv[kro::a52], nv[2] + nv[3])
             + cmul(cv[kro::a54], nv[4]);
    }
    if (M == M_ETD5 || M == M_ETD35) {
        if (S == 1) return cmul(cv[e5::E14], u) + cmul(cv[e5::a21], nv[1]);
        if (S == 2) return cmul(cv[e5::E14], u) + cmul(cv[e5::a31], nv[1]) + cmul(cv[e5::a32], nv[2]);
        if (S == 3) return cmul(cv[e5::E12], u) + cmul(cv[e5::a41], nv[1]) + cmul(cv[e5::a43], nv[3]);
        if (S == 4)
            return cmul(cv[e5::E34], u) + cmul(cv[e5::a51], nv[1]) + cmul(cv[e5::a52], nv[2] - nv[3])
                 + cmul(cv[e5::a54], nv[4]);
        if (S == 5)
            return cmul(cv[e5::E], u) + cmul(cv[e5::a61], nv[1])
                 + cmul(cv[e5::a62], nv[2] - (3.0 * nv[4]) / 2.0) + cmul(cv[e5::a63], nv[3])
                 + cmul(cv[e5::a65], nv[5]);
        return cmul(cv[e5::E], u) + cmul(cv[e5::a71], nv[1]) + cmul(cv[e5::a73], nv[3])
             + cmul(cv[e5::a74], nv[4]) + cmul(cv[e5::a75], nv[5]) + cmul(cv[e5::a76], nv[6]);
    }
    // IF45DP
    if (S == 1) return cmul(cv[dp::E15], u) + cmul(cv[dp::a21], nv[1]);
    if (S == 2) return cmul(cv[dp::E310], u) + cmul(cv[dp::a31], nv[1]) + cmul(cv[dp::a32], nv[2]);
    if (S == 3)
        return cmul(cv[dp::E45], u) + cmul(cv[dp::a41], nv[1]) + cmul(cv[dp::a42], nv[2])
             + cmul(cv[dp::a43], nv[3]);
    if (S == 4)
        return cmul(cv[dp::E89], u) + cmul(cv[dp::a51], nv[1]) + cmul(cv[dp::a52], nv[2])
             + cmul(cv[dp::a53], nv[3]) + cmul(cv[dp::a54], nv[4]);
    if (S == 5)
        return cmul(cv[dp::E], u) + cmul(cv[dp::a61], nv[1]) + cmul(cv[dp::a62], nv[2])
             + cmul(cv[dp::a63], nv[3]) + cmul(cv[dp::a64], nv[4]) + cmul(cv[dp::a65], nv[5]);
    return cmul(cv[dp::E], u) + cmul(cv[dp::a71], nv[1]) + cmul(cv[dp::a73], nv[3])
         + cmul(cv[dp::a74], nv[4]) + cmul(cv[dp::a75], nv[5]) + dp_a76(h) * nv[6];
}

// ETD35 error estimate, formed in the last stage kernel (etd35.py:344)
template <typename CT>
RKS_HD cplx etd35_err(const cplx* nv, const CT* cv) {
    return cmul(cv[e5::a75], -nv[1] + 4.0 * nv[3] - 6.0 * nv[4] + 4.0 * nv[5] - nv[6]);
}

// error estimates that need N(u+) and are therefore formed in the norm kernel
// (if34.py:130, etd34.py:190, if45dp.py:172-179)
template <int M, typename CT>
RKS_HD cplx embedded_err(const cplx* nv, const CT* cv, double h) {
    if (M == M_IF34) return div_const<6>(h * (nv[4] - nv[5]));
    if (M == M_ETD34) return cmul(cv[kro::a54], nv[4] - nv[5]);
    return cmul(cv[dp::r1], nv[1]) + cmul(cv[dp::r3], nv[3]) + cmul(cv[dp::r4], nv[4])
         + cmul(cv[dp::r5], nv[5]) + dp_r6(h) * nv[6] + dp_r7(h) * nv[7];
}

}  // namespace rks
