// Stage combine folded into the nonlinear-term kernel (K1 fused into K4's load prologue).
//
// A stage value k = E u + sum_j a_sj N_j is described at run time as a short list of terms
//     k = sum_t (c0_t + c1_t h) * coef[slot_t] (.) X[src_t]        (slot < 0: no coefficient array)
// so that ONE kernel per (row length, model, coefficient type) serves every method and stage.
// The list is built from the same tableaux as stages.cuh (reference lines cited there); products
// such as a62 (N2 - 3 N4 / 2) are expanded into two terms, which changes the association of the
// floating-point sum by a few ulp (well inside the 1e-12 per-step parity bar).
#pragma once
#include "common.cuh"

namespace rks {

constexpr int FUSE_MAX_TERMS = 7;

struct FuseDesc {
    int nterms;                       // 0: plain nonlinear evaluation of an existing array
    int src[FUSE_MAX_TERMS];          // 0 = u, 1..7 = logical N_j
    int slot[FUSE_MAX_TERMS];         // coefficient slot, -1 = none
    double c0[FUSE_MAX_TERMS], c1[FUSE_MAX_TERMS];
    int write_k;                      // final stage: also store k as the new state (candidate / in place)
    int track_max;                    // adaptive final stage: atomicMax of |k|^2 (solveras.py:451)
};

struct FuseBuilder {
    FuseDesc d;
    RKS_HD FuseBuilder() { d.nterms = 0; d.write_k = 0; d.track_max = 0; }
    RKS_HD FuseBuilder& term(int src, int slot, double c0 = 1.0, double c1 = 0.0) {
        d.src[d.nterms] = src; d.slot[d.nterms] = slot; d.c0[d.nterms] = c0; d.c1[d.nterms] = c1;
        d.nterms += 1;
        return *this;
    }
};

// descriptor of stage S (1-based) of method M
RKS_HD FuseDesc fuse_desc(int M, int S) {
    FuseBuilder b;
    if (M == M_IF4 || M == M_IF34) {                        // if4.py:112-120
        if (S == 1) b.term(0, ifc::E2).term(1, ifc::E2, 0.0, 0.5);
        else if (S == 2) b.term(0, ifc::E2).term(2, -1, 0.0, 0.5);
        else if (S == 3) b.term(0, ifc::E).term(3, ifc::E2, 0.0, 1.0);
        else b.term(0, ifc::E).term(1, ifc::E, 0.0, 1.0 / 6.0).term(2, ifc::E2, 0.0, 1.0 / 3.0)
              .term(3, ifc::E2, 0.0, 1.0 / 3.0).term(4, -1, 0.0, 1.0 / 6.0);
    } else if (M == M_ETD4 || M == M_ETD34) {               // etd4.py:167-173
        if (S == 1) b.term(0, kro::E2).term(1, kro::a21);
        else if (S == 2) b.term(0, kro::E2).term(1, kro::a31).term(2, kro::a32);
        else if (S == 3) b.term(0, kro::E).term(1, kro::a41).term(3, kro::a43);
        else b.term(0, kro::E).term(1, kro::a51).term(2, kro::a52).term(3, kro::a52).term(4, kro::a54);
    } else if (M == M_ETD5 || M == M_ETD35) {               // etd5.py:236-258
        if (S == 1) b.term(0, e5::E14).term(1, e5::a21);
        else if (S == 2) b.term(0, e5::E14).term(1, e5::a31).term(2, e5::a32);
        else if (S == 3) b.term(0, e5::E12).term(1, e5::a41).term(3, e5::a43);
        else if (S == 4) b.term(0, e5::E34).term(1, e5::a51).term(2, e5::a52).term(3, e5::a52, -1.0).term(4, e5::a54);
        else if (S == 5) b.term(0, e5::E).term(1, e5::a61).term(2, e5::a62).term(4, e5::a62, -1.5).term(3, e5::a63)
                          .term(5, e5::a65);
        else b.term(0, e5::E).term(1, e5::a71).term(3, e5::a73).term(4, e5::a74).term(5, e5::a75).term(6, e5::a76);
    } else {                                                // if45dp.py:140-171
        if (S == 1) b.term(0, dp::E15).term(1, dp::a21);
        else if (S == 2) b.term(0, dp::E310).term(1, dp::a31).term(2, dp::a32);
        else if (S == 3) b.term(0, dp::E45).term(1, dp::a41).term(2, dp::a42).term(3, dp::a43);
        else if (S == 4) b.term(0, dp::E89).term(1, dp::a51).term(2, dp::a52).term(3, dp::a53).term(4, dp::a54);
        else if (S == 5) b.term(0, dp::E).term(1, dp::a61).term(2, dp::a62).term(3, dp::a63).term(4, dp::a64)
                          .term(5, dp::a65);
        else b.term(0, dp::E).term(1, dp::a71).term(3, dp::a73).term(4, dp::a74).term(5, dp::a75)
              .term(6, -1, 0.0, 11.0 / 84.0);
    }
    return b.d;
}

// resolved form used inside the kernel: pointers instead of indices, scale with h folded in
template <typename CT>
struct FuseSource {
    const cplx* x[FUSE_MAX_TERMS];    // row base pointers of the sources
    const CT* c[FUSE_MAX_TERMS];      // coefficient arrays (nullptr: none)
    double sc[FUSE_MAX_TERMS];
    int nterms;
    RKS_HD cplx value(long long p) const {
        cplx acc = mk(0.0, 0.0);
#pragma unroll
        for (int t = 0; t < FUSE_MAX_TERMS; ++t) {
            if (t < nterms) {
#if defined(__CUDA_ARCH__)
                const double2 v = __ldcs(reinterpret_cast<const double2*>(x[t] + p));
                cplx term = mk(v.x, v.y);
                if (c[t]) term = cmul(ld_coef(c[t] + p), term);
#else
                cplx term = x[t][p];
                if (c[t]) term = cmul(c[t][p], term);
#endif
                if (sc[t] != 1.0) term = sc[t] * term;
                acc = t == 0 ? term : acc + term;
            }
        }
        return acc;
    }
#if defined(__CUDA_ARCH__)
    static __device__ __forceinline__ double ld_coef(const double* q) { return __ldg(q); }
    static __device__ __forceinline__ cplx ld_coef(const cplx* q) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(q));
        return mk(v.x, v.y);
    }
#endif
};

}  // namespace rks
