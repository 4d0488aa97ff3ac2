// EXPERIMENT (opt-in, RKS_K4_X2=1, not measured yet): the passes of fft_fast.cuh with TWO logical threads per physical
// thread.  nl_fast_kernel at n = 8192 runs 16 warps x 128 registers; ncu shows the FP64 pipe 53 % and the LSU 62 %
// busy with little overlap: a thread loads one butterfly, transforms it, stores it, and only then
// loads the next, and at 128 registers the compiler cannot hoist the second butterfly's loads (an earlier attempt
// spilled).  Here a CTA has 8 warps x 255 registers; physical thread t plays the logical threads of two adjacent
// logical warps (their two 512-point slices belong to the same physical warp, so the warp-local passes stay
// warp-local), and every pass is written "all loads, all butterflies, all stores" over the 32 values a thread now
// owns -- independent instruction streams the scheduler can interleave.  Same operations on the same data: the
// results are bit-identical to the one-thread-per-logical-thread passes (tests/host_check, test_device_math_host.py).
#pragma once
#include "fft_fast.cuh"

namespace rks {
namespace fast {

// logical threads of physical thread t: warps 2w and 2w + 1, same lane
RKS_HD int x2_first(int t) { return ((t >> 5) << 6) + (t & 31); }
RKS_HD int x2_second(int t) { return x2_first(t) + 32; }

template <int R, int Q, int SH, int NB2, int TS, bool GLOBAL_IN, class Model>
RKS_HD void dif_pass_x2(cplx* sm, const int (&p0)[NB2], const int (&j)[NB2], const cplx* tab, const Model& m) {
    cplx a[NB2][R];
#pragma unroll
    for (int b = 0; b < NB2; ++b) {
        if (GLOBAL_IN) bf_load_global<R, Q>(m, p0[b], a[b]);
        else bf_load<R, Q, SH>(sm, p0[b], a[b]);
    }
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_dif<R, Q, TS>(a[b], tab, j[b]);
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_store<R, Q, SH>(sm, p0[b], a[b]);
}
template <int R, int Q, int SH, int NB2, int TS, bool GLOBAL_OUT, class Model>
RKS_HD void dit_pass_x2(cplx* sm, const int (&p0)[NB2], const int (&j)[NB2], const cplx* tab, const Model& m) {
    cplx a[NB2][R];
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_load<R, Q, SH>(sm, p0[b], a[b]);
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_dit<R, Q, TS>(a[b], tab, j[b]);
#pragma unroll
    for (int b = 0; b < NB2; ++b) {
        if (GLOBAL_OUT) bf_store_global<R, Q>(m, p0[b], a[b]);
        else bf_store<R, Q, SH>(sm, p0[b], a[b]);
    }
}
template <int R, int SH, int NB2, class Model>
RKS_HD void core_pass_x2(cplx* sm, const int (&p0)[NB2], const Model& m) {
    cplx a[NB2][R];
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_load<R, 1, SH>(sm, p0[b], a[b]);
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_core<R>(a[b], m);
#pragma unroll
    for (int b = 0; b < NB2; ++b) bf_store<R, 1, SH>(sm, p0[b], a[b]);
}

// row-level passes (first / last): logical thread T owns the butterflies T + 32 W c
template <int N, int NB>
RKS_HD void x2_row_butterflies(int t, int (&p0)[2 * NB], int (&j)[2 * NB]) {
    using P = Plan<N>;
    const int Ta = x2_first(t), Tb = x2_second(t);
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        p0[c] = Ta + 32 * P::W * c; j[c] = p0[c];
        p0[NB + c] = Tb + 32 * P::W * c; j[NB + c] = p0[NB + c];
    }
}
// warp-local passes: the two logical warps' 512-point slices
template <int R, int Q, int NB>
RKS_HD void x2_warp_butterflies(int t, int (&p0)[2 * NB], int (&j)[2 * NB]) {
    int pa[NB], ja[NB], pb[NB], jb[NB];
    warp_butterflies<R, Q, NB>((x2_first(t) >> 5) * 512, t & 31, pa, ja);
    warp_butterflies<R, Q, NB>((x2_second(t) >> 5) * 512, t & 31, pb, jb);
#pragma unroll
    for (int c = 0; c < NB; ++c) { p0[c] = pa[c]; j[c] = ja[c]; p0[NB + c] = pb[c]; j[NB + c] = jb[c]; }
}

template <int N, class Model>
RKS_HD void phase_first_x2(cplx* sm, int t, const Twiddles& ti, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, NB = Q1 / (32 * P::W);
    int p0[2 * NB], j[2 * NB];
    x2_row_butterflies<N, NB>(t, p0, j);
    dif_pass_x2<P::R1, Q1, P::SH, 2 * NB, TW_S1, true>(sm, p0, j, ti.t1, m);
}
template <int N, class Model>
RKS_HD void phase_last_x2(cplx* sm, int t, const Twiddles& tf, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, NB = Q1 / (32 * P::W);
    int p0[2 * NB], j[2 * NB];
    x2_row_butterflies<N, NB>(t, p0, j);
    dit_pass_x2<P::R1, Q1, P::SH, 2 * NB, TW_S1, true>(sm, p0, j, tf.t1, m);
}
template <int N, int K, bool DIF, class Model>
RKS_HD void phase_middle_x2(cplx* sm, int t, const Twiddles& tw, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1;
    constexpr int R = K == 2 ? P::R2 : P::R3;
    constexpr int Q = K == 2 ? Q1 / P::R2 : Q1 / P::R2 / P::R3;
    constexpr int NB = 16 / R;
    int p0[2 * NB], j[2 * NB];
    x2_warp_butterflies<R, Q, NB>(t, p0, j);
    const cplx* tab = K == 2 ? tw.t2 : tw.t3;
    constexpr int TS = K == 2 ? TW_S2 : TW_S3;
    if (DIF) dif_pass_x2<R, Q, P::SH, 2 * NB, TS, false>(sm, p0, j, tab, m);
    else dit_pass_x2<R, Q, P::SH, 2 * NB, TS, false>(sm, p0, j, tab, m);
}
// first pass of a pre-transformed row (fft_fast.cuh phase_pre)
template <int N, class Model>
RKS_HD void phase_pre_x2(cplx* sm, int t, const Twiddles& ti, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, R = P::R2, Q = Q1 / P::R2, NB = 16 / R;
    int p0[2 * NB], j[2 * NB];
    x2_warp_butterflies<R, Q, NB>(t, p0, j);
    dif_pass_x2<R, Q, P::SH, 2 * NB, TW_S2, true>(sm, p0, j, ti.t2, m);
}
template <int N, class Model>
RKS_HD void phase_core_x2(cplx* sm, int t, const Model& m) {
    using P = Plan<N>;
    constexpr int R = P::R4 > 1 ? P::R4 : P::R3;
    constexpr int NB = 16 / R;
    int p0[2 * NB], j[2 * NB];
    x2_warp_butterflies<R, 1, NB>(t, p0, j);
    core_pass_x2<R, P::SH, 2 * NB>(sm, p0, m);
}

}  // namespace fast
}  // namespace rks
