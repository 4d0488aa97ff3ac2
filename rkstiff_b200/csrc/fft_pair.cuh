// K4 for the single-real-field model N = c rfft(irfft(u^)^3) (Allen-Cahn rows, 1-D and the last axis of 2-D grids):
// TWO ROWS PER COMPLEX TRANSFORM.  A real row run through the complex pipeline of fft_fast.cuh wastes half of both
// transforms (imaginary part identically zero on the way in, Hermitian redundancy on the way out).  Rows a and b
// are therefore packed into one complex row:
//     Z[k] = A^[k] + i B^[k]   (Hermitian extension of both half spectra; Im of DC / Nyquist dropped as c2r does)
//     ifft  ->  z = u_a + i u_b;   pointwise  ->  w = u_a^3 + i u_b^3;   fft  ->  W = F{u_a^3} + i F{u_b^3}
//     F{u_a^3}[k] = (W[k] + conj W[n-k]) / 2,     F{u_b^3}[k] = (W[k] - conj W[n-k]) / (2 i),   k = 0 .. n/2.
// Every pass is the unchanged fft_fast.cuh pass; only the first pass's loads (two rows) and the last pass's stores
// differ: the partner W[n-k] of an output in the lower half lives in another thread (the one that owns column
// Q1 - T of the last pass), so the upper half of W goes through the row's shared-memory slab once.  Half the FP64
// and shared-memory work per row of the full-length pair, no split twiddles.  The two rows exchange rounding errors
// at the 1e-16 level of the LARGER of them (they are neighbouring lines of one grid, or trajectories of one ensemble
// whose norms the controller combines anyway) -- inside the 1e-12 per-step bar, but a row's bits now depend on its
// neighbour.  tests/host_check runs these phases serially; tests/test_device_math_host.py pins them to NumPy.
#pragma once
#include "fft_fast.cuh"

namespace rks {
namespace fast {

struct PairedCubicModel {
    const cplx* in_a; const cplx* in_b;      // half spectra of the two rows (n/2 + 1 values each)
    cplx* out_a; cplx* out_b;
    double c; int n; bool on_a, on_b;        // row exists (odd tail: b is a row of zeros)
    RKS_HD cplx load(int p) const {
        const int hn = n >> 1;
        const int q = p <= hn ? p : n - p;
        const cplx va = row_ld(in_a + q);
        const cplx vb = on_b ? row_ld(in_b + q) : mk(0.0, 0.0);
        if (p == 0 || p == hn) return mk(va.x, vb.x);                  // c2r ignores Im of DC / Nyquist
        if (p < hn) return mk(va.x - vb.y, va.y + vb.x);               // A + i B
        return mk(va.x + vb.y, vb.x - va.y);                           // conj A + i conj B
    }
    RKS_HD cplx pointwise(cplx z) const {
        const double sc = 1.0 / (double)n;
        const double x = z.x * sc, y = z.y * sc;
        return mk(x * x * x, y * y * y);
    }
    RKS_HD void store(int, cplx) const {}    // (the generic last pass is not used: phase_pair_* below)
    // outputs k of both rows from W[k] and its partner W[n - k] (k = 0 and n/2: the partner is W[k] itself)
    RKS_HD void store_pair(int k, cplx wk, cplx wp) const {
        const double h = 0.5 * c;
        if (on_a) row_st(out_a + k, mk(h * (wk.x + wp.x), h * (wk.y - wp.y)));
        if (on_b) row_st(out_b + k, mk(h * (wk.y + wp.y), h * (wp.x - wk.x)));
    }
};

// last pass of a paired row, in three steps separated by row barriers (the caller places them):
//   load     the thread's NB last-pass butterflies leave the slab for registers
//   publish  forward butterflies; outputs in the upper half (k >= n/2) go back to the slab at position k
//   store    outputs k < n/2 (and k = n/2, owned by thread 0) meet their partners n - k and go to global memory
template <int N>
RKS_HD void phase_pair_load(const cplx* sm, int T, cplx* a) {
    using P = Plan<N>;
    constexpr int R1 = P::R1, Q1 = N / R1, TR = 32 * P::W, NB = Q1 / TR;
#pragma unroll
    for (int b = 0; b < NB; ++b) bf_load<R1, Q1, P::SH>(sm, T + TR * b, a + b * R1);
}
template <int N>
RKS_HD void phase_pair_publish(cplx* sm, int T, const Twiddles& tf, cplx* a) {
    using P = Plan<N>;
    constexpr int R1 = P::R1, Q1 = N / R1, TR = 32 * P::W, NB = Q1 / TR;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int p0 = T + TR * b;
        bf_dit<R1, Q1, TW_S1>(a + b * R1, tf.t1, p0);
#pragma unroll
        for (int r = R1 / 2; r < R1; ++r) sm[swz<P::SH>(p0 + Q1 * r)] = a[b * R1 + perm<R1>(r)];
    }
}
template <int N, class Model>
RKS_HD void phase_pair_store(const cplx* sm, int T, const cplx* a, const Model& m) {
    using P = Plan<N>;
    constexpr int R1 = P::R1, Q1 = N / R1, TR = 32 * P::W, NB = Q1 / TR;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int p0 = T + TR * b;
#pragma unroll
        for (int r = 0; r < R1 / 2; ++r) {
            const int k = p0 + Q1 * r;
            const cplx wk = a[b * R1 + perm<R1>(r)];
            const cplx wp = k == 0 ? wk : sm[swz<P::SH>(N - k)];
            m.store_pair(k, wk, wp);
        }
        if (p0 == 0) {
            const cplx wn = a[b * R1 + perm<R1>(R1 / 2)];               // k = n/2: its own partner
            m.store_pair(N / 2, wn, wn);
        }
    }
}

}  // namespace fast
}  // namespace rks
