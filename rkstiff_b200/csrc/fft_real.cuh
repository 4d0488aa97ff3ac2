// K4 for the real-field (rfft) models: forward transform of the REAL pointwise product as a half-length complex
// transform.  EXPERIMENT (DESIGN.md section 6 "what comes next"): this header holds the per-thread phase functions,
// tests/host_check runs them serially and tests/test_device_math_host.py pins them to NumPy; the only kernel that uses
// them is the opt-in, not yet measured nl_fast_real_kernel (kernels.cuh, RKS_RFFT_HALF=1).  The default kernels do not.
//
// Why: the u u_x / cubic models run a full-length complex inverse transform (it carries two real fields, or one with
// half the butterflies idle) and then a full-length complex FORWARD transform of a real signal w -- half of that
// work is redundant, and at n = 1024 the pair is FP64 bound.  With c[m] = w[2m] + i w[2m+1]:
//     C = FFT_{n/2}(c),   E[k] = (C[k] + conj C[n/2-k]) / 2,   O[k] = (C[k] - conj C[n/2-k]) / (2i),
//     W[k] = E[k] + w_n^k O[k],  k = 0 .. n/2     (C[n/2] := C[0]).
// How it fits the in-place digit-reversed layout of fft_fast.cuh: after the inverse DIF passes position p holds time
// sample j whose digits are those of p reversed; in particular the first-pass block b = p / Q1 is j mod R1, so the
// samples 2m and 2m+1 sit Q1 positions apart, in the blocks 2a and 2a+1 of one warp's slice (n <= 4096).  The core
// pass therefore takes the butterflies of an even block and of the odd block after it, forms c in registers and
// runs ONE forward butterfly, stored in the even block; the forward middle pass only visits even blocks; the last
// pass is a radix-R1/2 butterfly over the even blocks (twiddles w_{n/2}^{a k'}) followed by the split above, whose
// partner C[n/2-k] lives in another thread: one shared-memory exchange between two row barriers.
// Forward cost: half the butterflies of every pass, radix R1/2 instead of R1 in the last one, + ~10 flops per mode.
#pragma once
#include "fft_fast.cuh"

namespace rks {
namespace fast {

// v[slot(r)] *= w^r for r = 1..R-1 from the VALUE w (twiddle_scale reads it from a table); R <= 8
template <int R, bool INV, class Slot>
RKS_HD void twiddle_scale_value(cplx* v, cplx w, Slot slot) {
    static_assert(R <= 8, "radix of the half-length last pass");
    if (R == 1) return;
    const cplx w1 = cj<INV>(w);
    v[slot(1)] = v[slot(1)] * w1;
    if (R == 2) return;
    const cplx w2 = w1 * w1, w3 = w1 * w2;
    v[slot(2)] = v[slot(2)] * w2;
    v[slot(3)] = v[slot(3)] * w3;
    if (R == 4) return;
    const cplx w4 = w2 * w2;
    v[slot(4)] = v[slot(4)] * w4;
    v[slot(5)] = v[slot(5)] * (w4 * w1);
    v[slot(6)] = v[slot(6)] * (w4 * w2);
    v[slot(7)] = v[slot(7)] * (w4 * w3);
}

// exp(-2 pi i t / R1), t < R1 / 2: the factor between w_n^{k'} and w_n^{k' + Q1 t}
template <int R1>
RKS_HD cplx omega_first(int t) {
    if (R1 == 16) {
        switch (t) {
            case 0: return mk(1.0, 0.0);
            case 1: return mk(C8, -S8);
            case 2: return mk(SQH, -SQH);
            case 3: return mk(S8, -C8);
            case 4: return mk(0.0, -1.0);
            case 5: return mk(-S8, -C8);
            case 6: return mk(-SQH, -SQH);
            default: return mk(-C8, -S8);
        }
    }
    switch (t) {                         // R1 = 8
        case 0: return mk(1.0, 0.0);
        case 1: return mk(SQH, -SQH);
        case 2: return mk(0.0, -1.0);
        default: return mk(-SQH, -SQH);
    }
}

// core pass: inverse butterflies of an even block and of the odd block after it, the pointwise product of each
// (real), c = even + i odd, ONE forward butterfly, stored in the even block
template <int N, class Model>
RKS_HD void phase_core_pair(cplx* sm, int T, const Model& m) {
    using P = Plan<N>;
    static_assert(P::R4 == 1 && N <= 4096, "the blocks 2a and 2a+1 must lie in one warp's slice");
    constexpr int Q1 = N / P::R1, R = P::R3, BLK = 512 / Q1, PER = Q1 / R, PAIRS = (BLK / 2) * PER;
    const int chunk0 = (T >> 5) * 512, l = T & 31;
    for (int u = l; u < PAIRS; u += 32) {
        const int pe = chunk0 + 2 * (u / PER) * Q1 + (u % PER) * R;
        cplx x[R], c[R];
        bf_load<R, 1, P::SH>(sm, pe, x);
        dftR<R, true>(x);
#pragma unroll
        for (int r = 0; r < R; ++r) c[r].x = m.pointwise(x[perm<R>(r)]).x;
        bf_load<R, 1, P::SH>(sm, pe + Q1, x);
        dftR<R, true>(x);
#pragma unroll
        for (int r = 0; r < R; ++r) c[r].y = m.pointwise(x[perm<R>(r)]).x;
        dftR<R, false>(c);
        bf_store<R, 1, P::SH>(sm, pe, c);
    }
}

// forward middle pass (pass 2 of the plan) on the even blocks only
template <int N>
RKS_HD void phase_middle_even(cplx* sm, int T, const Twiddles& tf) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, R = P::R2, Q = Q1 / P::R2, BLK = 512 / Q1, BF = (BLK / 2) * Q;
    const int chunk0 = (T >> 5) * 512, l = T & 31;
    for (int u = l; u < BF; u += 32) {
        const int j = u % Q, p0 = chunk0 + 2 * (u / Q) * Q1 + j;
        cplx x[R];
        bf_load<R, Q, P::SH>(sm, p0, x);
        bf_dit<R, Q, TW_S2>(x, tf.t2, j);
        bf_store<R, Q, P::SH>(sm, p0, x);
    }
}

// last pass, part A: C[k' + Q1 t], t < R1/2, of the half-length transform in registers (slot perm<H>(t));
// k' = T + 32 W c.  Row barrier, then part B.
template <int N>
RKS_HD void phase_last_half_load(const cplx* sm, int T, const Twiddles& tf, cplx* x /*[NB][R1/2]*/) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, H = P::R1 / 2, NB = Q1 / (32 * P::W);
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const int kp = T + 32 * P::W * c;
        cplx* v = x + c * H;
#pragma unroll
        for (int a = 0; a < H; ++a) v[a] = sm[swz<P::SH>(kp + Q1 * 2 * a)];
        const cplx w1 = tw_ld(tf.t1 + kp);                   // w_n^{k'}
        twiddle_scale_value<H, false>(v, w1 * w1, SlotId());  // w_{n/2}^{a k'}
        dftR<H, false>(v);
    }
}
// part B: C to shared memory in natural order (positions k < n/2).  Row barrier, then part C.
template <int N>
RKS_HD void phase_last_half_exchange(cplx* sm, int T, const cplx* x) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, H = P::R1 / 2, NB = Q1 / (32 * P::W);
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const int kp = T + 32 * P::W * c;
#pragma unroll
        for (int t = 0; t < H; ++t) sm[swz<P::SH>(kp + Q1 * t)] = x[c * H + perm<H>(t)];
    }
}
// part C: the split with the partner C[n/2 - k] and the store of W[k] (and W[n/2] by the thread that owns k = 0)
template <int N, class Model>
RKS_HD void phase_last_half_split(const cplx* sm, int T, const Twiddles& tf, const cplx* x, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, H = P::R1 / 2, NB = Q1 / (32 * P::W), HN = N / 2;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const int kp = T + 32 * P::W * c;
        const cplx w1 = tw_ld(tf.t1 + kp);
#pragma unroll
        for (int t = 0; t < H; ++t) {
            const int k = kp + Q1 * t;
            const cplx ck = x[c * H + perm<H>(t)];
            const cplx cp = conj(sm[swz<P::SH>((HN - k) & (HN - 1))]);
            const cplx e = mk(0.5 * (ck.x + cp.x), 0.5 * (ck.y + cp.y));
            const cplx d = mk(ck.x - cp.x, ck.y - cp.y);
            const cplx o = mk(0.5 * d.y, -0.5 * d.x);                 // d / (2i)
            m.store(k, e + (w1 * omega_first<P::R1>(t)) * o);
            if (k == 0) m.store(HN, e - o);                            // w_n^{n/2} = -1
        }
    }
}

}  // namespace fast
}  // namespace rks
