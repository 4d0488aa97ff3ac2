// K4 fast path -- register-resident FP64 FFT pair (inverse, pointwise nonlinearity, forward)
// for rows of n = 512*W points, W in {1,2,4,8,16}  (n = 512 ... 8192).
//
// Layout.  A row is owned by W warps; every thread keeps 16 complex values in registers.
//   outer pass   radix-W across the W chunks of 512 points   (row-level exchange through smem)
//   inner passes radix-8 x 8 x 8 on each 512-point chunk, one warp per chunk, exchanges through
//                the warp's own 8 KB smem slice, synchronised with __syncwarp only
// The inverse transform is decimation-in-frequency and the forward one its mirrored
// decimation-in-time, so the time-domain data stay in digit-reversed order and the pointwise
// nonlinearity sits between the two innermost radix-8 butterflies without any exchange:
//   load -> [W] -x- [8] -x- [8] -x- [8] -> N(.) -> [8] -x- [8] -x- [8] -x- [W] -> store
// (-x- = shared-memory exchange: 2 row-level, 4 warp-level per nonlinear evaluation).
// Shared-memory positions are XOR-swizzled (swz) so that every 128-bit access of a quarter-warp
// is bank-conflict free.  The functions are __host__ __device__: tests/host_check executes the
// same phase sequence serially on the CPU.
#pragma once
#include "common.cuh"

namespace rks {
namespace fast {

constexpr double SQH = 0.70710678118654752440;   // sqrt(1/2)
constexpr double C8 = 0.92387953251128675613;    // cos(pi/8)
constexpr double S8 = 0.38268343236508977173;    // sin(pi/8)

RKS_HD int swz(int p) { return p ^ ((p >> 3) & 7); }

template <bool INV> RKS_HD cplx rot(cplx a) { return INV ? mul_i(a) : mul_mi(a); }     // * (+-i)
// multiply by exp(-+ i pi/4 * k)-type constants: w = (c, -s) forward, (c, +s) inverse
template <bool INV> RKS_HD cplx mulw(cplx a, double c, double s) {
    return INV ? mk(a.x * c - a.y * s, a.x * s + a.y * c) : mk(a.x * c + a.y * s, a.y * c - a.x * s);
}

// In-place DFTs on register arrays.  Output k of dftR lives in slot permR(k).
RKS_HD constexpr int perm2(int k) { return k; }
RKS_HD constexpr int perm4(int k) { return k; }
RKS_HD constexpr int perm8(int k) { return 2 * (k & 3) + (k >> 2); }
RKS_HD constexpr int perm16(int k) { return 4 * (k & 3) + (k >> 2); }
template <int R> RKS_HD constexpr int perm(int k) { return R == 16 ? perm16(k) : R == 8 ? perm8(k) : k; }

template <bool INV> RKS_HD void dft2(cplx& a, cplx& b) {
    const cplx t = a;
    a = t + b;
    b = t - b;
}
template <bool INV> RKS_HD void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
    const cplx t0 = x0 + x2, t1 = x0 - x2, t2 = x1 + x3, t3 = rot<INV>(x1 - x3);
    x0 = t0 + t2; x1 = t1 + t3; x2 = t0 - t2; x3 = t1 - t3;
}
template <bool INV> RKS_HD void dft8(cplx* v) {
    dft4<INV>(v[0], v[2], v[4], v[6]);                  // E[k1] -> slot 2 k1
    dft4<INV>(v[1], v[3], v[5], v[7]);                  // O[k1] -> slot 2 k1 + 1
    v[3] = mulw<INV>(v[3], SQH, SQH);                   // * w8^1
    v[5] = rot<INV>(v[5]);                              // * w8^2
    v[7] = mulw<INV>(v[7], -SQH, SQH);                  // * w8^3
    dft2<INV>(v[0], v[1]); dft2<INV>(v[2], v[3]); dft2<INV>(v[4], v[5]); dft2<INV>(v[6], v[7]);
}
template <bool INV> RKS_HD void dft16(cplx* v) {
    dft4<INV>(v[0], v[4], v[8], v[12]);                 // T[b][k1] -> slot 4 k1 + b
    dft4<INV>(v[1], v[5], v[9], v[13]);
    dft4<INV>(v[2], v[6], v[10], v[14]);
    dft4<INV>(v[3], v[7], v[11], v[15]);
    // * w16^(b k1)
    v[5] = mulw<INV>(v[5], C8, S8);      v[6] = mulw<INV>(v[6], SQH, SQH);     v[7] = mulw<INV>(v[7], S8, C8);
    v[9] = mulw<INV>(v[9], SQH, SQH);    v[10] = rot<INV>(v[10]);              v[11] = mulw<INV>(v[11], -SQH, SQH);
    v[13] = mulw<INV>(v[13], S8, C8);    v[14] = mulw<INV>(v[14], -SQH, SQH);  v[15] = mulw<INV>(v[15], -C8, -S8);
    dft4<INV>(v[0], v[1], v[2], v[3]);                  // y[k1 + 4 k2] -> slot 4 k1 + k2
    dft4<INV>(v[4], v[5], v[6], v[7]);
    dft4<INV>(v[8], v[9], v[10], v[11]);
    dft4<INV>(v[12], v[13], v[14], v[15]);
}
template <int R, bool INV> RKS_HD void dftR(cplx* v) {
    if (R == 2) dft2<INV>(v[0], v[1]);
    else if (R == 4) dft4<INV>(v[0], v[1], v[2], v[3]);
    else if (R == 8) dft8<INV>(v);
    else if (R == 16) dft16<INV>(v);
}

// Twiddle tables, laid out so that the lanes of a warp read consecutive entries (a gather
// tw[r*i] costs up to 32 L1 wavefronts per load, a contiguous read 4).  Only the powers
// 1, 2, 4, 8 are stored; the others are products of at most three of them.
//   o[k][i] = w_n^(k i),   k in {1,2,4,8}, i < 512     (outer radix-W pass)
//   a[k][i] = w_512^(k i), k in {1,2,4},   i < 64      (inner pass A)
//   b[k][i] = w_64^(k i),  k in {1,2,4},   i < 8       (inner pass B)
// with w_m = exp(-2 pi i / m).
struct Twiddles {
    const cplx* o;      // 4 x 512
    const cplx* a;      // 3 x 64
    const cplx* b;      // 3 x 8
};
constexpr int TW_O = 0, TW_A = 4 * 512, TW_B = TW_A + 3 * 64, TW_TOTAL = TW_B + 3 * 8;

RKS_HD cplx tw_ld(const cplx* p) {
#if defined(__CUDA_ARCH__)
    const double2 t = __ldg(reinterpret_cast<const double2*>(p));
    return mk(t.x, t.y);
#else
    return *p;
#endif
}
template <bool INV> RKS_HD cplx cj(cplx w) { return INV ? conj(w) : w; }
struct SlotId { RKS_HD int operator()(int r) const { return r; } };
template <int R> struct SlotPerm { RKS_HD int operator()(int r) const { return perm<R>(r); } };

// entry `idx` of the concatenated table block [o | a | b] for an n-point row
RKS_HD cplx twiddle_table_entry(int idx, int n) {
    int k, i, m;
    if (idx < TW_A) { k = 1 << (idx / 512); i = idx % 512; m = n; }
    else if (idx < TW_B) { const int t = idx - TW_A; k = 1 << (t / 64); i = t % 64; m = 512; }
    else { const int t = idx - TW_B; k = 1 << (t / 8); i = t % 8; m = 64; }
    const long long e = ((long long)k * i) % m;
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(-2.0 * (double)e / (double)m, &s, &c);
#else
    s = sin(-2.0 * M_PI * (double)e / (double)m); c = cos(-2.0 * M_PI * (double)e / (double)m);
#endif
    return mk(c, s);
}

// v[slot(r)] *= w^r for r = 1..R-1, with w^1, w^2, w^4, w^8 read from tab[k*stride + i]
template <int R, bool INV, class Slot>
RKS_HD void twiddle_scale(cplx* v, const cplx* tab, int stride, int i, Slot slot) {
    if (R == 1) return;
    const cplx w1 = cj<INV>(tw_ld(tab + i));
    v[slot(1)] = v[slot(1)] * w1;
    if (R == 2) return;
    const cplx w2 = cj<INV>(tw_ld(tab + stride + i));
    const cplx w3 = w1 * w2;
    v[slot(2)] = v[slot(2)] * w2;
    v[slot(3)] = v[slot(3)] * w3;
    if (R == 4) return;
    const cplx w4 = cj<INV>(tw_ld(tab + 2 * stride + i));
    v[slot(4)] = v[slot(4)] * w4;
    v[slot(5)] = v[slot(5)] * (w4 * w1);
    v[slot(6)] = v[slot(6)] * (w4 * w2);
    v[slot(7)] = v[slot(7)] * (w4 * w3);
    if (R == 8) return;
    const cplx w8 = cj<INV>(tw_ld(tab + 3 * stride + i));
    v[slot(8)] = v[slot(8)] * w8;
    v[slot(9)] = v[slot(9)] * (w8 * w1);
    v[slot(10)] = v[slot(10)] * (w8 * w2);
    v[slot(11)] = v[slot(11)] * (w8 * w3);
    const cplx w12 = w8 * w4;
    v[slot(12)] = v[slot(12)] * w12;
    v[slot(13)] = v[slot(13)] * (w12 * w1);
    v[slot(14)] = v[slot(14)] * (w12 * w2);
    v[slot(15)] = v[slot(15)] * (w12 * w3);
}

// ---------------------------------------------------------------------------------------
// model adaptors: how a row is read from / written to global memory and the pointwise N(.)
// ---------------------------------------------------------------------------------------
// global accesses of the row: streaming (each element is touched once per evaluation)
RKS_HD cplx row_ld(const cplx* p) {
#if defined(__CUDA_ARCH__)
    const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
    return mk(t.x, t.y);
#else
    return *p;
#endif
}
RKS_HD void row_st(cplx* p, cplx v) {
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<double2*>(p), make_double2(v.x, v.y));
#else
    *p = v;
#endif
}

struct NlsModel {          // N = i gamma fft(|f|^2 f), f = ifft(u^)      (demos/nls.ipynb)
    const cplx* in; cplx* out; double gamma; int n; bool on;
    RKS_HD cplx load(int p) const { return row_ld(in + p); }
    RKS_HD cplx pointwise(cplx z) const {
        const double sc = 1.0 / (double)n;
        const cplx f = mk(z.x * sc, z.y * sc);
        const double f2 = f.x * f.x + f.y * f.y;
        return mk(f2 * f.x, f2 * f.y);
    }
    RKS_HD void store(int p, cplx v) const { if (on) row_st(out + p, mk(-(gamma * v.y), gamma * v.x)); }
};
struct UuxModel {          // N = -c rfft(irfft(u^) irfft(i kx u^))       (models.py:140-143)
    const cplx* in; cplx* out; const double* kx; double c; int n; bool on;
    RKS_HD cplx load(int p) const {
        const int half = n >> 1;
        if (p <= half) {
            const cplx v = row_ld(in + p);
            const double kk = kx[p];
            if (p == 0 || p == half) return mk(v.x, -(kk * v.y));       // c2r ignores Im of DC / Nyquist
            return mk(v.x - kk * v.x, v.y - kk * v.y);                  // U^ + i (i k U^)
        }
        const cplx v = row_ld(in + (n - p));
        const double kk = kx[n - p];
        return mk(v.x + kk * v.x, -(v.y + kk * v.y));                   // Hermitian partner
    }
    RKS_HD cplx pointwise(cplx z) const {
        const double sc = 1.0 / ((double)n * (double)n);
        return mk((z.x * z.y) * sc, 0.0);
    }
    RKS_HD void store(int p, cplx v) const {
        if (on && p <= (n >> 1)) row_st(out + p, mk(-c * v.x, -c * v.y));
    }
};

// ---------------------------------------------------------------------------------------
// phases.  T = thread index within the row (0 .. 32 W - 1); sm = the row's n-element smem slab;
// chunk = the 512-element slice owned by this warp; l = lane.
//
// Every inner pass is IN PLACE: a butterfly writes its 8 results to the 8 positions it read, so
// a thread only ever keeps one butterfly (8 values) in registers, and the only hazards are
// between passes (separated by __syncwarp / the row barrier).
// Inverse-direction passes read twiddle copy `ti`, forward ones copy `tf` (two identical tables at
// different addresses: otherwise the compiler keeps the inverse half's twiddles alive in local
// memory across the whole transform instead of re-reading them from L1).
// ---------------------------------------------------------------------------------------
template <int W, class Model>
RKS_HD void p0_load_outer_dif(cplx* sm, int T, const Twiddles& ti, const Model& m) {
    constexpr int TR = 32 * W, NB = 16 / W;
    if (W == 1) return;                                   // pass A reads global memory directly
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = T + TR * b;
        cplx a[W];
#pragma unroll
        for (int s = 0; s < W; ++s) a[s] = m.load(i + 512 * s);
        dftR<W, true>(a);
        twiddle_scale<W, true>(a, ti.o, 512, i, SlotPerm<W>());
        cplx* dst = sm + swz(i);                          // swz(r*512 + i) = swz(i) + r*512
#pragma unroll
        for (int r = 0; r < W; ++r) dst[r * 512] = a[perm<W>(r)];
    }
}

// inner DIF pass A: butterflies i = l, l + 32 over positions i + 64 s
template <int W, class Model>
RKS_HD void p1_dif_a(cplx* chunk, int l, const Twiddles& ti, const Model& m) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int i = l + 32 * b;
        cplx* pos = chunk + swz(i);                       // swz(i + 64 s) = swz(i) + 64 s
        cplx a[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) a[s] = (W == 1) ? m.load(i + 64 * s) : pos[64 * s];
        dft8<true>(a);
        twiddle_scale<8, true>(a, ti.a, 64, i, SlotPerm<8>());
#pragma unroll
        for (int r = 0; r < 8; ++r) pos[64 * r] = a[perm8(r)];
    }
}
// inner DIF pass B: butterflies (blk, ii) over positions blk*64 + ii + 8 s
RKS_HD void p2_dif_b(cplx* chunk, int l, const Twiddles& ti) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int q = l + 32 * b, blk = q >> 3, ii = q & 7;
        cplx* base = chunk + blk * 64;
        cplx a[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) a[s] = base[8 * s + (ii ^ s)];          // swz: low bits ^ ((p >> 3) & 7) = ii ^ s
        dft8<true>(a);
        twiddle_scale<8, true>(a, ti.b, 8, ii, SlotPerm<8>());
#pragma unroll
        for (int r = 0; r < 8; ++r) base[8 * r + (ii ^ r)] = a[perm8(r)];
    }
}
// innermost: DIF radix-8, pointwise nonlinearity, DIT radix-8 on positions q*8 + s
template <class Model>
RKS_HD void p3_core(cplx* chunk, int l, const Model& m) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int q = l + 32 * b;
        cplx* base = chunk + q * 8;
        const int x = q & 7;                                                // (p >> 3) & 7 for p = q*8 + s
        cplx a[8], c[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) a[s] = base[s ^ x];
        dft8<true>(a);
#pragma unroll
        for (int r = 0; r < 8; ++r) c[r] = m.pointwise(a[perm8(r)]);
        dft8<false>(c);
#pragma unroll
        for (int k = 0; k < 8; ++k) base[k ^ x] = c[perm8(k)];
    }
}
// inner DIT pass B': inputs blk*64 + r*8 + ii (sub-block r), outputs blk*64 + ii + 8 k: same positions
RKS_HD void p4_dit_b(cplx* chunk, int l, const Twiddles& tf) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int q = l + 32 * b, blk = q >> 3, ii = q & 7;
        cplx* base = chunk + blk * 64;
        cplx a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = base[8 * r + (ii ^ r)];
        twiddle_scale<8, false>(a, tf.b, 8, ii, SlotId());
        dft8<false>(a);
#pragma unroll
        for (int k = 0; k < 8; ++k) base[8 * k + (ii ^ k)] = a[perm8(k)];
    }
}
// inner DIT pass A': inputs r*64 + i, outputs i + 64 k.  W == 1 stores straight to global memory.
template <int W, class Model>
RKS_HD void p5_dit_a(cplx* chunk, int l, const Twiddles& tf, const Model& m) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int i = l + 32 * b;
        cplx* pos = chunk + swz(i);
        cplx a[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = pos[64 * r];
        twiddle_scale<8, false>(a, tf.a, 64, i, SlotId());
        dft8<false>(a);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (W == 1) m.store(i + 64 * k, a[perm8(k)]);
            else pos[64 * k] = a[perm8(k)];
        }
    }
}
template <int W, class Model>
RKS_HD void p6_outer_dit_store(const cplx* sm, int T, const Twiddles& tf, const Model& m) {
    constexpr int TR = 32 * W, NB = 16 / W;
    if (W == 1) return;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int i = T + TR * b;
        const cplx* src = sm + swz(i);
        cplx a[W];
#pragma unroll
        for (int r = 0; r < W; ++r) a[r] = src[r * 512];
        twiddle_scale<W, false>(a, tf.o, 512, i, SlotId());
        dftR<W, false>(a);
#pragma unroll
        for (int k = 0; k < W; ++k) m.store(i + 512 * k, a[perm<W>(k)]);
    }
}

}  // namespace fast
}  // namespace rks
