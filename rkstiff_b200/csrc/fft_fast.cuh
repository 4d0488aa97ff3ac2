// K4 fast path -- shared-memory FP64 FFT pair (inverse, pointwise nonlinearity, forward) for rows
// of n = 512 ... 8192 points, built from IN-PLACE radix-8/16 passes.
//
// A row of n = 512 W points is owned by W warps (32 W threads) and lives in one smem slab.
//   pass 1       radix R1 over stride Q1 = n / R1, global memory -> smem (row-level barrier after it)
//   middle       radix R_k over stride Q_k inside blocks of R_k Q_k points; every block lies inside
//                one warp's 512-point slice, so these passes need __syncwarp only
//   core         the innermost radix-R_m butterfly of the inverse transform, the pointwise
//                nonlinearity, and the innermost butterfly of the forward transform, in registers
// The inverse transform is decimation-in-frequency and the forward one its mirrored
// decimation-in-time, so the time-domain data stay digit-reversed and nothing is ever reordered:
//   n = 512  : [8] [8] core8          n = 2048 : [16] [16] core8        n = 8192 : [16] [8] [8] core8
//   n = 1024 : [16] [8] core8         n = 4096 : [16] [16] core16
// i.e. 5 in-place smem passes per nonlinear evaluation (7 for n = 8192).  Every butterfly writes
// its results to the positions it read, so a thread holds one or two butterflies in registers.
// Positions are XOR-swizzled (swz<SH>) so that the 128-bit accesses of every quarter-warp are bank
// conflict free.  All functions are __host__ __device__: tests/host_check runs the same phase
// sequence serially on the CPU.
#pragma once
#include <type_traits>
#include <utility>
#include "common.cuh"

namespace rks {
namespace fast {

constexpr double SQH = 0.70710678118654752440;   // sqrt(1/2)
constexpr double C8 = 0.92387953251128675613;    // cos(pi/8)
constexpr double S8 = 0.38268343236508977173;    // sin(pi/8)

#ifndef RKS_TW_SQUARE
#define RKS_TW_SQUARE 1        // 1: read w^1 only and square; 0: read w^1, w^2, w^4, w^8 from tables
#endif

template <int SH> RKS_HD int swz(int p) { return p ^ ((p >> SH) & 7); }

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

// Twiddle tables: for every pass k with stride Q_k > 1 the Q_k values w_{L_k}^j, j < Q_k
// (L_k = R_k Q_k, w_m = exp(-2 pi i / m)), contiguous so that the lanes of a warp read consecutive
// entries.  Only the first power is stored; twiddle_scale builds the others by squaring/products.
struct Twiddles {
    const cplx* t1;     // pass 1
    const cplx* t2;     // pass 2
    const cplx* t3;     // pass 3 (n = 8192 only)
};
// each pass table holds the powers k = 1, 2, 4, 8 as four rows of TW_S{1,2,3} entries
constexpr int TW_S1 = 512, TW_S2 = 64, TW_S3 = 16;
constexpr int TW_T1 = 0, TW_T2 = 4 * TW_S1, TW_T3 = TW_T2 + 4 * TW_S2, TW_TOTAL = TW_T3 + 4 * TW_S3;

// static plan of an n-point row
template <int N> struct Plan;
// n < 512: a warp's 512-point slab holds 512 / n consecutive rows; only the first / last pass (radix n / 64
// inside each row) differ from the 512-point plan (phase_first_packed / phase_last_packed)
template <> struct Plan<64>   { static constexpr int W = 1,  SH = 3, R1 = 1,  R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct Plan<128>  { static constexpr int W = 1,  SH = 3, R1 = 2,  R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct Plan<256>  { static constexpr int W = 1,  SH = 3, R1 = 4,  R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct Plan<512>  { static constexpr int W = 1,  SH = 3, R1 = 8,  R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct Plan<1024> { static constexpr int W = 2,  SH = 3, R1 = 16, R2 = 8,  R3 = 8,  R4 = 1; };
template <> struct Plan<2048> { static constexpr int W = 4,  SH = 3, R1 = 16, R2 = 16, R3 = 8,  R4 = 1; };
template <> struct Plan<4096> { static constexpr int W = 8,  SH = 4, R1 = 16, R2 = 16, R3 = 16, R4 = 1; };
template <> struct Plan<8192> { static constexpr int W = 16, SH = 3, R1 = 16, R2 = 8,  R3 = 8,  R4 = 8; };

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

// entry `idx` of the concatenated table block [t1 | t2 | t3] for an n-point row
template <int N>
RKS_HD cplx twiddle_table_entry_n(int idx) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, Q2 = Q1 / P::R2, Q3 = Q2 / P::R3;
    int j, m, k, q;
    if (idx < TW_T2) { k = idx / TW_S1; j = idx % TW_S1; m = N; q = Q1; }
    else if (idx < TW_T3) { k = (idx - TW_T2) / TW_S2; j = (idx - TW_T2) % TW_S2; m = Q1; q = Q2; }
    else { k = (idx - TW_T3) / TW_S3; j = (idx - TW_T3) % TW_S3; m = Q2; q = Q3; }
    if (j >= q) j = 0;
    j = (int)(((long long)j << k) % m);              // w^(2^k j)
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(-2.0 * (double)j / (double)m, &s, &c);
#else
    s = sin(-2.0 * M_PI * (double)j / (double)m); c = cos(-2.0 * M_PI * (double)j / (double)m);
#endif
    return mk(c, s);
}
RKS_HD cplx twiddle_table_entry(int idx, int n) {
    switch (n) {
        case 64: return twiddle_table_entry_n<64>(idx);
        case 128: return twiddle_table_entry_n<128>(idx);
        case 256: return twiddle_table_entry_n<256>(idx);
        case 512: return twiddle_table_entry_n<512>(idx);
        case 1024: return twiddle_table_entry_n<1024>(idx);
        case 2048: return twiddle_table_entry_n<2048>(idx);
        case 4096: return twiddle_table_entry_n<4096>(idx);
        default: return twiddle_table_entry_n<8192>(idx);
    }
}

// v[slot(r)] *= w^r for r = 1..R-1.  RKS_TW_SQUARE: only w^1 is read and w^2, w^4, w^8 are
// squarings (error <= ~8 ulp on the twiddle); otherwise the four powers are read from the table
// rows.  The remaining powers are products of at most three of them.  Twiddle loads compete with
// the shared-memory passes for the LSU pipe.
template <int R, bool INV, class Slot>
RKS_HD void twiddle_scale(cplx* v, const cplx* tab, int stride, int i, Slot slot) {
    if (R == 1) return;
    const cplx w1 = cj<INV>(tw_ld(tab + i));
    v[slot(1)] = v[slot(1)] * w1;
    if (R == 2) return;
    const cplx w2 = RKS_TW_SQUARE ? w1 * w1 : cj<INV>(tw_ld(tab + stride + i));
    const cplx w3 = w1 * w2;
    v[slot(2)] = v[slot(2)] * w2;
    v[slot(3)] = v[slot(3)] * w3;
    if (R == 4) return;
    const cplx w4 = RKS_TW_SQUARE ? w2 * w2 : cj<INV>(tw_ld(tab + 2 * stride + i));
    v[slot(4)] = v[slot(4)] * w4;
    v[slot(5)] = v[slot(5)] * (w4 * w1);
    v[slot(6)] = v[slot(6)] * (w4 * w2);
    v[slot(7)] = v[slot(7)] * (w4 * w3);
    if (R == 8) return;
    const cplx w8 = RKS_TW_SQUARE ? w4 * w4 : cj<INV>(tw_ld(tab + 3 * stride + i));
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

// Where a row's input comes from: a plain array, or partly the TMA staging buffer (StagedRow).
struct ArraySource {
    const cplx* in;
    RKS_HD cplx value(long long p) const { return row_ld(in + p); }
};
// n = 8192 rows fill 128 KB of shared memory, so only one row is resident per SM and its HBM read
// cannot overlap another row's arithmetic.  The head of the NEXT row (`nst` elements; all of a half
// spectrum) is therefore copied by the TMA engine (cp.async.bulk, kernels.cuh) into the spare shared
// memory while the current row is transformed; the tail still comes from global (L2-prefetched).
struct StagedRow {
    const cplx* in; const cplx* stg; int nst;
    RKS_HD cplx value(long long p) const { return p < nst ? stg[p] : row_ld(in + p); }
    RKS_HD cplx get(int k) const { return k < nst ? stg[k] : row_ld(in + k); }
};
// pre-transformed rows: the first SL points of every warp's 512-point slice are staged (kernels.cuh stage_issue_sliced)
struct SlicedStagedRow {
    static constexpr int SL = 384;         // 16 x 384 = the 6144 points of the staging buffer (uneven shares -- more for the
                                           // warps that start later, or earlier -- measured slower: profiles/r02ac_*)
    const cplx* in; const cplx* stg;
    RKS_HD cplx value(long long p) const {
        const int o = (int)p & 511;
        return o < SL ? stg[((int)p >> 9) * SL + o] : row_ld(in + p);
    }
};
template <class Src>
struct NlsModelT {          // N = i gamma fft(|f|^2 f), f = ifft(u^)      (demos/nls.ipynb)
    Src src; cplx* out; double gamma; int n; bool on;
    RKS_HD cplx load(int p) const { return src.value(p); }
    RKS_HD cplx pointwise(cplx z) const {
        const double sc = 1.0 / (double)n;
        const cplx f = mk(z.x * sc, z.y * sc);
        const double f2 = f.x * f.x + f.y * f.y;
        return mk(f2 * f.x, f2 * f.y);
    }
    RKS_HD void store(int p, cplx v) const { if (on) row_st(out + p, mk(-(gamma * v.y), gamma * v.x)); }
};
using NlsModel = NlsModelT<ArraySource>;

// N = -c rfft(irfft(u^) irfft(i kx u^))  (models.py:140-143).  `Half` yields the half spectrum value at
// index k <= n/2: a global array or the TMA staging row.
struct GlobalHalf {
    const cplx* in;
    RKS_HD cplx get(int k) const { return row_ld(in + k); }
};
template <class Half>
struct UuxModelT {
    Half half; cplx* out; const double* kx; double c; int n; bool on;
    RKS_HD cplx load(int p) const {
        const int hn = n >> 1;
        if (p <= hn) {
            const cplx v = half.get(p);
            const double kk = kx[p];
            if (p == 0 || p == hn) return mk(v.x, -(kk * v.y));         // c2r ignores Im of DC / Nyquist
            return mk(v.x - kk * v.x, v.y - kk * v.y);                  // U^ + i (i k U^)
        }
        const cplx v = half.get(n - p);
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
using UuxModel = UuxModelT<GlobalHalf>;

// N = c rfft(irfft(u^)^3): the cubic term of the Fourier-diagonal Allen-Cahn equation
// u_t = eps u_xx + u - u^3 (L = 1 - eps k^2 carries the +u, mirroring rkstiff/models.py:240-244; c = -1).
template <class Half>
struct CubicModelT {
    Half half; cplx* out; double c; int n; bool on;
    RKS_HD cplx load(int p) const {
        const int hn = n >> 1;
        if (p <= hn) {
            const cplx v = half.get(p);
            if (p == 0 || p == hn) return mk(v.x, 0.0);                 // c2r ignores Im of DC / Nyquist
            return v;
        }
        return conj(half.get(n - p));                                   // Hermitian partner
    }
    RKS_HD cplx pointwise(cplx z) const {
        const double x = z.x * (1.0 / (double)n);
        return mk(x * x * x, 0.0);
    }
    RKS_HD void store(int p, cplx v) const {
        if (on && p <= (n >> 1)) row_st(out + p, mk(c * v.x, c * v.y));
    }
};
using CubicModel = CubicModelT<GlobalHalf>;

// Sine-Gordon phi_tt = phi_xx - sin(phi) in the first-order complex form psi = phi_t + i Omega phi,
// Omega = sqrt(1 + k^2):  psi^_t = i Omega psi^ + F{phi - sin phi},
// phi^(k) = (psi^(k) - conj(psi^(-k))) / (2 i Omega(k))   (`omega` holds Omega(k)).
struct SineGordonModel {
    const cplx* in; cplx* out; const double* omega; int n; bool on;
    RKS_HD cplx load(int p) const {
        const cplx a = row_ld(in + p), b = row_ld(in + ((n - p) & (n - 1)));
        const double scl = 1.0 / (2.0 * omega[p]);
        return mk((a.y + b.y) * scl, -((a.x - b.x) * scl));             // (a - conj b) / (2 i Omega)
    }
    RKS_HD cplx pointwise(cplx z) const {
        const double phi = z.x * (1.0 / (double)n);
        return mk(phi - sin(phi), 0.0);
    }
    RKS_HD void store(int p, cplx v) const { if (on) row_st(out + p, v); }
};

// ---- spectral derivative rows (rkstiff/derivatives.py:47-179): forward transform, (i kx)^m, inverse transform ----
// The K4 pipeline is inverse-DIF, pointwise, forward-DIT.  On conjugated data it computes the mirrored pair,
// fft(x) = conj(ifft_unnormalised(conj x)): the row is loaded conjugated, the "inverse" passes leave conj(X) in
// digit-reversed order, the pointwise step multiplies position p by mult[k(p)] = conj((i kx[k])^m) / n where k(p) is
// the frequency the position holds (DigitMap), the "forward" passes return conj(ifft(M X)), and the store conjugates
// again.  A model that needs the position offers pointwise_at(z, p) next to pointwise(z).
struct DigitMap {
    // in-place DIF with radices R1, R2, R3 (, R4): position digits d1 (top bits), d2, d3, d4 hold frequency
    // k = d1 + R1 (d2 + R2 (d3 + R3 d4)).  l1 < 0: the generic kernel's chain of radix-4 passes (+ one radix-2).
    int L, l1, l2, l3;
    RKS_HD int freq(int pos) const {
        if (l1 < 0) {
            int k = 0, sh = 0, rem = L;
            while (rem >= 2) { k |= ((pos >> (rem - 2)) & 3) << sh; sh += 2; rem -= 2; }
            if (rem) k |= (pos & 1) << sh;
            return k;
        }
        const int s1 = L - l1, s2 = s1 - l2, s3 = s2 - l3;
        const int d1 = pos >> s1, d2 = (pos >> s2) & ((1 << l2) - 1), d3 = (pos >> s3) & ((1 << l3) - 1);
        const int d4 = pos & ((1 << s3) - 1);
        return d1 | (d2 << l1) | (d3 << (l1 + l2)) | (d4 << (l1 + l2 + l3));
    }
};
RKS_HD constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
template <int N> RKS_HD DigitMap digit_map_n() {
    using P = Plan<N>;
    // packed rows (N < 512): radix N / 64 over stride 64 inside the row, then 8, 8
    return DigitMap{ilog2(N), N < 512 ? ilog2(N / 64) : ilog2(P::R1), ilog2(P::R2), ilog2(P::R3)};
}
RKS_HD DigitMap digit_map(int n, bool generic) {
    if (generic) return DigitMap{ilog2(n), -1, 0, 0};
    switch (n) {
        case 64: return digit_map_n<64>();
        case 128: return digit_map_n<128>();
        case 256: return digit_map_n<256>();
        case 512: return digit_map_n<512>();
        case 1024: return digit_map_n<1024>();
        case 2048: return digit_map_n<2048>();
        case 4096: return digit_map_n<4096>();
        default: return digit_map_n<8192>();
    }
}
template <class M, class = void> struct has_pos : std::false_type {};
template <class M>
struct has_pos<M, std::void_t<decltype(std::declval<const M&>().pointwise_at(std::declval<cplx>(), 0))>> : std::true_type {};
template <class M> RKS_HD cplx pointwise_of(const M& m, cplx z, int pos) {
    if constexpr (has_pos<M>::value) return m.pointwise_at(z, pos);
    else return m.pointwise(z);
}
struct DerivModel {          // complex rows (dx_fft, derivatives.py:126-179): out = ifft((i kx)^m fft(in))
    const cplx* in; cplx* out; const cplx* mult; DigitMap map; bool on;
    RKS_HD cplx load(int p) const { return conj(row_ld(in + p)); }
    RKS_HD cplx pointwise(cplx z) const { return z; }
    RKS_HD cplx pointwise_at(cplx z, int pos) const { return z * tw_ld(mult + map.freq(pos)); }
    RKS_HD void store(int p, cplx v) const { if (on) row_st(out + p, conj(v)); }
};
// real rows (dx_rfft, derivatives.py:47-123), two per complex transform: z = a + i b, and since the multiplier is
// Hermitian (M[n-k] = conj M[k], real at the Nyquist frequency: what irfft keeps) the result is a' + i b'.
// `in` / `out` are float64 arrays of n-point rows; row pair q = rows 2q, 2q+1 = 2n consecutive doubles.
struct DerivPairModel {
    const double* in; double* out; const cplx* mult; DigitMap map; int n; bool on;
    RKS_HD cplx load(int p) const { return mk(in[p], -in[n + p]); }
    RKS_HD cplx pointwise(cplx z) const { return z; }
    RKS_HD cplx pointwise_at(cplx z, int pos) const { return z * tw_ld(mult + map.freq(pos)); }
    RKS_HD void store(int p, cplx v) const { if (on) { out[p] = v.x; out[n + p] = -v.y; } }
};

// model ids of include/rkstiff_b200.h -> model objects reading/writing plain arrays
// (make_staged: input row partly staged in shared memory, models 1-3)
template <int MODEL> struct ModelOf;
template <> struct ModelOf<1> {
    using type = UuxModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double* kx, double p0, int n, bool on) {
        return type{GlobalHalf{in}, out, kx, p0, n, on};
    }
    RKS_HD static UuxModelT<StagedRow> make_staged(StagedRow s, cplx* out, const double* kx, double p0, int n, bool on) {
        return UuxModelT<StagedRow>{s, out, kx, p0, n, on};
    }
};
template <> struct ModelOf<2> {
    using type = NlsModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double*, double p0, int n, bool on) {
        return type{ArraySource{in}, out, p0, n, on};
    }
    RKS_HD static NlsModelT<StagedRow> make_staged(StagedRow s, cplx* out, const double*, double p0, int n, bool on) {
        return NlsModelT<StagedRow>{s, out, p0, n, on};
    }
};
template <> struct ModelOf<3> {
    using type = CubicModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double*, double p0, int n, bool on) {
        return type{GlobalHalf{in}, out, p0, n, on};
    }
    RKS_HD static CubicModelT<StagedRow> make_staged(StagedRow s, cplx* out, const double*, double p0, int n, bool on) {
        return CubicModelT<StagedRow>{s, out, p0, n, on};
    }
};
template <> struct ModelOf<4> {
    using type = SineGordonModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double* kx, double, int n, bool on) {
        return type{in, out, kx, n, on};
    }
};
// 5, 6: spectral derivatives; `kx` points at the complex multiplier table, p0 != 0 selects the generic kernel's digit order
template <> struct ModelOf<5> {
    using type = DerivModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double* kx, double p0, int n, bool on) {
        return type{in, out, reinterpret_cast<const cplx*>(kx), digit_map(n, p0 != 0.0), on};
    }
};
template <> struct ModelOf<6> {
    using type = DerivPairModel;
    RKS_HD static type make(const cplx* in, cplx* out, const double* kx, double p0, int n, bool on) {
        return type{reinterpret_cast<const double*>(in), reinterpret_cast<double*>(out), reinterpret_cast<const cplx*>(kx),
                    digit_map(n, p0 != 0.0), n, on};
    }
};

// 512 / N consecutive rows of an N-point model packed into one 512-point slab (N = 64, 128, 256):
// slab position p belongs to row p / N.  Rows past the end of the batch read zeros and store nothing.
template <int MODEL, int N>
struct PackedModel {
    const cplx* in0; cplx* out0; const double* kx; double p0; long long n_c; int nvalid;
    RKS_HD typename ModelOf<MODEL>::type row(int sub) const {
        return ModelOf<MODEL>::make(in0 + sub * n_c, out0 + sub * n_c, kx, p0, N, sub < nvalid);
    }
    RKS_HD cplx load(int p) const {
        const int sub = p / N;
        return sub < nvalid ? row(sub).load(p % N) : mk(0.0, 0.0);
    }
    RKS_HD cplx pointwise(cplx z) const { return row(0).pointwise(z); }
    template <class M = typename ModelOf<MODEL>::type, class = std::enable_if_t<has_pos<M>::value>>
    RKS_HD cplx pointwise_at(cplx z, int p) const { return row(0).pointwise_at(z, p % N); }
    RKS_HD void store(int p, cplx v) const {
        const int sub = p / N;
        if (sub < nvalid) row(sub).store(p % N, v);
    }
};

// ---------------------------------------------------------------------------------------
// generic in-place passes.  A butterfly is identified by the logical position p0 of its first
// element (elements p0 + Q s) and its twiddle index j (w_L^(r j), L = R Q).  NB butterflies of one
// thread are loaded first and then transformed/stored one after the other.
// Inverse-direction passes read twiddle copy `ti`, forward ones copy `tf` (two identical tables at
// different addresses: otherwise the compiler keeps the inverse half's twiddles alive in local
// memory across the whole transform instead of re-reading them from L1).
// ---------------------------------------------------------------------------------------
// ---- butterfly-level primitives (one butterfly = R values `a` of one thread) ----
template <int R, int Q, int SH>
RKS_HD void bf_load(const cplx* sm, int p0, cplx* a) {
#pragma unroll
    for (int s = 0; s < R; ++s) a[s] = sm[swz<SH>(p0 + Q * s)];
}
template <int R, int Q, class Model>
RKS_HD void bf_load_global(const Model& m, int p0, cplx* a) {
#pragma unroll
    for (int s = 0; s < R; ++s) a[s] = m.load(p0 + Q * s);
}
// results of bf_dif / bf_dit / bf_core sit in slot perm<R>(r); both stores undo that
template <int R, int Q, int SH>
RKS_HD void bf_store(cplx* sm, int p0, const cplx* a) {
#pragma unroll
    for (int r = 0; r < R; ++r) sm[swz<SH>(p0 + Q * r)] = a[perm<R>(r)];
}
template <int R, int Q, class Model>
RKS_HD void bf_store_global(const Model& m, int p0, const cplx* a) {
#pragma unroll
    for (int k = 0; k < R; ++k) m.store(p0 + Q * k, a[perm<R>(k)]);
}
template <int R, int Q, int TS>
RKS_HD void bf_dif(cplx* a, const cplx* tab, int j) {          // inverse butterfly, then twiddles on the outputs
    dftR<R, true>(a);
    if (Q > 1) twiddle_scale<R, true>(a, tab, TS, j, SlotPerm<R>());
}
template <int R, int Q, int TS>
RKS_HD void bf_dit(cplx* a, const cplx* tab, int j) {          // twiddles on the inputs, then forward butterfly
    if (Q > 1) twiddle_scale<R, false>(a, tab, TS, j, SlotId());
    dftR<R, false>(a);
}
template <int R, class Model>
RKS_HD void bf_core(cplx* a, const Model& m, int p0) {          // innermost inverse bf, N(.), innermost forward bf
    cplx c[R];
    dftR<R, true>(a);
#pragma unroll
    for (int r = 0; r < R; ++r) c[r] = pointwise_of(m, a[perm<R>(r)], p0 + r);
    dftR<R, false>(c);
#pragma unroll
    for (int r = 0; r < R; ++r) a[r] = c[r];
}

// ---- whole passes: NB butterflies per thread, one after the other ----
template <int R, int Q, int SH, int NB, int TS, bool GLOBAL_IN, class Model>
RKS_HD void dif_pass(cplx* sm, const int (&p0)[NB], const int (&j)[NB], const cplx* tab, const Model& m) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        cplx a[R];
        if (GLOBAL_IN) bf_load_global<R, Q>(m, p0[b], a);
        else bf_load<R, Q, SH>(sm, p0[b], a);
        bf_dif<R, Q, TS>(a, tab, j[b]);
        bf_store<R, Q, SH>(sm, p0[b], a);
    }
}
template <int R, int Q, int SH, int NB, int TS, bool GLOBAL_OUT, class Model>
RKS_HD void dit_pass(cplx* sm, const int (&p0)[NB], const int (&j)[NB], const cplx* tab, const Model& m) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        cplx a[R];
        bf_load<R, Q, SH>(sm, p0[b], a);
        bf_dit<R, Q, TS>(a, tab, j[b]);
        if (GLOBAL_OUT) bf_store_global<R, Q>(m, p0[b], a);
        else bf_store<R, Q, SH>(sm, p0[b], a);
    }
}
// innermost butterflies + pointwise nonlinearity (stride 1, no twiddles)
template <int R, int SH, int NB, class Model>
RKS_HD void core_pass(cplx* sm, const int (&p0)[NB], const Model& m) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        cplx a[R];
        bf_load<R, 1, SH>(sm, p0[b], a);
        bf_core<R>(a, m, p0[b]);
        bf_store<R, 1, SH>(sm, p0[b], a);
    }
}

// butterfly assignment of a warp-local pass: the warp's 512-point slice holds 512/R butterflies,
// lane l takes u = l + 32 c;  u -> (block u / Q, offset u % Q)
template <int R, int Q, int NB>
RKS_HD void warp_butterflies(int chunk0, int l, int (&p0)[NB], int (&j)[NB]) {
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const int u = l + 32 * c;
        j[c] = u % Q;
        p0[c] = chunk0 + (u / Q) * (R * Q) + j[c];
    }
}

// ---- the phase sequence of one row (T = thread index within the row, W = warps per row) ----
// phase ids for the host emulation: 0 first DIF pass, 1..: middle DIF, core, middle DIT, last DIT
template <int N, class Model>
RKS_HD void phase_first(cplx* sm, int T, const Twiddles& ti, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, NB = Q1 / (32 * P::W);
    int p0[NB], j[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) { p0[c] = T + 32 * P::W * c; j[c] = p0[c]; }
    dif_pass<P::R1, Q1, P::SH, NB, TW_S1, true>(sm, p0, j, ti.t1, m);
}
template <int N, class Model>
RKS_HD void phase_last(cplx* sm, int T, const Twiddles& tf, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, NB = Q1 / (32 * P::W);
    int p0[NB], j[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) { p0[c] = T + 32 * P::W * c; j[c] = p0[c]; }
    dit_pass<P::R1, Q1, P::SH, NB, TW_S1, true>(sm, p0, j, tf.t1, m);
}
// middle pass k = 2 (and 3 for n = 8192); DIR: true = inverse/DIF, false = forward/DIT
template <int N, int K, bool DIF, class Model>
RKS_HD void phase_middle(cplx* sm, int T, const Twiddles& tw, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1;
    constexpr int R = K == 2 ? P::R2 : P::R3;
    constexpr int Q = K == 2 ? Q1 / P::R2 : Q1 / P::R2 / P::R3;
    constexpr int NB = 16 / R;
    int p0[NB], j[NB];
    warp_butterflies<R, Q, NB>((T >> 5) * 512, T & 31, p0, j);
    const cplx* tab = K == 2 ? tw.t2 : tw.t3;
    constexpr int TS = K == 2 ? TW_S2 : TW_S3;
    if (DIF) dif_pass<R, Q, P::SH, NB, TS, false>(sm, p0, j, tab, m);
    else dit_pass<R, Q, P::SH, NB, TS, false>(sm, p0, j, tab, m);
}
// ---- pre-transformed rows (complex-field models) ----
// The stage kernel K1 is HBM bound and leaves the FP64 pipe idle, K4 is FP64/LSU bound and leaves HBM idle.
// For an intermediate stage value k (consumed by N(.) only) K1 therefore applies the FIRST inverse pass --
// the radix-R1 butterflies over stride Q1 = n / R1 with their twiddles -- while the 16 values of a butterfly
// are still in its registers (pre_butterfly) and stores the result at the logical positions phase_first
// would have written.  K4 then starts at the first warp-local pass, reading global memory (phase_pre):
// one shared-memory round trip, the largest twiddle pass and R1's butterflies leave the bound kernel.
// The arithmetic is the same operations in the same order: both routes agree to the last bit.
template <int R1>
RKS_HD void pre_butterfly(cplx* a, const cplx* t1, int j) {     // a[s] = k[j + Q1 s] -> a[perm(r)] = row[j + Q1 r]
    dftR<R1, true>(a);
    twiddle_scale<R1, true>(a, t1, TW_S1, j, SlotPerm<R1>());
}
// first pass of K4 on a pre-transformed row: middle pass 2 of the inverse transform, input from the model
template <int N, class Model>
RKS_HD void phase_pre(cplx* sm, int T, const Twiddles& ti, const Model& m) {
    using P = Plan<N>;
    constexpr int Q1 = N / P::R1, R = P::R2, Q = Q1 / P::R2, NB = 16 / R;
    int p0[NB], j[NB];
    warp_butterflies<R, Q, NB>((T >> 5) * 512, T & 31, p0, j);
    dif_pass<R, Q, P::SH, NB, TW_S2, true>(sm, p0, j, ti.t2, m);
}
template <int N, class Model>
RKS_HD void phase_core(cplx* sm, int T, const Model& m) {
    using P = Plan<N>;
    constexpr int R = P::R4 > 1 ? P::R4 : P::R3;
    constexpr int NB = 16 / R;
    int p0[NB], j[NB];
    warp_butterflies<R, 1, NB>((T >> 5) * 512, T & 31, p0, j);
    core_pass<R, P::SH, NB>(sm, p0, m);
}
// first / last pass of a packed slab (N < 512): radix N / 64 over stride 64 inside every row; lane l takes
// the butterflies u = l + 32 c of the slab, u -> (row u / 64, offset u % 64)
template <int N, class Model>
RKS_HD void phase_first_packed(cplx* sm, int T, const Twiddles& ti, const Model& m) {
    using P = Plan<N>;
    constexpr int NB = (512 / P::R1) / 32;
    int p0[NB], j[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) { const int u = T + 32 * c; j[c] = u & 63; p0[c] = (u >> 6) * N + j[c]; }
    dif_pass<P::R1, 64, P::SH, NB, TW_S1, true>(sm, p0, j, ti.t1, m);
}
template <int N, class Model>
RKS_HD void phase_last_packed(cplx* sm, int T, const Twiddles& tf, const Model& m) {
    using P = Plan<N>;
    constexpr int NB = (512 / P::R1) / 32;
    int p0[NB], j[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) { const int u = T + 32 * c; j[c] = u & 63; p0[c] = (u >> 6) * N + j[c]; }
    dit_pass<P::R1, 64, P::SH, NB, TW_S1, true>(sm, p0, j, tf.t1, m);
}
// number of warp-local middle passes per direction
template <int N> RKS_HD constexpr int middle_passes() { return Plan<N>::R4 > 1 ? 2 : 1; }

}  // namespace fast
}  // namespace rks
