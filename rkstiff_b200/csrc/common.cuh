// Shared definitions for the rkstiff_b200 CUDA engine (sm_100a).
// Everything marked RKS_HD is also compiled for the host by tests/host_check (logic checks
// without a GPU); the product path only ever runs the device versions.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define RKS_HD __host__ __device__ __forceinline__
#define RKS_D __device__ __forceinline__
#else
#define RKS_HD inline
#define RKS_D inline
#endif

namespace rks {

// ---------------------------------------------------------------------------------------
// complex128 value type: 16 bytes, 16-byte aligned => one LDG.E.128 / STG.E.128 per element
// ---------------------------------------------------------------------------------------
struct alignas(16) cplx {
    double x, y;
};

RKS_HD cplx mk(double x, double y) { cplx r; r.x = x; r.y = y; return r; }
RKS_HD cplx operator+(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
RKS_HD cplx operator-(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
RKS_HD cplx operator-(cplx a) { return mk(-a.x, -a.y); }
RKS_HD cplx operator*(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
RKS_HD cplx operator*(double s, cplx a) { return mk(s * a.x, s * a.y); }
RKS_HD cplx operator*(cplx a, double s) { return mk(s * a.x, s * a.y); }
RKS_HD cplx operator/(cplx a, double s) { return mk(a.x / s, a.y / s); }
// x / D for a small integer constant D, correctly rounded like the division it replaces but 3 FP64 instructions
// instead of a reciprocal iteration with a slow-path call: q = RN(x * RN(1/D)), r = x - D q (exact in an FMA),
// result RN(q + r RN(1/D)) (Markstein's correction step).  Checked against `/` on 2e9 random and adversarial
// mantissas for D = 3, 6 (0 mismatches); inf/NaN pass through, results in the subnormal range may differ by
// one subnormal ulp.  The IF4/IF34 last stage holds eight such divisions per element (if4.py:120-121).
template <int D> RKS_HD double div_const(double x) {
    constexpr double y = 1.0 / (double)D;
    const double q = x * y;
    const double r = fma(-(double)D, q, x);
    const double c = fma(r, y, q);
    return fabs(q) <= 1.7976931348623157e308 ? c : q;
}
template <int D> RKS_HD cplx div_const(cplx a) { return mk(div_const<D>(a.x), div_const<D>(a.y)); }
RKS_HD cplx operator+(cplx a, double s) { return mk(a.x + s, a.y); }
RKS_HD cplx operator-(cplx a, double s) { return mk(a.x - s, a.y); }
RKS_HD cplx conj(cplx a) { return mk(a.x, -a.y); }
RKS_HD cplx mul_i(cplx a) { return mk(-a.y, a.x); }      //  i * a
RKS_HD cplx mul_mi(cplx a) { return mk(a.y, -a.x); }     // -i * a
RKS_HD double abs2(cplx a) { return a.x * a.x + a.y * a.y; }

// Smith's complex division, the algorithm NumPy's complex128 divide loop uses.
RKS_HD cplx operator/(cplx a, cplx b) {
    const double br = fabs(b.x), bi = fabs(b.y);
    if (br >= bi) {
        if (br == 0.0 && bi == 0.0) return mk(a.x / br, a.y / bi);
        const double rat = b.y / b.x, scl = 1.0 / (b.x + b.y * rat);
        return mk((a.x + a.y * rat) * scl, (a.y - a.x * rat) * scl);
    }
    const double rat = b.x / b.y, scl = 1.0 / (b.y + b.x * rat);
    return mk((a.x * rat + a.y) * scl, (a.y * rat - a.x) * scl);
}

// complex exponential exp(a+ib) = e^a (cos b + i sin b)
RKS_HD cplx cexp(cplx z) {
    const double e = exp(z.x);
    double s, c;
#if defined(__CUDA_ARCH__)
    sincos(z.y, &s, &c);
#else
    s = sin(z.y); c = cos(z.y);
#endif
    if (z.y == 0.0) return mk(e, 0.0);       // keeps real arguments exactly real (as libm cexp does)
    return mk(e * c, e * s);
}

// coefficient scalar type helpers: IF methods keep a real lin_op real (if34.py:71, if45dp.py:204)
RKS_HD double cexp_t(double z) { return exp(z); }
RKS_HD cplx cexp_t(cplx z) { return cexp(z); }
RKS_HD cplx cmul(double a, cplx b) { return mk(a * b.x, a * b.y); }
RKS_HD cplx cmul(cplx a, cplx b) { return a * b; }
RKS_HD double cmul1(double a, double b) { return a * b; }
RKS_HD cplx cmul1(cplx a, cplx b) { return a * b; }
RKS_HD double scale(double s, double a) { return s * a; }
RKS_HD cplx scale(double s, cplx a) { return mk(s * a.x, s * a.y); }

// ---------------------------------------------------------------------------------------
// method ids -- must match include/rkstiff_b200.h
// ---------------------------------------------------------------------------------------
enum : int { M_IF4 = 0, M_ETD4 = 1, M_ETD5 = 2, M_IF34 = 3, M_ETD34 = 4, M_ETD35 = 5, M_IF45DP = 6 };

RKS_HD constexpr bool method_adaptive(int m) { return m >= M_IF34; }
RKS_HD constexpr bool method_is_if(int m) { return m == M_IF4 || m == M_IF34 || m == M_IF45DP; }
RKS_HD constexpr int method_stages(int m) {
    return (m == M_IF4 || m == M_ETD4 || m == M_IF34 || m == M_ETD34) ? 4 : 6;
}
// number of N buffers; FSAL methods carry one extra (N_last = N(u+))
RKS_HD constexpr int method_nl_buffers(int m) {
    return m == M_IF4 ? 4 : m == M_ETD4 ? 4 : m == M_ETD5 ? 6 : m == M_IF34 ? 5 : m == M_ETD34 ? 5
         : m == M_ETD35 ? 6 : 7;
}
RKS_HD constexpr bool method_fsal(int m) { return m == M_IF34 || m == M_ETD34 || m == M_IF45DP; }
RKS_HD constexpr int method_ncoef(int m) {
    return (m == M_IF4 || m == M_IF34) ? 2 : (m == M_ETD4 || m == M_ETD34) ? 10
         : (m == M_ETD5 || m == M_ETD35) ? 21 : 28;
}
RKS_HD constexpr int method_q(int m) { return m == M_IF45DP ? 5 : 4; }

// coefficient array slots -------------------------------------------------------------
namespace kro { enum { E = 0, E2, a21, a31, a32, a41, a43, a51, a52, a54, COUNT }; }          // ETD4/ETD34
namespace e5 { enum { E14 = 0, E12, E34, E, a21, a31, a32, a41, a43, a51, a52, a54, a61, a62, a63, a65,
                      a71, a73, a74, a75, a76, COUNT }; }                                      // ETD5/ETD35
namespace ifc { enum { E = 0, E2, COUNT }; }                                                   // IF4/IF34
namespace dp { enum { E15 = 0, E310, E45, E89, E, a21, a31, a32, a41, a42, a43, a51, a52, a53, a54,
                      a61, a62, a63, a64, a65, a71, a73, a74, a75, r1, r3, r4, r5, COUNT }; }   // IF45DP

// ---------------------------------------------------------------------------------------
// device control block: every scalar of the adaptive loop lives here
// ---------------------------------------------------------------------------------------
enum : int { ST_RUNNING = 0, ST_DONE = 1, ST_MAX_LOOPS = 2, ST_MIN_STEP = 3 };

constexpr int LOG_CAP = 4096;

struct TrialRec {
    double h, s, t_after;
    int32_t accepted, pad;
};

struct Ctrl {
    // --- time stepping state
    double h;        // step size of the next trial
    double h_last;   // step size of the last accepted trial
    double h_coeff;  // step size the coefficient arrays hold (NaN: none)
    double t, tf;
    double s_last;
    // --- SolverConfig scalars, refreshed by rks_set_config
    double epsilon, incr_f, decr_f, safety_f, adapt_cutoff, minh, inv_q;
    // --- ETDConfig
    double modecutoff, contour_radius;
    // --- reduction scalars (contiguous: {umax2, sum_u2, sum_e2} is what multi-GPU all-reduces)
    //     red[0] doubles as the atomicMax target: its bit pattern is monotone for values >= 0
    double red[3];
    // --- counters
    long long step_count, trial_count, nl_evals, coeff_updates, store_freq;
    unsigned int ticket;             // last-block-done counter of the norm kernel
    int contour_points, r4_fix;
    int status, accept, numloops, u_sel, n_sel, need_n1, n1_refresh, step_mode;
    int log_count, snap_count, snap_pending;
    int log_cap;                     // capacity of this plan's trial-log ring
};

// ---------------------------------------------------------------------------------------
// device view of a plan, passed by value to every kernel
// ---------------------------------------------------------------------------------------
struct DevPlan {
    Ctrl* ctrl;
    TrialRec* log;
    cplx* U[2];
    cplx* K;
    cplx* ERR;
    cplx* NL[8];       // NL[1..7]
    void* coef;        // ncoef arrays of lin_elems entries (double or cplx)
    const void* lin;   // lin_op copy (double or cplx)
    double* partials;  // per-block partial sums of the norm kernel (2 per block)
    int* pre_cnt;      // stage_pre_kernel work counters: {next row, finished warps} per column block (zero between launches)
    const cplx* tw;    // twiddles exp(-2 pi i j / n), j < n (generic NL kernel)
    const cplx* twf;   // fast-path twiddle tables (fft_fast.cuh: o | a | b)
    const double* kx;  // wavenumbers of the fused model
    const cplx* norm_u; // nullptr, or the array whose magnitudes drive the error controller (diagonalize=True:
                        // the physical S u+ while the estimate stays in the eigenbasis, etd35.py:495)
    long long batch, n_c, lin_elems, n;
    double model_p0;   // c (u u_x models) or gamma (NLS)
    int method, lin_complex, lin_full, model, log2n;
    // --- coefficient storage of large grids ("lin_op shaped like u" with many modes)
    int coef_mode;     // CM_COLUMN / CM_FLAT: `coef` = ncoef arrays of lin_elems entries (set by lin_full);
                       // CM_INDEXED: `coef` = one grouped record per DISTINCT lin_op value, `cidx` maps a mode to it;
                       // CM_SEPARABLE (IF methods, lin_op = sum of per-axis terms): `coef` = per-axis exponential
                       // tables [nq][sep_ntab], coefficient = cscale[slot] * prod_d table[q][off_d + i_d]
    const int* cidx;   // CM_INDEXED: n_c indices into the table of distinct values
    const double* cscale;   // CM_SEPARABLE: the rational * h factor of every coefficient slot
    int sep_nd, sep_dims[3], sep_ntab;   // CM_SEPARABLE: spectral grid dims (last = contiguous axis), sum of them
    int nd_inplace;    // row kernel of an N-D grid model: the rows of N_j are transformed in place (kernels.cuh nl_roles)
};
enum : int { CM_COLUMN = 0, CM_FLAT = 1, CM_INDEXED = 2, CM_SEPARABLE = 3 };

// select the physical N buffer for logical index j under the FSAL role swap
RKS_HD int nl_phys(int method, int j, int n_sel) {
    if (!method_fsal(method) || !n_sel) return j;
    const int last = method_nl_buffers(method);
    return j == 1 ? last : (j == last ? 1 : j);
}

}  // namespace rks
