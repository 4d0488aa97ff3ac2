// Device kernels of the rkstiff_b200 engine (sm_100a).  See DESIGN.md for the byte model of each.
#pragma once
#include <type_traits>
#include "common.cuh"
#include "coeffs.cuh"
#include "stages.cuh"
#include "errctl.cuh"
#include "fft.cuh"
#include "fft_fast.cuh"
#include "fft_real.cuh"
#include "fft_pair.cuh"
#include "fft_axis.cuh"

namespace rks {

// 128-bit global accesses (LDG.E.128 / STG.E.128)
RKS_D cplx ldg(const cplx* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return mk(v.x, v.y);
}
RKS_D cplx ldcs(const cplx* p) {     // streaming load: state arrays are touched once per kernel
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    return mk(v.x, v.y);
}
RKS_D cplx ld_plain(const cplx* p) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    return mk(v.x, v.y);
}
RKS_D void stg(cplx* p, cplx v) { *reinterpret_cast<double2*>(p) = make_double2(v.x, v.y); }
RKS_D double ldcoef(const double* p) { return __ldg(p); }
RKS_D cplx ldcoef(const cplx* p) { return ldg(p); }

RKS_D unsigned long long nonneg_bits(double v) {
    // bit pattern that orders like the value for v >= 0; NaN maps above +inf
    if (v != v) return 0x7ff8000000000000ull;
    return (unsigned long long)__double_as_longlong(v);
}

// ---------------------------------------------------------------------------------------
// control-block helpers
// ---------------------------------------------------------------------------------------
struct BeginArgs {
    double t0, tf, h;
    long long store_freq;
    int step_mode, keep_fsal, n1_refresh;
};

RKS_D void begin_ctrl(Ctrl* c, const BeginArgs& a);
__global__ void begin_kernel(Ctrl* c, BeginArgs a) { begin_ctrl(c, a); }
__global__ void begin_multi_kernel(const DevPlan* plans, int nplans, BeginArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nplans) begin_ctrl(plans[i].ctrl, a);
}
RKS_D void begin_ctrl(Ctrl* c, const BeginArgs& a) {
    c->t = a.t0; c->tf = a.tf; c->h = a.h; c->h_last = a.h;
    c->store_freq = a.store_freq;
    c->step_mode = a.step_mode;
    c->n1_refresh = a.n1_refresh;
    c->status = ST_RUNNING;
    c->numloops = 0;
    c->red[0] = c->red[1] = c->red[2] = 0.0;
    c->ticket = 0u;
    c->snap_pending = 0;
    if (!a.keep_fsal) {
        c->h_coeff = __longlong_as_double(0x7ff8000000000000ll);   // NaN: never equal
        c->accept = 0; c->need_n1 = 1; c->n_sel = 0;
        c->step_count = 0; c->trial_count = 0; c->nl_evals = 0; c->coeff_updates = 0;
        c->log_count = 0; c->snap_count = 0; c->s_last = 0.0;
    }
}

__global__ void set_h_kernel(Ctrl* c, double h) {
    c->h = h;
    c->status = ST_RUNNING;
    c->numloops = 0;
}
__global__ void set_h_multi_kernel(const DevPlan* plans, int nplans, double h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nplans) return;
    Ctrl* c = plans[i].ctrl;
    c->h = h; c->status = ST_RUNNING; c->numloops = 0;
}

struct CfgArgs {
    double epsilon, incr_f, decr_f, safety_f, adapt_cutoff, minh, inv_q, modecutoff, contour_radius;
    int contour_points, r4_fix;
};
RKS_D void set_config_ctrl(Ctrl* c, const CfgArgs& a);
__global__ void set_config_kernel(Ctrl* c, CfgArgs a) { set_config_ctrl(c, a); }
__global__ void set_config_multi_kernel(const DevPlan* plans, int nplans, CfgArgs a, int log_cap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nplans) return;
    set_config_ctrl(plans[i].ctrl, a);
    plans[i].ctrl->log_cap = log_cap;
}
// how many plans are still stepping (host polls this once per chunk of trials)
__global__ void count_running_kernel(const DevPlan* plans, int nplans, int* out) {
    // out[0] = rows still RUNNING, out[1] = worst status code of any row
    int mine = 0, worst = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nplans; i += gridDim.x * blockDim.x) {
        const int st = plans[i].ctrl->status;
        mine += st == ST_RUNNING;
        worst = st > worst ? st : worst;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mine += __shfl_xor_sync(0xffffffffu, mine, o);
        const int w = __shfl_xor_sync(0xffffffffu, worst, o);
        worst = w > worst ? w : worst;
    }
    if ((threadIdx.x & 31) == 0) {
        if (mine) atomicAdd(out, mine);
        if (worst) atomicMax(out + 1, worst);
    }
}
RKS_D void set_config_ctrl(Ctrl* c, const CfgArgs& a) {
    c->log_cap = LOG_CAP;
    c->epsilon = a.epsilon; c->incr_f = a.incr_f; c->decr_f = a.decr_f; c->safety_f = a.safety_f;
    c->adapt_cutoff = a.adapt_cutoff; c->minh = a.minh; c->inv_q = a.inv_q;
    c->modecutoff = a.modecutoff; c->contour_radius = a.contour_radius;
    c->contour_points = a.contour_points; c->r4_fix = a.r4_fix;
}

__global__ void fast_twiddle_kernel(cplx* twf, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 2 * fast::TW_TOTAL) twf[idx] = fast::twiddle_table_entry(idx % fast::TW_TOTAL, n);   // two copies
}

__global__ void twiddle_kernel(cplx* tw, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s, c;
    sincospi(-2.0 * (double)j / (double)n, &s, &c);
    tw[j] = mk(c, s);
}

// ---------------------------------------------------------------------------------------
// K2: coefficient kernel.  One lane per mode; modes below the cutoff are finished by the
// whole warp, one contour node per lane (M = 32 nodes == one warp by default).
// FAM: 0 = IF4/IF34 (E, E2), 1 = Krogstad ETD4/ETD34, 2 = ETD5/ETD35, 3 = IF45DP
// ---------------------------------------------------------------------------------------
template <typename T> RKS_D T load_lin(const void* lin, long long i);
template <> RKS_D double load_lin<double>(const void* lin, long long i) { return __ldg((const double*)lin + i); }
template <> RKS_D cplx load_lin<cplx>(const void* lin, long long i) { return ldg((const cplx*)lin + i); }

RKS_D cplx warp_sum(cplx v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

// where a mode's coefficients go: ncoef arrays of lin_elems entries (GROUPED = false), or one grouped record
// per distinct lin_op value (CM_INDEXED, stages.cuh group_*)
template <int M, typename CT, bool GROUPED>
RKS_D void emit_coefs(const DevPlan& p, long long i, const CT* arr) {
    constexpr int NC = method_ncoef(M), S = method_stages(M);
    CT* out = (CT*)p.coef;
    if (!GROUPED) {
#pragma unroll
        for (int s = 0; s < NC; ++s) out[s * p.lin_elems + i] = arr[s];
        return;
    }
    CT* rec = out + i * record_elems<CT>(M);
#pragma unroll
    for (int g = 1; g <= S + 1; ++g) {
#pragma unroll
        for (int s = 0; s < NC; ++s)
            if (group_mask(M, g) & (1u << s)) rec[group_off<CT>(M, g) + group_pos(group_mask(M, g), s)] = arr[s];
    }
}

template <int M, typename LT, bool GROUPED>
RKS_D void coef_kernel_body(const DevPlan& p, int force) {
    constexpr int FAM = (M == M_IF4 || M == M_IF34) ? 0 : (M == M_ETD4 || M == M_ETD34) ? 1 : (M == M_ETD5 || M == M_ETD35) ? 2 : 3;
    Ctrl* c = p.ctrl;
    const double h = c->h;
    if (!force) {
        if (c->status != ST_RUNNING) return;
        if (h == c->h_coeff) return;                      // etd35.py:851: exact float equality
    }
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < p.lin_elems;
    if (i == 0) atomicAdd((unsigned long long*)&c->coeff_updates, 1ull);

    if (FAM == 0) {
        if (!valid) return;
        LT arr[ifc::COUNT];
        const LT z = scale(h, load_lin<LT>(p.lin, i));
        arr[ifc::E] = cexp_t(z);
        arr[ifc::E2] = cexp_t(z / 2.0);
        emit_coefs<M, LT, GROUPED>(p, i, arr);
        return;
    }
    if (FAM == 3) {
        if (!valid) return;
        LT arr[dp::COUNT];
        tableau_if45dp<LT>(scale(h, load_lin<LT>(p.lin, i)), h, c->r4_fix, arr);
        emit_coefs<M, LT, GROUPED>(p, i, arr);
        return;
    }
    if (FAM == 1 || FAM == 2) {
        // ETD strategies cast lin_op to complex128 (etd35.py:124)
        constexpr bool FIVE = (FAM == 2);
        cplx z = mk(0.0, 0.0);
        if (valid) {
            if (p.lin_complex) z = scale(h, ldg((const cplx*)p.lin + i));
            else z = mk(h * __ldg((const double*)p.lin + i), 0.0);
        }
        const bool small = valid && (hypot(z.x, z.y) < c->modecutoff);    // np.abs(z) < modecutoff
        PsiSet ps = psi_zero();
        const ExpSet ez = exp_set<FIVE>(z);               // exp(z/4), exp(z/2), exp(3z/4), exp(z): one cexp
        if (valid && !small) {
            psi_accumulate<FIVE>(ps, z, ez);
            psi_scale(ps, h, 1.0);
        }
        // contour mean for the small modes of this warp (etd35.py:239-269)
        unsigned todo = __ballot_sync(0xffffffffu, small);
        const int lane = threadIdx.x & 31;
        const int m = c->contour_points;
        const double radius = c->contour_radius;
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            cplx zs;
            zs.x = __shfl_sync(0xffffffffu, z.x, src);
            zs.y = __shfl_sync(0xffffffffu, z.y, src);
            PsiSet acc = psi_zero();
            for (int j = lane; j < m; j += 32) psi_accumulate<FIVE>(acc, zs + contour_node(radius, j, m));
            acc.p1h = warp_sum(acc.p1h); acc.p2h = warp_sum(acc.p2h);
            acc.p1 = warp_sum(acc.p1); acc.p2 = warp_sum(acc.p2); acc.p3 = warp_sum(acc.p3);
            if (FIVE) {
                acc.p1q = warp_sum(acc.p1q); acc.p2q = warp_sum(acc.p2q);
                acc.p1t = warp_sum(acc.p1t); acc.p2t = warp_sum(acc.p2t);
            }
            if (lane == src) { ps = acc; psi_scale(ps, h, (double)m); }
        }
        if (!valid) return;
        if (FIVE) {
            cplx arr[e5::COUNT];
            arr[e5::E14] = ez.q;
            arr[e5::E12] = ez.h;
            arr[e5::E34] = ez.t;
            arr[e5::E] = ez.f;
            tableau_etd5(ps, arr);
            emit_coefs<M, cplx, GROUPED>(p, i, arr);
        } else {
            cplx arr[kro::COUNT];
            arr[kro::E] = ez.f;
            arr[kro::E2] = ez.h;
            tableau_krogstad(ps, arr);
            emit_coefs<M, cplx, GROUPED>(p, i, arr);
        }
    }
}
// 5 CTAs/SM (96 registers, a few spills): the exp / division chains are latency bound, occupancy pays
// (256^3 ETD35 coefficient set: 2.09 ms at 3 CTAs/SM, 1.55 ms at 5, 1.76 ms at 6)
template <int M, typename LT, bool GROUPED>
__global__ void __launch_bounds__(128, 5) coef_kernel(const __grid_constant__ DevPlan p, int force) { coef_kernel_body<M, LT, GROUPED>(p, force); }
// one plan per blockIdx.z: ensembles whose trajectories keep their own dt (rks_multi_*)
template <int M, typename LT>
__global__ void __launch_bounds__(128, 5) coef_kernel_multi(const DevPlan* plans, int force) { coef_kernel_body<M, LT, false>(plans[blockIdx.z], force); }

// K2 for CM_SEPARABLE plans (IF methods, lin_op = sum of per-axis terms a_d): the per-axis tables
// exp(q h a_d[i]) for every exponent id q of the method, and the rational * h factor of every slot.
// `lin` holds the concatenated axis terms (sep_ntab values); p.coef = [nq][sep_ntab].
template <int M, typename LT>
__global__ void __launch_bounds__(128) coef_sep_kernel(const __grid_constant__ DevPlan p, int force) {
    constexpr int NQ = sep_nq(M), NC = method_ncoef(M);
    Ctrl* c = p.ctrl;
    const double h = c->h;
    if (!force) {
        if (c->status != ST_RUNNING) return;
        if (h == c->h_coeff) return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        atomicAdd((unsigned long long*)&c->coeff_updates, 1ull);
        double* cs = const_cast<double*>(p.cscale);
        for (int s = 0; s < NC; ++s) cs[s] = sep_scale(M, s, h, c->r4_fix);
    }
    if (i >= p.sep_ntab) return;
    LT* out = (LT*)p.coef;
    const LT z = scale(h, load_lin<LT>(p.lin, i));
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q * p.sep_ntab + i] = sep_exp<LT>(M, q, z);
}

// ---------------------------------------------------------------------------------------
// K1: stage combine.  !FULL: block (32, 8), thread = one mode column x R batch rows, the
// coefficients of the column are loaded once.  FULL (lin_op shaped like u): block of 256,
// R elements per thread, coefficients per element.
// ---------------------------------------------------------------------------------------
constexpr int STAGE_R = 2;

// row index of a separable plan (rows = batch x the outer grid axes) -> product over the outer axes of the
// exponential tables, for every exponent id in QMASK
template <int M, typename CT, unsigned QMASK>
RKS_D void sep_row_factors(const DevPlan& p, long long row, CT* aq) {
    constexpr int NQ = sep_nq(M);
    const CT* __restrict__ tab = (const CT*)p.coef;
    const int ntab = p.sep_ntab;
    int i0, i1 = 0;
    if (p.sep_nd == 2) {
        i0 = (int)(row % p.sep_dims[0]);
    } else {
        i1 = (int)(row % p.sep_dims[1]);
        i0 = (int)((row / p.sep_dims[1]) % p.sep_dims[0]);
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        if (!(QMASK & (1u << q))) continue;
        CT v = ldcoef(tab + q * ntab + i0);
        if (p.sep_nd == 3) v = cmul1(v, ldcoef(tab + q * ntab + p.sep_dims[0] + i1));
        aq[q] = v;
    }
}
// the coefficient slots in CMASK from the row factors aq and the column factors bq
template <int M, typename CT, unsigned CMASK>
RKS_D void sep_coefs(const DevPlan& p, const CT* aq, const CT* bq, CT* cv) {
    constexpr int NC = method_ncoef(M);
#pragma unroll
    for (int s = 0; s < NC; ++s) {
        if (!(CMASK & (1u << s))) continue;
        const CT e = cmul1(aq[sep_q(M, s)], bq[sep_q(M, s)]);
        const bool pure = M != M_IF45DP || s <= dp::E;                 // the stage exponentials themselves
        cv[s] = pure ? e : scale(__ldg(p.cscale + s), e);
    }
}

template <int M, int S, typename CT, int CM>
RKS_D void stage_kernel_body(const DevPlan& p) {
    constexpr int R = STAGE_R;
    constexpr bool ADAPT = method_adaptive(M);
    constexpr int SMAX = method_stages(M);
    constexpr bool LAST = (S == SMAX);
    constexpr unsigned NMASK = stage_nl_mask(M, S);
    constexpr unsigned CMASK = stage_coef_mask(M, S);
    constexpr unsigned QMASK = CM == CM_SEPARABLE ? sep_qmask(M, S) : 0u;
    constexpr int NC = method_ncoef(M), NQ = sep_nq(M);
    constexpr bool PER_ELEM = CM == CM_FLAT || CM == CM_INDEXED || CM == CM_SEPARABLE;   // coefficients differ per element of a thread
    const Ctrl* c = p.ctrl;
    int u_sel = 0, n_sel = 0;
    if (ADAPT) {
        if (c->status != ST_RUNNING) return;
        u_sel = c->u_sel; n_sel = c->n_sel;
    }
    const double h = c->h;
    const cplx* u = p.U[u_sel];
    cplx* out = LAST ? (ADAPT ? p.U[1 - u_sel] : p.U[0]) : p.K;
    const CT* __restrict__ coef = (const CT*)p.coef;
    const long long cstride = p.lin_elems;

    long long didx[R], cidx[R], rowi[R];
    bool ok[R];
    if (CM == CM_FLAT) {
        const long long total = p.batch * p.n_c;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long e = ((long long)blockIdx.x * R + r) * 256 + threadIdx.y * 32 + threadIdx.x;
            didx[r] = e; cidx[r] = e; ok[r] = e < total;
        }
    } else if (CM == CM_INDEXED) {
        // x: modes of one trajectory, y: trajectory; the mode's index selects the record of its lin_op value
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long e = ((long long)blockIdx.x * R + r) * 256 + threadIdx.y * 32 + threadIdx.x;
            ok[r] = e < p.n_c;
            didx[r] = (long long)blockIdx.y * p.n_c + e;
            cidx[r] = ok[r] ? (long long)__ldg(p.cidx + e) : 0;
        }
    } else {
        // CM_COLUMN: (batch, n_c); CM_SEPARABLE: the same picture with rows = batch x outer grid axes, columns = last axis
        const long long ncol = CM == CM_SEPARABLE ? p.sep_dims[p.sep_nd - 1] : p.n_c;
        const long long nrow = CM == CM_SEPARABLE ? p.batch * (p.n_c / ncol) : p.batch;
        const long long col = (long long)blockIdx.x * 32 + threadIdx.x;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long row = ((long long)blockIdx.y * 8 + threadIdx.y) * R + r;
            didx[r] = row * ncol + col; cidx[r] = col; rowi[r] = row; ok[r] = (col < ncol) && (row < nrow);
        }
    }

    // ---- loads (all issued before any use)
    CT cv[R][NC];
    CT bq[NQ];
    cplx uv[R], nv[R][8];
    if (CM == CM_SEPARABLE && ok[0]) {
        const CT* __restrict__ tab = coef + (p.sep_ntab - p.sep_dims[p.sep_nd - 1]) + cidx[0];      // last axis' block of the table
#pragma unroll
        for (int q = 0; q < NQ; ++q)
            if (QMASK & (1u << q)) bq[q] = ldcoef(tab + q * p.sep_ntab);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!ok[r]) continue;
        if (CM == CM_SEPARABLE) {
            CT aq[NQ];
            sep_row_factors<M, CT, QMASK>(p, rowi[r], aq);
            sep_coefs<M, CT, CMASK>(p, aq, bq, cv[r]);
        } else if (CM == CM_INDEXED) {
            const CT* __restrict__ rec = coef + cidx[r] * record_elems<CT>(M) + group_off<CT>(M, S);
#pragma unroll
            for (int s = 0; s < NC; ++s)
                if (CMASK & (1u << s)) cv[r][s] = ldcoef(rec + group_pos(CMASK, s));
        } else if (CM == CM_FLAT || r == 0) {
#pragma unroll
            for (int s = 0; s < NC; ++s)
                if (CMASK & (1u << s)) cv[r][s] = ldcoef(coef + s * cstride + cidx[r]);
        }
        uv[r] = (LAST && !ADAPT) ? ld_plain(u + didx[r]) : ldcs(u + didx[r]);   // fixed step: u+ overwrites u in place
#pragma unroll
        for (int j = 1; j <= 7; ++j)
            if (NMASK & (1u << j)) nv[r][j] = ldcs(p.NL[nl_phys(M, j, n_sel)] + didx[r]);
    }
    // ---- combine + store
    unsigned long long mx = 0ull;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (!ok[r]) continue;
        const CT* cr = PER_ELEM ? cv[r] : cv[0];
        const cplx k = stage_combine<M, S, CT>(uv[r], nv[r], cr, h);
        stg(out + didx[r], k);
        if (LAST && ADAPT) {
            const unsigned long long b = nonneg_bits(abs2(k));
            mx = b > mx ? b : mx;
            if (M == M_ETD35) stg(p.ERR + didx[r], etd35_err<CT>(nv[r], cr));
        }
    }
    if (LAST && ADAPT) {
        // block max of |u+|^2 -> one atomicMax per block (solveras.py:451: magu.max())
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o);
            mx = t > mx ? t : mx;
        }
        __shared__ unsigned long long smx[8];
        const int w = threadIdx.y;
        if (threadIdx.x == 0) smx[w] = mx;
        __syncthreads();
        if (w == 0 && threadIdx.x == 0) {
            unsigned long long m = smx[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) m = smx[i] > m ? smx[i] : m;
            if (m) atomicMax((unsigned long long*)&p.ctrl->red[0], m);
        }
    }
}
// Coefficients shared by the batch: three CTAs per SM (<= 85 registers; the six-stage methods' last stages otherwise
// sit at 86-90 registers = two CTAs, where the IF45DP one ran at 0.90 of the roofline: 638 -> 563 us with three).
// Per-element coefficients (full-size arrays, gathered records, separable tables) need the registers.
// resident CTAs the separable / indexed coefficient variants are compiled for (tuning knobs; 1 = no register cap)
#ifndef RKS_STAGE_SEP_BLOCKS
#define RKS_STAGE_SEP_BLOCKS 3      // measured: cfg 4 trial 2.926 -> 2.844 ms (profiles/r02s_cfg4_*.json)
#endif
#ifndef RKS_STAGE_IDX_BLOCKS
#define RKS_STAGE_IDX_BLOCKS 1      // measured: 3 or 4 are slower (cfg 5: 44.1 -> 45.5 / 47.1 ms)
#endif
template <int M, int S, typename CT, int CM>
__global__ void __launch_bounds__(256, (CM == CM_COLUMN ? 3 : CM == CM_SEPARABLE ? RKS_STAGE_SEP_BLOCKS : CM == CM_INDEXED ? RKS_STAGE_IDX_BLOCKS : 1)) stage_kernel(const __grid_constant__ DevPlan p) { stage_kernel_body<M, S, CT, CM>(p); }
// one plan per blockIdx.z: ensembles whose trajectories keep their own dt (rks_multi_*)
template <int M, int S, typename CT, int CM>
__global__ void __launch_bounds__(256) stage_kernel_multi(const DevPlan* plans) { stage_kernel_body<M, S, CT, CM>(plans[blockIdx.z]); }


// K1 for an intermediate stage whose value only feeds the fused NLS-type nonlinearity (fft_fast.cuh, "pre-
// transformed rows").  Thread = one first-pass butterfly of one row, i.e. the R1 modes j + Q1 s of the row: the
// stage value of each mode is combined exactly as in stage_kernel, the radix-R1 inverse butterfly and its
// twiddles are applied in registers (FP64 work this HBM-bound kernel has room for) and the result goes to K in
// the layout the first pass of K4 would have produced.  Same byte model as stage_kernel (reads_s + 1 passes).
//
// A thread that keeps 16 accumulators cannot also keep enough loads in flight from registers (the first version
// did: 14 warps/SM, 15-22 warps stalled on the long scoreboard per issue, 0.88-0.97 of the HBM roofline), so the
// state loads go through shared memory with cp.async: persistent CTAs of 128 threads = 32 columns x 4 rows that
// walk down the batch, every thread streaming CHM modes x NIN arrays per commit group into its own slots of an
// NBUF-deep ring (no barrier: a thread only reads what it copied).  The copies of the next chunks -- of the next
// row, while the butterfly runs -- are always in flight: ~100-190 KB per SM.  The coefficients of the CTA's 32 x R1
// modes are shared by all rows (lin_elems == n_c): loaded into shared memory once.
RKS_D void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
RKS_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING> RKS_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

#ifndef RKS_PRE_CHM2
#define RKS_PRE_CHM2 4       // ring geometry of the two-array stages (stage 1 of every method)
#define RKS_PRE_NBUF2 3
#endif
#ifndef RKS_PRE_DESC
#define RKS_PRE_DESC 1
#endif
constexpr int PRE_THREADS = 128;
template <int M, int S, typename CT, int R1>
struct PreCfg {
    static constexpr unsigned NMASK = stage_nl_mask(M, S);
    static constexpr unsigned CMASK = stage_coef_mask(M, S);
    static constexpr int NIN = 1 + __builtin_popcount(NMASK);      // u and the N_j the stage reads
    static constexpr int NCU = __builtin_popcount(CMASK);          // coefficient arrays the stage reads
    static constexpr int CHM = NIN == 2 ? RKS_PRE_CHM2 : NIN <= 3 ? 4 : 2;     // modes per commit group
    static constexpr int NCH = R1 / CHM;                           // groups per row
    static constexpr int NBUF0 = NIN == 2 ? RKS_PRE_NBUF2 : NIN >= 6 ? 2 : 3;
    static constexpr int NBUF = NBUF0 < NCH ? NBUF0 : NCH;
    static constexpr size_t COEF_BYTES = (size_t)NCU * R1 * 32 * sizeof(CT);
    static constexpr size_t RING_BYTES = (size_t)NBUF * CHM * NIN * PRE_THREADS * sizeof(cplx);
    static constexpr size_t SMEM = COEF_BYTES + RING_BYTES;
};

template <int M, int S, typename CT, int R1>
__global__ void __launch_bounds__(PRE_THREADS) stage_pre_kernel(const __grid_constant__ DevPlan p) {
    using C = PreCfg<M, S, CT, R1>;
    constexpr bool ADAPT = method_adaptive(M);
    constexpr unsigned NMASK = C::NMASK, CMASK = C::CMASK;
    constexpr int NC = method_ncoef(M), NIN = C::NIN, CHM = C::CHM, NCH = C::NCH, NBUF = C::NBUF;
    static_assert(S < method_stages(M), "the last stage value is a state: it stays in natural order");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    CT* csm = reinterpret_cast<CT*>(smem_raw);                                   // [NCU][R1][32]
    cplx* ring = reinterpret_cast<cplx*>(smem_raw + C::COEF_BYTES);              // [NBUF][CHM][NIN][PRE_THREADS]
    const Ctrl* c = p.ctrl;
    int u_sel = 0, n_sel = 0;
    if (ADAPT) {
        if (c->status != ST_RUNNING) return;
        u_sel = c->u_sel; n_sel = c->n_sel;
    }
    const double h = c->h;
    const int q1 = (int)(p.n_c / R1);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int col0 = blockIdx.x * 32, col = col0 + lane;

    // coefficients of this column block, compacted to the slots the stage uses
    {
        const CT* __restrict__ coef = (const CT*)p.coef;
        int ci = 0;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (!(CMASK & (1u << k))) continue;
            for (int e = tid; e < R1 * 32; e += PRE_THREADS)
                csm[ci * R1 * 32 + e] = ldcoef(coef + (long long)k * p.lin_elems + col0 + (e >> 5) * q1 + (e & 31));
            ++ci;
        }
    }
    __syncthreads();

    // the arrays the stage reads: u, then N_j in ascending j
    const cplx* src[NIN];
    src[0] = p.U[u_sel] + col;
    {
        int a = 1;
#pragma unroll
        for (int j = 1; j <= 7; ++j)
            if (NMASK & (1u << j)) src[a++] = p.NL[nl_phys(M, j, n_sel)] + col;
    }
    // Rows are handed out dynamically, one at a time per warp (atomic counter per column block): every CTA does
    // the same amount of work, but SMs do not get the same share of HBM and a static split left 4-13 % of the
    // SM time idle at the end (ncu: smsp__cycles_active / sm__cycles_elapsed).  The warp whose grab ends the
    // column block last re-arms the counters for the next launch.
    int* cnt = p.pre_cnt + 2 * blockIdx.x;
    const int total_warps = (int)gridDim.y * (PRE_THREADS / 32);
    bool drained = false;
    auto grab = [&]() -> long long {
        if (drained) return p.batch;
        int r = 0;
        if (lane == 0) r = atomicAdd(cnt, 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        // K4 wrote the N this stage reads in ascending row order: walk down from the last row, whose lines are
        // the ones still in L2; K is then written towards row 0, where K4 starts reading
        if (r < p.batch) return RKS_PRE_DESC ? p.batch - 1 - r : r;
        drained = true;
        if (lane == 0 && atomicAdd(cnt + 1, 1) == total_warps - 1) {
            cnt[0] = 0; cnt[1] = 0;
            __threadfence();
        }
        return p.batch;
    };
    cplx* my = ring + tid;

    // commit group = CHM modes of the row the producer side is on; it runs NBUF groups ahead of the consumer,
    // so it moves on to the next row (prow) while the consumer is still on the current one
    long long prow = 0;
    int ichunk = 0, ibuf = 0;
    auto issue = [&]() {
        if (ichunk == 0) prow = grab();
        if (prow < p.batch) {
            const long long base = prow * p.n_c + (long long)(ichunk * CHM) * q1;
#pragma unroll
            for (int i = 0; i < CHM; ++i)
#pragma unroll
                for (int a = 0; a < NIN; ++a)
                    cp_async16(my + ((ibuf * CHM + i) * NIN + a) * PRE_THREADS, src[a] + base + (long long)i * q1);
            ibuf = ibuf + 1 == NBUF ? 0 : ibuf + 1;
        }
        ichunk = ichunk + 1 == NCH ? 0 : ichunk + 1;
        cp_async_commit();
    };
    issue();
    long long row = prow;                       // the row whose groups are consumed
#pragma unroll
    for (int g = 1; g < NBUF; ++g) issue();

    int rbuf = 0;
    while (row < p.batch) {
        cplx a[R1];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            cp_async_wait<NBUF - 1>();
#pragma unroll
            for (int i = 0; i < CHM; ++i) {
                const int s = ch * CHM + i;
                const cplx* slot = my + (rbuf * CHM + i) * NIN * PRE_THREADS;
                cplx nv[8];
                CT cv[NC];
                const cplx uv = slot[0];
                int ai = 1, ci = 0;
#pragma unroll
                for (int j = 1; j <= 7; ++j)
                    if (NMASK & (1u << j)) nv[j] = slot[(ai++) * PRE_THREADS];
#pragma unroll
                for (int kk = 0; kk < NC; ++kk)
                    if (CMASK & (1u << kk)) cv[kk] = csm[((ci++) * R1 + s) * 32 + lane];
                a[s] = stage_combine<M, S, CT>(uv, nv, cv, h);
            }
            rbuf = rbuf + 1 == NBUF ? 0 : rbuf + 1;
            issue();                     // refill the buffer just consumed (values are in registers)
        }
        fast::pre_butterfly<R1>(a, p.twf + fast::TW_T1, col);
        cplx* out = p.K + row * p.n_c + col;
#pragma unroll
        for (int r = 0; r < R1; ++r) stg(out + r * q1, a[fast::perm<R1>(r)]);
        row = prow;                      // the producer is on the next row by now (NBUF <= NCH groups ahead)
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------
// K4: fused spectral nonlinearity, one trajectory row per CTA slot, row resident in smem.
// j selects input/output roles; predicated on device so that no host sync is needed.
// ---------------------------------------------------------------------------------------
struct NlRoles {
    const cplx* in;
    cplx* out;
    bool run;
};

RKS_D NlRoles nl_roles(const DevPlan& p, int j, int force) {
    const Ctrl* c = p.ctrl;
    NlRoles r;
    if (c == nullptr) {                                  // standalone row transform (rks_rows_apply)
        r.in = p.U[0]; r.out = p.NL[1]; r.run = true;
        return r;
    }
    const int m = p.method;
    const bool adapt = method_adaptive(m);
    const int S = method_stages(m);
    const int u_sel = adapt ? c->u_sel : 0;
    const int n_sel = adapt ? c->n_sel : 0;
    r.run = force || (c->status == ST_RUNNING && (j != 1 || c->need_n1));
    if (j == 1) r.in = p.U[u_sel];
    else if (j <= S) r.in = p.K;
    else r.in = p.U[1 - u_sel];                      // FSAL: N(u+)
    r.out = p.NL[nl_phys(m, j, n_sel)];
    if (p.nd_inplace) r.in = r.out;                  // N-D grid: the strided-axis kernels already moved the input into N_j
    return r;
}

template <int MODEL>
RKS_D void nl_kernel_body(const DevPlan& p, int j, int force, int rows_per_cta) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cplx* smem = reinterpret_cast<cplx*>(smem_raw);
    const NlRoles roles = nl_roles(p, j, force);
    if (!roles.run) return;
    if (p.ctrl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&p.ctrl->nl_evals, 1ull);

    const int n = (int)p.n;
    const int log2n = p.log2n;
    const int tpr = blockDim.x / rows_per_cta;               // threads per row
    const int lrow = threadIdx.x / tpr, tid = threadIdx.x - lrow * tpr;
    const long long row = (long long)blockIdx.x * rows_per_cta + lrow;
    const bool active = row < p.batch;
    cplx* x = smem + (size_t)lrow * n;
    const long long rr = active ? row : p.batch - 1;
    const cplx* in = roles.in + rr * p.n_c;
    cplx* out = roles.out + rr * p.n_c;
    const auto m = fast::ModelOf<MODEL>::make(in, out, p.kx, p.model_p0, n, active);
    if (active)
        for (int q = tid; q < n; q += tpr) x[q] = m.load(q);
    __syncthreads();
    const int np = fft_num_passes(log2n);
    for (int q = 0; q < np; ++q) {
        if (active) ifft_dif_pass(x, log2n, q, p.tw, tid, tpr);
        __syncthreads();
    }
    if (active)
        for (int q = tid; q < n; q += tpr) x[q] = fast::pointwise_of(m, x[q], q);
    __syncthreads();
    for (int q = 0; q < np; ++q) {
        if (active) fft_dit_pass(x, log2n, q, p.tw, tid, tpr);
        __syncthreads();
    }
    if (active)
        for (int q = tid; q < n; q += tpr) m.store(q, x[q]);
}
template <int MODEL>
__global__ void __launch_bounds__(1024) nl_kernel(const __grid_constant__ DevPlan p, int j, int force, int rows_per_cta) { nl_kernel_body<MODEL>(p, j, force, rows_per_cta); }
// one plan per blockIdx.z: ensembles whose trajectories keep their own dt (rks_multi_*)
template <int MODEL>
__global__ void __launch_bounds__(1024) nl_kernel_multi(const DevPlan* plans, int j, int force, int rows_per_cta) { nl_kernel_body<MODEL>(plans[blockIdx.z], j, force, rows_per_cta); }



// ---------------------------------------------------------------------------------------
// K4 fast path (fft_fast.cuh): n = 512 W, W warps per row, 16 values per thread in registers,
// persistent CTAs looping over row groups.  RPC rows share one CTA when rows are short.
// ---------------------------------------------------------------------------------------
template <int TR>
RKS_D void row_barrier(int lrow, int rpc) {
    if (TR == 32) { __syncwarp(); return; }
    if (rpc == 1) { __syncthreads(); return; }
    asm volatile("bar.sync %0, %1;" ::"r"(lrow + 1), "r"(TR) : "memory");
}

struct NoHook { RKS_D void operator()() const {} };

// n = 8192 (one row per CTA, 16 warps): the warps leave every row barrier in step, so the four warps of a
// scheduler load together, compute together and store together and the LSU and the FP64 pipe are busy in turn
// instead of at the same time.  Before the last-pass butterflies of a row (the first FP64-heavy phase after the
// barrier) the warps of a scheduler (slot = warp / 4) are therefore put one phase apart: one computes while the
// previous one already stores and loads the next row, and the offsets survive the barrier-free part of that row.
// Measured in round 2 (profiles/r02t_row_stagger.md; same box, interleaved runs): pre-transformed rows
// 312-314 -> 285-288 us per evaluation, plain rows 338 -> 316 us; no effect at n = 4096 (two CTAs per SM are out
// of step anyway), so n = 8192 only.
//   RKS_ROW_STAGGER_MODE 1 (default): slot s spins s * RKS_ROW_STAGGER_CYC (default 1000; 800 ... 1200 measured alike,
//      1300 and more lose the gain) cycles;
//   2: slot s starts its butterfly when slot s - 1 has finished (named barriers 1..3, self-timed: 290-293 us);
//   0: off.  RKS_ROW_STAGGER_NP=0 switches the spin of the rows that are not pre-transformed off.
__constant__ int c_row_stagger_mode;
__constant__ int c_row_stagger_cyc;
__constant__ int c_row_stagger_np;
RKS_D void spin_cycles(long long c) {
    const long long t0 = clock64();
    while (clock64() - t0 < c) {}
}
template <int W>
RKS_D void row_stagger_spin(int T) {
    if (W != 16) return;
    const int slot = T >> 7;
    if (slot) spin_cycles((long long)slot * c_row_stagger_cyc);
}

// `after_first` runs once the first pass has consumed the row's input (staging buffer free again).
// PT: the row is pre-transformed (fft_fast.cuh pre_butterfly: K1 already applied the first inverse pass), so
// the first pass here is the warp-local middle pass reading global memory / the staging buffer.  `arrived`
// (PT with a staging buffer): shared counter of the warps that have consumed the buffer.
template <int N, bool PT = false, class Model, class Hook = NoHook>
RKS_D void nl_fast_row(cplx* sm, int T, int lrow, int rpc, const fast::Twiddles& ti, const fast::Twiddles& tf,
                       const Model& m, const Hook& after_first = Hook(), int* arrived = nullptr) {
    using P = fast::Plan<N>;
    constexpr int W = P::W, TR = 32 * W;
    if (PT) {
        fast::phase_pre<N>(sm, T, ti, m);
        __syncwarp();
        if (!std::is_same<Hook, NoHook>::value) {
            // No CTA barrier: the warp that consumes the staging buffer LAST refills it (the other warps run on into
            // their warp-local passes).  Release/acquire through the shared counter orders every warp's reads of
            // the buffer before the bulk copy that overwrites it.
            if ((T & 31) == 0) {
                __threadfence_block();
                if (atomicAdd(arrived, 1) == W - 1) {
                    atomicExch(arrived, 0);          // re-armed long before any warp gets here again (row barrier below)
                    __threadfence_block();
                    after_first();
                }
            }
        }
    } else {
        fast::phase_first<N>(sm, T, ti, m);
        row_barrier<TR>(lrow, rpc);
        after_first();
    }
    if (!PT) { fast::phase_middle<N, 2, true>(sm, T, ti, m);   __syncwarp(); }
    if (fast::middle_passes<N>() == 2) { fast::phase_middle<N, 3, true>(sm, T, ti, m);   __syncwarp(); }
    fast::phase_core<N>(sm, T, m);                  __syncwarp();
    if (fast::middle_passes<N>() == 2) { fast::phase_middle<N, 3, false>(sm, T, tf, m);  __syncwarp(); }
    fast::phase_middle<N, 2, false>(sm, T, tf, m);
    row_barrier<TR>(lrow, rpc);
    if constexpr (!PT) {
        if (W == 16 && c_row_stagger_np) row_stagger_spin<W>(T);
        fast::phase_last<N>(sm, T, tf, m);
        // no barrier here -- the last pass reads exactly the slab positions (T + 32 W c + Q1 s) that the same
        // thread overwrites in the first pass of its next row, so warps run on into the next row's loads.
    } else {
        // PT: the next row's first pass writes the warp's own 512-point slice, which other warps read in this pass.
        // The barrier sits right after the slab loads of the (single) last-pass butterfly of a thread, so the
        // butterfly, its 16 global stores and the next row's first pass are not held up by the slowest warp.
        constexpr int R1 = P::R1, Q1 = N / R1;
        static_assert(Q1 == 32 * W, "pre-transformed rows: one last-pass butterfly per thread");
        cplx a[R1];
        fast::bf_load<R1, Q1, P::SH>(sm, T, a);
        row_barrier<TR>(lrow, rpc);
        const int mode = W == 16 ? c_row_stagger_mode : 0;
        const int slot = T >> 7;
        if (mode == 1) row_stagger_spin<W>(T);
        if (mode == 2 && slot > 0) asm volatile("bar.sync %0, 256;" ::"r"(slot) : "memory");
        fast::bf_dit<R1, Q1, fast::TW_S1>(a, tf.t1, T);
        if (mode == 2 && slot < W / 4 - 1) asm volatile("bar.arrive %0, 256;" ::"r"(slot + 1) : "memory");
        fast::bf_store_global<R1, Q1>(m, T, a);
    }
}

// pull a row that will be needed soon from HBM into L2 (no registers, no smem)
template <int TR>
RKS_D void prefetch_row_l2(const void* row, int lines, int T) {
    const char* base = reinterpret_cast<const char*>(row);
    for (int q = T; q < lines; q += TR) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((size_t)q << 7)));
}

// TMA staging of the next input row (n = 8192: one row per SM, see fft_fast.cuh StagedRow).  One
// mbarrier, one elected thread: arm it with the byte count, issue the bulk copies; every thread
// waits on the phase parity before the first pass reads the staging buffer.
constexpr int NL_STAGE_ELEMS = 6144;                 // 96 KB next to the 128 KB row slab
constexpr int NL_STAGE_BYTES = NL_STAGE_ELEMS * 16 + 16;
RKS_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
RKS_D void stage_init(unsigned long long* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
RKS_D void stage_issue(cplx* stg, const cplx* src, unsigned bytes, unsigned long long* bar) {
    const unsigned b = smem_u32(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads of stg vs the async writes
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    constexpr unsigned CHUNK = 32768;
    for (unsigned off = 0; off < bytes; off += CHUNK) {
        const unsigned sz = bytes - off < CHUNK ? bytes - off : CHUNK;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stg) + off), "l"(reinterpret_cast<const char*>(src) + off), "r"(sz), "r"(b)
                     : "memory");
    }
}
RKS_D void stage_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RKS_STAGE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RKS_STAGE_DONE;\n"
        "bra RKS_STAGE_WAIT;\n"
        "RKS_STAGE_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Pre-transformed rows read their input slice by slice (warp w owns points 512 w ... 512 w + 511), so staging the HEAD
// of the row would leave the last warps -- the ones the stagger starts last -- reading global memory only.  Instead the
// first PT_SLICE_STAGED points of EVERY slice are staged: 16 bulk copies of 6 KB on the same mbarrier.
#ifndef RKS_PT_SLICED
#define RKS_PT_SLICED 1
#endif
constexpr int PT_SLICE_STAGED = fast::SlicedStagedRow::SL;
RKS_D void stage_issue_sliced(cplx* stg, const cplx* src, unsigned long long* bar) {
    const unsigned b = smem_u32(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(16u * PT_SLICE_STAGED * 16u) : "memory");
    for (int w = 0; w < 16; ++w)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(stg + w * PT_SLICE_STAGED)), "l"(src + w * 512), "r"(PT_SLICE_STAGED * 16u), "r"(b)
                     : "memory");
}
struct SlicedStageNext {
    cplx* stg; const cplx* src; unsigned long long* bar; bool go;
    RKS_D void operator()() const { if (go) stage_issue_sliced(stg, src, bar); }
};
struct StageNext {          // after the first pass: start copying the head of the next row
    cplx* stg; const cplx* src; unsigned bytes; unsigned long long* bar; bool go;
    RKS_D void operator()() const { if (go) stage_issue(stg, src, bytes, bar); }
};

// N_j = N(existing array) for rows of n = 512 W points.
// PT: the input row is a pre-transformed stage value (stage_pre_kernel above; complex-field models).
template <int W, int MODEL, bool PT = false>
RKS_D void nl_fast_kernel_body(const DevPlan& p, int j, int force) {
    static_assert(!PT || MODEL == 2, "pre-transformed rows: the NLS model only");
    constexpr int N = 512 * W;
    constexpr int TR = 32 * W;                       // threads per row
    constexpr int THREADS = W == 16 ? 512 : 256;
    constexpr int RPC = THREADS / TR;                // rows per CTA
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const NlRoles roles = nl_roles(p, j, force);
    if (!roles.run) return;
    if (p.ctrl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&p.ctrl->nl_evals, 1ull);

    const int lrow = threadIdx.x / TR, T = threadIdx.x - lrow * TR;
    cplx* sm = reinterpret_cast<cplx*>(smem_raw) + (size_t)lrow * N;
    const fast::Twiddles ti{p.twf + fast::TW_T1, p.twf + fast::TW_T2, p.twf + fast::TW_T3};
    const cplx* twf2 = p.twf + fast::TW_TOTAL;
    const fast::Twiddles tf{twf2 + fast::TW_T1, twf2 + fast::TW_T2, twf2 + fast::TW_T3};
    const int lines = (int)((p.n_c * 16 + 127) >> 7);
    const long long groups = (p.batch + RPC - 1) / RPC;

    // n = 8192: input rows arrive through the TMA staging buffer
    constexpr bool STAGED = W == 16 && MODEL >= 1 && MODEL <= 3;
    cplx* stg = reinterpret_cast<cplx*>(smem_raw) + (size_t)RPC * N;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(stg + NL_STAGE_ELEMS);
    int* arrived = reinterpret_cast<int*>(bar + 1);          // PT: warps that have consumed the staging buffer
    const int nst = p.n_c < NL_STAGE_ELEMS ? (int)p.n_c : NL_STAGE_ELEMS;
    unsigned parity = 0;
    if (STAGED) {
        if (threadIdx.x == 0) {
            *arrived = 0;
            stage_init(bar);
            if ((long long)blockIdx.x < groups) {
                if (PT && RKS_PT_SLICED) stage_issue_sliced(stg, roles.in + (long long)blockIdx.x * p.n_c, bar);
                else stage_issue(stg, roles.in + (long long)blockIdx.x * p.n_c, nst * 16u, bar);
            }
        }
        __syncthreads();
    }

    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long row = g * RPC + lrow;
        const bool on = row < p.batch;
        const long long rr = on ? row : p.batch - 1;
        const long long nrow = row + (long long)gridDim.x * RPC;     // the row this slot handles next
        const int nlines = nrow < p.batch ? lines : 0;
        cplx* out = roles.out + rr * p.n_c;
        if constexpr (STAGED && PT && RKS_PT_SLICED) {
            // L2-prefetch the unstaged tail of every slice of the next row: 16 lines of 128 bytes per slice
            if (nlines && T < 256) {
                const char* base = reinterpret_cast<const char*>(roles.in + nrow * p.n_c);
                const int w = T >> 4, l = T & 15;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + ((size_t)(512 * w + PT_SLICE_STAGED) << 4) + ((size_t)l << 7)));
            }
            stage_wait(bar, parity);
            parity ^= 1u;
            const fast::NlsModelT<fast::SlicedStagedRow> m{fast::SlicedStagedRow{roles.in + rr * p.n_c, stg}, out, p.model_p0, N, on};
            const SlicedStageNext next{stg, roles.in + (nlines ? nrow : rr) * p.n_c, bar, nlines != 0};
            nl_fast_row<N, PT>(sm, T, lrow, RPC, ti, tf, m, next, arrived);
        } else if constexpr (STAGED) {
            // L2-prefetch only the tail of the next row that does not fit the staging buffer
            const int head = (nst * 16) >> 7;
            if (nlines > head)
                prefetch_row_l2<TR>(reinterpret_cast<const char*>(roles.in + nrow * p.n_c) + ((size_t)head << 7), nlines - head, T);
            stage_wait(bar, parity);
            parity ^= 1u;
            const auto m = fast::ModelOf<MODEL>::make_staged(fast::StagedRow{roles.in + rr * p.n_c, stg, nst}, out, p.kx,
                                                             p.model_p0, N, on);
            // !PT: thread 0 issues the copy after the row barrier; PT: the lane the counter elects
            const StageNext next{stg, roles.in + (nlines ? nrow : rr) * p.n_c, nst * 16u, bar, nlines != 0 && (PT || threadIdx.x == 0)};
            nl_fast_row<N, PT>(sm, T, lrow, RPC, ti, tf, m, next, arrived);
        } else {
            prefetch_row_l2<TR>(roles.in + (nlines ? nrow : rr) * p.n_c, nlines, T);
            const auto m = fast::ModelOf<MODEL>::make(roles.in + rr * p.n_c, out, p.kx, p.model_p0, N, on);
            nl_fast_row<N, PT>(sm, T, lrow, RPC, ti, tf, m);
        }
    }
}
template <int W, int MODEL>
__global__ void __launch_bounds__((W == 16 ? 512 : 256), (W == 16 ? 1 : 2)) nl_fast_kernel(const __grid_constant__ DevPlan p, int j, int force) { nl_fast_kernel_body<W, MODEL>(p, j, force); }
// the same evaluation of a row K1 has pre-transformed (NLS model)
template <int W>
__global__ void __launch_bounds__((W == 16 ? 512 : 256), (W == 16 ? 1 : 2)) nl_fast_pre_kernel(const __grid_constant__ DevPlan p, int j, int force) { nl_fast_kernel_body<W, 2, true>(p, j, force); }
// one plan per blockIdx.z: ensembles whose trajectories keep their own dt (rks_multi_*)
template <int W, int MODEL>
__global__ void __launch_bounds__((W == 16 ? 512 : 256), (W == 16 ? 1 : 2)) nl_fast_kernel_multi(const DevPlan* plans, int j, int force) { nl_fast_kernel_body<W, MODEL>(plans[blockIdx.z], j, force); }


// ---------------------------------------------------------------------------------------
// EXPERIMENT, opt-in (RKS_RFFT_HALF=1), not measured yet: K4 for the real-field models (1 = u u_x, 3 = cubic) with the
// forward transform of the real product as a half-length complex transform (fft_real.cuh; CPU-pinned against NumPy
// in tests/test_device_math_host.py).  n = 512 ... 4096, plain evaluation only.  Half the forward butterflies, but
// three more row barriers per row (the E/O split needs C[n/2 - k] from another thread).
// ---------------------------------------------------------------------------------------
template <int N, class Model>
RKS_D void nl_fast_row_real(cplx* sm, int T, int lrow, int rpc, const fast::Twiddles& ti, const fast::Twiddles& tf,
                            const Model& m) {
    using P = fast::Plan<N>;
    constexpr int TR = 32 * P::W, H = P::R1 / 2, NB = (N / P::R1) / TR;
    fast::phase_first<N>(sm, T, ti, m);
    row_barrier<TR>(lrow, rpc);
    fast::phase_middle<N, 2, true>(sm, T, ti, m);   __syncwarp();
    fast::phase_core_pair<N>(sm, T, m);             __syncwarp();
    fast::phase_middle_even<N>(sm, T, tf);
    row_barrier<TR>(lrow, rpc);
    cplx x[NB * H];
    fast::phase_last_half_load<N>(sm, T, tf, x);
    row_barrier<TR>(lrow, rpc);                      // every thread has read its G values: the slab may take C
    fast::phase_last_half_exchange<N>(sm, T, x);
    row_barrier<TR>(lrow, rpc);
    fast::phase_last_half_split<N>(sm, T, tf, x, m);
    row_barrier<TR>(lrow, rpc);                      // partners' C values are read before the next row overwrites them
}
template <int W, int MODEL>
__global__ void __launch_bounds__(256, 2) nl_fast_real_kernel(const __grid_constant__ DevPlan p, int j, int force) {
    static_assert(W <= 8 && (MODEL == 1 || MODEL == 3), "real-field models, rows of 512 ... 4096 points");
    constexpr int N = 512 * W, TR = 32 * W, THREADS = 256, RPC = THREADS / TR;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const NlRoles roles = nl_roles(p, j, force);
    if (!roles.run) return;
    if (p.ctrl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&p.ctrl->nl_evals, 1ull);
    const int lrow = threadIdx.x / TR, T = threadIdx.x - lrow * TR;
    cplx* sm = reinterpret_cast<cplx*>(smem_raw) + (size_t)lrow * N;
    const fast::Twiddles ti{p.twf + fast::TW_T1, p.twf + fast::TW_T2, p.twf + fast::TW_T3};
    const cplx* twf2 = p.twf + fast::TW_TOTAL;
    const fast::Twiddles tf{twf2 + fast::TW_T1, twf2 + fast::TW_T2, twf2 + fast::TW_T3};
    const int lines = (int)((p.n_c * 16 + 127) >> 7);
    const long long groups = (p.batch + RPC - 1) / RPC;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long row = g * RPC + lrow;
        const bool on = row < p.batch;
        const long long rr = on ? row : p.batch - 1;
        const long long nrow = row + (long long)gridDim.x * RPC;
        const int nlines = nrow < p.batch ? lines : 0;
        prefetch_row_l2<TR>(roles.in + (nlines ? nrow : rr) * p.n_c, nlines, T);
        const auto m = fast::ModelOf<MODEL>::make(roles.in + rr * p.n_c, roles.out + rr * p.n_c, p.kx, p.model_p0, N, on);
        nl_fast_row_real<N>(sm, T, lrow, RPC, ti, tf, m);
    }
}

// ---------------------------------------------------------------------------------------
// K4 for the single-real-field cubic model, two rows per complex transform (fft_pair.cuh): n = 512 ... 4096.
// A slab holds a row PAIR; 256 threads = 256 / (32 W) pairs per CTA, two CTAs per SM, persistent over pair groups.
// ---------------------------------------------------------------------------------------
template <int N, class Model>
RKS_D void nl_fast_row_pair(cplx* sm, int T, int lrow, int rpc, const fast::Twiddles& ti, const fast::Twiddles& tf,
                            const Model& m) {
    using P = fast::Plan<N>;
    constexpr int TR = 32 * P::W, NB = (N / P::R1) / TR;
    fast::phase_first<N>(sm, T, ti, m);
    row_barrier<TR>(lrow, rpc);
    fast::phase_middle<N, 2, true>(sm, T, ti, m);   __syncwarp();
    fast::phase_core<N>(sm, T, m);                  __syncwarp();
    fast::phase_middle<N, 2, false>(sm, T, tf, m);
    row_barrier<TR>(lrow, rpc);
    cplx a[NB * P::R1];
    fast::phase_pair_load<N>(sm, T, a);
    row_barrier<TR>(lrow, rpc);                      // every thread holds its butterflies: the slab may take W
    fast::phase_pair_publish<N>(sm, T, tf, a);
    row_barrier<TR>(lrow, rpc);
    fast::phase_pair_store<N>(sm, T, a, m);
    row_barrier<TR>(lrow, rpc);                      // partners are read before the next pair's first pass overwrites them
}
template <int W>
RKS_D void nl_fast_pair_kernel_body(const DevPlan& p, int j, int force) {
    static_assert(W <= 8, "row pairs of 512 ... 4096 points");
    constexpr int N = 512 * W, TR = 32 * W, THREADS = 256, RPC = THREADS / TR;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const NlRoles roles = nl_roles(p, j, force);
    if (!roles.run) return;
    if (p.ctrl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&p.ctrl->nl_evals, 1ull);
    const int lrow = threadIdx.x / TR, T = threadIdx.x - lrow * TR;
    cplx* sm = reinterpret_cast<cplx*>(smem_raw) + (size_t)lrow * N;
    const fast::Twiddles ti{p.twf + fast::TW_T1, p.twf + fast::TW_T2, p.twf + fast::TW_T3};
    const cplx* twf2 = p.twf + fast::TW_TOTAL;
    const fast::Twiddles tf{twf2 + fast::TW_T1, twf2 + fast::TW_T2, twf2 + fast::TW_T3};
    const int lines = (int)((2 * p.n_c * 16 + 127) >> 7);           // a pair is two consecutive rows
    const long long pairs = (p.batch + 1) / 2;
    const long long groups = (pairs + RPC - 1) / RPC;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long pair = g * RPC + lrow;
        const long long ra = 2 * pair, rb = ra + 1;
        const bool on_a = ra < p.batch, on_b = rb < p.batch;
        const long long sa = on_a ? ra : p.batch - 1, sb = on_b ? rb : p.batch - 1;
        const long long npair = pair + (long long)gridDim.x * RPC;
        if (2 * npair < p.batch) {
            const int nl = 2 * npair + 1 < p.batch ? lines : (lines + 1) / 2;
            prefetch_row_l2<TR>(roles.in + 2 * npair * p.n_c, nl, T);
        }
        const fast::PairedCubicModel m{roles.in + sa * p.n_c, roles.in + sb * p.n_c, roles.out + sa * p.n_c,
                                       roles.out + sb * p.n_c, p.model_p0, N, on_a, on_b};
        nl_fast_row_pair<N>(sm, T, lrow, RPC, ti, tf, m);
    }
}
template <int W>
__global__ void __launch_bounds__(256, 2) nl_fast_pair_kernel(const __grid_constant__ DevPlan p, int j, int force) {
    nl_fast_pair_kernel_body<W>(p, j, force);
}

// ---------------------------------------------------------------------------------------
// K4 for short rows (n = 64, 128, 256): every warp is on its own -- it packs 512 / n consecutive rows into
// its 512-point slab and runs the 512-point pipeline on it (fft_fast.cuh PackedModel); no CTA barrier.
// ---------------------------------------------------------------------------------------
constexpr int NL_SMALL_WARPS = 8;
template <int N, int MODEL>
RKS_D void nl_small_kernel_body(const DevPlan& p, int j, int force) {
    constexpr int SUB = 512 / N;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const NlRoles roles = nl_roles(p, j, force);
    if (!roles.run) return;
    if (p.ctrl && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&p.ctrl->nl_evals, 1ull);
    const int w = threadIdx.x >> 5, T = threadIdx.x & 31;
    cplx* sm = reinterpret_cast<cplx*>(smem_raw) + (size_t)w * 512;
    const fast::Twiddles ti{p.twf + fast::TW_T1, p.twf + fast::TW_T2, p.twf + fast::TW_T3};
    const cplx* twf2 = p.twf + fast::TW_TOTAL;
    const fast::Twiddles tf{twf2 + fast::TW_T1, twf2 + fast::TW_T2, twf2 + fast::TW_T3};
    const long long slabs = (p.batch + SUB - 1) / SUB;
    for (long long s = (long long)blockIdx.x * NL_SMALL_WARPS + w; s < slabs; s += (long long)gridDim.x * NL_SMALL_WARPS) {
        const long long row0 = s * SUB;
        const long long left = p.batch - row0;
        const fast::PackedModel<MODEL, N> m{roles.in + row0 * p.n_c, roles.out + row0 * p.n_c, p.kx, p.model_p0, p.n_c,
                                            (int)(left < SUB ? left : SUB)};
        fast::phase_first_packed<N>(sm, T, ti, m);        __syncwarp();
        fast::phase_middle<N, 2, true>(sm, T, ti, m);     __syncwarp();
        fast::phase_core<N>(sm, T, m);                    __syncwarp();
        fast::phase_middle<N, 2, false>(sm, T, tf, m);    __syncwarp();
        fast::phase_last_packed<N>(sm, T, tf, m);
        // no barrier: the last pass reads the slab positions this thread overwrites in its next first pass
    }
}
template <int N, int MODEL>
__global__ void __launch_bounds__(32 * NL_SMALL_WARPS, 2) nl_small_kernel(const __grid_constant__ DevPlan p, int j, int force) {
    nl_small_kernel_body<N, MODEL>(p, j, force);
}

// ---------------------------------------------------------------------------------------
// K3: masked norms (solveras.py:451-454) + controller tail.
// grid (gx, gy): columns grid-stride over x, rows grid-stride over y; 128 threads.
// ---------------------------------------------------------------------------------------
template <int BT>
RKS_D double block_sum(double v, double* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < BT / 32; ++i) t += sh[i];
    }
    return t;     // valid in thread 0
}

template <int M, typename CT, int CM>
RKS_D void norm_kernel_body(const DevPlan& p, int fuse_controller) {
    constexpr unsigned NMASK = err_nl_mask(M);
    constexpr unsigned CMASK = err_coef_mask(M);
    constexpr unsigned QMASK = CM == CM_SEPARABLE ? sep_qmask(M, method_stages(M) + 1) : 0u;
    constexpr int NC = method_ncoef(M), NQ = sep_nq(M);
    constexpr bool FULL = CM == CM_FLAT;
    Ctrl* c = p.ctrl;
    if (c->status != ST_RUNNING) return;
    const int u_sel = c->u_sel, n_sel = c->n_sel;
    const double h = c->h;
    const double m = sqrt(c->red[0]);                       // max |u+|
    const double cutoff = c->adapt_cutoff;
    const cplx* __restrict__ un = p.norm_u ? p.norm_u : p.U[1 - u_sel];
    const CT* __restrict__ coef = (const CT*)p.coef;
    const long long cstride = p.lin_elems;
    // columns x rows: (n_c, batch); flat: one long row; separable: (last grid axis, batch x outer axes)
    const long long ncols = FULL ? p.batch * p.n_c : CM == CM_SEPARABLE ? p.sep_dims[p.sep_nd - 1] : p.n_c;
    const long long nrows = FULL ? 1 : CM == CM_SEPARABLE ? p.batch * (p.n_c / ncols) : p.batch;

    constexpr int NLOADS = 1 + (M == M_ETD35 ? 1 : __builtin_popcount(NMASK));
    constexpr int UN = (FULL || CM == CM_INDEXED) ? 1 : NLOADS <= 2 ? 4 : NLOADS <= 3 ? 2 : 1;   // flat / indexed: the column loop is the long one
    const double thr2 = (cutoff * m) * (cutoff * m);
    const bool band_ok = thr2 > 1e-290 && thr2 < 1e290;                         // else: always the exact test
    const double band_hi = band_ok ? thr2 * (1.0 + 1e-9) : __longlong_as_double(0x7ff0000000000000ll);
    const double band_lo = band_ok ? thr2 * (1.0 - 1e-9) : 0.0;

    double su = 0.0, se = 0.0;
    for (long long col = (long long)blockIdx.x * 128 + threadIdx.x; col < ncols; col += (long long)gridDim.x * 128) {
        CT cv[NC];
        CT bq[NQ];
        if (CM == CM_SEPARABLE) {
            const CT* __restrict__ tab = coef + (p.sep_ntab - ncols) + col;
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                if (QMASK & (1u << q)) bq[q] = ldcoef(tab + q * p.sep_ntab);
        } else if (CM == CM_INDEXED) {
            if (CMASK) {
                const CT* __restrict__ rec = coef + (long long)__ldg(p.cidx + col) * record_elems<CT>(M) + group_off<CT>(M, method_stages(M) + 1);
#pragma unroll
                for (int s = 0; s < NC; ++s)
                    if (CMASK & (1u << s)) cv[s] = ldcoef(rec + group_pos(CMASK, s));
            }
        } else {
#pragma unroll
            for (int s = 0; s < NC; ++s)
                if (CMASK & (1u << s)) cv[s] = ldcoef(coef + s * cstride + col);
        }
        // UN rows per iteration, all loads issued first (a thread with one row in flight keeps ~32 KB per SM on
        // the wire: 5.2 TB/s; with four it is the HBM roofline)
        for (long long row0 = blockIdx.y; row0 < nrows; row0 += (long long)UN * gridDim.y) {
            cplx uv[UN], ev[UN], nv[UN][8];
            bool ok[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                const long long row = row0 + (long long)q * gridDim.y;
                ok[q] = row < nrows;
                const long long d = (ok[q] ? row : row0) * ncols + col;
                uv[q] = ldcs(un + d);
                if (M == M_ETD35) {
                    ev[q] = ldcs(p.ERR + d);
                } else {
#pragma unroll
                    for (int j = 1; j <= 7; ++j)
                        if (NMASK & (1u << j)) nv[q][j] = ldcs(p.NL[nl_phys(M, j, n_sel)] + d);
                }
            }
#pragma unroll
            for (int q = 0; q < UN; ++q) {
                if (CM == CM_SEPARABLE && CMASK) {
                    CT aq[NQ];
                    const long long row = row0 + (long long)q * gridDim.y;
                    sep_row_factors<M, CT, QMASK>(p, ok[q] ? row : row0, aq);
                    sep_coefs<M, CT, CMASK>(p, aq, bq, cv);
                }
                if (M != M_ETD35) ev[q] = embedded_err<M, CT>(nv[q], cv, h);
                const double u2 = abs2(uv[q]);
                // idx = magu / magu.max() > adapt_cutoff   (solveras.py:452).  The square root and the division
                // decide only inside a 1e-9 band around (cutoff m)^2: outside it the comparison of the squares
                // gives the same answer (their rounding errors are ~1e-16), at a fraction of the FP64 work.
                const bool in = u2 > band_hi || (u2 >= band_lo && sqrt(u2) / m > cutoff);
                if (ok[q] && in) {
                    su += u2;
                    se += abs2(ev[q]);
                }
            }
        }
    }
    __shared__ double sh[4];
    __shared__ bool is_last;
    const double bu = block_sum<128>(su, sh);
    const double be = block_sum<128>(se, sh);
    const unsigned nblocks = gridDim.x * gridDim.y;
    const unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    if (threadIdx.x == 0) {
        p.partials[2 * bid] = bu;
        p.partials[2 * bid + 1] = be;
        __threadfence();
        const unsigned t = atomicAdd(&c->ticket, 1u);
        is_last = (t == nblocks - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // deterministic final reduction of the per-block partials
    double fu = 0.0, fe = 0.0;
    for (unsigned i = threadIdx.x; i < nblocks; i += 128) {
        fu += __ldcg(p.partials + 2 * i);
        fe += __ldcg(p.partials + 2 * i + 1);
    }
    const double tu = block_sum<128>(fu, sh);
    const double te = block_sum<128>(fe, sh);
    if (threadIdx.x == 0) {
        c->red[1] = tu;
        c->red[2] = te;
        c->ticket = 0u;
        if (fuse_controller) {
            Ctrl local = *c;
            controller_advance(local, p.log);
            *c = local;
        }
    }
}
template <int M, typename CT, int CM>
__global__ void __launch_bounds__(128) norm_kernel(const __grid_constant__ DevPlan p, int fuse_controller) { norm_kernel_body<M, CT, CM>(p, fuse_controller); }
// one plan per blockIdx.z: ensembles whose trajectories keep their own dt (rks_multi_*)
template <int M, typename CT, int CM>
__global__ void __launch_bounds__(128) norm_kernel_multi(const DevPlan* plans, int fuse_controller) { norm_kernel_body<M, CT, CM>(plans[blockIdx.z], fuse_controller); }


// max |x|^2 over an array -> ctrl.red[0], REPLACING what the last stage kernel accumulated there
// (diagonalize=True: the mask and tolerance of the controller come from the physical S u+)
__global__ void __launch_bounds__(1024) max_abs2_kernel(DevPlan p, const cplx* x, long long count) {
    Ctrl* c = p.ctrl;
    if (c->status != ST_RUNNING) return;
    unsigned long long best = 0ull;
    for (long long e = threadIdx.x; e < count; e += 1024) {
        const cplx v = x[e];
        const double a = v.x * v.x + v.y * v.y;
        const unsigned long long bits = (a != a) ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(a);
        best = bits > best ? bits : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t > best ? t : best;
    }
    __shared__ unsigned long long sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) best = sh[w] > best ? sh[w] : best;
        c->red[0] = __longlong_as_double((long long)best);
    }
}

__global__ void controller_kernel(DevPlan p) {
    Ctrl* c = p.ctrl;
    if (c->status != ST_RUNNING) return;
    Ctrl local = *c;
    controller_advance(local, p.log);
    *c = local;
}

// Dense complex matrix times vector(s), y[b] = A x[b]: the basis changes of diagonalize=True (etd35.py:463, 495:
// N'(k) = S^-1 N(S k), |S u+| for the controller).  One warp per matrix row, lanes stride the row (512-byte coalesced
// segments of A, x from L1/L2), shuffle reduction; grid (ceil(n / 4), batch).  Bound by the read of A (n^2 x 16 B).
__global__ void __launch_bounds__(128) gemv_kernel(const cplx* __restrict__ a, const cplx* __restrict__ x, cplx* __restrict__ y,
                                                   int n) {
    const int lane = threadIdx.x & 31, i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const cplx* row = a + (size_t)i * n;
    const cplx* xb = x + (size_t)blockIdx.y * n;
    double sr = 0.0, si = 0.0;
    for (int j = lane; j < n; j += 32) {
        const cplx av = ldg(row + j), xv = ldg(xb + j);
        sr += av.x * xv.x - av.y * xv.y;
        si += av.x * xv.y + av.y * xv.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, o);
        si += __shfl_xor_sync(0xffffffffu, si, o);
    }
    if (lane == 0) y[(size_t)blockIdx.y * n + i] = mk(sr, si);
}

// pointwise nonlinearity of the N-D models, applied between two library transforms (one read + one
// write instead of the half-dozen elementwise passes a torch expression costs)
__global__ void __launch_bounds__(256) pointwise_nls_kernel(const cplx* in, cplx* out, long long count, double gamma) {
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < count; e += (long long)gridDim.x * 256) {
        const cplx f = ldcs(in + e);
        const double f2 = f.x * f.x + f.y * f.y;
        stg(out + e, mk(-(gamma * (f2 * f.y)), gamma * (f2 * f.x)));          // i gamma |f|^2 f
    }
}
__global__ void __launch_bounds__(256) pointwise_cubic_kernel(const double* in, double* out, long long count, double c) {
    const long long pairs = count >> 1;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < pairs; e += (long long)gridDim.x * 256) {
        const double2 u = __ldcs(reinterpret_cast<const double2*>(in) + e);
        reinterpret_cast<double2*>(out)[e] = make_double2(c * (u.x * u.x * u.x), c * (u.y * u.y * u.y));
    }
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double u = in[count - 1];
        out[count - 1] = c * (u * u * u);
    }
}

// snapshot of the accepted state into the ring (solveras.py:643-645)
__global__ void __launch_bounds__(256) snapshot_kernel(DevPlan p, cplx* ring, double* ring_t, int cap) {
    const Ctrl* c = p.ctrl;
    if (!c->snap_pending) return;
    const long long total = p.batch * p.n_c;
    const int slot = (c->snap_count - 1) % cap;
    const cplx* src = p.U[c->u_sel];
    cplx* dst = ring + (size_t)slot * total;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256)
        stg(dst + e, ldcs(src + e));
    if (blockIdx.x == 0 && threadIdx.x == 0) ring_t[slot] = c->t;
}

// copy a caller array into / out of the role-selected state buffer
__global__ void __launch_bounds__(256) copy_u_kernel(DevPlan p, cplx* ext, int to_plan) {
    const int u_sel = method_adaptive(p.method) ? p.ctrl->u_sel : 0;
    cplx* mine = p.U[u_sel];
    const long long total = p.batch * p.n_c;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < total; e += (long long)gridDim.x * 256) {
        if (to_plan) stg(mine + e, ldcs(ext + e));
        else stg(ext + e, ldcs(mine + e));
    }
}

// row z of a caller array (nplans, n_c) <-> the state buffer of plan z
__global__ void __launch_bounds__(256) copy_u_multi_kernel(const DevPlan* plans, cplx* ext, int to_plan) {
    const DevPlan& p = plans[blockIdx.z];
    cplx* mine = p.U[p.ctrl->u_sel];
    cplx* row = ext + (size_t)blockIdx.z * p.n_c;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < p.n_c; e += (long long)gridDim.x * 256) {
        if (to_plan) stg(mine + e, ldcs(row + e));
        else stg(row + e, ldcs(mine + e));
    }
}

// ---------------------------------------------------------------------------------------
// K4, N-D grids: transform along a strided axis of [outer][N][inner] (fft_axis.cuh).  Persistent
// CTAs loop over (outer, column tile) pairs.  in == out is allowed: a CTA reads its whole tile
// before it writes any of it.
// ---------------------------------------------------------------------------------------
// otab / o_shift: scattered output (fft_axis.cuh Col): destination blocks are laid out [outer][1 << o_shift][inner]
template <int N, bool INV>
RKS_D void axis_fft_body(const cplx* in, cplx* out, long long outer, long long inner, const cplx* tw, double scale,
                         long long ostride, long long bstride, int rb_shift, const long long* otab = nullptr,
                         int o_shift = 31) {
    constexpr int C = axis::tile_cols<N>(), NBT = axis::tile_threads<N>() / C;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(smem_raw);
    const int col = threadIdx.x % C, bt = threadIdx.x / C;
    const long long tpo = (inner + C - 1) / C, tiles = outer * tpo;
    // N = 4096 (one 128 KB tile per SM): the first SROWS input rows of the NEXT tile are copied into the spare shared
    // memory by cp.async while this tile's levels and stores run (each thread copies and waits for its own chunks; the
    // barrier at the end of the iteration publishes them), so the first level reads 3/4 of its inputs from shared memory
    constexpr int SROWS = axis::stage_rows<N>(), SH = axis::tile_shift<N>();
    cplx* stg = tile + (size_t)N * C;
    auto stage_tile = [&](long long t) {
        if (SROWS == 0 || t >= tiles) return;
        const long long o = t / tpo, c0 = (t - o * tpo) * C;
        const cplx* base = in + o * ostride + c0;
        for (int q = threadIdx.x; q < SROWS * C; q += axis::tile_threads<N>()) {
            const int r = q / C, cc = q - r * C;
            if (c0 + cc < inner) {
                const long long roff = (long long)(r >> rb_shift) * bstride + (long long)(r & (int)((1u << rb_shift) - 1u)) * inner;
                cp_async16(stg + fast::swz<SH>(r) * C + cc, base + cc + roff);
            }
        }
        cp_async_commit();
    };
    if (SROWS > 0) {
        stage_tile(blockIdx.x);
        cp_async_wait<0>();
        __syncthreads();
    }
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const long long o = t / tpo, c0 = (t - o * tpo) * C;
        const long long off = o * ostride + c0 + col;
        const axis::Col c{in + off, out + off, inner, bstride, rb_shift, col, c0 + col < inner,
                          otab, o_shift, otab ? o * ((long long)inner << o_shift) + c0 + col : 0, SROWS > 0 ? stg : nullptr};
        axis::tile_level<N, INV, 0>(tile, tw, c, bt, NBT, scale);
        __syncthreads();
        stage_tile(t + gridDim.x);                   // the staged rows of this tile have all been read
        axis::tile_level<N, INV, 1>(tile, tw, c, bt, NBT, scale);
        __syncthreads();
        axis::tile_level<N, INV, 2>(tile, tw, c, bt, NBT, scale);
        if (SROWS > 0) cp_async_wait<0>();
        __syncthreads();
    }
}
template <int N, bool INV>
__global__ void __launch_bounds__(axis::tile_threads<N>(), axis::tile_blocks<N>())
axis_fft_kernel(const cplx* in, cplx* out, long long outer, long long inner, const cplx* tw, double scale,
                long long ostride, long long bstride, int rb_shift) {
    axis_fft_body<N, INV>(in, out, outer, inner, tw, scale, ostride, bstride, rb_shift);
}
// the same transform with its output rows scattered over the ranks of a slab decomposition (peer memory)
template <int N, bool INV>
__global__ void __launch_bounds__(axis::tile_threads<N>(), axis::tile_blocks<N>())
axis_fft_scatter_kernel(const cplx* in, long long outer, long long inner, const cplx* tw, double scale,
                        long long ostride, long long bstride, int rb_shift, const long long* otab, int o_shift) {
    axis_fft_body<N, INV>(in, nullptr, outer, inner, tw, scale, ostride, bstride, rb_shift, otab, o_shift);
}
// barrier between ranks that write into each other's memory (rks_peer_barrier): thread g signals rank g, then waits
// for rank g's signal.  Stream order makes the earlier kernels' stores happen-before the release.
__global__ void peer_barrier_kernel(const long long* flag_bases, int world, int rank, unsigned long long epoch) {
    const int g = threadIdx.x;
    if (g >= world) return;
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flag_bases[g]) + rank;
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(flag_bases[rank]) + g;
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(epoch) : "memory");
    unsigned long long seen;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    } while (seen < epoch);
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}
// the same transform as one step of the nonlinear term N_j of an N-D grid model (rks_set_model_nd): arrays and
// the run predicate come from the control block (no host sync, graph replay); the first step of an evaluation
// reads the stage value and writes N_j, the others work on N_j in place
template <int N, bool INV>
__global__ void __launch_bounds__(axis::tile_threads<N>(), axis::tile_blocks<N>())
axis_fft_plan_kernel(const __grid_constant__ DevPlan p, int j, int force, int first, long long outer, long long inner,
                     const cplx* tw, double scale) {
    const NlRoles r = nl_roles(p, j, force);
    if (!r.run) return;
    axis_fft_body<N, INV>(first ? r.in : r.out, r.out, outer, inner, tw, scale, (long long)N * inner, 0, 31);
}

}  // namespace rks
