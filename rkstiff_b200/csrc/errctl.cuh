// K3 tail -- the adaptive controller state machine, run by ONE device thread per trial.
// Restates rkstiff/solveras.py:336-410 (step), 412-455 (_compute_s), 457-506
// (_reject_step_size), 508-554 (_accept_step_size) and the bookkeeping of evolve()
// (solveras.py:614-645) in the same IEEE double operations, so that the accepted/rejected
// dt sequence matches the reference's host floats.  No FMA contraction is allowed here:
// every product/sum is a separately rounded operation (explicit _rn intrinsics on device).
#pragma once
#include "common.cuh"

namespace rks {

constexpr int CTRL_MAX_LOOPS = 50;     // solveras.py:255
constexpr double CTRL_MAX_S = 4.0;     // solveras.py:256
constexpr double CTRL_MIN_S = 0.25;    // solveras.py:257

RKS_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
RKS_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
RKS_HD double sub_rn(double a, double b) { return add_rn(a, -b); }
RKS_HD double div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

// s = safety_f * (epsilon * ||u[idx]|| / ||err[idx]||)^(1/q), solveras.py:453-454
RKS_HD double controller_scale(const Ctrl& c, double sum_u2, double sum_e2) {
    const double tol = mul_rn(c.epsilon, sqrt(sum_u2));
    const double ratio = div_rn(tol, sqrt(sum_e2));
    return mul_rn(c.safety_f, pow(ratio, c.inv_q));
}

// Consume the reduction results of one trial and advance the control block.
RKS_HD void controller_advance(Ctrl& c, TrialRec* log) {
    const double h = c.h;
    const double s = controller_scale(c, c.red[1], c.red[2]);
    c.s_last = s;
    c.h_coeff = h;                       // coefficient arrays now hold this h (etd35.py:851-853)
    c.trial_count += 1;
    c.snap_pending = 0;
    const bool bad = isinf(s) || isnan(s);
    TrialRec rec;
    rec.h = h; rec.s = s; rec.pad = 0;
    if (bad || s < 1.0) {
        // ---- reject: solveras.py:391-392, 492-506
        c.accept = 0;
        double hn;
        if (bad) {
            hn = mul_rn(CTRL_MIN_S, h);
        } else {
            double sc = s > CTRL_MIN_S ? s : CTRL_MIN_S;
            sc = sc < c.decr_f ? sc : c.decr_f;
            hn = mul_rn(sc, h);
        }
        c.h = hn;
        c.numloops += 1;
        if (c.numloops > CTRL_MAX_LOOPS) c.status = ST_MAX_LOOPS;       // checked first, solveras.py:399
        else if (hn < c.minh) c.status = ST_MIN_STEP;
        rec.accepted = 0; rec.t_after = c.t;
    } else {
        // ---- accept: solveras.py:394-397, 539-554
        c.accept = 1;
        c.numloops = 0;
        const double sc = s < CTRL_MAX_S ? s : CTRL_MAX_S;
        const double h_suggest = sc > c.incr_f ? mul_rn(sc, h) : h;
        c.h_last = h;
        c.u_sel ^= 1;                    // candidate buffer becomes the state
        c.n_sel ^= 1;                    // FSAL: N(u+) becomes N1 (ignored by non-FSAL methods)
        c.step_count += 1;
        if (c.step_mode) {
            c.h = h_suggest;
            c.status = ST_DONE;
        } else {
            // evolve(): solveras.py:626-645
            const double tc = add_rn(c.t, h);
            c.t = tc;
            c.h = add_rn(tc, h_suggest) > c.tf ? sub_rn(c.tf, tc) : h_suggest;
            if (c.store_freq > 0 && c.step_count % c.store_freq == 0) {
                c.snap_pending = 1;
                c.snap_count += 1;
            }
            if (!(tc < c.tf)) c.status = ST_DONE;
        }
        rec.accepted = 1; rec.t_after = c.t;
    }
    c.need_n1 = (c.accept && c.n1_refresh) ? 1 : 0;     // ETD35 is not FSAL: etd35.py:317-318
    log[c.log_count & (c.log_cap - 1)] = rec;          // ring capacities are powers of two
    c.log_count += 1;
    // reset the reduction scalars for the next trial
    c.red[0] = 0.0; c.red[1] = 0.0; c.red[2] = 0.0;
    c.ticket = 0u;
}

}  // namespace rks
