// CPU harness for the __host__ __device__ math of rkstiff_b200/csrc (tests only).
// It compiles the SAME headers the CUDA kernels use with g++ and drives the per-thread
// phase functions serially (one "thread" at a time, a barrier being the end of the loop),
// so the device algorithms (FFT passes, psi/tableau formulas, stage combinations,
// controller state machine) can be checked against the oracle without a GPU.
// This is not a product path: nothing in rkstiff_b200/ loads it.
#include <stdint.h>
#include <string.h>
#include <array>
#include <vector>

#include "../../rkstiff_b200/csrc/common.cuh"
#include "../../rkstiff_b200/csrc/coeffs.cuh"
#include "../../rkstiff_b200/csrc/stages.cuh"
#include "../../rkstiff_b200/csrc/errctl.cuh"
#include "../../rkstiff_b200/csrc/fft.cuh"
#include "../../rkstiff_b200/csrc/fft_fast.cuh"
#include "../../rkstiff_b200/csrc/fft_real.cuh"
#include "../../rkstiff_b200/csrc/fft_pair.cuh"
#include "../../rkstiff_b200/csrc/fft_axis.cuh"

using namespace rks;

template <int M, int S>
static void stage_all(int n, const cplx* u, const cplx* const* N, const cplx* coef, double h, cplx* out, cplx* err) {
    constexpr int NC = method_ncoef(M);
    for (int i = 0; i < n; ++i) {
        cplx nv[8], cv[NC];
        for (int j = 1; j <= 7; ++j) nv[j] = N[j] ? N[j][i] : mk(0, 0);
        for (int s = 0; s < NC; ++s) cv[s] = coef[s * n + i];
        out[i] = stage_combine<M, S, cplx>(u[i], nv, cv, h);
        if (M == M_ETD35 && S == 6 && err) err[i] = etd35_err<cplx>(nv, cv);
    }
}


// serial emulation of the fast NL kernel (fft_fast.cuh): one row, 32 W threads, phase by phase
template <int N, class Model>
static void fast_row(const Model& m, const fast::Twiddles& tw) {
    constexpr int TR = 32 * fast::Plan<N>::W;
    std::vector<cplx> sm(N);
    for (int T = 0; T < TR; ++T) fast::phase_first<N>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, true>(sm.data(), T, tw, m);
    if (fast::middle_passes<N>() == 2)
        for (int T = 0; T < TR; ++T) fast::phase_middle<N, 3, true>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_core<N>(sm.data(), T, m);
    if (fast::middle_passes<N>() == 2)
        for (int T = 0; T < TR; ++T) fast::phase_middle<N, 3, false>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, false>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_last<N>(sm.data(), T, tw, m);
}
// serial emulation of the pre-transformed route: stage_pre_kernel's butterflies on the row `k` (the stage value),
// then nl_fast_pre_kernel's phase sequence on the result
template <int N>
static void pre_row(const cplx* k, cplx* out, double gamma, const fast::Twiddles& tw) {
    using P = fast::Plan<N>;
    constexpr int TR = 32 * P::W, R1 = P::R1, Q1 = N / R1;
    std::vector<cplx> kt(N), sm(N);
    for (int col = 0; col < Q1; ++col) {                    // one K1 thread per first-pass butterfly
        cplx a[R1];
        for (int s = 0; s < R1; ++s) a[s] = k[col + s * Q1];
        fast::pre_butterfly<R1>(a, tw.t1, col);
        for (int r = 0; r < R1; ++r) kt[col + r * Q1] = a[fast::perm<R1>(r)];
    }
    const auto m = fast::ModelOf<2>::make(kt.data(), out, nullptr, gamma, N, true);
    for (int T = 0; T < TR; ++T) fast::phase_pre<N>(sm.data(), T, tw, m);
    if (fast::middle_passes<N>() == 2)
        for (int T = 0; T < TR; ++T) fast::phase_middle<N, 3, true>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_core<N>(sm.data(), T, m);
    if (fast::middle_passes<N>() == 2)
        for (int T = 0; T < TR; ++T) fast::phase_middle<N, 3, false>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, false>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_last<N>(sm.data(), T, tw, m);
}


// serial emulation of the real-field variant (fft_real.cuh): inverse passes as in fast_row, then the paired core
// pass, the forward middle pass on the even blocks and the half-length last pass with its exchange
template <int N, class Model>
static void fast_row_real(const Model& m, const fast::Twiddles& tw) {
    using P = fast::Plan<N>;
    constexpr int TR = 32 * P::W, H = P::R1 / 2, NB = (N / P::R1) / TR;
    std::vector<cplx> sm(N);
    for (int T = 0; T < TR; ++T) fast::phase_first<N>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, true>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_core_pair<N>(sm.data(), T, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle_even<N>(sm.data(), T, tw);
    std::vector<cplx> regs((size_t)TR * NB * H);
    for (int T = 0; T < TR; ++T) fast::phase_last_half_load<N>(sm.data(), T, tw, regs.data() + (size_t)T * NB * H);
    for (int T = 0; T < TR; ++T) fast::phase_last_half_exchange<N>(sm.data(), T, regs.data() + (size_t)T * NB * H);
    for (int T = 0; T < TR; ++T) fast::phase_last_half_split<N>(sm.data(), T, tw, regs.data() + (size_t)T * NB * H, m);
}
template <class Model>
static int fast_real_dispatch(int n, const Model& m, const fast::Twiddles& tw) {
    switch (n) {
        case 512: fast_row_real<512>(m, tw); return 0;
        case 1024: fast_row_real<1024>(m, tw); return 0;
        case 2048: fast_row_real<2048>(m, tw); return 0;
        case 4096: fast_row_real<4096>(m, tw); return 0;
    }
    return -1;
}

// serial emulation of the paired-row kernel (fft_pair.cuh): the unchanged passes on the packed row, then the three
// steps of the last pass (load / publish / store), each run for every thread before the next (the row barriers)
template <int N, class Model>
static void fast_row_pair(const Model& m, const fast::Twiddles& tw) {
    using P = fast::Plan<N>;
    constexpr int TR = 32 * P::W, NB = (N / P::R1) / TR, PER = NB * P::R1;
    std::vector<cplx> sm(N), regs((size_t)TR * PER);
    for (int T = 0; T < TR; ++T) fast::phase_first<N>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, true>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_core<N>(sm.data(), T, m);
    for (int T = 0; T < TR; ++T) fast::phase_middle<N, 2, false>(sm.data(), T, tw, m);
    for (int T = 0; T < TR; ++T) fast::phase_pair_load<N>(sm.data(), T, regs.data() + (size_t)T * PER);
    for (int T = 0; T < TR; ++T) fast::phase_pair_publish<N>(sm.data(), T, tw, regs.data() + (size_t)T * PER);
    for (int T = 0; T < TR; ++T) fast::phase_pair_store<N>(sm.data(), T, regs.data() + (size_t)T * PER, m);
}

template <class Model>
static int fast_dispatch(int n, const Model& m, const fast::Twiddles& tw) {
    switch (n) {
        case 512: fast_row<512>(m, tw); return 0;
        case 1024: fast_row<1024>(m, tw); return 0;
        case 2048: fast_row<2048>(m, tw); return 0;
        case 4096: fast_row<4096>(m, tw); return 0;
        case 8192: fast_row<8192>(m, tw); return 0;
    }
    return -1;
}

// serial emulation of the packed short-row kernel (nl_small_kernel): one warp, 512 / N rows per slab
template <int N, int MODEL>
static void packed_rows(int rows, long long n_c, const cplx* in, cplx* out, const double* kx, double p0,
                        const fast::Twiddles& tw) {
    constexpr int SUB = 512 / N;
    std::vector<cplx> sm(512);
    for (int row0 = 0; row0 < rows; row0 += SUB) {
        const int left = rows - row0;
        const fast::PackedModel<MODEL, N> m{in + row0 * n_c, out + row0 * n_c, kx, p0, n_c, left < SUB ? left : SUB};
        for (int T = 0; T < 32; ++T) fast::phase_first_packed<N>(sm.data(), T, tw, m);
        for (int T = 0; T < 32; ++T) fast::phase_middle<N, 2, true>(sm.data(), T, tw, m);
        for (int T = 0; T < 32; ++T) fast::phase_core<N>(sm.data(), T, m);
        for (int T = 0; T < 32; ++T) fast::phase_middle<N, 2, false>(sm.data(), T, tw, m);
        for (int T = 0; T < 32; ++T) fast::phase_last_packed<N>(sm.data(), T, tw, m);
    }
}
template <int MODEL>
static int packed_dispatch(int n, int rows, long long n_c, const cplx* in, cplx* out, const double* kx, double p0,
                           const fast::Twiddles& tw) {
    switch (n) {
        case 64: packed_rows<64, MODEL>(rows, n_c, in, out, kx, p0, tw); return 0;
        case 128: packed_rows<128, MODEL>(rows, n_c, in, out, kx, p0, tw); return 0;
        case 256: packed_rows<256, MODEL>(rows, n_c, in, out, kx, p0, tw); return 0;
    }
    return -1;
}

// fused NL of one row through the generic radix-4 passes (fft.cuh), emulating `nthreads` threads
template <class Model>
static void generic_row(const Model& m, int n, int nthreads) {
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    std::vector<cplx> tw(n), x(n);
    for (int j = 0; j < n; ++j) tw[j] = mk(cos(-2.0 * M_PI * j / n), sin(-2.0 * M_PI * j / n));
    for (int t = 0; t < nthreads; ++t)
        for (int q = t; q < n; q += nthreads) x[q] = m.load(q);
    const int np = fft_num_passes(log2n);
    for (int q = 0; q < np; ++q)
        for (int t = 0; t < nthreads; ++t) ifft_dif_pass(x.data(), log2n, q, tw.data(), t, nthreads);
    for (int q = 0; q < n; ++q) x[q] = fast::pointwise_of(m, x[q], q);
    for (int q = 0; q < np; ++q)
        for (int t = 0; t < nthreads; ++t) fft_dit_pass(x.data(), log2n, q, tw.data(), t, nthreads);
    for (int q = 0; q < n; ++q) m.store(q, x[q]);
}
// strided-axis transform of fft_axis.cuh, emulated thread by thread: every level is run for all
// threads of the CTA before the next one (that is what the __syncthreads between levels gives)
template <int N, bool INV, int LEVEL>
static void axis_level_all(cplx* tile, const cplx* tw, const cplx* in, cplx* out, long long inner, long long c0) {
    constexpr int C = axis::tile_cols<N>(), T = axis::tile_threads<N>(), NBT = T / C;
    for (int tid = 0; tid < T; ++tid) {
        const int col = tid % C, bt = tid / C;
        const axis::Col c{in + c0 + col, out + c0 + col, inner, 0, 31, col, c0 + col < inner};
        axis::tile_level<N, INV, LEVEL>(tile, tw, c, bt, NBT, INV ? 1.0 / (double)N : 1.0);
    }
}
template <int N, bool INV>
static void axis_all(const cplx* in, cplx* out, long long inner) {
    constexpr int C = axis::tile_cols<N>();
    std::vector<cplx> tw(N), tile((size_t)N * C);
    for (int j = 0; j < N; ++j) tw[j] = mk(cos(-2.0 * M_PI * j / N), sin(-2.0 * M_PI * j / N));
    for (long long c0 = 0; c0 < inner; c0 += C) {
        axis_level_all<N, INV, 0>(tile.data(), tw.data(), in, out, inner, c0);
        axis_level_all<N, INV, 1>(tile.data(), tw.data(), in, out, inner, c0);
        axis_level_all<N, INV, 2>(tile.data(), tw.data(), in, out, inner, c0);
    }
}
template <bool INV>
static int axis_dispatch(int n, const cplx* in, cplx* out, long long inner) {
    switch (n) {
        case 16: axis_all<16, INV>(in, out, inner); return 0;
        case 32: axis_all<32, INV>(in, out, inner); return 0;
        case 64: axis_all<64, INV>(in, out, inner); return 0;
        case 128: axis_all<128, INV>(in, out, inner); return 0;
        case 256: axis_all<256, INV>(in, out, inner); return 0;
        case 512: axis_all<512, INV>(in, out, inner); return 0;
        case 1024: axis_all<1024, INV>(in, out, inner); return 0;
        case 2048: axis_all<2048, INV>(in, out, inner); return 0;
        case 4096: axis_all<4096, INV>(in, out, inner); return 0;
        default: return -1;
    }
}

extern "C" {

int hc_nl_fast(int model, int n, const double* in, const double* kx, double p0, double* out) {
    std::vector<cplx> tab(fast::TW_TOTAL);
    for (int j = 0; j < fast::TW_TOTAL; ++j) tab[j] = fast::twiddle_table_entry(j, n);
    const fast::Twiddles tw{tab.data() + fast::TW_T1, tab.data() + fast::TW_T2, tab.data() + fast::TW_T3};
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* co = reinterpret_cast<cplx*>(out);
    if (model == 1) return fast_dispatch(n, fast::ModelOf<1>::make(cin, co, kx, p0, n, true), tw);
    if (model == 2) return fast_dispatch(n, fast::ModelOf<2>::make(cin, co, kx, p0, n, true), tw);
    if (model == 3) return fast_dispatch(n, fast::ModelOf<3>::make(cin, co, kx, p0, n, true), tw);
    if (model == 5) return fast_dispatch(n, fast::ModelOf<5>::make(cin, co, kx, 0.0, n, true), tw);    // derivative rows
    if (model == 6) return fast_dispatch(n, fast::ModelOf<6>::make(cin, co, kx, 0.0, n, true), tw);
    return fast_dispatch(n, fast::ModelOf<4>::make(cin, co, kx, p0, n, true), tw);
}

// real-field models (1 = u u_x, 3 = cubic) through the half-length forward transform of fft_real.cuh
int hc_nl_fast_real(int model, int n, const double* in, const double* kx, double p0, double* out) {
    std::vector<cplx> tab(fast::TW_TOTAL);
    for (int j = 0; j < fast::TW_TOTAL; ++j) tab[j] = fast::twiddle_table_entry(j, n);
    const fast::Twiddles tw{tab.data() + fast::TW_T1, tab.data() + fast::TW_T2, tab.data() + fast::TW_T3};
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* co = reinterpret_cast<cplx*>(out);
    if (model == 1) return fast_real_dispatch(n, fast::ModelOf<1>::make(cin, co, kx, p0, n, true), tw);
    if (model == 3) return fast_real_dispatch(n, fast::ModelOf<3>::make(cin, co, kx, p0, n, true), tw);
    return -1;
}

// cubic model on a row pair (in_b may be null: odd tail, the second row is zeros and nothing is stored for it)
int hc_nl_fast_pair(int n, const double* in_a, const double* in_b, double c, double* out_a, double* out_b) {
    std::vector<cplx> tab(fast::TW_TOTAL);
    for (int j = 0; j < fast::TW_TOTAL; ++j) tab[j] = fast::twiddle_table_entry(j, n);
    const fast::Twiddles tw{tab.data() + fast::TW_T1, tab.data() + fast::TW_T2, tab.data() + fast::TW_T3};
    const fast::PairedCubicModel m{reinterpret_cast<const cplx*>(in_a), reinterpret_cast<const cplx*>(in_b ? in_b : in_a),
                                   reinterpret_cast<cplx*>(out_a), reinterpret_cast<cplx*>(out_b), c, n, true, in_b != nullptr};
    switch (n) {
        case 512: fast_row_pair<512>(m, tw); return 0;
        case 1024: fast_row_pair<1024>(m, tw); return 0;
        case 2048: fast_row_pair<2048>(m, tw); return 0;
        case 4096: fast_row_pair<4096>(m, tw); return 0;
    }
    return -1;
}

// x / d through div_const<d> (common.cuh) for the divisors the kernels use
int hc_div_const(int d, int n, const double* x, double* out) {
    for (int i = 0; i < n; ++i) {
        switch (d) {
            case 3: out[i] = div_const<3>(x[i]); break;
            case 6: out[i] = div_const<6>(x[i]); break;
            case 40: out[i] = div_const<40>(x[i]); break;
            case 84: out[i] = div_const<84>(x[i]); break;
            case 525: out[i] = div_const<525>(x[i]); break;
            default: return -1;
        }
    }
    return 0;
}

// NLS evaluation of one row through the pre-transformed route (K1 applies the first inverse pass)
int hc_nl_pre(int n, const double* in, double gamma, double* out) {
    std::vector<cplx> tab(fast::TW_TOTAL);
    for (int j = 0; j < fast::TW_TOTAL; ++j) tab[j] = fast::twiddle_table_entry(j, n);
    const fast::Twiddles tw{tab.data() + fast::TW_T1, tab.data() + fast::TW_T2, tab.data() + fast::TW_T3};
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* co = reinterpret_cast<cplx*>(out);
    switch (n) {
        case 512: pre_row<512>(cin, co, gamma, tw); return 0;
        case 1024: pre_row<1024>(cin, co, gamma, tw); return 0;
        case 2048: pre_row<2048>(cin, co, gamma, tw); return 0;
        case 4096: pre_row<4096>(cin, co, gamma, tw); return 0;
        case 8192: pre_row<8192>(cin, co, gamma, tw); return 0;
    }
    return -1;
}

// `rows` consecutive rows of n_c elements through the packed short-row pipeline (n = 64, 128, 256)
int hc_nl_packed(int model, int n, int rows, const double* in, const double* kx, double p0, double* out) {
    std::vector<cplx> tab(fast::TW_TOTAL);
    for (int j = 0; j < fast::TW_TOTAL; ++j) tab[j] = fast::twiddle_table_entry(j, n);
    const fast::Twiddles tw{tab.data() + fast::TW_T1, tab.data() + fast::TW_T2, tab.data() + fast::TW_T3};
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* co = reinterpret_cast<cplx*>(out);
    const long long n_c = (model == 1 || model == 3) ? n / 2 + 1 : n;
    if (model == 1) return packed_dispatch<1>(n, rows, n_c, cin, co, kx, p0, tw);
    if (model == 2) return packed_dispatch<2>(n, rows, n_c, cin, co, kx, p0, tw);
    if (model == 3) return packed_dispatch<3>(n, rows, n_c, cin, co, kx, p0, tw);
    if (model == 5) return packed_dispatch<5>(n, rows, n_c, cin, co, kx, 0.0, tw);
    if (model == 6) return packed_dispatch<6>(n, rows, n_c, cin, co, kx, 0.0, tw);      // rows = row PAIRS
    return packed_dispatch<4>(n, rows, n_c, cin, co, kx, p0, tw);
}

void hc_nl(int model, int n, const double* in, const double* kx, double p0, double* out, int nthreads) {
    const cplx* cin = reinterpret_cast<const cplx*>(in);
    cplx* co = reinterpret_cast<cplx*>(out);
    if (model == 1) generic_row(fast::ModelOf<1>::make(cin, co, kx, p0, n, true), n, nthreads);
    else if (model == 2) generic_row(fast::ModelOf<2>::make(cin, co, kx, p0, n, true), n, nthreads);
    else if (model == 3) generic_row(fast::ModelOf<3>::make(cin, co, kx, p0, n, true), n, nthreads);
    else if (model == 5) generic_row(fast::ModelOf<5>::make(cin, co, kx, 1.0, n, true), n, nthreads);   // p0 != 0: generic digit order
    else if (model == 6) generic_row(fast::ModelOf<6>::make(cin, co, kx, 1.0, n, true), n, nthreads);
    else generic_row(fast::ModelOf<4>::make(cin, co, kx, p0, n, true), n, nthreads);
}

// inverse-DIF followed by forward-DIT must be n * identity; also exposes the raw transforms
void hc_fft_roundtrip(int n, const double* in, double* out) {
    int log2n = 0;
    while ((1 << log2n) < n) ++log2n;
    std::vector<cplx> tw(n), x(n);
    for (int j = 0; j < n; ++j) tw[j] = mk(cos(-2.0 * M_PI * j / n), sin(-2.0 * M_PI * j / n));
    memcpy(x.data(), in, sizeof(cplx) * n);
    const int np = fft_num_passes(log2n);
    for (int q = 0; q < np; ++q) ifft_dif_pass(x.data(), log2n, q, tw.data(), 0, 1);
    for (int q = 0; q < np; ++q) fft_dit_pass(x.data(), log2n, q, tw.data(), 0, 1);
    memcpy(out, x.data(), sizeof(cplx) * n);
}

// coefficient arrays of one method: out is [ncoef][n] complex (IF real coefficients are widened)
int hc_coeffs(int method, int n, const double* lin, int lin_complex, double h, double modecutoff,
              int contour_points, double contour_radius, int r4_fix, double* out) {
    cplx* o = reinterpret_cast<cplx*>(out);
    const int nc = method_ncoef(method);
    for (int i = 0; i < n; ++i) {
        const cplx L = lin_complex ? mk(lin[2 * i], lin[2 * i + 1]) : mk(lin[i], 0.0);
        if (method == M_IF4 || method == M_IF34) {
            if (lin_complex) {
                const cplx z = scale(h, L);
                o[ifc::E * n + i] = cexp_t(z);
                o[ifc::E2 * n + i] = cexp_t(z / 2.0);
            } else {
                const double z = h * L.x;
                o[ifc::E * n + i] = mk(exp(z), 0.0);
                o[ifc::E2 * n + i] = mk(exp(z / 2.0), 0.0);
            }
        } else if (method == M_IF45DP) {
            if (lin_complex) {
                cplx arr[dp::COUNT];
                tableau_if45dp<cplx>(scale(h, L), h, r4_fix, arr);
                for (int s = 0; s < nc; ++s) o[s * n + i] = arr[s];
            } else {
                double arr[dp::COUNT];
                tableau_if45dp<double>(h * L.x, h, r4_fix, arr);
                for (int s = 0; s < nc; ++s) o[s * n + i] = mk(arr[s], 0.0);
            }
        } else {
            const bool five = (method == M_ETD5 || method == M_ETD35);
            const cplx z = lin_complex ? scale(h, L) : mk(h * L.x, 0.0);
            PsiSet ps = psi_zero();
            if (hypot(z.x, z.y) < modecutoff) {
                for (int j = 0; j < contour_points; ++j) {
                    const cplx w = z + contour_node(contour_radius, j, contour_points);
                    if (five) psi_accumulate<true>(ps, w); else psi_accumulate<false>(ps, w);
                }
                psi_scale(ps, h, (double)contour_points);
            } else {
                if (five) psi_accumulate<true>(ps, z); else psi_accumulate<false>(ps, z);
                psi_scale(ps, h, 1.0);
            }
            if (five) {
                cplx arr[e5::COUNT];
                const ExpSet ez = exp_set<true>(z);
                arr[e5::E14] = ez.q; arr[e5::E12] = ez.h; arr[e5::E34] = ez.t; arr[e5::E] = ez.f;
                tableau_etd5(ps, arr);
                for (int s = 0; s < nc; ++s) o[s * n + i] = arr[s];
            } else {
                cplx arr[kro::COUNT];
                const ExpSet ez = exp_set<false>(z);
                arr[kro::E] = ez.f; arr[kro::E2] = ez.h;
                tableau_krogstad(ps, arr);
                for (int s = 0; s < nc; ++s) o[s * n + i] = arr[s];
            }
        }
    }
    return nc;
}

// CM_SEPARABLE: coefficient slots of an IF method on an (n0, n1) grid from the per-axis exponential tables,
// exactly as K2 (coef_sep_kernel) and K1 (sep_coefs) form them; out is [ncoef][n0*n1] complex
int hc_sep_coeffs(int method, int n0, int n1, const double* a0, const double* a1, int lin_complex, double h, int r4_fix,
                  double* out) {
    if (!method_is_if(method)) return -1;
    cplx* o = reinterpret_cast<cplx*>(out);
    const int nc = method_ncoef(method), nq = sep_nq(method);
    const long long n = (long long)n0 * n1;
    std::vector<cplx> t0((size_t)nq * n0), t1((size_t)nq * n1);
    for (int q = 0; q < nq; ++q) {
        for (int i = 0; i < n0; ++i)
            t0[(size_t)q * n0 + i] = lin_complex ? sep_exp<cplx>(method, q, scale(h, mk(a0[2 * i], a0[2 * i + 1])))
                                                 : mk(sep_exp<double>(method, q, h * a0[i]), 0.0);
        for (int j = 0; j < n1; ++j)
            t1[(size_t)q * n1 + j] = lin_complex ? sep_exp<cplx>(method, q, scale(h, mk(a1[2 * j], a1[2 * j + 1])))
                                                 : mk(sep_exp<double>(method, q, h * a1[j]), 0.0);
    }
    for (int s = 0; s < nc; ++s) {
        const int q = sep_q(method, s);
        const double sc = sep_scale(method, s, h, r4_fix);
        const bool pure = method != M_IF45DP || s <= dp::E;
        for (int i = 0; i < n0; ++i)
            for (int j = 0; j < n1; ++j) {
                const cplx e = t0[(size_t)q * n0 + i] * t1[(size_t)q * n1 + j];
                o[(long long)s * n + (long long)i * n1 + j] = pure ? e : scale(sc, e);
            }
    }
    return nc;
}

// CM_INDEXED record layout of a method: out[g] = offset of group g = 1..S+1, out[0] = record length (CT elements);
// masks[g] = the slots the group holds
int hc_record_layout(int method, int real_coef, int* out, unsigned* masks) {
    const int S = method_stages(method);
#define HC_LAYOUT(M) if (method == M) { \
        out[0] = real_coef ? record_elems<double>(M) : record_elems<cplx>(M); \
        for (int g = 1; g <= S + 1; ++g) { out[g] = real_coef ? group_off<double>(M, g) : group_off<cplx>(M, g); masks[g] = group_mask(M, g); } \
        return S + 1; }
    HC_LAYOUT(M_IF4) HC_LAYOUT(M_ETD4) HC_LAYOUT(M_ETD5) HC_LAYOUT(M_IF34) HC_LAYOUT(M_ETD34) HC_LAYOUT(M_ETD35) HC_LAYOUT(M_IF45DP)
#undef HC_LAYOUT
    return -1;
}

#define HC_STAGE(M, S) if (method == M && stage == S) { stage_all<M, S>(n, cu, Np, cc, h, co, ce); return 0; }
// one stage combine with complex coefficients; N is 8 pointers (index 1..7, may be null)
int hc_stage(int method, int stage, int n, const double* u, const double* const* N, const double* coef, double h,
             double* out, double* err) {
    const cplx* cu = reinterpret_cast<const cplx*>(u);
    const cplx* Np[8];
    for (int j = 0; j < 8; ++j) Np[j] = reinterpret_cast<const cplx*>(N[j]);
    const cplx* cc = reinterpret_cast<const cplx*>(coef);
    cplx* co = reinterpret_cast<cplx*>(out);
    cplx* ce = reinterpret_cast<cplx*>(err);
    HC_STAGE(M_IF4, 1) HC_STAGE(M_IF4, 2) HC_STAGE(M_IF4, 3) HC_STAGE(M_IF4, 4)
    HC_STAGE(M_IF34, 1) HC_STAGE(M_IF34, 2) HC_STAGE(M_IF34, 3) HC_STAGE(M_IF34, 4)
    HC_STAGE(M_ETD4, 1) HC_STAGE(M_ETD4, 2) HC_STAGE(M_ETD4, 3) HC_STAGE(M_ETD4, 4)
    HC_STAGE(M_ETD34, 1) HC_STAGE(M_ETD34, 2) HC_STAGE(M_ETD34, 3) HC_STAGE(M_ETD34, 4)
    HC_STAGE(M_ETD5, 1) HC_STAGE(M_ETD5, 2) HC_STAGE(M_ETD5, 3) HC_STAGE(M_ETD5, 4) HC_STAGE(M_ETD5, 5) HC_STAGE(M_ETD5, 6)
    HC_STAGE(M_ETD35, 1) HC_STAGE(M_ETD35, 2) HC_STAGE(M_ETD35, 3) HC_STAGE(M_ETD35, 4) HC_STAGE(M_ETD35, 5) HC_STAGE(M_ETD35, 6)
    HC_STAGE(M_IF45DP, 1) HC_STAGE(M_IF45DP, 2) HC_STAGE(M_IF45DP, 3) HC_STAGE(M_IF45DP, 4) HC_STAGE(M_IF45DP, 5) HC_STAGE(M_IF45DP, 6)
    return -1;
}

// embedded error estimate formed inside the norm kernel (IF34 / ETD34 / IF45DP)
int hc_embedded_err(int method, int n, const double* const* N, const double* coef, double h, double* err) {
    const cplx* Np[8];
    for (int j = 0; j < 8; ++j) Np[j] = reinterpret_cast<const cplx*>(N[j]);
    const cplx* cc = reinterpret_cast<const cplx*>(coef);
    cplx* ce = reinterpret_cast<cplx*>(err);
    for (int i = 0; i < n; ++i) {
        cplx nv[8], cv[32];
        for (int j = 1; j <= 7; ++j) nv[j] = Np[j] ? Np[j][i] : mk(0, 0);
        for (int s = 0; s < method_ncoef(method); ++s) cv[s] = cc[s * n + i];
        if (method == M_IF34) ce[i] = embedded_err<M_IF34, cplx>(nv, cv, h);
        else if (method == M_ETD34) ce[i] = embedded_err<M_ETD34, cplx>(nv, cv, h);
        else if (method == M_IF45DP) ce[i] = embedded_err<M_IF45DP, cplx>(nv, cv, h);
        else return -1;
    }
    return 0;
}

// controller: feed (sum_u2, sum_e2) of one trial; state is a caller-held opaque Ctrl blob
// in / out: [n][inner] complex128 (one `outer` slice); inverse: natural -> digit-reversed rows
int hc_axis_fft(int n, int inverse, const double* in, double* out, long long inner) {
    return inverse ? axis_dispatch<true>(n, (const cplx*)in, (cplx*)out, inner)
                   : axis_dispatch<false>(n, (const cplx*)in, (cplx*)out, inner);
}

int hc_ctrl_size(void) { return (int)sizeof(Ctrl); }
void hc_ctrl_init(void* blob, double t0, double tf, double h, long long store_freq, int step_mode, int n1_refresh,
                  double epsilon, double incr_f, double decr_f, double safety_f, double minh, int q) {
    Ctrl* c = reinterpret_cast<Ctrl*>(blob);
    memset(c, 0, sizeof(Ctrl));
    c->t = t0; c->tf = tf; c->h = h; c->h_last = h; c->store_freq = store_freq; c->step_mode = step_mode;
    c->n1_refresh = n1_refresh; c->need_n1 = 1;
    c->epsilon = epsilon; c->incr_f = incr_f; c->decr_f = decr_f; c->safety_f = safety_f; c->minh = minh;
    c->inv_q = 1.0 / (double)q;
    c->h_coeff = NAN;
    c->log_cap = LOG_CAP;
}
// out: {h, h_last, t, s_last, status, accept, numloops, u_sel, n_sel, need_n1, step_count, snap_pending, snap_count}
void hc_ctrl_advance(void* blob, double sum_u2, double sum_e2, double* out) {
    static TrialRec log[LOG_CAP];
    Ctrl* c = reinterpret_cast<Ctrl*>(blob);
    c->red[1] = sum_u2; c->red[2] = sum_e2;
    controller_advance(*c, log);
    out[0] = c->h; out[1] = c->h_last; out[2] = c->t; out[3] = c->s_last; out[4] = c->status; out[5] = c->accept;
    out[6] = c->numloops; out[7] = c->u_sel; out[8] = c->n_sel; out[9] = c->need_n1; out[10] = (double)c->step_count;
    out[11] = c->snap_pending; out[12] = c->snap_count;
}

}  // extern "C"
