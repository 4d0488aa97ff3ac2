// C ABI of the rkstiff_b200 engine: plan management, kernel dispatch, control-block I/O.
// Declared in include/rkstiff_b200.h.  No C++ exception crosses this boundary.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rkstiff_b200.h"
#include "kernels.cuh"

using namespace rks;

static thread_local char g_err[512] = "";
// per-device caches of function attributes / occupancy answers (a process may drive several GPUs)
constexpr int MAX_DEVICES = 64;
static std::mutex g_attr_mutex;

static int fail(int code, const char* fmt, const char* detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
#define CUDA_TRY(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return fail(RKS_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

static_assert(sizeof(rks_trial_rec) == sizeof(TrialRec), "log record layout");
static_assert(RKS_LOG_CAP == LOG_CAP, "log capacity");

// ---------------------------------------------------------------------------------------
// workspace layout
// ---------------------------------------------------------------------------------------
constexpr size_t ALIGN = 256;
constexpr int NORM_MAX_BLOCKS = 4096;
constexpr size_t PRE_CNT_BYTES = 256;       // stage_pre_kernel: {next row, finished warps} per column block (<= 16)
constexpr int MULTI_NORM_BLOCKS = 16;       // norm-kernel blocks per row in independent-dt mode
constexpr int MULTI_LOG_CAP = RKS_ROW_LOG_CAP;          // trial records kept per row in independent-dt mode
constexpr long long MODEL_MAX_N = 16384;        // longest row the smem-resident FFT handles

static size_t align_up(size_t v) { return (v + ALIGN - 1) / ALIGN * ALIGN; }

struct Layout {
    size_t ctrl, log, partials, tw, twf, kx, lin, coef, cidx, cscale, U[2], K, ERR, NL[8], total;
};

// CT elements of one grouped coefficient record (stages.cuh record_elems, runtime method id)
static size_t record_elems_rt(int m, bool real_coef) {
    switch (m) {
        case M_IF4: return real_coef ? record_elems<double>(M_IF4) : record_elems<cplx>(M_IF4);
        case M_IF34: return real_coef ? record_elems<double>(M_IF34) : record_elems<cplx>(M_IF34);
        case M_IF45DP: return real_coef ? record_elems<double>(M_IF45DP) : record_elems<cplx>(M_IF45DP);
        case M_ETD4: return record_elems<cplx>(M_ETD4);
        case M_ETD34: return record_elems<cplx>(M_ETD34);
        case M_ETD5: return record_elems<cplx>(M_ETD5);
        default: return record_elems<cplx>(M_ETD35);
    }
}

// mode CM_COLUMN / CM_FLAT: lin_elems = n_c or batch * n_c coefficient entries per array;
// CM_INDEXED: lin_elems = number of distinct lin_op values; CM_SEPARABLE: lin_elems = sum of the grid dims
static Layout make_layout(int method, long long batch, long long n_c, long long lin_elems, int lin_is_complex,
                          int mode = CM_COLUMN) {
    Layout L;
    memset(&L, 0, sizeof(L));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t state = (size_t)batch * (size_t)n_c * sizeof(cplx);
    const bool is_if = method_is_if(method);
    const size_t coef_elem = (is_if && !lin_is_complex) ? sizeof(double) : sizeof(cplx);
    const size_t coef_per_entry = mode == CM_INDEXED ? record_elems_rt(method, coef_elem == sizeof(double))
                                : mode == CM_SEPARABLE ? (size_t)sep_nq(method) : (size_t)method_ncoef(method);
    L.ctrl = take(sizeof(Ctrl));
    L.log = take(sizeof(TrialRec) * LOG_CAP);
    L.partials = take(sizeof(double) * 2 * NORM_MAX_BLOCKS + PRE_CNT_BYTES);    // + row counters of stage_pre_kernel
    const long long nmax = n_c <= MODEL_MAX_N ? 2 * n_c : 0;
    L.tw = take(sizeof(cplx) * (size_t)nmax);
    L.twf = take(sizeof(cplx) * (size_t)(n_c <= MODEL_MAX_N ? 2 * fast::TW_TOTAL : 0));
    L.kx = take(sizeof(double) * (size_t)(n_c <= MODEL_MAX_N ? n_c : 0));
    L.lin = take((lin_is_complex ? sizeof(cplx) : sizeof(double)) * (size_t)lin_elems);
    L.coef = take(coef_elem * (size_t)lin_elems * coef_per_entry);
    L.cidx = take(mode == CM_INDEXED ? sizeof(int) * (size_t)n_c : 0);
    L.cscale = take(mode == CM_SEPARABLE ? sizeof(double) * 32 : 0);
    L.U[0] = take(state);
    L.U[1] = method_adaptive(method) ? take(state) : L.U[0];
    L.K = take(state);
    L.ERR = method == M_ETD35 ? take(state) : 0;
    for (int j = 1; j <= method_nl_buffers(method); ++j) L.NL[j] = take(state);
    L.total = off;
    return L;
}

struct GraphCache {
    cudaGraphExec_t exec = nullptr;
    void* ring = nullptr;
    double* ring_t = nullptr;
    int cap = 0, kernels = 0;
    cudaStream_t stream = nullptr;
};

// SMs the persistent grids are sized for: all of them, or RKS_SM_LIMIT (a slab-decomposed run that overlaps its
// exchange with the transforms leaves a few SMs to the communication kernels)
static cudaError_t usable_sm_count(int* out, int device) {
    cudaError_t e = cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (const char* lim = getenv("RKS_SM_LIMIT")) {
        const int v = atoi(lim);
        if (v > 0 && v < *out) *out = v;
    }
    return cudaSuccess;
}

struct rks_plan {
    DevPlan d;
    GraphCache graph;
    bool use_graph;                 // replay one captured trial/step (RKS_NO_GRAPH=1 disables)
    // independent-dt ensembles: one single-row plan per trajectory, launched together (blockIdx.z = row)
    DevPlan* multi_dev = nullptr;   // device array of per-row plans (nullptr: ordinary plan)
    long long multi_n = 0;
    int* multi_count_dev = nullptr; // running-row counter
    int* multi_count_host = nullptr;
    Ctrl* multi_ctrl_host = nullptr;    // pinned staging of all row control blocks
    size_t multi_ctrl_stride = 0, multi_log_cap = 0;
    unsigned char* multi_ctrl_base = nullptr; TrialRec* multi_log_base = nullptr;
    std::vector<DevPlan> multi_host;
    Layout lay;
    unsigned char* ws;
    rks_config cfg;
    int method, device;
    int sm_count;
    long long launches;
    rks_ctrl_host* pinned_ctrl;     // pinned staging for rks_read_ctrl
    Ctrl* pinned_raw;
    TrialRec* pinned_log;
    double h_coeff_host;            // fixed-step methods: host-side cache key (etd4.py:392)
    bool have_h_coeff_host;
    int roles_u_sel, roles_n_sel;   // host mirror refreshed by rks_read_ctrl
    int nl_rows_per_cta, nl_threads;
    bool nl_fast;                   // n in {512..8192}: register-resident FFT kernel (fft_fast.cuh)
    bool nl_small;                  // n in {64, 128, 256}: the same pipeline on slabs of packed rows
    bool pretransform;              // intermediate NLS stages: K1 applies the first inverse FFT pass (RKS_PT=0 disables)
    bool rfft_half;                 // real-field models: half-length forward transform (default for n <= 1024)
    bool pair_rows;                 // cubic model, n = 512 ... 4096: two rows per complex transform (fft_pair.cuh)
    size_t nl_smem;
    // N-D grid model (rks_set_model_nd): strided-axis handles, the fused last-axis kernel, spectral grid dims
    struct rks_axis* nd_axes[2] = {nullptr, nullptr};
    struct rks_rows* nd_rows = nullptr;
    int nd = 0;
    long long nd_spec[3] = {0, 0, 0};
};

// smem of a fast NL launch: the row slabs
template <int W>
static size_t nl_fast_smem(int model) {
    constexpr int THREADS = W == 16 ? 512 : 256;
    constexpr int RPC = THREADS / (32 * W);
    size_t bytes = (size_t)RPC * 512 * W * sizeof(cplx);
    // n = 8192: TMA staging buffer for the next row + its mbarrier (kernels.cuh)
    if (W == 16 && model >= RKS_MODEL_UUX_RFFT && model <= RKS_MODEL_CUBIC_RFFT) bytes += NL_STAGE_BYTES;
    return bytes;
}

template <int W, int MODEL>
static void launch_nl_fast_t(rks_plan* p, int j, int force, cudaStream_t stream) {
    const DevPlan& d = p->d;
    constexpr int THREADS = W == 16 ? 512 : 256;
    constexpr int RPC = THREADS / (32 * W);
    const size_t smem = nl_fast_smem<W>(MODEL);
    if constexpr (MODEL <= 4) {      // independent-dt plans step the stepping models only
        if (p->multi_n) {
            nl_fast_kernel_multi<W, MODEL><<<dim3(1, 1, (unsigned)p->multi_n), THREADS, smem, stream>>>(p->multi_dev, j, force);
            return;
        }
    }
    const long long groups = (d.batch + RPC - 1) / RPC;
    const long long resident = (long long)p->sm_count * (W == 16 ? 1 : 2);
    const unsigned grid = (unsigned)(groups < resident ? groups : resident);
    nl_fast_kernel<W, MODEL><<<grid, THREADS, smem, stream>>>(d, j, force);
}

template <int W>
static void launch_nl_fast(rks_plan* p, int j, int force, cudaStream_t stream) {
    switch (p->d.model) {
        case RKS_MODEL_UUX_RFFT: launch_nl_fast_t<W, 1>(p, j, force, stream); break;
        case RKS_MODEL_NLS_FFT: launch_nl_fast_t<W, 2>(p, j, force, stream); break;
        case RKS_MODEL_CUBIC_RFFT: launch_nl_fast_t<W, 3>(p, j, force, stream); break;
        case RKS_MODEL_DERIV_FFT: launch_nl_fast_t<W, 5>(p, j, force, stream); break;
        case RKS_MODEL_DERIV_RFFT_PAIR: launch_nl_fast_t<W, 6>(p, j, force, stream); break;
        default: launch_nl_fast_t<W, 4>(p, j, force, stream); break;
    }
}

template <int W, int MODEL>
static cudaError_t prepare_nl_fast_t() {
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    cudaError_t e = cudaFuncSetAttribute(nl_fast_kernel<W, MODEL>, attr, (int)nl_fast_smem<W>(MODEL));
    if constexpr (MODEL <= 4)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(nl_fast_kernel_multi<W, MODEL>, attr, (int)nl_fast_smem<W>(MODEL));
    if (MODEL == 2 && W > 1 && e == cudaSuccess)
        e = cudaFuncSetAttribute(nl_fast_pre_kernel<(W > 1 ? W : 2)>, attr, (int)nl_fast_smem<W>(MODEL));
    return e;
}
template <int W>
static cudaError_t prepare_nl_fast(int model) {
    switch (model) {
        case RKS_MODEL_UUX_RFFT: return prepare_nl_fast_t<W, 1>();
        case RKS_MODEL_NLS_FFT: return prepare_nl_fast_t<W, 2>();
        case RKS_MODEL_CUBIC_RFFT: return prepare_nl_fast_t<W, 3>();
        case RKS_MODEL_DERIV_FFT: return prepare_nl_fast_t<W, 5>();
        case RKS_MODEL_DERIV_RFFT_PAIR: return prepare_nl_fast_t<W, 6>();
        default: return prepare_nl_fast_t<W, 4>();
    }
}

template <int N>
static cudaError_t prepare_nl_small(int model) {
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    const int smem = NL_SMALL_WARPS * 512 * (int)sizeof(cplx);
    switch (model) {
        case RKS_MODEL_UUX_RFFT: return cudaFuncSetAttribute(nl_small_kernel<N, 1>, attr, smem);
        case RKS_MODEL_NLS_FFT: return cudaFuncSetAttribute(nl_small_kernel<N, 2>, attr, smem);
        case RKS_MODEL_CUBIC_RFFT: return cudaFuncSetAttribute(nl_small_kernel<N, 3>, attr, smem);
        case RKS_MODEL_DERIV_FFT: return cudaFuncSetAttribute(nl_small_kernel<N, 5>, attr, smem);
        case RKS_MODEL_DERIV_RFFT_PAIR: return cudaFuncSetAttribute(nl_small_kernel<N, 6>, attr, smem);
        default: return cudaFuncSetAttribute(nl_small_kernel<N, 4>, attr, smem);
    }
}

// frees a partially built handle when a creation function returns early (CUDA_TRY)
template <class T, void (*DESTROY)(T*)>
struct HandleGuard {
    T* h;
    ~HandleGuard() { if (h) DESTROY(h); }
    void release() { h = nullptr; }
};

extern "C" int rks_abi_version(void) { return RKS_ABI_VERSION; }
extern "C" const char* rks_last_error(void) { return g_err; }

static bool valid_method(int m) { return m >= 0 && m <= 6; }
extern "C" int rks_num_stages(int method) { return valid_method(method) ? method_stages(method) : -1; }
extern "C" int rks_num_nl_buffers(int method) { return valid_method(method) ? method_nl_buffers(method) : -1; }
extern "C" int rks_is_adaptive(int method) { return valid_method(method) ? (int)method_adaptive(method) : -1; }

extern "C" size_t rks_workspace_bytes(int method, int64_t batch, int64_t n_c, int64_t lin_elems, int lin_is_complex) {
    if (!valid_method(method) || batch <= 0 || n_c <= 0) return 0;
    if (lin_elems != n_c && lin_elems != batch * n_c) return 0;
    return make_layout(method, batch, n_c, lin_elems, lin_is_complex).total;
}

static CfgArgs cfg_args(const rks_config& c, int method) {
    CfgArgs a;
    a.epsilon = c.epsilon; a.incr_f = c.incr_f; a.decr_f = c.decr_f; a.safety_f = c.safety_f;
    a.adapt_cutoff = c.adapt_cutoff; a.minh = c.minh;
    a.inv_q = 1.0 / (double)method_q(method);            // 1.0 / self._q(), solveras.py:454
    a.modecutoff = c.modecutoff; a.contour_radius = c.contour_radius;
    a.contour_points = c.contour_points; a.r4_fix = c.if45dp_r4_fix;
    return a;
}

static int check_cfg(const rks_config* c) {
    if (!c) return fail(RKS_ERR_ARG, "config is null");
    if (!(c->epsilon > 0) || !(c->incr_f > 1.0) || !(c->decr_f < 1.0) || !(c->safety_f <= 1.0) ||
        !(c->adapt_cutoff < 1.0) || !(c->minh > 0))
        return fail(RKS_ERR_ARG, "SolverConfig value out of range (solveras.py:99-169)");
    if (!(c->modecutoff > 0 && c->modecutoff <= 1.0) || c->contour_points <= 1 || !(c->contour_radius > 0))
        return fail(RKS_ERR_ARG, "ETDConfig value out of range (etd.py:95-131)");
    return RKS_OK;
}

// mode CM_COLUMN / CM_FLAT (chosen from lin_elems), CM_INDEXED (lin_op = the distinct values, `index` maps the
// n_c modes to them) or CM_SEPARABLE (lin_op = the concatenated per-axis terms of an nd-dimensional grid `dims`)
static int plan_create_impl(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* lin_op,
                            int lin_is_complex, int64_t lin_elems, int mode, const int32_t* index, int nd,
                            const int64_t* dims, const rks_config* cfg, void* workspace, size_t workspace_bytes,
                            void* stream_v) {
    if (!out) return fail(RKS_ERR_ARG, "out is null");
    *out = nullptr;
    if (!valid_method(method)) return fail(RKS_ERR_ARG, "unknown method id");
    if (batch <= 0 || n_c <= 0 || lin_elems <= 0) return fail(RKS_ERR_ARG, "batch, n_c and the lin_op length must be positive");
    if (!lin_op || !workspace) return fail(RKS_ERR_ARG, "null device pointer");
    if (((uintptr_t)workspace) % ALIGN) return fail(RKS_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_v;
    Layout L = make_layout(method, batch, n_c, lin_elems, lin_is_complex, mode);
    if (workspace_bytes < L.total) return fail(RKS_ERR_WORKSPACE, "workspace too small");

    rks_plan* p = new (std::nothrow) rks_plan();
    if (!p) return fail(RKS_ERR_ARG, "out of host memory");
    HandleGuard<rks_plan, rks_plan_destroy> guard{p};
    memset(&p->d, 0, sizeof(DevPlan));
    p->lay = L;
    p->ws = (unsigned char*)workspace;
    p->cfg = *cfg;
    p->method = method;
    p->launches = 0;
    p->have_h_coeff_host = false;
    p->roles_u_sel = p->roles_n_sel = 0;
    p->use_graph = getenv("RKS_NO_GRAPH") == nullptr;
    CUDA_TRY(cudaGetDevice(&p->device));
    CUDA_TRY(usable_sm_count(&p->sm_count, p->device));
    CUDA_TRY(cudaMallocHost((void**)&p->pinned_raw, sizeof(Ctrl)));
    CUDA_TRY(cudaMallocHost((void**)&p->pinned_log, sizeof(TrialRec) * LOG_CAP));

    DevPlan& d = p->d;
    unsigned char* w = p->ws;
    d.ctrl = (Ctrl*)(w + L.ctrl);
    d.log = (TrialRec*)(w + L.log);
    d.partials = (double*)(w + L.partials);
    d.pre_cnt = (int*)(d.partials + 2 * NORM_MAX_BLOCKS);
    d.tw = (const cplx*)(w + L.tw);
    d.twf = (const cplx*)(w + L.twf);
    d.kx = (const double*)(w + L.kx);
    d.lin = w + L.lin;
    d.coef = w + L.coef;
    d.U[0] = (cplx*)(w + L.U[0]);
    d.U[1] = (cplx*)(w + L.U[1]);
    d.K = (cplx*)(w + L.K);
    d.ERR = method == M_ETD35 ? (cplx*)(w + L.ERR) : nullptr;
    for (int j = 1; j <= method_nl_buffers(method); ++j) d.NL[j] = (cplx*)(w + L.NL[j]);
    d.batch = batch; d.n_c = n_c; d.lin_elems = lin_elems; d.n = 0;
    d.method = method; d.lin_complex = lin_is_complex;
    d.coef_mode = mode;
    d.lin_full = mode == CM_FLAT ? 1 : 0;
    d.model = RKS_MODEL_NONE; d.log2n = 0; d.model_p0 = 0.0;
    if (mode == CM_INDEXED) {
        d.cidx = (const int*)(w + L.cidx);
        CUDA_TRY(cudaMemcpyAsync(w + L.cidx, index, sizeof(int) * (size_t)n_c, cudaMemcpyDeviceToDevice, stream));
    }
    if (mode == CM_SEPARABLE) {
        d.cscale = (const double*)(w + L.cscale);
        d.sep_nd = nd;
        for (int k = 0; k < nd; ++k) d.sep_dims[k] = (int)dims[k];
        d.sep_ntab = (int)lin_elems;
    }

    CUDA_TRY(cudaMemsetAsync(w + L.ctrl, 0, L.partials + sizeof(double) * 2 * NORM_MAX_BLOCKS + PRE_CNT_BYTES - L.ctrl, stream));
    CUDA_TRY(cudaMemcpyAsync(w + L.lin, lin_op, (lin_is_complex ? sizeof(cplx) : sizeof(double)) * (size_t)lin_elems,
                             cudaMemcpyDeviceToDevice, stream));
    set_config_kernel<<<1, 1, 0, stream>>>(d.ctrl, cfg_args(*cfg, method));
    BeginArgs b;
    b.t0 = 0.0; b.tf = 0.0; b.h = 0.0; b.store_freq = 0; b.step_mode = 1; b.keep_fsal = 0;
    b.n1_refresh = method == M_ETD35;
    begin_kernel<<<1, 1, 0, stream>>>(d.ctrl, b);
    p->launches += 2;
    CUDA_TRY(cudaGetLastError());
    guard.release();
    *out = p;
    return RKS_OK;
}

extern "C" int rks_plan_create(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* lin_op,
                               int lin_is_complex, int64_t lin_elems, const rks_config* cfg, void* workspace,
                               size_t workspace_bytes, void* stream_v) {
    if (out) *out = nullptr;
    if (lin_elems != n_c && lin_elems != batch * n_c)
        return fail(RKS_ERR_ARG, "lin_op must have n_c or batch*n_c elements");
    // lin_elems == batch * n_c also for batch == 1: the flat kernels
    const int mode = (lin_elems == batch * n_c) ? CM_FLAT : CM_COLUMN;
    return plan_create_impl(out, method, batch, n_c, lin_op, lin_is_complex, lin_elems, mode, nullptr, 0, nullptr, cfg,
                            workspace, workspace_bytes, stream_v);
}

// Grids with many modes but few distinct lin_op values: coefficient records per distinct value
extern "C" size_t rks_workspace_bytes_indexed(int method, int64_t batch, int64_t n_c, int64_t n_values, int lin_is_complex) {
    if (!valid_method(method) || batch <= 0 || batch > 65535 || n_c <= 0 || n_values <= 0 || n_values > n_c) return 0;
    return make_layout(method, batch, n_c, n_values, lin_is_complex, CM_INDEXED).total;
}
extern "C" int rks_plan_create_indexed(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* values,
                                       int lin_is_complex, int64_t n_values, const int32_t* index, const rks_config* cfg,
                                       void* workspace, size_t workspace_bytes, void* stream_v) {
    if (out) *out = nullptr;
    if (!index) return fail(RKS_ERR_ARG, "index is null");
    if (n_values <= 0 || n_values > n_c) return fail(RKS_ERR_ARG, "n_values must be in 1..n_c");
    if (batch > 65535) return fail(RKS_ERR_UNSUPPORTED, "indexed plans take at most 65535 trajectories");
    return plan_create_impl(out, method, batch, n_c, values, lin_is_complex, n_values, CM_INDEXED, index, 0, nullptr, cfg,
                            workspace, workspace_bytes, stream_v);
}

// IF methods on an nd-dimensional grid whose lin_op is a sum of per-axis terms: per-axis exponential tables
static int64_t sep_total(int nd, const int64_t* dims, int64_t* modes) {
    if (nd < 2 || nd > 3 || !dims) return 0;
    int64_t sum = 0, prod = 1;
    for (int k = 0; k < nd; ++k) {
        if (dims[k] <= 0 || dims[k] > (1ll << 24)) return 0;
        sum += dims[k]; prod *= dims[k];
    }
    if (modes) *modes = prod;
    return sum;
}
extern "C" size_t rks_workspace_bytes_separable(int method, int64_t batch, int nd, const int64_t* dims, int lin_is_complex) {
    int64_t n_c = 0;
    const int64_t ntab = sep_total(nd, dims, &n_c);
    if (!valid_method(method) || !method_is_if(method) || batch <= 0 || !ntab) return 0;
    return make_layout(method, batch, n_c, ntab, lin_is_complex, CM_SEPARABLE).total;
}
extern "C" int rks_plan_create_separable(rks_plan** out, int method, int64_t batch, int nd, const int64_t* dims,
                                         const void* axis_terms, int lin_is_complex, const rks_config* cfg,
                                         void* workspace, size_t workspace_bytes, void* stream_v) {
    if (out) *out = nullptr;
    if (!valid_method(method) || !method_is_if(method))
        return fail(RKS_ERR_UNSUPPORTED, "separable coefficient tables exist for the IF methods only");
    int64_t n_c = 0;
    const int64_t ntab = sep_total(nd, dims, &n_c);
    if (!ntab) return fail(RKS_ERR_ARG, "separable plans take 2 or 3 positive grid dimensions");
    const int64_t rows = batch * (n_c / dims[nd - 1]);
    if (rows > 65535ll * 8 * STAGE_R) return fail(RKS_ERR_UNSUPPORTED, "too many grid rows for one launch");
    return plan_create_impl(out, method, batch, n_c, axis_terms, lin_is_complex, ntab, CM_SEPARABLE, nullptr, nd, dims, cfg,
                            workspace, workspace_bytes, stream_v);
}

// ---------------------------------------------------------------------------------------
// independent-dt ensembles (BASELINE cfg 2b): one single-row plan per trajectory -- its own control
// block, coefficient arrays and buffer roles -- all launched together (gridDim.z = rows)
// ---------------------------------------------------------------------------------------
struct MultiLayout {
    size_t tw, twf, kx, lin, plans, count, ctrl, ctrl_stride, log, partials, coef, coef_stride, U[2], K, ERR, NL[8], total;
};

static MultiLayout make_multi_layout(int method, long long batch, long long n_c, int lin_is_complex) {
    MultiLayout L;
    memset(&L, 0, sizeof(L));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    const size_t state = (size_t)batch * (size_t)n_c * sizeof(cplx);
    const size_t coef_elem = (method_is_if(method) && !lin_is_complex) ? sizeof(double) : sizeof(cplx);
    const long long nmax = n_c <= MODEL_MAX_N ? 2 * n_c : 0;
    L.tw = take(sizeof(cplx) * (size_t)nmax);
    L.twf = take(sizeof(cplx) * (size_t)(n_c <= MODEL_MAX_N ? 2 * fast::TW_TOTAL : 0));
    L.kx = take(sizeof(double) * (size_t)(n_c <= MODEL_MAX_N ? n_c : 0));
    L.lin = take((lin_is_complex ? sizeof(cplx) : sizeof(double)) * (size_t)n_c);
    L.plans = take(sizeof(DevPlan) * (size_t)batch);
    L.count = take(sizeof(int) * 4);
    L.ctrl_stride = align_up(sizeof(Ctrl));
    L.ctrl = take(L.ctrl_stride * (size_t)batch);
    L.log = take(sizeof(TrialRec) * MULTI_LOG_CAP * (size_t)batch);
    L.partials = take(sizeof(double) * 2 * MULTI_NORM_BLOCKS * (size_t)batch);
    L.coef_stride = coef_elem * (size_t)n_c * method_ncoef(method);
    L.coef = take(L.coef_stride * (size_t)batch);
    L.U[0] = take(state);
    L.U[1] = take(state);
    L.K = take(state);
    L.ERR = method == M_ETD35 ? take(state) : 0;
    for (int j = 1; j <= method_nl_buffers(method); ++j) L.NL[j] = take(state);
    L.total = off;
    return L;
}

extern "C" size_t rks_workspace_bytes_independent(int method, int64_t batch, int64_t n_c, int lin_is_complex) {
    if (!valid_method(method) || !method_adaptive(method) || batch <= 0 || batch > 65535 || n_c <= 0) return 0;
    return make_multi_layout(method, batch, n_c, lin_is_complex).total;
}

static int upload_multi_plans(rks_plan* p, cudaStream_t stream) {
    // every row shares the model fields of the prototype
    for (auto& q : p->multi_host) {
        q.model = p->d.model; q.n = p->d.n; q.log2n = p->d.log2n; q.model_p0 = p->d.model_p0;
    }
    CUDA_TRY(cudaMemcpyAsync(p->multi_dev, p->multi_host.data(), sizeof(DevPlan) * p->multi_host.size(),
                             cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return RKS_OK;
}

extern "C" int rks_plan_create_independent(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* lin_op,
                                           int lin_is_complex, const rks_config* cfg, void* workspace,
                                           size_t workspace_bytes, void* stream_v) {
    if (!out) return fail(RKS_ERR_ARG, "out is null");
    *out = nullptr;
    if (!valid_method(method) || !method_adaptive(method)) return fail(RKS_ERR_ARG, "independent dt needs an adaptive method");
    if (batch <= 0 || batch > 65535 || n_c <= 0) return fail(RKS_ERR_ARG, "batch must be in 1..65535 and n_c positive");
    if (!lin_op || !workspace) return fail(RKS_ERR_ARG, "null device pointer");
    if (((uintptr_t)workspace) % ALIGN) return fail(RKS_ERR_WORKSPACE, "workspace must be 256-byte aligned");
    if (int rc = check_cfg(cfg)) return rc;
    cudaStream_t stream = (cudaStream_t)stream_v;
    const MultiLayout L = make_multi_layout(method, batch, n_c, lin_is_complex);
    if (workspace_bytes < L.total) return fail(RKS_ERR_WORKSPACE, "workspace too small");
    rks_plan* p = new (std::nothrow) rks_plan();
    if (!p) return fail(RKS_ERR_ARG, "out of host memory");
    HandleGuard<rks_plan, rks_plan_destroy> guard{p};
    memset(&p->d, 0, sizeof(DevPlan));
    memset(&p->lay, 0, sizeof(Layout));
    p->ws = (unsigned char*)workspace;
    p->cfg = *cfg;
    p->method = method;
    p->launches = 0;
    p->have_h_coeff_host = false;
    p->roles_u_sel = p->roles_n_sel = 0;
    p->use_graph = getenv("RKS_NO_GRAPH") == nullptr;
    p->nl_fast = false;
    p->nl_small = false;
    CUDA_TRY(cudaGetDevice(&p->device));
    CUDA_TRY(usable_sm_count(&p->sm_count, p->device));
    CUDA_TRY(cudaMallocHost((void**)&p->pinned_raw, sizeof(Ctrl)));
    CUDA_TRY(cudaMallocHost((void**)&p->pinned_log, sizeof(TrialRec) * LOG_CAP));
    CUDA_TRY(cudaMallocHost((void**)&p->multi_ctrl_host, L.ctrl_stride * (size_t)batch));
    CUDA_TRY(cudaMallocHost((void**)&p->multi_count_host, sizeof(int) * 4));
    unsigned char* w = p->ws;
    p->lay.tw = L.tw; p->lay.twf = L.twf; p->lay.kx = L.kx; p->lay.lin = L.lin;
    p->multi_n = batch;
    p->multi_dev = (DevPlan*)(w + L.plans);
    p->multi_count_dev = (int*)(w + L.count);
    p->multi_ctrl_base = w + L.ctrl;
    p->multi_ctrl_stride = L.ctrl_stride;
    p->multi_log_base = (TrialRec*)(w + L.log);
    p->multi_log_cap = MULTI_LOG_CAP;
    p->multi_host.resize((size_t)batch);
    for (long long r = 0; r < batch; ++r) {
        DevPlan& d = p->multi_host[(size_t)r];
        memset(&d, 0, sizeof(DevPlan));
        d.ctrl = (Ctrl*)(w + L.ctrl + L.ctrl_stride * (size_t)r);
        d.log = (TrialRec*)(w + L.log) + (size_t)r * MULTI_LOG_CAP;
        d.partials = (double*)(w + L.partials) + (size_t)r * 2 * MULTI_NORM_BLOCKS;
        d.tw = (const cplx*)(w + L.tw);
        d.twf = (const cplx*)(w + L.twf);
        d.kx = (const double*)(w + L.kx);
        d.lin = w + L.lin;
        d.coef = w + L.coef + L.coef_stride * (size_t)r;
        const size_t ro = (size_t)r * (size_t)n_c;
        d.U[0] = (cplx*)(w + L.U[0]) + ro;
        d.U[1] = (cplx*)(w + L.U[1]) + ro;
        d.K = (cplx*)(w + L.K) + ro;
        d.ERR = method == M_ETD35 ? (cplx*)(w + L.ERR) + ro : nullptr;
        for (int j = 1; j <= method_nl_buffers(method); ++j) d.NL[j] = (cplx*)(w + L.NL[j]) + ro;
        d.batch = 1; d.n_c = n_c; d.lin_elems = n_c; d.n = 0;
        d.method = method; d.lin_complex = lin_is_complex; d.lin_full = 1; d.coef_mode = CM_FLAT;
        d.model = RKS_MODEL_NONE;
    }
    p->d = p->multi_host[0];
    CUDA_TRY(cudaMemsetAsync(w + L.count, 0, L.coef - L.count, stream));          // counters, ctrl, logs, partials
    CUDA_TRY(cudaMemcpyAsync(w + L.lin, lin_op, (lin_is_complex ? sizeof(cplx) : sizeof(double)) * (size_t)n_c,
                             cudaMemcpyDeviceToDevice, stream));
    if (int rc = upload_multi_plans(p, stream)) return rc;
    const unsigned g = (unsigned)((batch + 127) / 128);
    set_config_multi_kernel<<<g, 128, 0, stream>>>(p->multi_dev, (int)batch, cfg_args(*cfg, method), MULTI_LOG_CAP);
    BeginArgs b;
    b.t0 = 0.0; b.tf = 0.0; b.h = 0.0; b.store_freq = 0; b.step_mode = 1; b.keep_fsal = 0;
    b.n1_refresh = method == M_ETD35;
    begin_multi_kernel<<<g, 128, 0, stream>>>(p->multi_dev, (int)batch, b);
    p->launches += 2;
    CUDA_TRY(cudaGetLastError());
    guard.release();
    *out = p;
    return RKS_OK;
}

extern "C" void rks_plan_destroy(rks_plan* p) {
    if (!p) return;
    if (p->nd_axes[1] && p->nd_axes[1] != p->nd_axes[0]) rks_axis_destroy(p->nd_axes[1]);
    if (p->nd_axes[0]) rks_axis_destroy(p->nd_axes[0]);
    if (p->nd_rows) rks_rows_destroy(p->nd_rows);
    if (p->graph.exec) cudaGraphExecDestroy(p->graph.exec);
    if (p->graph.stream) cudaStreamDestroy(p->graph.stream);
    cudaFreeHost(p->pinned_raw);
    cudaFreeHost(p->pinned_log);
    if (p->multi_ctrl_host) cudaFreeHost(p->multi_ctrl_host);
    if (p->multi_count_host) cudaFreeHost(p->multi_count_host);
    delete p;
}

extern "C" int rks_set_config(rks_plan* p, const rks_config* cfg, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (int rc = check_cfg(cfg)) return rc;
    const bool etd_changed = cfg->modecutoff != p->cfg.modecutoff || cfg->contour_points != p->cfg.contour_points ||
                             cfg->contour_radius != p->cfg.contour_radius || cfg->if45dp_r4_fix != p->cfg.if45dp_r4_fix;
    p->cfg = *cfg;
    if (etd_changed) p->have_h_coeff_host = false;
    if (p->multi_n)
        set_config_multi_kernel<<<(unsigned)((p->multi_n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            p->multi_dev, (int)p->multi_n, cfg_args(*cfg, p->method), MULTI_LOG_CAP);
    else
        set_config_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->d.ctrl, cfg_args(*cfg, p->method));
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

static int prepare_nl_launch(rks_plan* p, int model, long long n, cplx* twf_dev, cudaStream_t stream) {
    DevPlan& d = p->d;
    p->nl_fast = (n >= 512 && n <= 8192) && !getenv("RKS_NL_GENERIC");
    // real-field models: half-length forward transform (fft_real.cuh).  Measured (profiles/r02_k4_ab.md): 7 % faster
    // than the full-length pair for n = 512 and 1024, slower for n >= 2048 (three more row barriers per row)
    const char* rh = getenv("RKS_RFFT_HALF");
    p->rfft_half = p->nl_fast && (rh ? (rh[0] != '0' && n <= 4096) : n <= 1024)
                   && (model == RKS_MODEL_UUX_RFFT || model == RKS_MODEL_CUBIC_RFFT);
    // single-real-field cubic model: two rows share one complex transform pair (RKS_PAIR_ROWS=0: off)
    const char* pr = getenv("RKS_PAIR_ROWS");
    p->pair_rows = p->nl_fast && n <= 4096 && model == RKS_MODEL_CUBIC_RFFT && !(pr && pr[0] == '0');
    if (p->pair_rows) {
        const auto a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        const int sm = 256 / 32 * 512 * (int)sizeof(cplx);          // pairs per CTA x n x 16 B = 64 KB for every n
        CUDA_TRY(n == 512 ? cudaFuncSetAttribute(nl_fast_pair_kernel<1>, a, sm) : n == 1024 ? cudaFuncSetAttribute(nl_fast_pair_kernel<2>, a, sm)
                 : n == 2048 ? cudaFuncSetAttribute(nl_fast_pair_kernel<4>, a, sm) : cudaFuncSetAttribute(nl_fast_pair_kernel<8>, a, sm));
    }
    if (p->rfft_half) {
        const auto a = cudaFuncAttributeMaxDynamicSharedMemorySize;
        const int sm = 256 / 32 * 512 * (int)sizeof(cplx);          // rows per CTA x n x 16 B = 64 KB for every n
        const bool uux = model == RKS_MODEL_UUX_RFFT;
        cudaError_t e = n == 512 ? (uux ? cudaFuncSetAttribute(nl_fast_real_kernel<1, 1>, a, sm) : cudaFuncSetAttribute(nl_fast_real_kernel<1, 3>, a, sm))
                      : n == 1024 ? (uux ? cudaFuncSetAttribute(nl_fast_real_kernel<2, 1>, a, sm) : cudaFuncSetAttribute(nl_fast_real_kernel<2, 3>, a, sm))
                      : n == 2048 ? (uux ? cudaFuncSetAttribute(nl_fast_real_kernel<4, 1>, a, sm) : cudaFuncSetAttribute(nl_fast_real_kernel<4, 3>, a, sm))
                                  : (uux ? cudaFuncSetAttribute(nl_fast_real_kernel<8, 1>, a, sm) : cudaFuncSetAttribute(nl_fast_real_kernel<8, 3>, a, sm));
        CUDA_TRY(e);
    }
    {   // per-row offsets between the warps of a scheduler in barrier-synchronised rows (kernels.cuh row_stagger_spin)
        const char* sm = getenv("RKS_ROW_STAGGER_MODE");
        const char* sg = getenv("RKS_ROW_STAGGER_CYC");
        const char* np = getenv("RKS_ROW_STAGGER_NP");
        const int mode = sm ? atoi(sm) : 1, v = sg ? atoi(sg) : 1000, vnp = np ? atoi(np) : 1;
        CUDA_TRY(cudaMemcpyToSymbol(c_row_stagger_mode, &mode, sizeof(int)));
        CUDA_TRY(cudaMemcpyToSymbol(c_row_stagger_cyc, &v, sizeof(int)));
        CUDA_TRY(cudaMemcpyToSymbol(c_row_stagger_np, &vnp, sizeof(int)));
    }
    const char* pt = getenv("RKS_PT");
    p->pretransform = !(pt && pt[0] == '0');           // pre-transformed intermediate stages
    p->nl_small = (n == 64 || n == 128 || n == 256) && !getenv("RKS_NL_GENERIC");
    if (p->nl_fast) {
        cudaError_t e = n == 512 ? prepare_nl_fast<1>(model) : n == 1024 ? prepare_nl_fast<2>(model)
                      : n == 2048 ? prepare_nl_fast<4>(model) : n == 4096 ? prepare_nl_fast<8>(model)
                      : prepare_nl_fast<16>(model);
        CUDA_TRY(e);
    }
    if (p->nl_small) CUDA_TRY(n == 64 ? prepare_nl_small<64>(model) : n == 128 ? prepare_nl_small<128>(model) : prepare_nl_small<256>(model));
    if (p->nl_fast || p->nl_small) {
        fast_twiddle_kernel<<<(2 * fast::TW_TOTAL + 255) / 256, 256, 0, stream>>>(twf_dev, (int)n);
        p->launches += 1;
    }
    // launch shape of the generic NL kernel: one row per CTA for long rows, several for short ones
    const size_t row_bytes = (size_t)n * sizeof(cplx);
    int tpr = (int)(n / 4);                       // one radix-4 butterfly per thread per pass
    if (tpr > 512) tpr = 512;
    if (tpr < 32) tpr = 32;
    int rows = 256 / tpr;
    if (rows < 1) rows = 1;
    while (rows > 1 && (long long)rows > d.batch) rows >>= 1;
    p->nl_rows_per_cta = rows;
    p->nl_threads = rows * tpr;
    p->nl_smem = row_bytes * rows;
    if (p->nl_smem > 227 * 1024) return fail(RKS_ERR_UNSUPPORTED, "row does not fit in shared memory");
    const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    if (model == RKS_MODEL_UUX_RFFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel<1>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_NLS_FFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel<2>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_CUBIC_RFFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel<3>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_DERIV_FFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel<5>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_DERIV_RFFT_PAIR) CUDA_TRY(cudaFuncSetAttribute(nl_kernel<6>, attr, (int)p->nl_smem));
    else CUDA_TRY(cudaFuncSetAttribute(nl_kernel<4>, attr, (int)p->nl_smem));
    if (model > RKS_MODEL_SINE_GORDON) return RKS_OK;          // no independent-dt variant of the derivative rows
    if (model == RKS_MODEL_UUX_RFFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel_multi<1>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_NLS_FFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel_multi<2>, attr, (int)p->nl_smem));
    else if (model == RKS_MODEL_CUBIC_RFFT) CUDA_TRY(cudaFuncSetAttribute(nl_kernel_multi<3>, attr, (int)p->nl_smem));
    else CUDA_TRY(cudaFuncSetAttribute(nl_kernel_multi<4>, attr, (int)p->nl_smem));
    return RKS_OK;
}

extern "C" int rks_set_model(rks_plan* p, int model, int64_t n, const double* kx, const double* params_host,
                             int nparams, void* stream_v) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    cudaStream_t stream = (cudaStream_t)stream_v;
    DevPlan& d = p->d;
    if (model == RKS_MODEL_NONE) { d.model = 0; return RKS_OK; }
    if (model < RKS_MODEL_UUX_RFFT || model > RKS_MODEL_SINE_GORDON) return fail(RKS_ERR_ARG, "unknown model id");
    if (d.lin_elems != d.n_c) return fail(RKS_ERR_UNSUPPORTED, "fused 1-D models need lin_op of n_c elements");
    if (n < 16 || n > MODEL_MAX_N || (n & (n - 1))) return fail(RKS_ERR_UNSUPPORTED, "n must be a power of two in [16, 16384]");
    const bool half_spectrum = model == RKS_MODEL_UUX_RFFT || model == RKS_MODEL_CUBIC_RFFT;
    const long long want_nc = half_spectrum ? n / 2 + 1 : n;
    if (want_nc != d.n_c) return fail(RKS_ERR_ARG, "n does not match n_c for this model");
    if (nparams < 1 || !params_host) return fail(RKS_ERR_ARG, "model needs one parameter");
    if ((model == RKS_MODEL_UUX_RFFT || model == RKS_MODEL_SINE_GORDON) && !kx) return fail(RKS_ERR_ARG, "kx is null");
    int log2n = 0;
    while ((1ll << log2n) < n) ++log2n;
    d.n = n; d.log2n = log2n; d.model = model; d.model_p0 = params_host[0];
    if (p->graph.exec) { cudaGraphExecDestroy(p->graph.exec); p->graph.exec = nullptr; }
    twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((cplx*)(p->ws + p->lay.tw), (int)n);
    p->launches += 1;
    if (kx) CUDA_TRY(cudaMemcpyAsync(p->ws + p->lay.kx, kx, sizeof(double) * (size_t)d.n_c, cudaMemcpyDeviceToDevice, stream));
    if (int rc = prepare_nl_launch(p, model, n, (cplx*)(p->ws + p->lay.twf), stream)) return rc;
    if (p->multi_n)
        if (int rc = upload_multi_plans(p, stream)) return rc;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_begin(rks_plan* p, double t0, double tf, double h, int64_t store_freq, int step_mode,
                         int keep_fsal, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    BeginArgs b;
    b.t0 = t0; b.tf = tf; b.h = h; b.store_freq = store_freq; b.step_mode = step_mode; b.keep_fsal = keep_fsal;
    b.n1_refresh = p->method == M_ETD35;
    if (!keep_fsal) { p->have_h_coeff_host = false; p->roles_n_sel = 0; }
    if (p->multi_n)
        begin_multi_kernel<<<(unsigned)((p->multi_n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p->multi_dev, (int)p->multi_n, b);
    else
        begin_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->d.ctrl, b);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_set_h(rks_plan* p, double h, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (p->multi_n)
        set_h_multi_kernel<<<(unsigned)((p->multi_n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p->multi_dev, (int)p->multi_n, h);
    else
        set_h_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->d.ctrl, h);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

static unsigned copy_grid(const rks_plan* p) {
    const long long total = p->d.batch * p->d.n_c;
    long long g = (total + 255) / 256;
    const long long cap = (long long)p->sm_count * 16;
    return (unsigned)(g < cap ? g : cap);
}

extern "C" int rks_set_u(rks_plan* p, const void* u, void* stream) {
    if (!p || !u) return fail(RKS_ERR_ARG, "null argument");
    if (p->multi_n)
        copy_u_multi_kernel<<<dim3((unsigned)((p->d.n_c + 255) / 256 > 64 ? 64 : (p->d.n_c + 255) / 256), 1, (unsigned)p->multi_n), 256, 0,
                              (cudaStream_t)stream>>>(p->multi_dev, (cplx*)u, 1);
    else
        copy_u_kernel<<<copy_grid(p), 256, 0, (cudaStream_t)stream>>>(p->d, (cplx*)u, 1);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_get_u(rks_plan* p, void* u_out, void* stream) {
    if (!p || !u_out) return fail(RKS_ERR_ARG, "null argument");
    if (p->multi_n)
        copy_u_multi_kernel<<<dim3((unsigned)((p->d.n_c + 255) / 256 > 64 ? 64 : (p->d.n_c + 255) / 256), 1, (unsigned)p->multi_n), 256, 0,
                              (cudaStream_t)stream>>>(p->multi_dev, (cplx*)u_out, 0);
    else
        copy_u_kernel<<<copy_grid(p), 256, 0, (cudaStream_t)stream>>>(p->d, (cplx*)u_out, 0);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// K2 dispatch
// ---------------------------------------------------------------------------------------
template <int M, typename LT>
static void launch_coef_t(rks_plan* p, int force, cudaStream_t stream) {
    const DevPlan& d = p->d;
    const unsigned grid = (unsigned)((d.lin_elems + 127) / 128);
    if (p->multi_n) coef_kernel_multi<M, LT><<<dim3(grid, 1, (unsigned)p->multi_n), 128, 0, stream>>>(p->multi_dev, force);
    else if (d.coef_mode == CM_INDEXED) coef_kernel<M, LT, true><<<grid, 128, 0, stream>>>(d, force);
    else if (d.coef_mode == CM_SEPARABLE) {
        if constexpr (method_is_if(M)) coef_sep_kernel<M, LT><<<(unsigned)((d.sep_ntab + 127) / 128), 128, 0, stream>>>(d, force);
    } else coef_kernel<M, LT, false><<<grid, 128, 0, stream>>>(d, force);
}

static int launch_coeffs(rks_plan* p, int force, cudaStream_t stream) {
    const bool cx = p->d.lin_complex != 0;
    switch (p->method) {
        case M_IF4: cx ? launch_coef_t<M_IF4, cplx>(p, force, stream) : launch_coef_t<M_IF4, double>(p, force, stream); break;
        case M_IF34: cx ? launch_coef_t<M_IF34, cplx>(p, force, stream) : launch_coef_t<M_IF34, double>(p, force, stream); break;
        case M_IF45DP: cx ? launch_coef_t<M_IF45DP, cplx>(p, force, stream) : launch_coef_t<M_IF45DP, double>(p, force, stream); break;
        case M_ETD4: launch_coef_t<M_ETD4, cplx>(p, force, stream); break;
        case M_ETD34: launch_coef_t<M_ETD34, cplx>(p, force, stream); break;
        case M_ETD5: launch_coef_t<M_ETD5, cplx>(p, force, stream); break;
        default: launch_coef_t<M_ETD35, cplx>(p, force, stream); break;
    }
    p->launches += 1;
    return RKS_OK;
}

extern "C" int rks_update_coeffs(rks_plan* p, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    // adaptive methods: the device compares ctrl.h with ctrl.h_coeff.  Fixed-step methods have no
    // controller kernel, so the host keeps the cache key (the caller sets h through rks_set_h).
    launch_coeffs(p, method_adaptive(p->method) ? 0 : 1, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// K1 dispatch
// ---------------------------------------------------------------------------------------
template <int M, int S, typename CT>
static void launch_stage_t(rks_plan* p, cudaStream_t stream) {
    const DevPlan& d = p->d;
    const dim3 block(32, 8);
    if (p->multi_n) {
        const unsigned gx = (unsigned)((d.n_c + 256 * STAGE_R - 1) / (256 * STAGE_R));
        stage_kernel_multi<M, S, CT, CM_FLAT><<<dim3(gx, 1, (unsigned)p->multi_n), block, 0, stream>>>(p->multi_dev);
    } else if (d.coef_mode == CM_FLAT) {
        const long long total = d.batch * d.n_c;
        const unsigned gx = (unsigned)((total + 256 * STAGE_R - 1) / (256 * STAGE_R));
        stage_kernel<M, S, CT, CM_FLAT><<<dim3(gx), block, 0, stream>>>(d);
    } else if (d.coef_mode == CM_INDEXED) {
        const unsigned gx = (unsigned)((d.n_c + 256 * STAGE_R - 1) / (256 * STAGE_R));
        stage_kernel<M, S, CT, CM_INDEXED><<<dim3(gx, (unsigned)d.batch), block, 0, stream>>>(d);
    } else if (d.coef_mode == CM_SEPARABLE) {
        if constexpr (method_is_if(M)) {
            const long long ncol = d.sep_dims[d.sep_nd - 1], nrow = d.batch * (d.n_c / ncol);
            const unsigned gx = (unsigned)((ncol + 31) / 32);
            const unsigned gy = (unsigned)((nrow + 8 * STAGE_R - 1) / (8 * STAGE_R));
            stage_kernel<M, S, CT, CM_SEPARABLE><<<dim3(gx, gy), block, 0, stream>>>(d);
        }
    } else {
        const unsigned gx = (unsigned)((d.n_c + 31) / 32);
        const unsigned gy = (unsigned)((d.batch + 8 * STAGE_R - 1) / (8 * STAGE_R));
        stage_kernel<M, S, CT, CM_COLUMN><<<dim3(gx, gy), block, 0, stream>>>(d);
    }
}

template <int M, typename CT>
static int launch_stage_m(rks_plan* p, int s, cudaStream_t stream) {
    constexpr int SMAX = method_stages(M);
    switch (s) {
        case 1: launch_stage_t<M, 1, CT>(p, stream); break;
        case 2: launch_stage_t<M, 2, CT>(p, stream); break;
        case 3: launch_stage_t<M, 3, CT>(p, stream); break;
        case 4: launch_stage_t<M, 4, CT>(p, stream); break;
        case 5: if (SMAX >= 5) launch_stage_t<M, (SMAX >= 5 ? 5 : 1), CT>(p, stream); break;
        case 6: if (SMAX >= 6) launch_stage_t<M, (SMAX >= 6 ? 6 : 1), CT>(p, stream); break;
        default: break;
    }
    return RKS_OK;
}

// K1 for a pre-transformed intermediate stage (stage_pre_kernel): R1 = radix of the first FFT pass of the row
template <int M, int S, typename CT, int R1>
static void launch_stage_pre_r(rks_plan* p, cudaStream_t stream) {
    using C = PreCfg<M, S, CT, R1>;
    // once per instantiation AND device (function attributes are per device): opt in to the dynamic shared
    // memory and ask how many CTAs an SM holds
    static int resident_dev[MAX_DEVICES] = {0};
    int resident;
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        int& slot = resident_dev[p->device % MAX_DEVICES];
        if (!slot) {
            cudaFuncSetAttribute(stage_pre_kernel<M, S, CT, R1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, stage_pre_kernel<M, S, CT, R1>, PRE_THREADS, C::SMEM);
            slot = nb > 0 ? nb : 1;
        }
        resident = slot;
    }
    const DevPlan& d = p->d;
    // persistent: column blocks x row groups fill the resident slots once; every CTA walks down the batch
    const long long gx = d.n / R1 / 32, slots = (long long)p->sm_count * resident;
    long long gy = slots / gx, need = (d.batch + 3) / 4;
    if (gy < 1) gy = 1;
    if (gy > need) gy = need;
    stage_pre_kernel<M, S, CT, R1><<<dim3((unsigned)gx, (unsigned)gy), PRE_THREADS, C::SMEM, stream>>>(d);
}
template <int M, int S, typename CT>
static void launch_stage_pre_t(rks_plan* p, cudaStream_t stream) {
    launch_stage_pre_r<M, S, CT, 16>(p, stream);       // fft_fast.cuh Plan<N>::R1 for n = 1024 ... 8192
}
template <int M, typename CT>
static void launch_stage_pre_m(rks_plan* p, int s, cudaStream_t stream) {
    constexpr int SMAX = method_stages(M);
    switch (s) {
        case 1: launch_stage_pre_t<M, 1, CT>(p, stream); break;
        case 2: launch_stage_pre_t<M, 2, CT>(p, stream); break;
        case 3: launch_stage_pre_t<M, 3, CT>(p, stream); break;
        case 4: if (SMAX > 4) launch_stage_pre_t<M, (SMAX > 4 ? 4 : 1), CT>(p, stream); break;
        case 5: if (SMAX > 5) launch_stage_pre_t<M, (SMAX > 5 ? 5 : 1), CT>(p, stream); break;
        default: break;
    }
}
static void launch_stage_pre(rks_plan* p, int s, cudaStream_t stream) {
    const bool cx = p->d.lin_complex != 0;
    switch (p->method) {
        case M_IF4: cx ? launch_stage_pre_m<M_IF4, cplx>(p, s, stream) : launch_stage_pre_m<M_IF4, double>(p, s, stream); break;
        case M_IF34: cx ? launch_stage_pre_m<M_IF34, cplx>(p, s, stream) : launch_stage_pre_m<M_IF34, double>(p, s, stream); break;
        case M_IF45DP: cx ? launch_stage_pre_m<M_IF45DP, cplx>(p, s, stream) : launch_stage_pre_m<M_IF45DP, double>(p, s, stream); break;
        case M_ETD4: launch_stage_pre_m<M_ETD4, cplx>(p, s, stream); break;
        case M_ETD34: launch_stage_pre_m<M_ETD34, cplx>(p, s, stream); break;
        case M_ETD5: launch_stage_pre_m<M_ETD5, cplx>(p, s, stream); break;
        case M_ETD35: launch_stage_pre_m<M_ETD35, cplx>(p, s, stream); break;
    }
    p->launches += 1;
}

extern "C" int rks_stage(rks_plan* p, int s, void* stream_v) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (s < 1 || s > method_stages(p->method)) return fail(RKS_ERR_ARG, "stage out of range");
    if (p->d.batch > 65535ll * 8 * STAGE_R && p->d.coef_mode == CM_COLUMN) return fail(RKS_ERR_UNSUPPORTED, "batch too large for one launch");
    cudaStream_t stream = (cudaStream_t)stream_v;
    const bool cx = p->d.lin_complex != 0;
    switch (p->method) {
        case M_IF4: cx ? launch_stage_m<M_IF4, cplx>(p, s, stream) : launch_stage_m<M_IF4, double>(p, s, stream); break;
        case M_IF34: cx ? launch_stage_m<M_IF34, cplx>(p, s, stream) : launch_stage_m<M_IF34, double>(p, s, stream); break;
        case M_IF45DP: cx ? launch_stage_m<M_IF45DP, cplx>(p, s, stream) : launch_stage_m<M_IF45DP, double>(p, s, stream); break;
        case M_ETD4: launch_stage_m<M_ETD4, cplx>(p, s, stream); break;
        case M_ETD34: launch_stage_m<M_ETD34, cplx>(p, s, stream); break;
        case M_ETD5: launch_stage_m<M_ETD5, cplx>(p, s, stream); break;
        case M_ETD35: launch_stage_m<M_ETD35, cplx>(p, s, stream); break;
    }
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// K4 dispatch
// ---------------------------------------------------------------------------------------
static void dispatch_nl_fast(rks_plan* p, int j, int force, cudaStream_t stream) {
    switch (p->d.n) {
        case 512: launch_nl_fast<1>(p, j, force, stream); break;
        case 1024: launch_nl_fast<2>(p, j, force, stream); break;
        case 2048: launch_nl_fast<4>(p, j, force, stream); break;
        case 4096: launch_nl_fast<8>(p, j, force, stream); break;
        default: launch_nl_fast<16>(p, j, force, stream); break;
    }
    p->launches += 1;
}

// K4 on a row stage_pre_kernel has pre-transformed (NLS model, plain evaluation)
template <int W>
static void launch_nl_fast_pre_t(rks_plan* p, int j, int force, cudaStream_t stream) {
    constexpr int THREADS = W == 16 ? 512 : 256;
    constexpr int RPC = THREADS / (32 * W);
    const long long groups = (p->d.batch + RPC - 1) / RPC;
    const long long resident = (long long)p->sm_count * (W == 16 ? 1 : 2);
    const unsigned grid = (unsigned)(groups < resident ? groups : resident);
    nl_fast_pre_kernel<W><<<grid, THREADS, nl_fast_smem<W>(RKS_MODEL_NLS_FFT), stream>>>(p->d, j, force);
}
static void dispatch_nl_fast_pre(rks_plan* p, int j, int force, cudaStream_t stream) {
    switch (p->d.n) {
        case 1024: launch_nl_fast_pre_t<2>(p, j, force, stream); break;
        case 2048: launch_nl_fast_pre_t<4>(p, j, force, stream); break;
        case 4096: launch_nl_fast_pre_t<8>(p, j, force, stream); break;
        default: launch_nl_fast_pre_t<16>(p, j, force, stream); break;
    }
    p->launches += 1;
}

template <int N, int MODEL>
static void launch_nl_small_t(rks_plan* p, int j, int force, cudaStream_t stream) {
    const long long slabs = (p->d.batch + 512 / N - 1) / (512 / N);
    const long long ctas = (slabs + NL_SMALL_WARPS - 1) / NL_SMALL_WARPS;
    const long long cap = (long long)p->sm_count * 3;
    nl_small_kernel<N, MODEL><<<(unsigned)(ctas < cap ? ctas : cap), 32 * NL_SMALL_WARPS, NL_SMALL_WARPS * 512 * sizeof(cplx), stream>>>(p->d, j, force);
}
template <int N>
static void launch_nl_small(rks_plan* p, int j, int force, cudaStream_t stream) {
    switch (p->d.model) {
        case RKS_MODEL_UUX_RFFT: launch_nl_small_t<N, 1>(p, j, force, stream); break;
        case RKS_MODEL_NLS_FFT: launch_nl_small_t<N, 2>(p, j, force, stream); break;
        case RKS_MODEL_CUBIC_RFFT: launch_nl_small_t<N, 3>(p, j, force, stream); break;
        case RKS_MODEL_DERIV_FFT: launch_nl_small_t<N, 5>(p, j, force, stream); break;
        case RKS_MODEL_DERIV_RFFT_PAIR: launch_nl_small_t<N, 6>(p, j, force, stream); break;
        default: launch_nl_small_t<N, 4>(p, j, force, stream); break;
    }
}
// EXPERIMENT (RKS_RFFT_HALF=1): real-field models through the half-length forward transform
template <int W, int MODEL>
static void launch_nl_fast_real_t(rks_plan* p, int j, int force, cudaStream_t stream) {
    constexpr int RPC = 256 / (32 * W);
    const long long groups = (p->d.batch + RPC - 1) / RPC, resident = (long long)p->sm_count * 2;
    const unsigned grid = (unsigned)(groups < resident ? groups : resident);
    nl_fast_real_kernel<W, MODEL><<<grid, 256, (size_t)RPC * 512 * W * sizeof(cplx), stream>>>(p->d, j, force);
}
static void launch_nl_fast_real(rks_plan* p, int j, int force, cudaStream_t stream) {
    const bool uux = p->d.model == RKS_MODEL_UUX_RFFT;
    switch (p->d.n) {
        case 512: uux ? launch_nl_fast_real_t<1, 1>(p, j, force, stream) : launch_nl_fast_real_t<1, 3>(p, j, force, stream); break;
        case 1024: uux ? launch_nl_fast_real_t<2, 1>(p, j, force, stream) : launch_nl_fast_real_t<2, 3>(p, j, force, stream); break;
        case 2048: uux ? launch_nl_fast_real_t<4, 1>(p, j, force, stream) : launch_nl_fast_real_t<4, 3>(p, j, force, stream); break;
        default: uux ? launch_nl_fast_real_t<8, 1>(p, j, force, stream) : launch_nl_fast_real_t<8, 3>(p, j, force, stream); break;
    }
    p->launches += 1;
}

// cubic model: row pairs (fft_pair.cuh)
template <int W>
static void launch_nl_fast_pair_t(rks_plan* p, int j, int force, cudaStream_t stream) {
    constexpr int RPC = 256 / (32 * W);
    const long long pairs = (p->d.batch + 1) / 2;
    const long long groups = (pairs + RPC - 1) / RPC, resident = (long long)p->sm_count * 2;
    const unsigned grid = (unsigned)(groups < resident ? groups : resident);
    nl_fast_pair_kernel<W><<<grid, 256, (size_t)RPC * 512 * W * sizeof(cplx), stream>>>(p->d, j, force);
}
static void launch_nl_fast_pair(rks_plan* p, int j, int force, cudaStream_t stream) {
    switch (p->d.n) {
        case 512: launch_nl_fast_pair_t<1>(p, j, force, stream); break;
        case 1024: launch_nl_fast_pair_t<2>(p, j, force, stream); break;
        case 2048: launch_nl_fast_pair_t<4>(p, j, force, stream); break;
        default: launch_nl_fast_pair_t<8>(p, j, force, stream); break;
    }
    p->launches += 1;
}

static int launch_nl_nd(rks_plan* p, int j, int force, cudaStream_t stream);

static int launch_nl(rks_plan* p, int j, int force, cudaStream_t stream) {
    const DevPlan& d = p->d;
    if (p->nd_rows) return launch_nl_nd(p, j, force, stream);
    if (p->nl_small && !p->multi_n) {
        if (d.n == 64) launch_nl_small<64>(p, j, force, stream);
        else if (d.n == 128) launch_nl_small<128>(p, j, force, stream);
        else launch_nl_small<256>(p, j, force, stream);
        p->launches += 1;
        return RKS_OK;
    }
    if (p->nl_fast && p->pair_rows && !p->multi_n) {
        launch_nl_fast_pair(p, j, force, stream);
        return RKS_OK;
    }
    if (p->nl_fast && p->rfft_half && !p->multi_n) {
        launch_nl_fast_real(p, j, force, stream);
        return RKS_OK;
    }

    if (p->nl_fast) {
        dispatch_nl_fast(p, j, force, stream);
        return RKS_OK;
    }
    if (p->multi_n) {
        const dim3 g(1, 1, (unsigned)p->multi_n);
        if (d.model == RKS_MODEL_UUX_RFFT) nl_kernel_multi<1><<<g, p->nl_threads, p->nl_smem, stream>>>(p->multi_dev, j, force, p->nl_rows_per_cta);
        else if (d.model == RKS_MODEL_NLS_FFT) nl_kernel_multi<2><<<g, p->nl_threads, p->nl_smem, stream>>>(p->multi_dev, j, force, p->nl_rows_per_cta);
        else if (d.model == RKS_MODEL_CUBIC_RFFT) nl_kernel_multi<3><<<g, p->nl_threads, p->nl_smem, stream>>>(p->multi_dev, j, force, p->nl_rows_per_cta);
        else nl_kernel_multi<4><<<g, p->nl_threads, p->nl_smem, stream>>>(p->multi_dev, j, force, p->nl_rows_per_cta);
        p->launches += 1;
        return RKS_OK;
    }
    const unsigned grid = (unsigned)((d.batch + p->nl_rows_per_cta - 1) / p->nl_rows_per_cta);
    if (d.model == RKS_MODEL_UUX_RFFT) nl_kernel<1><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    else if (d.model == RKS_MODEL_NLS_FFT) nl_kernel<2><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    else if (d.model == RKS_MODEL_CUBIC_RFFT) nl_kernel<3><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    else if (d.model == RKS_MODEL_DERIV_FFT) nl_kernel<5><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    else if (d.model == RKS_MODEL_DERIV_RFFT_PAIR) nl_kernel<6><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    else nl_kernel<4><<<grid, p->nl_threads, p->nl_smem, stream>>>(d, j, force, p->nl_rows_per_cta);
    p->launches += 1;
    return RKS_OK;
}

extern "C" int rks_nl(rks_plan* p, int j, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (p->d.model == RKS_MODEL_NONE) return fail(RKS_ERR_UNSUPPORTED, "no fused model set (rks_set_model)");
    const int S = method_stages(p->method);
    const int jmax = method_fsal(p->method) ? S + 1 : S;
    if (j < 1 || j > jmax) return fail(RKS_ERR_ARG, "N index out of range");
    // fixed-step methods have no device predicate: the caller decides when N1 is (re)computed
    launch_nl(p, j, method_adaptive(p->method) ? 0 : 1, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// Stage s followed by the nonlinear evaluation it feeds: K1 then K4 as separate kernels (the north_star
// decomposition; folding the combine into K4's load prologue was measured slower and removed).
// Intermediate stage of a fast NLS-type plan: its value only feeds N(.), so K1 may hand it over pre-transformed
// (stage_pre_kernel -> nl_fast_pre_kernel).  The last stage is a state and keeps the natural layout.
static bool can_pretransform(const rks_plan* p, int s) {
    // n = 512: a row is one warp's slice, K4 is already warp-local there and the pair measured 5-16 % slower
    return p->pretransform && p->nl_fast && !p->nd_rows && p->d.n >= 1024 && !p->multi_n && p->d.lin_elems == p->d.n_c && p->d.coef_mode == CM_COLUMN
        && p->d.model == RKS_MODEL_NLS_FFT && s < method_stages(p->method);
}

// part: 0 = stage s and the evaluation it feeds, 1 = the stage kernel only, 2 = the evaluation only (the
// kernels of part 0, separately: per-kernel timing in bench.py)
static int stage_nl_parts(rks_plan* p, int s, int part, void* stream_v) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (p->d.model == RKS_MODEL_NONE) return fail(RKS_ERR_UNSUPPORTED, "no fused model set (rks_set_model)");
    const int m = p->method, S = method_stages(m);
    if (s < 1 || s > S) return fail(RKS_ERR_ARG, "stage out of range");
    const bool adapt = method_adaptive(m);
    // which N the evaluation after stage s produces: N_{s+1}; after the last stage N1 (fixed step,
    // etd4.py:174) or N_last (FSAL methods, if34.py:129); none for ETD35
    const int j = s < S ? s + 1 : (adapt ? (method_fsal(m) ? S + 1 : 0) : 1);
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (can_pretransform(p, s)) {
        if (part != 2) launch_stage_pre(p, s, stream);
        if (part != 1) dispatch_nl_fast_pre(p, j, adapt ? 0 : 1, stream);
        CUDA_TRY(cudaGetLastError());
        return RKS_OK;
    }
    if (part != 2)
        if (int rc = rks_stage(p, s, stream_v)) return rc;
    return (j && part != 1) ? rks_nl(p, j, stream_v) : RKS_OK;
}
extern "C" int rks_stage_nl(rks_plan* p, int s, void* stream) { return stage_nl_parts(p, s, 0, stream); }
extern "C" int rks_stage_nl_part(rks_plan* p, int s, int part, void* stream) {
    if (part < 1 || part > 2) return fail(RKS_ERR_ARG, "part must be 1 (stage) or 2 (evaluation)");
    return stage_nl_parts(p, s, part, stream);
}

extern "C" void* rks_nl_input(rks_plan* p, int j) {
    if (!p) return nullptr;
    const int S = method_stages(p->method);
    const int u_sel = method_adaptive(p->method) ? p->roles_u_sel : 0;
    if (j == 1) return p->d.U[u_sel];
    if (j <= S) return p->d.K;
    return p->d.U[1 - u_sel];
}

extern "C" void* rks_nl_output(rks_plan* p, int j) {
    if (!p || j < 1 || j > method_nl_buffers(p->method)) return nullptr;
    const int n_sel = method_adaptive(p->method) ? p->roles_n_sel : 0;
    return p->d.NL[nl_phys(p->method, j, n_sel)];
}

// ---------------------------------------------------------------------------------------
// K3 dispatch
// ---------------------------------------------------------------------------------------
template <int M, typename CT>
static void launch_norm_t(rks_plan* p, int fuse, cudaStream_t stream) {
    const DevPlan& d = p->d;
    if (p->multi_n) {
        long long gx = (d.n_c + 127) / 128;
        if (gx > MULTI_NORM_BLOCKS) gx = MULTI_NORM_BLOCKS;
        norm_kernel_multi<M, CT, CM_FLAT><<<dim3((unsigned)gx, 1, (unsigned)p->multi_n), 128, 0, stream>>>(p->multi_dev, fuse);
        return;
    }
    const int mode = d.coef_mode;
    const long long sep_cols = mode == CM_SEPARABLE ? d.sep_dims[d.sep_nd - 1] : 1;
    const long long ncols = mode == CM_FLAT ? d.batch * d.n_c : mode == CM_SEPARABLE ? sep_cols : d.n_c;
    const long long nrows = mode == CM_FLAT ? 1 : mode == CM_SEPARABLE ? d.batch * (d.n_c / sep_cols) : d.batch;
    // one wave of resident 128-thread CTAs (a partial second wave costs as much as a full one)
    // (occupancy of the variant that is launched: the separable / indexed ones hold more registers)
    static int per_sm_dev[MAX_DEVICES][4] = {{0}};
    int per_sm;
    {
        std::lock_guard<std::mutex> lock(g_attr_mutex);
        int& slot = per_sm_dev[p->device % MAX_DEVICES][mode & 3];
        if (!slot) {
            int occ = 0;
            if (mode == CM_FLAT) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, norm_kernel<M, CT, CM_FLAT>, 128, 0);
            else if (mode == CM_INDEXED) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, norm_kernel<M, CT, CM_INDEXED>, 128, 0);
            else if (mode == CM_SEPARABLE) {
                if constexpr (method_is_if(M)) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, norm_kernel<M, CT, CM_SEPARABLE>, 128, 0);
            } else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, norm_kernel<M, CT, CM_COLUMN>, 128, 0);
            slot = occ < 1 ? 1 : occ > 8 ? 8 : occ;
        }
        per_sm = slot;
    }
    const long long target = (long long)p->sm_count * per_sm;
    long long gx = (ncols + 127) / 128;
    if (gx > target) gx = target;
    long long gy = target / gx;
    if (gy < 1) gy = 1;
    if (gy > nrows) gy = nrows;
    while (gx * gy > NORM_MAX_BLOCKS) { if (gy > 1) --gy; else --gx; }
    const dim3 grid((unsigned)gx, (unsigned)gy);
    if (mode == CM_FLAT) norm_kernel<M, CT, CM_FLAT><<<grid, 128, 0, stream>>>(d, fuse);
    else if (mode == CM_INDEXED) norm_kernel<M, CT, CM_INDEXED><<<grid, 128, 0, stream>>>(d, fuse);
    else if (mode == CM_SEPARABLE) {
        if constexpr (method_is_if(M)) norm_kernel<M, CT, CM_SEPARABLE><<<grid, 128, 0, stream>>>(d, fuse);
    } else norm_kernel<M, CT, CM_COLUMN><<<grid, 128, 0, stream>>>(d, fuse);
}

static int launch_norm(rks_plan* p, int fuse, cudaStream_t stream) {
    const bool cx = p->d.lin_complex != 0;
    switch (p->method) {
        case M_IF34: cx ? launch_norm_t<M_IF34, cplx>(p, fuse, stream) : launch_norm_t<M_IF34, double>(p, fuse, stream); break;
        case M_IF45DP: cx ? launch_norm_t<M_IF45DP, cplx>(p, fuse, stream) : launch_norm_t<M_IF45DP, double>(p, fuse, stream); break;
        case M_ETD34: launch_norm_t<M_ETD34, cplx>(p, fuse, stream); break;
        case M_ETD35: launch_norm_t<M_ETD35, cplx>(p, fuse, stream); break;
        default: return fail(RKS_ERR_UNSUPPORTED, "fixed-step methods have no error control");
    }
    p->launches += 1;
    return RKS_OK;
}

extern "C" int rks_error_control(rks_plan* p, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    const int rc = launch_norm(p, 1, (cudaStream_t)stream);
    p->d.norm_u = nullptr;                                 // an override lasts for one trial
    if (rc) return rc;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// The NEXT rks_error_control takes |u+| (max, mask, tolerance) from `u_phys` instead of the plan's own u+,
// while the error estimate stays the plan's: the reference's diagonalize=True strategies return the physical
// S k next to the eigenbasis estimate (etd35.py:495, etd34.py:300, if34.py:200) and _compute_s mixes the two.
extern "C" int rks_norm_override(rks_plan* p, const void* u_phys, void* stream) {
    if (!p || !u_phys) return fail(RKS_ERR_ARG, "null argument");
    if (!method_adaptive(p->method) || p->multi_n) return fail(RKS_ERR_UNSUPPORTED, "norm override needs a shared-dt adaptive plan");
    if (((uintptr_t)u_phys) & 15) return fail(RKS_ERR_ARG, "u_phys must be 16-byte aligned");
    max_abs2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p->d, (const cplx*)u_phys, p->d.batch * p->d.n_c);
    p->launches += 1;
    p->d.norm_u = (const cplx*)u_phys;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_error_sums(rks_plan* p, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (int rc = launch_norm(p, 0, (cudaStream_t)stream)) return rc;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_controller(rks_plan* p, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (!method_adaptive(p->method)) return fail(RKS_ERR_UNSUPPORTED, "fixed-step methods have no controller");
    if (p->multi_n) return fail(RKS_ERR_UNSUPPORTED, "independent-dt plans run the controller inside rks_error_control");
    controller_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p->d);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" double* rks_reduction_scalars(rks_plan* p) { return p ? p->d.ctrl->red : nullptr; }

// ---------------------------------------------------------------------------------------
// snapshots, whole trials, whole fixed steps
// ---------------------------------------------------------------------------------------
extern "C" int rks_snapshot(rks_plan* p, void* ring, double* ring_t, int cap, void* stream) {
    if (!p || !ring || !ring_t || cap < 1) return fail(RKS_ERR_ARG, "bad snapshot ring");
    if (p->multi_n) return fail(RKS_ERR_UNSUPPORTED, "snapshots are not available with independent dt");
    snapshot_kernel<<<copy_grid(p), 256, 0, (cudaStream_t)stream>>>(p->d, (cplx*)ring, ring_t, cap);
    p->launches += 1;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

static int enqueue_trial(rks_plan* p, void* ring, double* ring_t, int cap, void* stream) {
    const int m = p->method, S = method_stages(m);
    int rc;
    if ((rc = rks_update_coeffs(p, stream))) return rc;
    if ((rc = rks_nl(p, 1, stream))) return rc;                 // runs only when ctrl.need_n1
    for (int s = 1; s <= S; ++s)
        if ((rc = rks_stage_nl(p, s, stream))) return rc;
    if ((rc = rks_error_control(p, stream))) return rc;
    if (ring) return rks_snapshot(p, ring, ring_t, cap, stream);
    return RKS_OK;
}

// One trial / one fixed step is ~12 launches whose arguments never change (roles, h and predicates
// are read from the control block on the device), so the sequence is captured once into a CUDA
// graph and replayed: small problems are launch-latency bound (BASELINE cfg 1: 8 KB of state).
static int graph_replay(rks_plan* p, int count, int adaptive, void* ring, double* ring_t, int cap, cudaStream_t stream) {
    const int S = method_stages(p->method);
    GraphCache& g = p->graph;
    if (g.exec && (g.ring != ring || g.ring_t != ring_t || g.cap != cap)) {
        cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
    }
    if (!g.exec) {
        // capture on a plan-owned stream (the caller's may be the legacy default stream, which cannot
        // capture); the instantiated graph is then launched into the caller's stream
        if (!g.stream) CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        const long long before = p->launches;
        CUDA_TRY(cudaStreamBeginCapture(g.stream, cudaStreamCaptureModeRelaxed));
        int rc = RKS_OK;
        if (adaptive) rc = enqueue_trial(p, ring, ring_t, cap, g.stream);
        else
            for (int s = 1; s <= S && rc == RKS_OK; ++s) rc = rks_stage_nl(p, s, g.stream);
        cudaError_t e = cudaStreamEndCapture(g.stream, &graph);
        g.kernels = (int)(p->launches - before);
        p->launches = before;
        if (rc != RKS_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(RKS_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { g.exec = nullptr; return fail(RKS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
        g.ring = ring; g.ring_t = ring_t; g.cap = cap;
    }
    for (int i = 0; i < count; ++i) CUDA_TRY(cudaGraphLaunch(g.exec, stream));
    p->launches += (long long)count * g.kernels;
    return RKS_OK;
}

extern "C" int rks_run_trials(rks_plan* p, int ntrials, void* ring, double* ring_t, int cap, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (!method_adaptive(p->method)) return fail(RKS_ERR_UNSUPPORTED, "rks_run_trials needs an adaptive method");
    if (p->d.model == RKS_MODEL_NONE) return fail(RKS_ERR_UNSUPPORTED, "no fused model set (rks_set_model)");
    if (p->use_graph) return graph_replay(p, ntrials, 1, ring, ring_t, cap, (cudaStream_t)stream);
    for (int i = 0; i < ntrials; ++i)
        if (int rc = enqueue_trial(p, ring, ring_t, cap, stream)) return rc;
    return RKS_OK;
}

extern "C" int rks_run_fixed(rks_plan* p, int nsteps, void* stream) {
    if (!p) return fail(RKS_ERR_ARG, "plan is null");
    if (method_adaptive(p->method)) return fail(RKS_ERR_UNSUPPORTED, "rks_run_fixed needs a fixed-step method");
    if (p->d.model == RKS_MODEL_NONE) return fail(RKS_ERR_UNSUPPORTED, "no fused model set (rks_set_model)");
    if (p->use_graph) return graph_replay(p, nsteps, 0, nullptr, nullptr, 0, (cudaStream_t)stream);
    const int S = method_stages(p->method);
    int rc;
    for (int i = 0; i < nsteps; ++i) {
        // after the last stage U holds u+ and N1 <- N(u+)  (etd4.py:174)
        for (int s = 1; s <= S; ++s)
            if ((rc = rks_stage_nl(p, s, stream))) return rc;
    }
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// syncing reads
// ---------------------------------------------------------------------------------------
extern "C" int rks_read_ctrl(rks_plan* p, rks_ctrl_host* out, void* stream_v) {
    if (!p || !out) return fail(RKS_ERR_ARG, "null argument");
    cudaStream_t stream = (cudaStream_t)stream_v;
    int running = 0;
    if (p->multi_n) {
        // aggregate view: fields of row 0, status RUNNING while any row is still stepping
        CUDA_TRY(cudaMemsetAsync(p->multi_count_dev, 0, 2 * sizeof(int), stream));
        count_running_kernel<<<(unsigned)((p->multi_n + 255) / 256), 256, 0, stream>>>(p->multi_dev, (int)p->multi_n, p->multi_count_dev);
        p->launches += 1;
        CUDA_TRY(cudaMemcpyAsync(p->multi_count_host, p->multi_count_dev, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    }
    CUDA_TRY(cudaMemcpyAsync(p->pinned_raw, p->d.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (p->multi_n) running = p->multi_count_host[0];
    Ctrl& c = *p->pinned_raw;
    if (p->multi_n) c.status = running > 0 ? ST_RUNNING : p->multi_count_host[1];
    out->h = c.h; out->h_last = c.h_last; out->h_coeff = c.h_coeff; out->t = c.t; out->tf = c.tf;
    out->s_last = c.s_last;
    out->step_count = c.step_count; out->trial_count = c.trial_count; out->nl_evals = c.nl_evals;
    out->coeff_updates = c.coeff_updates;
    out->status = c.status; out->accept = c.accept; out->numloops = c.numloops;
    out->u_sel = c.u_sel; out->n_sel = c.n_sel; out->need_n1 = c.need_n1;
    out->log_count = c.log_count; out->snap_count = c.snap_count;
    p->roles_u_sel = c.u_sel;
    p->roles_n_sel = c.n_sel;
    return RKS_OK;
}

static void ctrl_to_host(const Ctrl& c, rks_ctrl_host* out) {
    out->h = c.h; out->h_last = c.h_last; out->h_coeff = c.h_coeff; out->t = c.t; out->tf = c.tf;
    out->s_last = c.s_last;
    out->step_count = c.step_count; out->trial_count = c.trial_count; out->nl_evals = c.nl_evals;
    out->coeff_updates = c.coeff_updates;
    out->status = c.status; out->accept = c.accept; out->numloops = c.numloops;
    out->u_sel = c.u_sel; out->n_sel = c.n_sel; out->need_n1 = c.need_n1;
    out->log_count = c.log_count; out->snap_count = c.snap_count;
}

// independent-dt ensembles: control block of every row, and the last trial records of one row
extern "C" int rks_read_rows(rks_plan* p, rks_ctrl_host* out, int64_t nrows, void* stream_v) {
    if (!p || !out || !p->multi_n || nrows != p->multi_n) return fail(RKS_ERR_ARG, "rks_read_rows needs an independent-dt plan and one slot per row");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CUDA_TRY(cudaMemcpyAsync(p->multi_ctrl_host, p->multi_ctrl_base, p->multi_ctrl_stride * (size_t)nrows, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int64_t r = 0; r < nrows; ++r)
        ctrl_to_host(*(const Ctrl*)((const unsigned char*)p->multi_ctrl_host + p->multi_ctrl_stride * (size_t)r), &out[r]);
    return RKS_OK;
}

extern "C" int rks_read_row_log(rks_plan* p, int64_t row, rks_trial_rec* out, int first, int count, void* stream_v) {
    if (!p || !out || !p->multi_n || row < 0 || row >= p->multi_n || first < 0 || count < 0 || count > MULTI_LOG_CAP)
        return fail(RKS_ERR_ARG, "bad row log range");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CUDA_TRY(cudaMemcpyAsync(p->pinned_log, p->multi_log_base + (size_t)row * MULTI_LOG_CAP, sizeof(TrialRec) * MULTI_LOG_CAP,
                             cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int i = 0; i < count; ++i) memcpy(&out[i], &p->pinned_log[(first + i) % MULTI_LOG_CAP], sizeof(TrialRec));
    return RKS_OK;
}

extern "C" int rks_read_log(rks_plan* p, rks_trial_rec* out, int first, int count, void* stream_v) {
    if (!p || !out || first < 0 || count < 0 || count > LOG_CAP) return fail(RKS_ERR_ARG, "bad log range");
    cudaStream_t stream = (cudaStream_t)stream_v;
    CUDA_TRY(cudaMemcpyAsync(p->pinned_log, p->d.log, sizeof(TrialRec) * LOG_CAP, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    for (int i = 0; i < count; ++i) memcpy(&out[i], &p->pinned_log[(first + i) % LOG_CAP], sizeof(TrialRec));
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// standalone fused row transform: out_row = F{ N( F^-1{in_row} ) } along the contiguous axis of
// any array (the innermost-axis part of an N-D nonlinear term; the outer axes are transformed by the
// caller).  Same kernels as rks_nl, no stepping plan / workspace.
// ---------------------------------------------------------------------------------------
struct rks_rows {
    rks_plan plan;            // only d (model fields), sm_count and the launch shape are used
    void* dev_mem = nullptr;
};

extern "C" int rks_rows_create(rks_rows** out, int model, int64_t n, const double* kx, double p0, void* stream_v) {
    if (!out) return fail(RKS_ERR_ARG, "out is null");
    *out = nullptr;
    if (model < RKS_MODEL_UUX_RFFT || model > RKS_MODEL_DERIV_RFFT_PAIR) return fail(RKS_ERR_ARG, "unknown model id");
    const bool deriv = model == RKS_MODEL_DERIV_FFT || model == RKS_MODEL_DERIV_RFFT_PAIR;
    if (n < 16 || n > MODEL_MAX_N || (n & (n - 1))) return fail(RKS_ERR_UNSUPPORTED, "n must be a power of two in [16, 16384]");
    if ((model == RKS_MODEL_UUX_RFFT || model == RKS_MODEL_SINE_GORDON || deriv) && !kx) return fail(RKS_ERR_ARG, "kx is null");
    cudaStream_t stream = (cudaStream_t)stream_v;
    rks_rows* r = new (std::nothrow) rks_rows();
    if (!r) return fail(RKS_ERR_ARG, "out of host memory");
    HandleGuard<rks_rows, rks_rows_destroy> guard{r};
    rks_plan* p = &r->plan;
    memset(&p->d, 0, sizeof(DevPlan));
    p->launches = 0; p->use_graph = false; p->pinned_raw = nullptr; p->pinned_log = nullptr;
    const bool half = model == RKS_MODEL_UUX_RFFT || model == RKS_MODEL_CUBIC_RFFT;
    const long long n_c = half ? n / 2 + 1 : n;
    const size_t tw_b = align_up(sizeof(cplx) * (size_t)n), twf_b = align_up(sizeof(cplx) * 2 * fast::TW_TOTAL);
    const size_t kx_elems = deriv ? 2 * (size_t)n : (size_t)n_c;      // derivative rows: n complex multipliers
    const size_t kx_b = align_up(sizeof(double) * kx_elems);
    CUDA_TRY(cudaGetDevice(&p->device));
    CUDA_TRY(usable_sm_count(&p->sm_count, p->device));
    CUDA_TRY(cudaMalloc(&r->dev_mem, tw_b + twf_b + kx_b));
    unsigned char* w = (unsigned char*)r->dev_mem;
    DevPlan& d = p->d;
    d.ctrl = nullptr;
    d.tw = (const cplx*)w; d.twf = (const cplx*)(w + tw_b); d.kx = (const double*)(w + tw_b + twf_b);
    d.n = n; d.n_c = n_c; d.lin_elems = n_c; d.batch = 1ll << 40; d.model = model; d.model_p0 = p0;   // batch: set per apply
    d.method = M_ETD4;
    int log2n = 0;
    while ((1ll << log2n) < n) ++log2n;
    d.log2n = log2n;
    twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((cplx*)w, (int)n);
    if (kx) CUDA_TRY(cudaMemcpyAsync(w + tw_b + twf_b, kx, sizeof(double) * kx_elems, cudaMemcpyDeviceToDevice, stream));
    if (int rc = prepare_nl_launch(p, model, n, (cplx*)(w + tw_b), stream)) return rc;
    // derivative rows: the frequency a position holds depends on the kernel family (fft_fast.cuh DigitMap)
    if (deriv) d.model_p0 = (p->nl_fast || p->nl_small) ? 0.0 : 1.0;
    CUDA_TRY(cudaGetLastError());
    guard.release();
    *out = r;
    return RKS_OK;
}

extern "C" int rks_rows_apply(rks_rows* r, const void* in, void* out, int64_t batch, void* stream) {
    if (!r || !in || !out || batch <= 0) return fail(RKS_ERR_ARG, "bad row-transform arguments");
    rks_plan* p = &r->plan;
    p->d.batch = batch;
    p->d.U[0] = (cplx*)in;
    p->d.NL[1] = (cplx*)out;
    // generic kernel: rows per CTA must not exceed THIS batch (the handle keeps its full launch shape)
    const int rpc0 = p->nl_rows_per_cta, thr0 = p->nl_threads;
    const size_t sm0 = p->nl_smem;
    if (!p->nl_fast) {
        while (p->nl_rows_per_cta > 1 && p->nl_rows_per_cta > batch) {
            p->nl_rows_per_cta >>= 1; p->nl_threads >>= 1; p->nl_smem >>= 1;
        }
    }
    launch_nl(p, 1, 1, (cudaStream_t)stream);
    p->nl_rows_per_cta = rpc0; p->nl_threads = thr0; p->nl_smem = sm0;
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" void rks_rows_destroy(rks_rows* r) {
    if (!r) return;
    cudaFree(r->dev_mem);
    delete r;
}

// ---------------------------------------------------------------------------------------
// strided-axis transforms of N-D grids (fft_axis.cuh)
// ---------------------------------------------------------------------------------------
struct rks_axis {
    cplx* tw = nullptr;
    long long n = 0;
    int sm_count = 0;
};

template <int N>
static cudaError_t axis_prepare() {
    const int smem = axis::tile_smem<N>();
    cudaError_t e = cudaFuncSetAttribute(axis_fft_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(axis_fft_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e;
}

template <int N>
static void axis_launch(const rks_axis* a, const cplx* in, cplx* out, long long outer, long long inner, int inverse,
                        long long ostride, long long bstride, int rb_shift, cudaStream_t stream) {
    constexpr int C = axis::tile_cols<N>();
    const size_t smem = (size_t)axis::tile_smem<N>();
    const long long tiles = outer * ((inner + C - 1) / C);
    const long long cap = (long long)a->sm_count * axis::tile_blocks<N>() * axis::tile_waves<N>();      // persistent CTAs, a few per slot
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    if (inverse) axis_fft_kernel<N, true><<<grid, axis::tile_threads<N>(), smem, stream>>>(in, out, outer, inner, a->tw, 1.0 / (double)N, ostride, bstride, rb_shift);
    else axis_fft_kernel<N, false><<<grid, axis::tile_threads<N>(), smem, stream>>>(in, out, outer, inner, a->tw, 1.0, ostride, bstride, rb_shift);
}

extern "C" int rks_axis_create(rks_axis** out, int64_t n, void* stream_v) {
    if (!out) return fail(RKS_ERR_ARG, "out is null");
    *out = nullptr;
    if (n < 16 || n > 4096 || (n & (n - 1))) return fail(RKS_ERR_UNSUPPORTED, "axis length must be a power of two in [16, 4096]");
    rks_axis* a = new (std::nothrow) rks_axis();
    if (!a) return fail(RKS_ERR_ARG, "out of host memory");
    HandleGuard<rks_axis, rks_axis_destroy> guard{a};
    a->n = n;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(usable_sm_count(&a->sm_count, dev));
    CUDA_TRY(cudaMalloc((void**)&a->tw, sizeof(cplx) * (size_t)n));
    twiddle_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_v>>>(a->tw, (int)n);
    cudaError_t e = cudaSuccess;
    switch (n) {
        case 16: e = axis_prepare<16>(); break;
        case 32: e = axis_prepare<32>(); break;
        case 64: e = axis_prepare<64>(); break;
        case 128: e = axis_prepare<128>(); break;
        case 256: e = axis_prepare<256>(); break;
        case 512: e = axis_prepare<512>(); break;
        case 1024: e = axis_prepare<1024>(); break;
        case 2048: e = axis_prepare<2048>(); break;
        default: e = axis_prepare<4096>(); break;
    }
    if (e != cudaSuccess) return fail(RKS_ERR_CUDA, "axis kernel attributes: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaGetLastError());
    guard.release();
    *out = a;
    return RKS_OK;
}

static int axis_apply(rks_axis* a, const void* in, void* out, int64_t outer, int64_t inner, int inverse,
                      long long ostride, long long bstride, int rb_shift, void* stream_v);

extern "C" int rks_axis_apply(rks_axis* a, const void* in, void* out, int64_t outer, int64_t inner, int inverse, void* stream_v) {
    if (!a) return fail(RKS_ERR_ARG, "axis handle is null");
    return axis_apply(a, in, out, outer, inner, inverse, a->n * inner, 0, 31, stream_v);
}

// The axis is split into `chunks` equal row blocks that are stored chunk-major:
// [chunks][outer][n / chunks][inner] -- the layout an all-to-all of a slab decomposition delivers.
extern "C" int rks_axis_apply_chunked(rks_axis* a, const void* in, void* out, int64_t outer, int64_t inner, int64_t chunks,
                                      int inverse, void* stream_v) {
    if (!a) return fail(RKS_ERR_ARG, "axis handle is null");
    if (chunks < 1 || (chunks & (chunks - 1)) || chunks > a->n) return fail(RKS_ERR_ARG, "chunks must be a power of two <= n");
    const long long rb = a->n / chunks;
    int sh = 0;
    while ((1ll << sh) < rb) ++sh;
    if (chunks == 1) sh = 31;
    return axis_apply(a, in, out, outer, inner, inverse, rb * inner, outer * rb * inner, sh, stream_v);
}

static int axis_apply(rks_axis* a, const void* in, void* out, int64_t outer, int64_t inner, int inverse,
                      long long ostride, long long bstride, int rb_shift, void* stream_v) {
    if (!in || !out || outer <= 0 || inner <= 0) return fail(RKS_ERR_ARG, "bad axis-transform arguments");
    if (((uintptr_t)in | (uintptr_t)out) & 15) return fail(RKS_ERR_ARG, "axis arrays must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_v;
    const cplx* i = (const cplx*)in;
    cplx* o = (cplx*)out;
    switch (a->n) {
        case 16: axis_launch<16>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 32: axis_launch<32>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 64: axis_launch<64>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 128: axis_launch<128>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 256: axis_launch<256>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 512: axis_launch<512>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 1024: axis_launch<1024>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        case 2048: axis_launch<2048>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
        default: axis_launch<4096>(a, i, o, outer, inner, inverse, ostride, bstride, rb_shift, stream); break;
    }
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

template <int N>
static void axis_launch_scatter(const rks_axis* a, const cplx* in, const long long* otab, long long outer, long long inner,
                                int inverse, long long ostride, long long bstride, int rb_shift, int o_shift, cudaStream_t stream) {
    constexpr int C = axis::tile_cols<N>();
    const size_t smem = (size_t)axis::tile_smem<N>();
    const long long tiles = outer * ((inner + C - 1) / C);
    const long long cap = (long long)a->sm_count * axis::tile_blocks<N>() * axis::tile_waves<N>();
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    // (per device and cheap: set on every launch rather than cached)
    if (inverse) cudaFuncSetAttribute(axis_fft_scatter_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else cudaFuncSetAttribute(axis_fft_scatter_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (inverse) axis_fft_scatter_kernel<N, true><<<grid, axis::tile_threads<N>(), smem, stream>>>(in, outer, inner, a->tw, 1.0 / (double)N, ostride, bstride, rb_shift, otab, o_shift);
    else axis_fft_scatter_kernel<N, false><<<grid, axis::tile_threads<N>(), smem, stream>>>(in, outer, inner, a->tw, 1.0, ostride, bstride, rb_shift, otab, o_shift);
}

extern "C" int rks_axis_apply_scatter(rks_axis* a, const void* in, const int64_t* out_bases, int64_t outer, int64_t inner,
                                      int64_t in_chunks, int64_t out_chunks, int inverse, void* stream_v) {
    if (!a) return fail(RKS_ERR_ARG, "axis handle is null");
    if (!in || !out_bases || outer <= 0 || inner <= 0) return fail(RKS_ERR_ARG, "bad axis-transform arguments");
    if ((uintptr_t)in & 15) return fail(RKS_ERR_ARG, "axis arrays must be 16-byte aligned");
    if (in_chunks < 1 || (in_chunks & (in_chunks - 1)) || in_chunks > a->n) return fail(RKS_ERR_ARG, "in_chunks must be a power of two <= n");
    if (out_chunks < 1 || (out_chunks & (out_chunks - 1)) || out_chunks > a->n) return fail(RKS_ERR_ARG, "out_chunks must be a power of two <= n");
    auto log2i = [](long long v) { int s = 0; while ((1ll << s) < v) ++s; return s; };
    const long long rb = a->n / in_chunks;
    const int rb_shift = in_chunks == 1 ? 31 : log2i(rb);
    const long long ostride = in_chunks == 1 ? a->n * inner : rb * inner;
    const long long bstride = in_chunks == 1 ? 0 : outer * rb * inner;
    const int o_shift = log2i(a->n / out_chunks);
    cudaStream_t stream = (cudaStream_t)stream_v;
    const cplx* i = (const cplx*)in;
    const long long* tab = (const long long*)out_bases;
#define RKS_CALL(N) axis_launch_scatter<N>(a, i, tab, outer, inner, inverse, ostride, bstride, rb_shift, o_shift, stream)
    switch (a->n) {
        case 16: RKS_CALL(16); break; case 32: RKS_CALL(32); break; case 64: RKS_CALL(64); break; case 128: RKS_CALL(128); break;
        case 256: RKS_CALL(256); break; case 512: RKS_CALL(512); break; case 1024: RKS_CALL(1024); break;
        case 2048: RKS_CALL(2048); break; default: RKS_CALL(4096); break;
    }
#undef RKS_CALL
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

extern "C" int rks_peer_barrier(const int64_t* flag_bases, int world, int rank, uint64_t epoch, void* stream_v) {
    if (!flag_bases || world < 1 || world > 32 || rank < 0 || rank >= world) return fail(RKS_ERR_ARG, "bad peer-barrier arguments");
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream_v>>>((const long long*)flag_bases, world, rank, (unsigned long long)epoch);
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// N-D grid models as engine models: N_j of a 2-D / 3-D spectral grid = inverse transforms over the strided axes
// (stage value -> N_j, then in place), the fused last-axis kernel on the rows of N_j, forward transforms over the
// strided axes.  Every kernel resolves its arrays and its run predicate from the control block, so whole trials
// are enqueued (and graph-replayed) without a host sync: rks_run_trials works for cfg 4 / cfg 5 like for 1-D rows.
// Replaces the N-D nl_func closures of demos/nls.ipynb:496-511 inside the trial loop solveras.py:379-410.
// ---------------------------------------------------------------------------------------
template <int N>
static void axis_launch_plan(const rks_axis* a, const DevPlan& d, int j, int force, int first, long long outer, long long inner,
                             int inverse, cudaStream_t stream) {
    constexpr int C = axis::tile_cols<N>();
    const size_t smem = (size_t)axis::tile_smem<N>();
    const long long tiles = outer * ((inner + C - 1) / C);
    const long long cap = (long long)a->sm_count * axis::tile_blocks<N>() * axis::tile_waves<N>();
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    if (inverse) axis_fft_plan_kernel<N, true><<<grid, axis::tile_threads<N>(), smem, stream>>>(d, j, force, first, outer, inner, a->tw, 1.0 / (double)N);
    else axis_fft_plan_kernel<N, false><<<grid, axis::tile_threads<N>(), smem, stream>>>(d, j, force, first, outer, inner, a->tw, 1.0);
}
template <int N>
static cudaError_t axis_plan_prepare() {
    const int smem = axis::tile_smem<N>();
    cudaError_t e = cudaFuncSetAttribute(axis_fft_plan_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(axis_fft_plan_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return e;
}
#define RKS_AXIS_SWITCH(n, CALL) \
    switch (n) { \
        case 16: CALL(16); break; case 32: CALL(32); break; case 64: CALL(64); break; case 128: CALL(128); break; \
        case 256: CALL(256); break; case 512: CALL(512); break; case 1024: CALL(1024); break; \
        case 2048: CALL(2048); break; default: CALL(4096); break; }

static void axis_step(rks_plan* p, int which, int j, int force, int first, long long outer, long long inner, int inverse,
                      cudaStream_t stream) {
    const rks_axis* a = p->nd_axes[which];
#define RKS_CALL(N) axis_launch_plan<N>(a, p->d, j, force, first, outer, inner, inverse, stream)
    RKS_AXIS_SWITCH(a->n, RKS_CALL)
#undef RKS_CALL
    p->launches += 1;
}

static int launch_nl_nd(rks_plan* p, int j, int force, cudaStream_t stream) {
    const DevPlan& d = p->d;
    const long long s0 = p->nd_spec[0], s1 = p->nd_spec[1], s2 = p->nd_spec[2];
    const long long slast = p->nd == 2 ? s1 : s2;
    axis_step(p, 0, j, force, 1, d.batch, d.n_c / s0, 1, stream);
    if (p->nd == 3) axis_step(p, 1, j, force, 0, d.batch * s0, s2, 1, stream);
    // rows of N_j, in place: the row kernels take the plan's arrays, control block and roles
    rks_plan* rp = &p->nd_rows->plan;
    DevPlan& r = rp->d;
    r.ctrl = d.ctrl; r.log = d.log; r.K = d.K; r.ERR = d.ERR;
    r.U[0] = d.U[0]; r.U[1] = d.U[1];
    for (int q = 0; q < 8; ++q) r.NL[q] = d.NL[q];
    r.method = d.method; r.batch = d.batch * (d.n_c / slast); r.nd_inplace = 1;
    const long long before = rp->launches;
    if (int rc = launch_nl(rp, j, force, stream)) return rc;
    p->launches += rp->launches - before;
    if (p->nd == 3) axis_step(p, 1, j, force, 0, d.batch * s0, s2, 0, stream);
    axis_step(p, 0, j, force, 0, d.batch, d.n_c / s0, 0, stream);
    return RKS_OK;
}

extern "C" int rks_set_model_nd(rks_plan* p, int model, int nd, const int64_t* grid, double p0, void* stream_v) {
    if (!p || !grid) return fail(RKS_ERR_ARG, "null argument");
    if (p->multi_n) return fail(RKS_ERR_UNSUPPORTED, "N-D models need a shared-dt plan");
    if (nd != 2 && nd != 3) return fail(RKS_ERR_UNSUPPORTED, "N-D models take 2 or 3 grid dimensions");
    if (model != RKS_MODEL_NLS_FFT && model != RKS_MODEL_CUBIC_RFFT)
        return fail(RKS_ERR_UNSUPPORTED, "N-D models: RKS_MODEL_NLS_FFT (complex field) or RKS_MODEL_CUBIC_RFFT (real field)");
    const bool half = model == RKS_MODEL_CUBIC_RFFT;
    long long spec[3] = {0, 0, 0}, modes = 1;
    for (int k = 0; k < nd; ++k) {
        const long long n = grid[k];
        const bool last = k == nd - 1;
        if (n < 16 || (n & (n - 1)) || n > (last ? MODEL_MAX_N : 4096))
            return fail(RKS_ERR_UNSUPPORTED, "grid axes must be powers of two: 16..4096, last axis 16..16384");
        spec[k] = last && half ? n / 2 + 1 : n;
        modes *= spec[k];
    }
    if (modes != p->d.n_c) return fail(RKS_ERR_ARG, "grid does not match the plan's modes per trajectory");
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (p->graph.exec) { cudaGraphExecDestroy(p->graph.exec); p->graph.exec = nullptr; }
    if (p->nd_rows) return fail(RKS_ERR_UNSUPPORTED, "the plan already has an N-D model");
    if (int rc = rks_axis_create(&p->nd_axes[0], grid[0], stream_v)) return rc;
    if (nd == 3) {
        if (grid[1] == grid[0]) p->nd_axes[1] = p->nd_axes[0];
        else if (int rc = rks_axis_create(&p->nd_axes[1], grid[1], stream_v)) return rc;
    }
    for (int k = 0; k < nd - 1; ++k) {
        cudaError_t e = cudaSuccess;
#define RKS_CALL(N) e = axis_plan_prepare<N>()
        RKS_AXIS_SWITCH(grid[k], RKS_CALL)
#undef RKS_CALL
        if (e != cudaSuccess) return fail(RKS_ERR_CUDA, "axis kernel attributes: %s", cudaGetErrorString(e));
    }
    if (int rc = rks_rows_create(&p->nd_rows, model, grid[nd - 1], nullptr, p0, stream_v)) return rc;
    p->nd = nd;
    for (int k = 0; k < 3; ++k) p->nd_spec[k] = spec[k];
    p->d.model = model; p->d.model_p0 = p0; p->d.n = grid[nd - 1];
    p->nl_fast = false; p->nl_small = false; p->pretransform = false;
    p->rfft_half = false;
    p->pair_rows = false;
    CUDA_TRY(cudaGetLastError());
    (void)stream;
    return RKS_OK;
}

extern "C" void rks_axis_destroy(rks_axis* a) {
    if (!a) return;
    cudaFree(a->tw);
    delete a;
}

// ---------------------------------------------------------------------------------------
// pointwise nonlinearities for N-D models whose transforms are done by a library FFT
// ---------------------------------------------------------------------------------------
extern "C" int rks_pointwise(int model, const void* in, void* out, int64_t count, double p0, void* stream_v) {
    if (!in || !out || count <= 0) return fail(RKS_ERR_ARG, "bad pointwise arguments");
    if (((uintptr_t)in | (uintptr_t)out) & 15) return fail(RKS_ERR_ARG, "pointwise arrays must be 16-byte aligned");
    cudaStream_t stream = (cudaStream_t)stream_v;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (count + 255) / 256;
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
    if (model == RKS_MODEL_NLS_FFT) pointwise_nls_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const cplx*)in, (cplx*)out, count, p0);
    else if (model == RKS_MODEL_CUBIC_RFFT) pointwise_cubic_kernel<<<(unsigned)blocks, 256, 0, stream>>>((const double*)in, (double*)out, count, p0);
    else return fail(RKS_ERR_UNSUPPORTED, "no pointwise kernel for this model");
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// dense basis change of diagonalize=True (etd35.py:463, 495)
extern "C" int rks_gemv(const void* a, const void* x, void* y, int64_t n, int64_t batch, void* stream_v) {
    if (!a || !x || !y || n <= 0 || batch <= 0 || n > (1 << 20) || batch > 65535) return fail(RKS_ERR_ARG, "bad gemv arguments");
    if (x == y) return fail(RKS_ERR_ARG, "gemv: x and y must not alias");
    if (((uintptr_t)a | (uintptr_t)x | (uintptr_t)y) & 15) return fail(RKS_ERR_ARG, "gemv arrays must be 16-byte aligned");
    gemv_kernel<<<dim3((unsigned)((n + 3) / 4), (unsigned)batch), 128, 0, (cudaStream_t)stream_v>>>((const cplx*)a, (const cplx*)x,
                                                                                              (cplx*)y, (int)n);
    CUDA_TRY(cudaGetLastError());
    return RKS_OK;
}

// ---------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------
static int coef_slot(int m, const std::string& s) {
    static const char* kro_n[] = {"E", "E2", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54"};
    static const char* e5_n[] = {"E14", "E12", "E34", "E", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54",
                                 "a61", "a62", "a63", "a65", "a71", "a73", "a74", "a75", "a76"};
    static const char* if_n[] = {"E", "E2"};
    static const char* dp_n[] = {"E15", "E310", "E45", "E89", "E", "a21", "a31", "a32", "a41", "a42", "a43", "a51",
                                 "a52", "a53", "a54", "a61", "a62", "a63", "a64", "a65", "a71", "a73", "a74", "a75",
                                 "r1", "r3", "r4", "r5"};
    const char** names = (m == M_IF4 || m == M_IF34) ? if_n : (m == M_ETD4 || m == M_ETD34) ? kro_n
                       : (m == M_ETD5 || m == M_ETD35) ? e5_n : dp_n;
    for (int i = 0; i < method_ncoef(m); ++i)
        if (s == names[i]) return i;
    return -1;
}

extern "C" void* rks_array(rks_plan* p, const char* name) {
    if (!p || !name) return nullptr;
    const std::string s(name);
    const DevPlan& d = p->d;
    if (s == "U0") return d.U[0];
    if (s == "U1") return d.U[1];
    if (s == "K") return d.K;
    if (s == "ERR") return d.ERR;
    if (s == "ctrl") return d.ctrl;
    if (s == "tw") return (void*)d.tw;
    if (s == "row_logs") return p->multi_n ? (void*)p->multi_log_base : nullptr;
    if (s.size() == 2 && s[0] == 'N' && s[1] >= '1' && s[1] <= '7') {
        const int j = s[1] - '0';
        return j <= method_nl_buffers(p->method) ? d.NL[j] : nullptr;
    }
    const int slot = coef_slot(p->method, s);
    if (slot < 0 || d.coef_mode == CM_INDEXED || d.coef_mode == CM_SEPARABLE) return nullptr;    // no per-slot arrays there
    const bool real_coef = method_is_if(p->method) && !d.lin_complex;
    const size_t elem = real_coef ? sizeof(double) : sizeof(cplx);
    return (unsigned char*)d.coef + elem * (size_t)d.lin_elems * slot;
}

extern "C" int64_t rks_kernel_launches(rks_plan* p) { return p ? p->launches : 0; }
