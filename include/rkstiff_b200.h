/*
 * rkstiff_b200 -- C ABI of the B200 stepping engine for rkstiff's diagonal ETD/IF
 * Runge-Kutta hot path.
 *
 * The reference (whalenpt/rkstiff) is a pure-Python class library with no FFI.  The seam
 * this library replaces is the per-method "strategy" object every public solver class
 * holds (`self._method`), whose interface is
 *     update_coeffs(h)            rkstiff/etd35.py:157, etd34.py:84, etd4.py:87, etd5.py:115,
 *                                 if34.py:81, if4.py:72, if45dp.py:183
 *     n1_init(u) / stage_init(u)  rkstiff/etd4.py:141, if34.py:95, etd35.py:290
 *     update_stages(u[,h][,acc])  rkstiff/etd4.py:152, etd5.py:219, etd34.py:162, etd35.py:301,
 *                                 if4.py:96, if34.py:99, if45dp.py:112
 * plus the controller the adaptive base class runs around it
 *     _compute_s / _reject_step_size / _accept_step_size / step / evolve
 *                                 rkstiff/solveras.py:412-455, 457-506, 508-554, 336-410, 556-650
 * and the spectral nonlinear closures the demos/models pass as nl_func
 *                                 rkstiff/models.py:140-143, 189-192; README.md:94-97; demos/nls.ipynb.
 *
 * Conventions
 *  - plain C, no torch/C++ types; all array arguments are DEVICE pointers unless named *_host.
 *  - every state array is complex128, row-major (batch, n_c); n_c = modes per trajectory.
 *  - lin_op is float64 or complex128 with lin_elems == n_c (shared by the batch) or
 *    lin_elems == batch*n_c ("shaped like u").
 *  - every call returns 0 on success or a negative rks_status; rks_last_error() gives text.
 *  - a plan is bound to one device and is not thread safe; all work is enqueued on the
 *    cudaStream_t passed in (void* here so that the header needs no CUDA include).
 *  - device-side failures of the adaptive loop are reported through rks_ctrl_host.status,
 *    never through return codes (no host sync inside the stepping calls).
 */
#ifndef RKSTIFF_B200_H
#define RKSTIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RKS_ABI_VERSION 1

/* method ids (one per public reference class) */
enum rks_method {
    RKS_IF4 = 0,    /* rkstiff/if4.py    IF4    fixed step  */
    RKS_ETD4 = 1,   /* rkstiff/etd4.py   ETD4   fixed step  */
    RKS_ETD5 = 2,   /* rkstiff/etd5.py   ETD5   fixed step  */
    RKS_IF34 = 3,   /* rkstiff/if34.py   IF34   adaptive    */
    RKS_ETD34 = 4,  /* rkstiff/etd34.py  ETD34  adaptive    */
    RKS_ETD35 = 5,  /* rkstiff/etd35.py  ETD35  adaptive    */
    RKS_IF45DP = 6  /* rkstiff/if45dp.py IF45DP adaptive    */
};

/* fused spectral nonlinearities (K4) */
enum rks_model {
    RKS_MODEL_NONE = 0,
    RKS_MODEL_UUX_RFFT = 1, /* N = -c * rfft(irfft(u^) * irfft(i kx u^)); KS/Burgers c=1, KdV c=6
                               models.py:140-143,189-192, README.md:94-97. n_c = n/2+1, params[0]=c */
    RKS_MODEL_NLS_FFT = 2,  /* N = i*gamma * fft(|ifft u^|^2 ifft u^); demos/nls.ipynb. n_c = n, params[0]=gamma */
    RKS_MODEL_CUBIC_RFFT = 3,   /* N = c * rfft(irfft(u^)^3): Fourier-diagonal Allen-Cahn (c = -1, L = 1 - eps k^2, the
                                   split of models.py:240-244). n_c = n/2+1, params[0]=c, kx unused */
    RKS_MODEL_SINE_GORDON = 4,  /* psi = phi_t + i Omega phi: N = fft(phi - sin phi), phi^ = (psi^(k) - conj psi^(-k))/(2 i Omega);
                                   n_c = n, kx = Omega(k) = sqrt(1 + k^2) */
    /* Row transforms only (rks_rows_*; not stepping models): spectral derivatives with the K4 transform pair run the
     * other way round -- forward transform, multiplier, inverse transform (rkstiff/derivatives.py:47-179).  `kx` of
     * rks_rows_create points at n complex128 multipliers mult[k] = conj((i kx[k])^m) / n in FFT order. */
    RKS_MODEL_DERIV_FFT = 5,        /* dx_fft, derivatives.py:126-179: complex rows of n points, out = ifft(M fft(in)) */
    RKS_MODEL_DERIV_RFFT_PAIR = 6   /* dx_rfft, derivatives.py:47-123: real rows, two per complex transform: `in`/`out`
                                       are float64 arrays, row pair q = rows 2q, 2q+1 (batch counts PAIRS); the
                                       multiplier table must be Hermitian (mult[n-k] = conj mult[k], real at k = n/2) */
};

enum rks_status {
    RKS_OK = 0,
    RKS_ERR_ARG = -1,
    RKS_ERR_CUDA = -2,
    RKS_ERR_UNSUPPORTED = -3,
    RKS_ERR_WORKSPACE = -4
};

/* values of rks_ctrl_host.status (device controller state) */
enum rks_ctrl_status {
    RKS_CTRL_RUNNING = 0,
    RKS_CTRL_DONE = 1,       /* evolve: t >= tf reached; step mode: trial accepted */
    RKS_CTRL_MAX_LOOPS = 2,  /* solveras.py:399-403 -> MaxLoopsExceeded      */
    RKS_CTRL_MIN_STEP = 3    /* solveras.py:405-410 -> MinimumStepReached    */
};

/* SolverConfig (solveras.py:71-94) + ETDConfig (etd.py:81-131) scalars */
typedef struct rks_config {
    double epsilon, incr_f, decr_f, safety_f, adapt_cutoff, minh;
    double modecutoff, contour_radius;
    int32_t contour_points;
    int32_t if45dp_r4_fix; /* 0 (default): r4 = 17h e^{z/5}/1920 as shipped (if45dp.py:234); 1: 71/1920 */
} rks_config;

/* host copy of the device control block */
typedef struct rks_ctrl_host {
    double h;          /* step size the next trial will use                                  */
    double h_last;     /* step size of the last accepted trial (step(): "h" return value)     */
    double h_coeff;    /* step size the coefficient arrays were built for                     */
    double t, tf;
    double s_last;     /* last controller scale factor                                        */
    int64_t step_count;  /* accepted steps since rks_begin                                    */
    int64_t trial_count; /* all trials since rks_begin                                        */
    int64_t nl_evals;    /* fused NL evaluations actually executed (predicated ones excluded) */
    int64_t coeff_updates;
    int32_t status;    /* rks_ctrl_status */
    int32_t accept;    /* last trial accepted                                                */
    int32_t numloops;
    int32_t u_sel;     /* which of the two state buffers holds the current u                  */
    int32_t n_sel;     /* FSAL role swap of N1 / N_last                                       */
    int32_t need_n1;
    int32_t log_count;   /* trial records written since rks_begin (ring of RKS_LOG_CAP)       */
    int32_t snap_count;  /* snapshots written since rks_begin                                 */
} rks_ctrl_host;

#define RKS_LOG_CAP 4096
typedef struct rks_trial_rec {
    double h;        /* step size tried */
    double s;        /* scale factor    */
    double t_after;  /* time after the trial */
    int32_t accepted;
    int32_t pad;
} rks_trial_rec;

typedef struct rks_plan rks_plan;

int rks_abi_version(void);
const char* rks_last_error(void);

/* number of stage-combine kernels S and of N-buffers for a method */
int rks_num_stages(int method);
int rks_num_nl_buffers(int method);
int rks_is_adaptive(int method);

/* bytes of device workspace a plan needs (caller allocates; 256-byte aligned) */
size_t rks_workspace_bytes(int method, int64_t batch, int64_t n_c, int64_t lin_elems, int lin_is_complex);

/* Build a plan.  lin_op is copied into the workspace.  Replaces the strategy-object
 * constructors (etd35.py:115-155 etc.). */
int rks_plan_create(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* lin_op,
                    int lin_is_complex, int64_t lin_elems, const rks_config* cfg, void* workspace,
                    size_t workspace_bytes, void* stream);
void rks_plan_destroy(rks_plan* plan);

/* Large grids ("lin_op shaped like u" on 2-D/3-D spectral grids, demos/nls.ipynb:496-511): a Fourier grid has far
 * fewer DISTINCT lin_op values than modes (|k|^2 takes <= 3 (n/2)^2 values on n^3 points), and the reference's
 * update_coeffs (etd35.py:157-288, if45dp.py:183-238) is a pure function of z = h * lin_op.  These two plan
 * flavours therefore never build full-size coefficient arrays; every other call takes them unchanged.
 *
 * indexed: `values` = the n_values distinct lin_op entries (device), `index` = n_c int32 (device) mapping each
 *   mode to its entry.  K2 builds one coefficient record per distinct value, K1/K3 gather it through L2.
 *   Any method; batch <= 65535 trajectories share lin_op. */
size_t rks_workspace_bytes_indexed(int method, int64_t batch, int64_t n_c, int64_t n_values, int lin_is_complex);
int rks_plan_create_indexed(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* values,
                            int lin_is_complex, int64_t n_values, const int32_t* index, const rks_config* cfg,
                            void* workspace, size_t workspace_bytes, void* stream);
/* separable (IF4 / IF34 / IF45DP): lin_op[i0, i1(, i2)] = sum_d a_d[i_d] on a grid of `dims` (nd = 2 or 3, last
 *   axis contiguous).  Every IF coefficient is rational * h * exp(q z) (if4.py:72-83, if45dp.py:204-237), so
 *   K2 builds the per-axis tables exp(q h a_d) (sum(dims) entries per exponent) and K1/K3 multiply them.
 *   `axis_terms` = a_0 | a_1 (| a_2) concatenated (device), the constant term folded into a_0. */
size_t rks_workspace_bytes_separable(int method, int64_t batch, int nd, const int64_t* dims, int lin_is_complex);
int rks_plan_create_separable(rks_plan** out, int method, int64_t batch, int nd, const int64_t* dims,
                              const void* axis_terms, int lin_is_complex, const rks_config* cfg, void* workspace,
                              size_t workspace_bytes, void* stream);

/* Independent-dt ensemble (BASELINE cfg 2b): `batch` trajectories of one 1-D problem, each with its
 * own controller, dt, coefficient arrays and buffer roles -- B separate reference solvers
 * (solveras.py:279-325 run once per trajectory) stepped by one set of launches (gridDim.z = row).
 * Adaptive methods and fused models only; every other call takes the plan unchanged
 * (rks_begin / rks_set_h broadcast to all rows, rks_set_u / rks_get_u move the (batch, n_c) block,
 * rks_read_ctrl reports row 0 with status RUNNING until every row has finished, else the worst
 * status of any row).  rks_snapshot, rks_controller and caller-side N evaluation are not available. */
#define RKS_ROW_LOG_CAP 64
size_t rks_workspace_bytes_independent(int method, int64_t batch, int64_t n_c, int lin_is_complex);
int rks_plan_create_independent(rks_plan** out, int method, int64_t batch, int64_t n_c, const void* lin_op,
                                int lin_is_complex, const rks_config* cfg, void* workspace,
                                size_t workspace_bytes, void* stream);

/* SolverConfig/ETDConfig are read live by the reference on every trial (solveras.py:452-454) */
int rks_set_config(rks_plan* plan, const rks_config* cfg, void* stream);

/* Select the fused nonlinearity: n = real-space points per trajectory, kx = wavenumber array
 * (n_c doubles, device; copied).  Afterwards rks_nl() and rks_run_*() are available. */
int rks_set_model(rks_plan* plan, int model, int64_t n, const double* kx, const double* params_host,
                  int nparams, void* stream);

/* Fused nonlinearity of a 2-D / 3-D spectral grid (the fft2 / fftn closures of demos/nls.ipynb:496-511 that the
 * reference is given on flattened arrays): `grid` = real-space points per axis (nd = 2 or 3, powers of two,
 * 16..4096, last axis 16..16384), last axis contiguous.  RKS_MODEL_NLS_FFT: N = i p0 fftn(|f|^2 f), f = ifftn(u^),
 * spectral dims = grid;  RKS_MODEL_CUBIC_RFFT: N = p0 rfftn(irfftn(u^)^3), last spectral dim grid[nd-1]/2+1.
 * Evaluated by the engine's own kernels -- inverse transforms over the strided axes, the fused last-axis kernel,
 * forward transforms over the strided axes -- all predicated on the device, so rks_nl / rks_stage_nl /
 * rks_run_trials / rks_run_fixed work as for 1-D rows (no host sync per trial, graph replay). */
int rks_set_model_nd(rks_plan* plan, int model, int nd, const int64_t* grid, double p0, void* stream);

/* (re)start: clears FSAL state / cached h (BaseSolver*.reset + _reset, solveras.py:306-312,
 * etd35.py:836-840) and arms the controller.  step_mode != 0: stop after the first accepted
 * trial (step()); otherwise integrate until t >= tf (evolve(), solveras.py:605-650). */
int rks_begin(rks_plan* plan, double t0, double tf, double h, int64_t store_freq, int step_mode,
              int keep_fsal, void* stream);
int rks_set_h(rks_plan* plan, double h, void* stream);

/* state in/out (device-to-device copies of batch*n_c complex128) */
int rks_set_u(rks_plan* plan, const void* u, void* stream);
int rks_get_u(rks_plan* plan, void* u_out, void* stream);   /* current u (after accept: the new state) */

/* K2: coefficient arrays for ctrl.h; no-op on device when ctrl.h == ctrl.h_coeff (etd35.py:851) */
int rks_update_coeffs(rks_plan* plan, void* stream);
/* K1: stage-combine kernel `stage` in 1..S; the last one also emits err (ETD35) and max|u+|^2 */
int rks_stage(rks_plan* plan, int stage, void* stream);
/* K4: N_j = N(input_j), j in 1..S+1 (see DESIGN.md for the input of each j); predicated on device */
int rks_nl(rks_plan* plan, int j, void* stream);
/* K1+K4: stage `stage` and the nonlinear evaluation it feeds (N_{stage+1}; after the last stage N1
 * for fixed-step methods, N_last for FSAL methods, none for ETD35).  One fused kernel when the plan
 * has a fused model with n in 512..8192: the stage value k is formed in the load prologue of the
 * FFT kernel and never written to HBM unless it is the new state. */
int rks_stage_nl(rks_plan* plan, int stage, void* stream);
/* The two kernels rks_stage_nl(plan, stage) launches, separately: part 1 = the stage kernel, part 2 = the
 * evaluation.  For intermediate stages of a complex-field (NLS) plan with n in 1024..8192 these are the
 * pre-transforming pair: the stage kernel also applies the first inverse FFT pass and the evaluation starts
 * one pass later (RKS_PT=0 in the environment selects the plain pair).  Results equal
 * rks_stage + rks_nl; the stage value left in K by part 1 is in that intermediate layout. */
int rks_stage_nl_part(rks_plan* plan, int stage, int part, void* stream);
/* for caller-supplied nl_func: device pointers valid for the roles last read by rks_read_ctrl */
void* rks_nl_input(rks_plan* plan, int j);
void* rks_nl_output(rks_plan* plan, int j);
/* K3: masked norms -> s -> accept/reject -> new h, t, roles, status, log record */
int rks_error_control(rks_plan* plan, void* stream);
/* diagonalize=True (dense lin_op diagonalised on the host, the plan stepping the eigenbasis state): the next
 * rks_error_control takes max|u+|, the mask and the tolerance from `u_phys` = S u+ (device array of the plan's
 * state shape) and the error norm from the plan's own eigenbasis estimate -- the mix the reference's
 * _compute_s sees (etd35.py:495, solveras.py:451-454).  One-shot: cleared by that rks_error_control. */
int rks_norm_override(rks_plan* plan, const void* u_phys, void* stream);
/* K3 split for multi-GPU shared-dt ensembles: local partial results live in three doubles
 * (max|u+|^2, sum|u+|^2, sum|err|^2) the caller all-reduces between the calls. */
int rks_error_sums(rks_plan* plan, void* stream);
int rks_controller(rks_plan* plan, void* stream);
double* rks_reduction_scalars(rks_plan* plan);   /* device ptr to {umax2, sum_u2, sum_e2} */

/* whole trials / steps with the fused nonlinearity, no host sync */
int rks_run_trials(rks_plan* plan, int ntrials, void* snap_ring, double* snap_t, int snap_cap, void* stream);
int rks_run_fixed(rks_plan* plan, int nsteps, void* stream);

/* snapshot ring: copies the accepted state into ring[(snap_count-1) % cap] when the last
 * controller call asked for one (solveras.py:643-645) */
int rks_snapshot(rks_plan* plan, void* snap_ring, double* snap_t, int snap_cap, void* stream);

/* Pointwise nonlinearity alone, for N-D grids whose transforms are done by a library FFT between
 * the engine's kernels: RKS_MODEL_NLS_FFT: out = i*p0*|in|^2 in (count complex128);
 * RKS_MODEL_CUBIC_RFFT: out = p0*in^3 (count float64).  in == out is allowed.  No plan needed. */
int rks_pointwise(int model, const void* in, void* out, int64_t count, double p0, void* stream);

/* y[b] = A x[b], A a dense n x n complex128 matrix (row-major), x and y `batch` vectors of n complex128: the two
 * basis changes per nonlinear evaluation of diagonalize=True, N'(k) = S^-1 N(S k) (rkstiff/etd35.py:463,
 * etd34.py:286, if34.py:209), and the physical |S u+| the controller looks at (etd35.py:495).  x != y. */
int rks_gemv(const void* a, const void* x, void* y, int64_t n, int64_t batch, void* stream);

/* Standalone fused row transform: out_row = F{ N( F^-1{ in_row } ) } along the contiguous axis of any
 * (batch, n_c) complex128 array -- the innermost-axis part of an N-D nonlinear term (the caller
 * transforms the outer axes).  Same kernels as rks_nl; no stepping plan.  in == out is allowed. */
typedef struct rks_rows rks_rows;
int rks_rows_create(rks_rows** out, int model, int64_t n, const double* kx, double p0, void* stream);
int rks_rows_apply(rks_rows* rows, const void* in, void* out, int64_t batch, void* stream);
void rks_rows_destroy(rks_rows* rows);

/* N-D grids: transform along a STRIDED axis of a contiguous complex128 array viewed as
 * [outer][n][inner] (n a power of two in 16..4096), the outer-axis part of the reference's
 * `np.fft.ifft2 / fft2 / ifftn / fftn` calls in N-D nl_func closures (demos/nls.ipynb:500-508).
 * inverse != 0: unnormalised-by-nothing inverse transform (scaled 1/n), natural order in,
 * DIGIT-REVERSED order out along the axis; inverse == 0: forward transform taking that
 * digit-reversed order back to natural order.  A pointwise nonlinearity in between does not
 * depend on the order, so the pair replaces ifft / fft along the axis.  in == out is allowed. */
typedef struct rks_axis rks_axis;
int rks_axis_create(rks_axis** out, int64_t n, void* stream);
int rks_axis_apply(rks_axis* axis, const void* in, void* out, int64_t outer, int64_t inner, int inverse, void* stream);
/* same transform when the axis arrives split into `chunks` equal row blocks stored chunk-major,
 * [chunks][outer][n/chunks][inner] -- what the all-to-all of a slab decomposition delivers -- so the
 * transposes either side of the exchange never have to be materialised */
int rks_axis_apply_chunked(rks_axis* axis, const void* in, void* out, int64_t outer, int64_t inner, int64_t chunks,
                           int inverse, void* stream);
/* same transform with the OUTPUT rows scattered over the ranks of a slab decomposition (single large N-D grids,
 * the distributed transpose of `fftn` -- demos/nls.ipynb:496-511 on a sharded grid): the n output rows
 * are split into `out_chunks` equal blocks, block g laid out [outer][n/out_chunks][inner] at the device address
 * out_bases[g] (an int64 array in device memory).  The addresses may be PEER mappings of other GPUs' buffers
 * (NVLink): the stores of the transform's last level then are the all-to-all exchange, and no collective runs.
 * Input: plain [outer][n][inner] (in_chunks = 1) or chunk-major [in_chunks][outer][n/in_chunks][inner].
 * The caller orders the ranks (a barrier before the destination buffers are read or rewritten). */
int rks_axis_apply_scatter(rks_axis* axis, const void* in, const int64_t* out_bases, int64_t outer, int64_t inner,
                           int64_t in_chunks, int64_t out_chunks, int inverse, void* stream);
void rks_axis_destroy(rks_axis* axis);
/* Barrier between the ranks that share peer-mapped buffers (the order points of rks_axis_apply_scatter), enqueued on
 * `stream`: flag_bases[g] (int64 array in device memory) is the address, in rank g's memory, of `world` uint64 flags
 * that start at zero; `epoch` must grow by one per call on every rank.  Rank r stores epoch into flag r of every rank
 * (system-scope release: everything this stream did before is visible to whoever sees the flag) and waits until its
 * own flags all reach epoch.  One 32-thread block, one NVLink round trip. */
int rks_peer_barrier(const int64_t* flag_bases, int world, int rank, uint64_t epoch, void* stream);

/* the only syncing calls */
int rks_read_ctrl(rks_plan* plan, rks_ctrl_host* out, void* stream);
int rks_read_log(rks_plan* plan, rks_trial_rec* out_host, int first, int count, void* stream);
/* independent-dt plans: control block of every row; trial records of one row (ring of RKS_ROW_LOG_CAP) */
int rks_read_rows(rks_plan* plan, rks_ctrl_host* out_host, int64_t nrows, void* stream);
int rks_read_row_log(rks_plan* plan, int64_t row, rks_trial_rec* out_host, int first, int count, void* stream);

/* introspection for tests: device pointer of a named array ("E", "a21", ..., "N1".., "K", "ERR", "U0", "U1";
 * independent-dt plans also "row_logs": [batch][RKS_ROW_LOG_CAP] rks_trial_rec) */
void* rks_array(rks_plan* plan, const char* name);
int64_t rks_kernel_launches(rks_plan* plan);   /* kernels launched through this plan so far */

#ifdef __cplusplus
}
#endif
#endif /* RKSTIFF_B200_H */
