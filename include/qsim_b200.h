/*
 * qsim_b200.h -- C ABI of the B200-native time-evolution hot path.
 *
 * This is the drop-in boundary for QSimulator.jl's four solver entry points
 *   unitary_propagator  (reference: src/time_evolution.jl:29-47)
 *   unitary_state       (reference: src/time_evolution.jl:63-79)
 *   me_propagator       (reference: src/time_evolution.jl:96-123)
 *   me_state            (reference: src/time_evolution.jl:140-164)
 * batched over the points of a parameter sweep (one trajectory per point).
 *
 * The reference has no FFI: its "interface" for this path is the four Julia
 * signatures plus the term lists of a CompositeQSystem
 * (src/composite_systems.jl:7-16).  The boundary therefore sits where those
 * term lists are final: the host (Julia shim via ccall, or the Python mirror
 * via ctypes) embeds every operator into the full Hilbert space
 * (embed_indices, src/composite_systems.jl:136-148) and hands over
 *   - the static Hamiltonian H0                 (hamiltonian(cqs), :151-157)
 *   - one descriptor per time-dependent term    (parametric_Hs, :40-50)
 *   - the Lindblad operators                    (fixed_Ls / parametric_Ls, :53-62)
 * Julia closures cannot cross a C ABI, so time dependence is expressed with the
 * closed descriptor family below, which covers every form the reference's own
 * operators, tests, benchmarks and notebooks generate.
 *
 * Conventions
 *   - all matrices are column-major, interleaved (re,im) FP64 == Julia
 *     Matrix{ComplexF64} == double _Complex == cuDoubleComplex
 *   - indices are 0-based inside the ABI
 *   - H is in GHz (cycles/ns), t in ns; the kernels apply the explicit 2*pi
 *     (src/time_evolution.jl:35)
 *   - every function returns 0 on success, a negative qsim_status otherwise;
 *     the message is available from qsim_last_error()
 *   - the caller owns every buffer it passes; the library keeps no pointer
 *     after a call returns
 *   - there is NO CPU fallback: solve calls fail with QSIM_ERR_NO_DEVICE when
 *     no CUDA device is usable.
 */
#ifndef QSIM_B200_H
#define QSIM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSIM_ABI_VERSION 1

typedef enum {
    QSIM_OK = 0,
    QSIM_ERR_ARG = -1,        /* bad argument (message says which)            */
    QSIM_ERR_NO_DEVICE = -2,  /* no usable CUDA device: no CPU fallback       */
    QSIM_ERR_CUDA = -3,       /* CUDA runtime error                           */
    QSIM_ERR_UNSUPPORTED = -4,/* problem shape not covered by a kernel        */
    QSIM_ERR_STATE = -5,      /* call sequence error (e.g. not finalized)     */
    QSIM_ERR_ALLOC = -6
} qsim_status;

/* Per-trajectory solver outcome; mirrors OrdinaryDiffEq's retcodes, which the
 * reference ignores (it returns sol.u unconditionally, time_evolution.jl:44). */
typedef enum {
    QSIM_RET_SUCCESS = 0,
    QSIM_RET_MAXITERS = 1,
    QSIM_RET_DT_UNDERFLOW = 2,
    QSIM_RET_NONFINITE = 3
} qsim_retcode;

/* A scalar argument: a literal, or a column of the per-trajectory parameter
 * row (params[traj * n_params + param]) when param >= 0.                      */
typedef struct {
    int32_t param;   /* < 0: use `value`; >= 0: index into the parameter row  */
    int32_t _pad;
    double value;
} qsim_slot;

enum { QSIM_CARRIER_NONE = 0, QSIM_CARRIER_COS = 1, QSIM_CARRIER_SIN = 2 };
enum { QSIM_ENV_NONE = 0, QSIM_ENV_GAUSSIAN = 1, QSIM_ENV_ERF_SQUARED = 2, QSIM_ENV_TABLE = 3 };
enum { QSIM_BLEND_NONE = 0, QSIM_BLEND_REST = 1 };

/*
 * A real scalar drive s(t).  With
 *   carrier(t) = 1 | cos((2*pi*freq)*t + phase) | sin((2*pi*freq)*t + phase)
 *   env(t)     = 1
 *              | exp(-0.5*((t - t1)/sigma)^2)                        (gaussian)
 *              | 0.5*(erf((t - t1)/sigma) - erf((t - t2)/sigma))     (erf_squared)
 *              | p(t), a tabulated piecewise polynomial                (table; see qsim_problem_add_table)
 * blend == NONE:  s = offset + amp * env * carrier
 * blend == REST:  s = env * (offset + amp * carrier) + (1 - env) * rest
 * These are the drive shapes of benchmark/benchmarks.jl:111,138,185,
 * test/test_time_evolution.jl:45,107-113,130, test/integration_tests.jl:40,102
 * and the doc notebooks.
 */
typedef struct {
    int32_t carrier;
    int32_t envelope;
    int32_t blend;
    int32_t table;   /* QSIM_ENV_TABLE: id returned by qsim_problem_add_table; 0 otherwise */
    qsim_slot offset, amp, freq, phase, t1, t2, sigma, rest;
} qsim_signal;

enum {
    /* H(t) += s(t) * A.  rwa_dipole (operators.jl:242-250): one term per
     * quadrature; dipole_drive with rotation_rate == 0 (operators.jl:220-226). */
    QSIM_TERM_SCALAR = 0,
    /* H(t) += s(t) * (e^{+i th} A + e^{-i th} B), th = 2*pi*(rate*t).
     * dipole_drive(q, drive, rate): A = embed(raising), B = embed(lowering)
     * (operators.jl:29-39,59-69,118,220-226); rotating-frame exchange
     * t -> c*X_Y([qa,qb],[da*t, db*t]) = 2c(e^{i th} a'b + e^{-i th} ab'),
     * rate = da - db (benchmark/benchmarks.jl:184,193).                       */
    QSIM_TERM_ROTATING = 1,
    /* parametric_drive on a DuffingTransmon (operators.jl:265-268 ->
     * systems.jl:189-206): diagonal, E_n = (s(t) - a/2) n + (a/2) n^2 - rot*n
     * on the basis states whose subsystem level is n (`levels`).  `a` is the
     * anharmonicity, `rot` the RotatingFrameSystem rate (systems.jl:248-265). */
    QSIM_TERM_DUFFING = 2,
    /* parametric_drive on a PerturbativeTransmon: phi = s(t) is the flux in
     * Phi0; f(phi), a(phi) from the perturbative series
     * (perturbative_transmon.jl:187-261, systems.jl:79-83,213-216), then E_n
     * as above.                                                               */
    QSIM_TERM_TRANSMON = 3,
    /* parametric_drive on a DiagonalChargeBasisTransmon (systems.jl:143-152, 231-237): H = diag(E_0 .. E_{dim-1}),
     * the lowest eigenvalues of the charge-basis Hamiltonian at flux phi = s(t).  They depend on the flux only
     * through z = cos(pi*phi) (EJ(phi)^2 = (EJ1 - EJ2)^2 + 4 EJ1 EJ2 z^2) and are smooth, even functions of z,
     * so the host tabulates E_k(z) on [-1, 1] once (qsim_problem_add_table, one table per level) instead of the
     * reference's 101 x 101 eigensolve per right-hand-side call; added with qsim_problem_add_spectrum_term.   */
    QSIM_TERM_SPECTRUM = 4
};

typedef struct {
    int32_t kind;        /* QSIM_TERM_*                                        */
    int32_t num_terms;   /* TRANSMON: perturbative num_terms (default 10)      */
    qsim_signal signal;  /* s(t)                                               */
    qsim_slot rate;      /* ROTATING: frame rate (GHz)                         */
    qsim_slot anharm;    /* DUFFING: anharmonicity                             */
    qsim_slot rot;       /* DUFFING/TRANSMON: rotating-frame rate, 0 if none   */
    qsim_slot EC, EJ1, EJ2; /* TRANSMON spec (systems.jl:47-51)                */
} qsim_term;

/* Solver options.  qsim_default_opts() fills in the literals the reference
 * hard-codes (reltol/abstol, time_evolution.jl:43) and OrdinaryDiffEq's
 * defaults for Tsit5.                                      */
typedef struct {
    double reltol;      /* 1e-6 */
    double abstol;      /* 1e-8 */
    double gamma;       /* 0.9  */
    double qmin;        /* 0.2  */
    double qmax;        /* 10   */
    double beta1;       /* 7/50 */
    double beta2;       /* 2/25 */
    double qoldinit;    /* 1e-4 */
    double dtmax;       /* <= 0: ts[end] - ts[1] */
    int64_t maxiters;   /* 1000000 */
    int32_t flags;      /* QSIM_FLAG_* */
    int32_t _pad;
    void *stream;       /* cudaStream_t to launch on (NULL: the ctx stream)    */
} qsim_solver_opts;

enum {
    /* params/ts/init/out/stats are DEVICE pointers on the ctx's device; the
     * call enqueues on the stream and returns without synchronising.          */
    QSIM_FLAG_DEVICE_PTRS = 1,
    /* Evaluate the drive descriptors directly at every stage time.  By default the kernels build
     * windowed Chebyshev interpolants of the drive channels (12 nodes over a few tens of steps, checked
     * against direct evaluation to 2e-14 relative, shrunk or abandoned per trajectory if the check
     * fails) and evaluate those at the stage times.                                                  */
    QSIM_FLAG_DIRECT_DRIVES = 2
};

typedef struct {
    int32_t retcode;    /* qsim_retcode */
    int32_t n_accept;
    int32_t n_reject;
    int32_t n_rhs;      /* right-hand-side evaluations: 3 + 6*(accept+reject)  */
    double dt_init;     /* the initial step the heuristic chose                */
    double t_final;     /* time reached                                        */
} qsim_traj_stats;

typedef struct qsim_ctx qsim_ctx;
typedef struct qsim_problem qsim_problem;

int qsim_version(void);
const char *qsim_last_error(void);
void qsim_default_opts(qsim_solver_opts *opts);

/* ---- context: one CUDA device + stream ----------------------------------- */
int qsim_create(int device, qsim_ctx **ctx);
int qsim_destroy(qsim_ctx *ctx);
int qsim_device_info(qsim_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                     double *sm_clock_mhz, int64_t *global_mem_bytes);
int qsim_synchronize(qsim_ctx *ctx);
/* Page-locked host buffers for params / out: host<->device copies of pinned memory run at full
 * PCIe rate and asynchronously.  Optional: pageable caller memory works too.  */
int qsim_alloc_pinned(size_t bytes, void **ptr);
int qsim_free_pinned(void *ptr);

/* ---- problem: the final term lists of one CompositeQSystem --------------- */
/* d = dimension(cqs) (composite_systems.jl:19).  Host-only; needs no device.  */
int qsim_problem_create(int d, int n_params, qsim_problem **prob);
int qsim_problem_destroy(qsim_problem *prob);
/* h0: d*d complex column-major = hamiltonian(cqs) (composite_systems.jl:151). */
int qsim_problem_set_h0(qsim_problem *prob, const double *h0);
/* One parametric_Hs entry (composite_systems.jl:40-50).
 *   SCALAR:   atom_a = embedded operator A (d*d complex), atom_b = NULL
 *   ROTATING: atom_a = A, atom_b = B
 *   DUFFING / TRANSMON: atoms NULL, levels[d] = subsystem level of each
 *   composite basis state (0-based).                                          */
int qsim_problem_add_term(qsim_problem *prob, const qsim_term *term,
                          const double *atom_a, const double *atom_b,
                          const int32_t *levels);
/* One Lindblad operator L(t) = scale(t) * lind (d*d complex, embedded);
 * scale == NULL means 1 (fixed_Ls; add_lindblad!, composite_systems.jl:53-62). */
/* parametric_drive on a DiagonalChargeBasisTransmon (QSIM_TERM_SPECTRUM above): flux = s(t) in Phi0; rot the
 * RotatingFrameSystem rate (NULL: none); levels[d] as for DUFFING / TRANSMON; table_ids[n_levels] the tables of
 * E_0(z) .. E_{n_levels-1}(z), z = cos(pi*flux) on [-1, 1], from qsim_problem_add_table.                    */
int qsim_problem_add_spectrum_term(qsim_problem *prob, const qsim_signal *flux, const qsim_slot *rot,
                                   const int32_t *levels, int n_levels, const int32_t *table_ids);
int qsim_problem_add_lindblad(qsim_problem *prob, const double *lind,
                              const qsim_signal *scale);
/* A tabulated envelope for QSIM_ENV_TABLE signals: the generic path for drives that are opaque closures in
 * the reference (arbitrary pulse shapes, add_hamiltonian!(cqs, t -> ..., q), composite_systems.jl:40-50;
 * SURVEY.md section 8f rank 3).  The host samples the closure into a piecewise polynomial: n_seg equal
 * segments over [t_start, t_end], segment k holding n_coef (2..16) Chebyshev coefficients c_j of
 * p(t) = sum_j c_j T_j(x), x = 2 (t - t_k)/h - 1, evaluated with the Clenshaw recurrence; outside
 * [t_start, t_end] the end values continue.  coef is [n_seg][n_coef].  The accuracy of the table is the
 * host's responsibility (the Python mirror refines n_seg until the closure is matched to 1e-13).      */
int qsim_problem_add_table(qsim_problem *prob, double t_start, double t_end, int n_seg, int n_coef,
                           const double *coef, int32_t *table_id);
int qsim_problem_finalize(qsim_problem *prob);

/* Introspection used by the tests: evaluate the assembled generator at time t
 * for one parameter row.  kind 0: H(t) (d*d complex); kind 1: the Lindblad
 * superoperator G(t) (d^2*d^2 complex) with du/dt = G u.  Host-only.          */
int qsim_problem_eval(qsim_problem *prob, int kind, double t,
                      const double *param_row, double *out);

/* ---- the four batched solver entry points -------------------------------- */
/*
 * n_traj trajectories; trajectory i uses params[i*n_params .. ] and the save
 * grid ts[i*ts_stride .. +n_ts) (ts_stride == 0: one grid shared by all).
 * ts_stride must be 0 or >= n_ts.  The state solvers read the initial state of
 * trajectory i at init[i*init_stride ..] (complex elements); init_stride must be 0
 * (one state shared by all) or >= the state length (d, or d*d for me_state).
 * ts must be sorted (the reference asserts issorted, time_evolution.jl:30);
 * integration starts at ts[0] from the identity / the given initial state and
 * one result is written per ts[k], ts[0] included (save_start=true).
 *
 * out layouts (complex, column-major inner matrices, contiguous):
 *   unitary_propagator: out[traj][k][d*d]
 *   unitary_state:      out[traj][k][d]        psi0[d]   (init_stride 0: shared)
 *   me_propagator:      out[traj][k][d^2*d^2]
 *   me_state:           out[traj][k][d*d]      rho0[d*d]
 * stats may be NULL.
 *
 * Every entry of every result is written.  When the generator's graph splits into
 * connected components (e.g. a conserved excitation-number parity), the entries
 * that couple different components are exactly zero -- in the reference as well:
 * its dense products (time_evolution.jl:31-37, 98-111) only ever add 0 * x there --
 * and the library writes them as zeros without integrating them; step sequences
 * and norms are those of the full state.  The same holds for state solvers whose
 * (host-side) initial states touch only some components.
 */
int qsim_unitary_propagator(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj,
                            const double *params, const double *ts, int n_ts,
                            int64_t ts_stride, const qsim_solver_opts *opts,
                            double *out, qsim_traj_stats *stats);
int qsim_unitary_state(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj,
                       const double *params, const double *ts, int n_ts,
                       int64_t ts_stride, const double *psi0,
                       int64_t init_stride, const qsim_solver_opts *opts,
                       double *out, qsim_traj_stats *stats);
int qsim_me_propagator(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj,
                       const double *params, const double *ts, int n_ts,
                       int64_t ts_stride, const qsim_solver_opts *opts,
                       double *out, qsim_traj_stats *stats);
int qsim_me_state(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj,
                  const double *params, const double *ts, int n_ts,
                  int64_t ts_stride, const double *rho0, int64_t init_stride,
                  const qsim_solver_opts *opts, double *out,
                  qsim_traj_stats *stats);

/* Reduced output: the same solves, but only selected entries of each saved
 * result leave the device -- what every sweep in the reference's notebooks keeps of a propagator
 * (populations abs2(u[i,j]), the qubit-subspace block; doc/Examples.ipynb cells 14-15, 31, 34-38,
 * test/test_time_evolution.jl:56-57,72).  which: 0 unitary_propagator, 1 unitary_state, 2 me_propagator,
 * 3 me_state.  Entry i is (sel_rows[i], sel_cols[i]) of the saved matrix (0-based; which = 1 saves a
 * vector and ignores sel_cols; which = 3 addresses the d*d density matrix, or, with sel_cols == NULL,
 * its column-major linear index).  sel_rows / sel_cols are HOST arrays even with
 * QSIM_FLAG_DEVICE_PTRS.  out[traj][k][n_sel]: complex (interleaved) when abs2 == 0, real |entry|^2
 * otherwise.  init / init_stride as in the state solvers (ignored for propagators).                  */
int qsim_solve_selected(qsim_ctx *ctx, qsim_problem *prob, int which, int64_t n_traj,
                        const double *params, const double *ts, int n_ts, int64_t ts_stride,
                        const double *init, int64_t init_stride, const qsim_solver_opts *opts,
                        int n_sel, const int32_t *sel_rows, const int32_t *sel_cols, int abs2,
                        double *out, qsim_traj_stats *stats);

/* One sweep over several devices from a single call: trajectories [0, n_traj) are split into n_ctx
 * contiguous blocks, block i is integrated through ctxs[i] (one ctx per device, each on its own host
 * thread inside the library) and written straight into the caller's out / stats at its offsets; there is
 * no collective and no ordering between blocks.  Host buffers only (QSIM_FLAG_DEVICE_PTRS and opts->stream
 * are rejected).  which / init / init_stride as in qsim_solve_selected.  The reference's sweep is a
 * user-level loop over unitary_propagator calls (doc/Examples.ipynb); this is that loop, sharded.       */
int qsim_solve_multi(qsim_ctx **ctxs, int n_ctx, qsim_problem *prob, int which, int64_t n_traj,
                     const double *params, const double *ts, int n_ts, int64_t ts_stride,
                     const double *init, int64_t init_stride, const qsim_solver_opts *opts,
                     double *out, qsim_traj_stats *stats);

/* Executed FP64 flop of the generator application per right-hand-side call of
 * one trajectory (8 flop per complex multiply-add actually issued; 4 when the
 * generator is purely imaginary, i.e. a real Hamiltonian), for the roofline
 * arithmetic.  which: 0 unitary_propagator, 1 unitary_state,
 * 2 me_propagator, 3 me_state.                                                */
int qsim_problem_flops_per_rhs(qsim_problem *prob, int which, double *flops);

/* Name of the kernel variant the dispatcher picks for this problem.          */
int qsim_problem_kernel_name(qsim_problem *prob, int which, char *buf, int buflen);

/* Dimension and number of sweep parameters the problem was created with.       */
int qsim_problem_dims(qsim_problem *prob, int *d, int *n_params);

/* Floquet propagation of a whole sweep in one call: replaces, batched and device-resident, the reference's
 * floquet_propagator / floquet_rise_fall_propagator (src/time_evolution.jl:188-220, 312-338) wrapped around
 * unitary_propagator (which = 0) or me_propagator (which = 2).  The system of trajectory i is taken to be periodic
 * with period t_period[i * period_stride] (period_stride 0: one period for all) between ts[0] + rise_time and
 * ts[n_ts-1] - fall_time; with rise_time = fall_time = 0 this is floquet_propagator.  rtol <= 0 selects the
 * reference's default eps(t_period) for merging the times modulo the period (unique_tol, :355-360).
 * Every segment (rise, the unique remainders plus one period, fall) of every trajectory is one trajectory of a
 * single batched solver launch; U(n tau + dt) = U(dt) U(tau)^n and the segment products are formed on the device
 * and only out[traj][k][n*n] (n = d or d*d, column-major) is delivered.  Host buffers only.  stats (may be NULL)
 * sums the step counts of a trajectory's segments.                                                        */
int qsim_floquet_propagator(qsim_ctx *ctx, qsim_problem *prob, int which, int64_t n_traj,
                            const double *params, const double *ts, int n_ts, int64_t ts_stride,
                            const double *t_period, int64_t period_stride, double rise_time,
                            double fall_time, double rtol, const qsim_solver_opts *opts,
                            double *out, qsim_traj_stats *stats);

/* Number of kernel launches issued through this ctx since creation.          */
int64_t qsim_launch_count(qsim_ctx *ctx);

/* Device time (CUDA events on the launch stream, milliseconds) from the start of the first to the end of
 * the last kernel of the most recent solver call on this ctx: the integration alone, without the
 * host<->device copies around it.  Waits for that call's kernels if they are still running.            */
int qsim_last_kernel_ms(qsim_ctx *ctx, double *ms);

/* Facts about the generator and the launch plan of one entry point (which: 0..3 as above), for reports and
 * tests.  info[0] state rows n, [1] columns integrated together, [2] stored generator entries (nnz),
 * [3] longest row, [4] value slots, [5] time-dependent value slots, [6] channels, [7] warps per trajectory,
 * [8] trajectories (teams) per CTA, [9] threads per CTA, [10] dynamic shared memory per CTA in bytes,
 * [11] 1 when the columns stream through scratch memory, [12] scratch bytes per trajectory in flight,
 * [13] value mode of the kernel, [14] 1 for the row-split kernels, [15] connected components of the generator's
 * graph when the propagator is integrated block by block (component pruning: entries that couple different
 * components are exactly zero in the reference too and are written as zeros), 0 when the whole state is integrated. */
int qsim_problem_plan_info(qsim_problem *prob, int which, int64_t info[16]);

/* FP64 FMA micro-benchmark: sustained DFMA TFLOP/s of the ctx's device,
 * measured with CUDA events over `ms_target` milliseconds of work.           */
int qsim_peak_fp64(qsim_ctx *ctx, double ms_target, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* QSIM_B200_H */
