// Batched adaptive Tsit5 integrator for du/dt = A(t) u on sm_100a: the hot path.
//
// Replaces, for a whole parameter sweep at once, the reference's per-point
//   solve(ODEProblem(ode, u0, tspan, p), Tsit5(); saveat=ts, save_start=true, reltol=1e-6, abstol=1e-8)
// (/root/reference/src/time_evolution.jl:42-43, 74-75, 118-119, 159-160) including the per-RHS
// assembly H(t) = H0 + sum f_k(t) O_k (composite_systems.jl:46-50) and the RHS products
// (time_evolution.jl:31-37, 98-111, 142-154).  Step control, error norm, initial step and dense
// output follow OrdinaryDiffEq's Tsit5 so that each trajectory takes the same
// step sequence as the reference solver.
//
// Mapping to the machine
//   * one trajectory = one parameter point of the sweep; a TEAM of `team_warps` warps integrates it.
//   * the state is a set of independent columns (columns of U, or of the superoperator propagator)
//     that share one adaptive step sequence: only the error norm couples them.  A warp owns a
//     BATCH of columns; during a step the batch (u, k1..k7) lives entirely in registers, the stage
//     vector is exchanged through shared memory, and A(t) is applied as a sparse row gather whose
//     (source, value-slot) lists sit in registers for the whole kernel.
//   * resident mode: one batch per warp, state never leaves registers (unitary d <= 32, me_state).
//     streamed mode: more batches than warps (me_propagator): u/k1 are read and u_new/k7 written
//     once per step per column through L2-resident scratch; everything else stays on chip.
//   * A(t): per step, the team evaluates the drive channels (FP64 sin/cos/sqrt/Horner) for the five
//     distinct stage times in parallel across lanes and refreshes only the time-dependent value
//     slots in shared memory.
//   * work distribution: persistent CTAs, dynamic queue over trajectories (atomic counter).
// No reductions use atomics, so results are bitwise reproducible and independent of the save grid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "qsim_eval.h"

namespace qsim {

struct DevGenerator {
    int n, max_nnz, n_slots, n_dyn_slots, n_chan, n_evals;
    const uint32_t *ell;      // [max_nnz][n]
    const uint8_t *row_len;   // [n]
    const int32_t *slot_ptr;  // [n_slots+1]
    const int32_t *pair_chan;
    const double2 *pair_w;
    const EvalDesc *evals;
    const SeriesTables *tab;
};

struct LaunchArgs {
    DevGenerator g;
    int ncols;            // columns of the state (1 for *_state)
    int identity_init;    // 1: u0 = I (propagators); 0: u0 = init (column 0)
    int n_params, n_ts;
    long long ts_stride, init_stride, n_traj;
    const double *params, *ts;
    const double2 *init;
    double2 *out;
    qsim_traj_stats *stats;
    unsigned long long *work_counter;
    double2 *scratch;     // streamed mode: [n_teams_total][4][n_batches*cpb*n]
    int team_warps, teams_per_cta, n_batches, groups, streamed;
    int nvp, ncp;         // padded slot / channel counts
    int team_smem_doubles;  // shared-memory doubles per team
    double reltol, abstol, gamma, qmin, qmax, beta1, beta2, qoldinit, dtmax;
    long long maxiters;
};

// stage-time fractions c2..c5 and c6 = c7 = 1 (stages 6 and 7 share one table)
static __device__ __constant__ double kStageC[5] = {0.161, 0.327, 0.9, 0.9800255409045097, 1.0};

// Tsit5 tableau (OrdinaryDiffEq Tsit5ConstantCache, Float64)
#define A21 0.161
#define A31 -0.008480655492356989
#define A32 0.335480655492357
#define A41 2.8971530571054935
#define A42 -6.359448489975075
#define A43 4.3622954328695815
#define A51 5.325864828439257
#define A52 -11.748883564062828
#define A53 7.4955393428898365
#define A54 -0.09249506636175525
#define A61 5.86145544294642
#define A62 -12.92096931784711
#define A63 8.159367898576159
#define A64 -0.071584973281401
#define A65 -0.028269050394068383
#define A71 0.09646076681806523
#define A72 0.01
#define A73 0.4798896504144996
#define A74 1.379008574103742
#define A75 -3.290069515436081
#define A76 2.324710524099774
#define BT1 -0.00178001105222577714
#define BT2 -0.0008164344596567469
#define BT3 0.007880878010261995
#define BT4 -0.1447110071732629
#define BT5 0.5823571654525552
#define BT6 -0.45808210592918697
#define BT7 0.015151515151515152

__device__ __forceinline__ void team_sync(int team_warps, int bar_id) {
    if (team_warps == 1)
        __syncwarp();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(team_warps * 32) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic team-wide sum of up to NV values held per thread.
template <int NV>
__device__ __forceinline__ void team_sum(double (&v)[NV], double *red, int team_warps, int warp_in_team, int lane,
                                         int bar_id) {
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
    if (team_warps == 1) return;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) red[warp_in_team * NV + i] = v[i];
    }
    team_sync(team_warps, bar_id);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
        for (int w = 0; w < team_warps; w++) s += red[w * NV + i];
        v[i] = s;
    }
    team_sync(team_warps, bar_id);
}

template <int RPL, int CPL, int MAXNNZ>
struct Solver {
    static constexpr int E = RPL * CPL;

    // ---- per-thread constants -------------------------------------------------
    const LaunchArgs &a;
    int lane, warp_in_team, team_thread, team_threads, bar_id;
    int g_idx, row0, cpb, colofs;
    bool lane_active;
    bool row_ok[RPL];
    int row_of[RPL];
    uint32_t ell[RPL][MAXNNZ];
    // shared memory
    double2 *V;      // [5][nvp]
    double *chan;    // [5][ncp]
    double *red;     // [team_warps*4 + 8]
    double2 *ybuf;   // per warp: [2][n*cpb]
    // per-trajectory
    const double *prow;
    double2 *sc_u[2], *sc_k[2];  // streamed scratch (cur/next)

    __device__ Solver(const LaunchArgs &args) : a(args) {}

    __device__ __forceinline__ void setup(double *team_smem, int team_global_idx) {
        const int n = a.g.n;
        lane = threadIdx.x & 31;
        team_threads = a.team_warps * 32;
        team_thread = threadIdx.x % team_threads;
        warp_in_team = team_thread >> 5;
        bar_id = 1 + threadIdx.x / team_threads;
        cpb = a.groups * CPL;
        if (a.groups > 1) {
            g_idx = lane / n;
            row0 = lane % n;
            lane_active = g_idx < a.groups;
        } else {
            g_idx = 0;
            row0 = lane;
            lane_active = true;
        }
        colofs = g_idx * CPL;
#pragma unroll
        for (int i = 0; i < RPL; i++) {
            row_of[i] = row0 + 32 * i;
            row_ok[i] = lane_active && row_of[i] < n;
#pragma unroll
            for (int j = 0; j < MAXNNZ; j++)
                ell[i][j] = (row_ok[i] && j < a.g.max_nnz) ? a.g.ell[(size_t)j * n + row_of[i]] : 0xFFFFFFFFu;
        }
        V = reinterpret_cast<double2 *>(team_smem);
        chan = team_smem + 2 * 5 * a.nvp;
        red = chan + 5 * a.ncp;
        ybuf = reinterpret_cast<double2 *>(red + a.team_warps * 4 + 8) + (size_t)warp_in_team * 2 * n * cpb;
        if (a.streamed) {
            const size_t buf = (size_t)a.n_batches * cpb * n;
            double2 *base = a.scratch + (size_t)team_global_idx * 4 * buf;
            sc_u[0] = base;
            sc_k[0] = base + buf;
            sc_u[1] = base + 2 * buf;
            sc_k[1] = base + 3 * buf;
        }
    }

    // ---- A(t) value tables -----------------------------------------------------
    // Refresh tables s < ntab for the times tbase + kStageC[s]*dtv; all==false touches only
    // time-dependent channels and slots.  Two team barriers.
    __device__ __forceinline__ void compute_tables(double tbase, double dtv, int ntab, bool all) {
        const DevGenerator &g = a.g;
        const int ne = g.n_evals;
        for (int it = team_thread; it < ntab * ne; it += team_threads) {
            const int s = it / ne, e = it - s * ne;
            const EvalDesc &ev = g.evals[e];
            if (all || ev.dynamic) run_evaluator(ev, *g.tab, tbase + kStageC[s] * dtv, prow, chan + s * a.ncp);
        }
        team_sync(a.team_warps, bar_id);
        const int nsl = all ? g.n_slots : g.n_dyn_slots;
        for (int it = team_thread; it < ntab * nsl; it += team_threads) {
            const int s = it / nsl, sl = it - s * nsl;
            const double *ch = chan + s * a.ncp;
            double vx = 0.0, vy = 0.0;
            for (int q = g.slot_ptr[sl]; q < g.slot_ptr[sl + 1]; q++) {
                const double c = ch[g.pair_chan[q]];
                const double2 w = g.pair_w[q];
                vx = fma(w.x, c, vx);
                vy = fma(w.y, c, vy);
            }
            V[s * a.nvp + sl] = make_double2(vx, vy);
        }
        team_sync(a.team_warps, bar_id);
    }

    // ---- stage exchange + sparse apply ------------------------------------------
    __device__ __forceinline__ void put_stage(const double2 (&y)[E], double2 *yb) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (row_ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) yb[row_of[i] * cpb + colofs + c] = y[i * CPL + c];
            }
        __syncwarp();
    }

    __device__ __forceinline__ void apply(const double2 *Vs, const double2 *yb, double2 (&k)[E]) {
#pragma unroll
        for (int i = 0; i < RPL; i++) {
            double2 acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; c++) acc[c] = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < MAXNNZ; j++) {
                const uint32_t e = ell[i][j];
                if (e != 0xFFFFFFFFu) {
                    const double2 v = Vs[e >> 16];
                    const double2 *yp = yb + (e & 0xFFFFu) * cpb + colofs;
#pragma unroll
                    for (int c = 0; c < CPL; c++) {
                        const double2 y = yp[c];
                        acc[c].x = fma(v.x, y.x, acc[c].x);
                        acc[c].x = fma(-v.y, y.y, acc[c].x);
                        acc[c].y = fma(v.x, y.y, acc[c].y);
                        acc[c].y = fma(v.y, y.x, acc[c].y);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < CPL; c++) k[i * CPL + c] = acc[c];
        }
    }

    // column index of local column c in batch b; validity
    __device__ __forceinline__ int col_of(int b, int c) const { return b * cpb + colofs + c; }

    __device__ __forceinline__ void initial_state(int b, long long traj, double2 (&u)[E]) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
#pragma unroll
            for (int c = 0; c < CPL; c++) {
                double2 v = make_double2(0.0, 0.0);
                const int col = col_of(b, c);
                if (row_ok[i] && col < a.ncols) {
                    if (a.identity_init)
                        v.x = (row_of[i] == col) ? 1.0 : 0.0;
                    else
                        v = a.init[(size_t)traj * a.init_stride + row_of[i]];
                }
                u[i * CPL + c] = v;
            }
    }

    __device__ __forceinline__ void load_batch(const double2 *buf, int b, double2 (&x)[E]) {
        const int n = a.g.n;
#pragma unroll
        for (int i = 0; i < RPL; i++)
#pragma unroll
            for (int c = 0; c < CPL; c++)
                x[i * CPL + c] = row_ok[i] ? buf[(size_t)col_of(b, c) * n + row_of[i]] : make_double2(0.0, 0.0);
    }
    __device__ __forceinline__ void store_batch(double2 *buf, int b, const double2 (&x)[E]) {
        const int n = a.g.n;
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (row_ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) buf[(size_t)col_of(b, c) * n + row_of[i]] = x[i * CPL + c];
            }
    }
    __device__ __forceinline__ void write_out(long long traj, int ksave, int b, const double2 (&x)[E]) {
        const int n = a.g.n;
        double2 *dst = a.out + ((size_t)traj * a.n_ts + ksave) * (size_t)a.ncols * n;
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (row_ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    const int col = col_of(b, c);
                    if (col < a.ncols) dst[(size_t)col * n + row_of[i]] = x[i * CPL + c];
                }
            }
    }

    // Dense output for the save points ks in [k_lo, k_hi) of the step [t, t+dt] (tnew = snapped end).
    __device__ __forceinline__ void save_points(long long traj, const double *ts, int k_lo, int k_hi, double t,
                                                double dt, double tnew, int b, const double2 (&u)[E],
                                                const double2 (&k1)[E], const double2 (&k2)[E],
                                                const double2 (&k3)[E], const double2 (&k4)[E],
                                                const double2 (&k5)[E], const double2 (&k6)[E],
                                                const double2 (&k7)[E], const double2 (&unew)[E]) {
        for (int ks = k_lo; ks < k_hi; ks++) {
            const double tsv = ts[ks];
            if (tsv == tnew) {
                write_out(traj, ks, b, unew);
            } else {
                const double th = (tsv - t) / dt;
                const double th2 = th * th;
                const double b1 = th * (1.0 + th * (-2.763706197274826 + th * (2.9132554618219126 + th * -1.0530884977290216)));
                const double b2 = th2 * (0.13169999999999998 + th * (-0.2234 + th * 0.1017));
                const double b3 = th2 * (3.9302962368947516 + th * (-5.941033872131505 + th * 2.490627285651253));
                const double b4 = th2 * (-12.411077166933676 + th * (30.33818863028232 + th * -16.548102889244902));
                const double b5 = th2 * (37.50931341651104 + th * (-88.1789048947664 + th * 47.37952196281928));
                const double b6 = th2 * (-27.896526289197286 + th * (65.09189467479366 + th * -34.87065786149661));
                const double b7 = th2 * (1.5 + th * (-4.0 + th * 2.5));
                double2 o[E];
#pragma unroll
                for (int e = 0; e < E; e++) {
                    double sx = k1[e].x * b1, sy = k1[e].y * b1;
                    sx += k2[e].x * b2; sy += k2[e].y * b2;
                    sx += k3[e].x * b3; sy += k3[e].y * b3;
                    sx += k4[e].x * b4; sy += k4[e].y * b4;
                    sx += k5[e].x * b5; sy += k5[e].y * b5;
                    sx += k6[e].x * b6; sy += k6[e].y * b6;
                    sx += k7[e].x * b7; sy += k7[e].y * b7;
                    o[e] = make_double2(u[e].x + dt * sx, u[e].y + dt * sy);
                }
                write_out(traj, ks, b, o);
            }
        }
    }

    // ---- one trajectory -----------------------------------------------------------
    __device__ void run(long long traj) {
        const int n = a.g.n;
        const double *ts = a.ts + (size_t)traj * a.ts_stride;
        prow = a.params + (size_t)traj * a.n_params;
        const double t0 = ts[0], tend = ts[a.n_ts - 1];
        const double ntot = (double)n * (double)a.ncols;
        double2 *yb0 = ybuf, *yb1 = ybuf + (size_t)n * cpb;
        const double2 *V0 = V, *V1 = V + a.nvp, *V2 = V + 2 * a.nvp, *V3 = V + 3 * a.nvp, *V4 = V + 4 * a.nvp;

        qsim_traj_stats st;
        st.retcode = QSIM_RET_SUCCESS;
        st.n_accept = st.n_reject = st.n_rhs = 0;
        st.dt_init = 0.0;
        st.t_final = t0;

        // constant channel + all evaluators at t0 (sets static channels and static slots in all 5 tables)
        for (int s = team_thread; s < 5; s += team_threads) chan[s * a.ncp] = 1.0;
        team_sync(a.team_warps, bar_id);
        compute_tables(t0, 0.0, 5, true);

        double2 u[E], k1[E], k2[E], k3[E], k4[E], k5[E], k6[E], k7[E], tmp[E];
        int cur = 0;
        int nsave = 0;
        while (nsave < a.n_ts && ts[nsave] <= t0) nsave++;  // save_start: every ts[k] == t0

        double t = t0, dt = 0.0;
        if (tend > t0) {
            const double dtmax = a.dtmax > 0.0 ? a.dtmax : tend - t0;
            // ---- initial step (DiffEqBase ode_determine_initdt) ----
            double acc2[2] = {0.0, 0.0};
            for (int b = warp_in_team; b < a.n_batches; b += a.team_warps) {
                initial_state(b, traj, u);
                for (int ks = 0; ks < nsave; ks++) write_out(traj, ks, b, u);
                put_stage(u, yb0);
                apply(V0, yb0, k1);
                if (a.streamed) {
                    store_batch(sc_u[0], b, u);
                    store_batch(sc_k[0], b, k1);
                }
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double sk = a.abstol + sqrt(u[e].x * u[e].x + u[e].y * u[e].y) * a.reltol;
                    const double ux = u[e].x / sk, uy = u[e].y / sk, fx = k1[e].x / sk, fy = k1[e].y / sk;
                    acc2[0] += ux * ux + uy * uy;
                    acc2[1] += fx * fx + fy * fy;
                }
            }
            team_sum<2>(acc2, red, a.team_warps, warp_in_team, lane, bar_id);
            const double d0 = sqrt(acc2[0] / ntot), d1 = sqrt(acc2[1] / ntot);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
            dt0 = fmin(dt0, dtmax);
            // table 1 <- t0 + dt0 (table 0 keeps t0): evaluate through offset pointers
            {
                double *chan_save = chan;
                double2 *V_save = V;
                chan = chan + a.ncp;
                V = V + a.nvp;
                compute_tables(t0 + dt0, 0.0, 1, false);
                chan = chan_save;
                V = V_save;
            }
            double acc1[1] = {0.0};
            for (int b = warp_in_team; b < a.n_batches; b += a.team_warps) {
                if (a.streamed) {
                    load_batch(sc_u[0], b, u);
                    load_batch(sc_k[0], b, k1);
                }
#pragma unroll
                for (int e = 0; e < E; e++) tmp[e] = make_double2(u[e].x + dt0 * k1[e].x, u[e].y + dt0 * k1[e].y);
                put_stage(tmp, yb1);
                apply(V1, yb1, k2);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double sk = a.abstol + sqrt(u[e].x * u[e].x + u[e].y * u[e].y) * a.reltol;
                    const double fx = (k2[e].x - k1[e].x) / sk, fy = (k2[e].y - k1[e].y) / sk;
                    acc1[0] += fx * fx + fy * fy;
                }
            }
            team_sum<1>(acc1, red, a.team_warps, warp_in_team, lane, bar_id);
            const double d2 = sqrt(acc1[0] / ntot) / dt0;
            const double md = fmax(d1, d2);
            const double dt1 = md <= 1e-15 ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(md)) / 5.0);
            dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
            st.dt_init = dt;
            st.n_rhs = 3;  // the reference evaluates f(u0,t0) twice (init-dt, then the integrator's first stage)

            // ---- main loop ----
            double qold = a.qoldinit;
            long long iters = 0;
            while (t < tend) {
                if (++iters > a.maxiters) { st.retcode = QSIM_RET_MAXITERS; break; }
                dt = fmin(dt, dtmax);
                dt = fmin(dt, tend - t);
                if (!(dt > fabs(t) * 2.220446049250313e-16)) { st.retcode = QSIM_RET_DT_UNDERFLOW; break; }
                compute_tables(t, dt, 5, false);
                double tnew = t + dt;
                {
                    const double mx = fmax(t, tend);
                    // 10*eps(max(t, tstop)): snap the last step onto the end point
                    const double ulp = __longlong_as_double(__double_as_longlong(mx) + 1) - mx;
                    if (fabs(tnew - tend) < 10.0 * ulp) tnew = tend;
                }
                int nsave_hi = nsave;
                while (nsave_hi < a.n_ts && ts[nsave_hi] <= tnew) nsave_hi++;

                double err[1] = {0.0};
                for (int b = warp_in_team; b < a.n_batches; b += a.team_warps) {
                    if (a.streamed) {
                        load_batch(sc_u[cur], b, u);
                        load_batch(sc_k[cur], b, k1);
                    }
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = A21 * k1[e].x, ay = A21 * k1[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(tmp, yb0);
                    apply(V0, yb0, k2);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ax = A31 * k1[e].x, ay = A31 * k1[e].y;
                        ax += A32 * k2[e].x; ay += A32 * k2[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(tmp, yb1);
                    apply(V1, yb1, k3);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ax = A41 * k1[e].x, ay = A41 * k1[e].y;
                        ax += A42 * k2[e].x; ay += A42 * k2[e].y;
                        ax += A43 * k3[e].x; ay += A43 * k3[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(tmp, yb0);
                    apply(V2, yb0, k4);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ax = A51 * k1[e].x, ay = A51 * k1[e].y;
                        ax += A52 * k2[e].x; ay += A52 * k2[e].y;
                        ax += A53 * k3[e].x; ay += A53 * k3[e].y;
                        ax += A54 * k4[e].x; ay += A54 * k4[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(tmp, yb1);
                    apply(V3, yb1, k5);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ax = A61 * k1[e].x, ay = A61 * k1[e].y;
                        ax += A62 * k2[e].x; ay += A62 * k2[e].y;
                        ax += A63 * k3[e].x; ay += A63 * k3[e].y;
                        ax += A64 * k4[e].x; ay += A64 * k4[e].y;
                        ax += A65 * k5[e].x; ay += A65 * k5[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(tmp, yb0);
                    apply(V4, yb0, k6);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ax = A71 * k1[e].x, ay = A71 * k1[e].y;
                        ax += A72 * k2[e].x; ay += A72 * k2[e].y;
                        ax += A73 * k3[e].x; ay += A73 * k3[e].y;
                        ax += A74 * k4[e].x; ay += A74 * k4[e].y;
                        ax += A75 * k5[e].x; ay += A75 * k5[e].y;
                        ax += A76 * k6[e].x; ay += A76 * k6[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);  // u_new
                    }
                    put_stage(tmp, yb1);
                    apply(V4, yb1, k7);
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        double ex = BT1 * k1[e].x, ey = BT1 * k1[e].y;
                        ex += BT2 * k2[e].x; ey += BT2 * k2[e].y;
                        ex += BT3 * k3[e].x; ey += BT3 * k3[e].y;
                        ex += BT4 * k4[e].x; ey += BT4 * k4[e].y;
                        ex += BT5 * k5[e].x; ey += BT5 * k5[e].y;
                        ex += BT6 * k6[e].x; ey += BT6 * k6[e].y;
                        ex += BT7 * k7[e].x; ey += BT7 * k7[e].y;
                        ex *= dt; ey *= dt;
                        const double m2 = fmax(u[e].x * u[e].x + u[e].y * u[e].y, tmp[e].x * tmp[e].x + tmp[e].y * tmp[e].y);
                        const double sk = a.abstol + sqrt(m2) * a.reltol;
                        ex /= sk; ey /= sk;
                        err[0] += ex * ex + ey * ey;
                    }
                    if (a.streamed) {
                        // tentative: next buffers are only adopted if the step is accepted; dense output
                        // slots are rewritten by the retried / later step if it is not
                        store_batch(sc_u[cur ^ 1], b, tmp);
                        store_batch(sc_k[cur ^ 1], b, k7);
                        save_points(traj, ts, nsave, nsave_hi, t, dt, tnew, b, u, k1, k2, k3, k4, k5, k6, k7, tmp);
                    }
                }
                st.n_rhs += 6;
                team_sum<1>(err, red, a.team_warps, warp_in_team, lane, bar_id);
                const double eest = sqrt(err[0] / ntot);
                if (!isfinite(eest)) { st.retcode = QSIM_RET_NONFINITE; break; }
                // PI controller (stepsize_controller! / step_accept_controller! / step_reject_controller!)
                double q, q11 = 0.0;
                if (eest == 0.0) {
                    q = 1.0 / a.qmax;
                } else {
                    q11 = pow(eest, a.beta1);
                    q = q11 / pow(qold, a.beta2);
                    q = fmax(1.0 / a.qmax, fmin(1.0 / a.qmin, q / a.gamma));
                }
                if (eest <= 1.0) {
                    st.n_accept++;
                    qold = fmax(eest, a.qoldinit);
                    if (!a.streamed) {
                        if (warp_in_team < a.n_batches)
                            save_points(traj, ts, nsave, nsave_hi, t, dt, tnew, warp_in_team, u, k1, k2, k3, k4, k5, k6,
                                        k7, tmp);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            u[e] = tmp[e];
                            k1[e] = k7[e];
                        }
                    } else {
                        cur ^= 1;
                    }
                    nsave = nsave_hi;
                    t = tnew;
                    dt = dt / q;
                } else {
                    st.n_reject++;
                    dt = dt / fmin(1.0 / a.qmin, q11 / a.gamma);
                }
            }
        } else {
            // zero-length span: only the start saves
            for (int b = warp_in_team; b < a.n_batches; b += a.team_warps) {
                initial_state(b, traj, u);
                for (int ks = 0; ks < nsave; ks++) write_out(traj, ks, b, u);
            }
        }
        // unreached save points of a failed trajectory: NaN
        if (nsave < a.n_ts) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int e = 0; e < E; e++) tmp[e] = make_double2(qnan, qnan);
            for (int b = warp_in_team; b < a.n_batches; b += a.team_warps)
                for (int ks = nsave; ks < a.n_ts; ks++) write_out(traj, ks, b, tmp);
        }
        st.t_final = t;
        if (team_thread == 0 && a.stats) a.stats[traj] = st;
    }
};

template <int RPL, int CPL, int MAXNNZ>
__global__ void __launch_bounds__(320) solve_kernel(const __grid_constant__ LaunchArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int team_threads = a.team_warps * 32;
    const int team_in_cta = threadIdx.x / team_threads;
    double *team_smem = smem + (size_t)team_in_cta * a.team_smem_doubles;
    Solver<RPL, CPL, MAXNNZ> s(a);
    s.setup(team_smem, blockIdx.x * a.teams_per_cta + team_in_cta);
    // work index broadcast slot: last 2 doubles of the team's `red` area
    unsigned long long *wslot = reinterpret_cast<unsigned long long *>(s.red + a.team_warps * 4 + 6);
    for (;;) {
        if (s.team_thread == 0) *wslot = atomicAdd(a.work_counter, 1ULL);
        team_sync(a.team_warps, s.bar_id);
        const unsigned long long traj = *wslot;
        team_sync(a.team_warps, s.bar_id);
        if (traj >= (unsigned long long)a.n_traj) break;
        s.run((long long)traj);
    }
}

// Host-side launcher for one (RPL, CPL, MAXNNZ) instantiation.
typedef cudaError_t (*launch_fn)(const LaunchArgs &, int grid, int block, size_t smem, cudaStream_t);

template <int RPL, int CPL, int MAXNNZ>
cudaError_t launch_solver(const LaunchArgs &a, int grid, int block, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(solve_kernel<RPL, CPL, MAXNNZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    solve_kernel<RPL, CPL, MAXNNZ><<<grid, block, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int RPL, int CPL, int MAXNNZ>
cudaError_t occupancy_solver(int block, size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(solve_kernel<RPL, CPL, MAXNNZ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, solve_kernel<RPL, CPL, MAXNNZ>, block, smem);
}

}  // namespace qsim
