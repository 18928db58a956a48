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
//     vector is exchanged through shared memory (explicit ld/st.shared on 32-bit addresses), and
//     A(t) is applied as a sparse row gather whose (source offset, value-slot offset) lists sit in
//     registers for the whole kernel.
//   * resident mode: one batch per warp, state never leaves registers (unitary d <= 32, me_state).
//     streamed mode: more batches than warps (me_propagator): u/k1 are read and u_new/k7 written
//     once per step per column through L2-resident scratch; everything else stays on chip.
//     row-split mode: state vectors longer than 96 entries (or rows longer than 16 entries): the rows
//     of a column are spread over the team's warps, rows come from global memory sorted by length.
//   * A(t): per trajectory the drive descriptors are resolved against the parameter row into
//     shared memory.  Each time-dependent channel is kept as a windowed Chebyshev interpolant
//     (12 nodes, built from the direct FP64 sin/cos/sqrt/Horner evaluator and checked against it);
//     per step the five distinct stage times are evaluated from the interpolants and only the
//     time-dependent value slots are refreshed (QSIM_FLAG_DIRECT_DRIVES: direct evaluation instead).
//   * step controller: EEst^b1 and qold^b2 from mantissa-interval Taylor tables passed in the
//     kernel parameters (constant bank); libm only outside the tables' range.
//   * lean kernels (Kern<..., LEAN>, resident teams): the stage vectors die as early as the tableau allows (five
//     live vectors in stage 7 instead of ten), which fits the step into 168 registers -- three warps per scheduler,
//     12 one-warp teams per SM -- and, for the seven-warp teams of 25..32-level propagators, makes room for the
//     static off-diagonal values next to four entries per lane.  The k's a dense-output point needs are parked in
//     lane-private shared memory on the steps that contain one.
//   * work distribution: persistent CTAs, dynamic queue over trajectories (atomic counter).
// No reductions use atomics, so results are bitwise reproducible and independent of the save grid.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "qsim_eval.h"

namespace qsim {


struct DevGenerator {
    int n, max_nnz, n_slots, n_dyn_slots, n_chan, n_evals;
    int n_vrows;              // rows of `ell` (>= n: rows n.. are the helper lanes' halves of split rows; see split_info)
    int n_dyn_pairs;          // pairs of the dynamic slots (a prefix of the pair arrays)
    const uint32_t *ell;      // [max_nnz][n]: src | slot << 16, 0xFFFFFFFF = padding
    const uint8_t *row_len;   // [n]
    const int32_t *slot_ptr;  // [n_slots+1]
    const int32_t *pair_chan;
    const double2 *pair_w;
    const EvalDesc *evals;
    const SeriesTables *tab;
    const double *tables;     // tabulated envelopes (flat; see Generator::tables) or nullptr
};

struct LaunchArgs {
    DevGenerator g;
    int ncols;            // columns of the state (1 for *_state)
    int n_out;            // rows of the state as the caller sees it (= g.n except with additively split rows, where g.n counts table rows)
    int identity_init;    // 1: u0 = I (propagators); 0: u0 = init (column 0)
    int n_params, n_ts;
    long long ts_stride, init_stride, n_traj;
    // Trajectory j of this launch reads inputs (params, ts, init) of sweep point in_first + j * in_step and writes
    // results and stats at position out_first + j * out_step: a block-cyclic shard of a sweep addresses the
    // caller's arrays in place; a chunk of a pipelined call writes a chunk-local result buffer.
    long long in_first, in_step, out_first, out_step;
    const double *params, *ts;
    const double2 *init;
    double2 *out;
    // selected-entry output (qsim_solve_selected): sel_index[col*n + row] = position of the entry in the
    // selection or -1; out is then [traj][k][n_sel] complex, or real |.|^2 when sel_abs2
    const int *sel_index;
    int n_sel, sel_abs2;
    qsim_traj_stats *stats;
    unsigned long long *work_counter;
    double2 *scratch;     // streamed mode: [n_teams_total][4][n_batches*cpb*n]
    const uint16_t *blk_slot;    // blocked variants: [n_blk_rows][(max_nnz-1)*RPL*RPL] slot ids of the block values
    int n_blk_rows;              // blocked and tiered variants: block rows (= lanes per column)
    const int *row_perm;         // tiered variants: state row of table row v (lane's block row * RPL + row-slot)
    const int *row_pos;          // tiered variants: position of table row v in the stage buffer (host-chosen against bank conflicts)
    // Row-split mode.  Rows are sorted by descending length and dealt out in chunks of 32: chunk c = slot * team_warps
    // + warp holds sorted rows [32c, 32c + 32), so every (warp, slot) group is homogeneous in length and the
    // warps of a team carry equal work.  ell_packed[j][q]: entry j of sorted row q, translated (source offset |
    // slot offset << 16), diagonal first, no interior padding; ell_rows = n + 8 (eight zero-slot padding rows
    // for lanes past the last row); rs_perm[q] = state row of sorted row q; rs_len[c] = entries per row of chunk c.
    const uint32_t *ell_packed;
    int ell_rows;
    const int *rs_perm;
    const int *rs_len;
    // Row-split tables, compact form: chunk c owns table rows [rs_off[c], rs_off[c] + rs_len[c]) of 32 entries each
    // (entry j of the chunk's lane l at ell_packed[(rs_off[c] + j) * 32 + l]); every chunk is padded with zero-slot
    // entries to 1 + an even number of off-diagonal entries plus two more rows, so the pairwise-pipelined gather may
    // read one pair ahead.  ell_in_smem: the table (ell_bytes) is staged in the CTA-wide shared region at off_ell.
    const int *rs_off;
    const int *rs_pos;       // position of sorted row q inside the stage buffer (chosen on the host against bank conflicts)
    int ell_in_smem, off_ell, ell_bytes;
    // streamed row-split teams: the next column's u and k1 are prefetched from the scratch into shared memory by
    // cp.async.bulk (two buffers of 2 * n * 16 bytes at off_pf inside the team region, completion on mbarriers)
    int rs_prefetch, off_pf;
    const uint8_t *slot_imag;    // value mode 4: per slot, 1 when the slot is purely imaginary (0: purely real)
    int team_warps, teams_per_cta, n_batches, groups, streamed;
    // Column-per-lane mapping of the row-split kernels (propagators with 17..32 rows): lane = column, every warp
    // holds a few whole rows (the same rows in all its lanes), so the sparse-row entries and values are
    // warp-uniform (no padding to the longest row of a warp's lanes, broadcast loads) and the operand gather of a
    // row reads 16 bytes per lane from consecutive addresses (no bank conflicts).
    int cu_map;
    // Split rows (one row per lane, fewer than 32 rows): the longest rows are cut in two and the second half is
    // given to an otherwise idle lane (a "helper": virtual row >= n in g.ell), so a warp executes half the entries of
    // its longest row; the helper's partial sums reach the row's owner by one shuffle per stage.
    // split_info[lane]: bits 0-4 partner lane, bit 5 lane is a helper, bit 6 lane owns a split row.  nullptr: off.
    const uint8_t *split_info;
    int y_row_stride, y_grp_stride;  // stage-buffer layout in 16-byte units (chosen on the host to avoid bank conflicts)
    int nvp, ncp;         // padded slot / channel counts (nvp > n_slots: slot n_slots is the zero slot)
    int common_bytes;     // CTA-wide shared region (series tables + Chebyshev block + dynamic pair lists)
    const double *cheb;   // [NN*NN] coefficient transform, [NN] nodes, [3] check points (NN = 12)
    unsigned dyn_chan_mask;  // channels produced by time-dependent evaluators
    int use_windows;      // 1: drive channels through windowed Chebyshev interpolants (see build_window)
    int dyn_fast;         // 1: every time-dependent slot has <= 3 pairs (64-byte records)
    int fast_tables;      // 1: the time-dependent slots are w0 + wA cA(t) + wB cB(t) with at most two
                          //    time-dependent channels A, B (records hold w0, wA, wB in that order): see fast_tables()
    int fast_chA, fast_chB;
    int off_dptr, off_dw, off_drec;  // byte offsets inside the common region (after tables, Chebyshev, power tables)
    const double *dyn_rec;           // [n_dyn_slots][8] records {w0, w1, w2 (double2), ch0, ch1, ch2, pad (int)}
    // Controller power tables (kernel parameters live in the constant bank; the index is warp-uniform):
    // Taylor coefficients of m^p about the centres of 16 sub-intervals of [1,2) and 2^(pE), for
    // p = -beta1/2 (first) and p = +beta2/2 (second).  Filled on the host per call.
    double pow_coef[2][16][10];
    double pow_exp2[2][128];
    // Component pruning (propagators, u0 = I).  Column c of a propagator is supported on the connected component of c
    // in the generator's graph, so every entry (r, c) with r and c in different components is exactly zero at all
    // times -- in the reference too: its dense products only ever add 0 * x there -- and contributes nothing to the
    // error norms.  The kernel then integrates a VIRTUAL system built on the host: warp w of a team holds g.n local
    // rows (g.ell has team_warps * g.n rows, sources are local row numbers) made of whole units (component, group of
    // CPL of its columns); cf_vrow[w * g.n + r] is the state row of local row r (-1: none) and
    // cf_vcol[(b * g.n + r) * CPL + c] the state column of entry (r, c) in batch b (-1: none; batch b runs on warp
    // b mod team_warps).  Norm divisors and results refer to the n_out x out_cols state the caller sees; the
    // structural zeros of every save point are written once per trajectory from zero_pos (positions col * n_out + row,
    // or positions inside the selection when sel_index is set).
    const int *cf_vrow, *cf_vcol, *zero_pos;
    // Component pruning in the row-split kernels: a batch is a PASS over up to 3 * team_warps chunks of 32 rows taken from
    // several units (component, one of its columns) at once.  The host expands the pass list into one 16-byte record
    // per (pass, chunk, lane): rs_pass[(b * 3 * team_warps + chunk) * 32 + lane] = {state row | state column << 16
    // (-1: no row), byte offset of the lane's entry in the stage buffer, byte offset of its first row-table entry,
    // entries per row | byte offset of the unit in the stage buffer << 8}.  The row tables hold sources relative to the
    // unit's position.  g.n = 96 * team_warps local rows per pass.
    const int4 *rs_pass;
    int n_zero;
    int out_cols;         // columns of the state as the caller sees it (= ncols except under component pruning)
    int team_bytes;       // shared bytes per team
    int off_park;         // lean kernels: offset of the park area inside the team region (5 * E * 512 bytes per warp)
    double reltol, abstol, gamma, qmin, qmax, beta1, beta2, qoldinit, dtmax;
    long long maxiters;
};

// stage-time fractions c2..c5 and c6 = c7 = 1 (stages 6 and 7 share one table)
static __device__ __constant__ double kStageC[5] = {0.161, 0.327, 0.9, 0.9800255409045097, 1.0};

// Tsit5 tableau (OrdinaryDiffEq Tsit5ConstantCache, Float64), in constant memory so that DFMA takes
// the coefficient straight from the constant bank instead of materialising 64-bit immediates.
static __device__ __constant__ double kTsit[28] = {
    0.161,                                                                                        // a21
    -0.008480655492356989, 0.335480655492357,                                                     // a31 a32
    2.8971530571054935, -6.359448489975075, 4.3622954328695815,                                   // a41..a43
    5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525,             // a51..a54
    5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383,  // a61..a65
    0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774,  // a71..a76
    -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
    0.5823571654525552, -0.45808210592918697, 0.015151515151515152};                              // btilde1..7
#define A21 kTsit[0]
#define A31 kTsit[1]
#define A32 kTsit[2]
#define A41 kTsit[3]
#define A42 kTsit[4]
#define A43 kTsit[5]
#define A51 kTsit[6]
#define A52 kTsit[7]
#define A53 kTsit[8]
#define A54 kTsit[9]
#define A61 kTsit[10]
#define A62 kTsit[11]
#define A63 kTsit[12]
#define A64 kTsit[13]
#define A65 kTsit[14]
#define A71 kTsit[15]
#define A72 kTsit[16]
#define A73 kTsit[17]
#define A74 kTsit[18]
#define A75 kTsit[19]
#define A76 kTsit[20]
#define BT1 kTsit[21]
#define BT2 kTsit[22]
#define BT3 kTsit[23]
#define BT4 kTsit[24]
#define BT5 kTsit[25]
#define BT6 kTsit[26]
#define BT7 kTsit[27]

// ---- shared memory through 32-bit addresses ------------------------------------------
__device__ __forceinline__ double2 lds2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t addr, double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double lds1(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts1(uint32_t addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ int ldsi(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t ldsu(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// ---- mbarriers and bulk copies (row-split teams) ----------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    for (int it = 0;; it++) {
        uint32_t ok;
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (it > (1 << 22)) __trap();  // a lost arrival must fail the launch, not hang the device
    }
}
// global -> shared bulk copy (TMA), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
// generic-proxy writes (st.global / st.shared) made visible to later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Branch-free reciprocal square root / reciprocal: hardware seed (MUFU.RSQ64H / MUFU.RCP64H, ~2^-22) and two
// Newton steps (full double precision up to an ulp or two).  The library versions carry special-case
// branches that serialise the three independent per-entry chains of the error norm; arguments here are
// positive normal numbers.
__device__ __forceinline__ double rsqrt_nr(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    return y;
}
__device__ __forceinline__ double rcp_nr(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// min / max of non-negative doubles through their bit patterns (IEEE order = integer order for x >= 0):
// a 64-bit integer compare + select instead of the NaN-aware DSETP.MIN/MAX + fix-up sequence.
__device__ __forceinline__ double pmin(double a, double b) {
    const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return __longlong_as_double(x < y ? x : y);
}
__device__ __forceinline__ double pmax(double a, double b) {
    const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
    return __longlong_as_double(x > y ? x : y);
}

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

// Deterministic team-wide sum of NV values held per thread (fixed order: no atomics).
template <int NV>
__device__ __forceinline__ void team_sum(double (&v)[NV], uint32_t red, int team_warps, int warp_in_team, int lane,
                                         int bar_id) {
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
    if (team_warps == 1) return;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) sts1(red + 8 * (warp_in_team * NV + i), v[i]);
    }
    team_sync(team_warps, bar_id);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double s = 0.0;
        for (int w = 0; w < team_warps; w++) s += lds1(red + 8 * (w * NV + i));
        v[i] = s;
    }
    team_sync(team_warps, bar_id);
}

// Team-wide sum of one value with a single barrier: the partial sums alternate between two buffers (`par`
// flips per call, identically in every warp), so a warp that runs ahead into the next call writes the other
// buffer and cannot reach the call after that before everyone has left this one.
__device__ __forceinline__ double team_sum_alt(double v, uint32_t red, int team_warps, int warp_in_team, int lane,
                                               int bar_id, int &par) {
    v = warp_sum(v);
    if (team_warps == 1) return v;
    const uint32_t buf = red + 8u * (uint32_t)((2 + par) * team_warps);
    par ^= 1;
    if (lane == 0) sts1(buf + 8 * warp_in_team, v);
    team_sync(team_warps, bar_id);
    double s = 0.0;
    for (int w = 0; w < team_warps; w++) s += lds1(buf + 8 * w);
    return s;
}

// Deterministic team-wide maximum.
__device__ __forceinline__ double team_max(double v, uint32_t red, int team_warps, int warp_in_team, int lane, int bar_id) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (team_warps == 1) return v;
    if (lane == 0) sts1(red + 8 * warp_in_team, v);
    team_sync(team_warps, bar_id);
    double m = lds1(red);
    for (int w = 1; w < team_warps; w++) m = fmax(m, lds1(red + 8 * w));
    team_sync(team_warps, bar_id);
    return m;
}

#define QSIM_POW_EMIN (-80)     // controller power tables cover EEst^2 in [2^-80, 2^48)
#define QSIM_POW_NE 128
#define QSIM_POW_K 10          // Taylor coefficients per mantissa sub-interval (16 sub-intervals of [1,2))
static_assert(QSIM_POW_NE == 128 && QSIM_POW_K == 10, "LaunchArgs::pow_coef / pow_exp2 are sized for these");
#ifdef QSIM_TIMING   // development build: per-phase cycle counts of trajectory 0 (printed from the kernel)
#define QSIM_TICK(i) { const long long c_ = clock64(); tk_[i] += c_ - tl_; tl_ = c_; }
#else
#define QSIM_TICK(i)
#endif
#define QSIM_NN 12             // Chebyshev nodes per drive window
#define QSIM_WIN_TOL 2e-14     // accepted interpolation error relative to the channel's magnitude

// Per-thread geometry (all scalars; arrays are kept separately so they stay in registers).
struct Geo {
    int lane, warp_in_team, team_thread, team_threads, bar_id;
    int cpb, colofs;
    // shared-memory byte addresses
    uint32_t sm_tab, sm_cheb, sm_pow, sm_dptr, sm_dchan, sm_dw, sm_drec;  // CTA-wide: series tables, Chebyshev block,
                                                                          // controller power tables, dynamic pair lists / records
    uint32_t sm_V, sm_chan, sm_rev, sm_win, sm_red;  // team (sm_win: per channel 12 nodes, 3 checks, 12 coefficients)
    uint32_t sm_y0, sm_y1;                      // this warp's two stage buffers (+ column offset)
    uint32_t sm_ctl;                            // this warp's control block: step-loop scalars parked in shared memory
    uint32_t sm_park;                           // lean kernels: this lane's slots for k2..k6 of a step with save points
    double stage_c;                             // kStageC[lane] for lanes 0..4 (fast_tables), set once per thread
    int blk_row;                                // blocked variants: this lane's block row
    uint32_t v_stride;                          // bytes between value tables
    // row-split teams: row table in shared memory (0: read it from global memory), prefetch buffers (0: none),
    // mbarriers {stage exchange, prefetch buffer 0, prefetch buffer 1}; phase bits of those three barriers
    uint32_t sm_ell, sm_pf, sm_mbar;
    mutable uint32_t ph_stage, ph_pf;
    int split_partner;                          // split rows: lane to take partial sums from (own lane: none)
    bool split_helper, split_owner;
};

// Stops the compiler from hoisting per-entry address arithmetic out of the step loop (register pressure).
#define QSIM_LAUNDER_ELL()                                                      \
    _Pragma("unroll") for (int i_ = 0; i_ < RPL; i_++) {                        \
        _Pragma("unroll") for (int j_ = 0; j_ < cap(i_); j_++)                  \
            asm volatile("" : "+r"(const_cast<uint32_t &>(ell[i_][j_])));       \
    }

// RS (row split): the rows of one column are spread over all warps of the team (n > 96, e.g. the
// d^2 = 729 Lindblad propagator): stage buffers are team-wide, stage exchange uses the team barrier,
// every warp visits every column, and the sparse rows are read from global memory (L1-resident).
// VM (value mode): 0 complex value slots; 1 purely imaginary generator (slots hold Im A, 2 DFMA per
// multiply-add); 2 as 1 with every off-diagonal entry time-independent: those values live in registers for
// the whole trajectory and only the diagonal is read from the per-stage table.
// LEAN: the Runge-Kutta stage vectors are consumed as early as the tableau allows -- after
// stage 6 the combinations for u_new and for the error estimate are formed and k2..k6 die, so stage 7 runs with five
// live vectors instead of ten -- and operands are requested two row entries at a time.  This fits the step in 168
// registers (three warps per scheduler instead of two).  The k's a dense-output point needs are parked in
// lane-private shared memory on the (rare) steps that contain a save point.
template <int RPL, int CPL, int MAXNNZ, int VM, bool RS, bool LEAN = false>
struct Kern {
    // VM == 5, "tiered": a lane owns RPL rows of one column (like the blocked variants), the rows sorted by length
    // into RPL tiers so that row-slot i of every lane has about the same number of entries; MAXNNZ packs the three
    // tier capacities as octal digits (0532: 5, 3 and 2 entries).  Purely imaginary generator, static off-diagonal
    // values in registers.  Against one row x three columns per lane every gathered operand is one complex number
    // instead of three and no row is padded to the longest row of the matrix.
    static constexpr bool TR = VM == 5;
    // VM == 6: value mode 2 with split rows (LaunchArgs::split_info): the longest rows are cut in two, the second
    // half runs on an otherwise idle lane of the same column group and reaches the owner by one shuffle per stage.
    static constexpr bool SPL = VM == 6;
    // VM == 7: value mode 2 on a generator with ADDITIVELY split rows.  A long row p is carried as two table rows whose
    // states add up to the physical entry, u_p = u_pa + u_pb: each part integrates the diagonal term on itself plus its
    // half of the row's couplings (the equation is linear, so the sum obeys the original one), and every row that
    // couples to p reads both parts.  The parts live on different lanes (the second on an otherwise idle one) and
    // exchange nothing during the stages; only the error norm, the initial-step norms and the results need the
    // physical entry and combine the parts by shuffles (once per step instead of once per stage as in mode 6).
    // LaunchArgs::split_info marks owner / helper lanes, row_perm maps table rows to state rows, n_out counts them.
    static constexpr bool AS = VM == 7;
    static constexpr bool IM = (VM >= 1 && VM <= 3) || TR || SPL || AS;
    static constexpr bool SV = VM == 2 || TR || SPL || AS;
    static constexpr int NZ = TR ? (MAXNNZ >> 6) : MAXNNZ;   // capacity of the per-lane row tables
    static __host__ __device__ constexpr int cap(int i) {
        return !TR ? MAXNNZ : i == 0 ? (MAXNNZ >> 6) : i == 1 ? ((MAXNNZ >> 3) & 7) : (MAXNNZ & 7);
    }
    // VM == 4, "mixed": a complex generator whose every OFF-DIAGONAL entry is purely real or purely imaginary
    // (Lindblad generators of real Hamiltonians and real collapse operators: couplings come from the commutator and
    // are imaginary, the sandwich terms L rho L' are real; only the diagonal mixes both).  A value table is stored
    // split -- real parts of all slots, then imaginary parts -- so an off-diagonal entry loads the one 8-byte part
    // it has (its byte offset points into the right half, bit 31 of the entry says which half) and costs 2 DFMA
    // instead of 4; the diagonal entry loads both halves.
    static constexpr bool MX = VM == 4;
    static constexpr uint32_t VSTRIDE = ((VM >= 1 && VM <= 3) || VM >= 5) ? 8u : 16u;  // table bytes per slot
    // VM == 3, "blocked": a lane owns RPL consecutive rows of one column and the off-diagonal part of the
    // generator is stored as dense RPL x RPL blocks (static, purely imaginary, values in registers).  Entry
    // 1 + b of row-slot k names row k of neighbour block b, so the RPL row-slots together load each
    // neighbour block exactly once: RPL * (MAXNNZ - 1) operand loads serve RPL * RPL * (MAXNNZ - 1)
    // multiply-adds (Kronecker-structured couplings reuse every operand RPL times).
    static constexpr bool BLK = VM == 3;
    static constexpr int NVS = BLK ? (MAXNNZ - 1) * RPL * RPL : RPL * NZ;  // static values held per lane
    static_assert(!(BLK || TR) || CPL == 1, "blocked and tiered variants hold one column per lane");
    static_assert(!TR || RPL == 3, "tiered variants: three rows per lane");
    static_assert(!AS || (LEAN && !RS && RPL == 1), "additively split rows: lean one-row-per-lane kernels");

    // the physical entry of an additively split row: the owner lane's part plus its helper's (other lanes: unchanged)
    static __device__ __forceinline__ double as_sum(const Geo &g, double v) {
        const double p = __shfl_sync(0xffffffffu, v, g.split_partner);
        return g.split_owner ? v + p : v;
    }
    static constexpr int E = RPL * CPL;
    static constexpr uint32_t VB = (IM || MX) ? 8u : 16u;  // bytes per value slot (IM / MX: one real number)
    typedef double2 Vec[E];

    // Refresh tables table0 + s, s < NTAB, for the times tbase + kStageC[s]*dtv.  all == false: only
    // the time-dependent channels and slots (pair lists in shared memory); all == true: everything
    // (pair lists from global memory; once per trajectory).  Two team barriers.
    // x^(-beta1/2) and x^(beta2/2) for x = EEst^2: split x = m 2^E, a degree-9 Taylor polynomial of m^p about the
    // centre of m's sub-interval (16 per octave; remainder < 1e-17 relative) times a tabulated 2^(pE).  Replaces a
    // log and an exp on the per-step critical path; false outside the tabulated exponent range (0, inf, nan too).
    static __device__ __forceinline__ bool pow_pair(const LaunchArgs &a, double x, double &ra, double &rb) {
        const long long b = __double_as_longlong(x);
        const int ex = (int)((b >> 52) & 0x7ff) - 1023;
        if (b < 0 || ex < QSIM_POW_EMIN || ex >= QSIM_POW_EMIN + QSIM_POW_NE) return false;
        const int idx = (int)((b >> 48) & 15);
        const double m = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
        const double t = m - (1.03125 + 0.0625 * idx);

        double ca[QSIM_POW_K], cb[QSIM_POW_K];
#pragma unroll
        for (int k = 0; k < QSIM_POW_K; k++) {  // constant-bank loads with a warp-uniform index
            ca[k] = a.pow_coef[0][idx][k];
            cb[k] = a.pow_coef[1][idx][k];
        }
        const double ta = a.pow_exp2[0][ex - QSIM_POW_EMIN], tb = a.pow_exp2[1][ex - QSIM_POW_EMIN];
        // Estrin evaluation of the two degree-9 polynomials (dependency depth 5 instead of 9)
        static_assert(QSIM_POW_K == 10, "Estrin scheme below is written for 10 coefficients");
        const double t2 = t * t, t4 = t2 * t2, t8 = t4 * t4;
        const double a01 = fma(ca[1], t, ca[0]), a23 = fma(ca[3], t, ca[2]), a45 = fma(ca[5], t, ca[4]),
                     a67 = fma(ca[7], t, ca[6]), a89 = fma(ca[9], t, ca[8]);
        const double b01 = fma(cb[1], t, cb[0]), b23 = fma(cb[3], t, cb[2]), b45 = fma(cb[5], t, cb[4]),
                     b67 = fma(cb[7], t, cb[6]), b89 = fma(cb[9], t, cb[8]);
        const double qa = fma(t8, a89, fma(t4, fma(t2, a67, a45), fma(t2, a23, a01)));
        const double qb = fma(t8, b89, fma(t4, fma(t2, b67, b45), fma(t2, b23, b01)));
        ra = qa * ta;
        rb = qb * tb;
        return true;
    }

    static __device__ __forceinline__ void load_resolved(uint32_t addr, ResolvedEval &r) {
        union {
            ResolvedEval r;
            double2 q[10];
        } ld;
#pragma unroll
        for (int i = 0; i < 10; i++) ld.q[i] = lds2(addr + 16 * i);
        r = ld.r;
    }

    // Chebyshev polynomials T_0..T_{NN-1} at x (three-term recurrence: one dependent FMA per degree).
    static __device__ __forceinline__ void cheb_basis(double x, double (&T)[QSIM_NN]) {
        const double tx = 2.0 * x;
        T[0] = 1.0;
        T[1] = x;
#pragma unroll
        for (int k = 2; k < QSIM_NN; k++) T[k] = fma(tx, T[k - 1], -T[k - 2]);
    }
    // sum_k c_k T_k with the coefficients in shared memory; two partial sums for instruction-level parallelism
    static __device__ __forceinline__ double cheb_dot(uint32_t coef, const double (&T)[QSIM_NN]) {
        double c[QSIM_NN];
#pragma unroll
        for (int k = 0; k < QSIM_NN; k++) c[k] = lds1(coef + 8 * k);  // all loads in flight before the first use
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int k = 0; k < QSIM_NN; k += 2) {
            p0 = fma(c[k], T[k], p0);
            p1 = fma(c[k + 1], T[k + 1], p1);
        }
        return p0 + p1;
    }

    // Drive window: evaluate every time-dependent channel at the 12 Chebyshev nodes of [Tw, Tw+W]
    // (one SIMT pass, a lane per node) plus 3 check points, turn the node values into Chebyshev
    // coefficients, and return the interpolant's worst error at the check points relative to the
    // channel's magnitude.  Stage times inside the window then cost a 12-term recurrence instead of
    // the sin/cos/sqrt/Horner chain; the caller shrinks the window until the check passes or falls
    // back to direct evaluation, so the interpolant is only used where it is exact to rounding.
    static __device__ __forceinline__ double build_window(const LaunchArgs &a, const Geo &g, const double *smem_gen,
                                                         double Tw, double W) {
        const int ne = a.g.n_evals;
        const SeriesTables *tab = reinterpret_cast<const SeriesTables *>(smem_gen);
        for (int it = g.team_thread; it < (QSIM_NN + 3) * ne; it += g.team_threads) {
            const int j = it % (QSIM_NN + 3), e = it / (QSIM_NN + 3);
            ResolvedEval r;
            load_resolved(g.sm_rev + (uint32_t)sizeof(ResolvedEval) * e, r);
            if (r.dynamic) {
                const double x = lds1(g.sm_cheb + 8 * (QSIM_NN * QSIM_NN + j));
                double o0, o1;
                run_resolved(r, tab->freq, tab->anh, Tw + 0.5 * W * (1.0 + x), &o0, &o1);
                sts1(g.sm_win + 256u * r.out_chan + 8 * j, o0);
                if (eval_outputs(r.kind) == 2) sts1(g.sm_win + 256u * (r.out_chan + 1) + 8 * j, o1);
            }
        }
        team_sync(a.team_warps, g.bar_id);
        for (int it = g.team_thread; it < QSIM_NN * a.g.n_chan; it += g.team_threads) {
            const int k = it % QSIM_NN, ch = it / QSIM_NN;
            if ((a.dyn_chan_mask >> ch) & 1u) {
                double c = 0.0;
#pragma unroll
                for (int j = 0; j < QSIM_NN; j++)
                    c = fma(lds1(g.sm_cheb + 8 * (k * QSIM_NN + j)), lds1(g.sm_win + 256u * ch + 8 * j), c);
                sts1(g.sm_win + 256u * ch + 128 + 8 * k, c);
            }
        }
        team_sync(a.team_warps, g.bar_id);
        double bad = 0.0;
        for (int it = g.team_thread; it < 3 * a.g.n_chan; it += g.team_threads) {
            const int i = it % 3, ch = it / 3;
            if ((a.dyn_chan_mask >> ch) & 1u) {
                const double x = lds1(g.sm_cheb + 8 * (QSIM_NN * QSIM_NN + QSIM_NN + i));
                double T[QSIM_NN];
                cheb_basis(x, T);
                const double p = cheb_dot(g.sm_win + 256u * ch + 128, T);
                const double f = lds1(g.sm_win + 256u * ch + 8 * (QSIM_NN + i));
                double scale = 0.0;
#pragma unroll
                for (int j = 0; j < QSIM_NN; j++) scale = fmax(scale, fabs(lds1(g.sm_win + 256u * ch + 8 * j)));
                const double err = fabs(p - f) / (scale > 0.0 ? scale : 1.0);
                bad = fmax(bad, err == err ? err : 1.0);  // NaN counts as failure
            }
        }
        return team_max(bad, g.sm_red, a.team_warps, g.warp_in_team, g.lane, g.bar_id);
    }

    // Stage tables for the common case (<= 2 time-dependent channels, drive window valid):
    // lane s < 5 evaluates both channels' Chebyshev interpolants at stage time s (basis by the double-step
    // recurrence T_{k+2} = 2 T_2 T_k - T_{k-2}: two independent chains), the values are broadcast by shuffle,
    // and every lane combines them with its slots' weights.  No shared-memory round trip, no barrier: the
    // stage exchange's __syncwarp orders the table stores before the first gather.  In a multi-warp team every
    // warp writes the whole (identical) table itself before it reads any of it, and the error-norm barrier of
    // the previous step separates these stores from that step's reads, so no team barrier is needed either.
    static __device__ __forceinline__ void fast_tables(const LaunchArgs &a, const Geo &g, double t, double dt,
                                                       double win_t, double win_sc) {
        double cA = 0.0, cB = 0.0;
        if (g.lane < 5) {
            const double x = fma((t + g.stage_c * dt) - win_t, win_sc, -1.0);  // stage-time fraction c_{lane+2}
            const uint32_t fa = g.sm_win + 256u * a.fast_chA + 128, fb = g.sm_win + 256u * a.fast_chB + 128;
            double ka[QSIM_NN], kb[QSIM_NN];
#pragma unroll
            for (int k = 0; k < QSIM_NN; k++) {  // all loads in flight first
                ka[k] = lds1(fa + 8 * k);
                kb[k] = lds1(fb + 8 * k);
            }
            double T[QSIM_NN];
            T[0] = 1.0;
            T[1] = x;
            const double c2 = fma(4.0 * x, x, -2.0);  // 2 T_2(x)
            T[2] = 0.5 * c2;
            T[3] = fma(c2, x, -x);
#pragma unroll
            for (int k = 4; k < QSIM_NN; k++) T[k] = fma(c2, T[k - 2], -T[k - 4]);
            double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
            for (int k = 0; k < QSIM_NN; k += 2) {
                a0 = fma(ka[k], T[k], a0);
                a1 = fma(ka[k + 1], T[k + 1], a1);
                b0 = fma(kb[k], T[k], b0);
                b1 = fma(kb[k + 1], T[k + 1], b1);
            }
            cA = a0 + a1;
            cB = b0 + b1;
        }
        const int n_items = 5 * a.g.n_dyn_slots;
        for (int base = 0; base < n_items; base += 32) {  // uniform trip count: every lane joins the shuffles
            const int it = base + g.lane;
            const int s = it % 5, sl = it / 5;
            const double fA = __shfl_sync(0xffffffffu, cA, s), fB = __shfl_sync(0xffffffffu, cB, s);
            if (it < n_items) {
                const uint32_t rec = g.sm_drec + 64u * sl;
                const double2 w0 = lds2(rec), w1 = lds2(rec + 16), w2 = lds2(rec + 32);
                const double vy = fma(w2.y, fB, fma(w1.y, fA, w0.y));
                if (IM) {
                    sts1(g.sm_V + s * g.v_stride + 8 * sl, vy);
                } else {
                    const double vx = fma(w2.x, fB, fma(w1.x, fA, w0.x));
                    if (MX)
                        {
                        sts1(g.sm_V + s * g.v_stride + 8 * sl, vx);
                        sts1(g.sm_V + s * g.v_stride + 8 * (a.nvp + sl), vy);
                    }
                    else
                        sts2(g.sm_V + s * g.v_stride + 16 * sl, make_double2(vx, vy));
                }
            }
        }
    }

    template <int NTAB>
    static __device__ __forceinline__ void compute_tables(const LaunchArgs &a, const Geo &g, const double *smem_gen,
                                                           double tbase, double dtv, bool all, int table0,
                                                           bool use_win = false, double win_t = 0.0, double win_w = 1.0) {
        const int ne = a.g.n_evals;
        const SeriesTables *tab = reinterpret_cast<const SeriesTables *>(smem_gen);
        if (use_win) {
            // stage times inside the current drive window: Chebyshev interpolants
            // one lane per stage time: the Chebyshev basis is shared by all channels
            const double sc = 2.0 / win_w;
            for (int s = g.team_thread; s < NTAB; s += g.team_threads) {
                double T[QSIM_NN];
                cheb_basis(fma((tbase + kStageC[s] * dtv) - win_t, sc, -1.0), T);
                for (unsigned m = a.dyn_chan_mask; m; m &= m - 1) {
                    const int ch = __ffs(m) - 1;
                    sts1(g.sm_chan + 8u * ((table0 + s) * a.ncp + ch), cheb_dot(g.sm_win + 256u * ch + 128, T));
                }
            }
        } else
        for (int it = g.team_thread; it < NTAB * ne; it += g.team_threads) {
            const int s = it % NTAB, e = it / NTAB;
            ResolvedEval r;
            load_resolved(g.sm_rev + (uint32_t)sizeof(ResolvedEval) * e, r);
            if (all || r.dynamic) {
                double o0, o1;
                run_resolved(r, tab->freq, tab->anh, tbase + kStageC[s] * dtv, &o0, &o1);
                const uint32_t ca = g.sm_chan + 8u * ((table0 + s) * a.ncp + r.out_chan);
                sts1(ca, o0);
                if (eval_outputs(r.kind) == 2) sts1(ca + 8, o1);
            }
        }
        team_sync(a.team_warps, g.bar_id);
        if (!all) {
            const int nsl = a.g.n_dyn_slots;
            if (a.dyn_fast) {
                // every time-dependent slot has at most 3 (channel, weight) pairs: fixed 64-byte records,
                // no dependent index loads
                for (int it = g.team_thread; it < NTAB * nsl; it += g.team_threads) {
                    const int s = it % NTAB, sl = it / NTAB;
                    const uint32_t ch = g.sm_chan + 8u * ((table0 + s) * a.ncp);
                    const uint32_t rec = g.sm_drec + 64u * sl;
                    const double2 w0 = lds2(rec), w1 = lds2(rec + 16), w2 = lds2(rec + 32);
                    const double c0 = lds1(ch + 8u * ldsi(rec + 48)), c1 = lds1(ch + 8u * ldsi(rec + 52)),
                                 c2 = lds1(ch + 8u * ldsi(rec + 56));
                    const double vy = fma(w2.y, c2, fma(w1.y, c1, w0.y * c0));
                    if (IM) {
                        sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vy);
                    } else {
                        const double vx = fma(w2.x, c2, fma(w1.x, c1, w0.x * c0));
                        if (MX) {
                            sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vx);
                            sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * (a.nvp + sl), vy);
                        } else
                            sts2(g.sm_V + (table0 + s) * g.v_stride + 16 * sl, make_double2(vx, vy));
                    }
                }
            } else
            for (int it = g.team_thread; it < NTAB * nsl; it += g.team_threads) {
                const int s = it % NTAB, sl = it / NTAB;
                const uint32_t ch = g.sm_chan + 8u * ((table0 + s) * a.ncp);
                double vx = 0.0, vy = 0.0;
                const int q1 = ldsi(g.sm_dptr + 4 * (sl + 1));
                for (int q = ldsi(g.sm_dptr + 4 * sl); q < q1; q++) {
                    const double c = lds1(ch + 8u * ldsi(g.sm_dchan + 4 * q));
                    const double2 w = lds2(g.sm_dw + 16 * q);
                    vx = fma(w.x, c, vx);
                    vy = fma(w.y, c, vy);
                }
                if (IM)
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vy);
                else if (MX) {
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vx);
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * (a.nvp + sl), vy);
                } else
                    sts2(g.sm_V + (table0 + s) * g.v_stride + 16 * sl, make_double2(vx, vy));
            }
        } else {
            const int nsl = a.g.n_slots;
            for (int it = g.team_thread; it < NTAB * nsl; it += g.team_threads) {
                const int s = it % NTAB, sl = it / NTAB;
                const uint32_t ch = g.sm_chan + 8u * ((table0 + s) * a.ncp);
                double vx = 0.0, vy = 0.0;
                for (int q = a.g.slot_ptr[sl]; q < a.g.slot_ptr[sl + 1]; q++) {
                    const double c = lds1(ch + 8u * a.g.pair_chan[q]);
                    const double2 w = a.g.pair_w[q];
                    vx = fma(w.x, c, vx);
                    vy = fma(w.y, c, vy);
                }
                if (IM)
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vy);
                else if (MX) {
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * sl, vx);
                    sts1(g.sm_V + (table0 + s) * g.v_stride + 8 * (a.nvp + sl), vy);
                } else
                    sts2(g.sm_V + (table0 + s) * g.v_stride + 16 * sl, make_double2(vx, vy));
            }
        }
        team_sync(a.team_warps, g.bar_id);
    }

    static __device__ __forceinline__ void put_stage(const LaunchArgs &a, const Geo &g, const Vec &y, uint32_t yb,
                                                     const uint32_t (&my_off)[RPL], const bool (&row_ok)[RPL]) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (row_ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) sts2(yb + my_off[i] + 16 * c, y[i * CPL + c]);
            }
        if (RS) {
            // split barrier: announce this warp's part of the stage vector now, wait for the others' just before
            // the gather (apply); the work in between (the next stage's partial combination) hides the latency
            __syncwarp();
            if (g.lane == 0) mbar_arrive(g.sm_mbar);
        } else
            __syncwarp();
    }

    // k = A(t_s) * y : sparse row gather, value table Vs, stage vector yb (both shared-memory addresses).
    // Entry 0 of every row is its diagonal: that operand is the lane's own stage value `yreg`.
    // value of a row entry (byte offset of its slot in bits 16..; MX: bit 31 is the "imaginary" flag)
    static __device__ __forceinline__ double2 ldv(uint32_t Vs, uint32_t e) {
        if (MX) return make_double2(0.0, lds1(Vs + ((e >> 16) & 0x7FFFu)));
        if (IM) return make_double2(0.0, lds1(Vs + (e >> 16)));
        return lds2(Vs + (e >> 16));
    }
    // the diagonal entry: complex in mixed mode (both halves of the split table)
    static __device__ __forceinline__ double2 ldv_diag(const LaunchArgs &a, uint32_t Vs, uint32_t e) {
        if (MX) return make_double2(lds1(Vs + (e >> 16)), lds1(Vs + (e >> 16) + 8u * a.nvp));
        return ldv(Vs, e);
    }
    static __device__ __forceinline__ void cfma_diag(double2 &acc, double2 v, double2 y) {
        if (IM) {
            acc.x = fma(-v.y, y.y, acc.x);
            acc.y = fma(v.y, y.x, acc.y);
        } else {
            acc.x = fma(v.x, y.x, acc.x);
            acc.y = fma(v.x, y.y, acc.y);
            acc.x = fma(-v.y, y.y, acc.x);
            acc.y = fma(v.y, y.x, acc.y);
        }
    }
    // acc += v * y for one row entry
    static __device__ __forceinline__ void cfma(double2 &acc, double2 v, uint32_t e, double2 y) {
        if (MX) {
            const bool im = (e >> 31) != 0;
            const double a0 = im ? -y.y : y.x, a1 = im ? y.x : y.y;
            acc.x = fma(v.y, a0, acc.x);
            acc.y = fma(v.y, a1, acc.y);
        } else if (IM) {
            // A = i B with B real: (iB)(yr + i yi) = -B yi + i B yr
            acc.x = fma(-v.y, y.y, acc.x);
            acc.y = fma(v.y, y.x, acc.y);
        } else {
            acc.x = fma(v.x, y.x, acc.x);
            acc.y = fma(v.x, y.y, acc.y);
            acc.x = fma(-v.y, y.y, acc.x);
            acc.y = fma(v.y, y.x, acc.y);
        }
    }

    static __device__ __forceinline__ void mac(const LaunchArgs &a, uint32_t Vs, uint32_t yb, uint32_t e, bool diag,
                                               const double2 *yreg, double2 (&acc)[CPL]) {
        const uint32_t ya = yb + (e & 0xFFFFu);
        if (diag) {
            const double2 v = ldv_diag(a, Vs, e);
#pragma unroll
            for (int c = 0; c < CPL; c++) cfma_diag(acc[c], v, yreg[c]);
        } else {
            const double2 v = ldv(Vs, e);
#pragma unroll
            for (int c = 0; c < CPL; c++) cfma(acc[c], v, e, lds2(ya + 16 * c));
        }
    }

    struct NoHook {
        __device__ __forceinline__ void operator()() const {}
    };
    // `hook` runs between the issue of a row's operand loads and their first use: independent work placed
    // there (the next stage's partial combination) executes under the shared-memory latency.
    // CHN: off-diagonal entries whose operands are requested together (0: the kernel's default -- QSIM_CH, or 2 in the
    // lean kernels, whose early stages, with few live vectors, ask for 4)
    template <int CHN = 0, typename Hook = NoHook>
    static __device__ __forceinline__ void apply(const LaunchArgs &a, const Geo &g, uint32_t Vs, uint32_t yb,
                                                 const uint32_t (&ell)[RPL][NZ], const double (&vstat)[NVS],
                                                 const int (&row_of)[RPL], const Vec &yreg, Vec &k, Hook hook = Hook()) {
        if (BLK) {
            constexpr int MB = MAXNNZ > 1 ? MAXNNZ - 1 : 1;  // (row-split variants instantiate MAXNNZ = 1; never blocked)
            double vd[RPL];
            double2 yy[MB][RPL], acc[RPL];
#pragma unroll
            for (int i = 0; i < RPL; i++) vd[i] = lds1(Vs + (ell[i][0] >> 16));
#pragma unroll
            for (int b = 0; b < MB; b++)
#pragma unroll
                for (int kk = 0; kk < RPL; kk++) yy[b][kk] = lds2(yb + (ell[kk][1 + b] & 0xFFFFu));
            hook();
#pragma unroll
            for (int i = 0; i < RPL; i++) {
                // A = i B with B real: (iB)(yr + i yi) = -B yi + i B yr
                acc[i].x = -vd[i] * yreg[i].y;
                acc[i].y = vd[i] * yreg[i].x;
            }
#pragma unroll
            for (int b = 0; b < MB; b++)
#pragma unroll
                for (int kk = 0; kk < RPL; kk++)
#pragma unroll
                    for (int i = 0; i < RPL; i++) {
                        const double v = vstat[(b * RPL + i) * RPL + kk];
                        acc[i].x = fma(-v, yy[b][kk].y, acc[i].x);
                        acc[i].y = fma(v, yy[b][kk].x, acc[i].y);
                    }
#pragma unroll
            for (int i = 0; i < RPL; i++) k[i] = acc[i];
            return;
        }
        if (TR) {
            // diagonal values (time-dependent: from the stage's table), then the operands of row-slot 0 and, as a
            // second group, of the shorter row-slots; `hook` runs under the first group's loads
            double vd[RPL];
            double2 acc[RPL];
#pragma unroll
            for (int i = 0; i < RPL; i++) vd[i] = lds1(Vs + (ell[i][0] >> 16));
            double2 ya[NZ > 1 ? NZ - 1 : 1];
#pragma unroll
            for (int j = 1; j < cap(0); j++) ya[j - 1] = lds2(yb + (ell[0][j] & 0xFFFFu));
            hook();
#pragma unroll
            for (int i = 0; i < RPL; i++) {
                acc[i].x = -vd[i] * yreg[i].y;
                acc[i].y = vd[i] * yreg[i].x;
            }
            double2 yr[RPL > 1 ? RPL - 1 : 1][NZ > 1 ? NZ - 1 : 1];
#pragma unroll
            for (int i = 1; i < RPL; i++)
#pragma unroll
                for (int j = 1; j < cap(i); j++) yr[i - 1][j - 1] = lds2(yb + (ell[i][j] & 0xFFFFu));
#pragma unroll
            for (int j = 1; j < cap(0); j++) {
                const double v = vstat[j];
                acc[0].x = fma(-v, ya[j - 1].y, acc[0].x);
                acc[0].y = fma(v, ya[j - 1].x, acc[0].y);
            }
#pragma unroll
            for (int i = 1; i < RPL; i++)
#pragma unroll
                for (int j = 1; j < cap(i); j++) {
                    const double v = vstat[i * NZ + j];
                    acc[i].x = fma(-v, yr[i - 1][j - 1].y, acc[i].x);
                    acc[i].y = fma(v, yr[i - 1][j - 1].x, acc[i].y);
                }
#pragma unroll
            for (int i = 0; i < RPL; i++) k[i] = acc[i];
            return;
        }
        if (RS) {
            // independent work first, then wait for every warp's part of the stage vector (put_stage arrived)
            hook();
            mbar_wait(g.sm_mbar, g.ph_stage);
            g.ph_stage ^= 1u;
        }
#pragma unroll
        for (int i = 0; i < RPL; i++) {
            double2 acc[CPL];
#pragma unroll
            for (int c = 0; c < CPL; c++) acc[c] = make_double2(0.0, 0.0);
            if (RS) {
                // Row table (shared memory, or global memory when it does not fit): entry j of this lane's row in
                // chunk (slot i, warp) sits 128 bytes after entry j - 1.  The chunk's rows all have `nnz` entries
                // (padded on the host to 1 + an even count, plus one spare pair), so the off-diagonal entries are
                // taken two at a time with the next pair's table words requested ahead of the multiply-adds.
                const int nnz = (int)ell[i][MAXNNZ > 1 ? 1 : 0];
                const uint32_t yb_full = yb;
                const uint32_t yb = yb_full + (MAXNNZ > 2 ? ell[i][MAXNNZ > 2 ? 2 : 0] : 0u);   // + the unit's position (component passes)
                if (g.sm_ell) {
                    const uint32_t tb = g.sm_ell + ell[i][0];
                    uint32_t ea = ldsu(tb + 128u), eb = ldsu(tb + 256u);
                    mac(a, Vs, yb, ldsu(tb), true, &yreg[i * CPL], acc);
                    for (int j = 1; j < nnz; j += 2) {
                        const uint32_t na = ldsu(tb + 128u * (uint32_t)(j + 2)), nb = ldsu(tb + 128u * (uint32_t)(j + 3));
                        const double2 va = ldv(Vs, ea), vb = ldv(Vs, eb);
                        double2 ya[CPL], yb2[CPL];
#pragma unroll
                        for (int c = 0; c < CPL; c++) {
                            ya[c] = lds2(yb + (ea & 0xFFFFu) + 16 * c);
                            yb2[c] = lds2(yb + (eb & 0xFFFFu) + 16 * c);
                        }
#pragma unroll
                        for (int c = 0; c < CPL; c++) {
                            cfma(acc[c], va, ea, ya[c]);
                            cfma(acc[c], vb, eb, yb2[c]);
                        }
                        ea = na;
                        eb = nb;
                    }
                } else {
                    const uint32_t *er = a.ell_packed + (ell[i][0] >> 2);
                    mac(a, Vs, yb, __ldg(er), true, &yreg[i * CPL], acc);
#pragma unroll 4
                    for (int j = 1; j < nnz; j++) mac(a, Vs, yb, __ldg(er + 32 * j), false, &yreg[i * CPL], acc);
                }
            } else {
                // no per-entry branch (short rows are padded with the zero slot).  Operands are requested in
                // chunks of up to four off-diagonal entries, all loads of a chunk ahead of its multiply-adds, so a
                // row of five exposes one shared-memory latency and longer rows stay within the register budget.
                {
                    // diagonal entry: operand from the lane's own registers
                    const uint32_t e = ell[i][0];
                    const double2 vd = ldv_diag(a, Vs, e);
                    if (i == RPL - 1 && MAXNNZ <= 1) hook();
#pragma unroll
                    for (int c = 0; c < CPL; c++) cfma_diag(acc[c], vd, yreg[i * CPL + c]);
                }
#ifndef QSIM_CH
#define QSIM_CH 4
#endif
#ifndef QSIM_LEAN_CH_EARLY   // lean kernels, stages 2-4 (few live vectors): entries requested together
#define QSIM_LEAN_CH_EARLY 2     /* (4 measured slower on cfg2: 571 ms against 555 ms) */
#endif
#ifndef QSIM_LEAN_CH         // lean kernels, stages 6-7
#define QSIM_LEAN_CH 2
#endif
#ifndef QSIM_LEAN_CH_MID     // lean kernels, stage 5
#define QSIM_LEAN_CH_MID 2
#endif
                constexpr int CH = CHN > 0 ? CHN : (LEAN ? QSIM_LEAN_CH : QSIM_CH);
#pragma unroll
                for (int j0 = 1; j0 < MAXNNZ; j0 += CH) {
                    double2 vv[CH], yy[CH][CPL];
#pragma unroll
                    for (int q = 0; q < CH; q++) {
                        const int j = j0 + q;
                        if (j < MAXNNZ) {
                            const uint32_t e = ell[i][j];
                            if (SV)
                                vv[q] = make_double2(0.0, vstat[i * NZ + j]);  // time-independent, loaded once per trajectory
                            else
                                vv[q] = ldv(Vs, e);
#pragma unroll
                            for (int c = 0; c < CPL; c++) yy[q][c] = lds2(yb + (e & 0xFFFFu) + 16 * c);
                        }
                    }
                    if (i == RPL - 1 && j0 == 1) hook();
#pragma unroll
                    for (int q = 0; q < CH; q++) {
                        if (j0 + q < MAXNNZ) {
#pragma unroll
                            for (int c = 0; c < CPL; c++) cfma(acc[c], vv[q], ell[i][j0 + q], yy[q][c]);
                        }
                    }
                }
            }
#ifdef QSIM_SPLIT_ROWS   // (experimental build: run-time switch for every row-per-lane kernel; it splits the gather's basic block)
            const bool do_split = !RS && !BLK && (SPL || a.split_info);
#else
            constexpr bool do_split = SPL;
#endif
            if (do_split) {
                // split rows: the owner adds its helper's half, the helper keeps no state of its own
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    const double px = __shfl_sync(0xffffffffu, acc[c].x, g.split_partner),
                                 py = __shfl_sync(0xffffffffu, acc[c].y, g.split_partner);
                    acc[c].x = g.split_helper ? 0.0 : acc[c].x + (g.split_owner ? px : 0.0);
                    acc[c].y = g.split_helper ? 0.0 : acc[c].y + (g.split_owner ? py : 0.0);
                }
            }
#pragma unroll
            for (int c = 0; c < CPL; c++) k[i * CPL + c] = acc[c];
        }
    }

    // state column of entry (row-slot i, c) of the batch starting at virtual column col0 (-1: none)
    static __device__ __forceinline__ int out_col(const LaunchArgs &a, int col0, int row_word, int c) {
        if (RS && a.rs_pass) return row_word >> 16;   // (row-split passes: the chunk's column rides in the row word)
        if (a.cf_vcol) return __ldg(a.cf_vcol + (size_t)col0 * a.g.n + (size_t)(row_word >> 16) * CPL + c);
        return col0 + c < a.ncols ? col0 + c : -1;
    }

    // row_of[i]: state row in bits 0..15 (component pruning: the warp-local row in bits 16..)
    static __device__ __forceinline__ void write_out(const LaunchArgs &a, const Geo &g, long long traj, int ksave, int col0,
                                                     const int (&row_of)[RPL], const bool (&row_ok)[RPL], const Vec &x_in) {
        const int n = a.n_out;
        Vec xs;
        bool ok[RPL];
#pragma unroll
        for (int i = 0; i < RPL; i++) ok[i] = row_ok[i] && !(AS && g.split_helper);
#pragma unroll
        for (int e = 0; e < E; e++)
            xs[e] = AS ? make_double2(as_sum(g, x_in[e].x), as_sum(g, x_in[e].y)) : x_in[e];
        const Vec &x = xs;
        if (a.sel_index) {
            // reduced output: only the selected entries leave the chip
            const size_t base = ((size_t)traj * a.n_ts + ksave) * (size_t)a.n_sel;
#pragma unroll
            for (int i = 0; i < RPL; i++)
                if (ok[i]) {
#pragma unroll
                    for (int c = 0; c < CPL; c++) {
                        const int col = out_col(a, col0, row_of[i], c);
                        if (col >= 0) {
                            const int si = __ldg(a.sel_index + (size_t)col * n + (row_of[i] & 0xFFFF));
                            const double2 v = x[i * CPL + c];
                            if (si >= 0) {
                                if (a.sel_abs2)
                                    reinterpret_cast<double *>(a.out)[base + si] = fma(v.x, v.x, v.y * v.y);
                                else
                                    a.out[base + si] = v;
                            }
                        }
                    }
                }
            return;
        }
        double2 *dst = a.out + ((size_t)traj * a.n_ts + ksave) * (size_t)a.out_cols * n;
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) {
                    const int col = out_col(a, col0, row_of[i], c);
                    if (col >= 0) dst[(size_t)col * n + (row_of[i] & 0xFFFF)] = x[i * CPL + c];
                }
            }
    }

    // Component pruning: the structurally zero entries of save points [k_lo, k_hi) (value 0, or NaN for the unreached
    // save points of a failed trajectory).  Positions are disjoint from every entry the lanes write.
    static __device__ __forceinline__ void fill_zeros(const LaunchArgs &a, const Geo &g, long long traj, int k_lo, int k_hi,
                                                      double val) {
        for (int ks = k_lo; ks < k_hi; ks++) {
            if (a.sel_index) {
                const size_t base = ((size_t)traj * a.n_ts + ks) * (size_t)a.n_sel;
                for (int z = g.team_thread; z < a.n_zero; z += g.team_threads) {
                    const int si = __ldg(a.zero_pos + z);
                    if (a.sel_abs2)
                        reinterpret_cast<double *>(a.out)[base + si] = val;
                    else
                        a.out[base + si] = make_double2(val, val);
                }
            } else {
                double2 *dst = a.out + ((size_t)traj * a.n_ts + ks) * (size_t)a.out_cols * a.n_out;
                for (int z = g.team_thread; z < a.n_zero; z += g.team_threads)
                    dst[__ldg(a.zero_pos + z)] = make_double2(val, val);
            }
        }
    }

    static __device__ __forceinline__ void initial_state(const LaunchArgs &a, const Geo &g, long long traj, int col0,
                                                         const int (&row_of)[RPL], const bool (&row_ok)[RPL], Vec &u) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
#pragma unroll
            for (int c = 0; c < CPL; c++) {
                double2 v = make_double2(0.0, 0.0);
                const int col = out_col(a, col0, row_of[i], c);
                if (row_ok[i] && col >= 0 && !(AS && g.split_helper)) {   // (a helper part starts at zero)
                    if (a.identity_init)
                        v.x = ((row_of[i] & 0xFFFF) == col) ? 1.0 : 0.0;
                    else
                        v = a.init[(size_t)traj * a.init_stride + (row_of[i] & 0xFFFF)];
                }
                u[i * CPL + c] = v;
            }
    }

    // Dense output (Tsit5 free interpolant) for save points [k_lo, k_hi) of the step [t, t+dt].
    static __device__ __forceinline__ void save_points(const LaunchArgs &a, const Geo &g, long long traj, const double *ts, int k_lo,
                                                       int k_hi, double t, double dt, double tnew, int col0,
                                                       const int (&row_of)[RPL], const bool (&row_ok)[RPL],
                                                       const Vec &u, const Vec &k1, const Vec &k2, const Vec &k3,
                                                       const Vec &k4, const Vec &k5, const Vec &k6, const Vec &k7,
                                                       const Vec &unew) {
        for (int ks = k_lo; ks < k_hi; ks++) {
            const double tsv = ts[ks];
            if (tsv == tnew) {
                write_out(a, g, traj, ks, col0, row_of, row_ok, unew);
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
                Vec o;
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
                write_out(a, g, traj, ks, col0, row_of, row_ok, o);
            }
        }
    }

    // Lean kernels: stage vector k_{v+2} of a step that contains save points, kept in lane-private shared-memory
    // slots (512 bytes per entry: conflict-free) until the step is accepted.
    static __device__ __forceinline__ void park(const Geo &g, int v, const Vec &k) {
#pragma unroll
        for (int e = 0; e < E; e++) sts2(g.sm_park + 512u * (uint32_t)(v * E + e), k[e]);
    }
    // Dense output of a lean step: k2..k6 come back from the park area, in the order of save_points' sums.
    static __device__ __forceinline__ void save_points_lean(const LaunchArgs &a, const Geo &g, long long traj, const double *ts,
                                                            int k_lo, int k_hi, double t, double dt, double tnew, int col0,
                                                            const int (&row_of)[RPL], const bool (&row_ok)[RPL],
                                                            const Vec &u, const Vec &k1, const Vec &k7, const Vec &unew) {
        for (int ks = k_lo; ks < k_hi; ks++) {
            const double tsv = ts[ks];
            if (tsv == tnew) {
                write_out(a, g, traj, ks, col0, row_of, row_ok, unew);
            } else {
                const double th = (tsv - t) / dt;
                const double th2 = th * th;
                double bb[7];
                bb[0] = th * (1.0 + th * (-2.763706197274826 + th * (2.9132554618219126 + th * -1.0530884977290216)));
                bb[1] = th2 * (0.13169999999999998 + th * (-0.2234 + th * 0.1017));
                bb[2] = th2 * (3.9302962368947516 + th * (-5.941033872131505 + th * 2.490627285651253));
                bb[3] = th2 * (-12.411077166933676 + th * (30.33818863028232 + th * -16.548102889244902));
                bb[4] = th2 * (37.50931341651104 + th * (-88.1789048947664 + th * 47.37952196281928));
                bb[5] = th2 * (-27.896526289197286 + th * (65.09189467479366 + th * -34.87065786149661));
                bb[6] = th2 * (1.5 + th * (-4.0 + th * 2.5));
                Vec o;
#pragma unroll
                for (int e = 0; e < E; e++) o[e] = make_double2(k1[e].x * bb[0], k1[e].y * bb[0]);
#pragma unroll
                for (int v = 0; v < 5; v++)
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double2 kk = lds2(g.sm_park + 512u * (uint32_t)(v * E + e));
                        o[e].x += kk.x * bb[1 + v];
                        o[e].y += kk.y * bb[1 + v];
                    }
#pragma unroll
                for (int e = 0; e < E; e++) {
                    o[e].x += k7[e].x * bb[6];
                    o[e].y += k7[e].y * bb[6];
                    o[e] = make_double2(u[e].x + dt * o[e].x, u[e].y + dt * o[e].y);
                }
                write_out(a, g, traj, ks, col0, row_of, row_ok, o);
            }
        }
    }

    static __device__ __forceinline__ void load_batch(const double2 *buf, int n, int col0, const int (&row_of)[RPL],
                                                      const bool (&row_ok)[RPL], Vec &x) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
#pragma unroll
            for (int c = 0; c < CPL; c++)
                x[i * CPL + c] = row_ok[i] ? buf[(size_t)(col0 + c) * n + row_of[i]] : make_double2(0.0, 0.0);
    }
    static __device__ __forceinline__ void store_batch(double2 *buf, int n, int col0, const int (&row_of)[RPL],
                                                       const bool (&row_ok)[RPL], const Vec &x) {
#pragma unroll
        for (int i = 0; i < RPL; i++)
            if (row_ok[i]) {
#pragma unroll
                for (int c = 0; c < CPL; c++) buf[(size_t)(col0 + c) * n + row_of[i]] = x[i * CPL + c];
            }
    }

    // ---- one trajectory -----------------------------------------------------------------
    static __device__ __forceinline__ void run(const LaunchArgs &a, const Geo &g, const double *smem_gen,
                                               const uint32_t (&ell)[RPL][NZ], const uint32_t (&my_off)[RPL],
                                               const int (&row_of)[RPL], const bool (&row_ok)[RPL],
                                               const int (&sidx)[RPL], double2 *sc_base, long long traj) {
        // sidx: this lane's row indices inside the scratch buffers (row-split teams keep the scratch in their
        // sorted row order so that a warp's accesses are contiguous; everyone else uses the state row)
        const int n = a.g.n;
        const long long tin = a.in_first + traj * a.in_step, tout = a.out_first + traj * a.out_step;
        const double *ts = a.ts + (size_t)tin * a.ts_stride;
        const double *prow = a.params + (size_t)tin * a.n_params;
        const double t0 = ts[0], tend = ts[a.n_ts - 1];
        const double ntot = (double)a.n_out * (double)a.out_cols;
        const uint32_t V0 = g.sm_V, V1 = V0 + g.v_stride, V2 = V1 + g.v_stride, V3 = V2 + g.v_stride, V4 = V3 + g.v_stride;
        const uint32_t yb0 = g.sm_y0, yb1 = g.sm_y1;
        const size_t sbuf = (size_t)a.n_batches * g.cpb * n;
        double2 *const sc_u0 = sc_base, *const sc_k0 = sc_base + sbuf;
        // Streamed row-split teams: column b's u and k1 are fetched by bulk copies into one of two shared-memory
        // buffers while column b - 1 is being integrated (issued by one thread, completion on the buffer's mbarrier).
        const uint32_t nb16 = 16u * (uint32_t)n;
        auto pf_issue = [&](int b, const double2 *su, const double2 *sk) {
            const uint32_t buf = g.sm_pf + (uint32_t)(b & 1) * 2u * nb16, mb = g.sm_mbar + 8u * (uint32_t)(1 + (b & 1));
            mbar_expect_tx(mb, 2u * nb16);
            bulk_g2s(buf, su + (size_t)b * n, nb16, mb);
            bulk_g2s(buf + nb16, sk + (size_t)b * n, nb16, mb);
        };
        // u and k1 of batch b (RS && prefetch: from the shared-memory buffer; otherwise straight from the scratch)
        auto fetch_uk = [&](int b, int col0, const double2 *su, const double2 *sk, Vec &uu, Vec &kk) {
            if (RS && g.sm_pf) {
                if (g.team_thread == 0 && b + 1 < a.n_batches) pf_issue(b + 1, su, sk);
                const uint32_t buf = g.sm_pf + (uint32_t)(b & 1) * 2u * nb16;
                mbar_wait(g.sm_mbar + 8u * (uint32_t)(1 + (b & 1)), (g.ph_pf >> (b & 1)) & 1u);
                g.ph_pf ^= 1u << (b & 1);
#pragma unroll
                for (int i = 0; i < RPL; i++) {
                    uu[i] = row_ok[i] ? lds2(buf + 16u * (uint32_t)sidx[i]) : make_double2(0.0, 0.0);
                    kk[i] = row_ok[i] ? lds2(buf + nb16 + 16u * (uint32_t)sidx[i]) : make_double2(0.0, 0.0);
                }
            } else {
                load_batch(su, n, col0, sidx, row_ok, uu);
                load_batch(sk, n, col0, sidx, row_ok, kk);
            }
        };

        // Component passes of the row-split kernels: batch b brings its own rows (one record per lane and row-slot,
        // LaunchArgs::rs_pass).  The per-lane tables are the kernel's own local arrays (never const objects), rewritten here.
        auto load_pass = [&](int b) {
            if (!(RS && MAXNNZ > 2) || !a.rs_pass) return;
#pragma unroll
            for (int i = 0; i < RPL; i++) {
                const int chunk = i * a.team_warps + g.warp_in_team;
                const int4 d = __ldg(a.rs_pass + ((size_t)b * (size_t)(RPL * a.team_warps) + chunk) * 32 + g.lane);
                const_cast<bool &>(row_ok[i]) = d.x >= 0;
                const_cast<int &>(row_of[i]) = d.x >= 0 ? d.x : 0;
                const_cast<int &>(sidx[i]) = chunk * 32 + g.lane;
                const_cast<uint32_t &>(my_off[i]) = (uint32_t)d.y;
                const_cast<uint32_t &>(ell[i][0]) = (uint32_t)d.z;
                const_cast<uint32_t &>(ell[i][MAXNNZ > 1 ? 1 : 0]) = (uint32_t)d.w & 0xFFu;
                const_cast<uint32_t &>(ell[i][MAXNNZ > 2 ? 2 : 0]) = (uint32_t)d.w >> 8;
                // the next pass's record (pass 0 of the next stage loop after the last one) on its way into L1: a pass
                // otherwise begins with an exposed L2 round trip on every lane
                const int bn = b + 1 < a.n_batches ? b + 1 : 0;
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.rs_pass + ((size_t)bn * (size_t)(RPL * a.team_warps) + chunk) * 32 + g.lane));
            }
        };

        qsim_traj_stats st;
        st.retcode = QSIM_RET_SUCCESS;
        st.n_accept = st.n_reject = 0;
        st.n_rhs = 0;
        st.dt_init = 0.0;

        // resolve the drive descriptors against this trajectory's parameter row; constant channel
        {
            ResolvedEval *rev = const_cast<ResolvedEval *>(
                reinterpret_cast<const ResolvedEval *>(smem_gen + (g.sm_rev - g.sm_tab) / 8));
            for (int e = g.team_thread; e < a.g.n_evals; e += g.team_threads) {
                ResolvedEval r;
                resolve_eval(a.g.evals[e], prow, a.g.tables, r);
                rev[e] = r;
            }
            for (int s = g.team_thread; s < 5; s += g.team_threads) sts1(g.sm_chan + 8u * (s * a.ncp), 1.0);
        }
        team_sync(a.team_warps, g.bar_id);
        compute_tables<5>(a, g, smem_gen, t0, 0.0, true, 0);
        double vstat[NVS];
        if (BLK) {
            // block values (per-trajectory constants): slot ids from the host-built table of this lane's block row
            const uint16_t *bs = a.blk_slot + (size_t)g.blk_row * NVS;
#pragma unroll
            for (int q = 0; q < NVS; q++) vstat[q] = lds1(g.sm_V + VB * (uint32_t)__ldg(bs + q));
        } else {
#pragma unroll
            for (int i = 0; i < RPL; i++)
#pragma unroll
                for (int j = 0; j < NZ; j++)
                    vstat[i * NZ + j] = (SV && j > 0 && j < cap(i)) ? lds1(g.sm_V + (ell[i][j] >> 16)) : 0.0;
        }

        Vec u, k1, k2, k3, k4, k5, k6, k7, tmp;
        int cur = 0;
        int nsave = 0;
        while (nsave < a.n_ts && ts[nsave] <= t0) nsave++;  // save_start: every ts[k] == t0
        // component pruning: the structural zeros of a save point are written when the save point is (spread over the
        // trajectory: a burst of them at the start of every trajectory backs up the write path of page-locked results)
        if (a.n_zero > 0) fill_zeros(a, g, tout, 0, nsave, 0.0);
        const double inf = __longlong_as_double(0x7ff0000000000000LL);
        double ts_next = nsave < a.n_ts ? ts[nsave] : inf;

        double t = t0, dt = 0.0;
        if (tend > t0) {
            const double dtmax = a.dtmax > 0.0 ? a.dtmax : tend - t0;
            // ---- initial step (DiffEqBase ode_determine_initdt) ----
            double acc2[2] = {0.0, 0.0};
            for (int b = RS ? 0 : g.warp_in_team; b < a.n_batches; b += RS ? 1 : a.team_warps) {
                load_pass(b);
                const int col0 = b * g.cpb + g.colofs;
                initial_state(a, g, tin, col0, row_of, row_ok, u);
                for (int ks = 0; ks < nsave; ks++) write_out(a, g, tout, ks, col0, row_of, row_ok, u);
                put_stage(a, g, u, yb0, my_off, row_ok);
                apply(a, g, V0, yb0, ell, vstat, row_of, u, k1);
                if (a.streamed) {
                    store_batch(sc_u0, n, col0, sidx, row_ok, u);
                    store_batch(sc_k0, n, col0, sidx, row_ok, k1);
                }
#pragma unroll
                for (int e = 0; e < E; e++) {
                    // (additively split rows: norms are taken over the physical entries; a helper part counts nothing)
                    const double2 up = AS ? make_double2(as_sum(g, u[e].x), as_sum(g, u[e].y)) : u[e];
                    const double2 kp = AS ? make_double2(as_sum(g, k1[e].x), as_sum(g, k1[e].y)) : k1[e];
                    const double sk = a.abstol + sqrt(up.x * up.x + up.y * up.y) * a.reltol;
                    const double ux = up.x / sk, uy = up.y / sk, fx = kp.x / sk, fy = kp.y / sk;
                    if (!(AS && g.split_helper)) {
                        acc2[0] += ux * ux + uy * uy;
                        acc2[1] += fx * fx + fy * fy;
                    }
                }
            }
            if (RS && g.sm_pf) fence_async_proxy();  // scratch stores above -> bulk-copy reads below
            team_sum<2>(acc2, g.sm_red, a.team_warps, g.warp_in_team, g.lane, g.bar_id);
            const double d0 = sqrt(acc2[0] / ntot), d1 = sqrt(acc2[1] / ntot);
            double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100.0;
            dt0 = fmin(dt0, dtmax);
            compute_tables<1>(a, g, smem_gen, t0 + dt0, 0.0, false, 1);  // table 1 <- t0 + dt0 (table 0 keeps t0)
            double acc1[1] = {0.0};
            if (RS && g.sm_pf && a.streamed) {
                if (a.team_warps == 1) __syncwarp();
                if (g.team_thread == 0) pf_issue(0, sc_u0, sc_k0);
            }
            for (int b = RS ? 0 : g.warp_in_team; b < a.n_batches; b += RS ? 1 : a.team_warps) {
                load_pass(b);
                const int col0 = b * g.cpb + g.colofs;
                if (a.streamed) fetch_uk(b, col0, sc_u0, sc_k0, u, k1);
#pragma unroll
                for (int e = 0; e < E; e++) tmp[e] = make_double2(u[e].x + dt0 * k1[e].x, u[e].y + dt0 * k1[e].y);
                put_stage(a, g, tmp, yb1, my_off, row_ok);
                apply(a, g, V1, yb1, ell, vstat, row_of, tmp, k2);
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double2 up = AS ? make_double2(as_sum(g, u[e].x), as_sum(g, u[e].y)) : u[e];
                    const double dx = k2[e].x - k1[e].x, dy = k2[e].y - k1[e].y;
                    const double2 dp = AS ? make_double2(as_sum(g, dx), as_sum(g, dy)) : make_double2(dx, dy);
                    const double sk = a.abstol + sqrt(up.x * up.x + up.y * up.y) * a.reltol;
                    const double fx = dp.x / sk, fy = dp.y / sk;
                    if (!(AS && g.split_helper)) acc1[0] += fx * fx + fy * fy;
                }
            }
            team_sum<1>(acc1, g.sm_red, a.team_warps, g.warp_in_team, g.lane, g.bar_id);
            const double d2 = sqrt(acc1[0] / ntot) / dt0;
            const double md = fmax(d1, d2);
            const double dt1 = md <= 1e-15 ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(md)) / 5.0);
            dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
            st.dt_init = dt;

            // ---- main loop ----
            // PI controller state: qold^beta2 is carried over (stepsize_controller! / step_accept_controller! /
            // step_reject_controller!: q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin], dt_new = dt/q)
            // Scalars that are only needed at the head and tail of a step live in the warp's control block in
            // shared memory and are fetched in two batches per step; keeping them in registers across the stage
            // code costs spills whose reloads each expose a full memory latency.
            enum { C_TEND, C_DTMAX, C_INVN, C_QIPOW, C_QI2, C_TSNEXT, C_WINT, C_WINW, C_WINNEXT, C_QOLDPOW, C_TNEW, C_WINSC };
#define CTL_LD(i) lds1(g.sm_ctl + 8 * (i))
#define CTL_ST(i, v) sts1(g.sm_ctl + 8 * (i), (v))
            CTL_ST(C_TEND, tend);
            CTL_ST(C_DTMAX, dtmax);
            CTL_ST(C_INVN, 1.0 / ntot);
            CTL_ST(C_QIPOW, pow(a.qoldinit, a.beta2));
            CTL_ST(C_QI2, a.qoldinit * a.qoldinit);
            CTL_ST(C_TSNEXT, ts_next);
            CTL_ST(C_WINT, 0.0);
            CTL_ST(C_WINW, 1.0);
            CTL_ST(C_WINNEXT, 0.0);
            CTL_ST(C_WINSC, 2.0);
            CTL_ST(C_QOLDPOW, pow(a.qoldinit, a.beta2));
            __syncwarp();
            long long iters = 0;
            int n_attempt = 0, red_par = 0;
            int win_state = a.use_windows ? 0 : -1;  // -1 direct evaluation, 0 no window yet, 1 window valid
#ifdef QSIM_TIMING
            long long tk_[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tl_ = clock64();
#endif
            for (;;) {
                QSIM_TICK(0)
                const double c_tend = CTL_LD(C_TEND), c_dtmax = CTL_LD(C_DTMAX), c_tsnext = CTL_LD(C_TSNEXT);
                double win_t = CTL_LD(C_WINT), win_w = CTL_LD(C_WINW), win_sc = CTL_LD(C_WINSC);
                if (!(t < c_tend)) break;
                if (++iters > a.maxiters) { st.retcode = QSIM_RET_MAXITERS; break; }
                if (!(dt > 0.0)) { st.retcode = dt == dt ? QSIM_RET_DT_UNDERFLOW : QSIM_RET_NONFINITE; break; }
                dt = pmin(pmin(dt, c_dtmax), c_tend - t);  // all three positive here
                if (!(dt > fabs(t) * 2.220446049250313e-16)) { st.retcode = QSIM_RET_DT_UNDERFLOW; break; }
                if (win_state >= 0 && (win_state == 0 || t + dt > win_t + win_w)) {
                    // (re)build the drive window starting at this step; shrink until the check passes
                    double wt = fmax(win_state == 0 ? 6.0 * dt : CTL_LD(C_WINNEXT), 1.25 * dt);
                    for (;;) {
                        const double werr = build_window(a, g, smem_gen, t, wt);
                        if (werr <= QSIM_WIN_TOL) {
                            win_t = t;
                            win_w = wt;
                            win_state = 1;
                            CTL_ST(C_WINT, win_t);
                            CTL_ST(C_WINW, win_w);
                            win_sc = 2.0 / win_w;
                            CTL_ST(C_WINSC, win_sc);
                            CTL_ST(C_WINNEXT, werr <= QSIM_WIN_TOL / 64.0 ? fmin(1.5 * wt, 128.0 * dt) : wt);
                            break;
                        }
                        if (wt <= 1.3 * dt) {  // not even one step fits: evaluate this trajectory's drives directly
                            win_state = -1;
                            break;
                        }
                        wt = fmax(0.5 * wt, 1.25 * dt);
                    }
                }
                QSIM_TICK(1)
                if (a.fast_tables && win_state == 1)
                    fast_tables(a, g, t, dt, win_t, win_sc);
                else
                    compute_tables<5>(a, g, smem_gen, t, dt, false, 0, win_state == 1, win_t, win_w);
                QSIM_TICK(2)
                QSIM_LAUNDER_ELL();
                int nsave_hi = nsave;
                {
                    double tnew = t + dt;
                    if (c_tend - tnew < 1e-9 * fabs(c_tend)) {
                        // 10*eps(max(t, tstop)): snap the last step onto the end point
                        const double mx = fmax(fabs(t), fabs(c_tend));
                        const double ulp = __longlong_as_double(__double_as_longlong(mx) + 1) - mx;
                        if (fabs(tnew - c_tend) < 10.0 * ulp) tnew = c_tend;
                    }
                    if (c_tsnext <= tnew) {
                        while (nsave_hi < a.n_ts && ts[nsave_hi] <= tnew) nsave_hi++;
                    }
                    CTL_ST(C_TNEW, tnew);
                }

                double err[1] = {0.0};
                QSIM_TICK(3)
                if (RS && g.sm_pf && a.streamed) {
                    if (a.team_warps == 1) __syncwarp();  // (larger teams passed the error-norm barrier)
                    if (g.team_thread == 0) pf_issue(0, sc_u0 + (size_t)(2 * cur) * sbuf, sc_k0 + (size_t)(2 * cur) * sbuf);
                }
                for (int b = RS ? 0 : g.warp_in_team; b < a.n_batches; b += RS ? 1 : a.team_warps) {
                    load_pass(b);
                    const int col0 = b * g.cpb + g.colofs;
                    if (a.streamed)
                        fetch_uk(b, col0, sc_u0 + (size_t)(2 * cur) * sbuf, sc_k0 + (size_t)(2 * cur) * sbuf, u, k1);
                    Vec part;  // partial combination of the next stage, formed while the current stage's operands load
                    double ex[E], ey[E];
                    if (LEAN) {
                        const bool saving = nsave_hi > nsave;   // warp-uniform: this step holds dense-output points
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = A21 * k1[e].x, ay = A21 * k1[e].y;
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                        }
                        put_stage(a, g, tmp, yb0, my_off, row_ok);
                        apply<QSIM_LEAN_CH_EARLY>(a, g, V0, yb0, ell, vstat, row_of, tmp, k2, [&] {
#pragma unroll
                            for (int e = 0; e < E; e++) part[e] = make_double2(A31 * k1[e].x, A31 * k1[e].y);
                        });
                        if (saving) park(g, 0, k2);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = fma(A32, k2[e].x, part[e].x), ay = fma(A32, k2[e].y, part[e].y);
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                        }
                        put_stage(a, g, tmp, yb1, my_off, row_ok);
                        apply<QSIM_LEAN_CH_EARLY>(a, g, V1, yb1, ell, vstat, row_of, tmp, k3, [&] {
#pragma unroll
                            for (int e = 0; e < E; e++)
                                part[e] = make_double2(fma(A42, k2[e].x, A41 * k1[e].x), fma(A42, k2[e].y, A41 * k1[e].y));
                        });
                        if (saving) park(g, 1, k3);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = fma(A43, k3[e].x, part[e].x), ay = fma(A43, k3[e].y, part[e].y);
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                        }
                        put_stage(a, g, tmp, yb0, my_off, row_ok);
                        apply<QSIM_LEAN_CH_EARLY>(a, g, V2, yb0, ell, vstat, row_of, tmp, k4, [&] {
#pragma unroll
                            for (int e = 0; e < E; e++)
                                part[e] = make_double2(fma(A53, k3[e].x, fma(A52, k2[e].x, A51 * k1[e].x)),
                                                       fma(A53, k3[e].y, fma(A52, k2[e].y, A51 * k1[e].y)));
                        });
                        if (saving) park(g, 2, k4);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = fma(A54, k4[e].x, part[e].x), ay = fma(A54, k4[e].y, part[e].y);
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                        }
                        put_stage(a, g, tmp, yb1, my_off, row_ok);
                        apply<QSIM_LEAN_CH_MID>(a, g, V3, yb1, ell, vstat, row_of, tmp, k5, [&] {
#pragma unroll
                            for (int e = 0; e < E; e++)
                                part[e] = make_double2(
                                    fma(A64, k4[e].x, fma(A63, k3[e].x, fma(A62, k2[e].x, A61 * k1[e].x))),
                                    fma(A64, k4[e].y, fma(A63, k3[e].y, fma(A62, k2[e].y, A61 * k1[e].y))));
                        });
                        if (saving) park(g, 3, k5);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = fma(A65, k5[e].x, part[e].x), ay = fma(A65, k5[e].y, part[e].y);
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                        }
                        put_stage(a, g, tmp, yb0, my_off, row_ok);
                        // stage 6; under its loads: the k1..k5 part of u_new's combination
                        apply(a, g, V4, yb0, ell, vstat, row_of, tmp, k6, [&] {
#pragma unroll
                            for (int e = 0; e < E; e++)
                                part[e] = make_double2(
                                    fma(A75, k5[e].x, fma(A74, k4[e].x, fma(A73, k3[e].x, fma(A72, k2[e].x, A71 * k1[e].x)))),
                                    fma(A75, k5[e].y, fma(A74, k4[e].y, fma(A73, k3[e].y, fma(A72, k2[e].y, A71 * k1[e].y)))));
                        });
                        if (saving) park(g, 4, k6);
                        // u_new, then the k1..k6 part of the error combination: k2..k6 are dead before stage 7's loads
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double ax = fma(A76, k6[e].x, part[e].x), ay = fma(A76, k6[e].y, part[e].y);
                            tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);  // u_new
                        }
                        put_stage(a, g, tmp, yb1, my_off, row_ok);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            ex[e] = fma(BT6, k6[e].x, fma(BT5, k5[e].x, fma(BT4, k4[e].x, fma(BT3, k3[e].x, fma(BT2, k2[e].x, BT1 * k1[e].x)))));
                            ey[e] = fma(BT6, k6[e].y, fma(BT5, k5[e].y, fma(BT4, k4[e].y, fma(BT3, k3[e].y, fma(BT2, k2[e].y, BT1 * k1[e].y)))));
                        }
                        apply(a, g, V4, yb1, ell, vstat, row_of, tmp, k7);
                    } else {
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = A21 * k1[e].x, ay = A21 * k1[e].y;
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(a, g, tmp, yb0, my_off, row_ok);
                    QSIM_TICK(4)
                    apply(a, g, V0, yb0, ell, vstat, row_of, tmp, k2, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++) part[e] = make_double2(A31 * k1[e].x, A31 * k1[e].y);
                    });
                    QSIM_TICK(5)
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = fma(A32, k2[e].x, part[e].x), ay = fma(A32, k2[e].y, part[e].y);
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(a, g, tmp, yb1, my_off, row_ok);
                    apply(a, g, V1, yb1, ell, vstat, row_of, tmp, k3, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++)
                            part[e] = make_double2(fma(A42, k2[e].x, A41 * k1[e].x), fma(A42, k2[e].y, A41 * k1[e].y));
                    });
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = fma(A43, k3[e].x, part[e].x), ay = fma(A43, k3[e].y, part[e].y);
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(a, g, tmp, yb0, my_off, row_ok);
                    apply(a, g, V2, yb0, ell, vstat, row_of, tmp, k4, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++)
                            part[e] = make_double2(fma(A53, k3[e].x, fma(A52, k2[e].x, A51 * k1[e].x)),
                                                   fma(A53, k3[e].y, fma(A52, k2[e].y, A51 * k1[e].y)));
                    });
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = fma(A54, k4[e].x, part[e].x), ay = fma(A54, k4[e].y, part[e].y);
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(a, g, tmp, yb1, my_off, row_ok);
                    apply(a, g, V3, yb1, ell, vstat, row_of, tmp, k5, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++)
                            part[e] = make_double2(
                                fma(A64, k4[e].x, fma(A63, k3[e].x, fma(A62, k2[e].x, A61 * k1[e].x))),
                                fma(A64, k4[e].y, fma(A63, k3[e].y, fma(A62, k2[e].y, A61 * k1[e].y))));
                    });
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = fma(A65, k5[e].x, part[e].x), ay = fma(A65, k5[e].y, part[e].y);
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);
                    }
                    put_stage(a, g, tmp, yb0, my_off, row_ok);
                    apply(a, g, V4, yb0, ell, vstat, row_of, tmp, k6, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++)
                            part[e] = make_double2(
                                fma(A75, k5[e].x, fma(A74, k4[e].x, fma(A73, k3[e].x, fma(A72, k2[e].x, A71 * k1[e].x)))),
                                fma(A75, k5[e].y, fma(A74, k4[e].y, fma(A73, k3[e].y, fma(A72, k2[e].y, A71 * k1[e].y)))));
                    });
#pragma unroll
                    for (int e = 0; e < E; e++) {
                        const double ax = fma(A76, k6[e].x, part[e].x), ay = fma(A76, k6[e].y, part[e].y);
                        tmp[e] = make_double2(u[e].x + dt * ax, u[e].y + dt * ay);  // u_new
                    }
                    put_stage(a, g, tmp, yb1, my_off, row_ok);
                    // while stage 7's operands load: the k1..k6 part of the error combination
                    apply(a, g, V4, yb1, ell, vstat, row_of, tmp, k7, [&] {
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            ex[e] = fma(BT6, k6[e].x, fma(BT5, k5[e].x, fma(BT4, k4[e].x, fma(BT3, k3[e].x, fma(BT2, k2[e].x, BT1 * k1[e].x)))));
                            ey[e] = fma(BT6, k6[e].y, fma(BT5, k5[e].y, fma(BT4, k4[e].y, fma(BT3, k3[e].y, fma(BT2, k2[e].y, BT1 * k1[e].y)))));
                        }
                    });
                    }
                    QSIM_TICK(6)
                    {
                        // Error norm: sum over this lane's entries of |u~|^2 / (abstol + max(|u|,|u_new|) reltol)^2.
                        // Written operation by operation across the entries so the independent per-entry
                        // chains (combination, reciprocal square root, reciprocal) overlap instead of running
                        // back to back.
                        double m2[E], ys[E], hx[E], dn[E], yr[E];
#pragma unroll
                        for (int e = 0; e < E; e++) { ex[e] = fma(BT7, k7[e].x, ex[e]); ey[e] = fma(BT7, k7[e].y, ey[e]); }
                        // additively split rows: error estimate and scale of the PHYSICAL entry (owner + helper part);
                        // the helper part itself counts nothing
                        double2 uc[E], tc[E];
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            uc[e] = u[e];
                            tc[e] = tmp[e];
                            if (AS) {
                                uc[e] = make_double2(as_sum(g, u[e].x), as_sum(g, u[e].y));
                                tc[e] = make_double2(as_sum(g, tmp[e].x), as_sum(g, tmp[e].y));
                                const double sx = as_sum(g, ex[e]), sy = as_sum(g, ey[e]);   // (every lane shuffles)
                                ex[e] = g.split_helper ? 0.0 : sx;
                                ey[e] = g.split_helper ? 0.0 : sy;
                            }
                        }
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double au = fma(uc[e].x, uc[e].x, uc[e].y * uc[e].y);
                            const double an = fma(tc[e].x, tc[e].x, tc[e].y * tc[e].y);
                            // max of non-negative doubles through their bit patterns (entries below 1e-140 count as that)
                            const long long bu = __double_as_longlong(au), bn = __double_as_longlong(an);
                            const long long bm = bu > bn ? bu : bn;
                            m2[e] = __longlong_as_double(bm > 0x05d0000000000000LL ? bm : 0x05d0000000000000LL);
                        }
#pragma unroll
                        for (int e = 0; e < E; e++) asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(ys[e]) : "d"(m2[e]));
#pragma unroll
                        for (int e = 0; e < E; e++) hx[e] = 0.5 * m2[e];
#pragma unroll
                        for (int e = 0; e < E; e++) ys[e] = ys[e] * fma(-hx[e] * ys[e], ys[e], 1.5);
#pragma unroll
                        for (int e = 0; e < E; e++) ys[e] = ys[e] * fma(-hx[e] * ys[e], ys[e], 1.5);
#pragma unroll
                        for (int e = 0; e < E; e++) dn[e] = fma(m2[e] * ys[e], a.reltol, a.abstol);  // abstol + |.| reltol
#pragma unroll
                        for (int e = 0; e < E; e++) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yr[e]) : "d"(dn[e]));
#pragma unroll
                        for (int e = 0; e < E; e++) yr[e] = fma(yr[e], fma(-dn[e], yr[e], 1.0), yr[e]);
#pragma unroll
                        for (int e = 0; e < E; e++) yr[e] = fma(yr[e], fma(-dn[e], yr[e], 1.0), yr[e]);
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            const double rs = dt * yr[e];
                            ex[e] *= rs;
                            ey[e] *= rs;
                        }
#pragma unroll
                        for (int e = 0; e < E; e++) err[0] += fma(ex[e], ex[e], ey[e] * ey[e]);
                    }
                    if (a.streamed) {
                        // tentative: the next buffers are adopted only if the step is accepted; dense-output
                        // slots are rewritten by the retried / a later step if it is not
                        store_batch(sc_u0 + (size_t)(2 * (cur ^ 1)) * sbuf, n, col0, sidx, row_ok, tmp);
                        store_batch(sc_k0 + (size_t)(2 * (cur ^ 1)) * sbuf, n, col0, sidx, row_ok, k7);
                        if (nsave_hi > nsave) {
                            if (LEAN)
                                save_points_lean(a, g, tout, ts, nsave, nsave_hi, t, dt, CTL_LD(C_TNEW), col0, row_of, row_ok,
                                                 u, k1, k7, tmp);
                            else
                                save_points(a, g, tout, ts, nsave, nsave_hi, t, dt, CTL_LD(C_TNEW), col0, row_of, row_ok, u, k1,
                                            k2, k3, k4, k5, k6, k7, tmp);
                        }
                    }
                }
                n_attempt++;
                QSIM_TICK(7)
                if (RS && g.sm_pf && a.streamed) fence_async_proxy();  // this step's scratch stores -> next step's bulk reads
                err[0] = team_sum_alt(err[0], g.sm_red, a.team_warps, g.warp_in_team, g.lane, g.bar_id, red_par);
                QSIM_TICK(8)
                // EEst^2 = mean |r_i|^2.  PI controller (stepsize_controller! / step_accept_controller! /
                // step_reject_controller!): q = EEst^beta1 / qold^beta2 / gamma clamped to [1/qmax, 1/qmin];
                // evaluated as dt * r with r = 1/q = gamma exp(beta2 log qold - beta1 log EEst).
                const double inv_ntot = CTL_LD(C_INVN), qold_pow = CTL_LD(C_QOLDPOW), qinit2 = CTL_LD(C_QI2),
                             qinit_pow = CTL_LD(C_QIPOW), tnew = CTL_LD(C_TNEW);
                const double e2 = err[0] * inv_ntot;
                if (!isfinite(e2)) { st.retcode = QSIM_RET_NONFINITE; break; }
                // ra = EEst^-beta1, rb = EEst^beta2 (EEst = sqrt(e2)) from the power tables; libm outside their range
                double ra, rb;
                if (!pow_pair(a, e2, ra, rb)) {
                    if (e2 == 0.0) {
                        ra = 1e300;   // "q = 1/qmax": largest growth after the clamp
                        rb = 0.0;
                    } else {
                        const double le = 0.5 * log(e2);
                        ra = exp(-a.beta1 * le);
                        rb = exp(a.beta2 * le);
                    }
                }
                if (e2 <= 1.0) {
                    st.n_accept++;
                    const double r = pmin(a.qmax, pmax(a.qmin, a.gamma * ra * qold_pow));
                    CTL_ST(C_QOLDPOW, e2 >= qinit2 ? rb : qinit_pow);  // qold = max(EEst, qoldinit)
                    if (!a.streamed) {
                        if (nsave_hi > nsave && (RS || g.warp_in_team < a.n_batches)) {
                            if (LEAN)
                                save_points_lean(a, g, tout, ts, nsave, nsave_hi, t, dt, tnew,
                                                 g.warp_in_team * g.cpb + g.colofs, row_of, row_ok, u, k1, k7, tmp);
                            else
                                save_points(a, g, tout, ts, nsave, nsave_hi, t, dt, tnew,
                                            (RS ? 0 : g.warp_in_team) * g.cpb + g.colofs,
                                            row_of, row_ok, u, k1, k2, k3, k4, k5, k6, k7, tmp);
                        }
#pragma unroll
                        for (int e = 0; e < E; e++) {
                            u[e] = tmp[e];
                            k1[e] = k7[e];
                        }
                    } else {
                        cur ^= 1;
                    }
                    if (nsave_hi > nsave) {
                        if (a.n_zero > 0) fill_zeros(a, g, tout, nsave, nsave_hi, 0.0);
                        nsave = nsave_hi;
                        CTL_ST(C_TSNEXT, nsave < a.n_ts ? ts[nsave] : inf);
                    }
                    t = tnew;
                    dt = dt * r;
                } else {
                    st.n_reject++;
                    // dt / min(1/qmin, q11/gamma), q11 = EEst^beta1
                    dt = dt * pmax(a.qmin, a.gamma * ra);
                }
            }
#ifdef QSIM_TIMING
            if (traj == 0 && g.team_thread == 0)
                printf("cycles/step: head %lld window %lld tables %lld head2 %lld combo+put(st2) %lld apply(st2) %lld "
                       "stages3-7 %lld error %lld reduce %lld controller+accept %lld   steps %d\n",
                       tk_[1] / n_attempt, 0LL, tk_[2] / n_attempt, tk_[3] / n_attempt, tk_[4] / n_attempt,
                       tk_[5] / n_attempt, tk_[6] / n_attempt, tk_[7] / n_attempt, tk_[8] / n_attempt, tk_[0] / n_attempt,
                       n_attempt);
#endif
            st.n_rhs = 3 + 6 * n_attempt;  // the reference also evaluates f(u0,t0) a second time
        } else {
            // zero-length span: only the start saves
            for (int b = RS ? 0 : g.warp_in_team; b < a.n_batches; b += RS ? 1 : a.team_warps) {
                load_pass(b);
                const int col0 = b * g.cpb + g.colofs;
                initial_state(a, g, tin, col0, row_of, row_ok, u);
                for (int ks = 0; ks < nsave; ks++) write_out(a, g, tout, ks, col0, row_of, row_ok, u);
            }
        }
        // unreached save points of a failed trajectory: NaN
        if (nsave < a.n_ts) {
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
            for (int e = 0; e < E; e++) tmp[e] = make_double2(qnan, qnan);
            for (int b = RS ? 0 : g.warp_in_team; b < a.n_batches; b += RS ? 1 : a.team_warps) {
                load_pass(b);
                for (int ks = nsave; ks < a.n_ts; ks++) write_out(a, g, tout, ks, b * g.cpb + g.colofs, row_of, row_ok, tmp);
            }
            if (a.n_zero > 0) fill_zeros(a, g, tout, nsave, a.n_ts, qnan);
        }
        st.t_final = t;
        if (g.team_thread == 0 && a.stats) a.stats[tout] = st;
    }
};

// MAXT: largest block the variant is launched with (128: one-warp teams, 4 per CTA, 2 CTAs = 8 warps per SM,
// up to 255 registers so nothing spills; 256: teams of 2-8 warps, one CTA per SM, 255 registers; 288: row-split
// teams of nine warps (865 > n > 768), 168 registers).
// MINB: minimum resident CTAs per SM the register allocation must allow; IM: purely imaginary generator
// (unitary evolution under a real Hamiltonian): value slots hold Im(A) only, 2 DFMA per multiply-add.
template <int RPL, int CPL, int MAXNNZ, int MAXT, int MINB, int VM, bool RS, bool LEAN = false>
__global__ void __launch_bounds__(MAXT, MINB) solve_kernel(const __grid_constant__ LaunchArgs a) {
    extern __shared__ __align__(16) double smem[];
    typedef Kern<RPL, CPL, MAXNNZ, VM, RS, LEAN> K;
    static_assert(!LEAN || (!RS && VM != 3), "lean kernels: resident row-per-lane teams");
    constexpr bool IM = K::IM;
    const int n = a.g.n;
    Geo g;
    g.lane = threadIdx.x & 31;
    g.team_threads = a.team_warps * 32;
    const int team_in_cta = threadIdx.x / g.team_threads;
    g.team_thread = threadIdx.x - team_in_cta * g.team_threads;
    g.warp_in_team = g.team_thread >> 5;
    g.bar_id = 1 + team_in_cta;
    g.cpb = a.groups * CPL;
    // lane -> (row, column group).  Lanes without a row mimic the addresses of the last active lane so
    // that their (discarded) gathers broadcast instead of adding bank conflicts.
    int g_idx = 0, row0 = g.lane;
    bool lane_active = true;
    g.blk_row = 0;
    if (K::BLK || K::TR) {
        // lane -> (block row, column group): RPL consecutive (virtual) rows of one column
        lane_active = g.lane < a.groups * a.n_blk_rows;
        const int le = lane_active ? g.lane : a.groups * a.n_blk_rows - 1;
        g_idx = le / a.n_blk_rows;
        g.blk_row = le - g_idx * a.n_blk_rows;
    } else if (RS && a.cu_map) {
        g_idx = g.lane;            // column (within the batch of 32); rows come from the table (rs_perm)
    } else if (a.groups > 1) {
        // lanes [0, groups * n): (column group, row); split rows: the next groups * nh lanes are the groups' helpers
        // (table rows n .. n + nh - 1, no state of their own); the rest mimic the last row lane
        const int nl = a.groups * n, nh = a.g.n_vrows - n;
        lane_active = g.lane < nl + a.groups * nh;
        if (g.lane >= nl && lane_active) {
            const int h = g.lane - nl;
            g_idx = h / nh;
            row0 = n + (h - g_idx * nh);
        } else {
            const int le = lane_active ? g.lane : nl - 1;
            g_idx = le / n;
            row0 = le - g_idx * n;
        }
    }
    g.colofs = g_idx * CPL;
    g.split_partner = g.lane;
    g.split_helper = g.split_owner = false;
    if (!RS && a.split_info) {
        const unsigned si = a.split_info[g.lane];
        g.split_partner = (int)(si & 31u);
        g.split_helper = (si & 32u) != 0;
        g.split_owner = (si & 64u) != 0;
    }
    g.stage_c = kStageC[g.lane < 5 ? g.lane : 4];

    // ---- shared memory map (byte addresses) ----
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem);
    g.sm_tab = sbase;
    g.sm_cheb = sbase + 512;                             // double[NN*NN + NN + 3] (1280 bytes reserved)
    g.sm_pow = sbase + 512 + 1280;                       // double[2][16][K] Taylor coefficients, double[2][NE] 2^(pE)
    g.sm_dptr = sbase + a.off_dptr;                      // int[n_dyn_slots + 1]
    g.sm_dchan = g.sm_dptr + 4 * (a.g.n_dyn_slots + 1);  // int[n_dyn_pairs]
    g.sm_dw = sbase + a.off_dw;                          // double2[n_dyn_pairs]
    g.sm_drec = sbase + a.off_drec;                      // 64-byte records of the time-dependent slots
    const uint32_t tbase = sbase + a.common_bytes + team_in_cta * a.team_bytes;
    g.v_stride = K::VSTRIDE * a.nvp;
    g.sm_V = tbase;
    g.sm_chan = g.sm_V + ((5 * g.v_stride + 15u) & ~15u);
    g.sm_rev = g.sm_chan + 8u * 5 * a.ncp;
    g.sm_win = g.sm_rev + (uint32_t)sizeof(ResolvedEval) * (a.g.n_evals > 0 ? a.g.n_evals : 1);
    g.sm_red = g.sm_win + 256u * a.ncp;
    const uint32_t ybytes = 16u * n * a.y_row_stride;
    g.sm_y0 = g.sm_red + 8u * (a.team_warps * 4 + 8) + (RS ? 0 : g.warp_in_team) * 2 * ybytes +
              16u * g_idx * a.y_grp_stride;
    g.sm_y1 = g.sm_y0 + ybytes;
    g.sm_ctl = g.sm_red + 8u * (a.team_warps * 4 + 8) + (RS ? 1 : a.team_warps) * 2 * ybytes + 128u * g.warp_in_team;
    g.sm_park = tbase + a.off_park + (uint32_t)g.warp_in_team * (5u * K::E * 512u) + 16u * g.lane;
    g.sm_ell = (RS && a.ell_in_smem) ? sbase + a.off_ell : 0u;
    g.sm_mbar = tbase + a.off_pf;                                   // three mbarriers (32 bytes reserved)
    g.sm_pf = (RS && a.rs_prefetch) ? g.sm_mbar + 32u : 0u;        // two buffers of 2 * n * 16 bytes
    g.ph_stage = 0u;
    g.ph_pf = 0u;

    // ---- CTA-wide constants: series tables and the dynamic slots' pair lists ----
    {
        const double *src = reinterpret_cast<const double *>(a.g.tab);
        for (int i = threadIdx.x; i < (int)(sizeof(SeriesTables) / 8); i += blockDim.x) sts1(g.sm_tab + 8 * i, src[i]);
        for (int i = threadIdx.x; i < QSIM_NN * QSIM_NN + QSIM_NN + 3; i += blockDim.x) sts1(g.sm_cheb + 8 * i, a.cheb[i]);
        if (a.dyn_fast)
            for (int i = threadIdx.x; i < 8 * a.g.n_dyn_slots; i += blockDim.x) sts1(g.sm_drec + 8 * i, a.dyn_rec[i]);
        int *dptr = reinterpret_cast<int *>(smem) + a.off_dptr / 4;
        for (int i = threadIdx.x; i <= a.g.n_dyn_slots; i += blockDim.x) dptr[i] = a.g.slot_ptr[i];
        for (int i = threadIdx.x; i < a.g.n_dyn_pairs; i += blockDim.x) {
            dptr[a.g.n_dyn_slots + 1 + i] = a.g.pair_chan[i];
            sts2(g.sm_dw + 16 * i, a.g.pair_w[i]);
        }
        if (RS) {
            if (a.ell_in_smem)
                for (int i = threadIdx.x; i < a.ell_bytes / 4; i += blockDim.x)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + a.off_ell + 4u * i), "r"(a.ell_packed[i]) : "memory");
            if (g.team_thread == 0) {
                mbar_init(g.sm_mbar, (uint32_t)a.team_warps);  // stage exchange: one arrival per warp
                mbar_init(g.sm_mbar + 8u, 1u);                 // prefetch buffers: the issuing thread + transaction bytes
                mbar_init(g.sm_mbar + 16u, 1u);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        // the zero slot of every table (padding entries of short rows read it)
        for (int s = g.team_thread; s < 5; s += g.team_threads)
            if (K::MX) {
                sts1(g.sm_V + s * g.v_stride + 8 * a.g.n_slots, 0.0);
                sts1(g.sm_V + s * g.v_stride + 8 * (a.nvp + a.g.n_slots), 0.0);
            } else if (IM)
                sts1(g.sm_V + s * g.v_stride + 8 * a.g.n_slots, 0.0);
            else
                sts2(g.sm_V + s * g.v_stride + 16 * a.g.n_slots, make_double2(0.0, 0.0));
    }
    __syncthreads();

    // ---- per-lane sparse rows, kept in registers for the whole kernel ----
    int row_of[RPL], sidx[RPL];
    bool row_ok[RPL];
    uint32_t my_off[RPL];
    uint32_t ell[RPL][K::NZ];
#pragma unroll
    for (int i = 0; i < RPL; i++) {
        if (RS) {
            // sorted row q of chunk (slot i, this warp); the stage buffer, the scratch and the packed table are
            // indexed by q, initial state and results by the state row rs_perm[q].  Lanes past the last row read
            // the table's zero-slot padding (the host fills the whole chunk).
            // (rs_perm / rs_pos hold 32 entries per chunk; -1 marks a lane without a row.  With the column-per-lane
            // mapping all lanes of a chunk name the same row and the lane selects the column instead.)
            if (MAXNNZ > 2) ell[i][MAXNNZ > 2 ? 2 : 0] = 0u;
            if (a.rs_pass) {   // component passes: every batch loads its own rows (Kern::run, load_pass)
                row_ok[i] = false;
                row_of[i] = sidx[i] = 0;
                my_off[i] = 0u;
                ell[i][0] = 0u;
                ell[i][MAXNNZ > 1 ? 1 : 0] = 1u;
                continue;
            }
            const int chunk = i * a.team_warps + g.warp_in_team;
            const int q = chunk * 32 + g.lane;
            const int rr = __ldg(a.rs_perm + q);
            row_ok[i] = rr >= 0;
            row_of[i] = row_ok[i] ? rr : 0;
            sidx[i] = row_ok[i] ? q : 0;
            my_off[i] = 16u * (uint32_t)(__ldg(a.rs_pos + q) * a.y_row_stride);
            ell[i][0] = 4u * (32u * (uint32_t)__ldg(a.rs_off + chunk) + (uint32_t)g.lane);  // byte offset of entry 0
            ell[i][MAXNNZ > 1 ? 1 : 0] = (uint32_t)__ldg(a.rs_len + chunk);                  // entries per row (0: no rows)
            continue;
        }
        sidx[i] = 0;
        row_of[i] = (K::BLK || K::TR) ? g.blk_row * RPL + i : row0 + 32 * i;   // (tiered: the virtual row, until below)
        row_ok[i] = lane_active && row_of[i] < n;
        // component pruning: warp w's local rows are rows [w * n, w * n + n) of the row table and of cf_vrow
        const int vbase = a.cf_vrow ? g.warp_in_team * n : 0;
        if (a.cf_vrow && row_ok[i]) row_ok[i] = __ldg(a.cf_vrow + vbase + row_of[i]) >= 0;
        sidx[i] = row_of[i];
        const int row_eff = row_of[i] < n ? row_of[i] : n - 1;
        my_off[i] = 16u * ((K::TR ? __ldg(a.row_pos + row_eff) : row_eff) * a.y_row_stride);
        // the row table may hold more rows than the state (helper lanes of split rows): those lanes gather too
        const bool ell_ok = lane_active && row_of[i] < (a.cf_vrow ? n : a.g.n_vrows);
        // (tiered: lanes without a row repeat the last lane's table rows, zero slot: their loads coincide with its loads)
        const int ell_row = vbase + ((ell_ok || (K::TR && row_of[i] < n)) ? row_of[i] : n - 1);
#pragma unroll
        for (int j = 0; j < K::NZ; j++) {
            // rows shorter than the variant's capacity: own row, zero slot
            uint32_t e = (uint32_t)row_eff | (0xFFFFu << 16);
            // (component forms: a local row without a state row is never stored to the stage buffer; the host pointed its
            // diagonal entry at a position that is, and the padding beyond the table's row length follows it)
            if (a.cf_vrow) e = (a.g.ell[ell_row] & 0xFFFFu) | (0xFFFFu << 16);
            if (!RS && j < a.g.max_nnz) e = a.g.ell[(size_t)j * a.g.n_vrows + ell_row];
            // slot 0xFFFF marks a zero entry (padding, or a row without a stored diagonal)
            const uint32_t sl = (!ell_ok || (e >> 16) == 0xFFFFu) ? (uint32_t)a.g.n_slots : (e >> 16);
            ell[i][j] = (16u * (e & 0xFFFFu) * a.y_row_stride) | ((K::VB * sl) << 16);
            // mixed mode: an imaginary off-diagonal entry reads the second half of the split table
            if (K::MX && j > 0 && sl < (uint32_t)a.g.n_slots && __ldg(a.slot_imag + sl))
                ell[i][j] = (ell[i][j] + ((8u * a.nvp) << 16)) | 0x80000000u;
        }
        // tiered: stage buffer and row table are indexed by the virtual row, initial state and results by the state row
        if (K::TR || K::AS) row_of[i] = row_ok[i] ? __ldg(a.row_perm + row_of[i]) : 0;
        // component pruning: state row for the initial state and the results, local row (bits 16..) for the column lookup
        if (a.cf_vrow) row_of[i] = (row_ok[i] ? __ldg(a.cf_vrow + vbase + sidx[i]) : 0) | (sidx[i] << 16);
    }
    double2 *sc_base = nullptr;
    if (a.streamed)
        sc_base = a.scratch + (size_t)(blockIdx.x * a.teams_per_cta + team_in_cta) * 4 * ((size_t)a.n_batches * g.cpb * n);

    const uint32_t wslot = g.sm_red + 8u * (a.team_warps * 4 + 6);
    for (;;) {
        if (g.team_thread == 0) {
            const unsigned long long w = atomicAdd(a.work_counter, 1ULL);
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(wslot), "l"(w) : "memory");
        }
        team_sync(a.team_warps, g.bar_id);
        unsigned long long traj;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(traj) : "r"(wslot));
        team_sync(a.team_warps, g.bar_id);
        if (traj >= (unsigned long long)a.n_traj) break;
        K::run(a, g, smem, ell, my_off, row_of, row_ok, sidx, sc_base, (long long)traj);
    }
}

// Host-side launcher for one (RPL, CPL, MAXNNZ) instantiation.
typedef cudaError_t (*launch_fn)(const LaunchArgs &, int grid, int block, size_t smem, cudaStream_t);

template <int RPL, int CPL, int MAXNNZ, int MAXT, int MINB, int VM, bool RS, bool LEAN = false>
cudaError_t launch_solver(const LaunchArgs &a, int grid, int block, size_t smem, cudaStream_t stream) {
    if (block > MAXT) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(solve_kernel<RPL, CPL, MAXNNZ, MAXT, MINB, VM, RS, LEAN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    solve_kernel<RPL, CPL, MAXNNZ, MAXT, MINB, VM, RS, LEAN><<<grid, block, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int RPL, int CPL, int MAXNNZ, int MAXT, int MINB, int VM, bool RS, bool LEAN = false>
cudaError_t occupancy_solver(int block, size_t smem, int *ctas_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(solve_kernel<RPL, CPL, MAXNNZ, MAXT, MINB, VM, RS, LEAN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm,
                                                         solve_kernel<RPL, CPL, MAXNNZ, MAXT, MINB, VM, RS, LEAN>, block, smem);
}

}  // namespace qsim
