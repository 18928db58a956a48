// Floquet propagation of a whole sweep in one call.
//
// Replaces, batched over the sweep and device-resident, the reference's wrappers
//   floquet_propagator(propagator_func, t_period, rtol)                  time_evolution.jl:188-220
//   floquet_rise_fall_propagator(propagator_func, t_period, rise, fall)  time_evolution.jl:312-338
//   decompose_times_floquet / unique_tol                                 time_evolution.jl:275-293, 355-360
// which use U(n tau + dt) = U(dt) U(tau)^n for a system that is tau-periodic between a rise and a fall segment.
//
// The reference runs three solves per system (rise times, the unique remainders of the periodic times plus one
// period, fall times) and combines them on the host, one system at a time.  Here every segment of every sweep
// point is one trajectory of ONE batched solver launch (per-trajectory save grids, padded to a common length;
// each trajectory may have its own period, e.g. when the drive frequency is a sweep axis), the results stay in
// device memory, and one kernel forms all the powers and products:
//   rise    t <  t0:          U(t)
//   middle  t0 <= t <= t1:    U_rem[idx(t)] * U_tau^q(t) * U(t0)          q non-decreasing along ts
//   fall    t >  t1:          U_fall(t) * U(t1)
// Only the final propagators leave the device.  Works for unitary_propagator (d x d) and me_propagator
// (d^2 x d^2, column-major vec), like the reference, which composes whatever its propagator function returns.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/qsim_b200.h"

// internal hooks of qsim_api.cu
int qsim_internal_fail(int code, const std::string &msg);
int qsim_internal_floquet_buffers(qsim_ctx *c, size_t seg_bytes, size_t work_bytes, size_t idx_bytes, void **seg, void **work,
                                  void **idx, cudaStream_t *stream, int *device);
void qsim_internal_count_launch(qsim_ctx *c);
int qsim_internal_solve_to_device(qsim_ctx *c, qsim_problem *p, int which, int64_t n_traj, const double *params, const double *ts,
                                  int n_ts, int64_t ts_stride, const qsim_solver_opts *opts, void *d_out, qsim_traj_stats *stats);

#define CUF(call)                                                                                                   \
    do {                                                                                                            \
        cudaError_t e__ = (call);                                                                                   \
        if (e__ != cudaSuccess) return qsim_internal_fail(QSIM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

namespace {

// per output (trajectory, k): which segment it comes from, which save of that segment, and the power of U_tau
struct OutRef {
    int32_t kind;   // 0 rise, 1 middle, 2 fall
    int32_t save;   // index into the segment's saves
    int32_t quot;   // middle: number of whole periods
    int32_t pad;
};
// per trajectory: its segments inside the solver's result buffer (-1: no such segment), where U(t0), U_tau and
// the reference point t1 sit
struct TrajRef {
    int32_t rise_seg, mid_seg, fall_seg;
    int32_t rise_t0_save;            // save index of t0 in the rise segment
    int32_t period_save;             // save index of t0 + tau in the middle segment
    int32_t t1_save, t1_quot;        // t1 = remainder save t1_save after t1_quot periods
    int32_t pad;
};

// C = A * B for n x n complex column-major matrices; all threads of the CTA take part (C must not alias A or B)
__device__ __forceinline__ void cta_matmul(int n, const double2 *A, const double2 *B, double2 *C) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int r = e % n, c = e / n;
        double sx = 0.0, sy = 0.0;
        for (int k = 0; k < n; k++) {
            const double2 a = A[(size_t)k * n + r], b = B[(size_t)c * n + k];
            sx = fma(a.x, b.x, fma(-a.y, b.y, sx));
            sy = fma(a.x, b.y, fma(a.y, b.x, sy));
        }
        C[e] = make_double2(sx, sy);
    }
}
__device__ __forceinline__ void cta_copy(int n, const double2 *A, double2 *C) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) C[e] = A[e];
}
__device__ __forceinline__ void cta_identity(int n, double2 *C) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) C[e] = make_double2((e % n) == (e / n) ? 1.0 : 0.0, 0.0);
}

// One CTA per sweep point.  seg[s][k] = propagator k of solver trajectory s (n*n complex each, n_save per trajectory);
// work: 4 matrices per CTA {P = U_tau^q, W = P U(t0), T (scratch), U1 = U(t1)}.
__global__ void __launch_bounds__(256) floquet_compose_kernel(int n, int n_ts, int n_save, const double2 *__restrict__ seg,
                                                              const TrajRef *__restrict__ trajs,
                                                              const OutRef *__restrict__ refs, double2 *work,
                                                              double2 *out, long long out_first, long long out_step) {
    const long long traj = blockIdx.x;
    const TrajRef tr = trajs[traj];
    const size_t nn = (size_t)n * n;
    double2 *P = work + (size_t)traj * 4 * nn, *W = P + nn, *T = W + nn, *U1 = T + nn;
    const double2 *mid = tr.mid_seg >= 0 ? seg + (size_t)tr.mid_seg * n_save * nn : nullptr;
    const double2 *u0 = tr.rise_seg >= 0 ? seg + ((size_t)tr.rise_seg * n_save + tr.rise_t0_save) * nn : nullptr;
    const double2 *utau = mid ? mid + (size_t)tr.period_save * nn : nullptr;
    double2 *dst = out + (size_t)(out_first + traj * out_step) * n_ts * nn;
    const OutRef *rf = refs + (size_t)traj * n_ts;

    // P = I, W = U(t0)
    cta_identity(n, P);
    if (u0) cta_copy(n, u0, W); else cta_identity(n, W);
    __syncthreads();
    int q = 0;
    auto advance = [&](int q_to) {   // P <- P U_tau^(q_to - q), W <- P U(t0); uniform over the CTA
        while (q < q_to) {
            cta_matmul(n, P, utau, T);
            __syncthreads();
            cta_copy(n, T, P);
            __syncthreads();
            q++;
        }
        if (u0) cta_matmul(n, P, u0, W); else cta_copy(n, P, W);
        __syncthreads();
    };
    // rise and middle outputs in order of time (quotients never decrease)
    for (int k = 0; k < n_ts; k++) {
        const OutRef r = rf[k];
        if (r.kind == 0) {
            cta_copy(n, seg + ((size_t)tr.rise_seg * n_save + r.save) * nn, dst + (size_t)k * nn);
        } else if (r.kind == 1) {
            if (r.quot > q) advance(r.quot);
            cta_matmul(n, mid + (size_t)r.save * nn, W, dst + (size_t)k * nn);
        }
    }
    // U(t1), then the fall outputs
    if (tr.fall_seg >= 0) {
        if (tr.t1_quot > q) advance(tr.t1_quot);
        cta_matmul(n, mid + (size_t)tr.t1_save * nn, W, U1);
        __syncthreads();
        for (int k = 0; k < n_ts; k++) {
            const OutRef r = rf[k];
            if (r.kind == 2) cta_matmul(n, seg + ((size_t)tr.fall_seg * n_save + r.save) * nn, U1, dst + (size_t)k * nn);
        }
    }
}

// unique_tol + sort of decompose_times_floquet for one trajectory's periodic times: tsm sorted, tsm[0] = t0.
// Returns the sorted unique remainders and, per time, (index into them, quotient).
void decompose(const std::vector<double> &tsm, double period, double rtol, std::vector<double> &uniq, std::vector<int> &idx,
               std::vector<int> &quot) {
    const size_t m = tsm.size();
    const double t0 = tsm[0], dt = rtol * period;
    std::vector<double> vals(m);
    quot.resize(m);
    for (size_t i = 0; i < m; i++) {
        // Julia: mod(x, y) is the exact remainder, fld(x, y) = round((x - mod(x, y)) / y) -- consistent by construction
        const double x = tsm[i] - t0;                            // >= 0: tsm is sorted and starts at t0
        const double rem = std::fmod(x, period);
        quot[i] = (int)std::nearbyint((x - rem) / period);
        vals[i] = std::nearbyint((rem + t0) / dt) * dt;          // unique_tol: round(ts / dt) * dt
    }
    uniq = vals;
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    idx.resize(m);
    for (size_t i = 0; i < m; i++) idx[i] = (int)(std::lower_bound(uniq.begin(), uniq.end(), vals[i]) - uniq.begin());
}

}  // namespace

extern "C" int qsim_floquet_propagator(qsim_ctx *ctx, qsim_problem *prob, int which, int64_t n_traj, const double *params,
                                       const double *ts, int n_ts, int64_t ts_stride, const double *t_period,
                                       int64_t period_stride, double rise_time, double fall_time, double rtol,
                                       const qsim_solver_opts *opts, double *out, qsim_traj_stats *stats) {
    if (!ctx) return qsim_internal_fail(QSIM_ERR_NO_DEVICE, "ctx is NULL (needs a CUDA device; no CPU fallback)");
    if (!prob || !ts || !t_period || !out) return qsim_internal_fail(QSIM_ERR_ARG, "NULL argument");
    if (which != 0 && which != 2) return qsim_internal_fail(QSIM_ERR_ARG, "which must be 0 (unitary_propagator) or 2 (me_propagator)");
    if (n_traj < 0 || n_ts < 1) return qsim_internal_fail(QSIM_ERR_ARG, "need n_traj >= 0 and at least one time");
    if (ts_stride != 0 && ts_stride < n_ts) return qsim_internal_fail(QSIM_ERR_ARG, "ts_stride must be 0 or >= n_ts");
    if (!(rise_time >= 0) || !(fall_time >= 0)) return qsim_internal_fail(QSIM_ERR_ARG, "rise_time and fall_time must be >= 0");
    if (opts && ((opts->flags & QSIM_FLAG_DEVICE_PTRS) || opts->stream))
        return qsim_internal_fail(QSIM_ERR_ARG, "qsim_floquet_propagator takes host buffers and uses the ctx's stream");
    if (n_traj == 0) return QSIM_OK;
    int d = 0, n_params = 0;
    if (int rc = qsim_problem_dims(prob, &d, &n_params)) return rc;
    const int n = which == 0 ? d : d * d;
    const size_t nn = (size_t)n * n;

    // ---- plan the segments of every sweep point (host) ----
    struct Seg {
        int64_t traj;
        std::vector<double> grid;
    };
    std::vector<Seg> segs;
    std::vector<TrajRef> trajs((size_t)n_traj);
    std::vector<OutRef> refs((size_t)n_traj * n_ts);
    size_t n_save = 2;
    for (int64_t i = 0; i < n_traj; i++) {
        const double *t = ts + (size_t)i * ts_stride;
        const double period = t_period[(size_t)i * period_stride];
        for (int k = 1; k < n_ts; k++)
            if (!(t[k - 1] <= t[k])) return qsim_internal_fail(QSIM_ERR_ARG, "ts must be sorted");
        if (!(period > 0) || !std::isfinite(period)) return qsim_internal_fail(QSIM_ERR_ARG, "t_period must be positive");
        const double t0 = t[0] + rise_time, t1 = t[n_ts - 1] - fall_time;
        if (!(t0 <= t1)) return qsim_internal_fail(QSIM_ERR_ARG, "rise_time + fall_time exceed the span of ts");
        const double tol = rtol > 0 ? rtol : (std::nextafter(period, std::numeric_limits<double>::infinity()) - period);  // eps(t_period)
        TrajRef tr;
        std::memset(&tr, 0, sizeof(tr));
        tr.rise_seg = tr.mid_seg = tr.fall_seg = -1;
        OutRef *rf = refs.data() + (size_t)i * n_ts;
        // rise: [ts < t0; t0]
        Seg rise{i, {}}, mid{i, {}}, fall{i, {}};
        std::vector<double> tsm{t0};
        std::vector<int> mid_k;
        fall.grid.push_back(t1);
        for (int k = 0; k < n_ts; k++) {
            if (t[k] < t0) {
                rf[k] = OutRef{0, (int32_t)rise.grid.size(), 0, 0};
                rise.grid.push_back(t[k]);
            } else if (t[k] <= t1) {
                mid_k.push_back(k);
                tsm.push_back(t[k]);
            } else {
                rf[k] = OutRef{2, (int32_t)fall.grid.size(), 0, 0};
                fall.grid.push_back(t[k]);
            }
        }
        tsm.push_back(t1);
        if (!rise.grid.empty()) {
            tr.rise_t0_save = (int32_t)rise.grid.size();
            rise.grid.push_back(t0);
            tr.rise_seg = (int32_t)segs.size();
            segs.push_back(std::move(rise));
        }
        std::vector<double> uniq;
        std::vector<int> idx, quot;
        decompose(tsm, period, tol, uniq, idx, quot);
        for (size_t j = 0; j < mid_k.size(); j++) rf[mid_k[j]] = OutRef{1, idx[j + 1], quot[j + 1], 0};
        tr.t1_save = idx.back();
        tr.t1_quot = quot.back();
        tr.period_save = (int32_t)uniq.size();
        mid.grid = uniq;
        mid.grid.push_back(period + t0);
        tr.mid_seg = (int32_t)segs.size();
        segs.push_back(std::move(mid));
        if (fall.grid.size() > 1) {
            tr.fall_seg = (int32_t)segs.size();
            segs.push_back(std::move(fall));
        }
        trajs[(size_t)i] = tr;
    }
    for (const auto &s : segs) n_save = std::max(n_save, s.grid.size());
    const int64_t n_seg = (int64_t)segs.size();
    const int np = std::max(n_params, 1);
    std::vector<double> seg_params((size_t)n_seg * np, 0.0), seg_ts((size_t)n_seg * n_save);
    for (int64_t s = 0; s < n_seg; s++) {
        if (params && n_params > 0)
            std::memcpy(&seg_params[(size_t)s * np], params + (size_t)segs[(size_t)s].traj * n_params, sizeof(double) * n_params);
        const auto &g = segs[(size_t)s].grid;
        for (size_t k = 0; k < n_save; k++) seg_ts[(size_t)s * n_save + k] = g[std::min(k, g.size() - 1)];   // pad: repeat the last time
    }

    // ---- one batched solve of all segments, results left on the device ----
    void *d_seg = nullptr, *d_work = nullptr, *d_idx = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
    const size_t seg_bytes = sizeof(double2) * nn * n_save * (size_t)n_seg;
    const size_t idx_bytes = sizeof(TrajRef) * (size_t)n_traj + sizeof(OutRef) * (size_t)n_traj * n_ts;
    if (int rc = qsim_internal_floquet_buffers(ctx, seg_bytes, sizeof(double2) * 4 * nn * (size_t)n_traj, idx_bytes, &d_seg, &d_work,
                                               &d_idx, &stream, &device))
        return rc;
    std::vector<qsim_traj_stats> seg_stats((size_t)n_seg);
    {
        qsim_solver_opts o;
        if (opts) o = *opts; else qsim_default_opts(&o);
        if (int rc = qsim_internal_solve_to_device(ctx, prob, which, n_seg, seg_params.data(), seg_ts.data(), (int)n_save,
                                                   (int64_t)n_save, &o, d_seg, seg_stats.data()))
            return rc;
    }
    // ---- powers and products on the device ----
    TrajRef *d_trajs = (TrajRef *)d_idx;
    OutRef *d_refs = (OutRef *)((char *)d_idx + sizeof(TrajRef) * (size_t)n_traj);
    CUF(cudaMemcpyAsync(d_trajs, trajs.data(), sizeof(TrajRef) * (size_t)n_traj, cudaMemcpyHostToDevice, stream));
    CUF(cudaMemcpyAsync(d_refs, refs.data(), sizeof(OutRef) * refs.size(), cudaMemcpyHostToDevice, stream));
    // results: straight into a page-locked caller buffer, else into a device buffer copied back afterwards
    cudaPointerAttributes at;
    double2 *d_out = nullptr;
    bool direct = false;
    if (cudaPointerGetAttributes(&at, out) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
        d_out = (double2 *)at.devicePointer;
        direct = true;
    } else {
        cudaGetLastError();
        CUF(cudaMalloc((void **)&d_out, sizeof(double2) * nn * (size_t)n_ts * (size_t)n_traj));
    }
    const int block = (int)std::min<size_t>(256, std::max<size_t>(32, ((nn + 31) / 32) * 32));
    floquet_compose_kernel<<<(unsigned)n_traj, block, 0, stream>>>(n, n_ts, (int)n_save, (const double2 *)d_seg, d_trajs, d_refs,
                                                                 (double2 *)d_work, d_out, 0, 1);
    CUF(cudaGetLastError());
    qsim_internal_count_launch(ctx);
    if (!direct) {
        CUF(cudaMemcpyAsync(out, d_out, sizeof(double2) * nn * (size_t)n_ts * (size_t)n_traj, cudaMemcpyDeviceToHost, stream));
        CUF(cudaStreamSynchronize(stream));
        CUF(cudaFree(d_out));
    } else {
        CUF(cudaStreamSynchronize(stream));
    }
    if (stats) {
        for (int64_t i = 0; i < n_traj; i++) {
            qsim_traj_stats st;
            std::memset(&st, 0, sizeof(st));
            const TrajRef &tr = trajs[(size_t)i];
            for (int sgi : {tr.rise_seg, tr.mid_seg, tr.fall_seg}) {
                if (sgi < 0) continue;
                const qsim_traj_stats &s = seg_stats[(size_t)sgi];
                if (s.retcode != QSIM_RET_SUCCESS && st.retcode == QSIM_RET_SUCCESS) st.retcode = s.retcode;
                st.n_accept += s.n_accept;
                st.n_reject += s.n_reject;
                st.n_rhs += s.n_rhs;
                if (sgi == tr.mid_seg) st.dt_init = s.dt_init;
            }
            st.t_final = ts[(size_t)i * ts_stride + n_ts - 1];
            stats[i] = st;
        }
    }
    return QSIM_OK;
}
