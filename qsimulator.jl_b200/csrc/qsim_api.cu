// C ABI of the B200 time-evolution library (include/qsim_b200.h): contexts, problem assembly,
// kernel dispatch, and the four batched solver entry points replacing
// unitary_propagator / unitary_state / me_propagator / me_state
// (/root/reference/src/time_evolution.jl:29-164).  No CPU fallback: solver calls need a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "qsim_registry.h"

namespace qsim {
void build_generator(const qsim_problem &p, int which, Generator &g);
// qsim_layout.cpp: bank-conflict-aware placement for the row-split kernels
void rowsplit_rows(const Generator &g, std::vector<std::vector<uint32_t>> &compact, std::vector<int> &perm);
void rowsplit_groups(const Generator &g, const std::vector<std::vector<uint32_t>> &compact, const std::vector<int> &perm,
                     int what, std::vector<std::vector<int>> &groups);
void anneal_colors(const std::vector<std::vector<int>> &groups, const std::vector<int> &cls, int n_colors, long iters,
                   std::vector<int> &color);
long layout_groups_cost(const std::vector<std::vector<int>> &groups, const std::vector<int> &color, int n_colors);
void optimize_value_slots(Generator &g, long iters);
void host_eval_generator(const Generator &g, double t, const double *row, cplx *dense);
const SeriesTables &series_tables();
}  // namespace qsim

using namespace qsim;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(QSIM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

#define QSIM_N_COUNTERS 4
struct qsim_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop;
    SeriesTables *d_tab = nullptr;
    double *d_cheb = nullptr;  // Chebyshev transform / nodes / check points for the drive windows
    cudaStream_t stream2 = nullptr;            // second launch stream of pipelined host-buffer calls
    unsigned long long *d_counter = nullptr;   // work-queue counters, one per launch in flight (QSIM_N_COUNTERS)
    DevBuf params, ts, init, out, out2, stats, scratch, scratch2, sel;
    DevBuf fl_seg, fl_work, fl_idx;            // qsim_floquet_propagator: segment propagators, power workspace, index tables
    void *stage[2] = {nullptr, nullptr};       // page-locked staging buffers of the pipelined result copies
    size_t stage_cap[2] = {0, 0};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // around the kernel launch(es) of the last solver call
    bool timed = false;
    int64_t launches = 0;
};

struct qsim_problem::DeviceCopy {
    int device = -1;
    uint32_t *ell = nullptr;
    uint8_t *row_len = nullptr;
    int32_t *slot_ptr = nullptr, *pair_chan = nullptr;
    double2 *pair_w = nullptr;
    EvalDesc *evals = nullptr;
    double *dyn_rec = nullptr;       // 64-byte records of the time-dependent slots
    double *tables = nullptr;        // tabulated envelopes (nullptr when the problem has none)
    int fast_ok = 0, fast_chA = 0, fast_chB = 0;  // records are in (constant, channel A, channel B) order
    // row-split kernels: sorted, translated rows; row permutation; chunk offsets and lengths.  One set per
    // (stage stride, team warps, value mode): me_propagator and me_state of one problem use different sets and may
    // run concurrently from different ctxs, so sets are never replaced, only added (freed with the problem).
    struct Packed {
        int stride, warps, vm, cu = 0;
        const ComponentForm *cf = nullptr;
        uint32_t *ell = nullptr;
        int *perm = nullptr, *off = nullptr, *len = nullptr, *pos = nullptr, *pass = nullptr;
    };
    std::vector<Packed> packed;
    uint8_t *slot_imag = nullptr;    // value mode 4
    uint32_t *split_ell = nullptr;   // split rows: row table with the helper rows appended, per-lane info
    uint8_t *split_info = nullptr;
    uint32_t *blk_ell = nullptr;     // blocked kernels: row table and value slot ids (built by make_plan)
    uint16_t *blk_slot = nullptr;
    uint32_t *tr_ell = nullptr;      // tiered kernels: row table over virtual rows, virtual -> state row
    int *tr_perm = nullptr, *tr_pos = nullptr;
    uint32_t *as_ell = nullptr;      // additively split rows: row table over the table rows, table row -> state row, lane info
    int *as_phys = nullptr;
    uint8_t *as_info = nullptr;
    // component forms (the propagator's, and one per support pattern of the initial states seen): per-warp row table,
    // local row -> state row, (batch, local row, c) -> state column, structurally zero entries
    struct CfDev {
        const ComponentForm *form = nullptr;
        uint32_t *ell = nullptr;
        int *vrow = nullptr, *vcol = nullptr, *zeros = nullptr;
    };
    std::vector<CfDev> cf_sets;
    ~DeviceCopy() { free_all(); }  // the owning device must be current
    void free_all() {
        for (auto &pk : packed) {
            cudaFree(pk.ell);
            cudaFree(pk.perm);
            cudaFree(pk.off);
            cudaFree(pk.len);
            cudaFree(pk.pos);
            cudaFree(pk.pass);
        }
        packed.clear();
        cudaFree(slot_imag);
        slot_imag = nullptr;
        cudaFree(split_ell);
        cudaFree(split_info);
        split_ell = nullptr;
        split_info = nullptr;
        cudaFree(blk_ell);
        cudaFree(blk_slot);
        blk_ell = nullptr;
        blk_slot = nullptr;
        cudaFree(tr_ell);
        cudaFree(tr_perm);
        cudaFree(tr_pos);
        tr_ell = nullptr;
        tr_perm = tr_pos = nullptr;
        cudaFree(as_ell);
        cudaFree(as_phys);
        cudaFree(as_info);
        as_ell = nullptr;
        as_phys = nullptr;
        as_info = nullptr;
        for (auto &q : cf_sets) {
            cudaFree(q.ell);
            cudaFree(q.vrow);
            cudaFree(q.vcol);
            cudaFree(q.zeros);
        }
        cf_sets.clear();
        cudaFree(dyn_rec);
        dyn_rec = nullptr;
        cudaFree(tables);
        tables = nullptr;
        cudaFree(ell); cudaFree(row_len); cudaFree(slot_ptr); cudaFree(pair_chan); cudaFree(pair_w); cudaFree(evals);
        ell = nullptr; row_len = nullptr; slot_ptr = nullptr; pair_chan = nullptr; pair_w = nullptr; evals = nullptr;
    }
};

extern "C" {

int qsim_version(void) { return QSIM_ABI_VERSION; }
const char *qsim_last_error(void) { return g_err.c_str(); }

void qsim_default_opts(qsim_solver_opts *o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->reltol = 1e-6;   // time_evolution.jl:43
    o->abstol = 1e-8;
    o->gamma = 0.9;     // OrdinaryDiffEq defaults for Tsit5
    o->qmin = 0.2;
    o->qmax = 10.0;
    o->beta1 = 7.0 / 50.0;
    o->beta2 = 2.0 / 25.0;
    o->qoldinit = 1e-4;
    o->dtmax = 0.0;
    o->maxiters = 1000000;
}

// ---------------------------------------------------------------- context
int qsim_create(int device, qsim_ctx **out) {
    if (!out) return fail(QSIM_ERR_ARG, "qsim_create: ctx is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(QSIM_ERR_NO_DEVICE, std::string("no usable CUDA device (") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                            "); this library has no CPU fallback");
    if (device < 0 || device >= count) return fail(QSIM_ERR_ARG, "qsim_create: device index out of range");
    CU(cudaSetDevice(device));
    qsim_ctx *c = new qsim_ctx();
    c->device = device;
    // (a failing step below returns through this guard: the partially built ctx is released, not leaked)
    struct Guard {
        qsim_ctx *c;
        ~Guard() { if (c) qsim_destroy(c); }
    } guard{c};
    CU(cudaGetDeviceProperties(&c->prop, device));
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&c->d_tab, sizeof(SeriesTables)));
    CU(cudaMemcpy(c->d_tab, &series_tables(), sizeof(SeriesTables), cudaMemcpyHostToDevice));
    CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CU(cudaMalloc(&c->d_counter, QSIM_N_COUNTERS * sizeof(unsigned long long)));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    CU(cudaEventCreateWithFlags(&c->ev_copy[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_copy[1], cudaEventDisableTiming));
    {
        // c_k = sum_j T[k][j] f(x_j), x_j = cos(pi (j + 1/2) / N): p(x) = sum_k c_k T_k(x)
        const int N = QSIM_NN;
        std::vector<double> cheb((size_t)N * N + N + 3);
        const double pi = 3.14159265358979323846;
        for (int k = 0; k < N; k++)
            for (int j = 0; j < N; j++)
                cheb[(size_t)k * N + j] = (k == 0 ? 1.0 : 2.0) / N * std::cos(pi * k * (j + 0.5) / N);
        for (int j = 0; j < N; j++) cheb[(size_t)N * N + j] = std::cos(pi * (j + 0.5) / N);
        cheb[(size_t)N * N + N + 0] = -0.85;
        cheb[(size_t)N * N + N + 1] = 0.1;
        cheb[(size_t)N * N + N + 2] = 0.92;
        CU(cudaMalloc(&c->d_cheb, sizeof(double) * cheb.size()));
        CU(cudaMemcpy(c->d_cheb, cheb.data(), sizeof(double) * cheb.size(), cudaMemcpyHostToDevice));
    }
    guard.c = nullptr;
    *out = c;
    return QSIM_OK;
}

int qsim_destroy(qsim_ctx *c) {
    if (!c) return QSIM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->stream2) cudaStreamSynchronize(c->stream2);
    c->params.release(); c->ts.release(); c->init.release(); c->out.release(); c->out2.release(); c->stats.release();
    c->scratch.release(); c->scratch2.release(); c->sel.release();
    c->fl_seg.release(); c->fl_work.release(); c->fl_idx.release();
    if (c->ev0) cudaEventDestroy(c->ev0);      // (members are null when qsim_create failed half way)
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (int i = 0; i < 2; i++) {
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
        if (c->stage[i]) cudaFreeHost(c->stage[i]);
    }
    if (c->stream2) cudaStreamDestroy(c->stream2);
    cudaFree(c->d_tab);
    cudaFree(c->d_cheb);
    cudaFree(c->d_counter);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return QSIM_OK;
}

int qsim_device_info(qsim_ctx *c, int *sm_count, int *cc_major, int *cc_minor, double *sm_clock_mhz,
                     int64_t *global_mem_bytes) {
    if (!c) return fail(QSIM_ERR_ARG, "ctx is NULL");
    if (sm_count) *sm_count = c->prop.multiProcessorCount;
    if (cc_major) *cc_major = c->prop.major;
    if (cc_minor) *cc_minor = c->prop.minor;
    if (sm_clock_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
        *sm_clock_mhz = khz / 1000.0;
    }
    if (global_mem_bytes) *global_mem_bytes = (int64_t)c->prop.totalGlobalMem;
    return QSIM_OK;
}

int qsim_synchronize(qsim_ctx *c) {
    if (!c) return fail(QSIM_ERR_ARG, "ctx is NULL");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return QSIM_OK;
}

int64_t qsim_launch_count(qsim_ctx *c) { return c ? c->launches : 0; }

int qsim_last_kernel_ms(qsim_ctx *c, double *ms) {
    if (!c || !ms) return fail(QSIM_ERR_ARG, "NULL argument");
    if (!c->timed) return fail(QSIM_ERR_STATE, "no solver call has completed on this ctx yet");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->ev1));
    float f = 0.0f;
    CU(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = (double)f;
    return QSIM_OK;
}

int qsim_alloc_pinned(size_t bytes, void **ptr) {
    if (!ptr) return fail(QSIM_ERR_ARG, "ptr is NULL");
    // portable + mapped: every device of a multi-ctx call may write results straight into it
    CU(cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    return QSIM_OK;
}
int qsim_free_pinned(void *ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return QSIM_OK;
}

// ---------------------------------------------------------------- problem
int qsim_problem_create(int d, int n_params, qsim_problem **out) {
    if (!out) return fail(QSIM_ERR_ARG, "prob is NULL");
    if (d < 1 || d > 255) return fail(QSIM_ERR_ARG, "qsim_problem_create: dimension must be in [1, 255]");
    if (n_params < 0) return fail(QSIM_ERR_ARG, "qsim_problem_create: n_params < 0");
    qsim_problem *p = new qsim_problem();
    p->d = d;
    p->n_params = n_params;
    p->h0.assign((size_t)d * d, cplx(0, 0));
    *out = p;
    return QSIM_OK;
}

int qsim_problem_destroy(qsim_problem *p) {
    if (!p) return QSIM_OK;
    for (int w = 0; w < 2; w++)
        for (auto *dc : p->dev[w]) {
            cudaSetDevice(dc->device);
            delete dc;
        }
    delete p;
    return QSIM_OK;
}

static int check_open(qsim_problem *p) {
    if (!p) return fail(QSIM_ERR_ARG, "prob is NULL");
    if (p->finalized) return fail(QSIM_ERR_STATE, "problem is already finalized");
    return QSIM_OK;
}
static int check_slot(const qsim_problem *p, const qsim_slot &s, const char *what) {
    if (s.param >= p->n_params) return fail(QSIM_ERR_ARG, std::string(what) + ": parameter index out of range");
    return QSIM_OK;
}
static int check_signal(const qsim_problem *p, const qsim_signal &s) {
    if (s.carrier < 0 || s.carrier > 2 || s.envelope < 0 || s.envelope > 3 || s.blend < 0 || s.blend > 1)
        return fail(QSIM_ERR_ARG, "signal: unknown carrier / envelope / blend");
    if (s.envelope == QSIM_ENV_TABLE && (s.table < 1 || s.table > (int)p->table_offset.size()))
        return fail(QSIM_ERR_ARG, "signal: table id does not name a table added with qsim_problem_add_table");
    const qsim_slot *sl[8] = {&s.offset, &s.amp, &s.freq, &s.phase, &s.t1, &s.t2, &s.sigma, &s.rest};
    for (auto q : sl)
        if (int rc = check_slot(p, *q, "signal")) return rc;
    return QSIM_OK;
}

int qsim_problem_add_table(qsim_problem *p, double t_start, double t_end, int n_seg, int n_coef, const double *coef,
                           int32_t *table_id) {
    if (int rc = check_open(p)) return rc;
    if (!coef || !table_id) return fail(QSIM_ERR_ARG, "NULL argument");
    if (!(t_end > t_start) || !std::isfinite(t_start) || !std::isfinite(t_end))
        return fail(QSIM_ERR_ARG, "table: need t_start < t_end");
    if (n_seg < 1 || n_seg > (1 << 20) || n_coef < 2 || n_coef > 16)
        return fail(QSIM_ERR_ARG, "table: n_seg must be in [1, 2^20] and n_coef in [2, 16]");
    for (size_t i = 0; i < (size_t)n_seg * n_coef; i++)
        if (!std::isfinite(coef[i])) return fail(QSIM_ERR_ARG, "table: non-finite coefficient");
    p->table_offset.push_back((int32_t)p->tables.size());
    p->tables.push_back(t_start);
    p->tables.push_back((double)n_seg / (t_end - t_start));
    p->tables.push_back((double)n_seg);
    p->tables.push_back((double)n_coef);
    p->tables.insert(p->tables.end(), coef, coef + (size_t)n_seg * n_coef);
    *table_id = (int32_t)p->table_offset.size();  // ids start at 1; 0 means "no table"
    return QSIM_OK;
}

int qsim_problem_set_h0(qsim_problem *p, const double *h0) {
    if (int rc = check_open(p)) return rc;
    if (!h0) return fail(QSIM_ERR_ARG, "h0 is NULL");
    std::memcpy(p->h0.data(), h0, sizeof(double) * 2 * p->d * p->d);
    return QSIM_OK;
}

int qsim_problem_add_term(qsim_problem *p, const qsim_term *t, const double *atom_a, const double *atom_b,
                          const int32_t *levels) {
    if (int rc = check_open(p)) return rc;
    if (!t) return fail(QSIM_ERR_ARG, "term is NULL");
    if (int rc = check_signal(p, t->signal)) return rc;
    const size_t dd = (size_t)p->d * p->d;
    qsim_problem::Term term;
    term.desc = *t;
    auto copy_atom = [&](const double *src, std::vector<cplx> &dst) {
        dst.resize(dd);
        std::memcpy(dst.data(), src, sizeof(double) * 2 * dd);
    };
    switch (t->kind) {
        case QSIM_TERM_SCALAR:
            if (!atom_a) return fail(QSIM_ERR_ARG, "SCALAR term needs atom_a");
            copy_atom(atom_a, term.a);
            break;
        case QSIM_TERM_ROTATING:
            if (!atom_a || !atom_b) return fail(QSIM_ERR_ARG, "ROTATING term needs atom_a and atom_b");
            if (int rc = check_slot(p, t->rate, "rate")) return rc;
            copy_atom(atom_a, term.a);
            copy_atom(atom_b, term.b);
            break;
        case QSIM_TERM_DUFFING:
        case QSIM_TERM_TRANSMON:
            if (!levels) return fail(QSIM_ERR_ARG, "DUFFING/TRANSMON term needs levels");
            if (int rc = check_slot(p, t->rot, "rot")) return rc;
            if (t->kind == QSIM_TERM_DUFFING) {
                if (int rc = check_slot(p, t->anharm, "anharm")) return rc;
            } else {
                if (t->num_terms < 1) return fail(QSIM_ERR_ARG, "num_terms < 1");  // perturbative_transmon.jl:225
                if (int rc = check_slot(p, t->EC, "EC")) return rc;
                if (int rc = check_slot(p, t->EJ1, "EJ1")) return rc;
                if (int rc = check_slot(p, t->EJ2, "EJ2")) return rc;
            }
            term.levels.assign(levels, levels + p->d);
            for (int v : term.levels)
                if (v < 0 || v > 1023) return fail(QSIM_ERR_ARG, "levels out of range");
            break;
        default:
            return fail(QSIM_ERR_ARG, "unknown term kind");
    }
    p->terms.push_back(std::move(term));
    return QSIM_OK;
}

int qsim_problem_add_spectrum_term(qsim_problem *p, const qsim_signal *flux, const qsim_slot *rot, const int32_t *levels,
                                   int n_levels, const int32_t *table_ids) {
    if (int rc = check_open(p)) return rc;
    if (!flux || !levels || !table_ids) return fail(QSIM_ERR_ARG, "NULL argument");
    if (n_levels < 1) return fail(QSIM_ERR_ARG, "spectrum term: n_levels < 1");
    if (int rc = check_signal(p, *flux)) return rc;
    qsim_problem::Term term;
    std::memset(&term.desc, 0, sizeof(term.desc));
    term.desc.kind = QSIM_TERM_SPECTRUM;
    term.desc.num_terms = n_levels;
    term.desc.signal = *flux;
    term.desc.rot.param = -1;
    if (rot) {
        if (int rc = check_slot(p, *rot, "rot")) return rc;
        term.desc.rot = *rot;
    }
    for (qsim_slot *sl : {&term.desc.rate, &term.desc.anharm, &term.desc.EC, &term.desc.EJ1, &term.desc.EJ2}) sl->param = -1;
    term.levels.assign(levels, levels + p->d);
    for (int v : term.levels)
        if (v < 0 || v >= n_levels) return fail(QSIM_ERR_ARG, "spectrum term: level without a table");
    term.tables.assign(table_ids, table_ids + n_levels);
    for (int v : term.tables)
        if (v < 1 || v > (int)p->table_offset.size())
            return fail(QSIM_ERR_ARG, "spectrum term: table id does not name a table added with qsim_problem_add_table");
    p->terms.push_back(std::move(term));
    return QSIM_OK;
}

int qsim_problem_add_lindblad(qsim_problem *p, const double *lind, const qsim_signal *scale) {
    if (int rc = check_open(p)) return rc;
    if (!lind) return fail(QSIM_ERR_ARG, "lind is NULL");
    qsim_problem::Lind l;
    const size_t dd = (size_t)p->d * p->d;
    l.op.resize(dd);
    std::memcpy(l.op.data(), lind, sizeof(double) * 2 * dd);
    l.has_scale = scale != nullptr;
    if (scale) {
        if (int rc = check_signal(p, *scale)) return rc;
        l.scale = *scale;
    } else {
        std::memset(&l.scale, 0, sizeof(l.scale));
    }
    p->linds.push_back(std::move(l));
    return QSIM_OK;
}

int qsim_problem_finalize(qsim_problem *p) {
    if (!p) return fail(QSIM_ERR_ARG, "prob is NULL");
    p->finalized = true;
    return QSIM_OK;
}

// Value mode 4 ("mixed"): every off-diagonal entry purely real or purely imaginary for all times and parameters
// (all channel weights of its slot real, or all imaginary); the diagonal may be complex.  slot_imag[sl] = 1 for the
// purely imaginary slots.
static bool generator_is_mixed(const Generator &g, std::vector<uint8_t> &slot_imag) {
    const int n = g.n;
    bool mixed = g.n_slots <= 2045;  // 8-byte offsets into both halves of the split table fit 15 bits
    slot_imag.assign((size_t)g.n_slots, 0);
    std::vector<uint8_t> off_diag((size_t)g.n_slots, 0);
    for (int j = 1; j < g.max_nnz; j++)
        for (int r = 0; r < n; r++) {
            const uint32_t sl = g.ell[(size_t)j * n + r] >> 16;
            if (sl != 0xFFFFu) off_diag[sl] = 1;
        }
    for (int sl = 0; sl < g.n_slots && mixed; sl++) {
        bool any_re = false, any_im = false;
        for (int q = g.slot_ptr[sl]; q < g.slot_ptr[sl + 1]; q++) {
            any_re = any_re || g.pair_w[(size_t)q].real() != 0.0;
            any_im = any_im || g.pair_w[(size_t)q].imag() != 0.0;
        }
        if (off_diag[(size_t)sl] && any_re && any_im) mixed = false;
        slot_imag[(size_t)sl] = any_im ? 1 : 0;
    }
    return mixed;
}

static long anneal_iters(int n_items) {
    if (const char *env = std::getenv("QSIM_ANNEAL_ITERS")) return std::atol(env);   // development: 0 disables
    return std::min<long>(1500000, 2000L * n_items);
}

static int ensure_generator(qsim_problem *p, int w) {
    if (!p->finalized) return fail(QSIM_ERR_STATE, "problem is not finalized");
    std::lock_guard<std::mutex> lock(p->dev_mutex);
    if (p->gen_built[w]) return QSIM_OK;
    try {
        build_generator(*p, w, p->gen[w]);
        Generator &g = p->gen[w];
        // Row-split kernels with a complex value table (16 bytes per slot): renumber the slots so that the values a
        // quarter warp reads together sit in distinct shared-memory bank groups (qsim_layout.cpp).
        std::vector<uint8_t> slot_imag;
        if ((g.n > 96 || g.max_nnz > 16) && !g.imag_only && !generator_is_mixed(g, slot_imag) && anneal_iters(g.n_slots) > 0)
            optimize_value_slots(g, anneal_iters(g.n_slots));
    } catch (const std::exception &e) {
        return fail(QSIM_ERR_UNSUPPORTED, std::string("generator assembly failed: ") + e.what());
    }
    p->gen_built[w] = true;
    return QSIM_OK;
}

int qsim_problem_eval(qsim_problem *p, int kind, double t, const double *param_row, double *out) {
    if (!p || !out) return fail(QSIM_ERR_ARG, "NULL argument");
    if (kind != 0 && kind != 1) return fail(QSIM_ERR_ARG, "kind must be 0 or 1");
    if (p->n_params > 0 && !param_row) return fail(QSIM_ERR_ARG, "param_row is NULL");
    if (int rc = ensure_generator(p, kind)) return rc;
    const Generator &g = p->gen[kind];
    std::vector<cplx> dense((size_t)g.n * g.n);
    host_eval_generator(g, t, param_row, dense.data());
    if (kind == 0) {
        // A = -2 pi i H  ->  H = A * i / (2 pi)
        const cplx f(0.0, 1.0 / QSIM_TWO_PI);
        for (auto &v : dense) v *= f;
    }
    std::memcpy(out, dense.data(), sizeof(cplx) * dense.size());
    return QSIM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- dispatch
struct Plan {
    const KernelVariant *kv = nullptr;
    int w = 0, ncols = 1, identity = 0;
    int groups = 1, n_batches = 1, team_warps = 1, teams_per_cta = 1, streamed = 0;
    int block = 32, grid = 1, nvp = 0, ncp = 0, common_bytes = 0, team_bytes = 0, n_dyn_pairs = 0;
    int y_row_stride = 1, y_grp_stride = 1, rowsplit = 0;
    int off_dptr = 0, off_dw = 0, off_drec = 0, dyn_fast = 0;
    size_t smem = 0, scratch_per_team = 0;
    int vm = 0;                        // value mode of the chosen variant
    // row-split teams: compact row table (see LaunchArgs::rs_off), row permutation, chunk offsets / lengths
    std::vector<uint32_t> rs_table;
    std::vector<int> rs_perm, rs_off, rs_len, rs_pos;
    std::vector<int> rs_lane;          // component passes: {row | col << 16, stage offset, table offset, len | unit offset << 8} per (pass, chunk, lane)
    int ell_in_smem = 0, off_ell = 0, ell_bytes = 0, rs_prefetch = 0, off_pf = 0;
    // split rows (LaunchArgs::split_info): virtual row table [split_nnz][split_nv], per-lane info
    int split_nv = 0, split_nnz = 0;
    std::vector<uint32_t> split_ell;
    std::vector<uint8_t> split_info;
    int cu = 0, rpl = 1;               // column-per-lane mapping (LaunchArgs::cu_map); rows per lane of the variant
    int lean = 0, off_park = 0;        // lean variant chosen (Kern<..., LEAN>); its park area inside the team region
    int tiered = 0;                    // tiered variant (value mode 5): blk_rows / blk_nnz hold its geometry
    const struct TieredForm *tf = nullptr;   // its table, row permutation and stage layout (cached in the problem)
    const struct AddSplitForm *af = nullptr; // additively split rows (value mode 7): table-row generator, maps (cached in the problem)
    const ComponentForm *cf = nullptr;       // component pruning of a propagator (qsim_components.cpp; cached in the problem)
    std::vector<uint8_t> slot_imag;    // value mode 4: per slot, 1 = purely imaginary, 0 = purely real
    // blocked variant (value mode 3): block rows, row table [1 + max blocks][n] and value slot ids
    int blk_rows = 0, blk_nnz = 0;
    std::vector<uint32_t> blk_ell;
    std::vector<uint16_t> blk_slot;
};

// Blocked form of a generator for the value-mode-3 kernels: rows grouped R at a time, every off-diagonal entry
// inside an R x R block (block row I, block column J != I) whose values are per-trajectory constants.
// Returns false when the generator does not have that shape.
static bool build_blocked(const Generator &g, int R, int capacity, Plan &pl) {
    const int n = g.n;
    if (n % R != 0 || !g.imag_only) return false;
    const int nb = n / R;
    std::vector<std::vector<int>> nbr((size_t)nb);
    for (int r = 0; r < n; r++)
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) == 0xFFFFu) continue;
            if ((int)(e >> 16) < g.n_dyn_slots) return false;       // time-dependent off-diagonal entry
            const int J = (int)(e & 0xFFFFu) / R, I = r / R;
            if (J == I) return false;                               // off-diagonal entry inside the diagonal block
            auto &v = nbr[(size_t)I];
            if (std::find(v.begin(), v.end(), J) == v.end()) v.push_back(J);
        }
    int mb = 0;
    for (auto &v : nbr) {
        std::sort(v.begin(), v.end());
        mb = std::max(mb, (int)v.size());
    }
    if (mb < 1 || mb + 1 > capacity) return false;
    const int nnz = capacity, MB = capacity - 1;
    pl.blk_rows = nb;
    pl.blk_nnz = nnz;
    pl.blk_ell.assign((size_t)nnz * n, 0);
    pl.blk_slot.assign((size_t)nb * MB * R * R, (uint16_t)g.n_slots);  // default: the zero slot
    for (int r = 0; r < n; r++) {
        const int I = r / R, k = r % R;
        pl.blk_ell[r] = g.ell[r];  // diagonal entry (own row, its slot)
        for (int b = 0; b < MB; b++) {
            const int J = b < (int)nbr[(size_t)I].size() ? nbr[(size_t)I][(size_t)b] : I;
            pl.blk_ell[(size_t)(1 + b) * n + r] = (uint32_t)(R * J + k) | (0xFFFFu << 16);
        }
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) == 0xFFFFu) continue;
            const int src = (int)(e & 0xFFFFu), J = src / R;
            const int b = (int)(std::find(nbr[(size_t)I].begin(), nbr[(size_t)I].end(), J) - nbr[(size_t)I].begin());
            pl.blk_slot[(((size_t)I * MB + b) * R + k) * R + (size_t)(src % R)] = (uint16_t)(e >> 16);
        }
    }
    return true;
}

// Tiered form of a generator for the value-mode-5 kernels (three rows of one column per lane).  Rows are sorted by
// their number of off-diagonal entries and cut into three tiers of n/3 rows; table row 3 b + i is the b-th row of
// tier i, so row-slot i of every lane carries rows of (nearly) the same length and the lanes of a warp execute
// caps[0] + caps[1] + caps[2] entries instead of 3 x the longest row.  Purely imaginary generators with static
// off-diagonal values only (the condition of value mode 2).
// The stage buffer holds state row r of column g at 16-byte unit pos[r] * S + g * Cg.  Which rows share a block,
// the order of a row's entries (entry j of the three blocks' row-slot i is ONE load instruction across the warp),
// the positions pos[] and the strides are chosen by a seeded annealing run against the exact wavefront count of the
// executed loads and stores (quarter warps of 8 lanes, 8 bank groups of 16 bytes), so the gathers run without bank
// conflicts where the pattern allows it.  The table is indexed like Generator::ell over table rows; the source
// field of an entry holds the stage position of its source row.  Result cached per generator (TieredForm).
struct TieredForm {
    bool ok = false;
    int caps[3] = {1, 1, 1}, nb = 0, nnz = 0, S = 1, Cg = 1;
    long cost = 0, floor = 0;
    std::vector<uint32_t> ell;   // [nnz][n] over table rows: stage position of the source | slot << 16
    std::vector<int> perm, pos;  // table row -> state row; table row -> stage position
};

static bool build_tiered(const Generator &g, int groups, TieredForm &tf) {
    const int n = g.n, R = 3;
    if (n % R != 0 || !g.imag_only || n > 32) return false;
    const int nb = n / R;
    std::vector<std::vector<uint32_t>> off((size_t)n);
    for (int r = 0; r < n; r++)
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) == 0xFFFFu) continue;
            if ((int)(e >> 16) < g.n_dyn_slots) return false;   // time-dependent off-diagonal entry
            off[(size_t)r].push_back(e);
        }
    std::vector<int> order((size_t)n);
    for (int r = 0; r < n; r++) order[(size_t)r] = r;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return off[(size_t)x].size() > off[(size_t)y].size(); });
    for (int i = 0; i < R; i++) {
        tf.caps[i] = 1;
        for (int b = 0; b < nb; b++) tf.caps[i] = std::max(tf.caps[i], 1 + (int)off[(size_t)order[(size_t)(i * nb + b)]].size());
    }
    if (tf.caps[0] > 7) return false;
    tf.nb = nb;
    tf.nnz = tf.caps[0];

    // search state: rows[i][b] = state row in (tier i, block b); ent[r] = entry order of row r; pos[r]; (S, Cg)
    struct State {
        std::vector<int> rows, pos;
        std::vector<std::vector<uint32_t>> ent;
        int S, Cg;
    };
    State cur;
    cur.rows = order;   // index i * nb + b
    cur.pos.resize((size_t)n);
    for (int r = 0; r < n; r++) cur.pos[(size_t)r] = r;
    cur.ent = off;
    cur.S = groups;
    cur.Cg = 1;
    const int nlanes = groups * nb;
    auto injective = [&](int S, int Cg) {
        std::vector<char> used((size_t)(n * S + groups * Cg + 1), 0);
        for (int r = 0; r < n; r++)
            for (int c = 0; c < groups; c++) {
                const int u = r * S + c * Cg;
                if (16 * u > 65535 - 16 || used[(size_t)u]) return false;
                used[(size_t)u] = 1;
            }
        return true;
    };
    auto cost_of = [&](const State &st) {
        long cost = 0;
        for (int i = 0; i < R; i++)
            for (int j = 0; j < tf.caps[i]; j++)      // j == 0: the lane's own store of row-slot i
                for (int q = 0; q < 4; q++) {
                    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[8][8];
                    bool any = false;
                    for (int l = 8 * q; l < 8 * q + 8; l++) {
                        const bool active = l < nlanes;
                        if (!active && j == 0) continue;            // inactive lanes do not store
                        const int le = active ? l : nlanes - 1;     // ... and mimic the last lane's loads
                        const int b = le % nb, c = le / nb;
                        const int r = st.rows[(size_t)(i * nb + b)];
                        int src = r;
                        if (j > 0) {
                            const auto &e = st.ent[(size_t)r];
                            src = j - 1 < (int)e.size() ? (int)(e[(size_t)(j - 1)] & 0xFFFFu) : r;   // padding: own row
                        }
                        const int unit = st.pos[(size_t)src] * st.S + c * st.Cg, bk = unit & 7;
                        bool dup = false;
                        for (int k = 0; k < cnt[bk]; k++) dup = dup || seen[bk][k] == unit;
                        if (!dup) seen[bk][cnt[bk]++] = unit;
                        any = true;
                    }
                    int w = 0;
                    for (int bk = 0; bk < 8; bk++) w = std::max(w, cnt[bk]);
                    cost += any ? w : 0;
                }
        return cost;
    };
    tf.floor = 0;
    for (int i = 0; i < R; i++) tf.floor += (long)tf.caps[i] * ((nlanes + 7) / 8 > 4 ? 4 : (nlanes + 7) / 8);
    // candidate strides: row-major (S >= groups) and column-major (Cg >= n) families with a little padding
    std::vector<std::pair<int, int>> strides;
    for (int pad = 0; pad <= 8; pad++) {
        if (injective(groups + pad, 1)) strides.push_back({groups + pad, 1});
        if (injective(1, n + pad)) strides.push_back({1, n + pad});
    }
    if (strides.empty()) return false;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto rnd = [&](int m) {
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        return (int)((rng >> 33) % (uint64_t)m);
    };
    State best = cur;
    long best_cost = -1;
    for (size_t sidx = 0; sidx < strides.size() && best_cost != tf.floor; sidx++) {
        State st = cur;
        st.S = strides[sidx].first;
        st.Cg = strides[sidx].second;
        long c = cost_of(st);
        if (best_cost < 0 || c < best_cost) { best = st; best_cost = c; }
        double temp = 2.0;
        for (int it = 0; it < 4000 && best_cost != tf.floor; it++, temp *= 0.9985) {
            State nx = st;
            const int mv = rnd(3);
            if (mv == 0) {                       // swap two stage positions
                const int x = rnd(n), y = rnd(n);
                std::swap(nx.pos[(size_t)x], nx.pos[(size_t)y]);
            } else if (mv == 1) {                // swap two rows inside a tier (equal capacities: any two of the tier)
                const int i = rnd(R), x = rnd(nb), y = rnd(nb);
                std::swap(nx.rows[(size_t)(i * nb + x)], nx.rows[(size_t)(i * nb + y)]);
            } else {                             // swap two entries of a row
                const int r = rnd(n);
                auto &e = nx.ent[(size_t)r];
                if (e.size() >= 2) std::swap(e[(size_t)rnd((int)e.size())], e[(size_t)rnd((int)e.size())]);
            }
            const long cn = cost_of(nx);
            // accept improvements always, deteriorations with a temperature-dependent chance (integer arithmetic: seeded, portable)
            const long d = cn - c;
            bool take = d <= 0;
            if (!take && temp > 0.05) take = rnd(1000) < (int)(1000.0 * std::exp(-(double)d / temp));
            if (take) {
                st = nx;
                c = cn;
                if (c < best_cost) { best = st; best_cost = c; }
            }
        }
    }
    tf.cost = best_cost;
    tf.S = best.S;
    tf.Cg = best.Cg;
    tf.perm.assign((size_t)n, 0);
    tf.pos.assign((size_t)n, 0);
    tf.ell.assign((size_t)tf.nnz * n, 0);
    for (int i = 0; i < R; i++)
        for (int b = 0; b < nb; b++) {
            const int v = R * b + i, r = best.rows[(size_t)(i * nb + b)];
            tf.perm[(size_t)v] = r;
            tf.pos[(size_t)v] = best.pos[(size_t)r];
            for (int j = 0; j < tf.nnz; j++) tf.ell[(size_t)j * n + v] = (uint32_t)best.pos[(size_t)r] | (0xFFFFu << 16);
            tf.ell[(size_t)v] = (uint32_t)best.pos[(size_t)r] | (g.ell[(size_t)r] & 0xFFFF0000u);   // diagonal: own row, its slot
            const auto &e = best.ent[(size_t)r];
            for (size_t j = 0; j < e.size(); j++)
                tf.ell[(1 + j) * (size_t)n + v] = (uint32_t)best.pos[(size_t)(e[j] & 0xFFFFu)] | (e[j] & 0xFFFF0000u);
        }
    tf.ok = true;
    return true;
}

// Additively split rows (value mode 7, Kern::AS).  With idle lanes in a grouped warp (groups * n < 32) the longest
// rows of the generator are carried as two table rows whose states add up to the physical entry: each part keeps the
// diagonal term on itself and half of the row's couplings, and every row that couples to a split row reads both parts
// (one more entry).  A row is split while that lowers the longest row of the table.  The form holds a generator copy
// over the n_v table rows (for the layout search), the table-row -> state-row map and the per-lane partner info.
struct AddSplitForm {
    bool ok = false;
    int n_v = 0, groups = 0, nnz = 0;
    Generator gv;                      // n = n_v, ell over table rows, everything else as the physical generator
    std::vector<int> phys;             // table row -> state row
    std::vector<uint8_t> lane_info;    // [32]: partner lane | 32 helper | 64 owner (as LaunchArgs::split_info)
};

static bool build_addsplit(const Generator &g, int ncols, AddSplitForm &af) {
    const int n = g.n;
    if (!g.imag_only) return false;
    std::vector<std::vector<uint32_t>> off((size_t)n);
    for (int r = 0; r < n; r++)
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) == 0xFFFFu) continue;
            if ((int)(e >> 16) < g.n_dyn_slots) return false;   // time-dependent off-diagonal entry
            off[(size_t)r].push_back(e);
        }
    // greedy: split the longest unsplit row while lanes remain and the longest table row gets shorter
    std::vector<int> split;   // physical rows carried as two parts
    auto lengths = [&](const std::vector<int> &sp, std::vector<int> *len_out) {
        int longest = 0;
        std::vector<int> len((size_t)n, 0);
        for (int r = 0; r < n; r++) {
            int L = 0;
            for (uint32_t e : off[(size_t)r])
                L += std::find(sp.begin(), sp.end(), (int)(e & 0xFFFFu)) != sp.end() ? 2 : 1;
            len[(size_t)r] = std::find(sp.begin(), sp.end(), r) != sp.end() ? (L + 1) / 2 : L;
            longest = std::max(longest, len[(size_t)r]);
        }
        if (len_out) *len_out = len;
        return longest;
    };
    int best = lengths(split, nullptr);
    for (;;) {
        const int n_v = n + (int)split.size() + 1;
        const int groups = std::min(32 / n_v, ncols);
        if (groups < 1 || groups * (ncols > 1 ? 3 : 1) < ncols) break;   // the whole propagator must stay in one warp
        std::vector<int> len;
        lengths(split, &len);
        int cand = -1;
        for (int r = 0; r < n; r++)
            if (std::find(split.begin(), split.end(), r) == split.end() && (cand < 0 || len[(size_t)r] > len[(size_t)cand])) cand = r;
        if (cand < 0) break;
        std::vector<int> trial = split;
        trial.push_back(cand);
        const int L = lengths(trial, nullptr);
        if (L >= best) break;
        best = L;
        split = trial;
    }
    if (split.empty()) return false;
    const int n_v = n + (int)split.size();
    af.n_v = n_v;
    af.groups = std::min(32 / n_v, ncols);
    af.nnz = 1 + best;
    af.phys.assign((size_t)n_v, 0);
    std::vector<int> helper_of((size_t)n, -1);
    for (int r = 0; r < n; r++) af.phys[(size_t)r] = r;
    for (size_t k = 0; k < split.size(); k++) {
        helper_of[(size_t)split[k]] = n + (int)k;
        af.phys[(size_t)n + k] = split[k];
    }
    af.gv = g;
    af.gv.n = n_v;
    af.gv.max_nnz = af.nnz;
    af.gv.ell.assign((size_t)af.nnz * n_v, 0);
    af.gv.row_len.assign((size_t)n_v, 0);
    for (int v = 0; v < n_v; v++)
        for (int j = 0; j < af.nnz; j++) af.gv.ell[(size_t)j * n_v + v] = (uint32_t)v | (0xFFFFu << 16);
    for (int r = 0; r < n; r++) {
        std::vector<uint32_t> ex;   // expanded entries: a split source appears as both of its parts
        for (uint32_t e : off[(size_t)r]) {
            const int src = (int)(e & 0xFFFFu);
            ex.push_back(e);
            if (helper_of[(size_t)src] >= 0) ex.push_back((uint32_t)helper_of[(size_t)src] | (e & 0xFFFF0000u));
        }
        const uint32_t dslot = g.ell[(size_t)r] & 0xFFFF0000u;
        const int h = helper_of[(size_t)r];
        const size_t first = h >= 0 ? (ex.size() + 1) / 2 : ex.size();
        af.gv.ell[(size_t)r] = (uint32_t)r | dslot;
        for (size_t j = 0; j < first; j++) af.gv.ell[(1 + j) * (size_t)n_v + r] = ex[j];
        af.gv.row_len[(size_t)r] = (uint8_t)(1 + first);
        if (h >= 0) {
            af.gv.ell[(size_t)h] = (uint32_t)h | dslot;   // the diagonal term acts on each part
            for (size_t j = first; j < ex.size(); j++) af.gv.ell[(1 + j - first) * (size_t)n_v + h] = ex[j];
            af.gv.row_len[(size_t)h] = (uint8_t)(1 + ex.size() - first);
        }
    }
    af.lane_info.assign(32, 0);
    for (int l = 0; l < 32; l++) af.lane_info[(size_t)l] = (uint8_t)l;
    for (int gi = 0; gi < af.groups; gi++)
        for (size_t k = 0; k < split.size(); k++) {
            const int lo = gi * n_v + split[k], lh = gi * n_v + n + (int)k;
            af.lane_info[(size_t)lo] = (uint8_t)(lh | 64);
            af.lane_info[(size_t)lh] = (uint8_t)(lo | 32);
        }
    af.ok = true;
    return true;
}

// Shared-memory wavefronts of the stage-vector gathers and stores for a candidate layout: the stage
// buffer holds element (row, group g, column c) at 16-byte unit row*S + g*Cg + c.  A 128-bit shared
// access is served a quarter warp (8 lanes) at a time; lanes hitting the same 16-byte bank group
// (unit mod 8) at different addresses serialise.
// blk_rows > 0: blocked variants (lane = block row + blk_rows * group, rows rpl*block_row + i) with their own
// row table `ell` of `max_nnz` entries per row.
static long layout_cost(const Generator &g, int rpl, int cpl, int groups, int S, int Cg, int blk_rows = 0,
                        const uint32_t *ell = nullptr, int max_nnz = 0, int table_rows = 0) {
    const int n = g.n;
    if (!ell) {
        ell = g.ell.data();
        max_nnz = g.max_nnz;
    }
    if (table_rows <= 0) table_rows = n;  // (split rows: the table also holds the helper lanes' rows n..table_rows-1)
    long cost = 0;
    for (int i = 0; i < rpl; i++) {
        for (int j = 0; j <= max_nnz; j++) {  // j == max_nnz: the lane's own store
            for (int q = 0; q < 4; q++) {
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int seen[8][8];
                for (int l = 8 * q; l < 8 * q + 8; l++) {
                    int gi = 0, r = l + 32 * i;
                    bool active = true;
                    if (blk_rows > 0) {
                        active = l < groups * blk_rows;
                        const int le = active ? l : groups * blk_rows - 1;
                        gi = le / blk_rows;
                        r = rpl * (le % blk_rows) + i;
                    } else if (groups > 1) {
                        const int nl = groups * n, nh = table_rows - n;   // split rows: groups * nh helper lanes follow
                        active = l < nl + groups * nh;
                        if (l >= nl && active) {
                            gi = (l - nl) / nh;
                            r = n + (l - nl) % nh;
                        } else {
                            const int le = active ? l : nl - 1;
                            gi = le / n;
                            r = le % n;
                        }
                    }
                    const bool helper = active && r >= n && r < table_rows;
                    if (r >= n && !helper) {
                        active = false;
                        r = n - 1;
                    }
                    int src = r;
                    if (j == max_nnz) {
                        if (!active || helper) continue;  // inactive lanes and helpers do not store
                    } else if (j == 0) {
                        continue;  // diagonal operand comes from registers
                    } else {
                        src = (int)(ell[(size_t)j * table_rows + r] & 0xFFFFu);
                    }
                    const int unit = src * S + gi * Cg;
                    const int b = unit & 7;
                    bool dup = false;
                    for (int k = 0; k < cnt[b]; k++) dup = dup || seen[b][k] == unit;
                    if (!dup) seen[b][cnt[b]++] = unit;
                }
                int w = 0;
                for (int b = 0; b < 8; b++) w = std::max(w, cnt[b]);
                cost += (long)w * cpl;
            }
        }
    }
    return cost;
}

// Split rows: with one row per lane and fewer than 32 rows, cut the longest rows in two and give the second half to
// an idle lane.  Chooses the smallest row length M such that every row longer than M can get a helper (at most
// 32 - n of them, each row at most 2 M long), and rebuilds the row table with n + helpers rows and 1 + M entries
// per row (diagonal first; helpers carry a zero diagonal).  Returns false when nothing is gained.
static bool build_split_rows(const Generator &g, Plan &pl, int groups = 1) {
    const int n = g.n, spare = (32 - groups * n) / groups;   // idle lanes per column group
    if (spare <= 0 || g.max_nnz < 3) return false;
    std::vector<std::vector<uint32_t>> off((size_t)n);
    int longest = 0;
    for (int r = 0; r < n; r++) {
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) != 0xFFFFu) off[(size_t)r].push_back(e);
        }
        longest = std::max(longest, (int)off[(size_t)r].size());
    }
    int M = longest;
    for (int m = (longest + 1) / 2; m < longest; m++) {
        int need = 0;
        for (int r = 0; r < n; r++) need += (int)off[(size_t)r].size() > m ? 1 : 0;
        if (need <= spare) {
            M = m;
            break;
        }
    }
    if (M >= longest) return false;
    std::vector<int> helper_of((size_t)n, -1);
    int nv = n;
    for (int r = 0; r < n; r++)
        if ((int)off[(size_t)r].size() > M) helper_of[(size_t)r] = nv++;
    pl.split_nv = nv;
    pl.split_nnz = 1 + M;
    pl.split_ell.assign((size_t)pl.split_nnz * nv, 0);
    pl.split_info.assign(32, 0);
    for (int l = 0; l < 32; l++) pl.split_info[(size_t)l] = (uint8_t)l;
    for (int v = 0; v < nv; v++) {
        // padding: own (owner's) row, zero slot
        for (int j = 0; j < pl.split_nnz; j++) pl.split_ell[(size_t)j * nv + v] = (uint32_t)std::min(v, n - 1) | (0xFFFFu << 16);
    }
    for (int r = 0; r < n; r++) {
        pl.split_ell[(size_t)r] = g.ell[(size_t)r];  // diagonal
        const auto &o = off[(size_t)r];
        const int h = helper_of[(size_t)r];
        const int first = h >= 0 ? ((int)o.size() + 1) / 2 : (int)o.size();
        for (int j = 0; j < first; j++) pl.split_ell[(size_t)(1 + j) * nv + r] = o[(size_t)j];
        if (h >= 0) {
            for (int j = first; j < (int)o.size(); j++) pl.split_ell[(size_t)(1 + j - first) * nv + h] = o[(size_t)j];
            for (int j = 0; j < pl.split_nnz; j++)   // helper's padding reads the owner's row (broadcast-friendly)
                if ((pl.split_ell[(size_t)j * nv + h] >> 16) == 0xFFFFu) pl.split_ell[(size_t)j * nv + h] = (uint32_t)r | (0xFFFFu << 16);
            // per-lane info: partner lane, helper / owner flags (lanes: group gi's rows at gi * n, its helpers after all rows)
            const int nh = nv - n;
            for (int gi = 0; gi < groups; gi++) {
                const int lo = groups > 1 ? gi * n + r : r, lh = groups > 1 ? groups * n + gi * nh + (h - n) : h;
                pl.split_info[(size_t)lo] = (uint8_t)(lh | 64);
                pl.split_info[(size_t)lh] = (uint8_t)(lo | 32);
            }
        }
    }
    return true;
}

static const KernelVariant *find_variant(int rpl, int cpl, int max_nnz, int block, int vm, int rs, int lean = 0) {
    const KernelVariant *(*lists[])(int *) = {variants_r1c1_t128, variants_r1c3_t128, variants_r1c3_t256,
                                              variants_r1c4_t256, variants_r2c1_t128,
                                              variants_r2c2_t256, variants_r3c1_t128, variants_r3c1_t256,
                                              variants_rowsplit, variants_colperlane, variants_blocked, variants_lean,
                                              variants_lean_t256, variants_lean_c2, variants_lean_c2_t256};
    // (development: QSIM_MAX_MINB caps the occupancy target of the chosen variant, e.g. 3 keeps the 168-register lean kernels)
    const int max_minb = std::getenv("QSIM_MAX_MINB") ? std::atoi(std::getenv("QSIM_MAX_MINB")) : 99;
    const KernelVariant *best = nullptr;
    for (auto fn : lists) {
        int cnt = 0;
        const KernelVariant *v = fn(&cnt);
        for (int i = 0; i < cnt; i++)
            if (v[i].rpl == rpl && v[i].cpl == cpl && (rs || v[i].maxnnz >= max_nnz) && v[i].maxt >= block &&
                v[i].vm == vm && v[i].rs == rs && v[i].lean == lean && v[i].minb <= max_minb) {
                // smallest row capacity, then smallest block bound, then the highest occupancy target
                auto key = [&](const KernelVariant &k) { return std::make_tuple(k.maxnnz, k.maxt, -k.minb); };
                if (!best || key(v[i]) < key(*best)) best = &v[i];
            }
    }
    return best;
}

// Row tables of the row-split kernels.  Rows are sorted by descending length and dealt out in chunks of 32
// (chunk c = slot * team_warps + warp), so each (warp, slot) group runs exactly as many entries as its longest row
// and the warps of a team carry equal work; without this every row would be padded to the longest row of the
// matrix (12 entries against an average of 6 for the d^2 = 729 superoperator).  Entries are translated to byte
// offsets into the stage buffer (indexed by sorted position) and the value table.  Compact layout: chunk c owns
// table rows [off[c], off[c] + len[c]) of 32 entries (one per lane); len[c] = 1 (diagonal) + the chunk's longest
// off-diagonal count rounded up to even; two more zero rows follow each chunk so the kernel's pairwise-pipelined
// gather may read one pair ahead.  Lanes past the last row and short rows read zero-slot entries.
// full == false: only the chunk offsets / lengths and the table size (make_plan runs on every call); full == true: also
// the entries, with the rows' positions inside the stage buffer chosen by anneal_colors() so that the sources a
// quarter warp gathers together fall into distinct bank groups (rs_pos[q] = position of sorted row q; the stage
// vector is the only thing stored in that order -- scratch, prefetch buffers and the table itself stay in sorted order).
static void build_rowsplit_table(const Generator &g, Plan &pl, bool full) {
    const int n = g.n, n_chunks = pl.rpl * pl.team_warps;
    const uint32_t vb = pl.vm >= 1 ? 8u : 16u;
    std::vector<std::vector<uint32_t>> compact;
    std::vector<int> perm;
    rowsplit_rows(g, compact, perm);
    if (pl.cu) {
        // Column-per-lane mapping: chunk (slot, warp) is ONE row, the same in all 32 lanes (the lane picks the
        // column).  Rows go to warps longest first, each to the warp with the fewest entries so far that still has a
        // free slot, so the warps of a team carry (nearly) equal work; within a warp the slots are filled in that
        // order.  The stage buffer keeps the state's own row order: row r at 16 * r * y_row_stride (+ 16 * column).
        const int W = pl.team_warps;
        std::vector<int> load((size_t)W, 0), used((size_t)W, 0), row_of_chunk((size_t)n_chunks, -1);
        for (int q = 0; q < n; q++) {
            int best = -1;
            for (int w = 0; w < W; w++)
                if (used[(size_t)w] < pl.rpl && (best < 0 || load[(size_t)w] < load[(size_t)best])) best = w;
            row_of_chunk[(size_t)(used[(size_t)best] * W + best)] = perm[(size_t)q];
            used[(size_t)best]++;
            load[(size_t)best] += (int)compact[(size_t)perm[(size_t)q]].size();
        }
        pl.rs_off.assign((size_t)n_chunks, 0);
        pl.rs_len.assign((size_t)n_chunks, 1);
        size_t total = 0;
        for (int c = 0; c < n_chunks; c++) {
            const int longest = row_of_chunk[(size_t)c] >= 0 ? (int)compact[(size_t)row_of_chunk[(size_t)c]].size() : 1;
            pl.rs_off[(size_t)c] = (int)(total / 32);
            pl.rs_len[(size_t)c] = 1 + 2 * (longest / 2);
            total += (size_t)32 * (pl.rs_len[(size_t)c] + 2);
        }
        pl.ell_bytes = (int)(total * sizeof(uint32_t));
        if (!full) return;
        pl.rs_perm.assign((size_t)n_chunks * 32, -1);
        pl.rs_pos.assign((size_t)n_chunks * 32, 0);
        const uint32_t zero = (vb * (uint32_t)g.n_slots) << 16;   // row 0, zero slot
        pl.rs_table.assign(total, zero);
        for (int c = 0; c < n_chunks; c++) {
            const int r = row_of_chunk[(size_t)c];
            if (r < 0) continue;
            uint32_t *tab = pl.rs_table.data() + (size_t)32 * pl.rs_off[(size_t)c];
            const auto &row = compact[(size_t)r];
            for (int l = 0; l < 32; l++) {
                pl.rs_perm[(size_t)c * 32 + l] = r;
                pl.rs_pos[(size_t)c * 32 + l] = r;
                for (size_t j = 0; j < row.size(); j++) {
                    const uint32_t sl = (row[j] >> 16) == 0xFFFFu ? (uint32_t)g.n_slots : (row[j] >> 16);
                    uint32_t e = (16u * (row[j] & 0xFFFFu) * (uint32_t)pl.y_row_stride) | ((vb * sl) << 16);
                    if (pl.vm == 4 && j > 0 && sl < (uint32_t)g.n_slots && pl.slot_imag[sl])
                        e = (e + ((8u * (uint32_t)pl.nvp) << 16)) | 0x80000000u;
                    tab[j * 32 + (size_t)l] = e;
                }
            }
        }
        return;
    }
    pl.rs_perm.assign((size_t)n_chunks * 32, -1);
    for (int q = 0; q < n; q++) pl.rs_perm[(size_t)q] = perm[(size_t)q];
    pl.rs_off.assign((size_t)n_chunks, 0);
    pl.rs_len.assign((size_t)n_chunks, 1);
    size_t total = 0;
    for (int c = 0; c < n_chunks; c++) {
        int longest = 1;
        if (c * 32 < n) longest = (int)compact[(size_t)perm[(size_t)c * 32]].size();
        pl.rs_off[(size_t)c] = (int)(total / 32);
        pl.rs_len[(size_t)c] = 1 + 2 * (longest / 2);  // 1 + (longest - 1) rounded up to even
        total += (size_t)32 * (pl.rs_len[(size_t)c] + 2);
    }
    pl.ell_bytes = (int)(total * sizeof(uint32_t));
    if (!full) return;

    // positions inside the stage buffer: pos[r] for state row r
    std::vector<int> pos((size_t)n);
    {
        std::vector<int> color((size_t)n), cls((size_t)n, 0);
        for (int q = 0; q < n; q++) color[(size_t)perm[(size_t)q]] = q % 8;
        const long iters = anneal_iters(n);
        if (iters > 0 && pl.y_row_stride % 2 == 1) {
            std::vector<std::vector<int>> groups;
            rowsplit_groups(g, compact, perm, 0, groups);
            const long before = layout_groups_cost(groups, color, 8);
            anneal_colors(groups, cls, 8, iters, color);
            if (std::getenv("QSIM_DEBUG"))
                std::fprintf(stderr, "[qsim] row-split stage layout: %zu quarter-warp accesses, wavefronts %ld -> %ld\n",
                             groups.size(), before, layout_groups_cost(groups, color, 8));
        }
        std::vector<std::vector<int>> free_pos(8);
        for (int p = n - 1; p >= 0; p--) free_pos[(size_t)(p % 8)].push_back(p);
        for (int q = 0; q < n; q++) {
            auto &v = free_pos[(size_t)color[(size_t)perm[(size_t)q]]];
            pos[(size_t)perm[(size_t)q]] = v.back();
            v.pop_back();
        }
    }
    pl.rs_pos.assign((size_t)n_chunks * 32, pos[(size_t)perm[(size_t)n - 1]]);
    for (int q = 0; q < n; q++) pl.rs_pos[(size_t)q] = pos[(size_t)perm[(size_t)q]];
    const uint32_t zero = (16u * (uint32_t)pos[(size_t)perm[(size_t)n - 1]] * (uint32_t)pl.y_row_stride) | ((vb * (uint32_t)g.n_slots) << 16);
    pl.rs_table.assign(total, zero);
    for (int c = 0; c < n_chunks; c++) {
        uint32_t *tab = pl.rs_table.data() + (size_t)32 * pl.rs_off[(size_t)c];
        for (int l = 0; l < 32 && c * 32 + l < n; l++) {
            const auto &row = compact[(size_t)perm[(size_t)c * 32 + l]];
            for (size_t j = 0; j < row.size(); j++) {
                const uint32_t src = (uint32_t)pos[(size_t)(row[j] & 0xFFFFu)];
                const uint32_t sl = (row[j] >> 16) == 0xFFFFu ? (uint32_t)g.n_slots : (row[j] >> 16);
                uint32_t e = (16u * src * (uint32_t)pl.y_row_stride) | ((vb * sl) << 16);
                // value mode 4: an imaginary off-diagonal entry reads the second half of the split value table
                if (pl.vm == 4 && j > 0 && sl < (uint32_t)g.n_slots && pl.slot_imag[sl])
                    e = (e + ((8u * (uint32_t)pl.nvp) << 16)) | 0x80000000u;
                tab[j * 32 + (size_t)l] = e;
            }
        }
    }
}

// which: 0 unitary_propagator, 1 unitary_state, 2 me_propagator, 3 me_state
// Row tables of a row-split component form: every component's own table (build_rowsplit_table on the component's
// generator: rows sorted by length in chunks of 32, sources as positions inside the unit, annealed against bank
// conflicts), concatenated; plus the empty chunk.  rs_perm holds state rows.
static void build_rowsplit_component_tables(const ComponentForm &cf, Plan &pl, bool full) {
    const uint32_t vb = pl.vm >= 1 ? 8u : 16u;
    pl.rs_table.clear();
    pl.rs_perm.clear();
    pl.rs_pos.clear();
    pl.rs_off.clear();
    pl.rs_len.clear();
    size_t total = 0;   // table entries so far
    for (size_t k = 0; k < cf.subs.size(); k++) {
        const Generator &sub = cf.subs[k];
        Plan q;
        q.rpl = 1;
        q.team_warps = (sub.n + 31) / 32;
        q.vm = pl.vm;
        q.nvp = pl.nvp;
        q.slot_imag = pl.slot_imag;
        q.y_row_stride = pl.y_row_stride;
        build_rowsplit_table(sub, q, full);
        for (int c = 0; c < q.team_warps; c++) {
            pl.rs_off.push_back(q.rs_off[(size_t)c] + (int)(total / 32));
            pl.rs_len.push_back(q.rs_len[(size_t)c]);
        }
        if (full) {
            pl.rs_table.insert(pl.rs_table.end(), q.rs_table.begin(), q.rs_table.end());
            for (int v : q.rs_perm) pl.rs_perm.push_back(v >= 0 ? cf.comp_rows[k][(size_t)v] : -1);
            pl.rs_pos.insert(pl.rs_pos.end(), q.rs_pos.begin(), q.rs_pos.end());
        }
        total += (size_t)q.ell_bytes / sizeof(uint32_t);
    }
    // the empty chunk: no rows, a zero diagonal and the two look-ahead rows
    pl.rs_off.push_back((int)(total / 32));
    pl.rs_len.push_back(1);
    if (full) {
        const uint32_t zero = (vb * (uint32_t)cf.gv.n_slots) << 16;   // position 0, zero slot
        pl.rs_table.insert(pl.rs_table.end(), (size_t)32 * 3, zero);
        pl.rs_perm.insert(pl.rs_perm.end(), 32, -1);
        pl.rs_pos.insert(pl.rs_pos.end(), 32, 0);
    }
    total += 32 * 3;
    pl.ell_bytes = (int)(total * sizeof(uint32_t));
    if (!full) return;
    // per (pass, chunk, lane) records for the kernel (LaunchArgs::rs_pass)
    const size_t n_rec = cf.pass.size() / 4;
    pl.rs_lane.assign(n_rec * 32 * 4, 0);
    for (size_t c = 0; c < n_rec; c++) {
        const int cc = cf.pass[4 * c], ubase = cf.pass[4 * c + 1], col = cf.pass[4 * c + 2];
        for (int l = 0; l < 32; l++) {
            const int q = cc * 32 + l, row = pl.rs_perm[(size_t)q];
            int *rec = &pl.rs_lane[(c * 32 + (size_t)l) * 4];
            rec[0] = row >= 0 ? (row | (col << 16)) : -1;
            rec[1] = 16 * (ubase + pl.rs_pos[(size_t)q]) * pl.y_row_stride;
            rec[2] = 4 * (32 * pl.rs_off[(size_t)cc] + l);
            rec[3] = pl.rs_len[(size_t)cc] | ((16 * ubase * pl.y_row_stride) << 8);
        }
    }
}

// every off-diagonal entry of the generator is a per-trajectory constant (only diagonal slots depend on time)
static bool offdiag_static(const Generator &g) {
    for (int j = 1; j < g.max_nnz; j++)
        for (int r = 0; r < g.n; r++) {
            const uint32_t sl = g.ell[(size_t)j * g.n + r] >> 16;
            if (sl != 0xFFFFu && (int)sl < g.n_dyn_slots) return false;
        }
    return true;
}

// Component pruning of a propagator problem: the kernel shape (rows x columns per lane), the packing and the kernel
// variant with the best estimated throughput, or nothing when the generator is connected.  The estimate: a Tsit5 step
// of a warp costs about 2200 + 650 E cycles (E = entries per lane; measured on cfg2) plus 340 per
// further warp of the team (barriers; fitted to cfg3: five warps x three columns 441 ms, four x four 470 ms, seven x two
// 484 ms per 1184 x 50 ns); one-warp teams run four per CTA and as many CTAs per SM as the variant's register target
// allows, multi-warp resident teams share a CTA of up to eight warps (two CTAs per SM where the variant asks for it),
// streamed teams run one per SM with their batches in sequence.
static void choose_component_form(const Generator &gp, std::shared_ptr<ComponentForm> &out) {
    const bool no_lean = std::getenv("QSIM_NO_LEAN") != nullptr;
    const bool stat = gp.imag_only && offdiag_static(gp);
    std::vector<uint8_t> slot_imag;
    const bool mixed = !gp.imag_only && generator_is_mixed(gp, slot_imag);
    double best = 0.0;
    for (int rpl = 1; rpl <= 3; rpl++)
        for (int cpl = 1; cpl <= (rpl == 1 ? 4 : rpl == 2 ? 2 : 1); cpl++) {
            auto cf = std::make_shared<ComponentForm>();
            if (!build_component_form(gp, rpl, cpl, 8, *cf)) continue;
            const int W = cf->team_warps, E = rpl * cpl;
            const int tpc = cf->streamed ? 1 : (W == 1 ? 4 : 8 / W);
            const int block = W * 32 * tpc;
            // the variant make_plan will pick for this shape
            int vm = gp.imag_only ? 1 : 0;
            const bool lean_ok = !cf->streamed && !no_lean;
            if (vm == 1 && stat && ((E <= 3 && find_variant(rpl, cpl, gp.max_nnz, block, 2, 0, 0)) ||
                                    (lean_ok && find_variant(rpl, cpl, gp.max_nnz, block, 2, 0, 1))))
                vm = 2;
            if (vm == 0 && mixed && find_variant(rpl, cpl, gp.max_nnz, block, 4, 0, 0)) vm = 4;
            const KernelVariant *lv = lean_ok ? find_variant(rpl, cpl, gp.max_nnz, block, vm, 0, 1) : nullptr;
            const KernelVariant *kv = lv ? lv : find_variant(rpl, cpl, gp.max_nnz, block, vm, 0, 0);
            if (!kv) continue;
            double score;
            if (cf->streamed)
                score = 1.0 / (2200.0 + (double)(cf->n_batches / W) * (650.0 * E + 400.0));
            else {
                // teams per SM: one-warp teams, four per CTA, as many CTAs as the variant's register target allows;
                // multi-warp teams share a CTA of up to eight warps, two CTAs per SM where the variant asks for it
                const int teams_per_sm = W == 1 ? 4 * kv->minb : kv->minb * tpc;
                score = (double)teams_per_sm / (2200.0 + 650.0 * E + 340.0 * (W - 1));   // (+ team barriers; measured on cfg3)
            }
            if (const char *force = std::getenv("QSIM_CF_SHAPE"))   // development: "r1c4" forces a shape
                if (std::string(force) == "r" + std::to_string(rpl) + "c" + std::to_string(cpl)) score = 1e9;
            if (std::getenv("QSIM_DEBUG"))
                std::fprintf(stderr, "[qsim] component form r%dc%d: %d components, %d warps%s, %d batches, %d local rows, %s: score %.3g\n",
                             rpl, cpl, cf->n_comp, W, cf->streamed ? " (streamed)" : "", cf->n_batches, cf->n_loc, kv->name, score * 1e3);
            if (score > best) {
                best = score;
                out = cf;
            }
        }
}

static int make_plan(qsim_problem *p, int which, Plan &pl, const ComponentForm *state_cf = nullptr, bool allow_prune = true) {
    pl.w = which >= 2 ? 1 : 0;
    if (int rc = ensure_generator(p, pl.w)) return rc;
    const Generator &gp = p->gen[pl.w];
    pl.identity = (which == 0 || which == 2) ? 1 : 0;
    pl.ncols = pl.identity ? gp.n : 1;
    // Additively split rows (value mode 7) for small propagators that leave lanes idle in their warp (the 2-transmon
    // gates: 9 rows x 3 column groups on 27 lanes; split, 10 table rows of 1 + 2 entries on 30 lanes instead of rows
    // padded to 1 + 4).  From here on the plan is made for the generator over the table rows; only the number of
    // columns, the results and the norms refer to the physical rows.  Opt-in (QSIM_ADDSPLIT=1): parity-green, 40 % fewer
    // generator multiply-adds and half the operand loads per stage, and still 2 % slower than the padded lean kernel
    // (cfg2: 566 ms against 555 ms) -- the shuffles that combine the parts for the error norm sit on the step's chain.
    if (pl.identity && gp.n <= 15 && gp.imag_only && std::getenv("QSIM_ADDSPLIT") && !std::getenv("QSIM_NO_LEAN") &&
        !std::getenv("QSIM_BLOCKED") && !std::getenv("QSIM_TIERED") && !std::getenv("QSIM_SPLIT")) {
        std::lock_guard<std::mutex> lk(p->plan_mutex);
        if (!p->addsplit[pl.w]) {
            p->addsplit[pl.w] = std::make_shared<AddSplitForm>();
            build_addsplit(gp, pl.ncols, *p->addsplit[pl.w]);
        }
        const AddSplitForm *af = p->addsplit[pl.w].get();
        if (af->ok && find_variant(1, 3, af->nnz, 128, 7, 0, 1)) pl.af = af;
    }
    // Component pruning (propagators): integrate only the blocks of the connected components of the generator's graph
    // (qsim_components.cpp).  QSIM_NO_PRUNE=1 keeps the full state (development A/B switch); the opt-in experiments
    // above and below run on the full state.
    if (state_cf) {
        pl.cf = state_cf;   // state evolution on the components the initial state touches (solve_common, state_form_for)
    } else if (pl.identity && allow_prune && !pl.af && !std::getenv("QSIM_NO_PRUNE") && !std::getenv("QSIM_BLOCKED") && !std::getenv("QSIM_TIERED") &&
        !std::getenv("QSIM_SPLIT") && !std::getenv("QSIM_SPLIT_ROWS") && !std::getenv("QSIM_COLPERLANE") && !std::getenv("QSIM_LEAN9") &&
        gp.n <= 9 * 96) {
        std::lock_guard<std::mutex> lk(p->plan_mutex);
        if (!p->cform_tried[pl.w]) {
            p->cform_tried[pl.w] = true;
            if (gp.n <= 96 && gp.max_nnz <= 16)
                choose_component_form(gp, p->cform[pl.w]);
            else {
                // row-split kernels: passes over several units at once, eight-warp teams with three rows per lane.
                // (Sixteen-warp teams with one row per lane at 128 registers -- 16 chunks per pass instead of 24, twice
                // the warps per SM -- were measured parity-green and 2.5 % slower on cfg5, 115.1 ms against 112.2 ms per
                // 592 x 2 ns: a pass costs about 1500 cycles per stage whatever a warp's share of the rows is.  Their
                // instantiations were dropped again; build_rowsplit_component_form still takes the rows per lane.)
                auto cf = std::make_shared<ComponentForm>();
                const bool ok = build_rowsplit_component_form(gp, 8, *cf);
                if (ok) p->cform[pl.w] = cf;
            }
        }
        if (p->cform[pl.w] && p->cform[pl.w]->ok) pl.cf = p->cform[pl.w].get();
    }
    const Generator &g = pl.af ? pl.af->gv : pl.cf ? pl.cf->gv : gp;
    const Generator &gs = pl.cf ? gp : g;   // the generator whose row table make_plan scans (a component form's table is per warp)
    const int n = g.n;
    int rpl, cpl;
    if (pl.cf) {
        rpl = pl.cf->rpl;
        cpl = pl.cf->cpl;
        pl.groups = 1;
        pl.rowsplit = pl.cf->rs;
    } else if (n <= 16) {
        rpl = 1;
        pl.groups = std::min(32 / n, pl.ncols);
        cpl = (pl.ncols + pl.groups - 1) / pl.groups > 1 ? 3 : 1;
    } else if (n <= 32 && pl.identity && std::getenv("QSIM_COLPERLANE")) {
        // Opt-in experiment (QSIM_COLPERLANE=1) for propagators with 17..32 rows: column-per-lane mapping of the
        // row-split machinery -- lane = column, four whole rows per warp, ceil(n/4) <= 8 warps, warp-uniform sparse
        // rows (LaunchArgs::cu_map).  Parity-green but measured 2x slower than the row-per-lane kernel on cfg3
        // (d = 27): half the shared-memory wavefronts, yet with one 7-warp team per SM the dynamic row loops and the
        // seven team-wide stage exchanges per step leave the step latency-bound (instruction-fetch stalls 38 %).
        rpl = 4;
        cpl = 1;
        pl.rowsplit = 1;
        pl.cu = 1;
        pl.groups = 32;
    } else if (n <= 32) {
        // three columns per lane up to 24 rows, four above: at most eight warps per trajectory.  Opt-in (QSIM_LEAN9=1):
        // 25..27 rows (the 3-transmon gates of cfg3) as nine-warp teams of the lean kernels with three columns per lane;
        // at the 168 registers nine warps leave each thread that variant spills and measures 5 % slower than the
        // seven-warp kernel on cfg3 (1.042 s against 0.994 s for 1184 x 50 ns).
        rpl = 1;
        const bool nine = n > 24 && n <= 27 && pl.ncols > 1 && g.imag_only && std::getenv("QSIM_LEAN9") && !std::getenv("QSIM_NO_LEAN") &&
                          find_variant(1, 3, g.max_nnz, 288, 1, 0, 1) != nullptr;
        cpl = pl.ncols > 1 ? ((n <= 24 || nine) ? 3 : 4) : 1;
    } else if (n <= 64) {
        rpl = 2;
        cpl = pl.ncols > 1 ? 2 : 1;
    } else if (n <= 96) {
        rpl = 3;
        cpl = 1;
    } else if (n <= 9 * 96) {
        // row split: one column at a time, its rows spread over ceil(n/96) warps
        rpl = 3;
        cpl = 1;
        pl.rowsplit = 1;
    } else {
        return fail(QSIM_ERR_UNSUPPORTED, "state vectors longer than 864 entries have no kernel yet");
    }
    if (!pl.rowsplit && g.max_nnz > 16) {
        // rows longer than the register-resident capacities (dense Hamiltonians of 17..96 levels): the row-split
        // kernels read their rows from global memory and take any row length
        rpl = 3;
        cpl = 1;
        pl.groups = 1;
        pl.rowsplit = 1;
    }
    int cpb = pl.groups * cpl;
    pl.n_batches = pl.cf ? pl.cf->n_batches : (pl.ncols + cpb - 1) / cpb;
    // Blocked variant (value mode 3) for small propagators whose couplings fall into 3 x 3 blocks (Kronecker
    // structure with a 3-level last subsystem: the 2-transmon gates of cfg1/cfg2): every lane owns 3 consecutive
    // rows of one column and each neighbour block is loaded once for its 9 multiply-adds, which cuts the
    // shared-memory wavefronts per step by a quarter.  Measured on B200 it is only 1.7 % faster than the
    // row-per-lane value-mode-2 kernel (the step is bound by its dependent chain, not by shared-memory
    // throughput) and its 18 block values per lane push it to 255 registers, so it is opt-in: QSIM_BLOCKED=1.
    bool blocked = false;
    if (pl.identity && !pl.rowsplit && n % 3 == 0 && std::getenv("QSIM_BLOCKED")) {
        const int nb = n / 3, gb = std::min(32 / nb, pl.ncols);
        if (gb >= pl.ncols) {  // the whole propagator in one warp
            for (int cap : {3, 5})
                if (!blocked && find_variant(3, 1, cap, 128, 3, 0) && build_blocked(g, 3, cap, pl)) blocked = true;
            if (blocked) {
                rpl = 3;
                cpl = 1;
                pl.groups = gb;
                cpb = gb;
                pl.n_batches = 1;
            }
        }
    }
    // Tiered variant (value mode 5) for small propagators that fit one warp with three rows of one column per
    // lane (the 2-transmon gates of cfg1/cfg2: 9 rows, tier capacities 5 / 3 / 2): 10 instead of 15 entries per lane
    // and stage, one complex operand per off-diagonal entry instead of three.  Opt-in (QSIM_TIERED=1): parity-green, but
    // at 168 registers its 10 row entries + 7 static values per lane spill, which costs more than the gathers save
    // (cfg2: 593 ms against 554 ms for the row-per-lane lean kernel; profiles/r2_cfg2_experiments.txt).
    const KernelVariant *tiered_kv = nullptr;
    if (pl.identity && !pl.rowsplit && !blocked && n % 3 == 0 && std::getenv("QSIM_TIERED") && !std::getenv("QSIM_NO_LEAN")) {
        const int nb = n / 3, gb = std::min(32 / std::max(nb, 1), pl.ncols);
        if (nb >= 1 && gb >= pl.ncols) {
            TieredForm *tf = nullptr;
            {
                std::lock_guard<std::mutex> lk(p->plan_mutex);
                if (!p->tiered[pl.w]) {
                    p->tiered[pl.w] = std::make_shared<TieredForm>();
                    build_tiered(g, gb, *p->tiered[pl.w]);
                    if (std::getenv("QSIM_DEBUG") && p->tiered[pl.w]->ok)
                        std::fprintf(stderr, "[qsim] tiered form: caps %d/%d/%d, stage layout S=%d Cg=%d, %ld wavefronts per stage (floor %ld)\n",
                                     p->tiered[pl.w]->caps[0], p->tiered[pl.w]->caps[1], p->tiered[pl.w]->caps[2], p->tiered[pl.w]->S,
                                     p->tiered[pl.w]->Cg, p->tiered[pl.w]->cost, p->tiered[pl.w]->floor);
                }
                tf = p->tiered[pl.w].get();
            }
            if (tf->ok) {
                int cnt = 0;
                const KernelVariant *v = variants_lean(&cnt);
                for (int i = 0; i < cnt; i++)
                    if (v[i].vm == 5 && v[i].rpl == 3 && (v[i].maxnnz >> 6) >= tf->caps[0] && ((v[i].maxnnz >> 3) & 7) >= tf->caps[1] &&
                        (v[i].maxnnz & 7) >= tf->caps[2] && !tiered_kv)
                        tiered_kv = &v[i];
            }
            if (tiered_kv) {
                pl.tiered = 1;
                pl.tf = tf;
                pl.blk_rows = tf->nb;
                pl.blk_nnz = tf->nnz;
                rpl = 3;
                cpl = 1;
                pl.groups = gb;
                cpb = gb;
                pl.n_batches = 1;
            }
        }
    }
    // Split rows (one row per lane, spare lanes): halves the longest rows; see build_split_rows.  Opt-in
    // (build with -DQSIM_SPLIT_ROWS, run with QSIM_SPLIT_ROWS=1): parity-green, but on cfg3 (d = 27: rows of 1 + 4 instead of 1 + 8 entries) the saved operand
    // loads are paid back by the per-stage shuffles of the helpers' partial sums and by bank conflicts of the compacted
    // rows (profiles/r2_cfg3_experiments.txt: 953 ms against 937 ms for 1184 x 50 ns).
    bool split = false;
#ifdef QSIM_SPLIT_ROWS   // experimental build (make EXTRA=-DQSIM_SPLIT_ROWS) + QSIM_SPLIT_ROWS=1 in the environment
    if (!pl.rowsplit && !blocked && !pl.tiered && rpl == 1 && pl.groups == 1 && n < 32 && std::getenv("QSIM_SPLIT_ROWS"))
        split = build_split_rows(g, pl);
#endif
    // Split rows in the lean value-mode-6 kernels (one-warp propagators with idle lanes: the 2-transmon gates run
    // 9 rows x 3 column groups on 27 lanes, and their one 5-entry row is shared with the 3 spare lanes): every lane
    // then executes rows of 1 + 2 entries instead of 1 + 4.  Opt-in (QSIM_SPLIT=1): parity-green and a third fewer
    // shared-memory wavefronts, but the six shuffles + selects per stage that bring the helper's partial sums to the
    // owner lengthen the dependent chain of the step (cfg2: 604 ms against 554 ms; profiles/r2_cfg2_experiments.txt).
    bool split6 = false;
    if (!split && !pl.rowsplit && !blocked && !pl.tiered && rpl == 1 && pl.n_batches == 1 && g.imag_only &&
        std::getenv("QSIM_SPLIT") && !std::getenv("QSIM_NO_LEAN")) {
        bool static_off = true;
        for (int j = 1; j < g.max_nnz && static_off; j++)
            for (int r = 0; r < n; r++) {
                const uint32_t sl = g.ell[(size_t)j * n + r] >> 16;
                if (sl != 0xFFFFu && (int)sl < g.n_dyn_slots) static_off = false;
            }
        if (static_off && build_split_rows(g, pl, pl.groups)) {
            if (find_variant(rpl, cpl, pl.split_nnz, 128, 6, 0, 1))
                split = split6 = true;
            else {
                pl.split_nv = pl.split_nnz = 0;
                pl.split_ell.clear();
                pl.split_info.clear();
            }
        }
    }
    const int eff_nnz = split ? pl.split_nnz : g.max_nnz;
    pl.rpl = rpl;
    if (pl.cf) {
        pl.team_warps = pl.cf->team_warps;
        pl.streamed = pl.cf->streamed;
    } else if (pl.cu) {
        pl.team_warps = (n + 3) / 4;
        pl.streamed = 0;   // all n <= 32 columns in one batch: the state never leaves registers
    } else if (pl.rowsplit) {
        pl.team_warps = (n + 95) / 96;
        pl.streamed = pl.n_batches > 1 ? 1 : 0;
    } else if (pl.n_batches <= 8 || (pl.n_batches == 9 && rpl == 1 && cpl == 3 && n > 24)) {
        pl.streamed = 0;
        pl.team_warps = pl.n_batches;
    } else {
        // streamed: eight warps share the batches (256 threads keep 255 registers per thread: no spills)
        pl.streamed = 1;
        pl.team_warps = 8;
    }
    // one-warp teams: 4 per CTA (128 threads; 2 CTAs = 8 warps per SM for the 255-register kernels, 3 CTAs = 12 warps
    // for the lean 168-register ones)
    pl.teams_per_cta = pl.team_warps == 1 ? 4 : 1;
    // (component forms: resident teams of 2-4 warps share the 256-thread CTA of the multi-warp kernels)
    if (pl.cf && !pl.streamed && pl.team_warps > 1) pl.teams_per_cta = 8 / pl.team_warps;
    if (const char *env = std::getenv("QSIM_TEAMS_PER_CTA")) {  // tuning knob (development)
        const int v = std::atoi(env);
        if (v >= 1 && v <= 15 && v * pl.team_warps * 32 <= 288) pl.teams_per_cta = v;
    }
    pl.block = pl.team_warps * 32 * pl.teams_per_cta;
    // value mode: 2 when the generator is purely imaginary and only diagonal entries depend on time
    int vm = g.imag_only ? 1 : 0;
    if (blocked) {
        vm = 3;
    } else if (pl.tiered) {
        vm = 5;
    } else if (split6) {
        vm = 6;
    } else if (pl.af) {
        vm = 7;
    } else if (vm == 1 && !pl.rowsplit) {
        const bool static_off = offdiag_static(gs);
        // (with four entries per lane the extra value registers spill in the 255-register kernels: mode 1 there; the
        // lean kernels of resident teams have room for them.  QSIM_LEAN_VM1=1: development A/B switch)
        const bool lean_ok = !pl.streamed && !split && !std::getenv("QSIM_NO_LEAN");
        if (static_off && rpl * cpl <= 3 && find_variant(rpl, cpl, eff_nnz, pl.block, 2, 0)) vm = 2;
        else if (static_off && lean_ok && !std::getenv("QSIM_LEAN_VM1") && find_variant(rpl, cpl, eff_nnz, pl.block, 2, 0, 1)) vm = 2;
    }
    if (vm == 0) {
        const bool mixed = generator_is_mixed(gs, pl.slot_imag);
        if (mixed && find_variant(rpl, cpl, eff_nnz, pl.block, 4, pl.rowsplit)) vm = 4;
    }
    pl.vm = vm;
    pl.kv = pl.tiered ? tiered_kv
            : pl.af   ? find_variant(rpl, cpl, pl.af->nnz, pl.block, 7, 0, 1)
            : split6  ? find_variant(rpl, cpl, pl.split_nnz, pl.block, 6, 0, 1)
                      : find_variant(rpl, cpl, blocked ? pl.blk_nnz : eff_nnz, pl.block, vm, pl.rowsplit);
    if (pl.tiered || split6 || pl.af) pl.lean = 1;
    if (!pl.kv && split) {   // (no variant with the shorter capacity and this value mode: fall back to whole rows)
        split = false;
        pl.split_nv = 0;
        pl.kv = find_variant(rpl, cpl, g.max_nnz, pl.block, vm, pl.rowsplit);
    }
    // Lean kernels (Kern<..., LEAN>: early-dying stage vectors, 168 registers, three warps per scheduler) for
    // resident teams when the shape has one; QSIM_NO_LEAN=1 keeps the 255-register kernels (development A/B switch).
    // (streamed teams stay on the 255-register kernels: the lean form of the d^2 = 81 kernel keeps its spills and
    // measures 3 % slower on cfg4, 776 ms against 751 ms for 1184 x 5 ns)
    if (!pl.streamed && !pl.rowsplit && !blocked && !split && !pl.tiered && !std::getenv("QSIM_NO_LEAN")) {
        if (const KernelVariant *lv = find_variant(rpl, cpl, eff_nnz, pl.block, vm, 0, 1)) {
            pl.kv = lv;
            pl.lean = 1;
        }
    }
    if (!pl.kv)
        return fail(QSIM_ERR_UNSUPPORTED, "generator rows with more than 16 entries have no kernel yet (row length " +
                                              std::to_string(g.max_nnz) + ")");
    // stage-buffer layout: smallest-cost (row stride, group stride) near the dense one
    if (pl.cu) {
        pl.y_row_stride = 32;   // a row of the stage buffer = 32 columns of 16 bytes: every gather reads one aligned 512-byte row
        pl.y_grp_stride = 1;
    } else if (pl.tiered) {
        pl.y_row_stride = pl.tf->S;     // chosen together with the row positions (build_tiered)
        pl.y_grp_stride = pl.tf->Cg;
    } else if (pl.cf) {
        pl.y_row_stride = pl.cf->stride;   // chosen together with the rows' positions (optimize_component_layout)
        pl.y_grp_stride = cpl;
        if (16 * n * pl.y_row_stride > 65535) return fail(QSIM_ERR_UNSUPPORTED, "state too large for the packed 16-bit row offsets");
        if (std::getenv("QSIM_DEBUG") && !pl.cf->rs)
            std::fprintf(stderr, "[qsim] component form: %d components, r%dc%d, %d warps x %d local rows, %d batches%s: stage stride %d, "
                                 "%ld wavefronts per stage\n", pl.cf->n_comp, rpl, cpl, pl.team_warps, n, pl.n_batches,
                         pl.streamed ? " (streamed)" : "", pl.y_row_stride, component_layout_cost(*pl.cf, pl.y_row_stride));
    } else {
        long best = -1;
        for (int Cg = cpl; Cg <= cpl + 16; Cg++)
            for (int S = (pl.groups - 1) * Cg + cpl; S <= (pl.groups - 1) * Cg + cpl + 16; S++) {
                if (pl.groups == 1 && Cg > cpl) break;  // no groups: only the row stride matters
                if (16 * n * S > 65535) continue;
                const long c = (blocked ? layout_cost(g, rpl, cpl, pl.groups, S, Cg, pl.blk_rows, pl.blk_ell.data(), pl.blk_nnz)
                                : split ? layout_cost(g, rpl, cpl, pl.groups, S, Cg, 0, pl.split_ell.data(), pl.split_nnz, pl.split_nv)
                                        : layout_cost(g, rpl, cpl, pl.groups, S, Cg)) * 64 + S;  // ties: smaller buffer
                if (best < 0 || c < best) {
                    best = c;
                    pl.y_row_stride = S;
                    pl.y_grp_stride = Cg;
                }
            }
        if (best < 0) return fail(QSIM_ERR_UNSUPPORTED, "state too large for the packed 16-bit row offsets");
        if (std::getenv("QSIM_DEBUG") && split)
            std::fprintf(stderr, "[qsim] split rows: %d helper lanes, rows of 1 + %d entries instead of 1 + %d\n", pl.split_nv - n,
                         pl.split_nnz - 1, g.max_nnz - 1);
        if (std::getenv("QSIM_DEBUG"))
            std::fprintf(stderr, "[qsim] n=%d groups=%d cpl=%d: stage layout S=%d Cg=%d wavefronts %ld (dense layout %ld)\n", n,
                         pl.groups, cpl, pl.y_row_stride, pl.y_grp_stride, best / 64,
                         blocked ? layout_cost(g, rpl, cpl, pl.groups, cpb, cpl, pl.blk_rows, pl.blk_ell.data(), pl.blk_nnz)
                                 : layout_cost(g, rpl, cpl, pl.groups, cpb, cpl));
    }
    if (g.n_slots > 4094)
        return fail(QSIM_ERR_UNSUPPORTED, "generator too large for the packed 16-bit slot offsets");
    pl.nvp = (g.n_slots + 1 + 1) & ~1;  // + the zero slot, even
    pl.ncp = (g.n_chan + 1) & ~1;
    pl.n_dyn_pairs = g.slot_ptr[g.n_dyn_slots];
    const int n_ev = std::max<int>((int)g.evals.size(), 1);
    // common shared region: series tables | Chebyshev block | controller power tables | dynamic pair lists | records
    pl.off_dptr = 512 + 1280 + 8 * (QSIM_POW_K * 32 + 2 * QSIM_POW_NE);
    pl.off_dw = pl.off_dptr + ((4 * (g.n_dyn_slots + 1 + pl.n_dyn_pairs) + 15) & ~15);
    pl.off_drec = pl.off_dw + 16 * std::max(pl.n_dyn_pairs, 1);
    pl.dyn_fast = 1;
    for (int sl = 0; sl < g.n_dyn_slots; sl++)
        if (g.slot_ptr[sl + 1] - g.slot_ptr[sl] > 3) pl.dyn_fast = 0;
    pl.common_bytes = pl.off_drec + (pl.dyn_fast ? 64 * std::max(g.n_dyn_slots, 1) : 16);
    pl.team_bytes = (((((pl.vm >= 1 && pl.vm <= 3) || pl.vm >= 5) ? 8 : 16) * 5 * pl.nvp + 15) & ~15) + 8 * 5 * pl.ncp + (int)sizeof(ResolvedEval) * n_ev + 256 * pl.ncp + 8 * (pl.team_warps * 4 + 8) +
                    (pl.rowsplit ? 1 : pl.team_warps) * 2 * 16 * n * pl.y_row_stride + 128 * pl.team_warps;
    if (pl.lean) {   // park area: k2..k6 of a step with save points, 512 bytes per entry and warp
        pl.off_park = (pl.team_bytes + 15) & ~15;
        pl.team_bytes = pl.off_park + pl.team_warps * 5 * rpl * cpl * 512;
    }
    const size_t smem_cap = 227 * 1024;   // opt-in dynamic shared memory per CTA on sm_100a
    if (pl.rowsplit) {
        if (pl.cf)
            build_rowsplit_component_tables(*pl.cf, pl, std::getenv("QSIM_DEBUG_LAYOUT") != nullptr);
        else
        build_rowsplit_table(g, pl, std::getenv("QSIM_DEBUG_LAYOUT") != nullptr);   // (debug: run the annealer without a GPU)
        // team region: + three mbarriers, + the two prefetch buffers (u, k1 of a column each) for streamed teams;
        // common region: + the row table.  Both are dropped (global-memory table, direct scratch loads) when the
        // CTA would not fit.
        pl.off_pf = (pl.team_bytes + 15) & ~15;
        pl.team_bytes = pl.off_pf + 32;
        pl.off_ell = (pl.common_bytes + 15) & ~15;
        const size_t base = (size_t)pl.off_ell + (size_t)pl.team_bytes * pl.teams_per_cta;
        const size_t pf = pl.streamed ? (size_t)4 * 16 * n : 0;
        const bool no_smem_table = std::getenv("QSIM_RS_GLOBAL_TABLE") != nullptr;   // development switches
        const bool no_prefetch = std::getenv("QSIM_RS_NO_PREFETCH") != nullptr;
        if (!no_smem_table && base + pl.ell_bytes <= smem_cap) {
            pl.ell_in_smem = 1;
            pl.common_bytes = pl.off_ell + pl.ell_bytes;
        }
        if (pf > 0 && !no_prefetch && (size_t)pl.common_bytes + (size_t)(pl.team_bytes + pf) * pl.teams_per_cta <= smem_cap) {
            pl.rs_prefetch = 1;
            pl.team_bytes += (int)pf;
        }
    }
    pl.smem = (size_t)pl.common_bytes + (size_t)pl.team_bytes * pl.teams_per_cta;
    // large value tables (dense Hamiltonians): fewer teams per CTA until the CTA fits
    while (pl.smem > smem_cap && pl.teams_per_cta > 1) {
        pl.teams_per_cta /= 2;
        pl.block = pl.team_warps * 32 * pl.teams_per_cta;
        pl.smem = (size_t)pl.common_bytes + (size_t)pl.team_bytes * pl.teams_per_cta;
    }
    if (std::getenv("QSIM_DEBUG"))
        std::fprintf(stderr, "[qsim] n=%d slots=%d (dyn %d) chan=%d: common %d B, team %d B x %d, smem %zu B%s%s\n", n, g.n_slots,
                     g.n_dyn_slots, g.n_chan, pl.common_bytes, pl.team_bytes, pl.teams_per_cta, pl.smem,
                     pl.ell_in_smem ? ", row table in smem" : "", pl.rs_prefetch ? ", bulk prefetch" : "");
    pl.scratch_per_team = pl.streamed ? (size_t)4 * pl.n_batches * cpb * n * sizeof(double2) : 0;
    if (pl.smem > smem_cap) return fail(QSIM_ERR_UNSUPPORTED, "problem needs more shared memory than one SM has");
    return QSIM_OK;
}

extern "C" int qsim_problem_kernel_name(qsim_problem *p, int which, char *buf, int buflen) {
    if (!p || !buf || buflen < 1 || which < 0 || which > 3) return fail(QSIM_ERR_ARG, "bad argument");
    Plan pl;
    if (int rc = make_plan(p, which, pl)) return rc;
    std::snprintf(buf, (size_t)buflen, "%s%s_tw%d", pl.kv->name, pl.streamed ? "_streamed" : "_resident", pl.team_warps);
    return QSIM_OK;
}

extern "C" int qsim_problem_plan_info(qsim_problem *p, int which, int64_t info[16]) {
    if (!p || !info || which < 0 || which > 3) return fail(QSIM_ERR_ARG, "bad argument");
    Plan pl;
    if (int rc = make_plan(p, which, pl)) return rc;
    const Generator &g = p->gen[pl.w];
    const int64_t v[16] = {g.n, pl.ncols, g.nnz, g.max_nnz, g.n_slots, g.n_dyn_slots, g.n_chan, pl.team_warps,
                           pl.teams_per_cta, pl.block, (int64_t)pl.smem, pl.streamed, (int64_t)pl.scratch_per_team, pl.vm,
                           pl.rowsplit, pl.cf ? pl.cf->n_comp : 0};
    std::memcpy(info, v, sizeof(v));
    return QSIM_OK;
}

extern "C" int qsim_problem_flops_per_rhs(qsim_problem *p, int which, double *flops) {
    if (!p || !flops || which < 0 || which > 3) return fail(QSIM_ERR_ARG, "bad argument");
    Plan pl;
    if (int rc = make_plan(p, which, pl)) return rc;
    const Generator &g = p->gen[pl.w];
    const double ncols = (which == 0 || which == 2) ? (double)g.n : 1.0;
    // a complex multiply-add is 4 DFMA (8 flop); value modes 1-4 (entries purely imaginary, or each purely
    // real or purely imaginary) issue 2 (4 flop)
    *flops = (pl.vm >= 1 ? 4.0 : 8.0) * (double)g.nnz * ncols;
    // component pruning: only the entries inside the components' blocks are integrated
    if (pl.cf) *flops = (pl.vm >= 1 ? 4.0 : 8.0) * (double)pl.cf->nnz_cols;
    return QSIM_OK;
}

template <typename T>
static cudaError_t upload(T **dst, const std::vector<T> &src) {
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    cudaError_t e = cudaMalloc((void **)dst, bytes);
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
    return e;
}

// The generator's device-resident tables on the ctx's device (uploaded on first use, then shared by every
// ctx on that device).  Called with the CUDA device already set.
static int ensure_device_copy(qsim_ctx *c, qsim_problem *p, int w, qsim_problem::DeviceCopy **out) {
    std::lock_guard<std::mutex> lock(p->dev_mutex);
    for (auto *have : p->dev[w])
        if (have->device == c->device) {
            *out = have;
            return QSIM_OK;
        }
    const Generator &g = p->gen[w];
    std::unique_ptr<qsim_problem::DeviceCopy> holder(new qsim_problem::DeviceCopy());
    qsim_problem::DeviceCopy *dc = holder.get();
    dc->device = c->device;
    std::vector<double2> w2(g.pair_w.size());
    for (size_t i = 0; i < w2.size(); i++) w2[i] = make_double2(g.pair_w[i].real(), g.pair_w[i].imag());
    CU(upload(&dc->ell, g.ell));
    CU(upload(&dc->row_len, g.row_len));
    CU(upload(&dc->slot_ptr, g.slot_ptr));
    CU(upload(&dc->pair_chan, g.pair_chan));
    CU(upload(&dc->pair_w, w2));
    CU(upload(&dc->evals, g.evals));
    if (!g.tables.empty()) CU(upload(&dc->tables, g.tables));
    {
        // {w0, w1, w2 (re, im), ch0, ch1, ch2, pad}: unused pairs carry weight 0 on the constant channel.
        // When the time-dependent slots involve only the constant channel and at most two time-dependent
        // channels A < B, the records are written in (constant, A, B) order so the kernels need no indices.
        std::vector<int> dyn_ch;
        for (const auto &e : g.evals)
            if (e.dynamic)
                for (int q = 0; q < eval_outputs(e.kind); q++) dyn_ch.push_back(e.out_chan + q);
        bool fast = !dyn_ch.empty() && dyn_ch.size() <= 2 && g.n_dyn_slots > 0;
        const int chA = dyn_ch.empty() ? 0 : dyn_ch[0], chB = dyn_ch.size() > 1 ? dyn_ch[1] : chA;
        for (int sl = 0; sl < g.n_dyn_slots && fast; sl++)
            for (int q = g.slot_ptr[sl]; q < g.slot_ptr[sl + 1]; q++)
                if (g.pair_chan[q] != 0 && g.pair_chan[q] != chA && g.pair_chan[q] != chB) fast = false;
        std::vector<double> rec((size_t)8 * std::max(g.n_dyn_slots, 1), 0.0);
        for (int sl = 0; sl < g.n_dyn_slots; sl++) {
            int32_t ch[4] = {0, 0, 0, 0};
            if (fast) {
                ch[1] = chA;
                ch[2] = chB;
                for (int q = g.slot_ptr[sl]; q < g.slot_ptr[sl + 1]; q++) {
                    const int pos = g.pair_chan[q] == 0 ? 0 : (g.pair_chan[q] == chA ? 1 : 2);
                    rec[(size_t)8 * sl + 2 * pos] += g.pair_w[q].real();
                    rec[(size_t)8 * sl + 2 * pos + 1] += g.pair_w[q].imag();
                }
            } else {
                const int np = std::min(3, g.slot_ptr[sl + 1] - g.slot_ptr[sl]);
                for (int q = 0; q < np; q++) {
                    rec[(size_t)8 * sl + 2 * q] = g.pair_w[g.slot_ptr[sl] + q].real();
                    rec[(size_t)8 * sl + 2 * q + 1] = g.pair_w[g.slot_ptr[sl] + q].imag();
                    ch[q] = g.pair_chan[g.slot_ptr[sl] + q];
                }
            }
            std::memcpy(&rec[(size_t)8 * sl + 6], ch, sizeof(ch));
        }
        CU(upload(&dc->dyn_rec, rec));
        dc->fast_ok = fast ? 1 : 0;
        dc->fast_chA = chA;
        dc->fast_chB = chB;
    }
    p->dev[w].push_back(holder.release());
    *out = dc;
    return QSIM_OK;
}

#define QSIM_FLAG_INTERNAL_DEVICE_OUT (1u << 30)   // library-internal: `out` is device memory, everything else host
struct Selection {
    int n_sel = 0, abs2 = 0;
    const int32_t *rows = nullptr, *cols = nullptr;
};
// A block-cyclic shard of a sweep: the call integrates points first, first + step, ... (n_traj of them) of arrays
// that hold `total` points; results go to the same positions of the caller's arrays.
struct Shard {
    int64_t first = 0, step = 1, total = 0;
};
static bool host_pointer_is_pinned(const void *ptr, void **dev_alias) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    *dev_alias = at.devicePointer;
    return true;
}

// State evolution (unitary_state / me_state) with host-side initial states: the components of the generator graph that
// the initial states touch (any of them, when every trajectory brings its own).  Entries of the other components stay
// exactly zero, in the reference too, so only the touched ones are integrated (csrc/qsim_components.cpp, state forms);
// forms are cached in the problem by the set of touched components.  *out stays nullptr when nothing can be pruned.
static int state_form_for(qsim_problem *p, int which, const double *init, int64_t init_stride, int64_t n_points,
                          const ComponentForm **out) {
    *out = nullptr;
    if (std::getenv("QSIM_NO_PRUNE")) return QSIM_OK;
    const int w = which >= 2 ? 1 : 0;
    if (int rc = ensure_generator(p, w)) return rc;
    const Generator &g = p->gen[w];
    if (g.n > 9 * 96) return QSIM_OK;
    std::lock_guard<std::mutex> lk(p->plan_mutex);
    if (p->n_comp[w] < 0) {
        std::vector<std::vector<int>> comps;
        generator_components(g, p->comp_of[w], comps);
        p->n_comp[w] = (int)comps.size();
    }
    if (p->n_comp[w] < 2) return QSIM_OK;
    std::vector<char> touched((size_t)p->n_comp[w], 0);
    int n_touched = 0;
    const int64_t n_init = init_stride ? n_points : 1;
    for (int64_t t = 0; t < n_init && n_touched < p->n_comp[w]; t++) {
        const double *v = init + 2 * (size_t)t * (size_t)init_stride;
        for (int r = 0; r < g.n; r++)
            if ((v[2 * r] != 0.0 || v[2 * r + 1] != 0.0) && !touched[(size_t)p->comp_of[w][(size_t)r]]) {
                touched[(size_t)p->comp_of[w][(size_t)r]] = 1;
                n_touched++;
            }
    }
    if (n_touched == 0 || n_touched == p->n_comp[w]) return QSIM_OK;
    for (const auto &q : p->sforms[w])
        if (q.first == touched) {
            if (q.second && q.second->ok) *out = q.second.get();
            return QSIM_OK;
        }
    auto cf = std::make_shared<ComponentForm>();
    const bool ok = build_state_form(g, touched, *cf);
    p->sforms[w].emplace_back(touched, ok ? cf : nullptr);
    if (ok) *out = cf.get();
    return QSIM_OK;
}

static int solve_common(qsim_ctx *c, qsim_problem *p, int which, int64_t n_traj, const double *params,
                        const double *ts, int n_ts, int64_t ts_stride, const double *init, int64_t init_stride,
                        const qsim_solver_opts *opts_in, double *out, qsim_traj_stats *stats,
                        const Selection *sel = nullptr, const Shard *shard = nullptr) {
    if (!c) return fail(QSIM_ERR_NO_DEVICE, "ctx is NULL: create one with qsim_create (needs a CUDA device; no CPU fallback)");
    if (!p) return fail(QSIM_ERR_ARG, "prob is NULL");
    if (n_traj < 0) return fail(QSIM_ERR_ARG, "n_traj < 0");
    if (n_ts < 1 || !ts) return fail(QSIM_ERR_ARG, "need at least one save time");
    if (!out) return fail(QSIM_ERR_ARG, "out is NULL");
    if (p->n_params > 0 && !params) return fail(QSIM_ERR_ARG, "params is NULL but the problem has sweep parameters");
    if ((which == 1 || which == 3) && !init) return fail(QSIM_ERR_ARG, "initial state is NULL");
    if (ts_stride != 0 && ts_stride < n_ts) return fail(QSIM_ERR_ARG, "ts_stride must be 0 or >= n_ts");
    if (which == 1 || which == 3) {
        // a per-trajectory initial state must hold the whole state vector (d or d*d complex numbers)
        const int64_t n_init = which == 1 ? p->d : (int64_t)p->d * p->d;
        if (init_stride != 0 && init_stride < n_init)
            return fail(QSIM_ERR_ARG, "init_stride must be 0 (shared initial state) or >= the state length");
    }
    qsim_solver_opts o;
    if (opts_in)
        o = *opts_in;
    else
        qsim_default_opts(&o);
    if (!(o.reltol > 0) || !(o.abstol >= 0) || !(o.gamma > 0) || !(o.qmin > 0) || !(o.qmax > 0))
        return fail(QSIM_ERR_ARG, "solver options out of range");
    const bool dev_ptrs = (o.flags & QSIM_FLAG_DEVICE_PTRS) != 0;
    Shard sh;
    if (shard) sh = *shard;
    else sh.total = n_traj;
    if (!dev_ptrs && !shard) {   // (a sharded call's arrays were checked once by its caller)
        // @assert issorted(ts)   (time_evolution.jl:30,64,97,141)
        const int64_t ngrids = ts_stride ? n_traj : 1;
        for (int64_t i = 0; i < ngrids; i++)
            for (int k = 1; k < n_ts; k++)
                if (!(ts[i * ts_stride + k - 1] <= ts[i * ts_stride + k]))
                    return fail(QSIM_ERR_ARG, "ts must be sorted");
    }
    if (n_traj == 0) return QSIM_OK;
    // (abstol = 0: the reference's error norm divides 0 by 0 on entries that are exactly zero; the full state is
    // integrated then, so that the library fails the same way instead of succeeding on the component blocks)
    const bool allow_prune = o.abstol > 0.0;
    const ComponentForm *state_cf = nullptr;
    if ((which == 1 || which == 3) && !dev_ptrs && init && allow_prune)
        if (int rc = state_form_for(p, which, init, init_stride, sh.total, &state_cf)) return rc;
    Plan pl;
    if (int rc = make_plan(p, which, pl, state_cf, allow_prune)) return rc;
    CU(cudaSetDevice(c->device));
    qsim_problem::DeviceCopy *dc = nullptr;
    if (int rc = ensure_device_copy(c, p, pl.w, &dc)) return rc;
    const Generator &g = p->gen[pl.w];
    // device tables that depend on the plan: looked up / uploaded under the problem's lock, and the pointers this
    // launch uses are copied out while it is held (another ctx may add sets concurrently; sets are never replaced)
    qsim_problem::DeviceCopy::Packed pk;
    const uint8_t *d_slot_imag = nullptr;
    const uint32_t *d_blk_ell = nullptr;
    const uint16_t *d_blk_slot = nullptr;
    const int *d_tr_perm = nullptr, *d_tr_pos = nullptr;
    qsim_problem::DeviceCopy::CfDev cfd;
    {
        std::lock_guard<std::mutex> packed_lock(p->dev_mutex);
        if (pl.rowsplit) {
            bool have = false;
            for (const auto &q : dc->packed)
                if (q.stride == pl.y_row_stride && q.warps == pl.team_warps && q.vm == pl.vm && q.cu == pl.cu && q.cf == pl.cf) {
                    pk = q;
                    have = true;
                }
            if (!have) {
                pk.stride = pl.y_row_stride;
                pk.warps = pl.team_warps;
                pk.vm = pl.vm;
                pk.cu = pl.cu;
                pk.cf = pl.cf;
                if (pl.cf) {
                    build_rowsplit_component_tables(*pl.cf, pl, true);
                    CU(upload(&pk.pass, pl.rs_lane));
                } else
                build_rowsplit_table(g, pl, true);
                CU(upload(&pk.ell, pl.rs_table));
                CU(upload(&pk.pos, pl.rs_pos));
                CU(upload(&pk.perm, pl.rs_perm));
                CU(upload(&pk.off, pl.rs_off));
                CU(upload(&pk.len, pl.rs_len));
                dc->packed.push_back(pk);
            }
        }
        if (pl.vm == 4 && !dc->slot_imag) CU(upload(&dc->slot_imag, pl.slot_imag));
        if (pl.blk_rows > 0 && !pl.tiered && !dc->blk_ell) {
            CU(upload(&dc->blk_ell, pl.blk_ell));
            CU(upload(&dc->blk_slot, pl.blk_slot));
        }
        if (pl.af && !dc->as_ell) {
            CU(upload(&dc->as_ell, pl.af->gv.ell));
            CU(upload(&dc->as_phys, pl.af->phys));
            CU(upload(&dc->as_info, pl.af->lane_info));
        }
        if (pl.cf) {
            for (const auto &q : dc->cf_sets)
                if (q.form == pl.cf) cfd = q;
            if (!cfd.form) {
                cfd.form = pl.cf;
                if (!pl.cf->rs) {
                    CU(upload(&cfd.ell, pl.cf->ell));
                    CU(upload(&cfd.vrow, pl.cf->vrow));
                    CU(upload(&cfd.vcol, pl.cf->vcol));
                }
                CU(upload(&cfd.zeros, pl.cf->zeros));
                dc->cf_sets.push_back(cfd);
            }
        }
        if (pl.tiered && !dc->tr_ell) {
            CU(upload(&dc->tr_ell, pl.tf->ell));
            CU(upload(&dc->tr_perm, pl.tf->perm));
            CU(upload(&dc->tr_pos, pl.tf->pos));
        }
        if (pl.split_nv > 0 && !dc->split_ell) {
            CU(upload(&dc->split_ell, pl.split_ell));
            CU(upload(&dc->split_info, pl.split_info));
        }
        d_slot_imag = dc->slot_imag;
        d_blk_ell = pl.tiered ? dc->tr_ell : dc->blk_ell;
        d_blk_slot = dc->blk_slot;
        d_tr_perm = dc->tr_perm;
        d_tr_pos = dc->tr_pos;
    }
    cudaStream_t stream = o.stream ? (cudaStream_t)o.stream : c->stream;

    int ctas_per_sm = 0;
    CU(pl.kv->occupancy(pl.block, pl.smem, &ctas_per_sm));
    if (ctas_per_sm < 1) return fail(QSIM_ERR_UNSUPPORTED, "kernel does not fit on an SM with this configuration");
    const int64_t need = (n_traj + pl.teams_per_cta - 1) / pl.teams_per_cta;
    pl.grid = (int)std::min<int64_t>((int64_t)ctas_per_sm * c->prop.multiProcessorCount, need);

    const size_t n_state = (size_t)g.n * pl.ncols;
    const size_t b_params = sizeof(double) * (size_t)std::max(p->n_params, 1) * sh.total;
    const size_t b_ts = sizeof(double) * (size_t)(ts_stride ? ts_stride * sh.total : n_ts);
    const size_t b_init = init ? sizeof(double2) * (size_t)(init_stride ? init_stride * sh.total : g.n) : 0;
    size_t traj_bytes = sizeof(double2) * n_state * n_ts;     // result bytes per trajectory

    LaunchArgs a;
    std::memset(&a, 0, sizeof(a));
    if (sel) {
        // position table of the selected entries (column-major state index -> position, -1 elsewhere)
        std::vector<int> index(n_state, -1);
        for (int i = 0; i < sel->n_sel; i++) {
            const int r = sel->rows[i], cc = sel->cols ? sel->cols[i] : 0;
            if (r < 0 || r >= g.n || cc < 0 || cc >= pl.ncols) return fail(QSIM_ERR_ARG, "selected entry out of range");
            if (index[(size_t)cc * g.n + r] >= 0) return fail(QSIM_ERR_ARG, "selected entries must be distinct");
            index[(size_t)cc * g.n + r] = i;
        }
        // component pruning: the selected entries that are structurally zero (positions inside the selection)
        size_t n_sel_zero = 0;
        if (pl.cf)
            for (int i = 0; i < sel->n_sel; i++)
                if (pl.cf->state ? !pl.cf->live[(size_t)sel->rows[i]]
                                 : pl.cf->comp_of[(size_t)sel->rows[i]] != pl.cf->comp_of[(size_t)(sel->cols ? sel->cols[i] : 0)]) {
                    index.push_back(i);
                    n_sel_zero++;
                }
        CU(c->sel.reserve(sizeof(int) * index.size()));
        CU(cudaMemcpyAsync(c->sel.p, index.data(), sizeof(int) * index.size(), cudaMemcpyHostToDevice, stream));
        CU(cudaStreamSynchronize(stream));  // `index` is pageable and goes out of scope
        a.sel_index = (const int *)c->sel.p;
        a.zero_pos = (const int *)c->sel.p + n_state;
        a.n_zero = (int)n_sel_zero;
        a.n_sel = sel->n_sel;
        a.sel_abs2 = sel->abs2;
        traj_bytes = (sel->abs2 ? sizeof(double) : sizeof(double2)) * (size_t)sel->n_sel * n_ts;
    }
    // How results leave the device (host buffers):
    //   zero-copy  the caller's `out` is page-locked (qsim_alloc_pinned / cudaHostRegister): the kernel writes it in
    //              place over PCIe while it integrates -- no device result buffer, no copy after the kernel;
    //   pipelined  pageable `out`, large results: the trajectories are cut into chunks launched alternately on two
    //              streams into two device buffers; chunk k's results are copied out while chunk k + 1 integrates;
    //   plain      small results: one launch, one copy.
    void *out_alias = nullptr, *stats_alias = nullptr;
    const bool device_out = (o.flags & QSIM_FLAG_INTERNAL_DEVICE_OUT) != 0;   // host inputs, results stay on the device
    if (device_out) out_alias = out;
    const bool zero_copy = device_out || (!dev_ptrs && !std::getenv("QSIM_NO_ZEROCOPY") && host_pointer_is_pinned(out, &out_alias));
    const bool stats_direct = zero_copy && stats && host_pointer_is_pinned(stats, &stats_alias);
    const int64_t capacity = (int64_t)ctas_per_sm * c->prop.multiProcessorCount * pl.teams_per_cta;  // teams in flight
    int n_chunks = 1;
    int64_t per_chunk = n_traj;
    if (!dev_ptrs && !zero_copy && !o.stream && (double)traj_bytes * (double)n_traj > 64e6 && !std::getenv("QSIM_NO_PIPELINE") &&
        n_traj > capacity) {
        // whole waves per chunk (a multiple of the teams the device holds), at most eight chunks; the last chunk is the
        // remainder, so the copy that nothing overlaps is as small as the sweep allows
        const int64_t waves = (n_traj + capacity - 1) / capacity;
        per_chunk = capacity * ((waves + 7) / 8);
        n_chunks = (int)((n_traj + per_chunk - 1) / per_chunk);
    }
    if (dev_ptrs) {
        a.params = params;
        a.ts = ts;
        a.init = reinterpret_cast<const double2 *>(init);
        a.out = reinterpret_cast<double2 *>(out);
        a.stats = stats;
    } else {
        CU(c->params.reserve(b_params));
        CU(c->ts.reserve(b_ts));
        if (!zero_copy) {
            CU(c->out.reserve(traj_bytes * (size_t)per_chunk));
            if (n_chunks > 1) CU(c->out2.reserve(traj_bytes * (size_t)per_chunk));
        }
        // the kernel writes a trajectory's stats where it writes its results: at the sweep point's own position when the
        // results go out in place (zero-copy), so a shard's device stats buffer spans the shard's positions
        if (!stats_direct)
            CU(c->stats.reserve(sizeof(qsim_traj_stats) * (size_t)(zero_copy ? sh.first + (n_traj - 1) * sh.step + 1 : n_traj)));
        if (params && p->n_params > 0) CU(cudaMemcpyAsync(c->params.p, params, b_params, cudaMemcpyHostToDevice, stream));
        CU(cudaMemcpyAsync(c->ts.p, ts, b_ts, cudaMemcpyHostToDevice, stream));
        if (init) {
            CU(c->init.reserve(b_init));
            CU(cudaMemcpyAsync(c->init.p, init, b_init, cudaMemcpyHostToDevice, stream));
        }
        a.params = (const double *)c->params.p;
        a.ts = (const double *)c->ts.p;
        a.init = (const double2 *)c->init.p;
        a.out = zero_copy ? (double2 *)out_alias : (double2 *)c->out.p;
        a.stats = stats_direct ? (qsim_traj_stats *)stats_alias : (qsim_traj_stats *)c->stats.p;
    }
    a.in_first = sh.first;
    a.in_step = sh.step;
    a.out_first = (dev_ptrs || zero_copy) ? sh.first : 0;
    a.out_step = (dev_ptrs || zero_copy) ? sh.step : 1;
    if (pl.streamed) {
        const size_t scratch_bytes = pl.scratch_per_team * (size_t)pl.grid * pl.teams_per_cta;
        CU(c->scratch.reserve(scratch_bytes));
        a.scratch = (double2 *)c->scratch.p;
        if (n_chunks > 1) CU(c->scratch2.reserve(pl.scratch_per_team * (size_t)pl.grid * pl.teams_per_cta));
    }

    a.g.n = g.n;
    a.n_out = g.n;
    a.g.max_nnz = g.max_nnz;
    a.g.n_slots = g.n_slots;
    a.g.n_dyn_slots = g.n_dyn_slots;
    a.g.n_chan = g.n_chan;
    a.g.n_evals = (int)g.evals.size();
    a.g.n_dyn_pairs = pl.n_dyn_pairs;
    a.g.ell = dc->ell;
    a.g.n_vrows = g.n;
    a.g.row_len = dc->row_len;
    a.g.slot_ptr = dc->slot_ptr;
    a.g.pair_chan = dc->pair_chan;
    a.g.pair_w = dc->pair_w;
    a.g.evals = dc->evals;
    a.g.tab = c->d_tab;
    a.g.tables = dc->tables;
    a.cheb = c->d_cheb;
    a.dyn_chan_mask = 0;
    bool any_dyn = false;
    for (const auto &e : g.evals)
        if (e.dynamic) {
            any_dyn = true;
            const int n_out = eval_outputs(e.kind);
            for (int q = 0; q < n_out; q++)
                if (e.out_chan + q < 32) a.dyn_chan_mask |= 1u << (e.out_chan + q);
        }
    a.use_windows = (any_dyn && g.n_chan <= 32 && !(o.flags & QSIM_FLAG_DIRECT_DRIVES)) ? 1 : 0;
    a.fast_tables = (a.use_windows && dc->fast_ok) ? 1 : 0;
    a.fast_chA = dc->fast_chA;
    a.fast_chB = dc->fast_chB;
    if (pl.blk_rows > 0) {
        a.g.ell = d_blk_ell;
        a.g.max_nnz = pl.blk_nnz;
        a.blk_slot = d_blk_slot;
        a.n_blk_rows = pl.blk_rows;
        a.row_perm = d_tr_perm;
        a.row_pos = d_tr_pos;
    }
    if (pl.af) {   // the kernel runs over the table rows; results, norms and the initial state refer to the state rows
        a.g.n = pl.af->n_v;
        a.g.n_vrows = pl.af->n_v;
        a.g.ell = dc->as_ell;
        a.g.max_nnz = pl.af->nnz;
        a.row_perm = dc->as_phys;
        a.split_info = dc->as_info;
    }
    if (pl.cf) {   // the kernel runs over each warp's local rows; results, norms and the initial state refer to the state
        a.g.n = pl.cf->n_loc;
        a.g.n_vrows = pl.cf->team_warps * pl.cf->n_loc;
        if (pl.cf->rs) {
            a.rs_pass = reinterpret_cast<const int4 *>(pk.pass);
        } else {
            a.g.ell = cfd.ell;
            a.cf_vrow = cfd.vrow;
            a.cf_vcol = cfd.vcol;
        }
        if (!sel) {
            a.zero_pos = cfd.zeros;
            a.n_zero = (int)pl.cf->zeros.size();
        }
    }
    if (pl.split_nv > 0) {
        a.g.ell = dc->split_ell;
        a.g.max_nnz = pl.split_nnz;
        a.g.n_vrows = pl.split_nv;
        a.split_info = dc->split_info;
    }
    a.ell_packed = pk.ell;
    a.ell_rows = g.n + 8;
    a.rs_perm = pk.perm;
    a.rs_off = pk.off;
    a.rs_pos = pk.pos;
    a.rs_len = pk.len;
    a.ell_in_smem = pl.ell_in_smem;
    a.off_ell = pl.off_ell;
    a.ell_bytes = pl.ell_bytes;
    a.rs_prefetch = pl.rs_prefetch;
    a.cu_map = pl.cu;
    a.off_pf = pl.off_pf;
    a.slot_imag = d_slot_imag;
    a.ncols = pl.cf ? pl.n_batches * pl.cf->cpl : pl.ncols;   // (component form: virtual columns, batch b holds [b * cpl, b * cpl + cpl))
    a.out_cols = pl.ncols;
    a.identity_init = pl.identity;
    a.n_params = p->n_params;
    a.n_ts = n_ts;
    a.ts_stride = ts_stride;
    a.init_stride = init_stride;
    a.n_traj = n_traj;
    a.work_counter = c->d_counter;
    a.team_warps = pl.team_warps;
    a.teams_per_cta = pl.teams_per_cta;
    a.n_batches = pl.n_batches;
    a.groups = pl.groups;
    a.y_row_stride = pl.y_row_stride;
    a.y_grp_stride = pl.y_grp_stride;
    a.streamed = pl.streamed;
    a.nvp = pl.nvp;
    a.ncp = pl.ncp;
    a.common_bytes = pl.common_bytes;
    a.off_dptr = pl.off_dptr;
    a.off_dw = pl.off_dw;
    a.off_drec = pl.off_drec;
    a.dyn_fast = pl.dyn_fast;
    a.dyn_rec = dc->dyn_rec;
    a.team_bytes = pl.team_bytes;
    a.off_park = pl.off_park;
    a.reltol = o.reltol; a.abstol = o.abstol; a.gamma = o.gamma; a.qmin = o.qmin; a.qmax = o.qmax;
    a.beta1 = o.beta1; a.beta2 = o.beta2; a.qoldinit = o.qoldinit; a.dtmax = o.dtmax;
    a.maxiters = o.maxiters > 0 ? o.maxiters : 1000000;
    for (int w = 0; w < 2; w++) {
        const double pw = w == 0 ? -0.5 * o.beta1 : 0.5 * o.beta2;
        for (int idx = 0; idx < 16; idx++) {
            const double cc = 1.03125 + 0.0625 * idx;
            double coef = std::pow(cc, pw);
            for (int k = 0; k < QSIM_POW_K; k++) {
                a.pow_coef[w][idx][k] = coef;
                coef = coef * (pw - k) / ((k + 1) * cc);  // binom(p, k+1) c^(p-k-1)
            }
        }
        for (int e = 0; e < QSIM_POW_NE; e++) a.pow_exp2[w][e] = std::exp2(pw * (double)(e + QSIM_POW_EMIN));
    }

    if (n_chunks == 1) {
        CU(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), stream));
        CU(cudaEventRecord(c->ev0, stream));
        CU(pl.kv->launch(a, pl.grid, pl.block, pl.smem, stream));
        CU(cudaEventRecord(c->ev1, stream));
        c->timed = true;
        c->launches++;
        if (!dev_ptrs) {
            // results: in place already (zero-copy), or one strided copy into the caller's positions
            if (!zero_copy)
                CU(cudaMemcpy2DAsync((char *)out + traj_bytes * (size_t)sh.first, traj_bytes * (size_t)sh.step, c->out.p, traj_bytes,
                                     traj_bytes, (size_t)n_traj, cudaMemcpyDeviceToHost, stream));
            if (stats && !stats_direct) {
                if (zero_copy)   // device stats at the points' own positions (see above)
                    CU(cudaMemcpy2DAsync(stats + sh.first, sizeof(qsim_traj_stats) * (size_t)sh.step,
                                         (const qsim_traj_stats *)c->stats.p + sh.first, sizeof(qsim_traj_stats) * (size_t)sh.step,
                                         sizeof(qsim_traj_stats), (size_t)n_traj, cudaMemcpyDeviceToHost, stream));
                else
                    CU(cudaMemcpy2DAsync(stats + sh.first, sizeof(qsim_traj_stats) * (size_t)sh.step, c->stats.p, sizeof(qsim_traj_stats),
                                         sizeof(qsim_traj_stats), (size_t)n_traj, cudaMemcpyDeviceToHost, stream));
            }
            CU(cudaStreamSynchronize(stream));
        }
        return QSIM_OK;
    }
    // pipelined: chunk k runs on stream k % 2 into device result buffer k % 2 (its own scratch and work counter).
    // While chunk k integrates, this thread drains chunk k - 1: pieces of at most QSIM_STAGE_BYTES are copied into two
    // small page-locked staging buffers (DMA at full rate) and moved from there to the caller's pageable, possibly
    // strided positions, double-buffered.  Inputs were queued on `stream`: make the second stream wait for them.
    const size_t stage_bytes = std::max<size_t>(traj_bytes, (size_t)32 << 20);
    for (int i = 0; i < 2; i++) {
        if (c->stage_cap[i] < stage_bytes) {
            if (c->stage[i]) CU(cudaFreeHost(c->stage[i]));
            c->stage[i] = nullptr;
            c->stage_cap[i] = 0;
            CU(cudaHostAlloc(&c->stage[i], stage_bytes, cudaHostAllocDefault));
            c->stage_cap[i] = stage_bytes;
        }
    }
    CU(cudaEventRecord(c->ev1, stream));
    CU(cudaStreamWaitEvent(c->stream2, c->ev1, 0));
    cudaStream_t streams[2] = {stream, c->stream2};
    const int64_t traj_per_piece = std::max<int64_t>(1, (int64_t)(stage_bytes / traj_bytes));
    auto move_out = [&](int k) -> int {   // device result buffer k % 2 -> staging -> the caller's arrays
        const int64_t lo = (int64_t)k * per_chunk, cnt = std::min<int64_t>(per_chunk, n_traj - lo);
        const char *dev = (const char *)((k % 2) ? c->out2.p : c->out.p);
        cudaStream_t cs = streams[k % 2];   // ordered behind chunk k's kernel; the next launch on it comes after this drain
        const int64_t n_pieces = (cnt + traj_per_piece - 1) / traj_per_piece;
        auto issue = [&](int64_t q) -> int {
            const int64_t p0 = q * traj_per_piece, pc = std::min<int64_t>(traj_per_piece, cnt - p0);
            CU(cudaMemcpyAsync(c->stage[q % 2], dev + traj_bytes * (size_t)p0, traj_bytes * (size_t)pc, cudaMemcpyDeviceToHost, cs));
            CU(cudaEventRecord(c->ev_copy[q % 2], cs));
            return QSIM_OK;
        };
        if (int rc = issue(0)) return rc;
        for (int64_t q = 0; q < n_pieces; q++) {
            if (q + 1 < n_pieces)
                if (int rc = issue(q + 1)) return rc;
            CU(cudaEventSynchronize(c->ev_copy[q % 2]));
            const int64_t p0 = q * traj_per_piece, pc = std::min<int64_t>(traj_per_piece, cnt - p0);
            const char *src = (const char *)c->stage[q % 2];
            char *dst = (char *)out + traj_bytes * (size_t)(sh.first + (lo + p0) * sh.step);
            if (sh.step == 1)
                std::memcpy(dst, src, traj_bytes * (size_t)pc);
            else
                for (int64_t j = 0; j < pc; j++) std::memcpy(dst + traj_bytes * (size_t)(j * sh.step), src + traj_bytes * (size_t)j, traj_bytes);
        }
        return QSIM_OK;
    };
    int last = -1;
    for (int k = 0; k < n_chunks; k++) {
        const int64_t lo = (int64_t)k * per_chunk, cnt = std::min<int64_t>(per_chunk, n_traj - lo);
        if (cnt <= 0) break;
        LaunchArgs ak = a;
        ak.n_traj = cnt;
        ak.in_first = sh.first + lo * sh.step;
        ak.out = (double2 *)((k % 2) ? c->out2.p : c->out.p);
        ak.stats = (qsim_traj_stats *)c->stats.p + lo;
        ak.work_counter = c->d_counter + (k % QSIM_N_COUNTERS);
        if (pl.streamed && (k % 2)) ak.scratch = (double2 *)c->scratch2.p;
        const int grid_k = (int)std::min<int64_t>(pl.grid, (cnt + pl.teams_per_cta - 1) / pl.teams_per_cta);
        CU(cudaMemsetAsync(ak.work_counter, 0, sizeof(unsigned long long), streams[k % 2]));
        if (k == 0) CU(cudaEventRecord(c->ev0, streams[0]));
        CU(pl.kv->launch(ak, grid_k, pl.block, pl.smem, streams[k % 2]));
        c->launches++;
        CU(cudaEventRecord(c->ev1, streams[k % 2]));   // (kept from the last chunk: end of the kernels)
        last = k;
        if (k >= 1)
            if (int rc = move_out(k - 1)) return rc;
    }
    c->timed = true;
    if (int rc = move_out(last)) return rc;
    if (stats)
        CU(cudaMemcpy2D(stats + sh.first, sizeof(qsim_traj_stats) * (size_t)sh.step, c->stats.p, sizeof(qsim_traj_stats),
                        sizeof(qsim_traj_stats), (size_t)n_traj, cudaMemcpyDeviceToHost));
    return QSIM_OK;
}

extern "C" {

int qsim_unitary_propagator(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj, const double *params,
                            const double *ts, int n_ts, int64_t ts_stride, const qsim_solver_opts *opts, double *out,
                            qsim_traj_stats *stats) {
    return solve_common(ctx, prob, 0, n_traj, params, ts, n_ts, ts_stride, nullptr, 0, opts, out, stats);
}
int qsim_unitary_state(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj, const double *params, const double *ts,
                       int n_ts, int64_t ts_stride, const double *psi0, int64_t init_stride,
                       const qsim_solver_opts *opts, double *out, qsim_traj_stats *stats) {
    return solve_common(ctx, prob, 1, n_traj, params, ts, n_ts, ts_stride, psi0, init_stride, opts, out, stats);
}
int qsim_me_propagator(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj, const double *params, const double *ts,
                       int n_ts, int64_t ts_stride, const qsim_solver_opts *opts, double *out,
                       qsim_traj_stats *stats) {
    return solve_common(ctx, prob, 2, n_traj, params, ts, n_ts, ts_stride, nullptr, 0, opts, out, stats);
}
int qsim_me_state(qsim_ctx *ctx, qsim_problem *prob, int64_t n_traj, const double *params, const double *ts,
                  int n_ts, int64_t ts_stride, const double *rho0, int64_t init_stride, const qsim_solver_opts *opts,
                  double *out, qsim_traj_stats *stats) {
    return solve_common(ctx, prob, 3, n_traj, params, ts, n_ts, ts_stride, rho0, init_stride, opts, out, stats);
}
int qsim_solve_selected(qsim_ctx *ctx, qsim_problem *prob, int which, int64_t n_traj, const double *params,
                        const double *ts, int n_ts, int64_t ts_stride, const double *init, int64_t init_stride,
                        const qsim_solver_opts *opts, int n_sel, const int32_t *sel_rows, const int32_t *sel_cols,
                        int abs2, double *out, qsim_traj_stats *stats) {
    if (which < 0 || which > 3) return fail(QSIM_ERR_ARG, "which must be 0..3");
    if (n_sel < 1 || !sel_rows) return fail(QSIM_ERR_ARG, "need at least one selected entry");
    if ((which == 0 || which == 2) && !sel_cols) return fail(QSIM_ERR_ARG, "propagator selections need sel_cols");
    if (!prob) return fail(QSIM_ERR_ARG, "prob is NULL");
    Selection sel;
    sel.n_sel = n_sel;
    sel.abs2 = abs2 ? 1 : 0;
    sel.rows = sel_rows;
    sel.cols = (which == 0 || which == 2) ? sel_cols : nullptr;
    std::vector<int32_t> lin;
    if (which == 3 && sel_cols) {
        // entries of the d x d density matrix: the solver's state is vec(rho), column-major
        lin.resize((size_t)n_sel);
        for (int i = 0; i < n_sel; i++) {
            if (sel_rows[i] < 0 || sel_rows[i] >= prob->d || sel_cols[i] < 0 || sel_cols[i] >= prob->d)
                return fail(QSIM_ERR_ARG, "selected entry out of range");
            lin[(size_t)i] = sel_cols[i] * prob->d + sel_rows[i];
        }
        sel.rows = lin.data();
    }
    return solve_common(ctx, prob, which, n_traj, params, ts, n_ts, ts_stride, (which == 1 || which == 3) ? init : nullptr,
                        init_stride, opts, out, stats, &sel);
}

// One sweep over several devices from one call (SURVEY.md section 8b "Threading": the fan-out happens inside
// the library so a Julia caller needs no threads).  Trajectories are dealt block-cyclically, shard i runs through
// ctxs[i] on its own host thread and delivers its results straight into the caller's buffers.
int qsim_solve_multi(qsim_ctx **ctxs, int n_ctx, qsim_problem *prob, int which, int64_t n_traj, const double *params,
                     const double *ts, int n_ts, int64_t ts_stride, const double *init, int64_t init_stride,
                     const qsim_solver_opts *opts, double *out, qsim_traj_stats *stats) {
    if (!ctxs || n_ctx < 1) return fail(QSIM_ERR_NO_DEVICE, "need at least one ctx (no CPU fallback)");
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]) return fail(QSIM_ERR_NO_DEVICE, "ctx is NULL: create one per device with qsim_create");
        for (int j = 0; j < i; j++)
            if (ctxs[j] == ctxs[i]) return fail(QSIM_ERR_ARG, "the same ctx appears twice");
    }
    if (which < 0 || which > 3) return fail(QSIM_ERR_ARG, "which must be 0..3");
    if (!prob) return fail(QSIM_ERR_ARG, "prob is NULL");
    if (opts && ((opts->flags & QSIM_FLAG_DEVICE_PTRS) || opts->stream))
        return fail(QSIM_ERR_ARG, "qsim_solve_multi takes host buffers and uses each ctx's own stream");
    if (n_traj < 0) return fail(QSIM_ERR_ARG, "n_traj < 0");
    const bool has_init = which == 1 || which == 3;
    if (n_ctx == 1 || n_traj <= 1)
        return solve_common(ctxs[0], prob, which, n_traj, params, ts, n_ts, ts_stride, has_init ? init : nullptr, init_stride,
                            opts, out, stats);
    if (n_ts < 1 || !ts) return fail(QSIM_ERR_ARG, "need at least one save time");
    if (ts_stride != 0 && ts_stride < n_ts) return fail(QSIM_ERR_ARG, "ts_stride must be 0 or >= n_ts");
    for (int64_t i = 0; i < (ts_stride ? n_traj : 1); i++)   // @assert issorted(ts), once for all shards
        for (int k = 1; k < n_ts; k++)
            if (!(ts[i * ts_stride + k - 1] <= ts[i * ts_stride + k])) return fail(QSIM_ERR_ARG, "ts must be sorted");
    // build the generator and check the plan once, on this thread (the workers then only read the problem)
    Plan pl;
    if (int rc = make_plan(prob, which, pl)) return rc;
    // Block-cyclic shards: ctx i integrates points i, i + n_ctx, ... and puts their results at
    // those positions of the caller's arrays -- in place when `out` is page-locked, otherwise through the pipelined
    // copy of solve_common, which hides the device-to-host transfer behind the next chunk's integration.
    std::vector<int> rcs((size_t)n_ctx, QSIM_OK);
    std::vector<std::string> msgs((size_t)n_ctx);
    std::vector<std::thread> workers;
    for (int i = 0; i < n_ctx; i++) {
        const int64_t cnt = (n_traj - i + n_ctx - 1) / n_ctx;
        if (cnt <= 0) continue;
        workers.emplace_back([=, &rcs, &msgs]() {
            Shard sh;
            sh.first = i;
            sh.step = n_ctx;
            sh.total = n_traj;
            rcs[(size_t)i] = solve_common(ctxs[i], prob, which, cnt, params, ts, n_ts, ts_stride, has_init ? init : nullptr,
                                          init_stride, opts, out, stats, nullptr, &sh);
            if (rcs[(size_t)i] != QSIM_OK) msgs[(size_t)i] = g_err;  // g_err is per thread
        });
    }
    for (auto &w : workers) w.join();
    for (int i = 0; i < n_ctx; i++)
        if (rcs[(size_t)i] != QSIM_OK) return fail(rcs[(size_t)i], "ctx " + std::to_string(i) + ": " + msgs[(size_t)i]);
    return QSIM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- FP64 peak probe
// 8 independent DFMA chains per thread, 2 flop per FMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *sink, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
            a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}

extern "C" int qsim_peak_fp64(qsim_ctx *c, double ms_target, double *tflops) {
    if (!c || !tflops) return fail(QSIM_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(c->device));
    double *sink = nullptr;
    CU(cudaMalloc(&sink, sizeof(double)));
    const int grid = c->prop.multiProcessorCount * 8, block = 256;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    int iters = 2000;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CU(cudaEventRecord(e0, c->stream));
        dfma_peak_kernel<<<grid, block, 0, c->stream>>>(sink, iters, 1.0);
        CU(cudaEventRecord(e1, c->stream));
        CU(cudaEventSynchronize(e1));
        c->launches++;
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)grid * block;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
        if (ms < ms_target && rep < 4) iters = (int)std::min<double>(2e8, iters * std::max(1.5, 0.9 * ms_target / std::max(ms, 1e-3f)));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops = best;
    return QSIM_OK;
}

// ---------------------------------------------------------------- hooks for qsim_floquet.cu
int qsim_internal_fail(int code, const std::string &msg) { return fail(code, msg); }
void qsim_internal_count_launch(qsim_ctx *c) { c->launches++; }
int qsim_internal_floquet_buffers(qsim_ctx *c, size_t seg_bytes, size_t work_bytes, size_t idx_bytes, void **seg, void **work,
                                  void **idx, cudaStream_t *stream, int *device) {
    CU(cudaSetDevice(c->device));
    CU(c->fl_seg.reserve(seg_bytes));
    CU(c->fl_work.reserve(work_bytes));
    CU(c->fl_idx.reserve(idx_bytes));
    *seg = c->fl_seg.p;
    *work = c->fl_work.p;
    *idx = c->fl_idx.p;
    *stream = c->stream;
    *device = c->device;
    return QSIM_OK;
}
int qsim_internal_solve_to_device(qsim_ctx *c, qsim_problem *p, int which, int64_t n_traj, const double *params, const double *ts,
                                  int n_ts, int64_t ts_stride, const qsim_solver_opts *opts, void *d_out, qsim_traj_stats *stats) {
    qsim_solver_opts o = *opts;
    o.flags |= QSIM_FLAG_INTERNAL_DEVICE_OUT;
    return solve_common(c, p, which, n_traj, params, ts, n_ts, ts_stride, nullptr, 0, &o, (double *)d_out, stats);
}
extern "C" int qsim_problem_dims(qsim_problem *p, int *d, int *n_params) {
    if (!p) return fail(QSIM_ERR_ARG, "prob is NULL");
    if (d) *d = p->d;
    if (n_params) *n_params = p->n_params;
    return QSIM_OK;
}
