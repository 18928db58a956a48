// Internal types shared by the host-side generator compiler and the CUDA kernels.
//
// Every solver call integrates  du/dt = A(t) u  column by column, where
//   unitary_*:  A(t) = -2 pi i H(t)                         (n = d)      time_evolution.jl:31-37
//   me_*:       A(t) = -2 pi i (I (x) H - H^T (x) I)
//                      + 2 pi sum_L (conj(L)(x)L - 1/2 I(x)L'L - 1/2 (L'L)^T(x)I)   (n = d^2)   :98-111
// A(t) is stored as a fixed sparse pattern (ELL rows of (source index, value slot)) plus a table of
// value slots, each a short linear combination of real "channels" c_k(t):
//   V[slot](t) = sum_j weight_j * c_{chan_j}(t)
// Channels are produced by evaluators: one per time-dependent term (qsim_term) or Lindblad rate.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <vector>
#include <complex>

#include "../../include/qsim_b200.h"

#if defined(__CUDACC__)
#define QSIM_HD __host__ __device__ __forceinline__
#else
#define QSIM_HD inline
#endif

namespace qsim {

typedef std::complex<double> cplx;

enum EvalKind : int32_t {
    EV_SIGNAL = 0,    // out[0] = s(t)
    EV_ROTATING = 1,  // out[0] = s cos(th), out[1] = s sin(th), th = 2 pi (rate t)
    EV_TRANSMON = 2,  // out[0] = f01(phi), out[1] = anharm(phi), phi = s(t)
    EV_SQUARE = 3,    // out[0] = s(t)^2            (Lindblad rate |scale|^2)
    EV_SLOT = 4,      // out[0] = slot value        (per-trajectory constant)
    EV_SPECTRUM = 5   // out[0] = table(cos(pi s(t)))   (a level of a DiagonalChargeBasisTransmon; num_terms = table ref)
};

// One channel evaluator (plain data, copied to the device as-is).
struct EvalDesc {
    int32_t kind;
    int32_t out_chan;     // first output channel
    int32_t dynamic;      // depends on t (else evaluated once per trajectory)
    int32_t num_terms;    // TRANSMON
    qsim_signal sig;
    qsim_slot rate;       // ROTATING
    qsim_slot EC, EJ1, EJ2;  // TRANSMON
    qsim_slot slot;       // SLOT
};

// The compiled generator for one state kind (n = d or d^2).
struct Generator {
    int n = 0;                 // rows of A
    int max_nnz = 0;           // longest ELL row
    int n_slots = 0;           // value slots
    int n_dyn_slots = 0;       // slots [0, n_dyn_slots) depend on a dynamic channel
    int n_chan = 0;            // channels; channel 0 is the constant 1
    std::vector<uint32_t> ell;     // [max_nnz][n]: src | slot << 16   (0xFFFFFFFF = padding)
    std::vector<uint8_t> row_len;  // [n]
    std::vector<int32_t> slot_ptr; // [n_slots + 1] into pair arrays
    std::vector<int32_t> pair_chan;
    std::vector<cplx> pair_w;
    std::vector<EvalDesc> evals;
    // tabulated envelopes, flat: per table {t_start, 1/h, n_seg, n_coef, coef[n_seg][n_coef]}; an evaluator's
    // sig.table holds 1 + the offset of its table in this array (0: none)
    std::vector<double> tables;
    int64_t nnz = 0;           // total stored entries (executed complex MACs per column per RHS)
    bool imag_only = false;    // every value slot is purely imaginary (A = i B, B real): real H, no dissipator
};

// Component form of a propagator problem (u0 = I).  Column c of the propagator is supported on the connected
// component of c in the generator's graph (union of the patterns at all times); the entries outside are exactly zero
// in the reference as well (time_evolution.jl:31-37, 98-111: dense products that only add 0 * x there) and add
// nothing to the error norms, so the kernels integrate only the component blocks: a UNIT is one component together
// with `cpl` of its columns, units are packed into the warps of a team (whole units, never split), and every warp
// sees a small virtual system over its local rows.  build_component_form() in qsim_components.cpp.
struct ComponentForm {
    bool ok = false;
    int n_comp = 0;
    int rpl = 1, cpl = 1;              // rows / columns per lane of the kernel shape the form was packed for
    int n_loc = 0;                     // local rows per warp (the kernel's n): at most 32 * rpl
    int team_warps = 1, n_batches = 1, streamed = 0;
    int max_nnz = 0;
    int stride = 1;                    // stage-buffer row stride in 16-byte units (chosen with the layout)
    std::vector<uint32_t> ell;         // [max_nnz][team_warps * n_loc]: local source row | slot << 16
    std::vector<int> vrow;             // [team_warps * n_loc]: state row of a local row, -1: none
    std::vector<int> vcol;             // [n_batches * n_loc * cpl]: state column of (batch, local row, c), -1: none
    std::vector<int> zeros;            // col * n + row of every structurally zero entry of the state
    int64_t nnz_cols = 0;              // sum over components of (stored entries x columns): executed complex MACs per RHS
    int64_t entries = 0;               // structurally non-zero entries of the state
    std::vector<int> comp_of;          // [n] component of every row
    Generator gv;                      // the generator as make_plan sees it: n = n_loc, ell = the per-warp table, slots and evaluators unchanged
    // Row-split form (rs = 1: state vectors longer than 96 entries or rows longer than 16 entries).  A batch is a PASS:
    // up to 3 * team_warps chunks of 32 rows taken from several units (component, ONE of its columns) at once.  Every
    // component keeps its rows sorted by length in chunks of 32 (`subs[k]`: the component's own generator, local row
    // numbers; the row tables are built from it by qsim_api.cu once the value mode is known); pass[4 * (b * 3 *
    // team_warps + c)] = {component chunk, position of the unit in the stage buffer, state column, 0}; component chunk
    // `n_cc` is the empty chunk.
    int rs = 0, n_cc = 0;
    // State form (state = 1; unitary_state / me_state): the single column is the initial state, supported on the
    // components it touches (live[r] = 1 for their rows); everything else stays exactly zero.  Same tables, one unit per
    // touched component with column 0.
    int state = 0;
    std::vector<char> live;
    std::vector<Generator> subs;       // per component: generator restricted to it
    std::vector<std::vector<int>> comp_rows;   // per component: its state rows (local row -> state row)
    std::vector<int> cc_first;         // per component: its first component chunk
    std::vector<int> pass;
};
bool build_component_form(const Generator &g, int rpl, int cpl, int max_team_warps, ComponentForm &cf);
bool build_rowsplit_component_form(const Generator &g, int team_warps, ComponentForm &cf, const std::vector<char> *touched = nullptr,
                                      int rpl = 3);
bool build_state_form(const Generator &g, const std::vector<char> &touched, ComponentForm &cf);
void optimize_component_layout(ComponentForm &cf);
long component_layout_cost(const ComponentForm &cf, int S);
void generator_components(const Generator &g, std::vector<int> &comp_of, std::vector<std::vector<int>> &comps);

}  // namespace qsim

struct qsim_problem {
    int d = 0, n_params = 0;
    bool finalized = false;
    std::vector<qsim::cplx> h0;
    struct Term {
        qsim_term desc;
        std::vector<qsim::cplx> a, b;
        std::vector<int32_t> levels;
        std::vector<int32_t> tables;   // SPECTRUM: table id of every level
    };
    struct Lind {
        std::vector<qsim::cplx> op;
        bool has_scale;
        qsim_signal scale;
    };
    std::vector<Term> terms;
    std::vector<Lind> linds;
    std::vector<double> tables;          // flat table storage (layout as Generator::tables)
    std::vector<int32_t> table_offset;   // table id -> offset into `tables`
    qsim::Generator gen[2];  // [0] unitary (n = d), [1] Lindblad (n = d^2)
    bool gen_built[2] = {false, false};
    // device copies of the generators, one per CUDA device that has solved this problem (shared by the
    // contexts on that device); created under `dev_mutex`
    struct DeviceCopy;
    std::vector<DeviceCopy *> dev[2];
    std::mutex dev_mutex;
    // tiered form of each generator (qsim_api.cu, build_tiered): built on first use, under plan_mutex
    std::shared_ptr<struct TieredForm> tiered[2];
    std::shared_ptr<struct AddSplitForm> addsplit[2];   // additively split rows (build_addsplit)
    std::shared_ptr<qsim::ComponentForm> cform[2];      // component form of the propagator problems (make_plan)
    bool cform_tried[2] = {false, false};
    // state forms by the set of components the initial state touches (solve_common), and the components themselves
    std::vector<std::pair<std::vector<char>, std::shared_ptr<qsim::ComponentForm>>> sforms[2];
    std::vector<int> comp_of[2];
    int n_comp[2] = {-1, -1};
    std::mutex plan_mutex;
};
