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
    std::mutex plan_mutex;
};
