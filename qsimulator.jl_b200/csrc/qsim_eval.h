// Host+device evaluation of the drive descriptors (the channel evaluators).
//
// Same formulas as the reference evaluates inside its RHS closures:
//   drive shapes                      operators.jl:220-268 and the census in SURVEY.md App. B
//   xi_effective                      perturbative_transmon.jl:187-189
//   Horner with muladd                perturbative_transmon.jl:193-201
//   perturbative_transmon_freq/anharm perturbative_transmon.jl:222-231, 253-261
// An evaluator is first RESOLVED against a trajectory's parameter row (slots -> numbers, the
// time-independent products 2*pi*freq, 2*EC, EJ1^2+EJ2^2, 2*EJ1*EJ2 formed once, in the same
// operation order as the reference) and then RUN at each stage time.  Both steps compile for the
// device (CUDA libm: FP64 sin/cos/erf/exp/sqrt) and for the host, where the introspection entry
// qsim_problem_eval uses them so the tests can check them without a GPU.
#pragma once
#include <math.h>

#include "qsim_types.h"

namespace qsim {

#define QSIM_TWO_PI 6.283185307179586

struct SeriesTables {
    double freq[31];
    double anh[31];
};

// One evaluator with its slots resolved for a trajectory (20 doubles; lives in shared memory).
struct ResolvedEval {
    int32_t kind, carrier, envelope, blend, num_terms, out_chan, dynamic, _pad;
    double offset, amp, w, phase, t1, t2, sigma, rest;  // w = 2*pi*freq
    double rate;                                        // ROTATING
    double EC, EC2, EJs, EJp;                           // TRANSMON: EC, 2*EC, EJ1^2+EJ2^2, 2*EJ1*EJ2
    double slot;                                        // SLOT
    const double *tab;                                  // ENV_TABLE: {t_start, 1/h, n_seg, n_coef, coef...}
    const double *tab2;                                 // SPECTRUM: the level's table E(z), z = cos(pi phi)
};
static_assert(sizeof(ResolvedEval) == 160, "ResolvedEval is copied as ten 16-byte words");

// Tabulated envelope: segment lookup + Clenshaw (qsim_problem_add_table).
QSIM_HD double eval_table(const double *T, double t) {
    const int n_seg = (int)T[2], nc = (int)T[3];
    const double u = (t - T[0]) * T[1];
    int k = (int)floor(u);
    k = k < 0 ? 0 : (k >= n_seg ? n_seg - 1 : k);
    double x = 2.0 * (u - (double)k) - 1.0;
    x = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);   // constant continuation outside [t_start, t_end]
    const double *c = T + 4 + (size_t)k * nc;
    double b1 = 0.0, b2 = 0.0;
    const double x2 = 2.0 * x;
    for (int j = nc - 1; j >= 1; j--) {
        const double b0 = fma(x2, b1, c[j] - b2);
        b2 = b1;
        b1 = b0;
    }
    return fma(x, b1, c[0] - b2);
}

QSIM_HD double slot_value(const qsim_slot &s, const double *row) { return s.param >= 0 ? row[s.param] : s.value; }

QSIM_HD bool signal_is_dynamic(const qsim_signal &s) {
    return s.carrier != QSIM_CARRIER_NONE || s.envelope != QSIM_ENV_NONE;
}

QSIM_HD void resolve_eval(const EvalDesc &e, const double *row, const double *tables, ResolvedEval &r) {
    r.kind = e.kind;
    r.carrier = e.sig.carrier;
    r.envelope = e.sig.envelope;
    r.blend = e.sig.blend;
    r.num_terms = e.num_terms;
    r.out_chan = e.out_chan;
    r.dynamic = e.dynamic;
    r._pad = 0;
    r.offset = slot_value(e.sig.offset, row);
    r.amp = slot_value(e.sig.amp, row);
    r.w = QSIM_TWO_PI * slot_value(e.sig.freq, row);
    r.phase = slot_value(e.sig.phase, row);
    r.t1 = slot_value(e.sig.t1, row);
    r.t2 = slot_value(e.sig.t2, row);
    r.sigma = slot_value(e.sig.sigma, row);
    r.rest = slot_value(e.sig.rest, row);
    r.rate = slot_value(e.rate, row);
    const double EC = slot_value(e.EC, row), EJ1 = slot_value(e.EJ1, row), EJ2 = slot_value(e.EJ2, row);
    r.EC = EC;
    r.EC2 = 2.0 * EC;
    r.EJs = EJ1 * EJ1 + EJ2 * EJ2;
    r.EJp = 2.0 * EJ1 * EJ2;
    r.slot = slot_value(e.slot, row);
    r.tab = (e.sig.envelope == QSIM_ENV_TABLE && e.sig.table > 0 && tables) ? tables + (e.sig.table - 1) : nullptr;
    r.tab2 = (e.kind == EV_SPECTRUM && e.num_terms > 0 && tables) ? tables + (e.num_terms - 1) : nullptr;
}

QSIM_HD double eval_signal(const ResolvedEval &s, double t) {
    double car = 1.0, env = 1.0;
    if (s.carrier != QSIM_CARRIER_NONE) {
        const double arg = s.w * t + s.phase;
        car = s.carrier == QSIM_CARRIER_COS ? cos(arg) : sin(arg);
    }
    if (s.envelope == QSIM_ENV_GAUSSIAN) {
        const double x = (t - s.t1) / s.sigma;
        env = exp(-0.5 * (x * x));
    } else if (s.envelope == QSIM_ENV_ERF_SQUARED) {
        env = 0.5 * (erf((t - s.t1) / s.sigma) - erf((t - s.t2) / s.sigma));
    } else if (s.envelope == QSIM_ENV_TABLE) {
        env = s.tab ? eval_table(s.tab, t) : 0.0;
    }
    if (s.blend == QSIM_BLEND_NONE) return s.offset + s.amp * env * car;
    return env * (s.offset + s.amp * car) + (1.0 - env) * s.rest;
}

// Two Horner chains interleaved (frequency and anharmonicity series share xi).
QSIM_HD void transmon_f_a(const ResolvedEval &r, const double *freq_tab, const double *anh_tab, double phi, double *f,
                          double *a) {
    const int nt = (r.num_terms < 32 ? r.num_terms : 32) - 1;
    const double xi = sqrt(r.EC2 / sqrt(r.EJs + r.EJp * cos(QSIM_TWO_PI * phi)));
    double pf = 0.0, pa = 0.0;
    if (nt > 0) {
        pf = freq_tab[nt - 1];
        pa = anh_tab[nt - 1];
        for (int i = nt - 2; i >= 0; i--) {
            pf = fma(xi, pf, freq_tab[i]);
            pa = fma(xi, pa, anh_tab[i]);
        }
    }
    *f = r.EC * (4.0 / xi + pf);
    *a = r.EC * pa;
}

// Run one resolved evaluator at time t: up to two channel values.
QSIM_HD void run_resolved(const ResolvedEval &e, const double *freq_tab, const double *anh_tab, double t, double *o0,
                          double *o1) {
    *o1 = 0.0;
    switch (e.kind) {
        case EV_SIGNAL:
            *o0 = eval_signal(e, t);
            break;
        case EV_ROTATING: {
            const double s = eval_signal(e, t);
            const double th = QSIM_TWO_PI * (e.rate * t);
            double sn, cs;
            sincos(th, &sn, &cs);
            *o0 = s * cs;
            *o1 = s * sn;
            break;
        }
        case EV_TRANSMON:
            transmon_f_a(e, freq_tab, anh_tab, eval_signal(e, t), o0, o1);
            break;
        case EV_SQUARE: {
            const double s = eval_signal(e, t);
            *o0 = s * s;
            break;
        }
        case EV_SPECTRUM:
            *o0 = e.tab2 ? eval_table(e.tab2, cos(0.5 * QSIM_TWO_PI * eval_signal(e, t))) : 0.0;
            break;
        default:
            *o0 = e.slot;
            break;
    }
}

QSIM_HD int eval_outputs(int kind) { return (kind == EV_ROTATING || kind == EV_TRANSMON) ? 2 : 1; }

}  // namespace qsim
