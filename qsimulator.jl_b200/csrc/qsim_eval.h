// Host+device evaluation of the drive descriptors (the channel evaluators).
//
// Same formulas as the reference evaluates inside its RHS closures:
//   drive shapes                      operators.jl:220-268 and the census in SURVEY.md App. B
//   xi_effective                      perturbative_transmon.jl:187-189
//   Horner with muladd                perturbative_transmon.jl:193-201
//   perturbative_transmon_freq/anharm perturbative_transmon.jl:222-231, 253-261
// Compiled for the device (CUDA libm: FP64 sin/cos/erf/exp/sqrt) and for the host, where the
// introspection entry qsim_problem_eval uses it so the tests can check it without a GPU.
#pragma once
#include <math.h>

#include "qsim_types.h"

namespace qsim {

#define QSIM_TWO_PI 6.283185307179586

struct SeriesTables {
    double freq[31];
    double anh[31];
};

QSIM_HD double slot_value(const qsim_slot &s, const double *row) { return s.param >= 0 ? row[s.param] : s.value; }

QSIM_HD double eval_signal(const qsim_signal &s, double t, const double *row) {
    double car = 1.0, env = 1.0;
    if (s.carrier != QSIM_CARRIER_NONE) {
        double arg = (QSIM_TWO_PI * slot_value(s.freq, row)) * t + slot_value(s.phase, row);
        car = s.carrier == QSIM_CARRIER_COS ? cos(arg) : sin(arg);
    }
    if (s.envelope == QSIM_ENV_GAUSSIAN) {
        double x = (t - slot_value(s.t1, row)) / slot_value(s.sigma, row);
        env = exp(-0.5 * (x * x));
    } else if (s.envelope == QSIM_ENV_ERF_SQUARED) {
        double sg = slot_value(s.sigma, row);
        env = 0.5 * (erf((t - slot_value(s.t1, row)) / sg) - erf((t - slot_value(s.t2, row)) / sg));
    }
    if (s.blend == QSIM_BLEND_NONE) return slot_value(s.offset, row) + slot_value(s.amp, row) * env * car;
    return env * (slot_value(s.offset, row) + slot_value(s.amp, row) * car) + (1.0 - env) * slot_value(s.rest, row);
}

QSIM_HD bool signal_is_dynamic(const qsim_signal &s) {
    return s.carrier != QSIM_CARRIER_NONE || s.envelope != QSIM_ENV_NONE;
}

QSIM_HD double horner(double x, const double *c, int n) {
    if (n <= 0) return 0.0;
    double ex = c[n - 1];
    for (int i = n - 2; i >= 0; i--) ex = fma(x, ex, c[i]);
    return ex;
}

QSIM_HD void transmon_f_a(const SeriesTables &tab, double EC, double EJ1, double EJ2, double phi, int num_terms,
                          double *f, double *a) {
    int t = num_terms < 32 ? num_terms : 32;
    double xi = sqrt((2.0 * EC) / sqrt((EJ1 * EJ1 + EJ2 * EJ2) + (2.0 * EJ1 * EJ2) * cos(QSIM_TWO_PI * phi)));
    *f = EC * (4.0 / xi + horner(xi, tab.freq, t - 1));
    *a = EC * horner(xi, tab.anh, t - 1);
}

// Evaluate one evaluator at time t into chan[] (indexed by absolute channel number).
QSIM_HD void run_evaluator(const EvalDesc &e, const SeriesTables &tab, double t, const double *row, double *chan) {
    switch (e.kind) {
        case EV_SIGNAL:
            chan[e.out_chan] = eval_signal(e.sig, t, row);
            break;
        case EV_ROTATING: {
            double s = eval_signal(e.sig, t, row);
            double th = QSIM_TWO_PI * (slot_value(e.rate, row) * t);
            double sn, cs;
            sincos(th, &sn, &cs);
            chan[e.out_chan] = s * cs;
            chan[e.out_chan + 1] = s * sn;
            break;
        }
        case EV_TRANSMON: {
            double f, a;
            transmon_f_a(tab, slot_value(e.EC, row), slot_value(e.EJ1, row), slot_value(e.EJ2, row),
                         eval_signal(e.sig, t, row), e.num_terms, &f, &a);
            chan[e.out_chan] = f;
            chan[e.out_chan + 1] = a;
            break;
        }
        case EV_SQUARE: {
            double s = eval_signal(e.sig, t, row);
            chan[e.out_chan] = s * s;
            break;
        }
        default:
            chan[e.out_chan] = slot_value(e.slot, row);
            break;
    }
}

}  // namespace qsim
