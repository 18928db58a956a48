/*
 * Plain-C consumer of include/qsim_b200.h: what a non-Python host (the Julia shim's ccall sequence,
 * INTEGRATION.md section 2) does, with no interpreter in the way.
 *
 *   abi_driver host   -- no device needed: assemble a problem, check qsim_problem_eval against the closed
 *                        form, check that a solve without a device fails with QSIM_ERR_NO_DEVICE or runs.
 *   abi_driver gpu    -- needs a CUDA device: Rabi-driven Duffing transmon (test/integration_tests.jl:32-64
 *                        system) swept over the drive amplitude through qsim_unitary_propagator, checked for
 *                        unitarity, identity at ts[0], and the rotating-wave population
 *                        P_g(t) = cos^2(pi * amp * t) within the reference's 1e-2.
 * Exit status 0 = all checks passed; prints one line per check.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qsim_b200.h"

#define D 3
static int failures = 0;

static void check(int ok, const char *what) {
    printf("%s %s\n", ok ? "ok  " : "FAIL", what);
    if (!ok) failures++;
}

static qsim_slot lit(double v) {
    qsim_slot s;
    s.param = -1;
    s._pad = 0;
    s.value = v;
    return s;
}

static qsim_slot par(int i) {
    qsim_slot s;
    s.param = i;
    s._pad = 0;
    s.value = 0.0;
    return s;
}

/* H0 = duffing_hamiltonian(5.0, -0.2, 3) (src/systems.jl:189-206); drive amp*cos(2 pi 5 t) on X = a + a'
 * (dipole_drive, src/operators.jl:220-226); amp is sweep parameter 0. */
static qsim_problem *build_problem(void) {
    qsim_problem *p = NULL;
    double h0[2 * D * D], x[2 * D * D];
    const double f = 5.0, a = -0.2;
    qsim_term term;
    int n;
    if (qsim_problem_create(D, 1, &p) != QSIM_OK) return NULL;
    memset(h0, 0, sizeof h0);
    memset(x, 0, sizeof x);
    for (n = 0; n < D; n++) h0[2 * (n * D + n)] = (f - 0.5 * a) * n + 0.5 * a * n * n;
    for (n = 1; n < D; n++) {
        x[2 * ((n - 1) * D + n)] = sqrt((double)n); /* raising: [n, n-1] column-major = (n-1)*D + n */
        x[2 * (n * D + (n - 1))] = sqrt((double)n); /* lowering */
    }
    memset(&term, 0, sizeof term);
    term.kind = QSIM_TERM_SCALAR;
    term.signal.carrier = QSIM_CARRIER_COS;
    term.signal.envelope = QSIM_ENV_NONE;
    term.signal.blend = QSIM_BLEND_NONE;
    term.signal.offset = lit(0.0);
    term.signal.amp = par(0);
    term.signal.freq = lit(f);
    term.signal.phase = lit(0.0);
    term.signal.t1 = term.signal.t2 = term.signal.rest = lit(0.0);
    term.signal.sigma = lit(1.0);
    term.rate = term.anharm = term.rot = term.EC = term.EJ1 = term.EJ2 = lit(0.0);
    if (qsim_problem_set_h0(p, h0) != QSIM_OK) return NULL;
    if (qsim_problem_add_term(p, &term, x, NULL, NULL) != QSIM_OK) return NULL;
    if (qsim_problem_finalize(p) != QSIM_OK) return NULL;
    return p;
}

static int host_checks(qsim_problem *p) {
    double h[2 * D * D], row[1] = {0.02}, t = 0.37, want;
    char name[128];
    double flops = 0;
    check(qsim_version() == QSIM_ABI_VERSION, "qsim_version == QSIM_ABI_VERSION");
    check(qsim_problem_eval(p, 0, t, row, h) == QSIM_OK, "qsim_problem_eval");
    want = 0.02 * cos(2.0 * M_PI * 5.0 * t);
    check(fabs(h[2 * (0 * D + 1)] - want) < 1e-15 && fabs(h[2 * (1 * D + 2)] - want * sqrt(2.0)) < 1e-15 &&
              fabs(h[2 * (1 * D + 1)] - 5.0) < 1e-15 && fabs(h[2 * (2 * D + 2)] - 9.8) < 1e-14,
          "H(t) = H0 + amp cos(2 pi f t) (a + a')");
    check(qsim_problem_kernel_name(p, 0, name, (int)sizeof name) == QSIM_OK && strstr(name, "tsit5_") == name,
          "qsim_problem_kernel_name");
    check(qsim_problem_flops_per_rhs(p, 0, &flops) == QSIM_OK && flops > 0, "qsim_problem_flops_per_rhs");
    check(qsim_problem_add_term(p, NULL, NULL, NULL, NULL) != QSIM_OK && strlen(qsim_last_error()) > 0,
          "argument error reported through qsim_last_error");
    return failures;
}

static int gpu_checks(qsim_problem *p) {
    enum { NT = 64, NS = 51 };
    qsim_ctx *ctx = NULL;
    qsim_solver_opts opts;
    double *params = NULL, *out = NULL, ts[NS];
    qsim_traj_stats *stats = malloc(sizeof(qsim_traj_stats) * NT);
    int i, k, r, c, j, rc;
    double worst_unit = 0, worst_pop = 0, worst_id = 0;
    rc = qsim_create(0, &ctx);
    check(rc == QSIM_OK, "qsim_create(0)");
    if (rc != QSIM_OK) {
        printf("     %s\n", qsim_last_error());
        return failures;
    }
    qsim_default_opts(&opts);
    check(opts.reltol == 1e-6 && opts.abstol == 1e-8, "default tolerances are the reference literals");
    check(qsim_alloc_pinned(sizeof(double) * NT, (void **)&params) == QSIM_OK, "qsim_alloc_pinned(params)");
    check(qsim_alloc_pinned(sizeof(double) * 2 * D * D * NS * NT, (void **)&out) == QSIM_OK, "qsim_alloc_pinned(out)");
    for (i = 0; i < NT; i++) params[i] = 0.01 + 0.0002 * i; /* nutation frequencies 10 .. 22.6 MHz */
    for (k = 0; k < NS; k++) ts[k] = (double)k;
    rc = qsim_unitary_propagator(ctx, p, NT, params, ts, NS, 0, &opts, out, stats);
    check(rc == QSIM_OK, "qsim_unitary_propagator");
    if (rc != QSIM_OK) printf("     %s\n", qsim_last_error());
    check(qsim_launch_count(ctx) >= 1, "qsim_launch_count counts the kernel launch");
    for (i = 0; i < NT && rc == QSIM_OK; i++) {
        if (stats[i].retcode != QSIM_RET_SUCCESS || stats[i].t_final != ts[NS - 1]) failures++;
        for (k = 0; k < NS; k++) {
            const double *u = out + 2 * (size_t)D * D * ((size_t)i * NS + k);
            for (r = 0; r < D; r++)
                for (c = 0; c < D; c++) {
                    double re = 0, im = 0; /* (U' U)[r][c] */
                    for (j = 0; j < D; j++) {
                        const double ar = u[2 * (r * D + j)], ai = u[2 * (r * D + j) + 1];
                        const double br = u[2 * (c * D + j)], bi = u[2 * (c * D + j) + 1];
                        re += ar * br + ai * bi;
                        im += ar * bi - ai * br;
                    }
                    re -= (r == c);
                    if (fabs(re) > worst_unit) worst_unit = fabs(re);
                    if (fabs(im) > worst_unit) worst_unit = fabs(im);
                    if (k == 0) {
                        const double e = fabs(u[2 * (c * D + r)] - (r == c)) + fabs(u[2 * (c * D + r) + 1]);
                        if (e > worst_id) worst_id = e;
                    }
                }
            {
                /* ground-state population of the driven qubit vs the rotating-wave answer; the third level and
                 * counter-rotating terms limit the agreement to ~1e-2 (test/integration_tests.jl:51) */
                const double pg = u[0] * u[0] + u[1] * u[1];
                const double want = pow(cos(M_PI * params[i] * ts[k]), 2.0);
                if (fabs(pg - want) > worst_pop) worst_pop = fabs(pg - want);
            }
        }
    }
    printf("     unitarity %.2e, identity at ts[0] %.2e, population vs RWA %.2e\n", worst_unit, worst_id, worst_pop);
    check(worst_id == 0.0, "U(ts[0]) == I exactly");
    check(worst_unit < 1e-3, "U'U = I within the solver tolerance");
    check(worst_pop < 3e-2, "Rabi population matches cos^2(pi amp t)");
    check(qsim_free_pinned(params) == QSIM_OK && qsim_free_pinned(out) == QSIM_OK, "qsim_free_pinned");
    check(qsim_destroy(ctx) == QSIM_OK, "qsim_destroy");
    free(stats);
    return failures;
}

int main(int argc, char **argv) {
    qsim_problem *p = build_problem();
    const char *mode = argc > 1 ? argv[1] : "host";
    if (!p) {
        printf("FAIL problem assembly: %s\n", qsim_last_error());
        return 1;
    }
    host_checks(p);
    if (!strcmp(mode, "gpu")) {
        gpu_checks(p);
    } else {
        /* without a device the solver must refuse (no CPU fallback); with one it must work */
        qsim_ctx *ctx = NULL;
        const int rc = qsim_create(0, &ctx);
        check(rc == QSIM_OK || rc == QSIM_ERR_NO_DEVICE, "qsim_create: device or QSIM_ERR_NO_DEVICE, nothing else");
        if (rc == QSIM_OK) qsim_destroy(ctx);
    }
    check(qsim_problem_destroy(p) == QSIM_OK, "qsim_problem_destroy");
    printf("%s\n", failures ? "FAILED" : "PASSED");
    return failures ? 1 : 0;
}
