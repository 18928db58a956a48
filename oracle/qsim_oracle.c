/*
 * qsim_oracle.c -- CPU restatement of the reference's time-evolution path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs may load this library; the product path
 * (qsimulator.jl_b200/csrc) never links, loads or calls it.
 *
 * PARITY UNPINNED at the 1e-8 level.  Julia is not installed in this image, so
 * the reference cannot run, and its integrator is the un-vendored third-party
 * package OrdinaryDiffEq (compat "5.29", with DiffEqBase "6.20";
 * /root/reference/Project.toml:8,11,18,20; no Manifest.toml).  This file
 * restates
 *   - the four right-hand sides as the reference writes them:
 *       unitary_propagator /root/reference/src/time_evolution.jl:31-37
 *       unitary_state      /root/reference/src/time_evolution.jl:65-71
 *       me_propagator      /root/reference/src/time_evolution.jl:98-111
 *       me_state           /root/reference/src/time_evolution.jl:142-154
 *   - per-RHS Hamiltonian assembly H(t) = H0 + sum_k embed(f_k(t))
 *       /root/reference/src/composite_systems.jl:46-50,65-72
 *     with the time-dependence forms of operators.jl:220-268, systems.jl:189-265
 *     and perturbative_transmon.jl:187-261 expressed through the descriptor
 *     family of include/qsim_b200.h
 *   - OrdinaryDiffEq's published Tsit5 algorithm with the options of
 *     time_evolution.jl:43 (saveat=ts, save_start=true, reltol=1e-6,
 *     abstol=1e-8): tableau, RMS error norm, PI controller, Hairer-Wanner
 *     initial step, free 4th-order interpolant.
 * It is pinned by the reference's known-answer tests, the Runge-Kutta order
 * conditions and the independent closure-driven numpy restatement
 * (oracle/np_oracle.py); see tests/test_oracle_*.py.
 *
 * mode 0 ("faithful"): dense assembly + dense GEMM; me_propagator forms the
 *   d^2 x d^2 Kronecker superoperators and does (1 + n_L) GEMMs per RHS call,
 *   like time_evolution.jl:104,109.  This is the timed CPU baseline.
 * mode 1 ("structured"): me_propagator applies the commutator / sandwich form
 *   column by column (mathematically identical, O(d^5) instead of O(d^6)); used
 *   to keep large parity cases within seconds.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/qsim_b200.h"

#define TWO_PI 6.283185307179586

/* ------------------------------------------------------------------ tables */
/* perturbative_transmon.jl:38-105, value = numerator / 2^b (filled at load)  */
static const char *FREQ_NUM[31] = {
    "-1", "-1", "-21", "-19", "-5319", "-6649", "-1180581", "-446287", "-1489138635", "-648381403",
    "-614557854099", "-75265839129", "-637411859250147", "-86690561488017", "-405768570324517701",
    "-15191635582891041", "-2497063196283456607731", "-102281923716042917215",
    "-2292687293949773041433127", "-25544408245062216574759", "-4971071120163260007203175705",
    "-59956026877695226936825271", "-6299936888270974385982624367587", "-20465345194746565030172477629",
    "-36984324599399309412347250837528543", "-128862667153189778842334459173303",
    "-62313306363032484263243187605857135455", "-57979140436623262897403437875845329",
    "-239123145585215826671902664445701932931163", "-236766982175345008541229542667501202161",
    "-518667194120793070334115565427490753019904133"};
static const int FREQ_EXP[31] = {0, 2, 7, 7, 15, 15, 22, 20, 31, 29, 38, 34, 46, 42, 53, 47, 63, 57, 70, 62, 78,
                                 70, 85, 75, 94, 84, 101, 89, 109, 97, 116};
static const char *ANH_NUM[31] = {
    "-1", "-9", "-81", "-3645", "-46899", "-1329129", "-20321361", "-2648273373", "-45579861135",
    "-1647988255539", "-31160327412879", "-2457206583272505", "-50387904068904927", "-2145673984043982897",
    "-47368663010124907041", "-17329540083222030375645", "-410048712835835979799431",
    "-20066784213453521778111375", "-507447585299180759749453827", "-53019019946496461235728807475",
    "-1429754157181172012054040903645", "-79571741391885949104006842758911",
    "-2283773190022904454409743892590327", "-540565733415401595950277192471356985",
    "-16479511149218202447739080120870460083", "-1034743270413623494225962156243473940687",
    "-33436402163767100825528499521512789823595", "-4445866752212713247387096882061024305480817",
    "-151943275517394187333713519962683890787229463", "-10671890872277478133986830879693284605566580129",
    "-384886904723357410697832985058012690322303378753"};
static const int ANH_EXP[31] = {0, 4, 7, 12, 15, 19, 22, 28, 31, 35, 38, 43, 46, 50, 53, 60, 63, 67, 70, 75, 78,
                                82, 85, 91, 94, 98, 101, 106, 109, 113, 116};
static double PT_FREQ[31], PT_ANH[31];
static pthread_once_t tables_once = PTHREAD_ONCE_INIT;
static void init_tables(void) {
    for (int i = 0; i < 31; i++) {
        PT_FREQ[i] = ldexp(strtod(FREQ_NUM[i], NULL), -FREQ_EXP[i]); /* strtod is correctly rounded */
        PT_ANH[i] = ldexp(strtod(ANH_NUM[i], NULL), -ANH_EXP[i]);
    }
}

/* ---------------------------------------------------------------- problem */
typedef struct {
    qsim_term desc;
    double *atom_a, *atom_b; /* d*d complex or NULL */
    int32_t *levels;         /* d or NULL */
    int n_levels;            /* SPECTRUM: dimension of the transmon (levels kept of the charge-basis spectrum) */
} o_term;
typedef struct {
    double *op; /* d*d complex */
    int has_scale;
    qsim_signal scale;
} o_lind;
typedef struct qso_problem {
    int d, n_params;
    double *h0;
    int n_terms, n_lind;
    o_term *terms;
    o_lind *linds;
    /* tabulated envelopes (qsim_problem_add_table): per table {t_start, 1/h, n_seg, n_coef, coef...} */
    int n_tables;
    double **tables;
} qso_problem;

qso_problem *qso_problem_create(int d, int n_params) {
    pthread_once(&tables_once, init_tables);
    qso_problem *p = (qso_problem *)calloc(1, sizeof(*p));
    p->d = d;
    p->n_params = n_params;
    p->h0 = (double *)calloc((size_t)2 * d * d, sizeof(double));
    return p;
}
void qso_problem_destroy(qso_problem *p) {
    if (!p) return;
    for (int i = 0; i < p->n_terms; i++) {
        free(p->terms[i].atom_a);
        free(p->terms[i].atom_b);
        free(p->terms[i].levels);
    }
    for (int i = 0; i < p->n_lind; i++) free(p->linds[i].op);
    for (int i = 0; i < p->n_tables; i++) free(p->tables[i]);
    free(p->tables);
    free(p->terms);
    free(p->linds);
    free(p->h0);
    free(p);
}
/* include/qsim_b200.h qsim_problem_add_table: returns the 1-based table id */
int qso_problem_add_table(qso_problem *p, double t_start, double t_end, int n_seg, int n_coef, const double *coef) {
    double *T = (double *)malloc(sizeof(double) * (4 + (size_t)n_seg * n_coef));
    T[0] = t_start;
    T[1] = (double)n_seg / (t_end - t_start);
    T[2] = (double)n_seg;
    T[3] = (double)n_coef;
    memcpy(T + 4, coef, sizeof(double) * (size_t)n_seg * n_coef);
    p->tables = (double **)realloc(p->tables, sizeof(double *) * (p->n_tables + 1));
    p->tables[p->n_tables++] = T;
    return p->n_tables;
}
static double *dup_mat(const double *m, int d) {
    if (!m) return NULL;
    double *r = (double *)malloc(sizeof(double) * 2 * d * d);
    memcpy(r, m, sizeof(double) * 2 * d * d);
    return r;
}
int qso_problem_set_h0(qso_problem *p, const double *h0) {
    memcpy(p->h0, h0, sizeof(double) * 2 * p->d * p->d);
    return 0;
}
int qso_problem_add_term(qso_problem *p, const qsim_term *t, const double *a, const double *b,
                         const int32_t *levels) {
    p->terms = (o_term *)realloc(p->terms, sizeof(o_term) * (p->n_terms + 1));
    o_term *ot = &p->terms[p->n_terms++];
    ot->desc = *t;
    ot->atom_a = dup_mat(a, p->d);
    ot->atom_b = dup_mat(b, p->d);
    ot->levels = NULL;
    ot->n_levels = 0;
    if (levels) {
        ot->levels = (int32_t *)malloc(sizeof(int32_t) * p->d);
        memcpy(ot->levels, levels, sizeof(int32_t) * p->d);
    }
    return 0;
}
/* parametric_drive on a DiagonalChargeBasisTransmon (systems.jl:143-152, 231-237).  t->kind = QSIM_TERM_SPECTRUM,
 * t->signal the flux, t->EC/EJ1/EJ2 the spec, t->num_terms the charge-basis size (101), t->rot the frame rate.
 * Unlike the product (which reads host-built tables of the level energies), the oracle does what the reference
 * does: an eigenvalue computation of the num_terms x num_terms charge-basis Hamiltonian on every RHS call.   */
int qso_problem_add_charge_term(qso_problem *p, const qsim_term *t, const int32_t *levels, int n_levels) {
    int rc = qso_problem_add_term(p, t, NULL, NULL, levels);
    p->terms[p->n_terms - 1].n_levels = n_levels;
    return rc;
}
int qso_problem_add_lindblad(qso_problem *p, const double *op, const qsim_signal *scale) {
    p->linds = (o_lind *)realloc(p->linds, sizeof(o_lind) * (p->n_lind + 1));
    o_lind *ol = &p->linds[p->n_lind++];
    ol->op = dup_mat(op, p->d);
    ol->has_scale = scale != NULL;
    if (scale) ol->scale = *scale;
    return 0;
}

/* ------------------------------------------------------- drive evaluation */
static inline double slot(const qsim_slot *s, const double *row) { return s->param >= 0 ? row[s->param] : s->value; }

/* piecewise Chebyshev series, Clenshaw recurrence; end values continue outside the table */
static double eval_table(const double *T, double t) {
    const int n_seg = (int)T[2], nc = (int)T[3];
    const double u = (t - T[0]) * T[1];
    int k = (int)floor(u);
    k = k < 0 ? 0 : (k >= n_seg ? n_seg - 1 : k);
    double x = 2.0 * (u - (double)k) - 1.0;
    x = x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);
    const double *c = T + 4 + (size_t)k * nc;
    double b1 = 0.0, b2 = 0.0;
    for (int j = nc - 1; j >= 1; j--) {
        double b0 = fma(2.0 * x, b1, c[j] - b2);
        b2 = b1;
        b1 = b0;
    }
    return fma(x, b1, c[0] - b2);
}

static double eval_signal(const qso_problem *p, const qsim_signal *s, double t, const double *row) {
    double car = 1.0, env = 1.0;
    if (s->carrier != QSIM_CARRIER_NONE) {
        double arg = (TWO_PI * slot(&s->freq, row)) * t + slot(&s->phase, row);
        car = s->carrier == QSIM_CARRIER_COS ? cos(arg) : sin(arg);
    }
    if (s->envelope == QSIM_ENV_GAUSSIAN) {
        double x = (t - slot(&s->t1, row)) / slot(&s->sigma, row);
        env = exp(-0.5 * (x * x));
    } else if (s->envelope == QSIM_ENV_ERF_SQUARED) {
        double sg = slot(&s->sigma, row);
        env = 0.5 * (erf((t - slot(&s->t1, row)) / sg) - erf((t - slot(&s->t2, row)) / sg));
    } else if (s->envelope == QSIM_ENV_TABLE) {
        env = (s->table >= 1 && s->table <= p->n_tables) ? eval_table(p->tables[s->table - 1], t) : 0.0;
    }
    if (s->blend == QSIM_BLEND_NONE) return slot(&s->offset, row) + slot(&s->amp, row) * env * car;
    return env * (slot(&s->offset, row) + slot(&s->amp, row) * car) + (1 - env) * slot(&s->rest, row);
}

/* perturbative_transmon.jl:193-201 (Horner with fused multiply-add) */
static double horner(double x, const double *c, int n) {
    if (n <= 0) return 0.0;
    double ex = c[n - 1];
    for (int i = n - 2; i >= 0; i--) ex = fma(x, ex, c[i]);
    return ex;
}
/* perturbative_transmon.jl:187-189, 222-231, 253-261 */
static void transmon_f_a(double EC, double EJ1, double EJ2, double phi, int num_terms, double *f, double *a) {
    int t = num_terms < 32 ? num_terms : 32;
    double xi = sqrt((2 * EC) / sqrt((EJ1 * EJ1 + EJ2 * EJ2) + (2 * EJ1 * EJ2) * cos(TWO_PI * phi)));
    *f = EC * (4 / xi + horner(xi, PT_FREQ, t - 1));
    *a = EC * horner(xi, PT_ANH, t - 1);
}

/* H(t) = H0 + sum of parametric terms, dense d x d (time_evolution.jl:32-34). */
/* Lowest `k` eigenvalues of the charge-basis transmon Hamiltonian (systems.jl:220-229): symmetric tridiagonal,
 * diagonal 4 EC n^2 for n = -N..N (size m = 2N + 1), off-diagonal -EJ/2.  Bisection on the Sturm count
 * (number of eigenvalues below x = number of negative pivots of T - x I), which is what sorting eigvals does for
 * the reference at :235, to the last bit the arithmetic allows.                                              */
static int sturm_count(int m, const double *diag, double off2, double x) {
    int cnt = 0;
    double q = diag[0] - x;
    if (q < 0) cnt++;
    for (int i = 1; i < m; i++) {
        if (q == 0.0) q = 1e-300;
        q = (diag[i] - x) - off2 / q;
        if (q < 0) cnt++;
    }
    return cnt;
}
static void charge_basis_levels(double EC, double EJ, int m, int k, double *out) {
    const int N = m / 2;
    double *diag = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < m; i++) diag[i] = 4.0 * EC * (double)((i - N) * (i - N));
    const double off = -0.5 * EJ, off2 = off * off;
    const double lo0 = -2.0 * fabs(off) - 1.0, hi0 = diag[0] + 2.0 * fabs(off) + 1.0; /* Gershgorin */
    for (int j = 0; j < k; j++) {
        double lo = lo0, hi = hi0;
        for (int it = 0; it < 200; it++) {
            const double mid = 0.5 * (lo + hi);
            if (mid <= lo || mid >= hi) break;
            if (sturm_count(m, diag, off2, mid) > j) hi = mid; else lo = mid;
        }
        out[j] = 0.5 * (lo + hi);
    }
    free(diag);
}

static void assemble_h(const qso_problem *p, double t, const double *row, double *ham) {
    const int d = p->d, dd = d * d;
    memcpy(ham, p->h0, sizeof(double) * 2 * dd);
    for (int k = 0; k < p->n_terms; k++) {
        const o_term *ot = &p->terms[k];
        const qsim_term *td = &ot->desc;
        double s = eval_signal(p, &td->signal, t, row);
        if (td->kind == QSIM_TERM_SCALAR) {
            for (int i = 0; i < dd; i++) {
                ham[2 * i] += s * ot->atom_a[2 * i];
                ham[2 * i + 1] += s * ot->atom_a[2 * i + 1];
            }
        } else if (td->kind == QSIM_TERM_ROTATING) {
            double th = TWO_PI * (slot(&td->rate, row) * t);
            double c = cos(th), sn = sin(th);
            for (int i = 0; i < dd; i++) {
                double ar = ot->atom_a[2 * i], ai = ot->atom_a[2 * i + 1];
                double br = ot->atom_b[2 * i], bi = ot->atom_b[2 * i + 1];
                /* e^{+i th} A + e^{-i th} B */
                double re = (c * ar - sn * ai) + (c * br + sn * bi);
                double im = (c * ai + sn * ar) + (c * bi - sn * br);
                ham[2 * i] += s * re;
                ham[2 * i + 1] += s * im;
            }
        } else if (td->kind == QSIM_TERM_SPECTRUM) {
            double lev[64];
            const double e1 = slot(&td->EJ1, row), e2 = slot(&td->EJ2, row);
            const double EJ = sqrt(e1 * e1 + e2 * e2 + 2.0 * e1 * e2 * cos(TWO_PI * s)); /* systems.jl:224 */
            charge_basis_levels(slot(&td->EC, row), EJ, td->num_terms, ot->n_levels, lev);
            double rot = slot(&td->rot, row);
            for (int i = 0; i < d; i++) {
                const int n = ot->levels[i];
                double e = lev[n];
                if (rot != 0) e -= (double)n * rot;                /* systems.jl:259-265 */
                ham[2 * (i + i * d)] += e;
            }
        } else {
            double f, a;
            if (td->kind == QSIM_TERM_DUFFING) {
                f = s;
                a = slot(&td->anharm, row);
            } else {
                transmon_f_a(slot(&td->EC, row), slot(&td->EJ1, row), slot(&td->EJ2, row), s, td->num_terms, &f, &a);
            }
            double rot = slot(&td->rot, row);
            for (int i = 0; i < d; i++) {
                double n = (double)ot->levels[i];
                double e = (f - 0.5 * a) * n + 0.5 * a * (n * n); /* systems.jl:192 */
                if (rot != 0) e -= n * rot;                        /* systems.jl:259-265 */
                ham[2 * (i + i * d)] += e;
            }
        }
    }
}

/* ------------------------------------------------------------ dense algebra */
/* C (+)= A(m x k) * B(k x n), column-major interleaved complex */
static void zgemm(int m, int n, int k, const double *A, const double *B, double *Cm, int accumulate) {
    if (!accumulate) memset(Cm, 0, sizeof(double) * 2 * m * n);
    for (int j = 0; j < n; j++) {
        double *cj = Cm + 2 * (size_t)j * m;
        for (int q = 0; q < k; q++) {
            const double br = B[2 * ((size_t)j * k + q)], bi = B[2 * ((size_t)j * k + q) + 1];
            if (br == 0.0 && bi == 0.0) continue;
            const double *aq = A + 2 * (size_t)q * m;
            for (int i = 0; i < m; i++) {
                double ar = aq[2 * i], ai = aq[2 * i + 1];
                cj[2 * i] += ar * br - ai * bi;
                cj[2 * i + 1] += ar * bi + ai * br;
            }
        }
    }
}
/* As zgemm, also skipping the zero entries of A (A is k columns of m; its nonzeros are listed once per call).
 * The surviving terms are added in the same order as in zgemm, so the result differs at most in the sign of
 * zeros.  Used by the structured me_propagator mode, where H and the collapse operators are sparse. */
static void zgemm_sa(int m, int n, int k, const double *A, const double *B, double *Cm, int accumulate, int *idx) {
    /* idx: workspace of k*(m+1) ints: per column q of A, count followed by the row indices of its nonzeros */
    for (int q = 0; q < k; q++) {
        int cnt = 0;
        for (int i = 0; i < m; i++)
            if (A[2 * ((size_t)q * m + i)] != 0.0 || A[2 * ((size_t)q * m + i) + 1] != 0.0) idx[q * (m + 1) + 1 + cnt++] = i;
        idx[q * (m + 1)] = cnt;
    }
    if (!accumulate) memset(Cm, 0, sizeof(double) * 2 * m * n);
    for (int j = 0; j < n; j++) {
        double *cj = Cm + 2 * (size_t)j * m;
        for (int q = 0; q < k; q++) {
            const double br = B[2 * ((size_t)j * k + q)], bi = B[2 * ((size_t)j * k + q) + 1];
            if (br == 0.0 && bi == 0.0) continue;
            const double *aq = A + 2 * (size_t)q * m;
            const int *iq = idx + q * (m + 1);
            for (int e = 1; e <= iq[0]; e++) {
                const int i = iq[e];
                double ar = aq[2 * i], ai = aq[2 * i + 1];
                cj[2 * i] += ar * br - ai * bi;
                cj[2 * i + 1] += ar * bi + ai * br;
            }
        }
    }
}
/* dense GEMM without the zero skip: used in faithful mode (BLAS does not skip) */
static void zgemm_dense(int m, int n, int k, const double *A, const double *B, double *Cm, int accumulate) {
    if (!accumulate) memset(Cm, 0, sizeof(double) * 2 * m * n);
    for (int j = 0; j < n; j++) {
        double *cj = Cm + 2 * (size_t)j * m;
        for (int q = 0; q < k; q++) {
            const double br = B[2 * ((size_t)j * k + q)], bi = B[2 * ((size_t)j * k + q) + 1];
            const double *aq = A + 2 * (size_t)q * m;
            for (int i = 0; i < m; i++) {
                double ar = aq[2 * i], ai = aq[2 * i + 1];
                cj[2 * i] += ar * br - ai * bi;
                cj[2 * i + 1] += ar * bi + ai * br;
            }
        }
    }
}
static void adjoint(int d, const double *A, double *Ah) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            Ah[2 * (j + i * d)] = A[2 * (i + j * d)];
            Ah[2 * (j + i * d) + 1] = -A[2 * (i + j * d) + 1];
        }
}
/* K (+)= alpha * kron(A, B) with optional transposes/conjugates applied by the caller */
static void kron_acc(int d, double are, double aim, const double *A, const double *B, double *K) {
    const int D = d * d;
    for (int ja = 0; ja < d; ja++)
        for (int ia = 0; ia < d; ia++) {
            double xr = A[2 * (ia + ja * d)], xi = A[2 * (ia + ja * d) + 1];
            if (xr == 0.0 && xi == 0.0) continue;
            double sr = are * xr - aim * xi, si = are * xi + aim * xr;
            for (int jb = 0; jb < d; jb++)
                for (int ib = 0; ib < d; ib++) {
                    double yr = B[2 * (ib + jb * d)], yi = B[2 * (ib + jb * d) + 1];
                    size_t idx = (size_t)(ia * d + ib) + (size_t)(ja * d + jb) * D;
                    K[2 * idx] += sr * yr - si * yi;
                    K[2 * idx + 1] += sr * yi + si * yr;
                }
        }
}

/* ------------------------------------------------------------------- RHS */
typedef struct {
    const qso_problem *p;
    const double *row;
    int which; /* 0 U-prop, 1 U-state, 2 me-prop, 3 me-state */
    int mode;
    int d;
    size_t n; /* complex unknowns */
    double *ham, *lm, *lh, *ldl, *tmp1, *tmp2, *eye, *sup, *xt;
    /* per-RHS-call collapse operators L_i(t), L_i', L_i'L_i (n_lind matrices each; me_prepare) */
    double *lms, *lhs, *ldls;
    int *spidx;
} rhs_ctx;

static void lind_matrix(const rhs_ctx *c, int i, double t, double *lm) {
    const o_lind *ol = &c->p->linds[i];
    const int dd = c->d * c->d;
    double s = ol->has_scale ? eval_signal(c->p, &ol->scale, t, c->row) : 1.0;
    for (int q = 0; q < dd; q++) {
        lm[2 * q] = s * ol->op[2 * q];
        lm[2 * q + 1] = s * ol->op[2 * q + 1];
    }
}

/* collapse operators at time t, formed once per RHS call (they are the same for every column) */
static void me_prepare(rhs_ctx *c, double t) {
    const int d = c->d, dd = d * d, nl = c->p->n_lind;
    if (!c->lms && nl > 0) {
        c->lms = (double *)malloc(sizeof(double) * 2 * dd * nl);
        c->lhs = (double *)malloc(sizeof(double) * 2 * dd * nl);
        c->ldls = (double *)malloc(sizeof(double) * 2 * dd * nl);
    }
    if (!c->spidx) c->spidx = (int *)malloc(sizeof(int) * (size_t)d * (d + 1));
    for (int i = 0; i < nl; i++) {
        double *lm = c->lms + 2 * (size_t)dd * i, *lh = c->lhs + 2 * (size_t)dd * i, *ldl = c->ldls + 2 * (size_t)dd * i;
        lind_matrix(c, i, t, lm);
        adjoint(d, lm, lh);
        zgemm(d, d, d, lh, lm, ldl, 0);
    }
}

/* d x d master-equation RHS on one matrix X (time_evolution.jl:146,152); me_prepare(c, t) first */
static void me_apply(rhs_ctx *c, const double *X, double *dX) {
    const int d = c->d, dd = d * d;
    /* -2 pi i (H X - X H) */
    zgemm_sa(d, d, d, c->ham, X, c->tmp1, 0, c->spidx);
    zgemm(d, d, d, X, c->ham, c->tmp2, 0);
    for (int q = 0; q < dd; q++) {
        double re = c->tmp1[2 * q] - c->tmp2[2 * q], im = c->tmp1[2 * q + 1] - c->tmp2[2 * q + 1];
        dX[2 * q] = TWO_PI * im;
        dX[2 * q + 1] = -TWO_PI * re;
    }
    for (int i = 0; i < c->p->n_lind; i++) {
        const double *lm = c->lms + 2 * (size_t)dd * i, *lh = c->lhs + 2 * (size_t)dd * i,
                     *ldl = c->ldls + 2 * (size_t)dd * i;
        /* L X L' */
        zgemm_sa(d, d, d, lm, X, c->tmp1, 0, c->spidx);
        zgemm(d, d, d, c->tmp1, lh, c->tmp2, 0);
        for (int q = 0; q < 2 * dd; q++) dX[q] += TWO_PI * c->tmp2[q];
        /* -1/2 L'L X - 1/2 X L'L */
        zgemm_sa(d, d, d, ldl, X, c->tmp1, 0, c->spidx);
        zgemm(d, d, d, X, ldl, c->tmp2, 0);
        for (int q = 0; q < 2 * dd; q++) dX[q] -= TWO_PI * 0.5 * (c->tmp1[q] + c->tmp2[q]);
    }
}

/* dense superoperator pieces of time_evolution.jl:104,109 */
static void build_super_h(rhs_ctx *c, double *K) {
    const int d = c->d;
    const size_t D = (size_t)d * d;
    memset(K, 0, sizeof(double) * 2 * D * D);
    /* -2 pi i (I (x) H - H^T (x) I) */
    double *ht = c->tmp1;
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            ht[2 * (j + i * d)] = c->ham[2 * (i + j * d)];
            ht[2 * (j + i * d) + 1] = c->ham[2 * (i + j * d) + 1];
        }
    kron_acc(d, 0.0, -TWO_PI, c->eye, c->ham, K);
    kron_acc(d, 0.0, TWO_PI, ht, c->eye, K);
}
static void build_super_l(rhs_ctx *c, double *K) {
    const int d = c->d;
    const size_t D = (size_t)d * d;
    memset(K, 0, sizeof(double) * 2 * D * D);
    adjoint(d, c->lm, c->lh);
    zgemm(d, d, d, c->lh, c->lm, c->ldl, 0);
    double *lc = c->tmp1, *ldlt = c->tmp2;
    for (int q = 0; q < d * d; q++) {
        lc[2 * q] = c->lm[2 * q];
        lc[2 * q + 1] = -c->lm[2 * q + 1];
    }
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            ldlt[2 * (j + i * d)] = c->ldl[2 * (i + j * d)];
            ldlt[2 * (j + i * d) + 1] = c->ldl[2 * (i + j * d) + 1];
        }
    kron_acc(d, TWO_PI, 0.0, lc, c->lm, K);
    kron_acc(d, -TWO_PI * 0.5, 0.0, c->eye, c->ldl, K);
    kron_acc(d, -TWO_PI * 0.5, 0.0, ldlt, c->eye, K);
}

static void rhs_eval(rhs_ctx *c, const double *u, double t, double *du) {
    const int d = c->d, dd = d * d;
    assemble_h(c->p, t, c->row, c->ham);
    if (c->which <= 1) {
        /* rmul!(ham, -2 pi i); mul!(du, ham, u)  (time_evolution.jl:35-36) */
        for (int q = 0; q < dd; q++) {
            double re = c->ham[2 * q], im = c->ham[2 * q + 1];
            c->ham[2 * q] = TWO_PI * im;
            c->ham[2 * q + 1] = -TWO_PI * re;
        }
        if (c->mode == 0)
            zgemm_dense(d, c->which == 0 ? d : 1, d, c->ham, u, du, 0);
        else
            zgemm(d, c->which == 0 ? d : 1, d, c->ham, u, du, 0);
    } else if (c->which == 3) {
        me_prepare(c, t);
        me_apply(c, u, du);
    } else if (c->mode == 1) {
        me_prepare(c, t);
        for (int j = 0; j < dd; j++) me_apply(c, u + 2 * (size_t)j * dd, du + 2 * (size_t)j * dd);
    } else {
        build_super_h(c, c->sup);
        zgemm_dense(dd, dd, dd, c->sup, u, du, 0);
        for (int i = 0; i < c->p->n_lind; i++) {
            lind_matrix(c, i, t, c->lm);
            build_super_l(c, c->sup);
            zgemm_dense(dd, dd, dd, c->sup, u, du, 1);
        }
    }
}

/* evaluation entry for tests: kind 0 H(t), kind 1 full superoperator G(t) */
int qso_problem_eval(const qso_problem *p, int kind, double t, const double *row, double *out) {
    rhs_ctx c;
    memset(&c, 0, sizeof(c));
    const int d = p->d, dd = d * d;
    c.p = p;
    c.row = row;
    c.d = d;
    c.ham = (double *)malloc(sizeof(double) * 2 * dd);
    assemble_h(p, t, row, c.ham);
    if (kind == 0) {
        memcpy(out, c.ham, sizeof(double) * 2 * dd);
        free(c.ham);
        return 0;
    }
    size_t D = (size_t)dd;
    c.lm = (double *)malloc(sizeof(double) * 2 * dd);
    c.lh = (double *)malloc(sizeof(double) * 2 * dd);
    c.ldl = (double *)malloc(sizeof(double) * 2 * dd);
    c.tmp1 = (double *)malloc(sizeof(double) * 2 * dd);
    c.tmp2 = (double *)malloc(sizeof(double) * 2 * dd);
    c.eye = (double *)calloc(2 * dd, sizeof(double));
    for (int i = 0; i < d; i++) c.eye[2 * (i + i * d)] = 1.0;
    c.sup = (double *)malloc(sizeof(double) * 2 * D * D);
    build_super_h(&c, out);
    for (int i = 0; i < p->n_lind; i++) {
        lind_matrix(&c, i, t, c.lm);
        build_super_l(&c, c.sup);
        for (size_t q = 0; q < 2 * D * D; q++) out[q] += c.sup[q];
    }
    free(c.ham); free(c.lm); free(c.lh); free(c.ldl); free(c.tmp1); free(c.tmp2); free(c.eye); free(c.sup);
    return 0;
}

/* ------------------------------------------------------------ Tsit5 solver */
static const double TC[7] = {0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0};
static const double TA[7][6] = {
    {0},
    {0.161},
    {-0.008480655492356989, 0.335480655492357},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}};
static const double TBT[7] = {-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                              -0.1447110071732629,     0.5823571654525552,     -0.45808210592918697,
                              0.015151515151515152};
static const double TR[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                {0.13169999999999998, -0.2234, 0.1017, 0},
                                {3.9302962368947516, -5.941033872131505, 2.490627285651253, 0},
                                {-12.411077166933676, 30.33818863028232, -16.548102889244902, 0},
                                {37.50931341651104, -88.1789048947664, 47.37952196281928, 0},
                                {-27.896526289197286, 65.09189467479366, -34.87065786149661, 0},
                                {1.5, -4.0, 2.5, 0}};

int qso_tsit5_constants(double *c, double *a, double *btilde, double *r) {
    memcpy(c, TC, sizeof(TC));
    memcpy(a, TA, sizeof(TA));
    memcpy(btilde, TBT, sizeof(TBT));
    memcpy(r, TR, sizeof(TR));
    return 0;
}

static double rms_scaled(size_t n, const double *x, const double *sk) {
    double s = 0.0;
    for (size_t i = 0; i < n; i++) {
        double re = x[2 * i] / sk[i], im = x[2 * i + 1] / sk[i];
        s += re * re + im * im;
    }
    return sqrt(s / (double)n);
}

static void solve_one(rhs_ctx *c, const double *u0, const double *ts, int n_ts, const qsim_solver_opts *o,
                      double *out, qsim_traj_stats *st) {
    const size_t n = c->n;
    qsim_traj_stats s;
    memset(&s, 0, sizeof(s));
    double *u = (double *)malloc(sizeof(double) * 2 * n);
    double *unew = (double *)malloc(sizeof(double) * 2 * n);
    double *tmp = (double *)malloc(sizeof(double) * 2 * n);
    double *sk = (double *)malloc(sizeof(double) * n);
    double *k[7];
    for (int i = 0; i < 7; i++) k[i] = (double *)malloc(sizeof(double) * 2 * n);
    memcpy(u, u0, sizeof(double) * 2 * n);
    const double t0 = ts[0], tend = ts[n_ts - 1];
    int nsave = 0;
    while (nsave < n_ts && ts[nsave] <= t0) { /* save_start */
        memcpy(out + 2 * n * (size_t)nsave, u, sizeof(double) * 2 * n);
        nsave++;
    }
    double t = t0;
    if (tend > t0) {
        const double dtmax = o->dtmax > 0 ? o->dtmax : tend - t0;
        /* initial step: DiffEqBase ode_determine_initdt */
        for (size_t i = 0; i < n; i++) sk[i] = o->abstol + hypot(u[2 * i], u[2 * i + 1]) * o->reltol;
        rhs_eval(c, u, t0, k[0]);
        s.n_rhs++;
        double d0 = rms_scaled(n, u, sk), d1 = rms_scaled(n, k[0], sk);
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : (d0 / d1) / 100;
        dt0 = fmin(dt0, dtmax);
        for (size_t i = 0; i < 2 * n; i++) tmp[i] = u[i] + dt0 * k[0][i];
        rhs_eval(c, tmp, t0 + dt0, k[1]);
        s.n_rhs++;
        for (size_t i = 0; i < 2 * n; i++) tmp[i] = k[1][i] - k[0][i];
        double d2 = rms_scaled(n, tmp, sk) / dt0;
        double md = fmax(d1, d2);
        double dt1 = md <= 1e-15 ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2 + log10(md)) / 5);
        double dt = fmin(fmin(100 * dt0, dt1), dtmax);
        s.dt_init = dt;
        rhs_eval(c, u, t0, k[0]); /* the integrator's own first stage */
        s.n_rhs++;
        double qold = o->qoldinit;
        int64_t iters = 0;
        while (t < tend) {
            if (++iters > o->maxiters) { s.retcode = QSIM_RET_MAXITERS; break; }
            dt = fmin(dt, dtmax);
            dt = fmin(dt, tend - t);
            if (!(dt > fabs(t) * 2.220446049250313e-16)) { s.retcode = QSIM_RET_DT_UNDERFLOW; break; }
            for (int i = 1; i < 6; i++) {
                for (size_t q = 0; q < 2 * n; q++) {
                    double acc = TA[i][0] * k[0][q];
                    for (int j = 1; j < i; j++) acc += TA[i][j] * k[j][q];
                    tmp[q] = u[q] + dt * acc;
                }
                rhs_eval(c, tmp, t + TC[i] * dt, k[i]);
            }
            for (size_t q = 0; q < 2 * n; q++) {
                double acc = TA[6][0] * k[0][q];
                for (int j = 1; j < 6; j++) acc += TA[6][j] * k[j][q];
                unew[q] = u[q] + dt * acc;
            }
            rhs_eval(c, unew, t + dt, k[6]);
            s.n_rhs += 6;
            double ss = 0.0;
            for (size_t i = 0; i < n; i++) {
                double er = 0, ei = 0;
                er = TBT[0] * k[0][2 * i];
                ei = TBT[0] * k[0][2 * i + 1];
                for (int j = 1; j < 7; j++) {
                    er += TBT[j] * k[j][2 * i];
                    ei += TBT[j] * k[j][2 * i + 1];
                }
                er *= dt;
                ei *= dt;
                double sc = o->abstol +
                            fmax(hypot(u[2 * i], u[2 * i + 1]), hypot(unew[2 * i], unew[2 * i + 1])) * o->reltol;
                er /= sc;
                ei /= sc;
                ss += er * er + ei * ei;
            }
            double eest = sqrt(ss / (double)n);
            if (!isfinite(eest)) { s.retcode = QSIM_RET_NONFINITE; break; }
            double q, q11 = 0.0;
            if (eest == 0.0) {
                q = 1.0 / o->qmax;
            } else {
                q11 = pow(eest, o->beta1);
                q = q11 / pow(qold, o->beta2);
                q = fmax(1.0 / o->qmax, fmin(1.0 / o->qmin, q / o->gamma));
            }
            if (eest <= 1.0) {
                s.n_accept++;
                qold = fmax(eest, o->qoldinit);
                double dtnew = dt / q;
                double tnew = t + dt;
                double mx = fmax(t, tend);
                if (fabs(tnew - tend) < 10 * (nextafter(mx, INFINITY) - mx)) tnew = tend;
                while (nsave < n_ts && ts[nsave] <= tnew) {
                    double *dst = out + 2 * n * (size_t)nsave;
                    if (ts[nsave] == tnew) {
                        memcpy(dst, unew, sizeof(double) * 2 * n);
                    } else {
                        double th = (ts[nsave] - t) / dt, b[7];
                        b[0] = th * (TR[0][0] + th * (TR[0][1] + th * (TR[0][2] + th * TR[0][3])));
                        for (int j = 1; j < 7; j++) b[j] = th * th * (TR[j][0] + th * (TR[j][1] + th * TR[j][2]));
                        for (size_t qq = 0; qq < 2 * n; qq++) {
                            double acc = k[0][qq] * b[0];
                            for (int j = 1; j < 7; j++) acc += k[j][qq] * b[j];
                            dst[qq] = u[qq] + dt * acc;
                        }
                    }
                    nsave++;
                }
                t = tnew;
                double *sw = u; u = unew; unew = sw;
                sw = k[0]; k[0] = k[6]; k[6] = sw;
                dt = dtnew;
            } else {
                s.n_reject++;
                dt = dt / fmin(1.0 / o->qmin, q11 / o->gamma);
            }
        }
    }
    /* unreached save points (failed trajectories) get NaN */
    for (; nsave < n_ts; nsave++)
        for (size_t q = 0; q < 2 * n; q++) out[2 * n * (size_t)nsave + q] = NAN;
    s.t_final = t;
    if (st) *st = s;
    free(u); free(unew); free(tmp); free(sk);
    for (int i = 0; i < 7; i++) free(k[i]);
}

/* --------------------------------------------------------- threaded sweep */
typedef struct {
    const qso_problem *p;
    int which, mode;
    int64_t n_traj;
    const double *params, *ts, *init;
    int n_ts;
    int64_t ts_stride, init_stride;
    const qsim_solver_opts *opts;
    double *out;
    qsim_traj_stats *stats;
    int64_t next;
    pthread_mutex_t mu;
} sweep_job;

static void *sweep_worker(void *arg) {
    sweep_job *job = (sweep_job *)arg;
    const qso_problem *p = job->p;
    const int d = p->d, dd = d * d;
    rhs_ctx c;
    memset(&c, 0, sizeof(c));
    c.p = p;
    c.which = job->which;
    c.mode = job->mode;
    c.d = d;
    c.n = job->which == 0 ? (size_t)dd : job->which == 1 ? (size_t)d : job->which == 2 ? (size_t)dd * dd : (size_t)dd;
    c.ham = (double *)malloc(sizeof(double) * 2 * dd);
    c.lm = (double *)malloc(sizeof(double) * 2 * dd);
    c.lh = (double *)malloc(sizeof(double) * 2 * dd);
    c.ldl = (double *)malloc(sizeof(double) * 2 * dd);
    c.tmp1 = (double *)malloc(sizeof(double) * 2 * dd);
    c.tmp2 = (double *)malloc(sizeof(double) * 2 * dd);
    c.eye = (double *)calloc(2 * dd, sizeof(double));
    for (int i = 0; i < d; i++) c.eye[2 * (i + i * d)] = 1.0;
    if (job->which == 2 && job->mode == 0) c.sup = (double *)malloc(sizeof(double) * 2 * (size_t)dd * dd);
    double *u0 = (double *)calloc(2 * c.n, sizeof(double));
    for (;;) {
        pthread_mutex_lock(&job->mu);
        int64_t i = job->next++;
        pthread_mutex_unlock(&job->mu);
        if (i >= job->n_traj) break;
        c.row = job->params ? job->params + (size_t)i * p->n_params : NULL;
        if (job->which == 0 || job->which == 2) {
            size_t m = job->which == 0 ? (size_t)d : (size_t)dd;
            memset(u0, 0, sizeof(double) * 2 * c.n);
            for (size_t q = 0; q < m; q++) u0[2 * (q + q * m)] = 1.0;
        } else {
            const double *src = job->init + 2 * (size_t)(job->init_stride ? i * job->init_stride : 0);
            memcpy(u0, src, sizeof(double) * 2 * c.n);
        }
        const double *ts = job->ts + (job->ts_stride ? i * job->ts_stride : 0);
        solve_one(&c, u0, ts, job->n_ts, job->opts, job->out + 2 * c.n * (size_t)job->n_ts * (size_t)i,
                  job->stats ? job->stats + i : NULL);
    }
    free(u0); free(c.ham); free(c.lm); free(c.lh); free(c.ldl); free(c.tmp1); free(c.tmp2); free(c.eye); free(c.sup);
    free(c.lms); free(c.lhs); free(c.ldls); free(c.spidx);
    return NULL;
}

/* init_stride is in complex elements per trajectory (0: shared initial state). */
int qso_solve(const qso_problem *p, int which, int mode, int64_t n_traj, const double *params, const double *ts,
              int n_ts, int64_t ts_stride, const double *init, int64_t init_stride, const qsim_solver_opts *opts,
              double *out, qsim_traj_stats *stats, int n_threads) {
    if (!p || which < 0 || which > 3 || n_ts < 1) return -1;
    if ((which == 1 || which == 3) && !init) return -1;
    for (int64_t i = 0; i < (ts_stride ? n_traj : 1); i++)
        for (int q = 1; q < n_ts; q++)
            if (ts[i * ts_stride + q] < ts[i * ts_stride + q - 1]) return -1; /* @assert issorted(ts) */
    qsim_solver_opts o;
    if (opts)
        o = *opts;
    else {
        memset(&o, 0, sizeof(o));
        o.reltol = 1e-6; o.abstol = 1e-8; o.gamma = 0.9; o.qmin = 0.2; o.qmax = 10.0;
        o.beta1 = 7.0 / 50; o.beta2 = 2.0 / 25; o.qoldinit = 1e-4; o.maxiters = 1000000;
    }
    sweep_job job;
    memset(&job, 0, sizeof(job));
    job.p = p; job.which = which; job.mode = mode; job.n_traj = n_traj; job.params = params; job.ts = ts;
    job.init = init; job.n_ts = n_ts; job.ts_stride = ts_stride; job.init_stride = init_stride; job.opts = &o;
    job.out = out; job.stats = stats;
    pthread_mutex_init(&job.mu, NULL);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_traj) n_threads = (int)n_traj;
    if (n_threads <= 1) {
        sweep_worker(&job);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
        for (int i = 0; i < n_threads; i++) pthread_create(&th[i], NULL, sweep_worker, &job);
        for (int i = 0; i < n_threads; i++) pthread_join(th[i], NULL);
        free(th);
    }
    pthread_mutex_destroy(&job.mu);
    return 0;
}
