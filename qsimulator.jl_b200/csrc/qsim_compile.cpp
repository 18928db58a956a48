// Host-side generator compiler: term lists of a CompositeQSystem -> sparse A(t) with value slots.
//
// Restates, symbolically, what the reference recomputes numerically on every RHS call:
//   H(t) = H0 + sum_k embed(f_k(t))                     composite_systems.jl:46-50, 65-72
//   unitary:   A = -2 pi i H                            time_evolution.jl:33-36
//   Lindblad:  A = -2 pi i (I(x)H - H^T(x)I)
//                + 2 pi sum (conj(L)(x)L - 1/2 I(x)L'L - 1/2 (L'L)^T(x)I)      time_evolution.jl:104-109
// with column-major vec, kron(P,Q)[(r,c),(k,l)] = P[c,l] Q[r,k].  Every matrix entry becomes a
// linear expression in the real channels c_k(t); identical expressions share one value slot.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "qsim_eval.h"

namespace qsim {

// perturbative_transmon.jl:38-105: coefficient = numerator / 2^b
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

const SeriesTables &series_tables() {
    static SeriesTables tab;
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 31; i++) {
            tab.freq[i] = std::ldexp(std::strtod(FREQ_NUM[i], nullptr), -FREQ_EXP[i]);
            tab.anh[i] = std::ldexp(std::strtod(ANH_NUM[i], nullptr), -ANH_EXP[i]);
        }
        init = true;
    }
    return tab;
}

namespace {

typedef std::map<int, cplx> Expr;  // channel -> weight

inline void expr_add(Expr &e, int chan, cplx w) {
    if (w == cplx(0.0, 0.0)) return;
    e[chan] += w;
}
inline void expr_axpy(Expr &dst, cplx a, const Expr &src) {
    for (auto &kv : src) expr_add(dst, kv.first, a * kv.second);
}

struct ChannelAlloc {
    std::vector<EvalDesc> evals;
    std::vector<bool> chan_dynamic;  // per channel
    ChannelAlloc() { chan_dynamic.push_back(false); }  // channel 0: constant 1
    int add(EvalDesc e, int n_out) {
        e.out_chan = (int)chan_dynamic.size();
        for (int i = 0; i < n_out; i++) chan_dynamic.push_back(e.dynamic != 0);
        evals.push_back(e);
        return e.out_chan;
    }
};

EvalDesc blank_eval() {
    EvalDesc e;
    std::memset(&e, 0, sizeof(e));
    return e;
}

// Channel carrying a slot: literal slots fold into channel 0 with the value as weight.
struct SlotRef {
    int chan;
    double scale;
};
SlotRef slot_channel(ChannelAlloc &ca, const qsim_slot &s) {
    if (s.param < 0) return {0, s.value};
    EvalDesc e = blank_eval();
    e.kind = EV_SLOT;
    e.dynamic = 0;
    e.slot = s;
    return {ca.add(e, 1), 1.0};
}

}  // namespace

static void build_h_exprs(const qsim_problem &p, ChannelAlloc &ca, std::vector<Expr> &H) {
    const int d = p.d;
    H.assign((size_t)d * d, Expr());
    for (int q = 0; q < d * d; q++) expr_add(H[q], 0, p.h0[q]);
    for (const auto &t : p.terms) {
        const qsim_term &td = t.desc;
        const bool sig_dyn = signal_is_dynamic(td.signal);
        if (td.kind == QSIM_TERM_SCALAR) {
            EvalDesc e = blank_eval();
            e.kind = EV_SIGNAL;
            e.sig = td.signal;
            e.dynamic = sig_dyn;
            int ch = ca.add(e, 1);
            for (int q = 0; q < d * d; q++) expr_add(H[q], ch, t.a[q]);
        } else if (td.kind == QSIM_TERM_ROTATING) {
            EvalDesc e = blank_eval();
            e.kind = EV_ROTATING;
            e.sig = td.signal;
            e.rate = td.rate;
            e.dynamic = sig_dyn || td.rate.param >= 0 || td.rate.value != 0.0;
            int ch = ca.add(e, 2);
            const cplx I(0.0, 1.0);
            for (int q = 0; q < d * d; q++) {
                expr_add(H[q], ch, t.a[q] + t.b[q]);            // cos th (A + B)
                expr_add(H[q], ch + 1, I * (t.a[q] - t.b[q]));  // sin th i(A - B)
            }
        } else if (td.kind == QSIM_TERM_SPECTRUM) {
            // one channel per level: E_k(cos(pi phi(t))) from the level's table; - rot * n like the other diagonal terms
            std::vector<int> ch(t.tables.size(), -1);
            SlotRef rr = slot_channel(ca, td.rot);
            for (int i = 0; i < d; i++) {
                const int n = t.levels[i];
                if (n < 0 || n >= (int)t.tables.size()) throw std::runtime_error("spectrum term: level without a table");
                if (ch[(size_t)n] < 0) {
                    EvalDesc e = blank_eval();
                    e.kind = EV_SPECTRUM;
                    e.sig = td.signal;
                    e.num_terms = t.tables[(size_t)n];   // table id, translated to an offset with the other tables
                    e.dynamic = sig_dyn;
                    ch[(size_t)n] = ca.add(e, 1);
                }
                Expr &ex = H[(size_t)i + (size_t)i * d];
                expr_add(ex, ch[(size_t)n], 1.0);
                expr_add(ex, rr.chan, -rr.scale * (double)n);
            }
        } else {
            int ch_f, ch_a;
            double a_scale = 1.0;
            if (td.kind == QSIM_TERM_DUFFING) {
                EvalDesc e = blank_eval();
                e.kind = EV_SIGNAL;
                e.sig = td.signal;
                e.dynamic = sig_dyn;
                ch_f = ca.add(e, 1);
                SlotRef ar = slot_channel(ca, td.anharm);
                ch_a = ar.chan;
                a_scale = ar.scale;
            } else {
                EvalDesc e = blank_eval();
                e.kind = EV_TRANSMON;
                e.sig = td.signal;
                e.EC = td.EC;
                e.EJ1 = td.EJ1;
                e.EJ2 = td.EJ2;
                e.num_terms = td.num_terms;
                e.dynamic = sig_dyn;
                ch_f = ca.add(e, 2);
                ch_a = ch_f + 1;
            }
            SlotRef rr = slot_channel(ca, td.rot);
            for (int i = 0; i < d; i++) {
                const double n = (double)t.levels[i];
                // E_n = (f - a/2) n + (a/2) n^2 - rot n = f n + a n(n-1)/2 - rot n   (systems.jl:192, 259-265)
                Expr &e = H[(size_t)i + (size_t)i * d];
                expr_add(e, ch_f, n);
                expr_add(e, ch_a, a_scale * (0.5 * n * (n - 1.0)));
                expr_add(e, rr.chan, -rr.scale * n);
            }
        }
    }
}

void build_generator(const qsim_problem &p, int which, Generator &g) {
    const int d = p.d;
    ChannelAlloc ca;
    std::vector<Expr> H;
    build_h_exprs(p, ca, H);
    const int n = which == 0 ? d : d * d;
    if (n > 65535) throw std::runtime_error("state dimension too large for 16-bit row indices");
    // row -> (src -> expr)
    std::vector<std::map<int, Expr>> rows((size_t)n);
    const cplx mI2pi(0.0, -QSIM_TWO_PI), pI2pi(0.0, QSIM_TWO_PI);
    if (which == 0) {
        for (int c = 0; c < d; c++)
            for (int r = 0; r < d; r++) {
                const Expr &e = H[(size_t)r + (size_t)c * d];
                if (!e.empty()) expr_axpy(rows[r][c], mI2pi, e);
            }
    } else {
        for (int k = 0; k < d; k++)
            for (int r = 0; r < d; r++) {
                const Expr &e = H[(size_t)r + (size_t)k * d];  // H[r,k]
                if (e.empty()) continue;
                // (H X)[r,c]: row (r,c) <- src (k,c)
                for (int c = 0; c < d; c++) expr_axpy(rows[r + c * d][k + c * d], mI2pi, e);
                // (X H)[rr,k]: H[r,k] multiplies X[rr,r]: row (rr,k) <- src (rr,r)
                for (int rr = 0; rr < d; rr++) expr_axpy(rows[rr + k * d][rr + r * d], pI2pi, e);
            }
        for (const auto &l : p.linds) {
            int ch = 0;
            if (l.has_scale) {
                EvalDesc e = blank_eval();
                e.kind = EV_SQUARE;
                e.sig = l.scale;
                e.dynamic = signal_is_dynamic(l.scale);
                ch = ca.add(e, 1);
            }
            const std::vector<cplx> &L = l.op;
            // sparse list of L
            std::vector<std::pair<int, int>> nz;
            for (int c = 0; c < d; c++)
                for (int r = 0; r < d; r++)
                    if (L[(size_t)r + (size_t)c * d] != cplx(0, 0)) nz.push_back({r, c});
            // M = L' L
            std::vector<cplx> M((size_t)d * d, cplx(0, 0));
            for (auto &a : nz)
                for (auto &b : nz)
                    if (a.first == b.first)  // M[ac, bc] += conj(L[r, ac]) L[r, bc]
                        M[(size_t)a.second + (size_t)b.second * d] +=
                            std::conj(L[(size_t)a.first + (size_t)a.second * d]) * L[(size_t)b.first + (size_t)b.second * d];
            // conj(L) (x) L : row (r,c) <- src (k,l) weight conj(L[c,l]) L[r,k]
            for (auto &a : nz)      // L[r,k]
                for (auto &b : nz)  // L[c,l]
                    expr_add(rows[a.first + b.first * d][a.second + b.second * d], ch,
                             QSIM_TWO_PI * (std::conj(L[(size_t)b.first + (size_t)b.second * d]) *
                                            L[(size_t)a.first + (size_t)a.second * d]));
            for (int k = 0; k < d; k++)
                for (int r = 0; r < d; r++) {
                    cplx m = M[(size_t)r + (size_t)k * d];  // M[r,k]
                    if (m == cplx(0, 0)) continue;
                    // -1/2 (M X)[r,c]: row (r,c) <- src (k,c)
                    for (int c = 0; c < d; c++) expr_add(rows[r + c * d][k + c * d], ch, -0.5 * QSIM_TWO_PI * m);
                    // -1/2 (X M)[rr,k]: row (rr,k) <- src (rr,r) weight M[r,k]
                    for (int rr = 0; rr < d; rr++) expr_add(rows[rr + k * d][rr + r * d], ch, -0.5 * QSIM_TWO_PI * m);
                }
        }
    }
    // dedupe expressions into slots
    typedef std::vector<std::pair<int, std::pair<double, double>>> Key;
    std::map<Key, int> slot_of;
    std::vector<Expr> slot_expr;
    std::vector<bool> slot_dyn;
    std::vector<std::vector<std::pair<int, int>>> row_entries((size_t)n);  // (src, provisional slot)
    for (int r = 0; r < n; r++) {
        for (auto &se : rows[r]) {
            Expr e;
            for (auto &kv : se.second)
                if (kv.second != cplx(0, 0)) e[kv.first] = kv.second;
            if (e.empty()) continue;
            Key key;
            bool dyn = false;
            for (auto &kv : e) {
                key.push_back({kv.first, {kv.second.real(), kv.second.imag()}});
                dyn = dyn || ca.chan_dynamic[kv.first];
            }
            auto it = slot_of.find(key);
            int s;
            if (it == slot_of.end()) {
                s = (int)slot_expr.size();
                slot_of[key] = s;
                slot_expr.push_back(e);
                slot_dyn.push_back(dyn);
            } else {
                s = it->second;
            }
            row_entries[r].push_back({se.first, s});
        }
    }
    // dynamic slots first
    const int ns = (int)slot_expr.size();
    if (ns > 65534) throw std::runtime_error("too many distinct generator values for 16-bit slots");
    std::vector<int> order, new_id((size_t)ns);
    for (int s = 0; s < ns; s++)
        if (slot_dyn[s]) order.push_back(s);
    const int n_dyn = (int)order.size();
    for (int s = 0; s < ns; s++)
        if (!slot_dyn[s]) order.push_back(s);
    for (int i = 0; i < ns; i++) new_id[order[i]] = i;

    g = Generator();
    g.n = n;
    g.n_slots = ns;
    g.n_dyn_slots = n_dyn;
    g.n_chan = (int)ca.chan_dynamic.size();
    g.evals = ca.evals;
    // tabulated envelopes: table id (1-based) -> 1 + offset into the flat table array
    g.tables = p.tables;
    for (auto &e : g.evals) {
        if (e.kind == EV_SPECTRUM) {
            if (e.num_terms < 1 || e.num_terms > (int)p.table_offset.size())
                throw std::runtime_error("spectrum term refers to a table that was not added");
            e.num_terms = 1 + p.table_offset[(size_t)e.num_terms - 1];
        }
        if (e.sig.envelope == QSIM_ENV_TABLE) {
            if (e.sig.table < 1 || e.sig.table > (int)p.table_offset.size())
                throw std::runtime_error("signal refers to a table that was not added");
            e.sig.table = 1 + p.table_offset[(size_t)e.sig.table - 1];
        } else {
            e.sig.table = 0;
        }
    }
    g.slot_ptr.push_back(0);
    for (int i = 0; i < ns; i++) {
        for (auto &kv : slot_expr[order[i]]) {
            g.pair_chan.push_back(kv.first);
            g.pair_w.push_back(kv.second);
        }
        g.slot_ptr.push_back((int32_t)g.pair_chan.size());
    }
    g.imag_only = true;
    for (const auto &w : g.pair_w)
        if (w.real() != 0.0) g.imag_only = false;
    // Row layout: entry 0 is the diagonal (an explicit zero, slot 0xFFFF, if the row stores none), so the
    // kernels can take that operand from registers.  Off-diagonal entries are aligned by diagonal offset
    // (src - row) when that does not lengthen the rows: entry j of every row then reads a shifted copy of
    // the row index, which keeps the shared-memory gathers of a quarter warp on distinct banks.
    g.nnz = 0;
    int plain_max = 0;
    std::vector<int> offsets;
    for (int r = 0; r < n; r++) {
        int cnt = 1;
        for (auto &e : row_entries[r])
            if (e.first != r) {
                cnt++;
                offsets.push_back(e.first - r);
            }
        plain_max = std::max(plain_max, cnt);
        g.nnz += (int64_t)row_entries[r].size();
    }
    std::sort(offsets.begin(), offsets.end());
    offsets.erase(std::unique(offsets.begin(), offsets.end()), offsets.end());
    const bool aligned = 1 + (int)offsets.size() <= plain_max;
    g.max_nnz = aligned ? 1 + (int)offsets.size() : plain_max;
    if (g.max_nnz > 255) throw std::runtime_error("generator row longer than 255 entries");
    g.row_len.assign((size_t)n, (uint8_t)g.max_nnz);
    g.ell.assign((size_t)g.max_nnz * n, 0xFFFFFFFFu);
    for (int r = 0; r < n; r++) {
        uint32_t diag = (uint32_t)r | (0xFFFFu << 16);
        int next = 1;
        for (auto &e : row_entries[r]) {
            const uint32_t packed = (uint32_t)e.first | ((uint32_t)new_id[e.second] << 16);
            if (e.first == r) {
                diag = packed;
            } else {
                const int j = aligned ? 1 + (int)(std::lower_bound(offsets.begin(), offsets.end(), e.first - r) -
                                                  offsets.begin())
                                      : next++;
                g.ell[(size_t)j * n + r] = packed;
            }
        }
        g.ell[r] = diag;
        // padding entries carry slot 0xFFFF (zero).  In aligned mode they still read "their" diagonal's
        // source row (shifted by multiples of 8 rows into range), which keeps the bank pattern regular.
        for (int j = 1; j < g.max_nnz; j++) {
            if (g.ell[(size_t)j * n + r] != 0xFFFFFFFFu) continue;
            int src = r;
            if (aligned && n >= 8) {
                src = r + offsets[j - 1];
                while (src >= n) src -= 8;
                while (src < 0) src += 8;
            }
            g.ell[(size_t)j * n + r] = (uint32_t)src | (0xFFFFu << 16);
        }
    }
}

// Dense A(t) from the compiled generator (tests / introspection).
void host_eval_generator(const Generator &g, double t, const double *row, cplx *dense) {
    const SeriesTables &tab = series_tables();
    std::vector<double> chan((size_t)g.n_chan, 0.0);
    chan[0] = 1.0;
    for (const auto &e : g.evals) {
        ResolvedEval r;
        resolve_eval(e, row, g.tables.empty() ? nullptr : g.tables.data(), r);
        double o0, o1;
        run_resolved(r, tab.freq, tab.anh, t, &o0, &o1);
        chan[r.out_chan] = o0;
        if (eval_outputs(r.kind) == 2) chan[r.out_chan + 1] = o1;
    }
    std::vector<cplx> V((size_t)g.n_slots);
    for (int s = 0; s < g.n_slots; s++) {
        cplx v(0, 0);
        for (int q = g.slot_ptr[s]; q < g.slot_ptr[s + 1]; q++) v += g.pair_w[q] * chan[g.pair_chan[q]];
        V[s] = v;
    }
    const int n = g.n;
    for (size_t q = 0; q < (size_t)n * n; q++) dense[q] = cplx(0, 0);
    for (int r = 0; r < n; r++)
        for (int j = 0; j < g.max_nnz; j++) {
            uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) == 0xFFFFu) continue;  // padding / explicit zero diagonal
            dense[(size_t)r + (size_t)(e & 0xFFFF) * n] += V[e >> 16];
        }
}

}  // namespace qsim
