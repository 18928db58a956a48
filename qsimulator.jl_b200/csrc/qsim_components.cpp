// Component pruning of propagator problems: connected components of the generator's graph and the packing of
// (component, column group) units into the warps of a team.  Host-side only.
//
// U(t) solves dU/dt = A(t) U, U(0) = I.  If the pattern of A (all stored entries, whatever their time dependence)
// splits the rows into connected components C_1..C_m, then A is block diagonal after a permutation and so is U:
// U[r, c] = 0 whenever r and c lie in different components.  The reference computes those entries too
// (time_evolution.jl:31-37 dense H*u, :98-111 dense superoperator), but every term it adds there is 0 * x, so they
// stay exactly zero, and they enter the error norms (DiffEqBase ODE_DEFAULT_NORM over all n^2 entries) as zeros.
// The 2- and 3-transmon gates of benchmarks.jl conserve the parity of the total excitation number (two components:
// half the entries), their rotating-frame exchange form conserves the excitation number itself (the d^2 = 729
// Lindblad generator of cfg5 has 13 components and 14 % of the propagator's entries are non-zero).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <vector>

#include "qsim_types.h"

namespace qsim {

void generator_components(const Generator &g, std::vector<int> &comp_of, std::vector<std::vector<int>> &comps) {
    const int n = g.n;
    std::vector<int> parent((size_t)n);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int x) {
        while (parent[(size_t)x] != x) x = parent[(size_t)x] = parent[(size_t)parent[(size_t)x]];
        return x;
    };
    for (int j = 0; j < g.max_nnz; j++)
        for (int r = 0; r < n; r++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if (e == 0xFFFFFFFFu || (e >> 16) == 0xFFFFu) continue;   // padding / zero entry
            const int a = find(r), b = find((int)(e & 0xFFFFu));
            if (a != b) parent[(size_t)std::max(a, b)] = std::min(a, b);
        }
    comp_of.assign((size_t)n, -1);
    comps.clear();
    for (int r = 0; r < n; r++) {
        const int root = find(r);
        if (comp_of[(size_t)root] < 0) {
            comp_of[(size_t)root] = (int)comps.size();
            comps.emplace_back();
        }
        comp_of[(size_t)r] = comp_of[(size_t)root];
        comps[(size_t)comp_of[(size_t)r]].push_back(r);
    }
}

// Shared-memory wavefronts per stage of a component form's gathers and stores for stage row stride S (16-byte units):
// a 128-bit shared access is served a quarter warp at a time, and lanes of a quarter that hit the same bank group
// (unit mod 8) at different addresses serialise.  Padding entries (zero slot) and lanes without a row are free: the
// tables make them repeat an address their quarter reads anyway (resolve_padding below).
static long warp_cost(const ComponentForm &cf, const std::vector<uint32_t> &ell, const std::vector<int> &vrow, int w, int S) {
    const int W = cf.team_warps, nl = cf.n_loc;
    long cost = 0;
    for (int i = 0; i < cf.rpl; i++)
        for (int q = 0; q < 4; q++) {
            if (8 * q + 32 * i >= nl) {   // lanes past the last local row repeat its addresses: one wavefront per access
                cost += cf.max_nnz;
                continue;
            }
            for (int j = 1; j <= cf.max_nnz; j++) {   // j == max_nnz: the lanes' own stores
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, seen[8][8];
                for (int l = 8 * q; l < 8 * q + 8; l++) {
                    const int r = l + 32 * i;
                    if (r >= nl || vrow[(size_t)w * nl + r] < 0) continue;
                    int src = r;
                    if (j < cf.max_nnz) {
                        const uint32_t e = ell[((size_t)j * W + w) * nl + r];
                        if ((e >> 16) == 0xFFFFu) continue;
                        src = (int)(e & 0xFFFFu);
                    }
                    const int unit = src * S, b = unit & 7;
                    bool dup = false;
                    for (int k = 0; k < cnt[b]; k++) dup = dup || seen[b][k] == unit;
                    if (!dup) seen[b][cnt[b]++] = unit;
                }
                int m = 1;
                for (int b = 0; b < 8; b++) m = std::max(m, cnt[b]);
                cost += m;
            }
        }
    return cost * cf.cpl;
}

long component_layout_cost(const ComponentForm &cf, int S) {
    long c = 0;
    for (int w = 0; w < cf.team_warps; w++) c += warp_cost(cf, cf.ell, cf.vrow, w, S);
    return c;
}

// Padding entries (zero slot) and local rows without a state row read an address that is read or written anyway: the
// first source their quarter warp really gathers at that entry (a broadcast: no extra wavefront), else the position
// of a row of the quarter (of the warp, if the quarter has none).  They must not read a position nobody stores to:
// the operand is multiplied by zero, but uninitialised shared memory may hold a NaN.
static void resolve_padding(ComponentForm &cf) {
    const int W = cf.team_warps, nl = cf.n_loc;
    for (int w = 0; w < W; w++) {
        int first_row = -1;
        for (int r = 0; r < nl && first_row < 0; r++)
            if (cf.vrow[(size_t)w * nl + r] >= 0) first_row = r;
        if (first_row < 0) first_row = 0;
        for (int j = 0; j < cf.max_nnz; j++)
            for (int r0 = 0; r0 < nl; r0 += 8) {
                const int r1 = std::min(nl, r0 + 8);
                int real = -1, anchor = first_row;
                for (int r = r1 - 1; r >= r0; r--)
                    if (cf.vrow[(size_t)w * nl + r] >= 0) anchor = r;
                for (int r = r0; r < r1 && real < 0 && j > 0; r++) {
                    const uint32_t e = cf.ell[((size_t)j * W + w) * nl + r];
                    if (cf.vrow[(size_t)w * nl + r] >= 0 && (e >> 16) != 0xFFFFu) real = (int)(e & 0xFFFFu);
                }
                for (int r = r0; r < r1; r++) {
                    uint32_t &e = cf.ell[((size_t)j * W + w) * nl + r];
                    const bool hole = cf.vrow[(size_t)w * nl + r] < 0;
                    if (j == 0) {
                        if (hole) e = (uint32_t)anchor | (0xFFFFu << 16);   // (the diagonal operand comes from registers)
                    } else if (hole || (e >> 16) == 0xFFFFu)
                        e = (uint32_t)(real >= 0 ? real : anchor) | (0xFFFFu << 16);
                }
            }
    }
}

// Chooses, per warp, which local row (= lane, = position in the stage buffer) every row of the warp's units takes, by
// simulated annealing over swaps against warp_cost, and the stage row stride; then resolves the padding entries.
// The local rows are first extended to a multiple of 8 (whole quarter warps), which gives the search room.
void optimize_component_layout(ComponentForm &cf) {
    const int W = cf.team_warps, cap = 32 * cf.rpl, old_nl = cf.n_loc;
    const int nl = std::min(cap, (old_nl + 7) & ~7);
    if (nl != old_nl) {
        std::vector<uint32_t> ell((size_t)cf.max_nnz * W * nl);
        std::vector<int> vrow((size_t)W * nl, -1), vcol((size_t)cf.n_batches * nl * cf.cpl, -1);
        for (int w = 0; w < W; w++)
            for (int r = 0; r < nl; r++)
                for (int j = 0; j < cf.max_nnz; j++)
                    ell[((size_t)j * W + w) * nl + r] = r < old_nl ? cf.ell[((size_t)j * W + w) * old_nl + r] : ((uint32_t)r | (0xFFFFu << 16));
        for (int w = 0; w < W; w++)
            for (int r = 0; r < old_nl; r++) vrow[(size_t)w * nl + r] = cf.vrow[(size_t)w * old_nl + r];
        for (int b = 0; b < cf.n_batches; b++)
            for (int r = 0; r < old_nl; r++)
                for (int q = 0; q < cf.cpl; q++) vcol[((size_t)b * nl + r) * cf.cpl + q] = cf.vcol[((size_t)b * old_nl + r) * cf.cpl + q];
        cf.ell.swap(ell);
        cf.vrow.swap(vrow);
        cf.vcol.swap(vcol);
        cf.n_loc = nl;
    }
    // stride: odd strides keep the lanes' own stores conflict-free; take the best of a few before the search
    int S = cf.cpl | 1;
    {
        long best = -1;
        for (int s = cf.cpl; s <= cf.cpl + 8; s++) {
            if (16 * nl * s > 65535) break;
            const long c = component_layout_cost(cf, s) * 64 + s;
            if (s % 2 == 1 && (best < 0 || c < best)) {
                best = c;
                S = s;
            }
        }
    }
    cf.stride = S;
    const char *env = std::getenv("QSIM_CF_ANNEAL");   // development: search iterations per local row (0: none)
    const long iters_per_row = env ? std::atol(env) : 400;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() {
        rng ^= rng << 13;
        rng ^= rng >> 7;
        rng ^= rng << 17;
        return (uint32_t)(rng >> 32);
    };
    const std::vector<uint32_t> nat_ell = cf.ell;   // tables in the packing's natural order
    const std::vector<int> nat_vrow = cf.vrow, nat_vcol = cf.vcol;
    std::vector<uint32_t> ell = nat_ell;
    std::vector<int> vrow = nat_vrow;
    // perm[t] = local row of natural row t; writes warp w's part of ell / vrow
    auto apply = [&](int w, const std::vector<int> &pm) {
        for (int t = 0; t < nl; t++) {
            vrow[(size_t)w * nl + pm[(size_t)t]] = nat_vrow[(size_t)w * nl + t];
            for (int j = 0; j < cf.max_nnz; j++) {
                const uint32_t e = nat_ell[((size_t)j * W + w) * nl + t];
                ell[((size_t)j * W + w) * nl + pm[(size_t)t]] = (uint32_t)pm[(size_t)(e & 0xFFFFu)] | (e & 0xFFFF0000u);
            }
        }
    };
    std::vector<std::vector<int>> perms((size_t)W);
    for (int w = 0; w < W; w++) {
        std::vector<int> &perm = perms[(size_t)w];
        // warps with identical tables (the warps of one streamed config) share one search
        int like = -1;
        for (int v = 0; v < w && like < 0; v++) {
            bool same = true;
            for (int r = 0; r < nl && same; r++) {
                same = nat_vrow[(size_t)v * nl + r] == nat_vrow[(size_t)w * nl + r];
                for (int j = 0; j < cf.max_nnz && same; j++)
                    same = nat_ell[((size_t)j * W + v) * nl + r] == nat_ell[((size_t)j * W + w) * nl + r];
            }
            if (same) like = v;
        }
        if (like >= 0) {
            perm = perms[(size_t)like];
            apply(w, perm);
            continue;
        }
        perm.resize((size_t)nl);
        std::iota(perm.begin(), perm.end(), 0);
        apply(w, perm);
        long cur = warp_cost(cf, ell, vrow, w, S), best = cur;
        std::vector<int> best_perm = perm;
        const long iters = iters_per_row * nl;
        for (long it = 0; it < iters; it++) {
            const int a = (int)(next() % (uint32_t)nl), b = (int)(next() % (uint32_t)nl);
            if (a == b) continue;
            std::swap(perm[(size_t)a], perm[(size_t)b]);
            apply(w, perm);
            const long c = warp_cost(cf, ell, vrow, w, S);
            const double T = 1.5 * (1.0 - (double)it / (double)iters) + 0.05;
            if (c <= cur || (next() >> 8) * (1.0 / 16777216.0) < std::exp(-(double)(c - cur) / (T * cf.cpl))) {
                cur = c;
                if (c < best) {
                    best = c;
                    best_perm = perm;
                }
            } else
                std::swap(perm[(size_t)a], perm[(size_t)b]);
        }
        perm = best_perm;
        apply(w, perm);
    }
    cf.ell = ell;
    cf.vrow = vrow;
    for (int b = 0; b < cf.n_batches; b++) {
        const std::vector<int> &pm = perms[(size_t)(b % W)];
        for (int t = 0; t < nl; t++)
            for (int q = 0; q < cf.cpl; q++) cf.vcol[((size_t)b * nl + pm[(size_t)t]) * cf.cpl + q] = nat_vcol[((size_t)b * nl + t) * cf.cpl + q];
    }
    resolve_padding(cf);
}

// Packs the units of a propagator into warps for a kernel shape of `rpl` rows x `cpl` columns per lane.
//   resident: every unit gets its own place, at most max_team_warps warps (first-fit, largest components first);
//   streamed: otherwise.  Components are bin-packed into CONFIGS (sets of components whose rows fit one warp
//             together); a batch of a config integrates one unit of each of its components; the team's warps
//             are dealt to the configs in proportion to their batch counts (a warp's rows are fixed for the whole
//             kernel), warp w runs batches w, w + team_warps, ...
// Returns false when the generator is connected (nothing to prune) or the form does not fit the shape.
bool build_component_form(const Generator &g, int rpl, int cpl, int max_team_warps, ComponentForm &cf) {
    const int n = g.n, cap = 32 * rpl;
    std::vector<std::vector<int>> comps;
    generator_components(g, cf.comp_of, comps);
    cf.ok = false;
    cf.n_comp = (int)comps.size();
    if (comps.size() < 2) return false;
    cf.rpl = rpl;
    cf.cpl = cpl;
    cf.max_nnz = g.max_nnz;
    std::vector<int> order(comps.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comps[(size_t)a].size() > comps[(size_t)b].size(); });
    if ((int)comps[(size_t)order[0]].size() > cap) return false;
    std::vector<int> index_in_comp((size_t)n, 0);
    cf.nnz_cols = 0;
    cf.entries = 0;
    for (const auto &c : comps) {
        for (size_t i = 0; i < c.size(); i++) index_in_comp[(size_t)c[i]] = (int)i;
        int64_t nnz = 0;
        for (int r : c)
            for (int j = 0; j < g.max_nnz; j++) {
                const uint32_t e = g.ell[(size_t)j * n + r];
                if (e != 0xFFFFFFFFu && (e >> 16) != 0xFFFFu) nnz++;
            }
        cf.nnz_cols += nnz * (int64_t)c.size();
        cf.entries += (int64_t)c.size() * (int64_t)c.size();
    }
    struct Place { int comp, first_col, base; };   // a unit inside a warp (resident) / a component inside a config (streamed)
    std::vector<std::vector<Place>> warps;          // per warp: its units / its config's components
    std::vector<int> used;
    // ---- resident packing: all units at once ----
    for (int k : order) {
        const int nk = (int)comps[(size_t)k].size();
        for (int c0 = 0; c0 < nk; c0 += cpl) {
            size_t w = 0;
            while (w < warps.size() && used[w] + nk > cap) w++;
            if (w == warps.size()) {
                warps.emplace_back();
                used.push_back(0);
            }
            warps[w].push_back({k, c0, used[w]});
            used[w] += nk;
        }
    }
    std::vector<int> batch_warp;    // streamed: config of every warp
    std::vector<std::vector<Place>> configs;
    std::vector<int> rounds;
    if ((int)warps.size() <= max_team_warps) {
        cf.streamed = 0;
        cf.team_warps = (int)warps.size();
        cf.n_batches = cf.team_warps;
    } else {
        // ---- streamed packing: components into configs, warps dealt to configs ----
        std::vector<int> cused;
        for (int k : order) {
            const int nk = (int)comps[(size_t)k].size();
            size_t c = 0;
            while (c < configs.size() && cused[c] + nk > cap) c++;
            if (c == configs.size()) {
                configs.emplace_back();
                cused.push_back(0);
            }
            configs[c].push_back({k, 0, cused[c]});
            cused[c] += nk;
        }
        const int nc = (int)configs.size();
        if (nc > max_team_warps) return false;
        rounds.assign((size_t)nc, 0);
        for (int c = 0; c < nc; c++)
            for (const auto &pl : configs[(size_t)c])
                rounds[(size_t)c] = std::max(rounds[(size_t)c], ((int)comps[(size_t)pl.comp].size() + cpl - 1) / cpl);
        // deal the warps: one each, then always to the config with the most batches per warp
        std::vector<int> nw((size_t)nc, 1);
        for (int left = max_team_warps - nc; left > 0; left--) {
            int best = 0;
            for (int c = 1; c < nc; c++)
                if ((rounds[(size_t)c] + nw[(size_t)c] - 1) / nw[(size_t)c] > (rounds[(size_t)best] + nw[(size_t)best] - 1) / nw[(size_t)best]) best = c;
            nw[(size_t)best]++;
        }
        cf.streamed = 1;
        cf.team_warps = max_team_warps;
        warps.clear();
        used.clear();
        int passes = 0;
        for (int c = 0; c < nc; c++) {
            passes = std::max(passes, (rounds[(size_t)c] + nw[(size_t)c] - 1) / nw[(size_t)c]);
            for (int q = 0; q < nw[(size_t)c]; q++) {
                warps.push_back(configs[(size_t)c]);
                used.push_back(cused[(size_t)c]);
                batch_warp.push_back(c);
            }
        }
        cf.n_batches = passes * cf.team_warps;
    }
    const int W = cf.team_warps;
    cf.n_loc = *std::max_element(used.begin(), used.end());
    const int nl = cf.n_loc;
    if (nl > cap) return false;
    cf.vrow.assign((size_t)W * nl, -1);
    cf.ell.assign((size_t)g.max_nnz * W * nl, 0);
    for (int w = 0; w < W; w++) {
        for (const auto &pl : warps[(size_t)w]) {
            const auto &c = comps[(size_t)pl.comp];
            for (size_t i = 0; i < c.size(); i++) {
                const int lr = pl.base + (int)i, r = c[i];
                cf.vrow[(size_t)w * nl + lr] = r;
                for (int j = 0; j < g.max_nnz; j++) {
                    const uint32_t e = g.ell[(size_t)j * n + r];
                    uint32_t v;
                    if (e == 0xFFFFFFFFu || (e >> 16) == 0xFFFFu)
                        v = (uint32_t)lr | (0xFFFFu << 16);   // padding / zero entry: own row, zero slot
                    else
                        v = (uint32_t)(pl.base + index_in_comp[(size_t)(e & 0xFFFFu)]) | (e & 0xFFFF0000u);
                    cf.ell[(size_t)j * W * nl + (size_t)w * nl + lr] = v;
                }
            }
        }
        // local rows past the warp's last one: the last row's position, zero slot (their loads coincide with its loads)
        const int last = std::max(used[(size_t)w] - 1, 0);
        for (int lr = used[(size_t)w]; lr < nl; lr++)
            for (int j = 0; j < g.max_nnz; j++) cf.ell[(size_t)j * W * nl + (size_t)w * nl + lr] = (uint32_t)last | (0xFFFFu << 16);
    }
    cf.vcol.assign((size_t)cf.n_batches * nl * cpl, -1);
    if (!cf.streamed) {
        for (int w = 0; w < W; w++)
            for (const auto &pl : warps[(size_t)w]) {
                const auto &c = comps[(size_t)pl.comp];
                for (size_t i = 0; i < c.size(); i++)
                    for (int q = 0; q < cpl; q++)
                        if (pl.first_col + q < (int)c.size())
                            cf.vcol[((size_t)w * nl + pl.base + i) * cpl + q] = c[(size_t)(pl.first_col + q)];
            }
    } else {
        // config c's warps share its rounds: the j-th warp of the config takes rounds j, j + nw, ... in its passes
        std::vector<int> seen((size_t)configs.size(), 0), nwc((size_t)configs.size(), 0);
        for (int w = 0; w < W; w++) nwc[(size_t)batch_warp[(size_t)w]]++;
        for (int w = 0; w < W; w++) {
            const int c = batch_warp[(size_t)w], j = seen[(size_t)c]++;
            for (int pass = 0; pass * W + w < cf.n_batches; pass++) {
                const int b = pass * W + w, round = pass * nwc[(size_t)c] + j;
                for (const auto &pl : configs[(size_t)c]) {
                    const auto &cc = comps[(size_t)pl.comp];
                    for (size_t i = 0; i < cc.size(); i++)
                        for (int q = 0; q < cpl; q++) {
                            const int col = round * cpl + q;
                            if (col < (int)cc.size()) cf.vcol[((size_t)b * nl + pl.base + i) * cpl + q] = cc[(size_t)col];
                        }
                }
            }
        }
    }
    optimize_component_layout(cf);
    cf.zeros.clear();
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++)
            if (cf.comp_of[(size_t)r] != cf.comp_of[(size_t)c]) cf.zeros.push_back(c * n + r);
    if (std::getenv("QSIM_DEBUG")) {
        // self-check: every entry inside a component block is held by exactly one (batch, local row, c)
        std::vector<int> cover((size_t)n * n, 0);
        const int nl2 = cf.n_loc;
        long bad = 0;
        for (int b = 0; b < cf.n_batches; b++)
            for (int r = 0; r < nl2; r++)
                for (int q = 0; q < cpl; q++) {
                    const int col = cf.vcol[((size_t)b * nl2 + r) * cpl + q], row = cf.vrow[(size_t)(b % W) * nl2 + r];
                    if (col >= 0 && row < 0) bad++;
                    if (col >= 0 && row >= 0) cover[(size_t)col * n + row]++;
                }
        for (int c = 0; c < n; c++)
            for (int r = 0; r < n; r++)
                if (cover[(size_t)c * n + r] != (cf.comp_of[(size_t)r] == cf.comp_of[(size_t)c] ? 1 : 0)) bad++;
        std::fprintf(stderr, "[qsim] component form r%dc%d self-check: %ld bad entries\n", rpl, cpl, bad);
    }
    cf.gv = g;
    cf.gv.n = cf.n_loc;
    cf.gv.ell = cf.ell;
    cf.gv.row_len.assign((size_t)W * cf.n_loc, (uint8_t)g.max_nnz);
    cf.ok = true;
    return true;
}

// Row-split form: passes over several units at once (see ComponentForm).  Units are placed largest component first;
// a pass is filled with as many units of the current component as fit its rpl * team_warps chunks, then topped up with
// units of smaller components.  Within a pass the chunks are dealt to (row-slot, warp) longest rows first, so the
// warps of the team carry about equal work.
bool build_rowsplit_component_form(const Generator &g, int team_warps, ComponentForm &cf, const std::vector<char> *touched, int rpl) {
    const int n = g.n, W = team_warps, slots = rpl * W;
    std::vector<std::vector<int>> comps;
    generator_components(g, cf.comp_of, comps);
    cf.ok = false;
    cf.n_comp = (int)comps.size();
    if (comps.size() < 2) return false;
    std::vector<int> order(comps.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return comps[(size_t)a].size() > comps[(size_t)b].size(); });
    if (((int)comps[(size_t)order[0]].size() + 31) / 32 > slots) return false;
    cf.rs = 1;
    cf.state = touched ? 1 : 0;
    if (touched) {
        cf.live.assign((size_t)n, 0);
        for (int r = 0; r < n; r++) cf.live[(size_t)r] = (*touched)[(size_t)cf.comp_of[(size_t)r]];
    }
    cf.rpl = rpl;
    cf.cpl = 1;
    cf.team_warps = W;
    cf.n_loc = 32 * rpl * W;
    cf.stride = 1;
    cf.max_nnz = g.max_nnz;
    // components in placement order, their generators and chunk lengths (rows sorted by length, as rowsplit_rows does)
    const int nk_total = (int)comps.size();
    cf.subs.assign((size_t)nk_total, Generator());
    cf.comp_rows.assign((size_t)nk_total, {});
    cf.cc_first.assign((size_t)nk_total, 0);
    std::vector<std::vector<int>> chunk_len((size_t)nk_total);
    std::vector<int> local((size_t)n, 0);
    cf.nnz_cols = 0;
    cf.entries = 0;
    int n_cc = 0;
    for (int ki = 0; ki < nk_total; ki++) {
        const auto &rows = comps[(size_t)order[(size_t)ki]];
        const int nk = (int)rows.size();
        cf.comp_rows[(size_t)ki] = rows;
        for (int i = 0; i < nk; i++) local[(size_t)rows[(size_t)i]] = i;
        Generator &sub = cf.subs[(size_t)ki];
        sub = g;
        sub.n = nk;
        sub.ell.assign((size_t)g.max_nnz * nk, 0);
        sub.row_len.assign((size_t)nk, (uint8_t)g.max_nnz);
        std::vector<int> len((size_t)nk, 1);
        int64_t nnz = 0;
        for (int i = 0; i < nk; i++)
            for (int j = 0; j < g.max_nnz; j++) {
                const uint32_t e = g.ell[(size_t)j * n + rows[(size_t)i]];
                const bool pad = e == 0xFFFFFFFFu || (e >> 16) == 0xFFFFu;
                sub.ell[(size_t)j * nk + i] = pad ? ((uint32_t)i | (0xFFFFu << 16)) : ((uint32_t)local[(size_t)(e & 0xFFFFu)] | (e & 0xFFFF0000u));
                if (!pad) nnz++;
                if (!pad && j > 0) len[(size_t)i]++;
            }
        sub.nnz = nnz;
        const bool live_k = !touched || (*touched)[(size_t)order[(size_t)ki]];
        cf.nnz_cols += touched ? (live_k ? nnz : 0) : nnz * nk;
        cf.entries += touched ? (live_k ? nk : 0) : (int64_t)nk * nk;
        std::sort(len.begin(), len.end(), std::greater<int>());
        cf.cc_first[(size_t)ki] = n_cc;
        for (int c0 = 0; c0 < nk; c0 += 32) chunk_len[(size_t)ki].push_back(len[(size_t)c0]);
        n_cc += (nk + 31) / 32;
    }
    cf.n_cc = n_cc;
    // ---- passes ----
    // units still to place per component: all its columns (propagators), or the one state column if the state touches it
    std::vector<int> next_col((size_t)nk_total, 0), n_units((size_t)nk_total, 0);
    for (int ki = 0; ki < nk_total; ki++)
        n_units[(size_t)ki] = touched ? ((*touched)[(size_t)order[(size_t)ki]] ? 1 : 0) : (int)cf.comp_rows[(size_t)ki].size();
    cf.pass.clear();
    int n_pass = 0;
    for (;;) {
        struct Chunk { int cc, ubase, col, len; };
        std::vector<Chunk> chunks;
        int free_slots = slots, ubase = 0;
        for (int ki = 0; ki < nk_total; ki++) {
            const int nk = (int)cf.comp_rows[(size_t)ki].size(), ncc = (nk + 31) / 32;
            while (next_col[(size_t)ki] < n_units[(size_t)ki] && ncc <= free_slots) {
                const int col = touched ? 0 : cf.comp_rows[(size_t)ki][(size_t)next_col[(size_t)ki]];
                next_col[(size_t)ki]++;
                for (int t = 0; t < ncc; t++) chunks.push_back({cf.cc_first[(size_t)ki] + t, ubase, col, chunk_len[(size_t)ki][(size_t)t]});
                free_slots -= ncc;
                ubase += (nk + 7) & ~7;
            }
        }
        if (chunks.empty()) break;
        std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk &a, const Chunk &b) { return a.len > b.len; });
        for (int c = 0; c < slots; c++) {
            if (c < (int)chunks.size()) {
                cf.pass.push_back(chunks[(size_t)c].cc);
                cf.pass.push_back(chunks[(size_t)c].ubase);
                cf.pass.push_back(chunks[(size_t)c].col);
            } else {
                cf.pass.push_back(n_cc);   // the empty chunk
                cf.pass.push_back(0);
                cf.pass.push_back(0);
            }
            cf.pass.push_back(0);
        }
        n_pass++;
    }
    if (n_pass == 0) return false;
    cf.n_batches = n_pass;
    cf.streamed = n_pass > 1 ? 1 : 0;
    cf.zeros.clear();
    if (touched) {
        for (int r = 0; r < n; r++)
            if (!cf.live[(size_t)r]) cf.zeros.push_back(r);
    } else
    for (int c = 0; c < n; c++)
        for (int r = 0; r < n; r++)
            if (cf.comp_of[(size_t)r] != cf.comp_of[(size_t)c]) cf.zeros.push_back(c * n + r);
    if (std::getenv("QSIM_DEBUG")) {
        // self-check: every (component, column) unit -- all of the component's chunks -- is placed exactly once
        // (state forms: column 0 of the touched components only)
        long bad = 0;
        std::vector<int> comp_of_cc((size_t)n_cc, 0);
        for (int ki = 0; ki < nk_total; ki++)
            for (int t = 0; t < ((int)cf.comp_rows[(size_t)ki].size() + 31) / 32; t++) comp_of_cc[(size_t)(cf.cc_first[(size_t)ki] + t)] = ki;
        std::vector<int> seen((size_t)n_cc * n, 0);
        for (size_t i = 0; i < cf.pass.size(); i += 4)
            if (cf.pass[i] != n_cc) seen[(size_t)cf.pass[i] * n + (size_t)cf.pass[i + 2]]++;
        for (int cc = 0; cc < n_cc; cc++) {
            const int ki = comp_of_cc[(size_t)cc];
            for (int c = 0; c < n; c++) {
                bool want;
                if (touched)
                    want = c == 0 && (*touched)[(size_t)order[(size_t)ki]];
                else
                    want = std::find(cf.comp_rows[(size_t)ki].begin(), cf.comp_rows[(size_t)ki].end(), c) != cf.comp_rows[(size_t)ki].end();
                if (seen[(size_t)cc * n + c] != (want ? 1 : 0)) bad++;
            }
        }
        std::fprintf(stderr, "[qsim] component form row-split self-check: %ld bad entries\n", bad);
        long used = 0;
        for (size_t i = 0; i < cf.pass.size(); i += 4) used += cf.pass[i] != n_cc ? 1 : 0;
        std::fprintf(stderr, "[qsim] row-split component form: %d components (largest %zu rows), %d component chunks, %d passes of %d chunk "
                             "slots (%ld filled) instead of %d columns\n", cf.n_comp, comps[(size_t)order[0]].size(), n_cc, n_pass, slots, used, n);
    }
    cf.gv = g;
    cf.gv.n = cf.n_loc;
    cf.ok = true;
    return true;
}

// State form: the rows of the components the initial state touches, one column.  Small live sets (at most 96 rows, rows
// of at most 16 entries) run as ONE warp of the row-per-lane kernels whatever the size of the full state; larger ones as
// passes of the row-split kernels (usually one).
bool build_state_form(const Generator &g, const std::vector<char> &touched, ComponentForm &cf) {
    const int n = g.n;
    std::vector<std::vector<int>> comps;
    generator_components(g, cf.comp_of, comps);
    cf.ok = false;
    cf.n_comp = (int)comps.size();
    if (comps.size() < 2 || touched.size() != comps.size()) return false;
    std::vector<int> rows;
    for (size_t k = 0; k < comps.size(); k++)
        if (touched[k]) rows.insert(rows.end(), comps[k].begin(), comps[k].end());
    const int nl = (int)rows.size();
    if (nl == 0 || nl == n) return false;
    if (nl > 96 || g.max_nnz > 16) {
        if (!build_rowsplit_component_form(g, 8, cf, &touched)) return false;
        if (std::getenv("QSIM_DEBUG"))
            std::fprintf(stderr, "[qsim] state form: %d of %d rows live (%d components), row-split, %d pass(es)\n", nl, n, cf.n_comp, cf.n_batches);
        return true;
    }
    std::sort(rows.begin(), rows.end());
    cf.state = 1;
    cf.rs = 0;
    cf.rpl = (nl + 31) / 32;
    cf.cpl = 1;
    cf.team_warps = 1;
    cf.n_batches = 1;
    cf.streamed = 0;
    cf.max_nnz = g.max_nnz;
    cf.n_loc = nl;
    cf.live.assign((size_t)n, 0);
    std::vector<int> local((size_t)n, -1);
    for (int i = 0; i < nl; i++) {
        cf.live[(size_t)rows[(size_t)i]] = 1;
        local[(size_t)rows[(size_t)i]] = i;
    }
    cf.vrow = rows;
    cf.vcol.assign((size_t)nl, 0);
    cf.ell.assign((size_t)g.max_nnz * nl, 0);
    cf.nnz_cols = 0;
    for (int i = 0; i < nl; i++)
        for (int j = 0; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + rows[(size_t)i]];
            const bool pad = e == 0xFFFFFFFFu || (e >> 16) == 0xFFFFu;
            cf.ell[(size_t)j * nl + i] = pad ? ((uint32_t)i | (0xFFFFu << 16)) : ((uint32_t)local[(size_t)(e & 0xFFFFu)] | (e & 0xFFFF0000u));
            if (!pad) cf.nnz_cols++;
        }
    cf.entries = nl;
    optimize_component_layout(cf);
    cf.zeros.clear();
    for (int r = 0; r < n; r++)
        if (!cf.live[(size_t)r]) cf.zeros.push_back(r);
    cf.gv = g;
    cf.gv.n = cf.n_loc;
    cf.gv.ell = cf.ell;
    cf.gv.row_len.assign((size_t)cf.n_loc, (uint8_t)g.max_nnz);
    cf.ok = true;
    if (std::getenv("QSIM_DEBUG"))
        std::fprintf(stderr, "[qsim] state form: %d of %d rows live (%d components), one warp, %d row(s) per lane\n", nl, n, cf.n_comp, cf.rpl);
    return true;
}

}  // namespace qsim
