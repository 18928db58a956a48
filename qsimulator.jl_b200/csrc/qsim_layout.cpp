// Host-side layout optimisation for the row-split kernels (state vectors longer than 96 entries).
//
// Those kernels are bound by shared-memory wavefronts: per stage and column every lane gathers, for each stored
// entry of its rows, a 16-byte operand from the stage vector and a 16-byte value from the per-stage value table.
// A 128-bit shared access is served a quarter warp (8 lanes) at a time and the 8 lanes conflict when two different
// addresses fall into the same 16-byte bank group (address / 16 mod 8).  Rows are dealt to lanes sorted by length
// (so that a warp executes no padding), which scatters the sources of a quarter warp over the stage vector: with
// the stage vector stored in lane order the measured cost is 1.9 wavefronts per ideal wavefront
// (profiles/r2_ncu_cfg5_*).  Both tables are only ever addressed through host-built offsets, so the position of a
// row inside the stage buffer and the position of a value slot inside the value table are free:
//   * anneal_colors() assigns every item (row / slot) one of 8 bank groups such that the items read together by a
//     quarter warp are spread over distinct groups (simulated annealing over colour swaps, deterministic seed);
//   * the caller turns colours into positions (position mod 8 == colour).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "qsim_types.h"

namespace qsim {

namespace {
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t next() {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        return (uint32_t)(s >> 32);
    }
    double uniform() { return (next() >> 8) * (1.0 / 16777216.0); }
};

inline int group_cost(const std::vector<int> &g, const std::vector<int> &color, int n_colors) {
    int cnt[16] = {0};
    int m = 0;
    for (int it : g) m = std::max(m, ++cnt[color[(size_t)it] % n_colors]);
    return m;
}
}  // namespace

// Sum over the groups of the largest number of members sharing one colour (= wavefronts of the accesses).
long layout_groups_cost(const std::vector<std::vector<int>> &groups, const std::vector<int> &color, int n_colors) {
    long t = 0;
    for (const auto &g : groups) t += group_cost(g, color, n_colors);
    return t;
}

// Swap-based simulated annealing.  color[i] in [0, n_colors); only items of the same class are swapped, so the
// number of items per (class, colour) -- which the caller's position ranges provide -- never changes.
void anneal_colors(const std::vector<std::vector<int>> &groups, const std::vector<int> &cls, int n_colors, long iters,
                   std::vector<int> &color) {
    const int n = (int)color.size();
    if (n < 2 || groups.empty()) return;
    std::vector<std::vector<int>> member((size_t)n);
    for (size_t gi = 0; gi < groups.size(); gi++)
        for (int it : groups[gi])
            if (member[(size_t)it].empty() || member[(size_t)it].back() != (int)gi) member[(size_t)it].push_back((int)gi);
    std::vector<int> gc(groups.size());
    for (size_t gi = 0; gi < groups.size(); gi++) gc[gi] = group_cost(groups[gi], color, n_colors);
    Rng rng(20261020);
    double T = 0.5;
    const double cool = 1.0 - 5.0 / (double)iters;
    std::vector<int> aff, newc;
    std::vector<int> best = color;
    long cur = 0, best_cost;
    for (int c : gc) cur += c;
    best_cost = cur;
    for (long it = 0; it < iters; it++) {
        const int a = (int)(rng.next() % (uint32_t)n), b = (int)(rng.next() % (uint32_t)n);
        if (color[(size_t)a] == color[(size_t)b] || cls[(size_t)a] != cls[(size_t)b]) continue;
        aff.clear();
        for (int gi : member[(size_t)a]) aff.push_back(gi);
        for (int gi : member[(size_t)b])
            if (std::find(member[(size_t)a].begin(), member[(size_t)a].end(), gi) == member[(size_t)a].end()) aff.push_back(gi);
        int old = 0;
        for (int gi : aff) old += gc[(size_t)gi];
        std::swap(color[(size_t)a], color[(size_t)b]);
        newc.resize(aff.size());
        int neu = 0;
        for (size_t k = 0; k < aff.size(); k++) {
            newc[k] = group_cost(groups[(size_t)aff[k]], color, n_colors);
            neu += newc[k];
        }
        if (neu <= old || rng.uniform() < std::exp(-(double)(neu - old) / T)) {
            for (size_t k = 0; k < aff.size(); k++) gc[(size_t)aff[k]] = newc[k];
            cur += neu - old;
            if (cur < best_cost) {
                best_cost = cur;
                best = color;
            }
        } else {
            std::swap(color[(size_t)a], color[(size_t)b]);
        }
        T = std::max(0.01, T * cool);
    }
    color = best;
}

// Rows of a generator without padding: entry 0 is the diagonal (slot 0xFFFF when the row stores none), then the
// stored off-diagonal entries; and the order the row-split kernels deal them to lanes in (descending length,
// stable): sorted position q holds row perm[q]; chunk c = rows perm[32c .. 32c+31].
void rowsplit_rows(const Generator &g, std::vector<std::vector<uint32_t>> &compact, std::vector<int> &perm) {
    const int n = g.n;
    compact.assign((size_t)n, {});
    for (int r = 0; r < n; r++) {
        compact[(size_t)r].push_back(g.ell[(size_t)r]);
        for (int j = 1; j < g.max_nnz; j++) {
            const uint32_t e = g.ell[(size_t)j * n + r];
            if ((e >> 16) != 0xFFFFu) compact[(size_t)r].push_back(e);
        }
    }
    perm.resize((size_t)n);
    for (int r = 0; r < n; r++) perm[(size_t)r] = r;
    std::stable_sort(perm.begin(), perm.end(),
                     [&](int x, int y) { return compact[(size_t)x].size() > compact[(size_t)y].size(); });
}

// Groups of items read together by a quarter warp in the row-split gather: what = 0 the source rows of the
// off-diagonal entries plus the lanes' own rows (the stage-vector store); what = 1 the distinct value slots of the
// off-diagonal entries.
void rowsplit_groups(const Generator &g, const std::vector<std::vector<uint32_t>> &compact, const std::vector<int> &perm,
                     int what, std::vector<std::vector<int>> &groups) {
    const int n = g.n;
    groups.clear();
    for (int c0 = 0; c0 < n; c0 += 32) {
        const int longest = (int)compact[(size_t)perm[(size_t)c0]].size();
        for (int q0 = c0; q0 < std::min(n, c0 + 32); q0 += 8) {
            const int q1 = std::min(n, q0 + 8);
            for (int j = 1; j < longest; j++) {
                std::vector<int> grp;
                for (int q = q0; q < q1; q++) {
                    const auto &row = compact[(size_t)perm[(size_t)q]];
                    if (j >= (int)row.size()) continue;
                    const int item = what == 0 ? (int)(row[(size_t)j] & 0xFFFFu) : (int)(row[(size_t)j] >> 16);
                    if (what == 1 && std::find(grp.begin(), grp.end(), item) != grp.end()) continue;  // broadcast
                    grp.push_back(item);
                }
                if (grp.size() > 1) groups.push_back(std::move(grp));
            }
            if (what == 0) {
                std::vector<int> grp;
                for (int q = q0; q < q1; q++) grp.push_back(perm[(size_t)q]);
                if (grp.size() > 1) groups.push_back(std::move(grp));
            }
        }
    }
}

// Renumber the value slots inside their two ranges (time-dependent first, then static) so that the slots a quarter
// warp reads together fall into distinct 16-byte bank groups.  Only for complex value tables (16 bytes per slot).
void optimize_value_slots(Generator &g, long iters) {
    const int ns = g.n_slots;
    if (ns < 9) return;
    std::vector<std::vector<uint32_t>> compact;
    std::vector<int> perm;
    rowsplit_rows(g, compact, perm);
    std::vector<std::vector<int>> groups;
    rowsplit_groups(g, compact, perm, 1, groups);
    std::vector<int> color((size_t)ns), cls((size_t)ns);
    for (int s = 0; s < ns; s++) {
        color[(size_t)s] = s % 8;
        cls[(size_t)s] = s < g.n_dyn_slots ? 0 : 1;
    }
    anneal_colors(groups, cls, 8, iters, color);
    // colours -> positions: within each class hand out that class's positions of the wanted colour in order
    std::vector<int> new_of((size_t)ns, -1);
    std::vector<std::vector<int>> free_pos(16);
    for (int c = 0; c < 2; c++) {
        for (auto &v : free_pos) v.clear();
        const int lo = c == 0 ? 0 : g.n_dyn_slots, hi = c == 0 ? g.n_dyn_slots : ns;
        for (int p = hi - 1; p >= lo; p--) free_pos[(size_t)(p % 8)].push_back(p);
        for (int s = lo; s < hi; s++) {
            auto &v = free_pos[(size_t)color[(size_t)s]];
            new_of[(size_t)s] = v.back();
            v.pop_back();
        }
    }
    // apply: ell entries, and the slots' pair lists
    for (auto &e : g.ell) {
        const uint32_t sl = e >> 16;
        if (sl != 0xFFFFu) e = (e & 0xFFFFu) | ((uint32_t)new_of[sl] << 16);
    }
    std::vector<int> old_of((size_t)ns);
    for (int s = 0; s < ns; s++) old_of[(size_t)new_of[(size_t)s]] = s;
    std::vector<int32_t> sp, pc;
    std::vector<cplx> pw;
    sp.push_back(0);
    for (int p = 0; p < ns; p++) {
        const int s = old_of[(size_t)p];
        for (int q = g.slot_ptr[(size_t)s]; q < g.slot_ptr[(size_t)s + 1]; q++) {
            pc.push_back(g.pair_chan[(size_t)q]);
            pw.push_back(g.pair_w[(size_t)q]);
        }
        sp.push_back((int32_t)pc.size());
    }
    g.slot_ptr = sp;
    g.pair_chan = pc;
    g.pair_w = pw;
}

}  // namespace qsim
