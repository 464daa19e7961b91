// tests/host/astar_host.cpp -- TEST INFRASTRUCTURE.
// Drives the per-cell rules of the general D8 path (hydro-tools_b200/csrc/ht_astar_rules.h,
// the very functions the sm_100a kernels call) with plain loops, so that the CPU test suite
// can compare them with the oracle where there is no GPU.  Not part of the product: the
// product path is ht_astar.cu and fails loudly without a device.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../hydro-tools_b200/csrc/ht_astar_rules.h"

using namespace astar;

// alt / W: GRASS integer millimetres and priority-flood level (INVALID = NULL), row-major.
// st: out, packed direction bytes.  stats[0] = rounds, [1] = episodes, [2] = largest episode,
// [3] = depression cells, [4] = largest generation.
extern "C" int astar_host_run(const int32_t *alt, const int32_t *W, int64_t rows, int64_t cols,
                              double ew, double ns, double diag, uint8_t *st, int64_t *stats)
{
    const int64_t n = rows * cols;
    std::vector<uint8_t> cls(n, 0), flags(n, 0), voted(n, 0), epi_done(n, 0);
    std::vector<int32_t> gen(n, GEN_UNSET), lab(n, NONE), trig(n, NONE), votes(n, 0), quorum(n, 0),
        seq(n, 0), size(n, 0);
    auto nb = [&](int64_t i, int ct) -> int64_t {
        const int64_t r = i / cols + ct_dr(ct), c = i % cols + ct_dc(ct);
        return (r < 0 || r >= rows || c < 0 || c >= cols) ? -1 : r * cols + c;
    };
    // ---- classes
    for (int64_t i = 0; i < n; i++) {
        if (alt[i] == INVALID)
            continue;
        const int64_t r = i / cols, c = i % cols;
        bool seed = r == 0 || c == 0 || r == rows - 1 || c == cols - 1, lower = false, rim = false;
        for (int ct = 0; ct < 8; ct++) {
            const int64_t j = nb(i, ct);
            if (j < 0)
                continue;
            if (alt[j] == INVALID) {
                seed = true;
                continue;
            }
            lower |= W[j] < W[i];
            rim |= W[j] == W[i] && alt[j] < W[j];
        }
        if (alt[i] < W[i])
            cls[i] = CLS_DEPR;
        else
            cls[i] = (seed ? CLS_SEED : (lower ? 0 : CLS_INTERIOR)) | (rim ? CLS_RIM : 0);
    }
    // ---- depression components (8-connected), representative = smallest index
    int64_t ndep = 0;
    for (int64_t i = 0; i < n; i++) {
        if (!(cls[i] & CLS_DEPR) || lab[i] != NONE)
            continue;
        std::vector<int64_t> stack{i};
        lab[i] = (int32_t)i;
        while (!stack.empty()) {
            const int64_t u = stack.back();
            stack.pop_back();
            size[i]++;
            ndep++;
            for (int ct = 0; ct < 8; ct++) {
                const int64_t j = nb(u, ct);
                if (j >= 0 && (cls[j] & CLS_DEPR) && lab[j] == NONE) {
                    lab[j] = (int32_t)i;
                    stack.push_back(j);
                }
            }
        }
    }
    // ---- generations: level cells with a lower-level neighbour and seeds are generation 0
    for (int64_t i = 0; i < n; i++)
        if (alt[i] != INVALID && !(cls[i] & (CLS_DEPR | CLS_INTERIOR)))
            gen[i] = 0;
    int32_t maxgen = 0;
    for (bool changed = true; changed;) {
        changed = false;
        for (int64_t i = 0; i < n; i++) {
            if (cls[i] & CLS_DEPR) { // component: the earliest generation among its level-L rim
                for (int ct = 0; ct < 8; ct++) {
                    const int64_t j = nb(i, ct);
                    if (j < 0 || alt[j] == INVALID || (cls[j] & CLS_DEPR) || W[j] != W[i])
                        continue;
                    if (gen[j] < gen[lab[i]]) {
                        gen[lab[i]] = gen[j];
                        changed = true;
                    }
                }
            }
            else if (cls[i] & CLS_INTERIOR) {
                int32_t best = gen[i];
                for (int ct = 0; ct < 8; ct++) {
                    const int64_t j = nb(i, ct);
                    if (j < 0 || alt[j] == INVALID || W[j] != W[i])
                        continue;
                    const int32_t gj = (cls[j] & CLS_DEPR) ? gen[lab[j]] : gen[j];
                    if (gj != GEN_UNSET && gj + 1 < best)
                        best = gj + 1;
                }
                if (best < gen[i]) {
                    gen[i] = best;
                    changed = true;
                    maxgen = std::max(maxgen, best);
                }
            }
        }
    }
    // ---- quorum: level cells next to a component, of the component's generation
    auto adjacent_reps = [&](int64_t u, int32_t *reps) -> int {
        int k = 0;
        for (int ct = 0; ct < 8; ct++) {
            const int64_t j = nb(u, ct);
            if (j < 0 || !(cls[j] & CLS_DEPR) || W[j] != W[u] || gen[lab[j]] != gen[u])
                continue;
            bool seen = false;
            for (int q = 0; q < k; q++)
                seen |= reps[q] == lab[j];
            if (!seen)
                reps[k++] = lab[j];
        }
        return k;
    };
    for (int64_t i = 0; i < n; i++)
        if (cls[i] & CLS_RIM) {
            int32_t reps[8];
            const int k = adjacent_reps(i, reps);
            for (int q = 0; q < k; q++)
                quorum[reps[q]]++;
        }
    memset(st, ST_UNKNOWN, n);
    for (int64_t i = 0; i < n; i++)
        if (alt[i] == INVALID)
            st[i] = F_NULL;
    Grid g{alt, W, cls.data(), st, gen.data(), lab.data(), trig.data(), votes.data(), quorum.data(),
           seq.data(), rows, cols, ew, ns, diag};
    std::vector<int64_t> act;
    for (int64_t i = 0; i < n; i++)
        if (alt[i] != INVALID && !(cls[i] & CLS_DEPR))
            act.push_back(i);
    std::vector<HeapItem> heap;
    int64_t rounds = 0, episodes = 0, biggest = 0;
    while (!act.empty()) {
        rounds++;
        bool progress = false;
        std::vector<int64_t> rest;
        for (int64_t i : act) {
            // 1. direction
            if (st[i] == ST_UNKNOWN) {
                const uint8_t b = resolve_level_cell(g, i);
                if (b != ST_UNKNOWN) {
                    st[i] = b;
                    progress = true;
                }
            }
            bool pending = st[i] == ST_UNKNOWN;
            if (cls[i] & CLS_RIM) {
                int32_t reps[8];
                const int k = adjacent_reps(i, reps);
                // 2. votes: a rim cell runs for trigger of every component of its generation
                // (its FIFO position is known once its pusher is; a seed's always)
                if (voted[i] < k && ((cls[i] & CLS_SEED) || st[i] != ST_UNKNOWN)) {
                    while (voted[i] < k) {
                        const int32_t rep = reps[voted[i]];
                        const int32_t cur = trig[rep];
                        int cmp = cur == NONE ? -1 : compare(g, i, cur);
                        if (cmp == 0)
                            break;
                        if (cmp < 0)
                            trig[rep] = (int32_t)i;
                        votes[rep]++;
                        voted[i]++;
                        progress = true;
                    }
                }
                pending |= voted[i] < k;
                // 3. episode: once every adjacent vote is closed
                if (voted[i] == k && !epi_done[i]) {
                    bool closed = true, mine = false;
                    int64_t cap = 0;
                    for (int q = 0; q < k; q++) {
                        closed &= votes[reps[q]] == quorum[reps[q]];
                        if (trig[reps[q]] == (int32_t)i) {
                            mine = true;
                            cap += size[reps[q]];
                        }
                    }
                    if (closed) {
                        if (mine) {
                            heap.resize(cap);
                            const int32_t got = episode_run(g, i, heap.data(), (int32_t)cap, flags.data());
                            if (got != cap)
                                return -5;
                            episodes++;
                            biggest = std::max<int64_t>(biggest, got);
                        }
                        epi_done[i] = 1;
                        progress = true;
                    }
                }
                pending |= !epi_done[i];
            }
            if (pending)
                rest.push_back(i);
        }
        if (!progress)
            return -4; // the rules dead-locked
        act.swap(rest);
    }
    for (int64_t i = 0; i < n; i++)
        if (st[i] == ST_UNKNOWN)
            return -6; // a depression cell no episode reached
    if (stats) {
        stats[0] = rounds; stats[1] = episodes; stats[2] = biggest; stats[3] = ndep; stats[4] = maxgen;
    }
    return 0;
}
