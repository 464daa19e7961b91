// ht_astar_rules.h -- per-cell rules of the GENERAL D8 path (DEMs with ties, flats and
// depressions), shared by the sm_100a kernels (ht_astar.cu) and by the host harness the
// CPU tests compile (tests/host/astar_host.cpp: the same functions driven by plain loops,
// so the rules are checked against the oracle where no GPU is available).
//
// What is reproduced, exactly: GRASS r.watershed ram/SFD on ANY DEM
// (/root/reference/src/hydrotools/watershed.py:64-75; rules R1-R4 of oracle/rwatershed.c).
// r.watershed's A* pops the cell with the smallest (elevation, insertion counter) from a
// heap seeded with the region border / NULL-adjacent cells, and every cell drains to the
// cell that pushed it.  Let W(c) be the priority-flood level of c (the running maximum of
// the popped elevations when c pops; W = alt outside depressions).  Then:
//   * cells pop in ascending W;
//   * at one level L the LEVEL cells (alt = W = L) pop first-in-first-out: seeds first (they
//     enter the heap before the search, in row-major order), every other cell after the cell
//     that pushed it, in the pusher's neighbour scan order;
//   * a DEPRESSION cell (alt < W = L) belongs to a connected component C of such cells.  C is
//     entered when the first level-L cell next to it pops (its TRIGGER X); the whole of C
//     then pops before the next level cell does, lowest elevation first -- an "episode",
//     which touches nothing outside C and is replayed sequentially by one thread
//     (episode_run below: the very heap walk of r.watershed, restricted to the
//     components X triggers).
// The pop time of a cell is therefore the key
//     t(c) = (W, FIFO position of c's anchor at level W, position inside the episode)
// with anchor = c itself for a level cell, the trigger for a depression cell; FIFO
// positions compare through the push events (t(pusher), scan position), recursively.
// compare() walks two such chains in lock step until they differ.  The direction of a cell
// is the earliest-popped neighbour that pushes it (r.watershed's diagonal rule may make a
// neighbour skip it), so every decision needs only cells that pop EARLIER: all cells are
// resolved in parallel rounds, a cell whose comparison meets an unresolved cell waits.
// Breadth-first "generations" of the level cells that have no lower-level neighbour
// (flats, and cells reached through a depression) tell a cell which neighbours can
// precede it at all, so that waiting never dead-locks.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HT_HD __host__ __device__ __forceinline__
#define HT_D __device__ __forceinline__
#else
#define HT_HD static inline
#define HT_D static inline
#endif

namespace astar {

constexpr int32_t INVALID = 0x7FFFFFFF; // NULL cell / outside the raster
constexpr uint8_t ST_UNKNOWN = 0xFF;    // direction byte not resolved yet
constexpr uint8_t CLS_SEED = 1;         // region border or NULL 8-neighbour (always a level cell)
constexpr uint8_t CLS_INTERIOR = 2;     // level cell, not a seed, no neighbour of a lower level
constexpr uint8_t CLS_DEPR = 4;         // depression cell: alt < W
constexpr uint8_t CLS_RIM = 8;          // level cell with a depression cell of its own level next to it
constexpr int32_t GEN_UNSET = 0x7FFFFFFF;
constexpr int32_t NONE = -1;

// packed direction byte (include/hydrotools_b200.h)
constexpr uint8_t F_CODE = 0x0F, F_NEG = 0x10, F_EDGE = 0x20, F_TAINT = 0x40, F_NULL = 0x80;

// r.watershed's neighbour scan order ct = 0..7: S, N, W, E, NE, SW, SE, NW
HT_HD int ct_dr(int ct) { return (int)((0x2852u >> (2 * ct)) & 3u) - 1; }  // {1,-1,0,0,-1,1,1,-1}
HT_HD int ct_dc(int ct) { return (int)((0x2285u >> (2 * ct)) & 3u) - 1; }  // {0,0,-1,1,1,-1,1,-1}
// GRASS code of the neighbour in scan position ct: {6, 2, 4, 8, 1, 5, 7, 3}
HT_HD int ct2code(int ct) { return (int)((0x37518426u >> (4 * ct)) & 0xFu); }
// scan position of the neighbour GRASS code 1..8 points at: code -> ct
HT_HD int code2ct(int code) { return (int)((0x360527140ull >> (4 * code)) & 0xFull); }
// row / column step of GRASS code 1..8 (0 for code 0)
HT_HD int code_dr(int code) { return (int)((0x1A901u >> (2 * code)) & 3u) - 1; }
HT_HD int code_dc(int code) { return (int)((0x29019u >> (2 * code)) & 3u) - 1; }
// cardinal flanks of a diagonal scan position (r.watershed's nbr_ew / nbr_ns)
HT_HD int flank_ew(int ct) { return (int)((0x1001u >> (4 * (ct - 4))) & 0xFu); } // {1,0,0,1}
HT_HD int flank_ns(int ct) { return (int)((0x2323u >> (4 * (ct - 4))) & 0xFu); } // {3,2,3,2}

// volatile / fenced accessors: inside one launch a cell may read what another thread has
// just published (results only ever go from "unknown" to their final value)
#ifdef __CUDA_ARCH__
HT_D uint8_t ld_u8(const uint8_t *p) { return *reinterpret_cast<const volatile uint8_t *>(p); }
HT_D int32_t ld_i32(const int32_t *p) { return *reinterpret_cast<const volatile int32_t *>(p); }
HT_D void fence() { __threadfence(); }
#else
HT_D uint8_t ld_u8(const uint8_t *p) { return *p; }
HT_D int32_t ld_i32(const int32_t *p) { return *p; }
HT_D void fence() {}
#endif

struct Grid {
    const int32_t *alt;   // GRASS integer millimetres; INVALID = NULL
    const int32_t *W;     // priority-flood level
    const uint8_t *cls;   // CLS_* flags
    uint8_t *st;          // direction bytes (ST_UNKNOWN = pending)
    const int32_t *gen;   // level cells: generation; component representative: the component's
    const int32_t *lab;   // depression cells: representative (smallest cell index) of the component
    int32_t *trig;        // at the representative: trigger candidate (NONE = no vote yet)
    int32_t *votes;       // at the representative: votes received
    const int32_t *quorum; // at the representative: votes to wait for
    int32_t *seq;         // depression cells: pop position inside their episode
    int64_t rows, cols;
    double ew, ns, diag;
};

HT_HD int64_t nbr_of(const Grid &g, int64_t i, int ct, bool &inside)
{
    const int64_t r = i / g.cols, c = i - r * g.cols;
    const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
    inside = rr >= 0 && rr < g.rows && cc >= 0 && cc < g.cols;
    return rr * g.cols + cc;
}

// the trigger of the component of depression cell d, or NONE while the vote is still open
HT_D int32_t trigger_of(const Grid &g, int64_t d)
{
    const int32_t rep = g.lab[d];
    if (ld_i32(g.votes + rep) != g.quorum[rep])
        return NONE;
    fence();
    return ld_i32(g.trig + rep);
}

// pop order of two distinct valid cells: -1 = a first, +1 = b first, 0 = not known yet
HT_D int compare(const Grid &g, int64_t a, int64_t b)
{
    for (int64_t guard = 0; guard < g.rows * g.cols + 8; guard++) {
        const int32_t Wa = g.W[a], Wb = g.W[b];
        if (Wa != Wb)
            return Wa < Wb ? -1 : 1;
        // anchors: the level cell whose FIFO position dates the pop
        int64_t xa = a, xb = b;
        int32_t sa = 0, sb = 0;
        if (g.cls[a] & CLS_DEPR) {
            if (ld_u8(g.st + a) == ST_UNKNOWN)
                return 0; // episode not replayed yet
            fence();
            sa = 1 + g.seq[a];
            xa = trigger_of(g, a);
        }
        if (g.cls[b] & CLS_DEPR) {
            if (ld_u8(g.st + b) == ST_UNKNOWN)
                return 0;
            fence();
            sb = 1 + g.seq[b];
            xb = trigger_of(g, b);
        }
        if (xa == xb)
            return sa < sb ? -1 : 1;
        const bool seed_a = g.cls[xa] & CLS_SEED, seed_b = g.cls[xb] & CLS_SEED;
        if (seed_a || seed_b) {
            if (seed_a && seed_b)
                return xa < xb ? -1 : 1; // seeds enter the heap in row-major order
            return seed_a ? -1 : 1;      // ... and before every cell pushed later
        }
        const uint8_t ta = ld_u8(g.st + xa), tb = ld_u8(g.st + xb);
        if (ta == ST_UNKNOWN || tb == ST_UNKNOWN)
            return 0;
        const int ca = ta & F_CODE, cb = tb & F_CODE;
        const int64_t pa = xa + (int64_t)code_dr(ca) * g.cols + code_dc(ca);
        const int64_t pb = xb + (int64_t)code_dr(cb) * g.cols + code_dc(cb);
        if (pa == pb) {
            // pushed by the same cell: its scan order decides (the pusher sees a cell in the
            // position opposite to the one the cell sees the pusher in: ct ^ 1)
            return (code2ct(ca) ^ 1) < (code2ct(cb) ^ 1) ? -1 : 1;
        }
        a = pa;
        b = pb;
    }
    return 0;
}

// does the popping cell D (diagonal neighbour of U in U's scan position ct) skip U?
// r.watershed: a cardinal flank E of D next to U that is still unworked, higher than D and
// from which U is steeper.  1 = skips, 0 = pushes, -1 = not known yet.
HT_D int skips(const Grid &g, int64_t u, int ct, const int64_t *in, const int32_t *an)
{
    const int32_t au = g.alt[u], ad = an[ct];
    if (ct < 4 || au <= ad)
        return 0;
    const double sd = (double)(au - ad) / g.diag;
    // U's cardinal neighbours on D's side are D's flanks: (0, dc) an ew step from U, (dr, 0) an ns step
    const int fl[2] = {ct_dc(ct) < 0 ? 2 : 3, ct_dr(ct) > 0 ? 0 : 1};
    const double dist[2] = {g.ew, g.ns};
    for (int k = 0; k < 2; k++) {
        const int e = fl[k];
        if (an[e] == INVALID || an[e] <= ad)
            continue; // slope[e] = 0: NULL / outside / not higher than D
        if (au <= an[e] || !(sd < (double)(au - an[e]) / dist[k]))
            continue; // no steeper way through E: no need to know whether E is worked
        const int cmp = compare(g, in[e], in[ct]);
        if (cmp == 0)
            return -1;
        if (cmp > 0)
            return 1; // E pops after D: unworked when D pops
    }
    return 0;
}

// direction byte of LEVEL cell i (seed or not), or ST_UNKNOWN if a comparison has to wait
HT_D uint8_t resolve_level_cell(const Grid &g, int64_t i)
{
    const int64_t r = i / g.cols, c = i - r * g.cols;
    const int32_t L = g.W[i];
    int32_t an[8], wn[8];
    int64_t in[8];
    for (int ct = 0; ct < 8; ct++) {
        bool inside;
        in[ct] = nbr_of(g, i, ct, inside);
        an[ct] = inside ? g.alt[in[ct]] : INVALID;
        wn[ct] = inside && an[ct] != INVALID ? g.W[in[ct]] : INVALID;
        if (!inside)
            in[ct] = -1;
    }
    const uint8_t cl = g.cls[i];
    int best = -1;
    bool wait = false;
    auto offer = [&](int ct) {
        if (best < 0) {
            best = ct;
            return;
        }
        const int cmp = compare(g, in[ct], in[best]);
        if (cmp == 0)
            wait = true;
        else if (cmp < 0)
            best = ct;
    };
    if (cl & CLS_SEED) {
        // rule R2 / R4: a seed is re-pointed by the first neighbour that pops before it
        for (int ct = 0; ct < 8 && !wait; ct++) {
            if (an[ct] == INVALID)
                continue;
            bool before = wn[ct] < L;
            if (wn[ct] == L) {
                const uint8_t cn = g.cls[in[ct]];
                if (cn & CLS_DEPR) {
                    // its episode runs right after its trigger: before this seed only if the
                    // trigger is an earlier seed
                    const int32_t x = trigger_of(g, in[ct]);
                    if (x == NONE) {
                        wait = true;
                        break;
                    }
                    before = (g.cls[x] & CLS_SEED) && x < i;
                    if (before && ld_u8(g.st + in[ct]) == ST_UNKNOWN)
                        wait = true;
                }
                else
                    before = (cn & CLS_SEED) && in[ct] < i;
            }
            if (before && !wait)
                offer(ct);
        }
        if (wait)
            return ST_UNKNOWN;
        if (best >= 0)
            return (uint8_t)(ct2code(best) | F_EDGE);
        int code;
        if (r == 0)
            code = 2;
        else if (c == 0)
            code = 4;
        else if (r == g.rows - 1)
            code = 6;
        else if (c == g.cols - 1)
            code = 8;
        else {
            code = 0;
            for (int ct = 7; ct >= 0; ct--) // first NULL neighbour in scan order
                if (an[ct] == INVALID)
                    code = ct2code(ct);
        }
        return (uint8_t)(code | F_EDGE | F_NEG);
    }
    if (cl & CLS_INTERIOR) {
        // generation k >= 1: pushed by a cell of its own level one generation earlier -- a
        // level cell of equal elevation, or a cell of a depression whose trigger is one
        const int32_t k = g.gen[i];
        if (k == GEN_UNSET || k <= 0)
            return ST_UNKNOWN;
        for (int ct = 0; ct < 8 && !wait; ct++) {
            if (wn[ct] != L)
                continue;
            const uint8_t cn = g.cls[in[ct]];
            if (cn & CLS_DEPR) {
                if (g.gen[g.lab[in[ct]]] != k - 1)
                    continue;
                if (ld_u8(g.st + in[ct]) == ST_UNKNOWN) {
                    wait = true;
                    break;
                }
                const int sk = skips(g, i, ct, in, an);
                if (sk < 0)
                    wait = true;
                else if (sk == 0)
                    offer(ct);
            }
            else if (g.gen[in[ct]] == k - 1)
                offer(ct);
        }
        if (wait || best < 0)
            return ST_UNKNOWN;
        return (uint8_t)ct2code(best);
    }
    // generation 0: the earliest neighbour of a lower level that does not skip it
    for (int ct = 0; ct < 8 && !wait; ct++) {
        if (wn[ct] >= L) // INVALID is larger than every level
            continue;
        const int sk = skips(g, i, ct, in, an);
        if (sk < 0)
            wait = true;
        else if (sk == 0)
            offer(ct);
    }
    if (wait || best < 0)
        return ST_UNKNOWN;
    return (uint8_t)ct2code(best);
}

// ---------------------------------------------------------------------------
// Episode: the walk of r.watershed's heap through the depression components that level
// cell x triggers.  Sequential by nature (lowest elevation first, first-in-first-out among
// equals); independent of everything outside the components, so one thread replays it.
// heap: caller-provided scratch of `cap` entries (cap >= cells of the components).
// flags: per cell, bit 0 in the list, bit 1 worked (zero before the first episode).
// ---------------------------------------------------------------------------
struct HeapItem {
    int32_t alt, added, cell;
};

HT_HD bool heap_less(const HeapItem &a, const HeapItem &b)
{
    return a.alt != b.alt ? a.alt < b.alt : a.added < b.added;
}

// 8-ary heap: a sift-down level costs one round of independent loads (the eight children
// are contiguous) instead of two dependent ones per binary level -- the episode is one
// thread chasing global-memory latency
constexpr int HEAP_ARITY = 8;

HT_D void heap_push(HeapItem *h, int32_t &n, HeapItem v)
{
    int32_t i = n++;
    while (i > 0) {
        const int32_t p = (i - 1) / HEAP_ARITY;
        const HeapItem hp = h[p];
        if (!heap_less(v, hp))
            break;
        h[i] = hp;
        i = p;
    }
    h[i] = v;
}

HT_D HeapItem heap_pop(HeapItem *h, int32_t &n)
{
    const HeapItem top = h[0];
    const HeapItem v = h[--n];
    int32_t i = 0;
    for (;;) {
        const int32_t first = HEAP_ARITY * i + 1;
        if (first >= n)
            break;
        const int32_t last = first + HEAP_ARITY < n ? first + HEAP_ARITY : n;
        int32_t m = first;
        HeapItem hm = h[first];
        for (int32_t c = first + 1; c < last; c++) {
            const HeapItem hc = h[c];
            if (heap_less(hc, hm)) {
                hm = hc;
                m = c;
            }
        }
        if (!heap_less(hm, v))
            break;
        h[i] = hm;
        i = m;
    }
    if (n > 0)
        h[i] = v;
    return top;
}

// returns the number of cells popped, or -1 if the scratch is too small
HT_D int32_t episode_run(const Grid &g, int64_t x, HeapItem *heap, int32_t cap, uint8_t *flags)
{
    const int32_t L = g.W[x];
    int32_t n = 0, counter = 0, popped = 0;
    // x pops: its depression neighbours that it triggers enter the heap in scan order (they are
    // lower than x, so nothing is skipped)
    for (int ct = 0; ct < 8; ct++) {
        bool inside;
        const int64_t u = nbr_of(g, x, ct, inside);
        if (!inside || g.alt[u] == INVALID || !(g.cls[u] & CLS_DEPR) || g.W[u] != L)
            continue;
        if (g.trig[g.lab[u]] != (int32_t)x || (flags[u] & 1))
            continue;
        if (n >= cap)
            return -1;
        flags[u] |= 1;
        g.st[u] = ST_UNKNOWN; // published after the episode, direction kept in seq[] meanwhile
        g.seq[u] = -(ct2code(ct ^ 1)) - 1;
        heap_push(heap, n, HeapItem{g.alt[u], counter++, (int32_t)u});
    }
    const double dist[8] = {g.ns, g.ns, g.ew, g.ew, g.diag, g.diag, g.diag, g.diag};
    while (n > 0) {
        const HeapItem top = heap_pop(heap, n);
        const int64_t d = top.cell;
        const int32_t ad = top.alt;
        flags[d] |= 2;
        double slope[8];
        int32_t alt_nbr[8];
        for (int ct = 0; ct < 8; ct++) {
            slope[ct] = 0.0;
            alt_nbr[ct] = 0;
            bool inside;
            const int64_t u = nbr_of(g, d, ct, inside);
            if (!inside)
                continue;
            const int32_t au = g.alt[u];
            const bool member = au != INVALID && (g.cls[u] & CLS_DEPR); // same component as d
            // worked: NULL cells, the trigger, members already popped; every other neighbour
            // of a depression cell pops after the episode
            const bool worked = au == INVALID || u == x || (member && (flags[u] & 2));
            if (!worked) {
                alt_nbr[ct] = au;
                slope[ct] = au > ad ? (double)(au - ad) / dist[ct] : 0.0;
            }
            if (!member || (flags[u] & 1))
                continue; // cells outside the component are pushed by the general rule
            bool skip = false;
            if (ct > 3 && slope[ct] > 0.0) {
                const int e = flank_ew(ct), s = flank_ns(ct);
                if (slope[e] > 0.0 && alt_nbr[ct] > alt_nbr[e] &&
                    slope[ct] < (double)(alt_nbr[ct] - alt_nbr[e]) / g.ew)
                    skip = true;
                if (!skip && slope[s] > 0.0 && alt_nbr[ct] > alt_nbr[s] &&
                    slope[ct] < (double)(alt_nbr[ct] - alt_nbr[s]) / g.ns)
                    skip = true;
            }
            if (skip)
                continue;
            if (n >= cap)
                return -1;
            flags[u] |= 1;
            g.seq[u] = -(ct2code(ct ^ 1)) - 1; // u drains to d
            heap_push(heap, n, HeapItem{au, counter++, (int32_t)u});
        }
        // pop position; the direction travelled in seq[] until now
        const int code = -(g.seq[d] + 1);
        g.seq[d] = popped++;
        fence();
        g.st[d] = (uint8_t)code;
    }
    return popped;
}

} // namespace astar
