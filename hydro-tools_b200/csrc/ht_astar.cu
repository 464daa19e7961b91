// ht_astar.cu -- GENERAL D8 direction path: GRASS r.watershed -s on ANY DEM (equal
// elevations, flats, depressions), bit-exact, in parallel on the device.
//
// Replaces the direction half of /root/reference/src/hydrotools/watershed.py:64-75 for the
// DEMs the local-rule kernel (ht_d8.cu) refuses -- which is every raw DEM, the reference's
// own test fixture included (tests/data-files/dem.tif: 2 326 pits, 25 % of the cells inside
// depressions).  r.watershed is a sequential least-cost search; ht_astar_rules.h states
// what fixes each cell's direction in terms of cells that pop earlier, so that the search
// becomes rounds over all pending cells:
//   K1  alt_kernel        float metres -> GRASS integer millimetres (rule R1)
//   K2  fill_*            priority-flood level W by tile-wise relaxation in shared memory
//                         (W = max(alt, min W of the neighbours), seeds fixed), repeated
//                         over the tiles whose ring changed until nothing moves
//   K3  classify_kernel   seeds, level / depression cells, flat interiors, rims
//   K4  ccl_*             depression components (label = smallest cell index)
//   K5  generation_kernel breadth-first generations through flats and components
//   K6  quorum_kernel     how many rim cells run for trigger of each component
//   K7  round_kernel      every pending level cell: direction (resolve_level_cell), vote for
//                         trigger (atomicCAS tournament), and -- for the cell that wins a
//                         component -- the sequential replay of its episode
//   K8  finish_kernel     taint flags (rule R4) and the packed direction byte
// The accumulation stages (ht_acc.cu) then run on the packed grid unchanged.
// Host loops (fill passes, label / generation sweeps, rounds) read one counter per
// iteration; the entry point is synchronous.
#include "ht_common.cuh"
#include "ht_astar_rules.h"

namespace {

using namespace astar;

constexpr int FT = 64;          // fill tile edge
constexpr int FR = FT + 2;      // with ring
constexpr int32_t BIG = 0x7FFFFFFE; // not reached yet (valid, not a seed)

struct Work {
    int32_t *alt, *W, *gen, *lab, *trig, *votes, *quorum, *seq, *size;
    uint8_t *cls, *st, *flags, *voted;
    int32_t *listA, *listB;  // cell lists (ping-pong)
    uint8_t *dirty;          // [2][tiles] fill tiles to revisit
    HeapItem *heap;          // episode scratch, one slice per episode
    unsigned long long *ctr; // [16] device counters
    int64_t n;
};

enum { C_RANGE = 0, C_CHANGED, C_NLIST, C_NDEP, C_NSPECIAL, C_PROGRESS, C_HEAP, C_EPISODES, C_BIGGEST, C_ERR };

size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

size_t work_bytes(int64_t rows, int64_t cols)
{
    const size_t n = (size_t)rows * cols;
    const size_t tiles = (size_t)ht::cdiv(rows, FT) * ht::cdiv(cols, FT);
    return 11 * align_up(n * 4) + 4 * align_up(n) + align_up(2 * tiles) + align_up(n * sizeof(HeapItem)) +
           align_up(16 * 8) + 512;
}

int carve(void *ws, size_t ws_bytes, int64_t rows, int64_t cols, Work &w)
{
    if (!ws || ws_bytes < work_bytes(rows, cols))
        return HT_ERR_WORKSPACE;
    const size_t n = (size_t)rows * cols;
    const size_t tiles = (size_t)ht::cdiv(rows, FT) * ht::cdiv(cols, FT);
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    int32_t **i32s[] = {&w.alt, &w.W, &w.gen, &w.lab, &w.trig, &w.votes, &w.quorum, &w.seq, &w.size,
                        &w.listA, &w.listB};
    for (auto p : i32s) {
        *p = (int32_t *)b;
        b += align_up(n * 4);
    }
    uint8_t **u8s[] = {&w.cls, &w.st, &w.flags, &w.voted};
    for (auto p : u8s) {
        *p = b;
        b += align_up(n);
    }
    w.dirty = b; b += align_up(2 * tiles);
    w.heap = (HeapItem *)b; b += align_up(n * sizeof(HeapItem));
    w.ctr = (unsigned long long *)b;
    w.n = (int64_t)n;
    return 0;
}

// ---------------------------------------------------------------- K1
__global__ void alt_kernel(const float *dem, int64_t ld, int64_t rows, int64_t cols, int has_nodata,
                           float nodata, int32_t *alt, unsigned long long *ctr)
{
    const int64_t n = rows * cols;
    unsigned range = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const float z = dem[r * ld + c];
        int32_t a = INVALID;
        if (!((has_nodata && z == nodata) || z != z)) {
            if (fabsf(z) < 16777.2f) {
                // rule R1: (int)(z * 1000 +- 0.5) in double; the half takes z's sign
                const double d = (double)z * 1000.0;
                a = (int32_t)(d > 0 ? d + 0.5 : d - 0.5);
            }
            else { // out of the supported range: clamped and counted (the caller refuses)
                a = z > 0 ? (1 << 24) - 1 : 1 - (1 << 24);
                range++;
            }
        }
        alt[i] = a;
    }
    if (range)
        atomicAdd(&ctr[C_RANGE], (unsigned long long)range);
}

// ---------------------------------------------------------------- K2
__global__ void fill_init_kernel(const int32_t *alt, int64_t rows, int64_t cols, int32_t *W)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = alt[i];
        if (a == INVALID) {
            W[i] = INVALID;
            continue;
        }
        const int64_t r = i / cols, c = i - r * cols;
        bool seed = r == 0 || c == 0 || r == rows - 1 || c == cols - 1;
        if (!seed)
            for (int ct = 0; ct < 8; ct++)
                seed |= alt[(r + ct_dr(ct)) * cols + c + ct_dc(ct)] == INVALID;
        W[i] = seed ? a : BIG;
    }
}

// one tile: relax in shared memory until nothing moves; W only ever decreases
__global__ void __launch_bounds__(256)
fill_tile_kernel(const int32_t *alt, int32_t *W, int64_t rows, int64_t cols, int tiles_x, int tiles_y,
                 const uint8_t *dirty_in, uint8_t *dirty_out, unsigned long long *ctr)
{
    __shared__ int32_t sW[FR * FR];
    __shared__ int32_t sA[FT * FT];
    const int tile = blockIdx.x;
    if (!dirty_in[tile])
        return;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int64_t x0 = (int64_t)tx * FT, y0 = (int64_t)ty * FT;
    const int tid = threadIdx.x;
    for (int k = tid; k < FR * FR; k += 256) {
        const int yy = k / FR, xx = k - yy * FR;
        const int64_t r = y0 - 1 + yy, c = x0 - 1 + xx;
        sW[k] = (r >= 0 && r < rows && c >= 0 && c < cols) ? W[r * cols + c] : INVALID;
    }
    for (int k = tid; k < FT * FT; k += 256) {
        const int yy = k >> 6, xx = k & 63;
        const int64_t r = y0 + yy, c = x0 + xx;
        sA[k] = (r < rows && c < cols) ? alt[r * cols + c] : INVALID;
    }
    __syncthreads();
    bool any = false;
    for (;;) {
        bool changed = false;
#pragma unroll 4
        for (int k = tid; k < FT * FT; k += 256) {
            const int yy = k >> 6, xx = k & 63;
            const int32_t a = sA[k];
            const int p = (yy + 1) * FR + xx + 1;
            const int32_t cur = sW[p];
            if (a == INVALID || cur == a)
                continue; // NULL, or already at its floor (seeds start there)
            int32_t m = min(min(min(sW[p - FR - 1], sW[p - FR]), min(sW[p - FR + 1], sW[p - 1])),
                            min(min(sW[p + 1], sW[p + FR - 1]), min(sW[p + FR], sW[p + FR + 1])));
            const int32_t nw = max(a, m);
            if (nw < cur) {
                sW[p] = nw;
                changed = true;
            }
        }
        if (!__syncthreads_or(changed))
            break;
        any = true;
    }
    if (!any)
        return;
    for (int k = tid; k < FT * FT; k += 256) {
        const int yy = k >> 6, xx = k & 63;
        const int64_t r = y0 + yy, c = x0 + xx;
        if (r < rows && c < cols)
            W[r * cols + c] = sW[(yy + 1) * FR + xx + 1];
    }
    if (tid < 9) { // the neighbours' rings changed
        const int nx = tx + tid % 3 - 1, ny = ty + tid / 3 - 1;
        if (tid != 4 && nx >= 0 && nx < tiles_x && ny >= 0 && ny < tiles_y)
            dirty_out[ny * tiles_x + nx] = 1;
    }
    if (tid == 0)
        atomicAdd(&ctr[C_CHANGED], 1ull);
}

// ---------------------------------------------------------------- K3
__global__ void classify_kernel(Work w, int64_t rows, int64_t cols)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t a = w.alt[i];
        uint8_t cl = 0, st = ST_UNKNOWN;
        int32_t gen = GEN_UNSET, lab = NONE;
        if (a == INVALID)
            st = F_NULL;
        else {
            const int64_t r = i / cols, c = i - r * cols;
            const int32_t L = w.W[i];
            bool seed = r == 0 || c == 0 || r == rows - 1 || c == cols - 1, lower = false, rim = false;
            for (int ct = 0; ct < 8; ct++) {
                const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                    continue;
                const int64_t j = rr * cols + cc;
                const int32_t aj = w.alt[j];
                if (aj == INVALID) {
                    seed = true;
                    continue;
                }
                const int32_t wj = w.W[j];
                lower |= wj < L;
                rim |= wj == L && aj < wj;
            }
            if (a < L) {
                cl = CLS_DEPR;
                lab = (int32_t)i;
            }
            else {
                cl = (seed ? CLS_SEED : (lower ? 0 : CLS_INTERIOR)) | (rim ? CLS_RIM : 0);
                if (!(cl & CLS_INTERIOR))
                    gen = 0;
            }
        }
        w.cls[i] = cl;
        w.st[i] = st;
        w.gen[i] = gen;
        w.lab[i] = lab;
        w.trig[i] = NONE;
        w.votes[i] = 0;
        w.quorum[i] = 0;
        w.size[i] = 0;
        w.seq[i] = 0;
        w.flags[i] = 0;
        w.voted[i] = 0;
    }
}

// append the cells whose class has one of the `mask` bits (or, mask = 0, every valid level cell)
__global__ void collect_kernel(Work w, uint8_t mask, int32_t *list, unsigned long long *count)
{
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~31ll; i0 < w.n;
         i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + lane;
        bool in = false;
        if (i < w.n && w.alt[i] != INVALID) {
            const uint8_t cl = w.cls[i];
            in = mask ? (cl & mask) != 0 : !(cl & CLS_DEPR);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (!m)
            continue;
        unsigned long long at = 0;
        if (lane == 0)
            at = atomicAdd(count, (unsigned long long)__popc(m));
        at = __shfl_sync(0xffffffffu, at, 0) + __popc(m & ((1u << lane) - 1u));
        if (in)
            list[at] = (int32_t)i;
    }
}

// ---------------------------------------------------------------- K4: components
__device__ __forceinline__ int32_t find_root(const int32_t *lab, int32_t x)
{
    for (;;) {
        const int32_t p = ld_i32(lab + x);
        if (p == x)
            return x;
        x = p;
    }
}

// hook: the smaller root adopts the larger (labels only ever decrease; roots are members)
__global__ void ccl_hook_kernel(Work w, const int32_t *list, int64_t nlist, int64_t cols, int64_t rows)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nlist;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = list[k];
        if (!(w.cls[i] & CLS_DEPR))
            continue;
        const int64_t r = i / cols, c = i - r * cols;
        int32_t ri = find_root(w.lab, i);
        bool changed = false;
        for (int ct = 0; ct < 8; ct++) {
            const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
            if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                continue;
            const int64_t j = rr * cols + cc;
            if (w.alt[j] == INVALID || !(w.cls[j] & CLS_DEPR))
                continue;
            int32_t rj = find_root(w.lab, (int32_t)j);
            while (ri != rj) { // union by smaller index
                const int32_t hi = max(ri, rj), lo = min(ri, rj);
                const int32_t old = atomicMin(w.lab + hi, lo);
                if (old == hi) {
                    changed = true;
                    ri = rj = lo;
                }
                else { // somebody else re-rooted `hi`: chase
                    ri = find_root(w.lab, lo);
                    rj = find_root(w.lab, old);
                }
            }
        }
        if (changed)
            atomicAdd(&w.ctr[C_CHANGED], 1ull);
    }
}

__global__ void ccl_flatten_kernel(Work w, const int32_t *list, int64_t nlist)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nlist;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = list[k];
        if (!(w.cls[i] & CLS_DEPR))
            continue;
        const int32_t root = find_root(w.lab, i);
        w.lab[i] = root;
        atomicAdd(w.size + root, 1);
    }
}

// ---------------------------------------------------------------- K5: generations
__global__ void generation_kernel(Work w, const int32_t *list, int64_t nlist, int64_t rows, int64_t cols)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nlist;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = list[k];
        const uint8_t cl = w.cls[i];
        const int64_t r = i / cols, c = i - r * cols;
        const int32_t L = w.W[i];
        bool changed = false;
        if (cl & CLS_DEPR) { // the component: earliest generation among its level-L rim
            const int32_t rep = w.lab[i];
            int32_t best = GEN_UNSET;
            for (int ct = 0; ct < 8; ct++) {
                const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                    continue;
                const int64_t j = rr * cols + cc;
                if (w.alt[j] == INVALID || (w.cls[j] & CLS_DEPR) || w.W[j] != L)
                    continue;
                best = min(best, ld_i32(w.gen + j));
            }
            if (best < ld_i32(w.gen + rep))
                changed = atomicMin(w.gen + rep, best) > best;
        }
        else if (cl & CLS_INTERIOR) { // one more than the earliest source of its own level
            int32_t best = ld_i32(w.gen + i);
            const int32_t before = best;
            for (int ct = 0; ct < 8; ct++) {
                const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                    continue;
                const int64_t j = rr * cols + cc;
                if (w.alt[j] == INVALID || w.W[j] != L)
                    continue;
                const int32_t gj = ld_i32(w.gen + ((w.cls[j] & CLS_DEPR) ? w.lab[j] : (int32_t)j));
                if (gj != GEN_UNSET && gj + 1 < best)
                    best = gj + 1;
            }
            if (best < before) {
                w.gen[i] = best;
                changed = true;
            }
        }
        if (changed)
            atomicAdd(&w.ctr[C_CHANGED], 1ull);
    }
}

// distinct components of u's own generation next to rim cell u
__device__ __forceinline__ int adjacent_reps(const Work &w, int64_t u, int64_t rows, int64_t cols,
                                             int32_t *reps)
{
    const int64_t r = u / cols, c = u - r * cols;
    const int32_t L = w.W[u], gu = w.gen[u];
    int k = 0;
    for (int ct = 0; ct < 8; ct++) {
        const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
        if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
            continue;
        const int64_t j = rr * cols + cc;
        if (w.alt[j] == INVALID || !(w.cls[j] & CLS_DEPR) || w.W[j] != L)
            continue;
        const int32_t rep = w.lab[j];
        if (w.gen[rep] != gu)
            continue;
        bool seen = false;
        for (int q = 0; q < k; q++)
            seen |= reps[q] == rep;
        if (!seen)
            reps[k++] = rep;
    }
    return k;
}

// ---------------------------------------------------------------- K6
__global__ void quorum_kernel(Work w, const int32_t *list, int64_t nlist, int64_t rows, int64_t cols)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nlist;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = list[k];
        if (!(w.cls[i] & CLS_RIM))
            continue;
        int32_t reps[8];
        const int n = adjacent_reps(w, i, rows, cols, reps);
        for (int q = 0; q < n; q++)
            atomicAdd(w.quorum + reps[q], 1);
    }
}

// ---------------------------------------------------------------- K7: one round
// voted[]: low nibble = votes cast, bit 7 = episode question settled
__global__ void __launch_bounds__(128)
round_kernel(Work w, Grid g, const int32_t *list, int64_t nlist, int32_t *next,
             unsigned long long *next_count)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nlist;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t i = list[k];
        const uint8_t cl = w.cls[i];
        bool progress = false;
        // 1. direction
        if (ld_u8(w.st + i) == ST_UNKNOWN) {
            const uint8_t b = resolve_level_cell(g, i);
            if (b != ST_UNKNOWN) {
                fence();
                w.st[i] = b;
                progress = true;
            }
        }
        bool pending = w.st[i] == ST_UNKNOWN;
        if (cl & CLS_RIM) {
            int32_t reps[8];
            const int nrep = adjacent_reps(w, i, g.rows, g.cols, reps);
            uint8_t v = w.voted[i];
            int cast = v & 15;
            // 2. votes: run for trigger of every adjacent component of this generation (the
            // cell's FIFO position is known once its pusher is; a seed's always)
            if (cast < nrep && ((cl & CLS_SEED) || !pending)) {
                while (cast < nrep) {
                    const int32_t rep = reps[cast];
                    bool settled = false;
                    for (;;) {
                        const int32_t cur = ld_i32(w.trig + rep);
                        if (cur == NONE) {
                            if (atomicCAS(w.trig + rep, NONE, i) == NONE) {
                                settled = true;
                                break;
                            }
                            continue;
                        }
                        const int cmp = compare(g, i, cur);
                        if (cmp == 0)
                            break; // wait for the chains to resolve
                        if (cmp > 0) {
                            settled = true;
                            break;
                        }
                        if (atomicCAS(w.trig + rep, cur, i) == cur) {
                            settled = true;
                            break;
                        }
                    }
                    if (!settled)
                        break;
                    fence();
                    atomicAdd(w.votes + rep, 1);
                    cast++;
                    progress = true;
                }
                v = (uint8_t)((v & 0xF0) | cast);
                w.voted[i] = v;
            }
            pending |= cast < nrep;
            // 3. episode: once every vote this cell took part in is closed
            if (cast == nrep && !(v & 0x80)) {
                bool closed = true, mine = false;
                long long cap = 0;
                for (int q = 0; q < nrep; q++) {
                    closed &= ld_i32(w.votes + reps[q]) == w.quorum[reps[q]];
                }
                if (closed) {
                    fence();
                    for (int q = 0; q < nrep; q++)
                        if (ld_i32(w.trig + reps[q]) == i) {
                            mine = true;
                            cap += w.size[reps[q]];
                        }
                    if (mine) {
                        const unsigned long long at = atomicAdd(&w.ctr[C_HEAP], (unsigned long long)cap);
                        const int32_t got = episode_run(g, i, w.heap + at, (int32_t)cap, w.flags);
                        if (got != (int32_t)cap)
                            atomicAdd(&w.ctr[C_ERR], 1ull);
                        atomicAdd(&w.ctr[C_EPISODES], 1ull);
                        atomicMax(&w.ctr[C_BIGGEST], (unsigned long long)cap);
                    }
                    w.voted[i] = (uint8_t)(v | 0x80);
                    v |= 0x80;
                    progress = true;
                }
            }
            pending |= !(v & 0x80);
        }
        if (pending)
            next[atomicAdd(next_count, 1ull)] = i;
        if (progress)
            atomicAdd(&w.ctr[C_PROGRESS], 1ull);
    }
}

// ---------------------------------------------------------------- K8
__global__ void finish_kernel(Work w, int64_t rows, int64_t cols, uint8_t *packed, int64_t packed_ld)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        uint8_t b = w.st[i];
        if (b == ST_UNKNOWN) { // must not happen: reported, the cell ends the path
            atomicAdd(&w.ctr[C_ERR], 1ull);
            b = 0;
        }
        if (!(b & F_NULL)) {
            // rule R4: the cell an unworked seed was re-pointed at has its sign flipped
            for (int ct = 0; ct < 8; ct++) {
                const int64_t rr = r + ct_dr(ct), cc = c + ct_dc(ct);
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                    continue;
                const uint8_t u = w.st[rr * cols + cc];
                if ((u & (F_EDGE | F_NEG | F_NULL)) == F_EDGE && (u & F_CODE) == ct2code(ct ^ 1))
                    b |= F_TAINT;
            }
        }
        packed[r * packed_ld + c] = b;
    }
}

int grid_for(int64_t n) { return (int)min((int64_t)ht::kSMs * 16, (n + 255) / 256 > 0 ? (n + 255) / 256 : 1); }

} // namespace

extern "C" size_t ht_d8_general_workspace_bytes(int64_t rows, int64_t cols)
{
    if (rows <= 0 || cols <= 0)
        return 256;
    return work_bytes(rows, cols);
}

// stats_host[8]: [0] rounds, [1] episodes, [2] largest episode (cells), [3] depression cells,
// [4] fill passes, [5] generation sweeps, [6] cells out of the supported range, [7] errors
extern "C" int ht_d8_general_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld,
                                 int has_nodata, float nodata, double ewres, double nsres,
                                 uint8_t *packed, int64_t packed_ld, void *ws, size_t ws_bytes,
                                 long long *stats_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!dem || !packed || rows <= 0 || cols <= 0 || ld < cols || packed_ld < cols || ewres <= 0.0 ||
        nsres <= 0.0 || rows * cols >= 0x7FFFFFF0ll)
        return HT_ERR_ARG;
    Work w;
    int rc = carve(ws, ws_bytes, rows, cols, w);
    if (rc)
        return rc;
    const int64_t n = rows * cols;
    const int tiles_x = ht::cdiv(cols, FT), tiles_y = ht::cdiv(rows, FT), tiles = tiles_x * tiles_y;
    unsigned long long h[16];
    auto read_ctr = [&]() -> int {
        HT_CUDA_TRY(cudaMemcpyAsync(h, w.ctr, sizeof(h), cudaMemcpyDeviceToHost, st));
        HT_CUDA_TRY(cudaStreamSynchronize(st));
        return 0;
    };
    auto zero_ctr = [&](int k) -> int {
        HT_CUDA_TRY(cudaMemsetAsync(w.ctr + k, 0, 8, st));
        return 0;
    };
    long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    HT_CUDA_TRY(cudaMemsetAsync(w.ctr, 0, 16 * 8, st));
    alt_kernel<<<grid_for(n), 256, 0, st>>>(dem, ld, rows, cols, has_nodata, nodata, w.alt, w.ctr);
    HT_LAUNCH_CHECK();

    // ---- K2: fill
    fill_init_kernel<<<grid_for(n), 256, 0, st>>>(w.alt, rows, cols, w.W);
    HT_LAUNCH_CHECK();
    HT_CUDA_TRY(cudaMemsetAsync(w.dirty, 1, tiles, st));
    HT_CUDA_TRY(cudaMemsetAsync(w.dirty + tiles, 0, tiles, st));
    for (int pass = 0;; pass++) {
        uint8_t *din = w.dirty + (size_t)(pass & 1) * tiles, *dout = w.dirty + (size_t)((pass + 1) & 1) * tiles;
        if ((rc = zero_ctr(C_CHANGED)))
            return rc;
        fill_tile_kernel<<<tiles, 256, 0, st>>>(w.alt, w.W, rows, cols, tiles_x, tiles_y, din, dout, w.ctr);
        HT_LAUNCH_CHECK();
        HT_CUDA_TRY(cudaMemsetAsync(din, 0, tiles, st));
        if ((rc = read_ctr()))
            return rc;
        stats[4] = pass + 1;
        if (!h[C_CHANGED])
            break;
        if (pass > 4 * (tiles_x + tiles_y) * FT + 64)
            return HT_ERR_ARG; // cannot happen: W decreases monotonically
    }
    stats[6] = (long long)h[C_RANGE];

    // ---- K3 / K4
    classify_kernel<<<grid_for(n), 256, 0, st>>>(w, rows, cols);
    HT_LAUNCH_CHECK();
    if ((rc = zero_ctr(C_NSPECIAL)))
        return rc;
    collect_kernel<<<grid_for(n), 256, 0, st>>>(w, (uint8_t)(CLS_DEPR | CLS_INTERIOR | CLS_RIM), w.listB,
                                                w.ctr + C_NSPECIAL);
    HT_LAUNCH_CHECK();
    if ((rc = read_ctr()))
        return rc;
    const int64_t nspecial = (int64_t)h[C_NSPECIAL];
    if (nspecial) {
        for (int it = 0; it < 1000; it++) {
            if ((rc = zero_ctr(C_CHANGED)))
                return rc;
            ccl_hook_kernel<<<grid_for(nspecial), 256, 0, st>>>(w, w.listB, nspecial, cols, rows);
            HT_LAUNCH_CHECK();
            if ((rc = read_ctr()))
                return rc;
            if (!h[C_CHANGED])
                break;
        }
        ccl_flatten_kernel<<<grid_for(nspecial), 256, 0, st>>>(w, w.listB, nspecial);
        HT_LAUNCH_CHECK();
        // ---- K5
        for (int it = 0;; it++) {
            if ((rc = zero_ctr(C_CHANGED)))
                return rc;
            generation_kernel<<<grid_for(nspecial), 256, 0, st>>>(w, w.listB, nspecial, rows, cols);
            HT_LAUNCH_CHECK();
            if ((rc = read_ctr()))
                return rc;
            stats[5] = it + 1;
            if (!h[C_CHANGED])
                break;
            if (it > n)
                return HT_ERR_ARG;
        }
        quorum_kernel<<<grid_for(nspecial), 256, 0, st>>>(w, w.listB, nspecial, rows, cols);
        HT_LAUNCH_CHECK();
    }

    // ---- K7: rounds over the level cells
    if ((rc = zero_ctr(C_NLIST)))
        return rc;
    collect_kernel<<<grid_for(n), 256, 0, st>>>(w, (uint8_t)0, w.listA, w.ctr + C_NLIST);
    HT_LAUNCH_CHECK();
    if ((rc = read_ctr()))
        return rc;
    int64_t nlist = (int64_t)h[C_NLIST];
    Grid g;
    g.alt = w.alt; g.W = w.W; g.cls = w.cls; g.st = w.st; g.gen = w.gen; g.lab = w.lab;
    g.trig = w.trig; g.votes = w.votes; g.quorum = w.quorum; g.seq = w.seq;
    g.rows = rows; g.cols = cols; g.ew = ewres; g.ns = nsres; g.diag = sqrt(ewres * ewres + nsres * nsres);
    int32_t *cur = w.listA, *nxt = w.listB;
    while (nlist > 0) {
        if ((rc = zero_ctr(C_NLIST)) || (rc = zero_ctr(C_PROGRESS)))
            return rc;
        const int grid = (int)min((int64_t)ht::kSMs * 16, (nlist + 127) / 128);
        round_kernel<<<grid, 128, 0, st>>>(w, g, cur, nlist, nxt, w.ctr + C_NLIST);
        HT_LAUNCH_CHECK();
        if ((rc = read_ctr()))
            return rc;
        stats[0]++;
        if (!h[C_PROGRESS] && h[C_NLIST])
            return HT_ERR_DRIVER; // the rules dead-locked: must not happen
        nlist = (int64_t)h[C_NLIST];
        int32_t *t = cur; cur = nxt; nxt = t;
    }
    finish_kernel<<<grid_for(n), 256, 0, st>>>(w, rows, cols, packed, packed_ld);
    HT_LAUNCH_CHECK();
    if ((rc = read_ctr()))
        return rc;
    stats[1] = (long long)h[C_EPISODES];
    stats[2] = (long long)h[C_BIGGEST];
    stats[3] = (long long)h[C_HEAP];
    stats[7] = (long long)h[C_ERR];
    if (stats_host)
        for (int k = 0; k < 8; k++)
            stats_host[k] = stats[k];
    return h[C_ERR] ? HT_ERR_DRIVER : HT_OK;
}
