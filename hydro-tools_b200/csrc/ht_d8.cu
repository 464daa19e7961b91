// ht_d8.cu -- D8 single-flow-direction stencil, integer direction codes.
//
// Replaces the direction half of GRASS `r.watershed -s` as called by the
// reference (/root/reference/src/hydrotools/watershed.py:64-75, drainage="fd").
// On a hydrologically conditioned DEM (pit-free and tie-free after GRASS's
// round(z*1000) integer scaling -- what BASELINE.json's north_star stipulates)
// r.watershed's A* pops cells in ascending elevation, so the direction of every
// cell is a function of its 3x3 window (SURVEY.md 8(c) "local-rule corollary";
// rules R1-R4 of oracle/rwatershed.c):
//   * seed (region border, or NULL 8-neighbour): re-pointed at its LOWEST lower
//     valid neighbour if it has one (and that neighbour's accumulation sign is
//     tainted), else it keeps the negative outward code;
//   * other cells: lower neighbours in ascending order; a cardinal is taken at
//     once, a diagonal unless the flanking cardinal that is still unworked gives
//     a steeper slope (double arithmetic as in GRASS, behind an fp32 filter).
//
// Kernel: persistent CTAs, 2-stage TMA pipeline of (32+4) x (128+8) fp32 tiles
// (2-cell halo: the taint flag of a cell needs its neighbours' directions).
// Pass 0 converts the tile in place to GRASS integer millimetres, held as
// integer-valued floats (|alt| < 2^24 mm, so every difference is exact in fp32).
//  * plain tiles (no NULL cell in the box, no raster border, not the first /
//    last tile row of a band): every thread owns a 4 x 4 patch -- 128-bit LDS,
//    neighbours by warp shuffle, rolling 3-row register window, and a loop-free
//    form of the local rule in fp32 (d8_plain_cell); the four codes of a row go
//    out as one 32-bit store.  No seeds => no taint pass.
//  * all other tiles: the box is converted to int32 and the general
//    per-cell rule (seeds, NULL neighbours, taint flag) runs as three passes
//    through shared memory.
// 4 B read + 1 B written per cell (5 B/cell algorithmic).
#include <atomic>

#include "ht_common.cuh"

namespace {

constexpr int TW = 128, TH = 32;
constexpr int PADL = 4;
constexpr int BW = TW + 2 * PADL; // 136
constexpr int BH = TH + 4;        // 36
constexpr int STAGES = 2;
constexpr int THREADS = 256;
constexpr uint32_t STAGE_BYTES = BW * BH * 4;
constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
constexpr int DW = TW + 2, DH = TH + 2; // cells whose direction is computed
constexpr int DX0 = 15;                 // dirs column of ring column 0: tile column 0 sits at 16
constexpr int DP = 160;                 // pitch of the direction byte tile
constexpr int32_t INVALID = 1 << 30;    // NULL or outside the region
constexpr int32_t ALT_LIMIT = 1 << 24;  // valid |alt| < 2^24 mm: exact as fp32
constexpr int32_t DCLAMP = 1 << 27;     // differences are compared after clamping to +-2^27
constexpr float F_INVALID = 3.0e38f;    // NULL / outside, while the box holds floats

struct D8Args {
    int64_t rows, cols;       // owned rows, columns
    int64_t row0, total_rows; // position of owned row 0 in the whole raster
    int halo_top, halo_bot;   // rows available above / below (0 or >= 2)
    int ring_top, ring_bot;   // also store the packed byte of row -1 / row `rows` (neighbour band's edge row)
    int tiles_x, tiles_y;
    int has_nodata;
    float nodata;
    double ew, ns, diag;
    float inv_ew, inv_ns, inv_diag;
    float r_ew, r_ns;         // diag / ew, diag / ns
    uint8_t *packed;
    int64_t packed_ld;
    unsigned long long *counters; // [0] pits, [1] ties, [2] elevations out of range
    unsigned *next_tile;          // dynamic tile scheduler of this launch (zeroed before it)
};

// r.watershed's neighbour order ct = 0..7: S,N,W,E,NE,SW,SE,NW (cardinals first)
// (window cell, 3x3 row-major, of neighbour ct = {7, 1, 3, 5, 2, 6, 8, 0})
// GRASS direction code of neighbour ct = {6, 2, 4, 8, 1, 5, 7, 3}
__device__ __forceinline__ int ct2code(int ct) { return (0x37518426u >> (4 * ct)) & 0xF; }
// GRASS direction code that points at window cell k = {3, 2, 1, 4, 0, 8, 5, 6, 7}
__device__ __forceinline__ int win2code(int k) { return (int)((0x765804123ULL >> (4 * k)) & 0xF); }

// is fl(a/da) < fl(b/db) in double?  fp32 filter first (relative slack 1e-6
// dwarfs the 2e-7 float error), exact GRASS arithmetic only when it is close.
__device__ __forceinline__ bool slope_less(int32_t a, double da, float inv_da, int32_t b,
                                           double db, float inv_db)
{
    const float x = (float)a * inv_da, y = (float)b * inv_db;
    if (x < y * 0.999999f)
        return true;
    if (x > y * 1.000001f)
        return false;
    return (double)a / da < (double)b / db;
}

// Packed direction byte of one cell from its 3x3 integer window w0..w8.
// Neighbours are ranked through keys (clamped difference << 3 | ct): one integer
// min finds the lowest neighbour and which one it is; equal elevations (not a
// conditioned DEM) fall to the earlier neighbour in r.watershed's order.
template <bool VALIDATE>
__device__ __forceinline__ uint8_t d8_cell(int32_t w0, int32_t w1, int32_t w2, int32_t w3,
                                           int32_t c, int32_t w5, int32_t w6, int32_t w7,
                                           int32_t w8, bool border, int border_code,
                                           const D8Args &p, unsigned &pits, unsigned &ties)
{
    if (c == INVALID)
        return HT_DIR_NULL;
    auto key = [&](int32_t a, int ct) -> int32_t {
        const int32_t d = max(min(a - c, DCLAMP - 1), -DCLAMP);
        return (d << 3) | ct;
    };
    int32_t k0 = key(w7, 0), k1 = key(w1, 1), k2 = key(w3, 2), k3 = key(w5, 3);
    int32_t k4 = key(w2, 4), k5 = key(w6, 5), k6 = key(w8, 6), k7 = key(w0, 7);
    int32_t kmin = min(min(min(k0, k1), min(k2, k3)), min(min(k4, k5), min(k6, k7)));
    const int32_t nmax = max(max(max(w0, w1), max(w2, w3)), max(max(w5, w6), max(w7, w8)));
    const bool any_invalid = nmax == INVALID;
    if (VALIDATE) {
        const int32_t w[9] = {w0, w1, w2, w3, c, w5, w6, w7, w8};
        bool tie = false;
#pragma unroll
        for (int i = 0; i < 9; i++)
#pragma unroll
            for (int j = i + 1; j < 9; j++)
                tie |= (w[i] != INVALID && w[i] == w[j]);
        ties += tie;
        pits += (!border && !any_invalid && kmin >= 0);
    }
    if (border || any_invalid) {
        // rule R2 + the re-pointing half of rule R4
        if (kmin < 0)
            return (uint8_t)(ct2code(kmin & 7) | HT_DIR_EDGE);
        int code = border_code;
        if (!border) {
            // first NULL neighbour in the order S,N,W,E,NE,SW,SE,NW
            code = w0 == INVALID ? 3 : 0;
            code = w8 == INVALID ? 7 : code;
            code = w6 == INVALID ? 5 : code;
            code = w2 == INVALID ? 1 : code;
            code = w5 == INVALID ? 8 : code;
            code = w3 == INVALID ? 4 : code;
            code = w1 == INVALID ? 2 : code;
            code = w7 == INVALID ? 6 : code;
        }
        return (uint8_t)(code | HT_DIR_EDGE | HT_DIR_NEG);
    }
    // rule R4, local form: lower neighbours in ascending order; a cardinal is taken
    // at once, a diagonal unless a flanking cardinal between it and the centre
    // gives a steeper slope
    for (int it = 0; it < 5; it++) {
        if (kmin >= 0)
            return 0; // pit: not a conditioned DEM (ht_d8_validate_conditioned reports it)
        const int ct = kmin & 7;
        if (ct < 4)
            return (uint8_t)ct2code(ct);
        // diagonal ct: 4 NE, 5 SW, 6 SE, 7 NW; flanks in the centre's row / column
        const int32_t a_d = ct == 4 ? w2 : ct == 5 ? w6 : ct == 6 ? w8 : w0;
        const int32_t a_ew = (ct == 4 || ct == 6) ? w5 : w3;
        const int32_t a_ns = (ct == 4 || ct == 7) ? w1 : w7;
        const int32_t rise = c - a_d;
        bool skip = false;
        if (a_ew > a_d && a_ew < c)
            skip = slope_less(rise, p.diag, p.inv_diag, c - a_ew, p.ew, p.inv_ew);
        if (!skip && a_ns > a_d && a_ns < c)
            skip = slope_less(rise, p.diag, p.inv_diag, c - a_ns, p.ns, p.inv_ns);
        if (!skip)
            return (uint8_t)ct2code(ct);
        // drop this diagonal and look at the next lowest neighbour
        k4 = ct == 4 ? INT32_MAX : k4;
        k5 = ct == 5 ? INT32_MAX : k5;
        k6 = ct == 6 ? INT32_MAX : k6;
        k7 = ct == 7 ? INT32_MAX : k7;
        kmin = min(min(min(k0, k1), min(k2, k3)), min(min(k4, k5), min(k6, k7)));
    }
    return 0;
}


// ---------------------------------------------------------------------------
// plain tiles: all nine cells valid, no seed in sight
// ---------------------------------------------------------------------------
// Loop-free form of rule R4 on a tie-free window.  d_k = a_k - c (negative =
// lower).  r.watershed walks the lower neighbours in ascending order, takes a
// cardinal at once and skips a diagonal whose slope is smaller than that of a
// flanking cardinal which is still unworked (= higher than the diagonal).  A
// flank lower than the diagonal wins the walk anyway, so the cell chosen is the
// lowest of {cardinals} + {diagonals with slope >= both flanks}:
//     diagonal usable  <=>  d_diag <= min(d_ew * diag/ew, d_ns * diag/ns).
// Keys 8*d + ct rank the candidates (exact while |d| < 2^21); `amb` asks for the
// exact integer / double evaluation (d8_cell) when a slope comparison is within
// 1e-6 relative of equality or a drop is too large for the fp32 key.
// exact evaluation of a plain window (rare: a slope comparison too close for fp32)
__device__ __noinline__ int d8_exact_cell(float w0, float w1, float w2, float w3, float c, float w5,
                                          float w6, float w7, float w8, double ew, double ns,
                                          double diag, float inv_ew, float inv_ns, float inv_diag)
{
    D8Args p; // only the slope terms are read
    p.ew = ew; p.ns = ns; p.diag = diag;
    p.inv_ew = inv_ew; p.inv_ns = inv_ns; p.inv_diag = inv_diag;
    unsigned pp = 0, tt = 0;
    return d8_cell<false>((int)w0, (int)w1, (int)w2, (int)w3, (int)c, (int)w5, (int)w6, (int)w7,
                          (int)w8, false, 0, p, pp, tt);
}

struct Row6f {
    float v[6]; // columns -1 .. 4 relative to the thread's quad
};

__device__ __forceinline__ Row6f load_row6(const float *srow, int lane)
{
    Row6f r;
    const float4 c = *reinterpret_cast<const float4 *>(srow + PADL + 4 * lane);
    float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    float rt = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0)
        l = srow[PADL - 1];
    if (lane == 31)
        rt = srow[PADL + TW];
    r.v[0] = l; r.v[1] = c.x; r.v[2] = c.y; r.v[3] = c.z; r.v[4] = c.w; r.v[5] = rt;
    return r;
}

__device__ __forceinline__ int d8_plain_cell(float w0, float w1, float w2, float w3, float c,
                                             float w5, float w6, float w7, float w8,
                                             const D8Args &p, bool &amb)
{
    constexpr float EPS = 1e-6f, BLOCKED = 1e30f;
    const float dS = w7 - c, dN = w1 - c, dW = w3 - c, dE = w5 - c;
    const float dNE = w2 - c, dSW = w6 - c, dSE = w8 - c, dNW = w0 - c;
    // flank drops scaled to the diagonal's run
    const float sS = dS * p.r_ns, sN = dN * p.r_ns, sW = dW * p.r_ew, sE = dE * p.r_ew;
    float t_amb = 1.0f; // <= 0 if any comparison is too close to call
    auto diag_key = [&](float d, float s_ew, float s_ns, float ct) -> float {
        const float m = fminf(s_ew, s_ns);
        const float e = d - m; // > 0: a flank is steeper, the diagonal is skipped
        t_amb = fminf(t_amb, fmaf(-EPS, fabsf(m), fabsf(e)));
        // e > 0 pushes the key out of reach (e >= 2^-19 whenever it is not ambiguous)
        return fmaf(fmaxf(e, 0.0f), BLOCKED, fmaf(d, 8.0f, ct));
    };
    const float k0 = fmaf(dS, 8.0f, 0.0f), k1 = fmaf(dN, 8.0f, 1.0f);
    const float k2 = fmaf(dW, 8.0f, 2.0f), k3 = fmaf(dE, 8.0f, 3.0f);
    const float k4 = diag_key(dNE, sE, sN, 4.0f), k5 = diag_key(dSW, sW, sS, 5.0f);
    const float k6 = diag_key(dSE, sE, sS, 6.0f), k7 = diag_key(dNW, sW, sN, 7.0f);
    const float kmin = fminf(fminf(fminf(k0, k1), fminf(k2, k3)), fminf(fminf(k4, k5), fminf(k6, k7)));
    amb = t_amb <= 0.0f || kmin <= -16777208.0f;
    const int ct = __float2int_rn(kmin) & 7;
    return kmin < 0.0f ? ct2code(ct) : 0; // 0 = pit: not a conditioned DEM
}

template <bool VALIDATE>
__global__ void __launch_bounds__(THREADS, 3)
d8_kernel(const __grid_constant__ CUtensorMap map, const D8Args p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[STAGES];
    __shared__ __align__(16) uint8_t dirs[DH * DP];
    __shared__ __align__(16) uint8_t tmap[DH * DP]; // 1 where a re-pointed seed drains into the cell
    __shared__ int any_edge;

    // Tiles are handed out by a counter, not by a fixed stride: a CTA that starts late (a
    // collective's kernel held its SM when the grid launched) then simply takes fewer tiles
    // instead of finishing its fixed share late.  A CTA's first tile is blockIdx.x.
    __shared__ int stile[STAGES];
    const int tid = threadIdx.x;
    const int ntiles = p.tiles_x * p.tiles_y;
    if (tid == 0) {
        ht::tma_prefetch_desc(&map);
        for (int s = 0; s < STAGES; s++) {
            ht::mbar_init(&full[s], 1);
            stile[s] = s == 0 ? (int)blockIdx.x : (int)(gridDim.x + atomicAdd(p.next_tile, 1u));
        }
        ht::fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int s) {
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        ht::mbar_expect_tx(&full[s], STAGE_BYTES);
        ht::tma_load_2d(smem_raw + (size_t)s * STAGE_STRIDE, &map, tx * TW - PADL,
                        ty * TH - 2 + p.halo_top, &full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < STAGES; s++)
            if (stile[s] < ntiles)
                issue(stile[s], s);

    unsigned pits = 0, ties = 0, range = 0;
    for (int k = 0;; k++) {
        const int s = k % STAGES;
        // written by thread 0 before the barriers of the previous iteration / of the prologue
        const int tile = stile[s];
        if (tile >= ntiles)
            break; // the counter only grows: every later stage is past the end as well
        int32_t *alt = reinterpret_cast<int32_t *>(smem_raw + (size_t)s * STAGE_STRIDE);
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int64_t x0 = (int64_t)tx * TW, y0 = (int64_t)ty * TH;
        // rows / columns of the staged box that hold data of this raster
        // (box row yy <-> owned row y0 - 2 + yy, box column xx <-> column x0 - PADL + xx)
        const int ay_lo = (int)max((int64_t)0, 2 - y0 - p.halo_top);
        const int ay_hi = (int)min((int64_t)BH - 1, p.rows + p.halo_bot - 1 - y0 + 2);
        const int ax_lo = (int)max((int64_t)0, PADL - x0);
        const int ax_hi = (int)min((int64_t)BW - 1, p.cols - 1 - x0 + PADL);
        // ring coordinates (yy, xx) of the raster's border rows / columns, if they
        // fall into this tile's ring (else an index no cell has)
        const int64_t gy0 = p.row0 + y0 - 1, gx0 = x0 - 1; // raster coords of ring (0, 0)
        const int yy_top = (int)max((int64_t)-1, min((int64_t)DH, -gy0));
        const int yy_bot = (int)max((int64_t)-1, min((int64_t)DH, p.total_rows - 1 - gy0));
        const int xx_left = (int)max((int64_t)-1, min((int64_t)DW, -gx0));
        const int xx_right = (int)max((int64_t)-1, min((int64_t)DW, p.cols - 1 - gx0));
        if (tid == 0)
            any_edge = 0;
        ht::mbar_wait(&full[s], (k / STAGES) & 1);

        // ---- pass 0: float metres -> GRASS integer millimetres (integer-valued
        // floats), in place.  Outside the raster the TMA filled NaN.
        if (!VALIDATE)
            for (int i = tid; i < DH * DP / 16; i += THREADS)
                reinterpret_cast<uint4 *>(tmap)[i] = make_uint4(0u, 0u, 0u, 0u);
        bool saw_invalid = false;
        {
            float4 *box = reinterpret_cast<float4 *>(alt);
            for (int i = tid; i < BW * BH / 4; i += THREADS) {
                float4 q = box[i];
                float *z = reinterpret_cast<float *>(&q);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // |z| < 16 777.2 m  =>  |alt| < 2^24 - 15; false for NaN
                    const bool small = fabsf(z[j]) < 16777.2f;
                    const bool nd = p.has_nodata && z[j] == p.nodata;
                    float a = F_INVALID;
                    if (small && !nd) {
                        // rule R1: (int)(z * 1000 +- 0.5) in double; the half takes z's sign.
                        // (An exact fp32 formulation -- product, FMA residual, rint and two
                        // half-way tests -- was measured: 0.38 ms against 0.34, the kernel is
                        // bound by instruction issue and the fp64 / conversion pipes are idle.)
                        const double half = __hiloint2double(
                            0x3FE00000 | (int)(__float_as_uint(z[j]) & 0x80000000u), 0);
                        a = (float)(int)((double)z[j] * 1000.0 + half);
                    }
                    else if (!nd && z[j] == z[j]) { // out of range: clamped and counted
                        a = z[j] > 0 ? (float)(ALT_LIMIT - 1) : (float)(1 - ALT_LIMIT);
                        range++;
                    }
                    else
                        saw_invalid = true;
                    z[j] = a;
                }
                box[i] = q;
            }
        }
        // plain tile: nothing but valid cells in the box, no raster border in the
        // ring, and not a tile row that also stores a neighbour band's halo row
        const bool framed = yy_top < 0 && xx_left < 0 && yy_bot >= DH && xx_right >= DW &&
                            !(ty == 0 && p.ring_top) && !(ty == p.tiles_y - 1 && p.ring_bot) &&
                            y0 + TH <= p.rows && x0 + TW <= p.cols;
        const bool plain = !__syncthreads_or(saw_invalid) && framed && !VALIDATE;

        if (plain) {
            // ---- warp -> 4 rows, lane -> 4 columns, rolling 3-row window
            const int warp = tid >> 5, lane = tid & 31;
            const float *fb = reinterpret_cast<const float *>(alt);
            const int yyc = 2 + 4 * warp; // box row of the first centre row
            Row6f ra = load_row6(fb + (yyc - 1) * BW, lane);
            Row6f rb = load_row6(fb + yyc * BW, lane);
            uint8_t *dst = p.packed + (y0 + 4 * warp) * p.packed_ld + x0 + 4 * lane;
#pragma unroll
            for (int i = 0; i < 4; i++, dst += p.packed_ld) {
                const Row6f rc = load_row6(fb + (yyc + 1 + i) * BW, lane);
                uint32_t out = 0;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    bool amb;
                    int code = d8_plain_cell(ra.v[c], ra.v[c + 1], ra.v[c + 2], rb.v[c], rb.v[c + 1],
                                             rb.v[c + 2], rc.v[c], rc.v[c + 1], rc.v[c + 2], p, amb);
                    if (amb)
                        code = d8_exact_cell(ra.v[c], ra.v[c + 1], ra.v[c + 2], rb.v[c], rb.v[c + 1],
                                             rb.v[c + 2], rc.v[c], rc.v[c + 1], rc.v[c + 2], p.ew, p.ns,
                                             p.diag, p.inv_ew, p.inv_ns, p.inv_diag);
                    out |= (uint32_t)code << (8 * c);
                }
                *reinterpret_cast<uint32_t *>(dst) = out;
                ra = rb;
                rb = rc;
            }
        }
        else {
        // ---- general tiles: three passes through shared memory.  Cells whose window is
        // all valid and off the raster border still take the fp32 rule; the integer rule
        // (seeds, NULL neighbours) runs only next to NULL cells and borders.  The
        // validating instantiation compares integers (ties) and converts the box first.
        if (VALIDATE) {
            for (int i = tid; i < BW * BH; i += THREADS) {
                const float f = __int_as_float(alt[i]);
                alt[i] = f == F_INVALID ? INVALID : (int32_t)f;
            }
            __syncthreads();
        }

        // ---- pass 1: direction byte of the tile and a 1-cell ring around it
        bool edge_seen = false;
        for (int i = tid; i < DW * DH; i += THREADS) {
            const int yy = i / DW, xx = i - yy * DW; // ring coords; box coords +1, +PADL-1
            const int32_t *a = alt + (yy + 1) * BW + (xx + PADL - 1);
            const bool border = yy == yy_top || xx == xx_left || yy == yy_bot || xx == xx_right;
            const int border_code = yy == yy_top ? 2 : xx == xx_left ? 4 : yy == yy_bot ? 6 : 8;
            unsigned pp = 0, tt = 0;
            uint8_t b;
            if (VALIDATE)
                b = d8_cell<VALIDATE>(a[-BW - 1], a[-BW], a[-BW + 1], a[-1], a[0], a[1], a[BW - 1],
                                      a[BW], a[BW + 1], border, border_code, p, pp, tt);
            else {
                const float *f = reinterpret_cast<const float *>(a);
                const float w0 = f[-BW - 1], w1 = f[-BW], w2 = f[-BW + 1], w3 = f[-1], c = f[0],
                            w5 = f[1], w6 = f[BW - 1], w7 = f[BW], w8 = f[BW + 1];
                const float mx = fmaxf(fmaxf(fmaxf(w0, w1), fmaxf(w2, w3)), fmaxf(fmaxf(w5, w6), fmaxf(w7, w8)));
                if (c == F_INVALID)
                    b = HT_DIR_NULL;
                else if (!border && mx < F_INVALID) {
                    bool amb;
                    int code = d8_plain_cell(w0, w1, w2, w3, c, w5, w6, w7, w8, p, amb);
                    if (amb)
                        code = d8_exact_cell(w0, w1, w2, w3, c, w5, w6, w7, w8, p.ew, p.ns, p.diag,
                                             p.inv_ew, p.inv_ns, p.inv_diag);
                    b = (uint8_t)code;
                }
                else {
                    auto iv = [](float v) -> int32_t { return v == F_INVALID ? INVALID : (int32_t)v; };
                    b = d8_cell<false>(iv(w0), iv(w1), iv(w2), iv(w3), iv(c), iv(w5), iv(w6), iv(w7),
                                       iv(w8), border, border_code, p, pp, tt);
                }
            }
            dirs[yy * DP + xx + DX0] = b;
            edge_seen |= (b & HT_DIR_EDGE) != 0;
            if (!VALIDATE && (b & (HT_DIR_EDGE | HT_DIR_NEG | HT_DIR_NULL)) == HT_DIR_EDGE) {
                // rule R4: the cell an unworked seed is re-pointed at has its accumulation
                // sign flipped.  The seed tells its receiver (all writers store the same 1).
                const int code = b & HT_DIR_CODE;
                const int ty2 = yy + (int)((0x1A900u >> (2 * code)) & 3u) - 1;
                const int tx2 = xx + (int)((0x29018u >> (2 * code)) & 3u) - 1;
                if (ty2 >= 0 && ty2 < DH && tx2 >= 0 && tx2 < DW)
                    tmap[ty2 * DP + tx2 + DX0] = 1;
            }
            if (VALIDATE) {
                // only cells of the tile itself are counted
                const bool own = yy >= 1 && yy <= TH && xx >= 1 && xx <= TW &&
                                 y0 - 1 + yy < p.rows && x0 - 1 + xx < p.cols;
                if (own) {
                    pits += pp;
                    ties += tt;
                }
            }
        }
        if (edge_seen)
            any_edge = 1;
        __syncthreads();

        // ---- pass 2: taint flags collected in pass 1, 128-bit row stores
        if (!VALIDATE) {
            const bool need_taint = any_edge != 0;
            // item -> 16 consecutive cells of one row (8 items per row).  Rows -1 and
            // `rows` (first halo row of a neighbouring band) are stored too, without
            // the taint flag, for the accumulation stages' ring.
            for (int item = tid; item < (TH + 2) * 8; item += THREADS) {
                const int yy = item >> 3, seg = item & 7;
                const int64_t lr = y0 - 1 + yy;
                const bool own = yy >= 1 && yy <= TH && lr < p.rows;
                const bool halo_row = (lr == -1 && p.ring_top) || (lr == p.rows && p.ring_bot);
                const int64_t gc = x0 + seg * 16;
                if (!(own || halo_row) || gc >= p.cols)
                    continue;
                const uint8_t *row = dirs + yy * DP + DX0 + 1 + seg * 16; // 16-B aligned
                uint4 v = *reinterpret_cast<const uint4 *>(row);
                if (need_taint && own) {
                    const uint4 t = *reinterpret_cast<const uint4 *>(tmap + yy * DP + DX0 + 1 + seg * 16);
                    v.x |= t.x << 6; v.y |= t.y << 6; v.z |= t.z << 6; v.w |= t.w << 6; // HT_DIR_TAINT
                }
                uint8_t *dst = p.packed + lr * p.packed_ld + gc;
                if (gc + 16 <= p.cols)
                    *reinterpret_cast<uint4 *>(dst) = v;
                else {
                    const uint32_t out[4] = {v.x, v.y, v.z, v.w};
                    for (int j = 0; j < 16 && gc + j < p.cols; j++)
                        dst[j] = (uint8_t)(out[j >> 2] >> (8 * (j & 3)));
                }
            }
        }

        } // general tile

        ht::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const int next = (int)(gridDim.x + atomicAdd(p.next_tile, 1u));
            stile[s] = next;
            if (next < ntiles)
                issue(next, s);
        }
    }
    if (VALIDATE) {
        if (pits)
            atomicAdd(&p.counters[0], (unsigned long long)pits);
        if (ties)
            atomicAdd(&p.counters[1], (unsigned long long)ties);
        if (range)
            atomicAdd(&p.counters[2], (unsigned long long)range);
    }
}

int setup(const float *dem, int64_t rows, int64_t cols, int64_t ld, int has_nodata,
          float nodata, double ewres, double nsres, int64_t row0, int64_t total_rows,
          int halo_top, int halo_bot, D8Args &p, CUtensorMap &map)
{
    if (!dem || rows <= 0 || cols <= 0 || ld < cols || ewres <= 0.0 || nsres <= 0.0 ||
        row0 < 0 || total_rows < row0 + rows || halo_top < 0 || halo_bot < 0)
        return HT_ERR_ARG;
    // a side without halo rows must be the edge of the whole raster
    if ((halo_top == 0 && row0 != 0) || (halo_bot == 0 && row0 + rows != total_rows) ||
        (halo_top != 0 && halo_top < 2 && row0 >= 2) ||
        (halo_bot != 0 && halo_bot < 2 && total_rows - row0 - rows >= 2))
        return HT_ERR_ARG;
    p.rows = rows; p.cols = cols; p.row0 = row0; p.total_rows = total_rows;
    p.halo_top = halo_top > 2 ? 2 : halo_top;
    p.halo_bot = halo_bot > 2 ? 2 : halo_bot;
    p.ring_top = p.halo_top ? 1 : 0;
    p.ring_bot = p.halo_bot ? 1 : 0;
    p.tiles_x = ht::cdiv(cols, TW); p.tiles_y = ht::cdiv(rows, TH);
    p.has_nodata = has_nodata ? 1 : 0;
    p.nodata = nodata;
    p.ew = ewres; p.ns = nsres; p.diag = sqrt(ewres * ewres + nsres * nsres);
    p.inv_ew = (float)(1.0 / p.ew); p.inv_ns = (float)(1.0 / p.ns);
    p.inv_diag = (float)(1.0 / p.diag);
    p.r_ew = (float)(p.diag / p.ew); p.r_ns = (float)(p.diag / p.ns);
    p.packed = nullptr; p.packed_ld = 0; p.counters = nullptr; p.next_tile = nullptr;
    const float *base = dem - (int64_t)p.halo_top * ld;
    return ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base,
                           rows + p.halo_top + p.halo_bot, cols, ld, BW, BH, true);
}

// One scheduler counter per launch in flight: a ring of 256, so that launches on different
// streams (or queued behind each other) never share one.
__device__ unsigned g_next_tile[256];

template <bool VALIDATE>
int launch(const CUtensorMap &map, D8Args p, cudaStream_t st)
{
    static std::atomic<unsigned> ring{0};
    unsigned *counters = nullptr;
    HT_CUDA_TRY(cudaGetSymbolAddress((void **)&counters, g_next_tile));
    p.next_tile = counters + (ring.fetch_add(1u) & 255u);
    HT_CUDA_TRY(cudaMemsetAsync(p.next_tile, 0, sizeof(unsigned), st));
    auto kern = d8_kernel<VALIDATE>;
    const size_t smem = (size_t)STAGES * STAGE_STRIDE;
    HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = p.tiles_x * p.tiles_y;
    int grid = 3 * ht::kSMs;
    if (grid > ntiles)
        grid = ntiles;
    kern<<<grid, THREADS, smem, st>>>(map, p);
    HT_LAUNCH_CHECK();
    return 0;
}

} // namespace

// A run of rows of a band: like ht_d8_f32, but the packed byte of the row above / below the
// run is stored only on request (ring_top / ring_bot) -- a row-band rank computes its
// interior rows while the halo rows of its neighbours are still in flight, then the first
// and last tile rows, and only those two calls also store the neighbours' edge rows.
extern "C" int ht_d8_part_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld,
                              int has_nodata, float nodata, double ewres, double nsres,
                              int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                              int ring_top, int ring_bot, uint8_t *packed, int64_t packed_ld,
                              void *stream)
{
    if (rows == 0 || cols == 0)
        return HT_OK;
    D8Args p;
    CUtensorMap map;
    int rc = setup(dem, rows, cols, ld, has_nodata, nodata, ewres, nsres, row0, total_rows,
                   halo_top, halo_bot, p, map);
    if (rc)
        return rc;
    if (!packed || packed_ld < cols || (ring_top && !halo_top) || (ring_bot && !halo_bot))
        return HT_ERR_ARG;
    if (((uintptr_t)packed & 15) || (packed_ld & 15))
        return HT_ERR_ALIGN;
    p.packed = packed; p.packed_ld = packed_ld;
    p.ring_top = ring_top ? 1 : 0;
    p.ring_bot = ring_bot ? 1 : 0;
    return launch<false>(map, p, (cudaStream_t)stream);
}

extern "C" int ht_d8_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld,
                         int has_nodata, float nodata, double ewres, double nsres,
                         int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                         uint8_t *packed, int64_t packed_ld, void *stream)
{
    return ht_d8_part_f32(dem, rows, cols, ld, has_nodata, nodata, ewres, nsres, row0, total_rows,
                          halo_top, halo_bot, halo_top != 0, halo_bot != 0, packed, packed_ld, stream);
}

extern "C" int ht_d8_validate_conditioned_f32(const float *dem, int64_t rows, int64_t cols,
                                              int64_t ld, int has_nodata, float nodata,
                                              int64_t row0, int64_t total_rows, int halo_top,
                                              int halo_bot, unsigned long long *counts3,
                                              void *stream)
{
    if (!counts3)
        return HT_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    HT_CUDA_TRY(cudaMemsetAsync(counts3, 0, 3 * sizeof(unsigned long long), st));
    if (rows == 0 || cols == 0)
        return HT_OK;
    D8Args p;
    CUtensorMap map;
    int rc = setup(dem, rows, cols, ld, has_nodata, nodata, 1.0, 1.0, row0, total_rows,
                   halo_top, halo_bot, p, map);
    if (rc)
        return rc;
    p.counters = counts3;
    return launch<true>(map, p, st);
}
