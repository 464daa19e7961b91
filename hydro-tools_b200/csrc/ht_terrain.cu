// ht_terrain.cu -- fused slope + aspect + TRI over ONE read of the DEM tile.
//
// Replaces three gdaldem subprocesses of the reference
// (/root/reference/src/hydrotools/elevation.py:41-54 slope, :70-78 aspect,
//  :118-126 TRI; each `gdaldem MODE -compute_edges`), following GDAL 3.11
// apps/gdaldem_lib.cpp semantics (rules G1-G6 in oracle/gdaldem.c).
//
// Kernel: persistent CTAs, 3-stage TMA (cp.async.bulk.tensor.2d) pipeline of
// (32+2) x (128+8) fp32 halo tiles in shared memory, 256 threads, each thread
// a 4x4 patch: 128-bit LDS of the centre quad, neighbours by warp shuffle,
// 128-bit streaming stores.  HBM-bound by design: 4 B read + 12 B written per
// cell (16 B/cell algorithmic); fp32 math with MUFU sqrt/rcp and a degree-6
// minimax atan core (relative error < 1e-6, inside the 1e-5 contract).
#include <math.h>

#include "ht_common.cuh"

namespace {

constexpr int TW = 128;         // tile width  (cells)
constexpr int TH = 32;          // tile height (cells)
constexpr int PADL = 4;         // left pad keeps the centre quads 16-B aligned
constexpr int BW = TW + 2 * PADL; // 136 floats = 544 B per box row
constexpr int BH = TH + 2;
constexpr int STAGES = 3;
constexpr int THREADS = 256;
constexpr uint32_t STAGE_BYTES = BW * BH * sizeof(float);
// TMA destinations must be 128-B aligned
constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 127) / 128 * 128;
constexpr uint32_t STAGE_FLOATS = STAGE_STRIDE / sizeof(float);

struct TerrainArgs {
    int64_t rows, cols;  // owned rows, columns
    int halo_top, halo_bot;
    int tiles_x, tiles_y;
    int has_nodata, nodata_is_nan;
    float nodata;
    float nd_lo, nd_hi;  // closed interval of floats that ARE_REAL_EQUAL to nodata
    float inv_ew, inv_ns, inv_8scale;
    int percent;
    float *slope, *aspect, *tri;
    int64_t out_ld;
};

// ARE_REAL_EQUAL(v, nodata) of gdal_priv.h (== or within 2 float ulps); NaN
// nodata -> isnan (rule G2).  The set of floats equal to nodata in that sense is
// an interval around it, found once on the host (nodata_interval).
__host__ __device__ __forceinline__ bool are_real_equal(float a, float b)
{
    return a == b || fabsf(a - b) < 1.1920929e-07f * fabsf(a + b) * 2.0f;
}

__device__ __forceinline__ bool is_nd(float v, const TerrainArgs &p)
{
    if (p.nodata_is_nan)
        return v != v;
    return v >= p.nd_lo && v <= p.nd_hi;
}

// INTERPOL(a, b) of gdaldem_lib.cpp (rule G3)
__device__ __forceinline__ float interpol(float a, float b, const TerrainArgs &p)
{
    if (p.has_nodata && !p.nodata_is_nan && (is_nd(a, p) || is_nd(b, p)))
        return p.nodata;
    return 2.0f * a - b; // NaN nodata propagates as NaN by itself
}

// atan(t)/t on [0,1], minimax in relative error (7.4e-7)
__device__ __forceinline__ float atan01(float t)
{
    const float u = t * t;
    float p = 0.008007033728063107f;
    p = fmaf(p, u, -0.03744340315461159f);
    p = fmaf(p, u, 0.08435460180044174f);
    p = fmaf(p, u, -0.13512206077575684f);
    p = fmaf(p, u, 0.19887329638004303f);
    p = fmaf(p, u, -0.3332701325416565f);
    p = fmaf(p, u, 0.9999994039535522f);
    return p * t;
}

constexpr float kRad2Deg = 57.29577951308232f;
constexpr float kHalfPi = 1.5707963267948966f;
constexpr float kPi = 3.14159265358979f;

// atan of g >= 0, in degrees
__device__ __forceinline__ float atan_pos_deg(float g)
{
    const bool big = g > 1.0f;
    const float t = big ? ht::fast_rcp(g) : g;
    float r = atan01(t);
    r = big ? kHalfPi - r : r;
    return r * kRad2Deg;
}

// atan2(y, x) in degrees, (-180, 180]
__device__ __forceinline__ float atan2_deg(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float r = atan01(mn * ht::fast_rcp(mx));
    r = ay > ax ? kHalfPi - r : r;
    r = x < 0.0f ? kPi - r : r;
    r = r * kRad2Deg;
    return y < 0.0f ? -r : r;
}

struct Out3 {
    float s, a, t;
};

// Rules G4-G6 on a window whose nodata has already been resolved.
template <bool WS, bool WA, bool WT>
__device__ __forceinline__ Out3 horn(const float w0, const float w1, const float w2,
                                     const float w3, const float w4, const float w5,
                                     const float w6, const float w7, const float w8,
                                     const TerrainArgs &p)
{
    Out3 o;
    o.s = o.a = o.t = 0.0f;
    if (WS || WA) {
        // float sums in gdaldem's order (the products would be exact in double,
        // the sums are what rounds)
        const float left = __fadd_rn(__fadd_rn(__fadd_rn(w0, w3), w3), w6);
        const float right = __fadd_rn(__fadd_rn(__fadd_rn(w2, w5), w5), w8);
        const float bot = __fadd_rn(__fadd_rn(__fadd_rn(w6, w7), w7), w8);
        const float top = __fadd_rn(__fadd_rn(__fadd_rn(w0, w1), w1), w2);
        const float dxs = __fsub_rn(left, right); // slope dx numerator = -(aspect dx)
        const float dy = __fsub_rn(bot, top);
        if (WS) {
            const float gx = dxs * p.inv_ew, gy = dy * p.inv_ns;
            const float g = ht::fast_sqrt(fmaf(gx, gx, gy * gy)) * p.inv_8scale;
            o.s = p.percent ? 100.0f * g : atan_pos_deg(g);
        }
        if (WA) {
            // aspect dx = right - left = -dxs;  atan2(dy, -dx) = atan2(dy, dxs)
            float a = atan2_deg(dy, dxs);
            if (dxs == 0.0f && dy == 0.0f)
                a = HT_TERRAIN_NODATA;
            else
                a = a > 90.0f ? 450.0f - a : 90.0f - a;
            if (a == 360.0f)
                a = 0.0f;
            o.a = a;
        }
    }
    if (WT) {
        float d, s;
        d = w0 - w4; s = d * d;
        d = w1 - w4; s = fmaf(d, d, s);
        d = w2 - w4; s = fmaf(d, d, s);
        d = w3 - w4; s = fmaf(d, d, s);
        d = w5 - w4; s = fmaf(d, d, s);
        d = w6 - w4; s = fmaf(d, d, s);
        d = w7 - w4; s = fmaf(d, d, s);
        d = w8 - w4; s = fmaf(d, d, s);
        o.t = ht::fast_sqrt(s);
    }
    return o;
}

struct Row6 {
    float v[6];    // columns -1 .. 4 relative to the thread's quad
    unsigned nd;   // bit k: v[k] is nodata
};

template <bool ND>
__device__ __forceinline__ Row6 load_row(const float *srow, int lane,
                                         const TerrainArgs &p)
{
    Row6 r;
    const float4 c = *reinterpret_cast<const float4 *>(srow + PADL + 4 * lane);
    float l = __shfl_up_sync(0xffffffffu, c.w, 1);
    float rt = __shfl_down_sync(0xffffffffu, c.x, 1);
    if (lane == 0)
        l = srow[PADL - 1];
    if (lane == 31)
        rt = srow[PADL + TW];
    r.v[0] = l; r.v[1] = c.x; r.v[2] = c.y; r.v[3] = c.z; r.v[4] = c.w; r.v[5] = rt;
    r.nd = 0;
    if (ND) {
#pragma unroll
        for (int k = 0; k < 6; k++)
            r.nd |= is_nd(r.v[k], p) ? (1u << k) : 0u;
    }
    return r;
}

// Stencil over one staged tile: warp -> 4 rows, lane -> 4 columns.
template <bool ND, bool WS, bool WA, bool WT>
__device__ __forceinline__ void stencil_tile(const float *sm, int warp, int lane, int64_t x0,
                                             int64_t y0, const TerrainArgs &p)
{
    const int yyc = 1 + 4 * warp; // smem row of the first centre row
    Row6 ra = load_row<ND>(sm + (yyc - 1) * BW, lane, p);
    Row6 rb = load_row<ND>(sm + yyc * BW, lane, p);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const Row6 rc = load_row<ND>(sm + (yyc + 1 + i) * BW, lane, p);
        const int64_t gr = y0 + 4 * warp + i;
        float os[4], oa[4], ot[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float w0 = ra.v[c], w1 = ra.v[c + 1], w2 = ra.v[c + 2];
            float w3 = rb.v[c], w4 = rb.v[c + 1], w5 = rb.v[c + 2];
            float w6 = rc.v[c], w7 = rc.v[c + 1], w8 = rc.v[c + 2];
            bool centre_nd = false;
            if (ND) {
                centre_nd = (rb.nd >> (c + 1)) & 1u;
                const unsigned ma = ra.nd >> c, mb = rb.nd >> c, mc = rc.nd >> c;
                w0 = (ma & 1u) ? w4 : w0; w1 = (ma & 2u) ? w4 : w1; w2 = (ma & 4u) ? w4 : w2;
                w3 = (mb & 1u) ? w4 : w3;                           w5 = (mb & 4u) ? w4 : w5;
                w6 = (mc & 1u) ? w4 : w6; w7 = (mc & 2u) ? w4 : w7; w8 = (mc & 4u) ? w4 : w8;
            }
            Out3 o = horn<WS, WA, WT>(w0, w1, w2, w3, w4, w5, w6, w7, w8, p);
            if (ND && centre_nd)
                o.s = o.a = o.t = HT_TERRAIN_NODATA;
            os[c] = o.s; oa[c] = o.a; ot[c] = o.t;
        }
        const int64_t gc = x0 + 4 * lane;
        if (gr < p.rows && gc < p.cols) {
            const int64_t off = gr * p.out_ld + gc;
            if (gc + 3 < p.cols) {
                if (WS) ht::st_stream_f4(p.slope + off, make_float4(os[0], os[1], os[2], os[3]));
                if (WA) ht::st_stream_f4(p.aspect + off, make_float4(oa[0], oa[1], oa[2], oa[3]));
                if (WT) ht::st_stream_f4(p.tri + off, make_float4(ot[0], ot[1], ot[2], ot[3]));
            }
            else {
                for (int c = 0; c < 4 && gc + c < p.cols; c++) {
                    if (WS) p.slope[off + c] = os[c];
                    if (WA) p.aspect[off + c] = oa[c];
                    if (WT) p.tri[off + c] = ot[c];
                }
            }
        }
        ra = rb;
        rb = rc;
    }
}

template <bool ND, bool WS, bool WA, bool WT>
__global__ void __launch_bounds__(THREADS, 3)
terrain_kernel(const __grid_constant__ CUtensorMap map, const TerrainArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage0 = reinterpret_cast<float *>(smem_raw);
    __shared__ __align__(8) uint64_t full[STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntiles = p.tiles_x * p.tiles_y;

    if (tid == 0) {
        ht::tma_prefetch_desc(&map);
        for (int s = 0; s < STAGES; s++)
            ht::mbar_init(&full[s], 1);
        ht::fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int s) {
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        ht::mbar_expect_tx(&full[s], STAGE_BYTES);
        ht::tma_load_2d(stage0 + (size_t)s * STAGE_FLOATS, &map, tx * TW - PADL,
                        ty * TH - 1 + p.halo_top, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) {
            const int t = blockIdx.x + s * gridDim.x;
            if (t < ntiles)
                issue(t, s);
        }
    }

    int k = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
        const int s = k % STAGES;
        float *sm = stage0 + (size_t)s * STAGE_FLOATS;
        const int tx = tile % p.tiles_x, ty = tile / p.tiles_x;
        const int64_t x0 = (int64_t)tx * TW, y0 = (int64_t)ty * TH;
        ht::mbar_wait(&full[s], (k / STAGES) & 1);

        // ---- rule G3: synthesize the missing line / column at raster edges
        const bool top_edge = (y0 == 0 && p.halo_top == 0);
        const bool bot_edge = (y0 + TH >= p.rows && p.halo_bot == 0);
        const bool left_edge = (x0 == 0);
        const bool right_edge = (x0 + TW >= p.cols);
        if (top_edge || bot_edge || left_edge || right_edge) {
            // real rows held in this stage: yy in [yy_lo, yy_hi]
            const int yy_lo = top_edge ? 1 : 0;
            const int yy_last = (int)(p.rows - 1 - y0) + 1; // last owned row
            const int yy_hi = bot_edge ? yy_last : BH - 1;
            const int xx_last = (int)(p.cols - 1 - x0) + PADL;
            if (left_edge)
                for (int yy = yy_lo + tid; yy <= yy_hi; yy += THREADS)
                    sm[yy * BW + PADL - 1] =
                        interpol(sm[yy * BW + PADL], sm[yy * BW + PADL + 1], p);
            if (right_edge)
                for (int yy = yy_lo + tid; yy <= yy_hi; yy += THREADS)
                    sm[yy * BW + xx_last + 1] =
                        interpol(sm[yy * BW + xx_last], sm[yy * BW + xx_last - 1], p);
            const int xx_lo = left_edge ? PADL : 0;
            const int xx_hi = right_edge ? xx_last : BW - 1;
            if (top_edge)
                for (int xx = xx_lo + tid; xx <= xx_hi; xx += THREADS)
                    sm[xx] = interpol(sm[BW + xx], sm[2 * BW + xx], p);
            if (bot_edge)
                for (int xx = xx_lo + tid; xx <= xx_hi; xx += THREADS)
                    sm[(yy_last + 1) * BW + xx] =
                        interpol(sm[yy_last * BW + xx], sm[(yy_last - 1) * BW + xx], p);
            __syncthreads();
        }

        // ---- stencil; tiles that hold no nodata at all take the plain path
        if (ND) {
            bool mine = false;
            for (int i = tid; i < BW * BH; i += THREADS)
                mine |= is_nd(sm[i], p);
            if (__syncthreads_or(mine))
                stencil_tile<true, WS, WA, WT>(sm, warp, lane, x0, y0, p);
            else
                stencil_tile<false, WS, WA, WT>(sm, warp, lane, x0, y0, p);
        }
        else
            stencil_tile<false, WS, WA, WT>(sm, warp, lane, x0, y0, p);

        // ---- recycle the stage: everyone is done reading (and patching) it
        ht::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            const int next = tile + STAGES * gridDim.x;
            if (next < ntiles)
                issue(next, s);
        }
    }
}

// Rule G3, corner cells: on the first/last line gdaldem CLAMPS the column index
// at j == 0 and j == n-1 instead of extrapolating.  Four cells per raster, so
// they are recomputed here instead of taxing the streaming kernel.
template <bool ND>
__global__ void terrain_corner_kernel(const float *dem, int64_t ld, const TerrainArgs p)
{
    const int which = threadIdx.x; // 0 TL, 1 TR, 2 BL, 3 BR
    const bool top = which < 2, leftc = !(which & 1);
    if (top && p.halo_top)
        return;
    if (!top && p.halo_bot)
        return;
    const int64_t i = top ? 0 : p.rows - 1;
    const int64_t inb = top ? 1 : p.rows - 2; // neighbour line (may be a halo row)
    const int64_t j = leftc ? 0 : p.cols - 1;
    const int64_t jm = leftc ? j : j - 1, jp = leftc ? j + 1 : j;
    const int64_t jj[3] = {jm, j, jp};
    float w[9];
    for (int k = 0; k < 3; k++) {
        const float a = dem[i * ld + jj[k]], b = dem[inb * ld + jj[k]];
        const float e = interpol(a, b, p);
        if (top) { w[k] = e; w[3 + k] = a; w[6 + k] = b; }
        else     { w[k] = b; w[3 + k] = a; w[6 + k] = e; }
    }
    Out3 o;
    if (ND && is_nd(w[4], p)) {
        o.s = o.a = o.t = HT_TERRAIN_NODATA;
    }
    else {
        if (ND)
            for (int k = 0; k < 9; k++)
                if (is_nd(w[k], p))
                    w[k] = w[4];
        o = horn<true, true, true>(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], w[8], p);
    }
    const int64_t off = i * p.out_ld + j;
    if (p.slope) p.slope[off] = o.s;
    if (p.aspect) p.aspect[off] = o.a;
    if (p.tri) p.tri[off] = o.t;
}

__global__ void fill_f32_kernel(float *out, int64_t rows, int64_t cols, int64_t ld, float v)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[(i / cols) * ld + (i % cols)] = v;
}

template <bool ND, bool WS, bool WA, bool WT>
int launch(const CUtensorMap &map, const TerrainArgs &p, cudaStream_t st)
{
    auto kern = terrain_kernel<ND, WS, WA, WT>;
    const size_t smem = (size_t)STAGES * STAGE_STRIDE;
    HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = p.tiles_x * p.tiles_y;
    int grid = 3 * ht::kSMs; // three resident CTAs per SM, persistent
    if (grid > ntiles)
        grid = ntiles;
    kern<<<grid, THREADS, smem, st>>>(map, p);
    HT_LAUNCH_CHECK();
    return 0;
}

template <bool ND>
int dispatch(const CUtensorMap &map, const TerrainArgs &p, cudaStream_t st)
{
    const int m = (p.slope ? 1 : 0) | (p.aspect ? 2 : 0) | (p.tri ? 4 : 0);
    switch (m) {
    case 1: return launch<ND, true, false, false>(map, p, st);
    case 2: return launch<ND, false, true, false>(map, p, st);
    case 3: return launch<ND, true, true, false>(map, p, st);
    case 4: return launch<ND, false, false, true>(map, p, st);
    case 5: return launch<ND, true, false, true>(map, p, st);
    case 6: return launch<ND, false, true, true>(map, p, st);
    case 7: return launch<ND, true, true, true>(map, p, st);
    }
    return HT_ERR_ARG;
}

} // namespace

extern "C" int ht_terrain_fused_f32(const float *dem, int64_t rows, int64_t cols,
                                    int64_t ld, int has_nodata, float nodata,
                                    double ewres, double nsres, double zscale,
                                    int slope_percent, float *slope, float *aspect,
                                    float *tri, int64_t out_ld, int halo_top,
                                    int halo_bot, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!dem || rows < 0 || cols < 0 || ld < cols || out_ld < cols || halo_top < 0 ||
        halo_bot < 0 || (!slope && !aspect && !tri) || ewres == 0.0 || nsres == 0.0 ||
        zscale == 0.0)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    float *outs[3] = {slope, aspect, tri};
    for (float *o : outs)
        if (o && (((uintptr_t)o & 15) || (out_ld & 3)))
            return HT_ERR_ALIGN;
    if (rows + halo_top + halo_bot < 2 || cols < 2) { // rule G3, degenerate raster
        for (float *o : outs)
            if (o) {
                fill_f32_kernel<<<ht::kSMs, 256, 0, st>>>(o, rows, cols, out_ld, HT_TERRAIN_NODATA);
                HT_LAUNCH_CHECK();
            }
        return HT_OK;
    }

    TerrainArgs p;
    p.rows = rows; p.cols = cols;
    p.halo_top = halo_top ? 1 : 0; p.halo_bot = halo_bot ? 1 : 0;
    p.tiles_x = ht::cdiv(cols, TW); p.tiles_y = ht::cdiv(rows, TH);
    p.has_nodata = has_nodata ? 1 : 0;
    p.nodata = nodata;
    p.nodata_is_nan = (has_nodata && nodata != nodata) ? 1 : 0;
    p.nd_lo = p.nd_hi = nodata;
    if (has_nodata && !p.nodata_is_nan && !isinf(nodata)) {
        // widen ulp by ulp while ARE_REAL_EQUAL still holds (a handful of steps)
        for (int i = 0; i < 64; i++) {
            const float lo = nextafterf(p.nd_lo, -INFINITY), hi = nextafterf(p.nd_hi, INFINITY);
            bool grew = false;
            if (!isinf(lo) && are_real_equal(lo, nodata)) { p.nd_lo = lo; grew = true; }
            if (!isinf(hi) && are_real_equal(hi, nodata)) { p.nd_hi = hi; grew = true; }
            if (!grew)
                break;
        }
    }
    p.inv_ew = (float)(1.0 / ewres);
    p.inv_ns = (float)(1.0 / nsres);
    p.inv_8scale = (float)(1.0 / (8.0 * zscale));
    p.percent = slope_percent ? 1 : 0;
    p.slope = slope; p.aspect = aspect; p.tri = tri; p.out_ld = out_ld;

    CUtensorMap map;
    const float *base = dem - (int64_t)p.halo_top * ld;
    int rc = ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base,
                             rows + p.halo_top + p.halo_bot, cols, ld, BW, BH, true);
    if (rc)
        return rc;
    rc = has_nodata ? dispatch<true>(map, p, st) : dispatch<false>(map, p, st);
    if (rc)
        return rc;
    if (has_nodata)
        terrain_corner_kernel<true><<<1, 4, 0, st>>>(dem, ld, p);
    else
        terrain_corner_kernel<false><<<1, 4, 0, st>>>(dem, ld, p);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
