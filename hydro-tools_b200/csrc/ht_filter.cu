// ht_filter.cu -- moving-window filter over an arbitrary boolean kernel.
//
// Replaces the numba kernel of hydrotools.interpolate.custom_raster_filter /
// raster_filter (/root/reference/src/hydrotools/interpolate.py:42-64 `apply_filter`
// under `da.map_overlap(depth, boundary=nan)`, :94-107; reducers :147-222) for the
// reducers sum, min, max, diff, mean, std, variance: for every non-NaN cell the
// reducer of the cells the kernel selects in the window
//     rows i - i_start .. i + i_end,  cols j - j_start .. j + j_end,
//     i_start = ceil((kh - 1) / 2), i_end = kh - 1 - i_start   (likewise for columns)
// with NaN outside the raster; sample and reducer in fp64 like the reference
// (np.nansum / nanmean / nanstd / nanvar on a float64 sample), result fp32, NaN
// where the centre is NaN or no selected cell holds data.  
//
// Kernel: one CTA per 32 x 128 output tile; the tile plus its halo (kernel edge
// <= 33 => halo <= 16) arrives as ONE TMA box, out-of-raster cells filled with NaN
// by the TMA unit -- the reference's `boundary=nan` for free.  256 threads, 16
// cells each (lane -> consecutive columns: conflict-free LDS), the set kernel cells
// as a table of shared-memory offsets in the kernel parameters (constant bank).
// 4 B read + 4 B written per cell; HBM-bound for small kernels, bound by the fp64
// adds / LDS for large ones (work = cells x set kernel cells).
#include <algorithm>
#include <new>

#include "ht_common.cuh"

namespace {

constexpr int TW = 128, TH = 32, THREADS = 256;
constexpr int MAXK = 33;  // largest kernel edge
constexpr int CELLS = TW * TH / THREADS;
enum { M_SUM = 0, M_MIN, M_MAX, M_DIFF, M_MEAN, M_STD, M_VAR, M_MEDIAN, M_MODE, M_COUNT };

struct FilterArgs {
    int64_t rows, cols;
    int halo_top;
    int tiles_x, tiles_y;
    int has_nodata;
    float nodata;
    int n_off;          // set kernel cells
    int bw, bh;         // TMA box
    int hx0a, hy0;      // left halo (rounded up to 4 columns), top halo
    float *out;
    int64_t out_ld;
    int16_t off[MAXK * MAXK]; // shared-memory offsets of the set kernel cells, relative to the centre
};

template <int METHOD>
__global__ void __launch_bounds__(THREADS, 2)
filter_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ FilterArgs p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    float *s = reinterpret_cast<float *>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int64_t x0 = (int64_t)tx * TW, y0 = (int64_t)ty * TH;
    if (tid == 0) {
        ht::tma_prefetch_desc(&map);
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, (uint32_t)(p.bw * p.bh * 4));
        ht::tma_load_2d(s, &map, (int)x0 - p.hx0a, (int)y0 - p.hy0 + p.halo_top, &bar);
    }
    __syncthreads();
    ht::mbar_wait(&bar, 0);
    if (p.has_nodata) { // the reference fills masked cells with NaN before filtering
        for (int i = tid; i < p.bw * p.bh; i += THREADS)
            if (s[i] == p.nodata)
                s[i] = __int_as_float(0x7fc00000);
        __syncthreads();
    }

    // thread -> rows 4 * warp + i, columns lane + 32 * j
    int c[CELLS];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
            c[4 * i + j] = (p.hy0 + 4 * warp + i) * p.bw + p.hx0a + lane + 32 * j;

    float res[CELLS];
    if (METHOD == M_MEDIAN) {
        // np.nanmedian (interpolate.py:215-219): the middle value of the valid samples, the
        // mean of the two middle ones (fp64) for an even count.  No sample is stored: the
        // r-th smallest key is found bit by bit -- the largest v with #{key < v} <= r -- in 32
        // counting passes over the window (keys = floats mapped to unsigned in their order).
        auto key_of = [](float v) -> unsigned {
            const unsigned b = __float_as_uint(v);
            return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
        };
        int n[CELLS];
        unsigned pre[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            n[k] = 0;
            pre[k] = 0u;
        }
        for (int o = 0; o < p.n_off; o++) {
            const int d = p.off[o];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                const float v = s[c[k] + d];
                n[k] += v == v;
            }
        }
        for (int bit = 31; bit >= 0; bit--) {
            int cnt[CELLS];
#pragma unroll
            for (int k = 0; k < CELLS; k++)
                cnt[k] = 0;
            for (int o = 0; o < p.n_off; o++) {
                const int d = p.off[o];
#pragma unroll
                for (int k = 0; k < CELLS; k++) {
                    const float v = s[c[k] + d];
                    cnt[k] += (v == v) && key_of(v) < (pre[k] | (1u << bit));
                }
            }
#pragma unroll
            for (int k = 0; k < CELLS; k++)
                if (n[k] && cnt[k] <= (n[k] - 1) / 2)
                    pre[k] |= 1u << bit;
        }
        // pre = key of the lower middle value; the upper one is the same value if enough
        // samples are <= it, else the smallest key above it
        int le[CELLS];
        unsigned nxt[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            le[k] = 0;
            nxt[k] = 0xFFFFFFFFu;
        }
        for (int o = 0; o < p.n_off; o++) {
            const int d = p.off[o];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                const float v = s[c[k] + d];
                if (v == v) {
                    const unsigned q = key_of(v);
                    le[k] += q <= pre[k];
                    if (q > pre[k])
                        nxt[k] = min(nxt[k], q);
                }
            }
        }
        auto val_of = [](unsigned q) -> float {
            return __uint_as_float(q ^ ((q >> 31) ? 0x80000000u : 0xFFFFFFFFu));
        };
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            if (!n[k]) {
                res[k] = __int_as_float(0x7fc00000);
                continue;
            }
            const float lo = val_of(pre[k]);
            const float hi = (n[k] & 1) || le[k] > n[k] / 2 ? lo : val_of(nxt[k]);
            res[k] = (n[k] & 1) ? lo : (float)(((double)lo + (double)hi) * 0.5);
        }
    }
    else if (METHOD == M_MODE) {
        // the reference counts the values in a dictionary in scan order and keeps the first
        // key whose count is strictly the largest (interpolate.py:174-201): the first value,
        // in window scan order, of maximal multiplicity
        int best[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            best[k] = 0;
            res[k] = __int_as_float(0x7fc00000);
        }
        for (int o = 0; o < p.n_off; o++) {
            const int d = p.off[o];
            float cand[CELLS];
            int cnt[CELLS];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                cand[k] = s[c[k] + d];
                cnt[k] = 0;
            }
            for (int o2 = 0; o2 < p.n_off; o2++) {
                const int d2 = p.off[o2];
#pragma unroll
                for (int k = 0; k < CELLS; k++)
                    cnt[k] += s[c[k] + d2] == cand[k]; // false for NaN
            }
#pragma unroll
            for (int k = 0; k < CELLS; k++)
                if (cnt[k] > best[k]) {
                    best[k] = cnt[k];
                    res[k] = cand[k];
                }
        }
    }
    else if (METHOD == M_MIN || METHOD == M_MAX || METHOD == M_DIFF) {
        float mn[CELLS], mx[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            mn[k] = __int_as_float(0x7f800000);
            mx[k] = __int_as_float(0xff800000);
        }
        for (int o = 0; o < p.n_off; o++) {
            const int d = p.off[o];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                const float v = s[c[k] + d];
                mn[k] = fminf(mn[k], v); // fminf / fmaxf ignore NaN
                mx[k] = fmaxf(mx[k], v);
            }
        }
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            const bool any = mx[k] >= mn[k];
            const float r = METHOD == M_MIN ? mn[k] : METHOD == M_MAX ? mx[k]
                          : (float)((double)mx[k] - (double)mn[k]);
            res[k] = any ? r : __int_as_float(0x7fc00000);
        }
    }
    else {
        double sum[CELLS];
        int n[CELLS];
#pragma unroll
        for (int k = 0; k < CELLS; k++) {
            sum[k] = 0.0;
            n[k] = 0;
        }
        for (int o = 0; o < p.n_off; o++) {
            const int d = p.off[o];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                const float v = s[c[k] + d];
                if (v == v) {
                    sum[k] += (double)v;
                    n[k]++;
                }
            }
        }
        if (METHOD == M_SUM || METHOD == M_MEAN) {
#pragma unroll
            for (int k = 0; k < CELLS; k++)
                res[k] = n[k] ? (float)(METHOD == M_SUM ? sum[k] : sum[k] / (double)n[k])
                              : __int_as_float(0x7fc00000);
        }
        else { // two passes like np.nanstd / np.nanvar: mean, then squared deviations
            double ss[CELLS];
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                sum[k] = n[k] ? sum[k] / (double)n[k] : 0.0; // mean
                ss[k] = 0.0;
            }
            for (int o = 0; o < p.n_off; o++) {
                const int d = p.off[o];
#pragma unroll
                for (int k = 0; k < CELLS; k++) {
                    const float v = s[c[k] + d];
                    if (v == v) {
                        const double e = (double)v - sum[k];
                        ss[k] = fma(e, e, ss[k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < CELLS; k++) {
                const double var = n[k] ? ss[k] / (double)n[k] : 0.0;
                res[k] = n[k] ? (float)(METHOD == M_VAR ? var : sqrt(var)) : __int_as_float(0x7fc00000);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int64_t gr = y0 + 4 * warp + i;
        if (gr >= p.rows)
            break;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t gc = x0 + lane + 32 * j;
            if (gc >= p.cols)
                continue;
            const float centre = s[c[4 * i + j]];
            p.out[gr * p.out_ld + gc] = centre == centre ? res[4 * i + j] : __int_as_float(0x7fc00000);
        }
    }
}


// ---------------------------------------------------------------------------
// Large kernels, sum / mean: row prefix sums + one difference per kernel-row run.
// The reference's callers filter with discs built from a distance
// (kernel_from_distance, /root/reference/src/hydrotools/utils.py:408-426;
// morphology.py:282-289 "mean", :1373-1374 "sum"), i.e. with hundreds to tens of
// thousands of samples per cell.  With P[y][x] = sum of the valid values of row y left
// of column x (fp64) and C[y][x] = their number, a run [a, b] of selected cells in
// kernel row ky contributes P[y][x + b + 1] - P[y][x + a] -- two loads instead of
// b - a + 1, and no limit on the kernel size.  Columns are clamped to the raster, which
// is the reference's NaN boundary (nothing to add outside).
// ---------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4;

// one CTA per row: exclusive prefix of the valid values / their count, cols + 1 entries
__global__ void __launch_bounds__(SCAN_THREADS)
row_prefix_kernel(const float *src, int64_t ld, int64_t rows_all, int64_t cols, int has_nodata,
                  float nodata, double *P, unsigned *Cn, int64_t pld)
{
    __shared__ double wsum[SCAN_THREADS / 32];
    __shared__ unsigned wcnt[SCAN_THREADS / 32];
    __shared__ double carry_s;
    __shared__ unsigned carry_c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t y = blockIdx.x; y < rows_all; y += gridDim.x) {
        const float *row = src + y * ld;
        double *Pr = P + y * pld;
        unsigned *Cr = Cn + y * pld;
        if (tid == 0) {
            carry_s = 0.0;
            carry_c = 0u;
            Pr[0] = 0.0;
            Cr[0] = 0u;
        }
        __syncthreads();
        for (int64_t x0 = 0; x0 < cols; x0 += SCAN_THREADS * SCAN_ITEMS) {
            double v[SCAN_ITEMS];
            unsigned c[SCAN_ITEMS];
            double ts = 0.0;
            unsigned tc = 0u;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) {
                const int64_t x = x0 + (int64_t)tid * SCAN_ITEMS + k;
                float f = x < cols ? row[x] : __int_as_float(0x7fc00000);
                const bool ok = f == f && !(has_nodata && f == nodata);
                ts += ok ? (double)f : 0.0;
                tc += ok ? 1u : 0u;
                v[k] = ts; // inclusive within the thread
                c[k] = tc;
            }
            // exclusive scan of the per-thread totals across the CTA
            double ws = ts;
            unsigned wc = tc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double s2 = __shfl_up_sync(0xffffffffu, ws, o);
                const unsigned c2 = __shfl_up_sync(0xffffffffu, wc, o);
                if (lane >= o) {
                    ws += s2;
                    wc += c2;
                }
            }
            if (lane == 31) {
                wsum[warp] = ws;
                wcnt[warp] = wc;
            }
            __syncthreads();
            double base = carry_s;
            unsigned cbase = carry_c;
            for (int w = 0; w < warp; w++) {
                base += wsum[w];
                cbase += wcnt[w];
            }
            base += ws - ts; // exclusive prefix of this thread
            cbase += wc - tc;
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; k++) {
                const int64_t x = x0 + (int64_t)tid * SCAN_ITEMS + k;
                if (x < cols) {
                    Pr[x + 1] = base + v[k];
                    Cr[x + 1] = cbase + c[k];
                }
            }
            __syncthreads();
            if (tid == SCAN_THREADS - 1) {
                carry_s = base + ts;
                carry_c = cbase + tc;
            }
            __syncthreads();
        }
    }
}

struct Run { int16_t ky, a, b; }; // kernel row, first / last selected column (kernel coordinates)

// thread -> one output cell; lanes along x.  P / Cn cover rows -halo_top .. rows + halo_bot.
__global__ void __launch_bounds__(256)
run_filter_kernel(const float *src, int64_t ld, int64_t rows, int64_t cols, int halo_top,
                  int halo_bot, int has_nodata, float nodata, const double *P, const unsigned *Cn,
                  int64_t pld, const Run *runs, int n_runs, int i_start, int j_start, int mean,
                  float *out, int64_t out_ld)
{
    extern __shared__ Run srun[];
    for (int i = threadIdx.x; i < n_runs; i += blockDim.x)
        srun[i] = runs[i];
    __syncthreads();
    const int64_t x = (int64_t)blockIdx.x * 128 + (threadIdx.x & 127);
    const int64_t y0 = ((int64_t)blockIdx.y * 2 + (threadIdx.x >> 7)) * 16;
    if (x >= cols)
        return;
    for (int64_t y = y0; y < min(y0 + 16, rows); y++) {
        const float centre = src[y * ld + x];
        double sum = 0.0;
        unsigned n = 0u;
        for (int r = 0; r < n_runs; r++) {
            const Run q = srun[r];
            const int64_t yy = y + q.ky - i_start;
            if (yy < -(int64_t)halo_top || yy >= rows + halo_bot)
                continue;
            const int64_t xa = min(max(x + q.a - j_start, (int64_t)0), cols);
            const int64_t xb = min(max(x + q.b - j_start + 1, (int64_t)0), cols);
            const int64_t at = (yy + halo_top) * pld;
            sum += P[at + xb] - P[at + xa];
            n += Cn[at + xb] - Cn[at + xa];
        }
        const bool ok = centre == centre && !(has_nodata && centre == nodata) && n;
        out[y * out_ld + x] = ok ? (float)(mean ? sum / (double)n : sum) : __int_as_float(0x7fc00000);
    }
}

// The same sums with the prefix rows of one output tile staged in shared memory: every P / Cn
// entry is read from L2 once per tile instead of once per kernel row that touches it (a disc
// of radius 10 reads each entry ~20 times), and the clamping of the run ends to the raster is
// done while staging.  Tile = th rows x 128 columns; staged block = (th + kh - 1) x (128 + kw)
// entries of 12 B.  Used when that fits shared memory; larger kernels keep the kernel above.
__global__ void __launch_bounds__(256)
run_filter_tiled_kernel(const float *src, int64_t ld, int64_t rows, int64_t cols, int halo_top,
                        int halo_bot, int has_nodata, float nodata, const double *P, const unsigned *Cn,
                        int64_t pld, const Run *runs, int n_runs, int i_start, int j_start, int mean,
                        float *out, int64_t out_ld, int th, int R, int C)
{
    extern __shared__ __align__(16) unsigned char fsm[];
    double *sP = reinterpret_cast<double *>(fsm);
    unsigned *sC = reinterpret_cast<unsigned *>(fsm + (size_t)R * C * 8);
    Run *srun = reinterpret_cast<Run *>(fsm + (size_t)R * C * 12);
    const int64_t x0 = (int64_t)blockIdx.x * 128, y0 = (int64_t)blockIdx.y * th;
    for (int i = threadIdx.x; i < n_runs; i += 256)
        srun[i] = runs[i];
    for (int idx = threadIdx.x; idx < R * C; idx += 256) {
        const int rr = idx / C, cc = idx - rr * C;
        const int64_t yy = y0 - i_start + rr;
        double pv = 0.0;
        unsigned cv = 0u;
        if (yy >= -(int64_t)halo_top && yy < rows + halo_bot) {
            const int64_t X = min(max(x0 - j_start + cc, (int64_t)0), cols);
            const int64_t at = (yy + halo_top) * pld + X;
            pv = P[at];
            cv = Cn[at];
        }
        sP[idx] = pv;
        sC[idx] = cv;
    }
    __syncthreads();
    const int lx = threadIdx.x & 127;
    const int64_t x = x0 + lx;
    if (x >= cols)
        return;
    const int half = th / 2, ly0 = (threadIdx.x >> 7) * half;
    for (int ly = ly0; ly < ly0 + half; ly++) {
        const int64_t y = y0 + ly;
        if (y >= rows)
            break;
        double sum = 0.0;
        unsigned n = 0u;
        for (int r = 0; r < n_runs; r++) {
            const Run q = srun[r];
            const int at = (ly + q.ky) * C + lx;
            sum += sP[at + q.b + 1] - sP[at + q.a];
            n += sC[at + q.b + 1] - sC[at + q.a];
        }
        const float centre = src[y * ld + x];
        const bool ok = centre == centre && !(has_nodata && centre == nodata) && n;
        out[y * out_ld + x] = ok ? (float)(mean ? sum / (double)n : sum) : __int_as_float(0x7fc00000);
    }
}

// Self-contained variant for kernels whose staged block fits shared memory (discs up to a
// radius of ~25): the CTA reads the RAW tile plus halo (4 B per staged cell instead of 12 B of
// global prefix entries, and no prefix pass over the raster at all), builds the row prefix
// sums LOCALLY -- a run difference does not care where the prefix starts -- and, when no
// staged cell is invalid and the block lies inside the raster, skips the counts: n is then
// the number of selected cells.  Local prefixes are also shorter sums, i.e. more accurate.
constexpr int MAX_LOCAL_RUNS = 128;
struct LocalRuns {
    int n, cells;                 // runs, selected cells
    int off_a[MAX_LOCAL_RUNS];    // ky * C + a       (prefix sums, pitch C)
    int off_b[MAX_LOCAL_RUNS];    // ky * C + b + 1
    int cnt_a[MAX_LOCAL_RUNS];    // ky * bw + a      (counts, pitch bw)
    int cnt_b[MAX_LOCAL_RUNS];    // ky * bw + b + 1
};

__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(uint32_t addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int TH_, int KU>
__global__ void __launch_bounds__(256)
run_filter_local_kernel(const __grid_constant__ CUtensorMap map, const float *src, int64_t ld,
                        int64_t rows, int64_t cols, int halo_top, int has_nodata, float nodata,
                        const __grid_constant__ LocalRuns runs, int i_start, int shift, int hx0a,
                        int mean, float *out, int64_t out_ld, int R, int C, int bw)
{
    constexpr int th = TH_, HALF = TH_ / 2;
    // sP: R x C doubles.  Behind it the raw box (R x bw floats, one TMA load, NaN outside the
    // raster); the counts of a row overwrite that row's raw values once the row is scanned.
    // (a name of its own: the 16-byte alignment of the other kernel's `fsm` would win)
    extern __shared__ __align__(128) unsigned char lsm[];
    __shared__ __align__(8) uint64_t bar;
    float *raw = reinterpret_cast<float *>(lsm);
    unsigned *sC = reinterpret_cast<unsigned *>(lsm);
    double *sP = reinterpret_cast<double *>(lsm + (((size_t)R * bw * 4 + 127) / 128) * 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t x0 = (int64_t)blockIdx.x * 128, y0 = (int64_t)blockIdx.y * th;
    if (tid == 0) {
        ht::tma_prefetch_desc(&map);
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, (uint32_t)(R * bw * 4));
        ht::tma_load_2d(raw, &map, (int)x0 - hx0a, (int)y0 - i_start + halo_top, &bar);
    }
    __syncthreads();
    ht::mbar_wait(&bar, 0);
    // ---- row prefixes over the C - 1 raw cells of a row: warp -> rows, lane -> K consecutive cells.
    // +-inf are data (np.nansum gives inf for a window that holds one, NaN for both signs) but
    // must stay out of the prefix SUMS, or every difference to their right would be inf - inf:
    // they are counted instead, in the same word as the valid cells (bits 0..11 valid cells,
    // 12..21 +inf, 22..31 -inf; prefixes are monotone per field, so differences never borrow).
    const int K = (C - 1 + 31) / 32; // <= KU
    bool dirty = false;
    for (int rr = warp; rr < R; rr += 8) {
        float f[KU];
        double run = 0.0;
        unsigned cnt = 0u;
#pragma unroll
        for (int k = 0; k < KU; k++) {
            const int cc = lane * K + k;
            f[k] = __int_as_float(0x7fc00000);
            if (k < K && cc < C - 1) {
                f[k] = raw[rr * bw + shift + cc];
                const bool ok = f[k] == f[k] && !(has_nodata && f[k] == nodata);
                const bool fin = fabsf(f[k]) <= 3.4028234e38f;
                if (!ok)
                    f[k] = __int_as_float(0x7fc00000);
                dirty |= !ok || !fin;
                run += ok && fin ? (double)f[k] : 0.0;
                cnt += ok ? (fin ? 1u : (f[k] > 0.0f ? 0x1001u : 0x400001u)) : 0u;
            }
        }
        double ws = run;
        unsigned wc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double s2 = __shfl_up_sync(0xffffffffu, ws, o);
            const unsigned c2 = __shfl_up_sync(0xffffffffu, wc, o);
            if (lane >= o) {
                ws += s2;
                wc += c2;
            }
        }
        double base = ws - run; // exclusive prefix of the lane
        unsigned cbase = wc - cnt;
        __syncwarp(); // every lane has read its raw cells: the row now becomes the counts
        if (lane == 0) {
            sP[rr * C] = 0.0;
            sC[rr * bw] = 0u;
        }
#pragma unroll
        for (int k = 0; k < KU; k++) {
            const int cc = lane * K + k;
            if (k < K && cc < C - 1) {
                if (f[k] == f[k]) {
                    const bool fin = fabsf(f[k]) <= 3.4028234e38f;
                    if (fin)
                        base += (double)f[k];
                    cbase += fin ? 1u : (f[k] > 0.0f ? 0x1001u : 0x400001u);
                }
                sP[rr * C + cc + 1] = base;
                sC[rr * bw + cc + 1] = cbase;
            }
        }
    }
    const bool clean = !__syncthreads_or(dirty);
    const int lx = tid & 127;
    const int64_t x = x0 + lx;
    if (x >= cols)
        return;
    // thread -> one column, HALF consecutive rows; runs in the OUTER loop, so that a run's
    // offsets are fetched once for the thread's rows.  32-bit shared addresses, stepped by
    // the row pitch: two loads, two adds and two fp64 adds per cell and run.
    const int ly0 = (tid >> 7) * HALF;
    double sum[HALF];
    unsigned n[HALF];
#pragma unroll
    for (int c = 0; c < HALF; c++) {
        sum[c] = 0.0;
        n[c] = clean ? (unsigned)runs.cells : 0u;
    }
    const uint32_t aP = ht::smem_u32(sP) + 8u * (unsigned)(ly0 * C + lx), stepP = 8u * (unsigned)C;
    const uint32_t aC = ht::smem_u32(sC) + 4u * (unsigned)(ly0 * bw + lx), stepC = 4u * (unsigned)bw;
    if (clean) {
        for (int r = 0; r < runs.n; r++) {
            uint32_t pb = aP + 8u * (unsigned)runs.off_b[r], pa = aP + 8u * (unsigned)runs.off_a[r];
#pragma unroll
            for (int c = 0; c < HALF; c++, pb += stepP, pa += stepP)
                sum[c] += lds_f64(pb) - lds_f64(pa);
        }
    } else {
        for (int r = 0; r < runs.n; r++) {
            uint32_t pb = aP + 8u * (unsigned)runs.off_b[r], pa = aP + 8u * (unsigned)runs.off_a[r];
            uint32_t cb = aC + 4u * (unsigned)runs.cnt_b[r], ca = aC + 4u * (unsigned)runs.cnt_a[r];
#pragma unroll
            for (int c = 0; c < HALF; c++, pb += stepP, pa += stepP, cb += stepC, ca += stepC) {
                sum[c] += lds_f64(pb) - lds_f64(pa);
                n[c] += lds_u32(cb) - lds_u32(ca);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < HALF; c++) {
        const int64_t y = y0 + ly0 + c;
        if (y >= rows)
            break;
        const float centre = src[y * ld + x];
        const unsigned nv = n[c] & 0xFFFu, np = (n[c] >> 12) & 0x3FFu, nn = n[c] >> 22;
        const bool ok = centre == centre && !(has_nodata && centre == nodata) && nv;
        float r = (float)(mean ? sum[c] / (double)nv : sum[c]);
        if (np | nn) // the window holds +-inf: inf, or NaN for both signs (np.nansum / nanmean)
            r = (np && nn) ? __int_as_float(0x7fc00000) : __int_as_float(np ? 0x7f800000 : 0xff800000);
        out[y * out_ld + x] = ok ? r : __int_as_float(0x7fc00000);
    }
}

template <int METHOD>
int launch(const CUtensorMap &map, const FilterArgs &p, cudaStream_t st)
{
    auto kern = filter_kernel<METHOD>;
    const int smem = p.bw * p.bh * 4;
    HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<p.tiles_x * p.tiles_y, THREADS, smem, st>>>(map, p);
    HT_LAUNCH_CHECK();
    return 0;
}

} // namespace

// kernel_host: kh x kw bytes in HOST memory (non-zero = cell selected), kh, kw <= 33.
// method: 0 sum, 1 min, 2 max, 3 diff, 4 mean, 5 std, 6 variance, 7 median, 8 mode.
// halo_top / halo_bot: valid rows of the neighbour bands directly above / below
// `src` in the same allocation (0 = that side is the edge of the raster, NaN beyond).
extern "C" int ht_window_filter_f32(const float *src, int64_t rows, int64_t cols, int64_t ld,
                                    int has_nodata, float nodata,
                                    const unsigned char *kernel_host, int kh, int kw, int method,
                                    float *out, int64_t out_ld, int halo_top, int halo_bot,
                                    void *stream)
{
    if (!src || !out || !kernel_host || rows < 0 || cols < 0 || ld < cols || out_ld < cols ||
        kh < 1 || kw < 1 || kh > MAXK || kw > MAXK || method < 0 || method >= M_COUNT ||
        halo_top < 0 || halo_bot < 0)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    FilterArgs p;
    const int i_start = kh / 2, i_end = kh - 1 - i_start; // ceil((kh - 1) / 2) == kh / 2
    const int j_start = kw / 2, j_end = kw - 1 - j_start;
    p.rows = rows; p.cols = cols;
    p.halo_top = halo_top;
    p.tiles_x = ht::cdiv(cols, TW); p.tiles_y = ht::cdiv(rows, TH);
    p.has_nodata = (has_nodata && nodata == nodata) ? 1 : 0;
    p.nodata = nodata;
    p.hy0 = i_start;
    p.hx0a = (j_start + 3) / 4 * 4;
    p.bw = TW + p.hx0a + (j_end + 3) / 4 * 4;
    p.bh = TH + i_start + i_end;
    p.out = out; p.out_ld = out_ld;
    p.n_off = 0;
    for (int ki = 0; ki < kh; ki++)
        for (int kj = 0; kj < kw; kj++)
            if (kernel_host[ki * kw + kj])
                p.off[p.n_off++] = (int16_t)((ki - i_start) * p.bw + (kj - j_start));
    CUtensorMap map;
    const float *base = src - (int64_t)halo_top * ld;
    int rc = ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base,
                             rows + halo_top + halo_bot, cols, ld, p.bw, p.bh, true);
    if (rc)
        return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (method) {
    case M_SUM: return launch<M_SUM>(map, p, st);
    case M_MIN: return launch<M_MIN>(map, p, st);
    case M_MAX: return launch<M_MAX>(map, p, st);
    case M_DIFF: return launch<M_DIFF>(map, p, st);
    case M_MEAN: return launch<M_MEAN>(map, p, st);
    case M_STD: return launch<M_STD>(map, p, st);
    case M_VAR: return launch<M_VAR>(map, p, st);
    case M_MEDIAN: return launch<M_MEDIAN>(map, p, st);
    case M_MODE: return launch<M_MODE>(map, p, st);
    }
    return HT_ERR_ARG;
}

// Scratch ht_window_filter_runs_f32 needs: row prefix sums (fp64) and counts (u32) of the
// rows it is given, cols + 1 entries per row (pitch rounded up to 2), plus the run table.
extern "C" size_t ht_window_filter_runs_workspace_bytes(int64_t rows_with_halos, int64_t cols,
                                                        int kh)
{
    if (rows_with_halos <= 0 || cols <= 0 || kh <= 0)
        return 256;
    const int64_t pld = (cols + 2) / 2 * 2;
    return (size_t)rows_with_halos * pld * 12 + (size_t)kh * 64 * sizeof(Run) + 1024;
}

// sum (method 0) / mean (method 4) over a boolean kernel of ANY size by row prefix sums.
// At most 64 runs per kernel row (a disc has one).  ws: device scratch, see above.
extern "C" int ht_window_filter_runs_f32(const float *src, int64_t rows, int64_t cols, int64_t ld,
                                         int has_nodata, float nodata,
                                         const unsigned char *kernel_host, int kh, int kw,
                                         int method, float *out, int64_t out_ld, int halo_top,
                                         int halo_bot, void *ws, size_t ws_bytes, void *stream)
{
    if (!src || !out || !kernel_host || !ws || rows < 0 || cols < 0 || ld < cols || out_ld < cols ||
        kh < 1 || kw < 1 || kh > 32767 || kw > 32767 || (method != M_SUM && method != M_MEAN) ||
        halo_top < 0 || halo_bot < 0)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    const int64_t rows_all = rows + halo_top + halo_bot;
    if (ws_bytes < ht_window_filter_runs_workspace_bytes(rows_all, cols, kh))
        return HT_ERR_WORKSPACE;
    // runs of selected cells per kernel row
    Run *hruns = new (std::nothrow) Run[(size_t)kh * 64];
    if (!hruns)
        return HT_ERR_ARG;
    int n_runs = 0;
    for (int ki = 0; ki < kh; ki++) {
        int in_row = 0;
        for (int kj = 0; kj < kw;) {
            if (!kernel_host[(size_t)ki * kw + kj]) {
                kj++;
                continue;
            }
            int e = kj;
            while (e + 1 < kw && kernel_host[(size_t)ki * kw + e + 1])
                e++;
            if (++in_row > 64) {
                delete[] hruns;
                return HT_ERR_ARG;
            }
            hruns[n_runs].ky = (int16_t)ki; hruns[n_runs].a = (int16_t)kj; hruns[n_runs].b = (int16_t)e;
            n_runs++;
            kj = e + 1;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    // kernels whose staged block fits shared memory: local prefixes, no pass over the raster
    int n_cells = 0;
    for (int r = 0; r < n_runs; r++)
        n_cells += hruns[r].b - hruns[r].a + 1;
    // (the count word of the local kernel has 12 bits for the cells of a window)
    if (n_runs <= MAX_LOCAL_RUNS && kw <= 100 && n_cells <= 4095) {
        const int i_start = kh / 2, j_start = kw / 2, j_end = kw - 1 - j_start;
        const int hx0a = (j_start + 3) / 4 * 4; // the TMA box starts on a 16-byte boundary
        for (int th = 32; th >= 8; th >>= 1) {
            const int R = th + kh - 1, C = 128 + kw;
            const int bw = (std::max(128 + hx0a + (j_end + 3) / 4 * 4, C) + 3) / 4 * 4;
            const size_t need = ((size_t)R * bw * 4 + 127) / 128 * 128 + (size_t)R * C * 8;
            if (need > (th == 32 ? 110u : 200u) * 1024u || bw > 256 || R > 256)
                continue;
            CUtensorMap map;
            if (ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, src - (int64_t)halo_top * ld,
                                rows_all, cols, ld, bw, R, true))
                break; // rows not 16-byte aligned: the prefix path below takes any layout
            LocalRuns lr;
            lr.n = n_runs;
            lr.cells = 0;
            for (int r = 0; r < n_runs; r++) {
                lr.off_a[r] = hruns[r].ky * C + hruns[r].a;
                lr.off_b[r] = hruns[r].ky * C + hruns[r].b + 1;
                lr.cnt_a[r] = hruns[r].ky * bw + hruns[r].a;
                lr.cnt_b[r] = hruns[r].ky * bw + hruns[r].b + 1;
                lr.cells += hruns[r].b - hruns[r].a + 1;
            }
            delete[] hruns;
            const bool k5 = (C - 1 + 31) / 32 <= 5; // cells per lane in the row scans
            auto kern = th == 32 ? (k5 ? run_filter_local_kernel<32, 5> : run_filter_local_kernel<32, 8>)
                        : th == 16 ? (k5 ? run_filter_local_kernel<16, 5> : run_filter_local_kernel<16, 8>)
                                   : (k5 ? run_filter_local_kernel<8, 5> : run_filter_local_kernel<8, 8>);
            HT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            dim3 tgrid((unsigned)ht::cdiv(cols, 128), (unsigned)ht::cdiv(rows, th));
            kern<<<tgrid, 256, need, st>>>(map, src, ld, rows, cols, halo_top,
                                          has_nodata && nodata == nodata, nodata, lr, i_start,
                                          hx0a - j_start, hx0a, method == M_MEAN, out, out_ld, R, C, bw);
            HT_LAUNCH_CHECK();
            return HT_OK;
        }
    }
    const int64_t pld = (cols + 2) / 2 * 2;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    double *P = (double *)b;
    unsigned *Cn = (unsigned *)(b + (size_t)rows_all * pld * 8);
    Run *druns = (Run *)(b + (size_t)rows_all * pld * 12);
    cudaError_t e = cudaMemcpyAsync(druns, hruns, (size_t)n_runs * sizeof(Run), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(st); // hruns is about to go away
    delete[] hruns;
    if (e != cudaSuccess)
        return (int)e;
    const float *base = src - (int64_t)halo_top * ld;
    row_prefix_kernel<<<(unsigned)std::min<int64_t>(rows_all, 148 * 8), SCAN_THREADS, 0, st>>>(
        base, ld, rows_all, cols, has_nodata && nodata == nodata, nodata, P, Cn, pld);
    HT_LAUNCH_CHECK();
    const int i_start = kh / 2, j_start = kw / 2;
    // tile rows: the largest of 32 / 16 / 8 whose staged prefix block fits (two CTAs per SM if possible)
    for (int th = 32; th >= 8; th >>= 1) {
        const int R = th + kh - 1, C = 128 + kw;
        const size_t need = (size_t)R * C * 12 + (size_t)n_runs * sizeof(Run) + 16;
        if (need > (th == 32 ? 110u : 200u) * 1024u)
            continue;
        HT_CUDA_TRY(cudaFuncSetAttribute(run_filter_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)need));
        dim3 tgrid((unsigned)ht::cdiv(cols, 128), (unsigned)ht::cdiv(rows, th));
        run_filter_tiled_kernel<<<tgrid, 256, need, st>>>(src, ld, rows, cols, halo_top, halo_bot,
                                                          has_nodata && nodata == nodata, nodata, P, Cn, pld,
                                                          druns, n_runs, i_start, j_start, method == M_MEAN,
                                                          out, out_ld, th, R, C);
        HT_LAUNCH_CHECK();
        return HT_OK;
    }
    const size_t smem = (size_t)n_runs * sizeof(Run);
    if (smem > 48 * 1024)
        HT_CUDA_TRY(cudaFuncSetAttribute(run_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ht::cdiv(cols, 128), (unsigned)ht::cdiv(rows, 32));
    run_filter_kernel<<<grid, 256, smem, st>>>(src, ld, rows, cols, halo_top, halo_bot,
                                              has_nodata && nodata == nodata, nodata, P, Cn, pld,
                                              druns, n_runs, i_start, j_start, method == M_MEAN, out,
                                              out_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
