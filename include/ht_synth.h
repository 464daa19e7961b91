/*
 * include/ht_synth.h -- synthetic conditioned fractal DEM, integer-exact.
 *
 * Workload generator for BASELINE.json configs 2-5 (SURVEY.md section 8(d)
 * "Synthetic DEM spec").  It is a pure integer function of (seed,row,col) so a
 * CUDA kernel (each GPU generating its own row band) and the CPU build used
 * by the tests produce the same bits.  Not part of the reference; the
 * reference ships no generator (its only DEM is tests/data-files/dem.tif).
 *
 * Construction (all integer):
 *   q(r,c)  = tilt*(R-1-r) + ((V*tri_q8(c + meander(r)) << 8) + noise_q16) >> 16
 *   z_mm    = 1000 + 9*q + (r%3)*3 + (c%3)            (< 4 000 000)
 *   dem     = (float)(z_mm / 1000.0)   so that round(dem*1000) == z_mm
 * - tie-free: the residue (r%3)*3+(c%3) is distinct for the 9 cells of any
 *   3x3 window and 9*q keeps residues from colliding;
 * - pit-free: tilt exceeds the bound on the per-row change of the valley and
 *   noise terms, so the cell below is always strictly lower.
 * `ht_d8_validate_conditioned` re-checks both on the device before timing.
 */
#ifndef HT_SYNTH_H
#define HT_SYNTH_H
#include <stdint.h>

#ifdef __CUDACC__
#define HT_HD __host__ __device__ __forceinline__
#else
#define HT_HD static inline
#endif

#define HT_SYNTH_OCTAVES 8 /* wavelengths 4 .. 512 cells */

typedef struct {
    uint32_t seed;
    int32_t tilt;        /* q-levels per row */
    int32_t valley;      /* V: q-levels per column toward the valley line */
    int32_t period_log2; /* valley spacing P = 1 << period_log2 columns */
    int32_t meander_q8;  /* meander amplitude (columns, Q8); 0 = straight */
    int32_t amp_q8[HT_SYNTH_OCTAVES]; /* noise amplitude per octave, Q8 */
    int64_t total_rows;  /* R of the whole (unsharded) grid */
    int32_t with_nulls;  /* 1: carve NULL blobs (tests only) */
} ht_synth_params;

HT_HD uint32_t ht_mix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

HT_HD uint32_t ht_hash3(uint32_t seed, uint32_t k, uint32_t i, uint32_t j)
{
    uint32_t h = ht_mix32(seed ^ (k * 0x9E3779B9U));
    h = ht_mix32(h + i * 0x85EBCA6BU);
    h = ht_mix32(h ^ (j * 0xC2B2AE35U));
    return h;
}

/* smoothstep weight, t in Q16 -> w in Q16 */
HT_HD int64_t ht_smooth_q16(int64_t t)
{
    return (t * t * (3 * 65536 - 2 * t)) >> 32;
}

/* 2-D value noise of wavelength (1<<sr) rows x (1<<sc) cols, result Q16 */
HT_HD int64_t ht_vnoise_q16(uint32_t seed, uint32_t k, int64_t r, int64_t c,
                            int sr, int sc)
{
    uint32_t i0 = (uint32_t)(r >> sr), j0 = (uint32_t)(c >> sc);
    int64_t wr = ht_smooth_q16((r & ((1LL << sr) - 1)) << (16 - sr));
    int64_t wc = ht_smooth_q16((c & ((1LL << sc) - 1)) << (16 - sc));
    int64_t v00 = ht_hash3(seed, k, i0, j0) >> 16;
    int64_t v01 = ht_hash3(seed, k, i0, j0 + 1) >> 16;
    int64_t v10 = ht_hash3(seed, k, i0 + 1, j0) >> 16;
    int64_t v11 = ht_hash3(seed, k, i0 + 1, j0 + 1) >> 16;
    int64_t top = (v00 * (65536 - wc) + v01 * wc) >> 16;
    int64_t bot = (v10 * (65536 - wc) + v11 * wc) >> 16;
    return (top * (65536 - wr) + bot * wr) >> 16;
}

/* integer millimetre elevation; row is the GLOBAL row index */
HT_HD int32_t ht_synth_zmm(const ht_synth_params *p, int64_t row, int64_t col)
{
    int64_t noise_q16 = 0;
    for (int s = 0; s < HT_SYNTH_OCTAVES; s++) {
        /* rows are stretched 16x: the vertical gradient is what pit-freeness
         * bounds, so spend the budget on cross-slope structure */
        int64_t v = ht_vnoise_q16(p->seed, 16 + s, row, col, s + 6, s + 2);
        noise_q16 += ((int64_t)p->amp_q8[s] * v) >> 8;
    }
    int64_t m_q8 = 0;
    if (p->meander_q8) {
        int64_t v = ht_vnoise_q16(p->seed, 7, row, 0, 11, 11); /* 2048 rows */
        m_q8 = ((int64_t)p->meander_q8 * v) >> 16;
    }
    int64_t period_q8 = 1LL << (p->period_log2 + 8);
    int64_t x = ((col << 8) + m_q8) & (period_q8 - 1);
    int64_t tri_q8 = x - (period_q8 >> 1);
    if (tri_q8 < 0)
        tri_q8 = -tri_q8;
    int64_t fine_q16 = (((int64_t)p->valley * tri_q8) << 8) + noise_q16;
    int64_t q = (int64_t)p->tilt * (p->total_rows - 1 - row) + (fine_q16 >> 16);
    return (int32_t)(1000 + 9 * q + (row % 3) * 3 + (col % 3));
}

HT_HD int ht_synth_is_null(const ht_synth_params *p, int64_t row, int64_t col)
{
    if (!p->with_nulls)
        return 0;
    int64_t v = ht_vnoise_q16(p->seed, 3, row, col, 5, 5);
    return v > 46000; /* roughly a tenth of the area, in 32-cell blobs */
}

HT_HD float ht_synth_dem_value(const ht_synth_params *p, int64_t row,
                               int64_t col)
{
    return (float)((double)ht_synth_zmm(p, row, col) / 1000.0);
}

/* precipitation-like weight, 1650.000 .. 1699.999 (reference fixture
 * tests/data-files/precip.tif spans 1658.77 .. 1700.42) */
HT_HD float ht_synth_weight_value(uint32_t seed, int64_t row, int64_t col)
{
    uint32_t h = ht_hash3(seed + 1U, 99, (uint32_t)row, (uint32_t)col);
    return (float)(1650000 + (int32_t)(h % 50000U)) / 1000.0f;
}

/* Parameter choice for a grid of total_rows rows: spend 80% of the 444k
 * q-level budget (z_mm < 4e6) on the regional tilt, bound the per-row change
 * of everything else below it.  Host-side, integer only. */
static inline void ht_synth_make_params(uint32_t seed, int64_t total_rows,
                                        int with_nulls, ht_synth_params *p)
{
    /* lambda^0.8 for lambda = 4..512, Q4 */
    static const int32_t pow08_q4[HT_SYNTH_OCTAVES] = {48,  84,  147, 256,
                                                       446, 776, 1351, 2353};
    int64_t budget = (4000000 - 1000 - 9) / 9;
    int64_t rows1 = total_rows > 1 ? total_rows - 1 : 1;
    int64_t tilt = (budget * 8 / 10) / rows1;
    if (tilt > 64)
        tilt = 64;
    if (tilt < 3)
        tilt = 3;
    p->seed = seed;
    p->tilt = (int32_t)tilt;
    p->total_rows = total_rows;
    p->with_nulls = with_nulls;
    p->period_log2 = 10;
    p->valley = (int32_t)tilt;
    /* meander: amplitude A columns over a 2048-row wavelength moves the
     * valley line by <= 1.5*A/2048 columns per row, i.e. V*1.5*A/2048
     * q-levels per row; A = 192 columns -> 0.141*V */
    int64_t g_valley_q8 = 0;
    p->meander_q8 = 0;
    if (tilt >= 8) {
        p->meander_q8 = 192 << 8;
        g_valley_q8 = ((int64_t)p->valley * 3 * 192 * 256) / (2 * 2048) + 256;
    }
    /* noise: vertical wavelength of octave s is 64<<s rows, so its per-row
     * change is <= 1.5*A_s/(64<<s).  Give the noise what is left of
     * tilt - 2 after the valley term. */
    int64_t g_left_q8 = (tilt - 2) * 256 - g_valley_q8 - 128;
    if (g_left_q8 < 0)
        g_left_q8 = 0;
    /* sum_s 1.5 * a*pow08[s] / (64<<s)  with a in Q8 */
    int64_t denom_q12 = 0; /* sum 1.5*pow08_q4[s]/(64<<s) in Q12 */
    for (int s = 0; s < HT_SYNTH_OCTAVES; s++)
        denom_q12 += (3 * (int64_t)pow08_q4[s] * 256) / (2 * (64LL << s));
    int64_t a_q8 = denom_q12 > 0 ? (g_left_q8 * 4096) / denom_q12 : 0;
    for (int s = 0; s < HT_SYNTH_OCTAVES; s++) {
        int64_t amp = (a_q8 * pow08_q4[s]) >> 4; /* Q8 */
        /* keep valley + noise within the remaining 20% of the budget */
        int64_t cap = (budget / 10 / HT_SYNTH_OCTAVES) << 8;
        if (amp > cap)
            amp = cap;
        p->amp_q8[s] = (int32_t)amp;
    }
}

#endif /* HT_SYNTH_H */
