/*
 * oracle/gdaldem.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of GDAL 3.11 `gdaldem {slope,aspect,TRI} -compute_edges`
 * (apps/gdaldem_lib.cpp: GDALGeneric3x3Processing<float>, ComputeVal,
 * GDALSlopeHornAlg, GDALAspectAlg, GDALTRIAlgRiley), which is what the
 * reference shells out to:
 *   /root/reference/src/hydrotools/elevation.py:41-54   slope  (-s scale, -p)
 *   /root/reference/src/hydrotools/elevation.py:70-78   aspect
 *   /root/reference/src/hydrotools/elevation.py:118-126 TRI
 * GDAL is not vendored under /root/reference (Dockerfile:1 pins
 * ghcr.io/osgeo/gdal:ubuntu-small-3.11.0) and is absent from this image, so
 * the algorithm is restated from the published source.  PARITY UNPINNED: the
 * reference has no test for these three functions.
 *
 * Rules:
 *  G1 window   w0..w8 row-major, w4 centre; Float32 source keeps the window
 *              as float and the bracketed sums are float arithmetic.
 *  G2 nodata   ARE_REAL_EQUAL(v, nodata) (== or within 2 float ulps); centre
 *              nodata -> -9999; with -compute_edges any other nodata window
 *              value is replaced by the centre value.
 *  G3 edges    first/last line: missing line = 2*a - b of the two nearest
 *              lines (nodata if either is), column index CLAMPED at j==0 and
 *              j==n-1; other lines, first/last column: missing column =
 *              2*a - b of the two nearest columns.  Any dimension < 2 -> all
 *              -9999.
 *  G4 slope    Horn; degrees atan(sqrt(dx^2+dy^2)/(8*scale))*180/pi or
 *              percent 100*sqrt()/(8*scale).
 *  G5 aspect   atan2(dy,-dx) in degrees; flat -> -9999; azimuth conversion;
 *              360 -> 0.
 *  G6 TRI      Riley (GDAL >= 3.3 default): sqrt(sum (w_k - w_4)^2).
 */
#include <math.h>
#include <float.h>
#include <pthread.h>
#include <stdint.h>

#include "oracle.h"

#define DST_NODATA (-9999.0f)

static int ora_threads = 1;
void ora_set_threads(int n) { ora_threads = n > 0 ? n : 1; }

static int are_real_equal(float a, float b)
{
    return a == b || fabsf(a - b) < FLT_EPSILON * fabsf(a + b) * 2;
}

typedef struct {
    int has_nodata;
    float nodata;
    int nodata_is_nan;
} nd_t;

static int is_nd(const nd_t *nd, float v)
{
    if (!nd->has_nodata)
        return 0;
    return nd->nodata_is_nan ? isnan(v) : are_real_equal(v, nd->nodata);
}

static float interpol(const nd_t *nd, float a, float b)
{
    if (nd->has_nodata &&
        (are_real_equal(a, nd->nodata) || are_real_equal(b, nd->nodata)))
        return nd->nodata;
    return 2 * a - b;
}

static const double RAD2DEG = 180.0 / 3.14159265358979323846;

static float slope_horn(const float *w, double ewres, double nsres,
                        double scale, int percent)
{
    const double dx =
        ((w[0] + w[3] + w[3] + w[6]) - (w[2] + w[5] + w[5] + w[8])) / ewres;
    const double dy =
        ((w[6] + w[7] + w[7] + w[8]) - (w[0] + w[1] + w[1] + w[2])) / nsres;
    const double key = dx * dx + dy * dy;
    if (!percent)
        return (float)(atan(sqrt(key) / (8 * scale)) * RAD2DEG);
    return (float)(100 * (sqrt(key) / (8 * scale)));
}

static float aspect_horn(const float *w)
{
    const double dx = (w[2] + w[5] + w[5] + w[8]) - (w[0] + w[3] + w[3] + w[6]);
    const double dy = (w[6] + w[7] + w[7] + w[8]) - (w[0] + w[1] + w[1] + w[2]);
    float a = (float)(atan2(dy, -dx) * RAD2DEG);
    if (dx == 0 && dy == 0)
        a = DST_NODATA;
    else if (a > 90.0f)
        a = 450.0f - a;
    else
        a = 90.0f - a;
    if (a == 360.0f)
        a = 0.0f;
    return a;
}

static float tri_riley(const float *w)
{
    double s = 0;
    for (int k = 0; k < 9; k++) {
        if (k == 4)
            continue;
        double d = w[k] - w[4];
        s += d * d;
    }
    return (float)sqrt(s);
}

typedef struct {
    const float *dem;
    int64_t rows, cols, row_lo, row_hi;
    int has_nodata;
    float nodata;
    double ewres, nsres, scale;
    int slope_percent;
    float *slope, *aspect, *tri;
} job_t;

static int gdaldem_rows(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                        float nodata, double ewres, double nsres, double scale,
                        int slope_percent, float *slope, float *aspect, float *tri,
                        int64_t row_lo, int64_t row_hi);

static void *job_main(void *arg)
{
    job_t *j = (job_t *)arg;
    gdaldem_rows(j->dem, j->rows, j->cols, j->has_nodata, j->nodata, j->ewres, j->nsres,
                 j->scale, j->slope_percent, j->slope, j->aspect, j->tri, j->row_lo, j->row_hi);
    return NULL;
}

/* Rows are independent.  gdaldem itself is single-threaded (ora_set_threads(1), the
 * default); bench.py's all-core CPU baseline asks for more (pthreads, one block of rows
 * per thread). */
int ora_gdaldem(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                float nodata, double ewres, double nsres, double scale,
                int slope_percent, float *slope, float *aspect, float *tri)
{
    int nt = ora_threads;
    if (nt > 256)
        nt = 256;
    if (nt <= 1 || rows < 2 * nt)
        return gdaldem_rows(dem, rows, cols, has_nodata, nodata, ewres, nsres, scale,
                            slope_percent, slope, aspect, tri, 0, rows);
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nt; t++) {
        job_t j = {dem, rows, cols, rows * t / nt, rows * (t + 1) / nt, has_nodata, nodata,
                   ewres, nsres, scale, slope_percent, slope, aspect, tri};
        jobs[t] = j;
        pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < nt; t++)
        pthread_join(th[t], NULL);
    return 0;
}

static int gdaldem_rows(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                        float nodata, double ewres, double nsres, double scale,
                        int slope_percent, float *slope, float *aspect, float *tri,
                        int64_t row_lo, int64_t row_hi)
{
    nd_t nd = {has_nodata, nodata, has_nodata && isnan(nodata)};
    if (rows < 2 || cols < 2) {
        for (int64_t i = 0; i < rows * cols; i++) {
            if (slope)
                slope[i] = DST_NODATA;
            if (aspect)
                aspect[i] = DST_NODATA;
            if (tri)
                tri[i] = DST_NODATA;
        }
        return 0;
    }
#define L(i, j) dem[(int64_t)(i) * cols + (j)]
    for (int64_t i = row_lo; i < row_hi; i++) {
        for (int64_t j = 0; j < cols; j++) {
            float w[9];
            if (i == 0 || i == rows - 1) {
                int64_t jm = j == 0 ? j : j - 1, jp = j == cols - 1 ? j : j + 1;
                int64_t jj[3] = {jm, j, jp};
                int64_t in = i, nb = i == 0 ? 1 : rows - 2;
                for (int k = 0; k < 3; k++) {
                    float e = interpol(&nd, L(in, jj[k]), L(nb, jj[k]));
                    if (i == 0) {
                        w[k] = e;
                        w[3 + k] = L(in, jj[k]);
                        w[6 + k] = L(nb, jj[k]);
                    }
                    else {
                        w[k] = L(nb, jj[k]);
                        w[3 + k] = L(in, jj[k]);
                        w[6 + k] = e;
                    }
                }
            }
            else if (j == 0) {
                for (int k = 0; k < 3; k++) {
                    int64_t ii = i - 1 + k;
                    w[3 * k] = interpol(&nd, L(ii, 0), L(ii, 1));
                    w[3 * k + 1] = L(ii, 0);
                    w[3 * k + 2] = L(ii, 1);
                }
            }
            else if (j == cols - 1) {
                for (int k = 0; k < 3; k++) {
                    int64_t ii = i - 1 + k;
                    w[3 * k] = L(ii, j - 1);
                    w[3 * k + 1] = L(ii, j);
                    w[3 * k + 2] = interpol(&nd, L(ii, j), L(ii, j - 1));
                }
            }
            else {
                for (int k = 0; k < 3; k++)
                    for (int m = 0; m < 3; m++)
                        w[3 * k + m] = L(i - 1 + k, j - 1 + m);
            }
            int64_t o = i * cols + j;
            if (is_nd(&nd, w[4])) {
                if (slope)
                    slope[o] = DST_NODATA;
                if (aspect)
                    aspect[o] = DST_NODATA;
                if (tri)
                    tri[o] = DST_NODATA;
                continue;
            }
            for (int k = 0; k < 9; k++)
                if (is_nd(&nd, w[k]))
                    w[k] = w[4];
            if (slope)
                slope[o] = slope_horn(w, ewres, nsres, scale, slope_percent);
            if (aspect)
                aspect[o] = aspect_horn(w);
            if (tri)
                tri[o] = tri_riley(w);
        }
    }
#undef L
    return 0;
}
