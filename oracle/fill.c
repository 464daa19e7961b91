/*
 * oracle/fill.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Priority-flood depression filling (Barnes, Lehman & Mulla 2014) on GRASS's integer
 * millimetre elevations, seeded like r.watershed seeds its A* search (region border and
 * NULL-adjacent cells; rule R2 of rwatershed.c):
 *     W(seed) = alt(seed),   W(c) = max(alt(c), min over 8-neighbours W(n))
 * i.e. the level at which r.watershed's search reaches c (the running maximum of the
 * popped elevations).  The general D8 path of the product (csrc/ht_astar.cu) fills on the
 * device by relaxation; this sequential version is what the tests check it against, and
 * what the filled-DEM expectation of r.watershed is built from
 * (/root/reference/src/hydrotools/watershed.py:64-75 on a DEM with depressions).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "oracle.h"

typedef struct {
    int32_t w;
    int64_t cell;
} item_t;

static void sift_up(item_t *h, int64_t i)
{
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (h[p].w <= h[i].w)
            break;
        item_t t = h[p];
        h[p] = h[i];
        h[i] = t;
        i = p;
    }
}

static void sift_down(item_t *h, int64_t n, int64_t i)
{
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && h[l].w < h[m].w)
            m = l;
        if (r < n && h[r].w < h[m].w)
            m = r;
        if (m == i)
            break;
        item_t t = h[m];
        h[m] = h[i];
        h[i] = t;
        i = m;
    }
}

int ora_fill_mm(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                float nodata, int32_t *w_out)
{
    const int64_t n = rows * cols;
    if (n <= 0)
        return 0;
    uint8_t *done = (uint8_t *)calloc(n, 1);
    item_t *heap = (item_t *)malloc(n * sizeof(item_t));
    if (!done || !heap)
        return -1;
    int64_t hn = 0;
    for (int64_t i = 0; i < n; i++) {
        float z = dem[i];
        w_out[i] = ((has_nodata && z == nodata) || isnan(z)) ? ORA_ALT_NULL : ora_alt_from_f32(z);
    }
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            int64_t i = r * cols + c;
            if (w_out[i] == ORA_ALT_NULL) {
                done[i] = 1;
                continue;
            }
            int seed = r == 0 || c == 0 || r == rows - 1 || c == cols - 1;
            for (int dr = -1; dr <= 1 && !seed; dr++)
                for (int dc = -1; dc <= 1; dc++)
                    if (w_out[(r + dr) * cols + c + dc] == ORA_ALT_NULL)
                        seed = 1;
            if (seed) {
                heap[hn].w = w_out[i];
                heap[hn].cell = i;
                sift_up(heap, hn++);
                done[i] = 1;
            }
        }
    while (hn > 0) {
        item_t top = heap[0];
        heap[0] = heap[--hn];
        sift_down(heap, hn, 0);
        int64_t r = top.cell / cols, c = top.cell % cols;
        for (int dr = -1; dr <= 1; dr++)
            for (int dc = -1; dc <= 1; dc++) {
                int64_t rr = r + dr, cc = c + dc;
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
                    continue;
                int64_t j = rr * cols + cc;
                if (done[j])
                    continue;
                done[j] = 1;
                if (w_out[j] < top.w)
                    w_out[j] = top.w;
                heap[hn].w = w_out[j];
                heap[hn].cell = j;
                sift_up(heap, hn++);
            }
    }
    free(done);
    free(heap);
    return 0;
}
