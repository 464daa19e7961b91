/*
 * oracle/flowacc.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the GRASS addon r.flowaccumulation as the reference uses
 * it (/root/reference/src/hydrotools/watershed.py:341-355 cell counts
 * "modals"; :369-382 weighted sum, type=DCELL):
 *     acc(c) = w(c) + sum of acc(u) over cells u whose D8 receiver is c
 * on the direction code 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE 8=E
 * (/root/reference/src/hydrotools/watershed.py:575-585); codes <= 0 and NULL
 * are terminal, exactly how every consumer in the reference treats them
 * (watershed.py:589 `fd[i, j] <= 0`, streams.py:100,122, morphology.py:117).
 * The addon is not vendored and not installed by the reference Dockerfile
 * (only at run time, utils.py:232-239).  PARITY UNPINNED.  Assumptions made
 * here and recorded in DESIGN.md: flow into a NULL-direction cell or out of
 * the region vanishes; a NULL weight contributes 0; output is NULL (NaN)
 * where the direction is NULL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "oracle.h"

static const int asp_r[9] = {0, -1, -1, -1, 0, 1, 1, 1, 0};
static const int asp_c[9] = {0, 1, 0, -1, -1, -1, 0, 1, 1};

static int64_t receiver(const int8_t *dir, int64_t rows, int64_t cols,
                        int64_t i)
{
    int d = dir[i];
    if (d <= 0 || d > 8)
        return -1;
    int64_t r = i / cols + asp_r[d], c = i % cols + asp_c[d];
    if (r < 0 || r >= rows || c < 0 || c >= cols)
        return -1;
    int64_t j = r * cols + c;
    if (dir[j] == ORA_DIR_NULL)
        return -1;
    return j;
}

static int accumulate(const int8_t *dir, int64_t rows, int64_t cols,
                      double *acc)
{
    int64_t n = rows * cols;
    uint8_t *indeg = (uint8_t *)calloc(n ? n : 1, 1);
    int64_t *queue = (int64_t *)malloc((n ? n : 1) * sizeof(int64_t));
    if (!indeg || !queue)
        return -1;
    for (int64_t i = 0; i < n; i++) {
        int64_t j = receiver(dir, rows, cols, i);
        if (j >= 0 && dir[i] != ORA_DIR_NULL)
            indeg[j]++;
    }
    int64_t head = 0, tail = 0;
    for (int64_t i = 0; i < n; i++)
        if (!indeg[i] && dir[i] != ORA_DIR_NULL)
            queue[tail++] = i;
    while (head < tail) {
        int64_t i = queue[head++];
        int64_t j = receiver(dir, rows, cols, i);
        if (j < 0)
            continue;
        acc[j] += acc[i];
        if (--indeg[j] == 0)
            queue[tail++] = j;
    }
    for (int64_t i = 0; i < n; i++)
        if (dir[i] == ORA_DIR_NULL)
            acc[i] = NAN;
    free(indeg);
    free(queue);
    return 0;
}

int ora_flowacc_count(const int8_t *dir, int64_t rows, int64_t cols,
                      double *acc)
{
    for (int64_t i = 0; i < rows * cols; i++)
        acc[i] = 1.0;
    return accumulate(dir, rows, cols, acc);
}

int ora_flowacc_weighted(const int8_t *dir, const float *w, int has_wnodata,
                         float wnodata, int64_t rows, int64_t cols,
                         double *acc)
{
    for (int64_t i = 0; i < rows * cols; i++) {
        float v = w[i];
        acc[i] = ((has_wnodata && v == wnodata) || isnan(v)) ? 0.0 : (double)v;
    }
    return accumulate(dir, rows, cols, acc);
}
