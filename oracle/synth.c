/*
 * oracle/synth.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU build of the integer-exact synthetic DEM / weight generator declared in
 * include/ht_synth.h, so tests can feed the oracle and the CUDA path the same
 * bits (the product generates its DEMs on the device with the same header).
 */
#include <math.h>
#include "../include/ht_synth.h"
#include "oracle.h"

void ora_synth_dem(uint64_t seed, int64_t rows, int64_t cols, int64_t row0,
                   int64_t total_rows, int with_nulls, float nodata, float *out)
{
    ht_synth_params p;
    ht_synth_make_params((uint32_t)seed, total_rows, with_nulls, &p);
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++)
            out[r * cols + c] = ht_synth_is_null(&p, row0 + r, c)
                                    ? nodata
                                    : ht_synth_dem_value(&p, row0 + r, c);
}

void ora_synth_weight(uint64_t seed, int64_t rows, int64_t cols, int64_t row0,
                      float *out)
{
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++)
            out[r * cols + c] = ht_synth_weight_value((uint32_t)seed, row0 + r, c);
}
