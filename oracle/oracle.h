/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatements of the third-party engines behind the reference's hot path
 * (GRASS r.watershed / r.flowaccumulation, GDAL gdaldem).  See each .c file
 * for the reference call site it follows.  PARITY UNPINNED (see rwatershed.c).
 */
#ifndef HT_ORACLE_H
#define HT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_DIR_NULL (-128) /* hydrotools.utils.infer_nodata("int8") */

int32_t ora_alt_from_f32(float z);

/* rwatershed.c */
int ora_rwatershed_sfd(const float *dem, int64_t rows, int64_t cols,
                       int has_nodata, float nodata, double ewres, double nsres,
                       int flag_a, int edge_rule, int8_t *dir_out,
                       double *acc_out, int64_t *order_out);
int ora_check_conditioned(const float *dem, int64_t rows, int64_t cols,
                          int has_nodata, float nodata, int64_t *n_pits,
                          int64_t *n_ties);
int ora_condition_by_rank(const float *dem, int64_t rows, int64_t cols,
                          int has_nodata, float nodata, double ewres,
                          double nsres, float *out);

/* fill.c */
#define ORA_ALT_NULL 0x7FFFFFFF
int ora_fill_mm(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                float nodata, int32_t *w_out);

/* gdaldem.c */
int ora_gdaldem(const float *dem, int64_t rows, int64_t cols, int has_nodata,
                float nodata, double ewres, double nsres, double scale,
                int slope_percent, float *slope, float *aspect, float *tri);

void ora_set_threads(int n); /* threads of ora_gdaldem (default 1) */

/* flowacc.c */
int ora_flowacc_count(const int8_t *dir, int64_t rows, int64_t cols,
                      double *acc);
int ora_flowacc_weighted(const int8_t *dir, const float *w, int has_wnodata,
                         float wnodata, int64_t rows, int64_t cols,
                         double *acc);

/* synth.c */
void ora_synth_dem(uint64_t seed, int64_t rows, int64_t cols, int64_t row0,
                   int64_t total_rows, int with_nulls, float nodata,
                   float *out);
void ora_synth_weight(uint64_t seed, int64_t rows, int64_t cols, int64_t row0,
                      float *out);

#ifdef __cplusplus
}
#endif
#endif
