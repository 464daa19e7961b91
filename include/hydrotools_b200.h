/*
 * hydrotools_b200.h -- C-ABI of libhydrotools_b200.so (sm_100a).
 *
 * The reference (bluegeo/hydro-tools) has no FFI: its hot path is Python that
 * shells out to GRASS / GDAL binaries.  Each entry point below names the
 * reference call it replaces; `hydrotools.cuda` (Python, ctypes) is the only
 * intended caller and mirrors the reference signatures on top of these.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *    the caller owns all buffers, the library never frees caller memory;
 *  - rasters are row-major, `ld` = row pitch in ELEMENTS.  Float rasters read
 *    through TMA need a 16-byte aligned base and `ld % 4 == 0`
 *    (HT_ERR_ALIGN otherwise; the Python host pads);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *    the call returns without synchronising unless stated;
 *  - return 0 = ok, negative = HT_ERR_* argument/driver error, positive =
 *    cudaError_t.  No C++ exception crosses the ABI.  Re-entrant; no global
 *    state besides a cached driver entry point.
 *  - row-band sharding: a rank passes a pointer to its FIRST OWNED row plus
 *    `halo_top` / `halo_bot` = how many valid rows of its neighbours' data sit
 *    directly above / below in the same allocation (0 = that side is the edge
 *    of the whole raster), and `row0` / `total_rows` for whole-raster position.
 */
#ifndef HYDROTOOLS_B200_H
#define HYDROTOOLS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HT_OK 0
#define HT_ERR_ARG (-1)    /* bad size / null pointer / unsupported option */
#define HT_ERR_ALIGN (-2)  /* base or pitch alignment not met */
#define HT_ERR_DRIVER (-3) /* cuTensorMapEncodeTiled unavailable or failed */
#define HT_ERR_WORKSPACE (-4)

/* nodata written by the terrain kernels: gdaldem's fixed destination nodata */
#define HT_TERRAIN_NODATA (-9999.0f)

/* device direction byte ("packed dir"): low nibble = |GRASS code| 0..8
 * (1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE 8=E,
 *  /root/reference/src/hydrotools/watershed.py:575-585) */
#define HT_DIR_CODE 0x0F
#define HT_DIR_NEG 0x10   /* code is negative (points off-region / into NULL) */
#define HT_DIR_EDGE 0x20  /* r.watershed "edge" cell: does not pass flow on */
#define HT_DIR_TAINT 0x40 /* r.watershed flipped this cell's sign negative */
#define HT_DIR_NULL 0x80  /* NULL cell */

int ht_version(void);
const char *ht_error_string(int code);

/* ---- elevation.slope / aspect / terrain_ruggedness_index ------------------
 * replaces  subprocess gdaldem {slope,aspect,TRI} -compute_edges
 *   /root/reference/src/hydrotools/elevation.py:41-54, 70-78, 118-126
 * One read of the DEM tile, up to three fp32 outputs (any may be NULL).
 * slope_percent: 0 = degrees, 1 = percent (gdaldem -p); zscale = gdaldem -s. */
int ht_terrain_fused_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld,
                         int has_nodata, float nodata, double ewres,
                         double nsres, double zscale, int slope_percent,
                         float *slope, float *aspect, float *tri, int64_t out_ld,
                         int halo_top, int halo_bot, void *stream);

/* ---- synthetic workloads (BASELINE.json configs 2-5) ----------------------
 * integer-exact fractal DEM / weights, see include/ht_synth.h */
int ht_synth_dem_f32(float *dem, int64_t rows, int64_t cols, int64_t ld,
                     uint64_t seed, int64_t row0, int64_t total_rows,
                     int with_nulls, float nodata, void *stream);
int ht_synth_weight_f32(float *w, int64_t rows, int64_t cols, int64_t ld,
                        uint64_t seed, int64_t row0, void *stream);

#ifdef __cplusplus
}
#endif
#endif
