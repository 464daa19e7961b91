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

/* ---- watershed.flow_direction_accumulation, direction half ------------------
 * replaces  GRASS r.watershed -s  drainage=fd
 *   /root/reference/src/hydrotools/watershed.py:64-75 (utils.py:157-159 runs it)
 * Input must be hydrologically conditioned (pit-free, tie-free after GRASS's
 * round(z*1000)); ht_d8_validate_conditioned_f32 counts violations.
 * Needs 2 halo rows per shared side when sharded.  Writes the packed byte of
 * the owned rows and, when a side has halo rows, of the first halo row beyond it
 * (packed row -1 / row `rows`), which stage A/C of the accumulation read.
 * `packed`: 16-B aligned, packed_ld % 16 == 0. */
int ht_d8_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld, int has_nodata,
              float nodata, double ewres, double nsres, int64_t row0,
              int64_t total_rows, int halo_top, int halo_bot, uint8_t *packed,
              int64_t packed_ld, void *stream);
/* a run of rows of a band: the packed byte of the row above / below the run is stored only
 * with ring_top / ring_bot (a rank computes its interior rows while its neighbours' halo
 * rows are in flight, then its first and last tile rows) */
int ht_d8_part_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld, int has_nodata,
                   float nodata, double ewres, double nsres, int64_t row0,
                   int64_t total_rows, int halo_top, int halo_bot, int ring_top,
                   int ring_bot, uint8_t *packed, int64_t packed_ld, void *stream);
/* counts3[0] = pits (non-seed cells without a strictly lower neighbour),
 * counts3[1] = cells whose 3x3 window holds two equal integer elevations,
 * counts3[2] = cells whose |elevation| >= 2^24 mm = 16 777 m (unsupported; clamped) */
int ht_d8_validate_conditioned_f32(const float *dem, int64_t rows, int64_t cols,
                                   int64_t ld, int has_nodata, float nodata,
                                   int64_t row0, int64_t total_rows, int halo_top,
                                   int halo_bot, unsigned long long *counts3,
                                   void *stream);

/* ---- the same for ANY DEM: equal elevations, flats, depressions ------------------
 * r.watershed's least-cost search (heap keyed by elevation, first-in-first-out among
 * equals; /root/reference/src/hydrotools/watershed.py:64-75 on a raw DEM such as the
 * reference's own tests/data-files/dem.tif) reproduced bit for bit in parallel rounds:
 * priority-flood levels by relaxation, then every cell's direction from the cells that
 * pop before it; the cells of a depression are replayed sequentially by the thread of the
 * cell that enters it (csrc/ht_astar_rules.h).  Whole rasters only (no row bands), fewer
 * than 2^31 cells.  Synchronous: host loops read a device counter per round.
 * stats8_host (may be NULL): [0] rounds, [1] depression episodes, [2] cells of the largest,
 * [3] depression cells, [4] fill passes, [5] generation sweeps, [6] cells outside the
 * supported elevation range (|z| >= 16 777 m; the result is then not valid), [7] errors. */
size_t ht_d8_general_workspace_bytes(int64_t rows, int64_t cols);
int ht_d8_general_f32(const float *dem, int64_t rows, int64_t cols, int64_t ld, int has_nodata,
                      float nodata, double ewres, double nsres, uint8_t *packed,
                      int64_t packed_ld, void *ws, size_t ws_bytes, long long *stats8_host,
                      void *stream);

/* GRASS int8 codes (-128 = NULL) <-> packed byte; used for direction rasters that
 * come from disk (FlowAccumulation(direction),
 * /root/reference/src/hydrotools/watershed.py:329-339).  A code 1..8 whose
 * receiver is NULL or outside the raster is flagged HT_DIR_EDGE (does not pass
 * flow on), which is what r.flowaccumulation does with it. */
int ht_dir_pack_i8(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                   uint8_t *packed, int64_t packed_ld, void *stream);
/* row-band form: `dir` / `packed` point at the band's first owned row; with halo_top /
 * halo_bot (0 / 1) the edge row of the neighbour band sits directly above / below in both
 * allocations (the direction row arrived by exchange; its packed byte is written too) */
int ht_dir_pack_band_i8(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                        int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                        uint8_t *packed, int64_t packed_ld, void *stream);
int ht_dir_unpack_i8(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                     int64_t cols, int8_t *dir, int64_t ld, void *stream);

/* ---- accumulation -----------------------------------------------------------
 * replaces  GRASS r.watershed accumulation=fa [-a]   (watershed.py:64-75)
 *      and  GRASS r.flowaccumulation [weight=]      (watershed.py:341-382)
 * mode: 0 = cell count, 1 = count of tainted cells upstream (sign step),
 *       2 = fp32 weights, summed exactly in fixed point, reported as fp64.
 * Stage A (ht_acc_local) builds the tile boundary graph in the workspace,
 * stage B (ht_acc_resolve) resolves it by pointer doubling until no pointer
 * changes, stage C (ht_acc_final) writes acc (fp64; NaN where NULL) and, if
 * dir_out != NULL, the final GRASS direction codes (int8, -128 = NULL).
 * Count modes: acc 16-B aligned with acc_ld % 2 == 0, dir_out 4-B aligned with
 * dir_ld % 4 == 0 (HT_ERR_ALIGN otherwise).
 * ht_acc_f64 = A + B + C for an unsharded raster. */
size_t ht_acc_workspace_bytes(int64_t rows, int64_t cols); /* count modes */
/* mode 2 carries 128-bit sums and 8 B/cell of tile-local sums: a larger workspace */
size_t ht_acc_workspace_bytes_mode(int64_t rows, int64_t cols, int mode);
/* weighted mode works in exact fixed point (bit-reproducible, no fp64 atomics); the
 * unit follows from the raster's largest finite |weight|.  ht_acc_weight_max writes its
 * bits (a non-negative float stored in an unsigned, so that row bands combine theirs with
 * an integer MAX all-reduce), ht_acc_set_weight_max hands the result to the workspace
 * before ht_acc_local(mode 2).  ht_acc_f64 does both itself. */
int ht_acc_weight_max(const float *w, int64_t w_ld, int64_t rows, int64_t cols, int has_wnd,
                      float wnd, unsigned *max_bits, void *stream);
int ht_acc_set_weight_max(int64_t rows, int64_t cols, void *ws, size_t ws_bytes,
                          const unsigned *max_bits, void *stream);
/* byte offsets of the boundary-graph arrays inside a COUNT-mode workspace (tests and
 * tools): layout11[0..8] = base0, aggA, aggB (resolved inflow), next0, ptrA, ptrB,
 * rootA, rootB (last node of each slot's path), flags; [9] = slots; [10] = first
 * virtual slot (cells of the band above, then of the band below, `cols` each). */
int ht_acc_workspace_layout(int64_t rows, int64_t cols, void *ws, size_t ws_bytes,
                            int64_t *layout11_host);
/* subtree sums over a forest given by next[] (< 0 = root); the band-level
 * boundary graph of a row-band sharded raster is resolved with it.
 * vtype: 0 = u64, 1 = fp64, 2 = 128-bit integers {u64 lo, i64 hi} (weighted mode).
 * tmp_bytes >= (3 * value size + 8) * n + 4096. */
int ht_forest_resolve(const void *base, const int32_t *next, int64_t n, int vtype,
                      void *agg_out, void *tmp, size_t tmp_bytes, void *stream);
int ht_acc_local(const uint8_t *packed, int64_t packed_ld, int64_t rows, int64_t cols,
                 int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                 int mode, const float *w, int64_t w_ld, int has_wnd, float wnd,
                 void *ws, size_t ws_bytes, void *stream);
int ht_acc_resolve(int64_t rows, int64_t cols, int mode, int want_roots, void *ws,
                   size_t ws_bytes, int *rounds_dev, void *stream);
/* row-band sharding, after ht_acc_local: tag the boundary-graph slots of the band's first
 * (halo_top) / last (halo_bot) row with a mark (1 << 40 added to the u64 counts; 1 added to
 * the low word of the 128-bit weighted sums, whose low 24 bits are free) that rides through
 * ht_acc_resolve to every node downstream */
int ht_acc_mark_edges(int64_t rows, int64_t cols, int mode, int halo_top, int halo_bot,
                      void *ws, size_t ws_bytes, void *stream);
/* row-band sharding: after ht_acc_resolve on a marked graph, add the effect of the flows
 * the neighbour bands deliver into the band's first / last row (imp_top / imp_bot: cols
 * values of the mode's type each -- u64 or 128-bit -- or NULL) and clear the marks: a
 * resolve over the nodes downstream of the band edge instead of a second full one */
int ht_acc_resolve_delta(int64_t rows, int64_t cols, int mode, const void *imp_top,
                         const void *imp_bot, void *ws, size_t ws_bytes, void *stream);
/* row-band sharding, any mode: add the imported flows (cols values of the mode's type, u64
 * or 128-bit, or NULL) to the boundary graph; a second full ht_acc_resolve follows */
int ht_acc_add_imports(int64_t rows, int64_t cols, int mode, const void *imp_top,
                       const void *imp_bot, void *ws, size_t ws_bytes, void *stream);
/* row-band sharding, after ht_acc_resolve(want_roots = 1): the tables a band publishes.
 * count modes: out[4][cols] (int64): [0] / [1] flow delivered into column c of the edge
 * row of the band above / below, [2] / [3] for the band's own first / last row cell c the
 * foreign cell its path finally reaches (side * cols + column, side 0 = band above) or -1.
 * mode 2: out[6][cols]: [0],[1] low / high word of the 128-bit delivery to the band above,
 * [2],[3] to the band below, [4],[5] the links. */
int ht_band_tables(int64_t rows, int64_t cols, int mode, void *ws, size_t ws_bytes,
                   long long *out, void *stream);
/* the band-level forest from the all-gathered tables every[bands][4 or 6][cols]:
 * base[bands][2][cols] (u64, or 128-bit in mode 2), next[bands][2][cols] for
 * ht_forest_resolve */
int ht_band_graph(const long long *every, int bands, int64_t cols, int mode, void *base,
                  int32_t *next, void *stream);
int ht_acc_final(const uint8_t *packed, int64_t packed_ld, int64_t rows, int64_t cols,
                 int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                 int mode, const float *w, int64_t w_ld, int has_wnd, float wnd,
                 double *acc, int64_t acc_ld, int8_t *dir_out, int64_t dir_ld, void *ws,
                 size_t ws_bytes, void *stream);
/* row-band sharding: stage C in two passes, so that most of it runs while the band tables
 * travel.  ht_acc_tile_phases (after ht_acc_resolve on a marked graph) leaves one byte per
 * tile: 1 = an entry slot of the tile carries a band-edge mark, i.e. its inflow changes with
 * what the neighbour bands deliver.  ht_acc_final_phase(phase = 0) writes the tiles that do
 * not depend on the exchange, phase = 1 (after ht_acc_resolve_delta) the others; phase < 0 =
 * every tile (= ht_acc_final). */
int ht_acc_tile_phases(int64_t rows, int64_t cols, int mode, void *ws, size_t ws_bytes,
                       void *stream);
int ht_acc_final_phase(const uint8_t *packed, int64_t packed_ld, int64_t rows, int64_t cols,
                       int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                       int mode, const float *w, int64_t w_ld, int has_wnd, float wnd,
                       double *acc, int64_t acc_ld, int8_t *dir_out, int64_t dir_ld, void *ws,
                       size_t ws_bytes, int phase, void *stream);
/* stage B of the count modes has two variants with identical results: one flat pointer-
 * doubling resolve over all nodes, or two levels (super-tiles of 8 x 4 tiles resolved in
 * shared memory, the flat resolve on the nodes that receive flow from another super-tile,
 * a walk that hands the external inflow down).  mode 0 = automatic (two levels from 128
 * tiles on), 1 = flat, 2 = two levels; returns the previous mode.  Test hook; the
 * environment variable HT_STAGE_B = flat | two sets the initial mode. */
int ht_acc_stage_b_mode(int mode);
int ht_acc_f64(const uint8_t *packed, int64_t packed_ld, int64_t rows, int64_t cols,
               int mode, const float *w, int64_t w_ld, int has_wnd, float wnd,
               double *acc, int64_t acc_ld, int8_t *dir_out, int64_t dir_ld, void *ws,
               size_t ws_bytes, void *stream);
/* positive_only=False: negate acc at edge cells and below tainted cells
 * (taint_acc = result of mode 1) */
int ht_acc_apply_sign(double *acc, int64_t acc_ld, const double *taint_acc,
                      int64_t taint_ld, const uint8_t *packed, int64_t packed_ld,
                      int64_t rows, int64_t cols, void *stream);

/* ---- downstream consumers of the direction grid -------------------
 * fd: int8 GRASS codes (1 = NE ... 8 = E, <= 0 and -128 end a path).  All three synchronise
 * the stream (pointer-jumping sweeps / an inlet count are read back).  Fewer than 2^31 cells.
 * ws: ht_streams_workspace_bytes(rows, cols, max_window, max_inlets) bytes (0, 0 for labels
 * and basin). */
size_t ht_streams_workspace_bytes(int64_t rows, int64_t cols, int max_window, int64_t max_inlets);
/* replaces the numba walker `_fa_label_task`
 *   /root/reference/src/hydrotools/streams.py:320-388
 * labels (uint32, 0 = no reach) number the connected runs of cells that hold one reach value
 * (!= nodata) and are linked by the direction grid, in row-major order of their first cell. */
int ht_reach_labels_i32(const int32_t *reaches, int64_t r_ld, const int8_t *fd, int64_t fd_ld,
                        int64_t rows, int64_t cols, int32_t nodata, uint32_t *labels,
                        int64_t l_ld, void *ws, size_t ws_bytes, void *stream);
/* replaces the numba walker `_upstream_compute`
 *   /root/reference/src/hydrotools/morphology.py:50-153 (upstream_stats, :292-360)
 * streams: nonzero = stream cell.  From every inlet (stream cell no stream cell drains
 * into) down the stream: out[cell] = mean of `values` over the cells of the path within
 * `distance` map units upstream (fp64 running sum in the reference's order, stored fp32);
 * where paths overlap the inlet latest in row-major order prevails, as in the reference.
 * `out` must arrive filled (the reference starts from NaN); cells never visited keep it.
 * max_window = ceil(distance / min(csx, csy)) + 2 (morphology.py:331). */
int ht_upstream_mean_f64(const double *values, int64_t v_ld, const uint8_t *streams, int64_t s_ld,
                         const int8_t *fd, int64_t fd_ld, int64_t rows, int64_t cols, double csx,
                         double csy, double distance, int max_window, int64_t max_inlets,
                         float *out, int64_t o_ld, void *ws, size_t ws_bytes,
                         long long *ninlets_host, void *stream);
/* replaces GRASS r.water.outlet (/root/reference/src/hydrotools/watershed.py:292-318):
 * out = 1 where the cell's path passes through (outlet_row, outlet_col), the outlet
 * included, else 0. */
int ht_basin_u8(const int8_t *fd, int64_t fd_ld, int64_t rows, int64_t cols, int64_t outlet_row,
                int64_t outlet_col, uint8_t *out, int64_t o_ld, void *ws, size_t ws_bytes,
                void *stream);

/* ---- interpolate.raster_filter / custom_raster_filter ------------------------
 * replaces the numba moving-window kernel `apply_filter` under
 *   da.map_overlap(depth, boundary=nan)
 *   /root/reference/src/hydrotools/interpolate.py:42-64, 94-107; reducers :147-222
 * for an arbitrary boolean kernel (kernel_host: kh x kw bytes in HOST memory, nonzero =
 * selected, kh, kw <= 33).  Window rows i - kh/2 .. i + kh - 1 - kh/2 (likewise columns),
 * NaN outside the raster; fp64 sample and reducer, fp32 result, NaN where the centre is
 * NaN (or equals `nodata`) or no selected cell holds data.
 * method: 0 sum, 1 min, 2 max, 3 diff (max - min), 4 mean, 5 std, 6 variance (ddof 0),
 * 7 median (np.nanmedian), 8 mode (first value, in window scan order, of the largest
 * multiplicity -- the reference's dictionary walk, interpolate.py:174-201). */
int ht_window_filter_f32(const float *src, int64_t rows, int64_t cols, int64_t ld,
                         int has_nodata, float nodata, const unsigned char *kernel_host,
                         int kh, int kw, int method, float *out, int64_t out_ld,
                         int halo_top, int halo_bot, void *stream);

/* the same filter for sum (0) / mean (4) over a boolean kernel of ANY size (the discs
 * of kernel_from_distance, /root/reference/src/hydrotools/utils.py:408-426, as used at
 * morphology.py:282-289 and :1373-1374): row prefix sums in fp64 + one difference per
 * run of selected cells in a kernel row (<= 64 runs per row).  ws: device scratch of
 * ht_window_filter_runs_workspace_bytes(rows + halo_top + halo_bot, cols, kh) bytes.
 * Synchronises the stream once (upload of the run table). */
size_t ht_window_filter_runs_workspace_bytes(int64_t rows_with_halos, int64_t cols, int kh);
int ht_window_filter_runs_f32(const float *src, int64_t rows, int64_t cols, int64_t ld,
                              int has_nodata, float nodata, const unsigned char *kernel_host,
                              int kh, int kw, int method, float *out, int64_t out_ld,
                              int halo_top, int halo_bot, void *ws, size_t ws_bytes,
                              void *stream);

/* ---- raster.to_raster, the on-disk step --------------------------
 * replaces the compression half of  to_raster -> tiled 512 x 512 GeoTIFF -> gdal_translate COG
 *   /root/reference/src/hydrotools/raster.py:794-846, utils.py:78-98, config.py:23-45
 * A device raster (row pitch ld_bytes, samples of 1 / 2 / 4 / 8 bytes) becomes one zlib
 * stream per 512 x 512 tile -- what TIFF compression 8 (deflate) stores -- so that only the
 * compressed bytes cross PCIe; the host writes the BigTIFF / COG container
 * (hydrotools/cuda/_cog.py).  pad_value: the sample stored beyond the raster edge inside
 * edge tiles (little endian in the low bytes).  out: ht_tiff_encode_bound() bytes; tile t
 * (row-major over the tile grid) is out[tile_off[t] .. + tile_size[t]), *total bytes in all.
 * tile_off / tile_size / total are DEVICE memory; nothing synchronises. */
size_t ht_tiff_encode_bound(int64_t rows, int64_t cols, int elem_bytes);
size_t ht_tiff_encode_workspace_bytes(int64_t rows, int64_t cols, int elem_bytes);
int ht_tiff_encode_tiles(const void *src, int64_t ld_bytes, int64_t rows, int64_t cols,
                         int elem_bytes, unsigned long long pad_value, uint8_t *out,
                         size_t out_cap, unsigned long long *tile_off,
                         unsigned long long *tile_size, unsigned long long *total, void *ws,
                         size_t ws_bytes, void *stream);

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
