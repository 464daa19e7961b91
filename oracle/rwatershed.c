/*
 * oracle/rwatershed.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of GRASS GIS r.watershed, ram mode, single flow direction
 * (flags -s [-a]), which is what the reference calls for its hot path:
 *   /root/reference/src/hydrotools/watershed.py:64-75
 *     gr.run_command("r.watershed", elevation="dem", drainage="fd",
 *                    accumulation="fa", flags="a"+"s")
 * GRASS itself is NOT vendored under /root/reference (Dockerfile:14-16 installs
 * Ubuntu 24.04's apt `grass` = 8.3.x), so this file restates the published
 * algorithm of raster/r.watershed/ram/{init_vars.c, do_astar.c, do_cum.c,
 * close_maps.c} from memory of the upstream sources.
 *
 * PARITY UNPINNED: the reference's only test of this path
 * (tests/test_watershed.py:10-16) asserts nothing, and neither GRASS nor the
 * reference's Python stack can run in this image, so no golden output of the
 * real r.watershed exists to pin this restatement.  Every rule below is a
 * named step so that it can be corrected in one place once a box with GRASS
 * is available (tools/regen_goldens_with_grass.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may call this.  The product path never does.
 *
 * Steps (names used in DESIGN.md):
 *  R1 load      alt = (int) round-half-away(z * 1000) for FCELL input
 *               (init_vars.c: ele_scale = 1000 for FP maps; ele_round()).
 *               NULL = (z == nodata) || isnan(z)  (r.external + GDAL nodata).
 *  R2 seeds     row-major scan: region-border cells get asp -2/-4/-6/-8
 *               (r==0, c==0, r==nrows-1, c==ncols-1, in that priority),
 *               else cells with a NULL 8-neighbour get
 *               asp = -drain[r-rn+1][c-cn+1] of the FIRST NULL neighbour in
 *               the order S,N,W,E,NE,SW,SE,NW; seeds go on the heap with
 *               wat = -1 ("likely underestimate").
 *  R3 heap      min-heap keyed (alt, insertion counter): equal alt pops FIFO.
 *  R4 A*        pop D; neighbours U in the order above; slope[ct] computed
 *               for not-yet-worked U; diagonal-bias skip; first visitor sets
 *               asp[U] = drain[ur-r+1][uc-c+1]; an unworked seed U (asp<0)
 *               is re-pointed at D and D's wat sign is flipped negative.
 *  R5 do_cum    reverse pop order; receiver = cell + offset(|asp|); an EDGE
 *               cell (any 8-neighbour outside the region or NULL) does not
 *               pass its flow on ("do not distribute flow along edges"), and
 *               if it is a swale (|wat| >= 60, the default threshold) with a
 *               positive asp, its asp is reset to minus the direction of the
 *               first outside/NULL neighbour; otherwise magnitudes add and
 *               negativity is sticky.
 *  R6 output    drainage = asp (negatives kept even with -a); accumulation =
 *               wat, fabs() with -a; both NULL where the DEM is NULL.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* neighbour scan order of r.watershed: S, N, W, E, NE, SW, SE, NW */
static const int nextdr[8] = {1, -1, 0, 0, -1, 1, 1, -1};
static const int nextdc[8] = {0, 0, -1, 1, 1, -1, 1, -1};
/* cardinal neighbours of the popped cell that flank diagonal ct (ew: the one
 * an ew_res step away from the diagonal cell; ns: an ns_res step away) */
static const int nbr_ew[8] = {0, 1, 2, 3, 1, 0, 0, 1};
static const int nbr_ns[8] = {0, 1, 2, 3, 3, 2, 3, 2};
/* drain[dr+1][dc+1]: code of a cell that flows by (-dr,-dc); same table as
 * /root/reference/src/hydrotools/watershed.py:617 */
static const int drain[3][3] = {{7, 6, 5}, {8, 0, 4}, {1, 2, 3}};
/* offsets of code 1..8: 1=NE 2=N 3=NW 4=W 5=SW 6=S 7=SE 8=E
 * (/root/reference/src/hydrotools/watershed.py:575-585) */
static const int asp_r[9] = {0, -1, -1, -1, 0, 1, 1, 1, 0};
static const int asp_c[9] = {0, 1, 0, -1, -1, -1, 0, 1, 1};

typedef struct {
    int64_t *idx;   /* cell index */
    int64_t *added; /* insertion counter */
    int64_t size, cap, counter;
    const int32_t *alt;
} heap_t;

static int heap_less(const heap_t *h, int64_t a, int64_t b)
{
    int32_t ea = h->alt[h->idx[a]], eb = h->alt[h->idx[b]];
    if (ea != eb)
        return ea < eb;
    return h->added[a] < h->added[b];
}

static void heap_swap(heap_t *h, int64_t a, int64_t b)
{
    int64_t t = h->idx[a];
    h->idx[a] = h->idx[b];
    h->idx[b] = t;
    t = h->added[a];
    h->added[a] = h->added[b];
    h->added[b] = t;
}

static void heap_push(heap_t *h, int64_t cell)
{
    int64_t i = h->size++;
    h->idx[i] = cell;
    h->added[i] = h->counter++;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!heap_less(h, i, p))
            break;
        heap_swap(h, i, p);
        i = p;
    }
}

static int64_t heap_pop(heap_t *h)
{
    int64_t top = h->idx[0];
    h->size--;
    if (h->size > 0) {
        h->idx[0] = h->idx[h->size];
        h->added[0] = h->added[h->size];
        int64_t i = 0;
        for (;;) {
            int64_t l = 2 * i + 1, r = l + 1, m = i;
            if (l < h->size && heap_less(h, l, m))
                m = l;
            if (r < h->size && heap_less(h, r, m))
                m = r;
            if (m == i)
                break;
            heap_swap(h, i, m);
            i = m;
        }
    }
    return top;
}

/* R1: GRASS ele_round() on (double)z * ele_scale */
int32_t ora_alt_from_f32(float z)
{
    double d = (double)z * 1000.0;
    return (int32_t)(d > 0 ? d + 0.5 : d - 0.5);
}

static double get_slope2(int32_t ele, int32_t up_ele, double dist)
{
    if (ele >= up_ele)
        return 0.0;
    return (double)(up_ele - ele) / dist;
}

int ora_rwatershed_sfd(const float *dem, int64_t rows, int64_t cols,
                       int has_nodata, float nodata, double ewres, double nsres,
                       int flag_a, int edge_rule, int8_t *dir_out,
                       double *acc_out, int64_t *order_out)
{
    if (rows <= 0 || cols <= 0)
        return 0;
    const int64_t n = rows * cols;
    int32_t *alt = (int32_t *)malloc(n * sizeof(int32_t));
    int8_t *asp = (int8_t *)calloc(n, 1);
    double *wat = (double *)malloc(n * sizeof(double));
    uint8_t *flg = (uint8_t *)calloc(n, 1); /* 1 null, 2 in_list, 4 worked */
    int64_t *order = (int64_t *)malloc(n * sizeof(int64_t));
    heap_t h;
    h.idx = (int64_t *)malloc(n * sizeof(int64_t));
    h.added = (int64_t *)malloc(n * sizeof(int64_t));
    h.size = 0;
    h.cap = n;
    h.counter = 0;
    h.alt = alt;
    if (!alt || !asp || !wat || !flg || !order || !h.idx || !h.added)
        return -1;

    double dist[8];
    for (int ct = 0; ct < 8; ct++) {
        double dy = abs(nextdr[ct]) * nsres, dx = abs(nextdc[ct]) * ewres;
        dist[ct] = ct < 4 ? dx + dy : sqrt(dx * dx + dy * dy);
    }

    /* R1 load */
    int64_t do_points = 0;
    for (int64_t i = 0; i < n; i++) {
        float z = dem[i];
        if ((has_nodata && z == nodata) || isnan(z)) {
            flg[i] = 1 | 2 | 4;
            alt[i] = 0;
            wat[i] = 0.0;
        }
        else {
            alt[i] = ora_alt_from_f32(z);
            wat[i] = 1.0;
            do_points++;
        }
    }

    /* R2 seeds */
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t c = 0; c < cols; c++) {
            int64_t i = r * cols + c;
            if (flg[i] & 1)
                continue;
            if (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) {
                if (wat[i] > 0)
                    wat[i] = -wat[i];
                if (r == 0)
                    asp[i] = -2;
                else if (c == 0)
                    asp[i] = -4;
                else if (r == rows - 1)
                    asp[i] = -6;
                else
                    asp[i] = -8;
                heap_push(&h, i);
                flg[i] |= 2;
            }
            else {
                for (int ct = 0; ct < 8; ct++) {
                    int64_t rn = r + nextdr[ct], cn = c + nextdc[ct];
                    if (flg[rn * cols + cn] & 1) {
                        asp[i] = (int8_t)(-drain[r - rn + 1][c - cn + 1]);
                        if (wat[i] > 0)
                            wat[i] = -wat[i];
                        heap_push(&h, i);
                        flg[i] |= 2;
                        break;
                    }
                }
            }
        }
    }

    /* R3/R4 A* */
    int64_t npop = 0;
    while (h.size > 0) {
        int64_t d = heap_pop(&h);
        int64_t r = d / cols, c = d % cols;
        int32_t alt_val = alt[d];
        double slope[8];
        int32_t alt_nbr[8];
        order[npop++] = d;
        flg[d] |= 4;
        for (int ct = 0; ct < 8; ct++) {
            int64_t ur = r + nextdr[ct], uc = c + nextdc[ct];
            slope[ct] = 0.0;
            alt_nbr[ct] = 0;
            if (ur < 0 || ur >= rows || uc < 0 || uc >= cols)
                continue;
            int64_t u = ur * cols + uc;
            int in_list = (flg[u] & 2) != 0, worked = (flg[u] & 4) != 0;
            int skip_diag = 0;
            if (!worked) {
                alt_nbr[ct] = alt[u];
                slope[ct] = get_slope2(alt_val, alt_nbr[ct], dist[ct]);
            }
            if (!in_list) {
                if (ct > 3 && slope[ct] > 0) {
                    if (slope[nbr_ew[ct]] > 0) {
                        if (slope[ct] < get_slope2(alt_nbr[nbr_ew[ct]],
                                                   alt_nbr[ct], ewres))
                            skip_diag = 1;
                    }
                    if (!skip_diag && slope[nbr_ns[ct]] > 0) {
                        if (slope[ct] < get_slope2(alt_nbr[nbr_ns[ct]],
                                                   alt_nbr[ct], nsres))
                            skip_diag = 1;
                    }
                }
            }
            if (!in_list && !skip_diag) {
                heap_push(&h, u);
                flg[u] |= 2;
                asp[u] = (int8_t)drain[ur - r + 1][uc - c + 1];
            }
            else if (in_list && !worked) {
                /* neighbour is an edge cell in the list, not yet worked */
                if (asp[u] < 0) {
                    asp[u] = (int8_t)drain[ur - r + 1][uc - c + 1];
                    if (wat[d] > 0)
                        wat[d] = -wat[d];
                }
            }
        }
    }

    /* R5 do_cum: from the last popped to the first popped */
    const double threshold = 60.0;
    for (int64_t k = npop - 1; k >= 0; k--) {
        int64_t i = order[k];
        int64_t r = i / cols, c = i % cols;
        int a = asp[i];
        if (!a)
            continue;
        int64_t dr = r + asp_r[abs(a)], dc = c + asp_c[abs(a)];
        if (dr < 0 || dr >= rows || dc < 0 || dc >= cols)
            continue;
        int64_t down = dr * cols + dc;
        double value = wat[i], valued = wat[down];
        int is_swale = fabs(value) >= threshold;
        if (edge_rule) {
            int edge = 0;
            int64_t rn = 0, cn = 0;
            for (int ct = 0; ct < 8; ct++) {
                rn = r + nextdr[ct];
                cn = c + nextdc[ct];
                if (rn >= 0 && rn < rows && cn >= 0 && cn < cols) {
                    if (flg[rn * cols + cn] & 1)
                        edge = 1;
                }
                else
                    edge = 1;
                if (edge)
                    break;
            }
            if (edge) {
                if (is_swale && a > 0)
                    asp[i] = (int8_t)(-drain[r - rn + 1][c - cn + 1]);
                continue;
            }
        }
        if (value > 0) {
            if (valued > 0)
                valued += value;
            else
                valued -= value;
        }
        else {
            if (valued < 0)
                valued += value;
            else
                valued = value - valued;
        }
        wat[down] = valued;
    }

    /* R6 output */
    for (int64_t i = 0; i < n; i++) {
        if (flg[i] & 1) {
            if (dir_out)
                dir_out[i] = ORA_DIR_NULL;
            if (acc_out)
                acc_out[i] = NAN;
        }
        else {
            if (dir_out)
                dir_out[i] = asp[i];
            if (acc_out)
                acc_out[i] = flag_a ? fabs(wat[i]) : wat[i];
        }
    }
    if (order_out) {
        memcpy(order_out, order, npop * sizeof(int64_t));
        for (int64_t k = npop; k < n; k++)
            order_out[k] = -1;
    }
    free(alt);
    free(asp);
    free(wat);
    free(flg);
    free(order);
    free(h.idx);
    free(h.added);
    return 0;
}

/*
 * Conditioning check used by tests and by the bench to certify inputs:
 * counts (a) "pits": non-seed valid cells with no strictly lower valid
 * 8-neighbour, and (b) "ties": valid cells whose 3x3 window holds another valid
 * cell of the same integer alt.  With both zero, R4's pop order is ascending
 * alt and the direction of every cell is a function of its 3x3 window
 * ("local-rule corollary").
 */
int ora_check_conditioned(const float *dem, int64_t rows, int64_t cols,
                          int has_nodata, float nodata, int64_t *n_pits,
                          int64_t *n_ties)
{
    int64_t pits = 0, ties = 0;
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t c = 0; c < cols; c++) {
            float z = dem[r * cols + c];
            if ((has_nodata && z == nodata) || isnan(z))
                continue;
            int32_t a = ora_alt_from_f32(z);
            int seed = (r == 0 || c == 0 || r == rows - 1 || c == cols - 1);
            int lower = 0, tie = 0;
            for (int ct = 0; ct < 8; ct++) {
                int64_t rn = r + nextdr[ct], cn = c + nextdc[ct];
                if (rn < 0 || rn >= rows || cn < 0 || cn >= cols)
                    continue;
                float zn = dem[rn * cols + cn];
                if ((has_nodata && zn == nodata) || isnan(zn)) {
                    seed = 1;
                    continue;
                }
                int32_t an = ora_alt_from_f32(zn);
                if (an < a)
                    lower = 1;
                if (an == a)
                    tie = 1;
            }
            /* ties between two neighbours of the same cell */
            for (int p = 0; p < 8 && !tie; p++) {
                int64_t rp = r + nextdr[p], cp = c + nextdc[p];
                if (rp < 0 || rp >= rows || cp < 0 || cp >= cols)
                    continue;
                float zp = dem[rp * cols + cp];
                if ((has_nodata && zp == nodata) || isnan(zp))
                    continue;
                for (int q = p + 1; q < 8; q++) {
                    int64_t rq = r + nextdr[q], cq = c + nextdc[q];
                    if (rq < 0 || rq >= rows || cq < 0 || cq >= cols)
                        continue;
                    float zq = dem[rq * cols + cq];
                    if ((has_nodata && zq == nodata) || isnan(zq))
                        continue;
                    if (ora_alt_from_f32(zp) == ora_alt_from_f32(zq)) {
                        tie = 1;
                        break;
                    }
                }
            }
            if (!seed && !lower)
                pits++;
            if (tie)
                ties++;
        }
    }
    *n_pits = pits;
    *n_ties = ties;
    return 0;
}

/*
 * Test helper: make an arbitrary DEM "hydrologically conditioned" the way
 * north_star stipulates, by replacing every valid cell's elevation with its
 * rank in r.watershed's own A* pop order (millimetres).  Ranks are unique
 * (tie-free) and every non-seed cell was reached from an earlier-popped
 * neighbour (pit-free).  rank < 4e6 keeps float32 <-> int-mm exact.
 */
int ora_condition_by_rank(const float *dem, int64_t rows, int64_t cols,
                          int has_nodata, float nodata, double ewres,
                          double nsres, float *out)
{
    int64_t n = rows * cols;
    int64_t *order = (int64_t *)malloc(n * sizeof(int64_t));
    if (!order)
        return -1;
    int rc = ora_rwatershed_sfd(dem, rows, cols, has_nodata, nodata, ewres,
                                nsres, 1, 1, NULL, NULL, order);
    if (rc) {
        free(order);
        return rc;
    }
    for (int64_t i = 0; i < n; i++)
        out[i] = has_nodata ? nodata : NAN;
    for (int64_t k = 0; k < n && order[k] >= 0; k++)
        out[order[k]] = (float)((double)(k + 1) / 1000.0);
    free(order);
    return 0;
}
