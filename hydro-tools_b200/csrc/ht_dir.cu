// ht_dir.cu -- conversions between GRASS direction codes (what the reference's
// rasters hold: 1..8 CCW from E starting NE, <= 0 terminal, NULL;
// /root/reference/src/hydrotools/watershed.py:575-589) and the packed device
// byte, and the r.watershed sign step for positive_only=False
// (/root/reference/src/hydrotools/watershed.py:52-54: flag "a" absent).
#include "ht_common.cuh"

namespace {

// rows [-halo_top, rows + halo_bot) of the allocation are packed (halo rows = the edge rows of
// the neighbour bands, which the accumulation stages read as their ring); `dir` points at
// owned row 0, global row = row0 + local row
__global__ void pack_i8_kernel(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                               int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                               uint8_t *packed, int64_t packed_ld)
{
    const int64_t n = (rows + halo_top + halo_bot) * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols - halo_top, c = i - (r + halo_top) * cols;
        const int d = dir[r * ld + c];
        uint8_t b;
        if (d == -128)
            b = HT_DIR_NULL;
        else if (d >= 1 && d <= 8) {
            // a receiver outside the raster or NULL ends the path: the cell "does not
            // pass flow on" (the accumulation kernels then never re-check a target)
            const int64_t rr = r + (int)((0x1A900u >> (2 * d)) & 3u) - 1;
            const int64_t cc = c + (int)((0x29018u >> (2 * d)) & 3u) - 1;
            bool gone = row0 + rr < 0 || row0 + rr >= total_rows || cc < 0 || cc >= cols;
            // a halo row's receiver may lie beyond the rows this band holds: nobody asks
            if (!gone && rr >= -halo_top && rr < rows + halo_bot)
                gone = dir[rr * ld + cc] == -128;
            b = (uint8_t)d | (gone ? HT_DIR_EDGE : 0);
        }
        else if (d <= -1 && d >= -8)
            b = (uint8_t)(-d) | HT_DIR_NEG;
        else
            b = 0;
        packed[r * packed_ld + c] = b;
    }
}

__global__ void unpack_i8_kernel(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                                 int64_t cols, int8_t *dir, int64_t ld)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const uint8_t b = packed[r * packed_ld + c];
        const int code = b & HT_DIR_CODE;
        dir[r * ld + c] = (b & HT_DIR_NULL) ? (int8_t)-128 : (b & HT_DIR_NEG) ? (int8_t)-code : (int8_t)code;
    }
}

// r.watershed without -a: accumulation is negative ("likely underestimate") at
// every edge cell and at every cell with a tainted cell upstream
__global__ void apply_sign_kernel(double *acc, int64_t acc_ld, const double *taint,
                                  int64_t taint_ld, const uint8_t *packed, int64_t packed_ld,
                                  int64_t rows, int64_t cols)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const uint8_t b = packed[r * packed_ld + c];
        if (b & HT_DIR_NULL)
            continue;
        if ((b & HT_DIR_EDGE) || taint[r * taint_ld + c] > 0.0)
            acc[r * acc_ld + c] = -acc[r * acc_ld + c];
    }
}

} // namespace

extern "C" int ht_dir_pack_i8(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                              uint8_t *packed, int64_t packed_ld, void *stream)
{
    if (!dir || !packed || rows < 0 || cols < 0 || ld < cols || packed_ld < cols)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    pack_i8_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(dir, ld, rows, cols, 0, rows, 0, 0,
                                                                  packed, packed_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

// Row-band form: `dir` / `packed` point at the band's first OWNED row; with halo_top /
// halo_bot the edge row of the neighbour band sits directly above / below in both
// allocations (the direction row arrives by exchange, the packed row is written here).
extern "C" int ht_dir_pack_band_i8(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                                   int64_t row0, int64_t total_rows, int halo_top, int halo_bot,
                                   uint8_t *packed, int64_t packed_ld, void *stream)
{
    if (!dir || !packed || rows < 0 || cols < 0 || ld < cols || packed_ld < cols || row0 < 0 ||
        total_rows < row0 + rows || (!halo_top && row0 != 0) || (!halo_bot && row0 + rows != total_rows))
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    pack_i8_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(
        dir, ld, rows, cols, row0, total_rows, halo_top ? 1 : 0, halo_bot ? 1 : 0, packed, packed_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_dir_unpack_i8(const uint8_t *packed, int64_t packed_ld, int64_t rows,
                                int64_t cols, int8_t *dir, int64_t ld, void *stream)
{
    if (!dir || !packed || rows < 0 || cols < 0 || ld < cols || packed_ld < cols)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    unpack_i8_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(packed, packed_ld, rows, cols, dir, ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_acc_apply_sign(double *acc, int64_t acc_ld, const double *taint_acc,
                                 int64_t taint_ld, const uint8_t *packed, int64_t packed_ld,
                                 int64_t rows, int64_t cols, void *stream)
{
    if (!acc || !taint_acc || !packed || rows < 0 || cols < 0 || acc_ld < cols ||
        taint_ld < cols || packed_ld < cols)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    apply_sign_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(acc, acc_ld, taint_acc, taint_ld,
                                                                     packed, packed_ld, rows, cols);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
