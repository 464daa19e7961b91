// ht_dir.cu -- conversions between GRASS direction codes (what the reference's
// rasters hold: 1..8 CCW from E starting NE, <= 0 terminal, NULL;
// /root/reference/src/hydrotools/watershed.py:575-589) and the packed device
// byte, and the r.watershed sign step for positive_only=False
// (/root/reference/src/hydrotools/watershed.py:52-54: flag "a" absent).
#include "ht_common.cuh"

namespace {

__global__ void pack_i8_kernel(const int8_t *dir, int64_t ld, int64_t rows, int64_t cols,
                               uint8_t *packed, int64_t packed_ld)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const int d = dir[r * ld + c];
        uint8_t b;
        if (d == -128)
            b = HT_DIR_NULL;
        else if (d >= 1 && d <= 8) {
            // a receiver outside the raster or NULL ends the path: the cell "does not
            // pass flow on" (the accumulation kernels then never re-check a target)
            const int64_t rr = r + (int)((0x1A900u >> (2 * d)) & 3u) - 1;
            const int64_t cc = c + (int)((0x29018u >> (2 * d)) & 3u) - 1;
            const bool gone = rr < 0 || rr >= rows || cc < 0 || cc >= cols ||
                              dir[rr * ld + cc] == -128;
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
    pack_i8_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(dir, ld, rows, cols, packed, packed_ld);
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
