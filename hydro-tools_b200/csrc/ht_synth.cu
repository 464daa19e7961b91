// ht_synth.cu -- device-side generator of the synthetic conditioned fractal DEM
// and precipitation-like weights (BASELINE.json configs 2-5).
// Same integer function as the CPU build (include/ht_synth.h), so each GPU can
// generate its own row band with no host traffic.
#include "ht_common.cuh"
#include "../../include/ht_synth.h"

namespace {

__global__ void synth_dem_kernel(float *dem, int64_t rows, int64_t cols, int64_t ld,
                                 ht_synth_params p, int64_t row0, float nodata)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        dem[r * ld + c] = ht_synth_is_null(&p, row0 + r, c)
                              ? nodata
                              : ht_synth_dem_value(&p, row0 + r, c);
    }
}

__global__ void synth_weight_kernel(float *w, int64_t rows, int64_t cols, int64_t ld,
                                    uint32_t seed, int64_t row0)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        w[r * ld + c] = ht_synth_weight_value(seed, row0 + r, c);
    }
}

} // namespace

extern "C" int ht_synth_dem_f32(float *dem, int64_t rows, int64_t cols, int64_t ld,
                                uint64_t seed, int64_t row0, int64_t total_rows,
                                int with_nulls, float nodata, void *stream)
{
    if (!dem || rows < 0 || cols < 0 || ld < cols || total_rows < row0 + rows)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    ht_synth_params p;
    ht_synth_make_params((uint32_t)seed, total_rows, with_nulls, &p);
    synth_dem_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(dem, rows, cols, ld, p,
                                                                    row0, nodata);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_synth_weight_f32(float *w, int64_t rows, int64_t cols, int64_t ld,
                                   uint64_t seed, int64_t row0, void *stream)
{
    if (!w || rows < 0 || cols < 0 || ld < cols)
        return HT_ERR_ARG;
    if (rows == 0 || cols == 0)
        return HT_OK;
    synth_weight_kernel<<<ht::kSMs * 8, 256, 0, (cudaStream_t)stream>>>(w, rows, cols, ld,
                                                                       (uint32_t)seed, row0);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
