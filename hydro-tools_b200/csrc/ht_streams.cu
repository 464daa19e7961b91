// ht_streams.cu -- downstream consumers of the D8 direction grid: the
// reference walks the grid cell by cell in numba; here the same results come from pointer
// jumping and path walkers on the device.
//   * ht_reach_labels_i32     replaces `_fa_label_task`
//       (/root/reference/src/hydrotools/streams.py:320-388): connected runs of cells with
//       one reach value, connected along direction links, numbered in row-major order of
//       their first cell;
//   * ht_upstream_mean_f64    replaces `_upstream_compute`
//       (/root/reference/src/hydrotools/morphology.py:50-153): from every stream inlet, a
//       moving mean over the last `distance` map units along the path downstream; where the
//       paths of several inlets overlap the reference's last writer (the inlet latest in
//       row-major order) prevails;
//   * ht_basin_u8             replaces GRASS `r.water.outlet`
//       (/root/reference/src/hydrotools/watershed.py:292-318): every cell whose path passes
//       through the outlet cell.
// Directions are int8 GRASS codes (1 = NE ... 8 = E counter-clockwise, <= 0 / -128 end a path).
#include "ht_common.cuh"

namespace {

__device__ __forceinline__ int code_dr(int code) { return (int)((0x1A901u >> (2 * code)) & 3u) - 1; }
__device__ __forceinline__ int code_dc(int code) { return (int)((0x29019u >> (2 * code)) & 3u) - 1; }

// receiver of cell (r, c), or -1: code outside 1..8 or receiver outside the raster
__device__ __forceinline__ int64_t receiver(const int8_t *fd, int64_t fd_ld, int64_t rows, int64_t cols,
                                            int64_t r, int64_t c)
{
    const int d = fd[r * fd_ld + c];
    if (d < 1 || d > 8)
        return -1;
    const int64_t rr = r + code_dr(d), cc = c + code_dc(d);
    if (rr < 0 || rr >= rows || cc < 0 || cc >= cols)
        return -1;
    return rr * cols + cc;
}

// ------------------------------------------------------------------ reach labels
__global__ void reach_parent_kernel(const int32_t *reaches, int64_t r_ld, const int8_t *fd, int64_t fd_ld,
                                    int64_t rows, int64_t cols, int32_t nodata, int32_t *parent,
                                    int32_t *minidx)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const int32_t v = reaches[r * r_ld + c];
        int32_t p = -1; // not a reach cell
        if (v != nodata) {
            p = (int32_t)i;
            const int64_t j = receiver(fd, fd_ld, rows, cols, r, c);
            if (j >= 0 && reaches[(j / cols) * r_ld + j % cols] == v)
                p = (int32_t)j;
        }
        parent[i] = p;
        minidx[i] = 0x7FFFFFFF;
    }
}

// pointer jumping: parent <- parent of parent, until every cell points at the last cell of
// its run (a run is a subtree of the direction forest, so it has exactly one)
__global__ void jump_kernel(int32_t *parent, int64_t n, unsigned long long *changed)
{
    bool any = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = parent[i];
        if (p < 0)
            continue;
        const int32_t q = parent[p];
        if (q != p) {
            parent[i] = q;
            any = true;
        }
    }
    if (any)
        atomicAdd(changed, 1ull);
}

__global__ void reach_min_kernel(const int32_t *parent, int64_t n, int32_t *minidx)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        if (parent[i] >= 0)
            atomicMin(minidx + parent[i], (int32_t)i);
}

// block-wise exclusive scan of "cell i is the first cell of its run" (1024 cells per block)
constexpr int SCAN_BLOCK = 1024;
__global__ void __launch_bounds__(256)
first_scan_kernel(const int32_t *parent, const int32_t *minidx, int64_t n, uint32_t *rank,
                  uint32_t *block_sum)
{
    __shared__ uint32_t warp_tot[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    uint32_t f[4], run = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int64_t i = base + k;
        f[k] = (i < n && parent[i] >= 0 && minidx[parent[i]] == (int32_t)i) ? 1u : 0u;
        run += f[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = run;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d)
            inc += t;
    }
    if (lane == 31)
        warp_tot[warp] = inc;
    __syncthreads();
    uint32_t off = 0;
    for (int w = 0; w < warp; w++)
        off += warp_tot[w];
    uint32_t ex = off + inc - run;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n)
            rank[base + k] = ex;
        ex += f[k];
    }
    if (threadIdx.x == 255)
        block_sum[blockIdx.x] = off + inc;
}

__global__ void block_offsets_kernel(uint32_t *block_sum, int64_t nblocks)
{
    // one block: sequential chunks of 1024 block sums
    __shared__ uint32_t carry, tot[1024];
    if (threadIdx.x == 0)
        carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        tot[threadIdx.x] = i < nblocks ? block_sum[i] : 0u;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) {
            const uint32_t t = threadIdx.x >= d ? tot[threadIdx.x - d] : 0u;
            __syncthreads();
            tot[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks)
            block_sum[i] = carry + tot[threadIdx.x] - block_sum[i]; // exclusive
        __syncthreads();
        if (threadIdx.x == 1023)
            carry += tot[1023];
        __syncthreads();
    }
}

__global__ void reach_label_kernel(const int32_t *parent, const int32_t *minidx, const uint32_t *rank,
                                   const uint32_t *block_off, int64_t rows, int64_t cols,
                                   uint32_t *labels, int64_t l_ld)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t p = parent[i];
        uint32_t lab = 0;
        if (p >= 0) {
            const int32_t first = minidx[p];
            lab = 1u + rank[first] + block_off[first / SCAN_BLOCK];
        }
        labels[(i / cols) * l_ld + i % cols] = lab;
    }
}

// ------------------------------------------------------------------ upstream moving mean
__device__ __forceinline__ bool on_stream(const uint8_t *streams, int64_t s_ld, int64_t r, int64_t c)
{
    return streams[r * s_ld + c] != 0;
}

// inlet = stream cell without a stream cell draining into it (morphology.py:80-103)
__global__ void inlet_kernel(const uint8_t *streams, int64_t s_ld, const int8_t *fd, int64_t fd_ld,
                             int64_t rows, int64_t cols, int32_t *inlets, unsigned long long *count,
                             int32_t *owner)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        owner[i] = -1;
        const int64_t r = i / cols, c = i - r * cols;
        if (!on_stream(streams, s_ld, r, c))
            continue;
        bool inlet = true;
        for (int di = -1; di <= 1 && inlet; di++)
            for (int dj = -1; dj <= 1; dj++) {
                if (!di && !dj)
                    continue;
                const int64_t rr = r + di, cc = c + dj;
                if (rr < 0 || rr >= rows || cc < 0 || cc >= cols || !on_stream(streams, s_ld, rr, cc))
                    continue;
                if (receiver(fd, fd_ld, rows, cols, rr, cc) == i) {
                    inlet = false;
                    break;
                }
            }
        if (inlet)
            inlets[atomicAdd(count, 1ull)] = (int32_t)i;
    }
}

// pass 1: the last inlet (row-major) whose path passes through each cell
__global__ void owner_kernel(const uint8_t *streams, int64_t s_ld, const int8_t *fd, int64_t fd_ld,
                             int64_t rows, int64_t cols, const int32_t *inlets, int64_t ninlets,
                             int32_t *owner)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ninlets;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t me = inlets[k];
        int64_t cur = me;
        for (int64_t guard = 0; guard < rows * cols; guard++) {
            const int64_t nx = receiver(fd, fd_ld, rows, cols, cur / cols, cur % cols);
            if (nx < 0 || !on_stream(streams, s_ld, nx / cols, nx % cols))
                break;
            // a later inlet has been here: it owns everything further down as well
            if (atomicMax(owner + nx, me) > me)
                break;
            cur = nx;
        }
    }
}

// pass 2: the reference's walk, value by value; an inlet writes the cells it owns
__global__ void upstream_walk_kernel(const double *values, int64_t v_ld, const uint8_t *streams,
                                     int64_t s_ld, const int8_t *fd, int64_t fd_ld, int64_t rows,
                                     int64_t cols, double csx, double csy, double distance,
                                     int max_window, const int32_t *inlets, int64_t ninlets,
                                     const int32_t *owner, double *val_buf, double *dist_buf,
                                     float *out, int64_t o_ld)
{
    const double diag = sqrt(csx * csx + csy * csy);
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < ninlets;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int32_t me = inlets[k];
        double *vb = val_buf + k * max_window, *db = dist_buf + k * max_window;
        int head = 0, tail = 1, size = 1;
        double cum = 0.0;
        int64_t cur = me;
        vb[0] = values[(cur / cols) * v_ld + cur % cols];
        db[0] = 0.0;
        double sum = vb[0];
        for (int64_t guard = 0; guard < rows * cols; guard++) {
            const int64_t r = cur / cols, c = cur % cols;
            const int d = fd[r * fd_ld + c];
            const int64_t nx = receiver(fd, fd_ld, rows, cols, r, c);
            if (nx < 0 || !on_stream(streams, s_ld, nx / cols, nx % cols))
                break;
            const int di = code_dr(d), dj = code_dc(d);
            cum += (di != 0 && dj != 0) ? diag : (dj != 0 ? csx : csy);
            const double v = values[(nx / cols) * v_ld + nx % cols];
            vb[tail] = v;
            db[tail] = cum;
            tail = (tail + 1) % max_window;
            size++;
            sum += v;
            while (cum - db[head] > distance) {
                sum -= vb[head];
                head = (head + 1) % max_window;
                size--;
            }
            const int32_t own = owner[nx];
            if (own == me)
                out[(nx / cols) * o_ld + nx % cols] = (float)(sum / size);
            else if (own > me)
                break; // everything further down belongs to later inlets
            cur = nx;
        }
    }
}

// ------------------------------------------------------------------ basin
__global__ void basin_parent_kernel(const int8_t *fd, int64_t fd_ld, int64_t rows, int64_t cols,
                                    int64_t outlet, int32_t *parent)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = i == outlet ? -1 : receiver(fd, fd_ld, rows, cols, i / cols, i % cols);
        // NULL cells are no part of any basin; paths stop at the outlet
        parent[i] = fd[(i / cols) * fd_ld + i % cols] == -128 ? -1 : (j >= 0 ? (int32_t)j : (int32_t)i);
    }
}

__global__ void basin_mark_kernel(const int32_t *parent, int64_t rows, int64_t cols, int64_t outlet,
                                  uint8_t *out, int64_t o_ld)
{
    const int64_t n = rows * cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[(i / cols) * o_ld + i % cols] = parent[i] == (int32_t)outlet ? 1 : 0;
}

int grid_for(int64_t n) { return (int)min((int64_t)ht::kSMs * 16, (n + 255) / 256 > 0 ? (n + 255) / 256 : 1); }
size_t align_up(size_t x) { return (x + 255) / 256 * 256; }

// pointer jumping until nothing moves (a run / path of L cells needs ~log2 L sweeps)
int jump_to_roots(int32_t *parent, int64_t n, unsigned long long *flag, cudaStream_t st)
{
    for (int it = 0; it < 64; it++) {
        HT_CUDA_TRY(cudaMemsetAsync(flag, 0, 8, st));
        jump_kernel<<<grid_for(n), 256, 0, st>>>(parent, n, flag);
        HT_LAUNCH_CHECK();
        unsigned long long h = 0;
        HT_CUDA_TRY(cudaMemcpyAsync(&h, flag, 8, cudaMemcpyDeviceToHost, st));
        HT_CUDA_TRY(cudaStreamSynchronize(st));
        if (!h)
            return 0;
    }
    return HT_ERR_ARG; // a cycle in the direction raster
}

} // namespace

// labels / basin: max_window = max_inlets = 0.  upstream mean: two ring buffers of max_window
// doubles per inlet; max_inlets = any upper bound (the number of stream cells will do)
extern "C" size_t ht_streams_workspace_bytes(int64_t rows, int64_t cols, int max_window,
                                             int64_t max_inlets)
{
    if (rows <= 0 || cols <= 0)
        return 256;
    const size_t n = (size_t)rows * cols;
    const size_t win = max_window > 0 && max_inlets > 0 ? (size_t)max_window * (size_t)max_inlets : 0;
    return 3 * align_up(n * 4) + align_up((n / 1024 + 2) * 4) + 2 * align_up(win * 8) + 1024;
}

extern "C" int ht_reach_labels_i32(const int32_t *reaches, int64_t r_ld, const int8_t *fd, int64_t fd_ld,
                                   int64_t rows, int64_t cols, int32_t nodata, uint32_t *labels,
                                   int64_t l_ld, void *ws, size_t ws_bytes, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!reaches || !fd || !labels || rows <= 0 || cols <= 0 || r_ld < cols || fd_ld < cols || l_ld < cols ||
        rows * cols >= 0x7FFFFFF0ll)
        return HT_ERR_ARG;
    if (!ws || ws_bytes < ht_streams_workspace_bytes(rows, cols, 0, 0))
        return HT_ERR_WORKSPACE;
    const int64_t n = rows * cols, nblocks = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    int32_t *parent = (int32_t *)b; b += align_up(n * 4);
    int32_t *minidx = (int32_t *)b; b += align_up(n * 4);
    uint32_t *rank = (uint32_t *)b; b += align_up(n * 4);
    uint32_t *block_sum = (uint32_t *)b; b += align_up((n / 1024 + 2) * 4);
    unsigned long long *flag = (unsigned long long *)b;
    reach_parent_kernel<<<grid_for(n), 256, 0, st>>>(reaches, r_ld, fd, fd_ld, rows, cols, nodata, parent, minidx);
    HT_LAUNCH_CHECK();
    int rc = jump_to_roots(parent, n, flag, st);
    if (rc)
        return rc;
    reach_min_kernel<<<grid_for(n), 256, 0, st>>>(parent, n, minidx);
    first_scan_kernel<<<(unsigned)nblocks, 256, 0, st>>>(parent, minidx, n, rank, block_sum);
    block_offsets_kernel<<<1, 1024, 0, st>>>(block_sum, nblocks);
    reach_label_kernel<<<grid_for(n), 256, 0, st>>>(parent, minidx, rank, block_sum, rows, cols, labels, l_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_upstream_mean_f64(const double *values, int64_t v_ld, const uint8_t *streams,
                                    int64_t s_ld, const int8_t *fd, int64_t fd_ld, int64_t rows,
                                    int64_t cols, double csx, double csy, double distance,
                                    int max_window, int64_t max_inlets, float *out, int64_t o_ld,
                                    void *ws, size_t ws_bytes, long long *ninlets_host, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!values || !streams || !fd || !out || rows <= 0 || cols <= 0 || v_ld < cols || s_ld < cols ||
        fd_ld < cols || o_ld < cols || max_window < 2 || !(distance >= 0.0) || csx <= 0.0 || csy <= 0.0 ||
        rows * cols >= 0x7FFFFFF0ll)
        return HT_ERR_ARG;
    if (!ws || max_inlets < 0 || ws_bytes < ht_streams_workspace_bytes(rows, cols, max_window, max_inlets))
        return HT_ERR_WORKSPACE;
    const int64_t n = rows * cols;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    int32_t *owner = (int32_t *)b; b += align_up(n * 4);
    int32_t *inlets = (int32_t *)b; b += align_up(n * 4);
    unsigned long long *count = (unsigned long long *)b; b += align_up(n * 4);
    b += align_up((n / 1024 + 2) * 4);
    double *val_buf = (double *)b; b += align_up((size_t)max_inlets * (size_t)max_window * 8);
    double *dist_buf = (double *)b;
    HT_CUDA_TRY(cudaMemsetAsync(count, 0, 8, st));
    inlet_kernel<<<grid_for(n), 256, 0, st>>>(streams, s_ld, fd, fd_ld, rows, cols, inlets, count, owner);
    HT_LAUNCH_CHECK();
    unsigned long long h = 0;
    HT_CUDA_TRY(cudaMemcpyAsync(&h, count, 8, cudaMemcpyDeviceToHost, st));
    HT_CUDA_TRY(cudaStreamSynchronize(st));
    if (ninlets_host)
        *ninlets_host = (long long)h;
    if (h == 0)
        return HT_OK;
    if (h > (unsigned long long)max_inlets)
        return HT_ERR_WORKSPACE; // *ninlets_host tells how many there are
    const int64_t ni = (int64_t)h;
    const int grid = (int)min((int64_t)ht::kSMs * 8, (ni + 63) / 64);
    owner_kernel<<<grid, 64, 0, st>>>(streams, s_ld, fd, fd_ld, rows, cols, inlets, ni, owner);
    upstream_walk_kernel<<<grid, 64, 0, st>>>(values, v_ld, streams, s_ld, fd, fd_ld, rows, cols, csx, csy,
                                              distance, max_window, inlets, ni, owner, val_buf, dist_buf,
                                              out, o_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}

extern "C" int ht_basin_u8(const int8_t *fd, int64_t fd_ld, int64_t rows, int64_t cols, int64_t outlet_row,
                           int64_t outlet_col, uint8_t *out, int64_t o_ld, void *ws, size_t ws_bytes,
                           void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!fd || !out || rows <= 0 || cols <= 0 || fd_ld < cols || o_ld < cols || outlet_row < 0 ||
        outlet_row >= rows || outlet_col < 0 || outlet_col >= cols || rows * cols >= 0x7FFFFFF0ll)
        return HT_ERR_ARG;
    if (!ws || ws_bytes < ht_streams_workspace_bytes(rows, cols, 0, 0))
        return HT_ERR_WORKSPACE;
    const int64_t n = rows * cols, outlet = outlet_row * cols + outlet_col;
    uint8_t *b = (uint8_t *)(((uintptr_t)ws + 255) / 256 * 256);
    int32_t *parent = (int32_t *)b; b += align_up(n * 4);
    unsigned long long *flag = (unsigned long long *)b;
    basin_parent_kernel<<<grid_for(n), 256, 0, st>>>(fd, fd_ld, rows, cols, outlet, parent);
    HT_LAUNCH_CHECK();
    int rc = jump_to_roots(parent, n, flag, st);
    if (rc)
        return rc;
    basin_mark_kernel<<<grid_for(n), 256, 0, st>>>(parent, rows, cols, outlet, out, o_ld);
    HT_LAUNCH_CHECK();
    return HT_OK;
}
