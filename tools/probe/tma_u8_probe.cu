// probe: which uint8 TMA box shapes work on sm_100a? usage: probe BOXW BOXH X Y
#include <cstdio>
#include <cstdlib>
#include "../../hydro-tools_b200/csrc/ht_common.cuh"
__global__ void k(const __grid_constant__ CUtensorMap map, int bw, int bh, int x, int y, unsigned *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        ht::mbar_init(&bar, 1);
        ht::fence_mbar_init();
        ht::mbar_expect_tx(&bar, bw * bh);
        ht::tma_load_2d(sm, &map, x, y, &bar);
    }
    __syncthreads();
    ht::mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) s += sm[i];
    atomicAdd(out, s);
}
int main(int argc, char **argv)
{
    int bw = atoi(argv[1]), bh = atoi(argv[2]), x = atoi(argv[3]), y = atoi(argv[4]);
    int rows = 100, cols = 150, ld = 160;
    unsigned char *d; unsigned *out;
    cudaMalloc(&d, rows * ld); cudaMemset(d, 1, rows * ld);
    cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    CUtensorMap map;
    int rc = ht::make_map_2d(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, d, rows, cols, ld, bw, bh, false);
    printf("box %dx%d at (%d,%d): encode rc=%d ", bw, bh, x, y, rc);
    if (rc) { printf("\n"); return 0; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<1, 128, 65536>>>(map, bw, bh, x, y, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
    printf("run: %s sum=%u\n", cudaGetErrorString(e), h);
    return 0;
}
