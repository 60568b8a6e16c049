// Standalone probe of the 3D TMA tile load used by mg3d_smooth_tma.cu (development aid).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../pde_multigrid_b200/csrc/mg_tma.cuh"
using namespace mgtma;

template <typename T, int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, T* out, int c0, int c1, int c2, int variant)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + ((BW * BH * sizeof(T) + 127) / 128 * 128));
    if (threadIdx.x == 0) {
        if (variant & 1) prefetch_tensormap(&map);
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, BW * BH * sizeof(T));
        tma_load_3d(tile, &map, bar, c0, c1, c2);
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T, int BW, int BH>
int run(encode_fn enc, int hp, int n, int nz, int c0, int c1, int c2, int variant)
{
    size_t tot = (size_t)hp * n * nz;
    T* h = (T*)malloc(tot * sizeof(T));
    for (size_t i = 0; i < tot; i++) h[i] = (T)(i % 100003);
    T *d, *dout;
    cudaMalloc(&d, tot * sizeof(T));
    cudaMalloc(&dout, BW * BH * sizeof(T));
    cudaMemcpy(d, h, tot * sizeof(T), cudaMemcpyHostToDevice);
    cuuint64_t dims[3] = {(cuuint64_t)hp, (cuuint64_t)n, (cuuint64_t)nz};
    cuuint64_t strides[2] = {(cuuint64_t)hp * sizeof(T), (cuuint64_t)hp * n * sizeof(T)};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUtensorMap map;
    CUresult r = enc(&map, sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("T=%zu BW=%d BH=%d hp=%d n=%d c=(%d,%d,%d) variant=%d encode=%d ", sizeof(T), BW, BH, hp, n, c0, c1, c2, variant, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    size_t smem = (BW * BH * sizeof(T) + 127) / 128 * 128 + 64;
    cudaFuncSetAttribute(probe<T, BW, BH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<T, BW, BH><<<1, 128, smem>>>(map, dout, c0, c1, c2, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        T* o = (T*)malloc(BW * BH * sizeof(T));
        cudaMemcpy(o, dout, BW * BH * sizeof(T), cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int y = 0; y < BH; y++)
            for (int x = 0; x < BW; x++) {
                int gx = c0 + x, gy = c1 + y;
                T want = (gx < 0 || gx >= hp || gy < 0 || gy >= n || c2 < 0 || c2 >= nz) ? (T)0 : h[(size_t)c2 * hp * n + (size_t)gy * hp + gx];
                if (o[y * BW + x] != want) bad++;
            }
        printf("mismatches=%ld", bad);
        free(o);
    }
    printf("\n");
    free(h);
    cudaFree(d); cudaFree(dout);
    return e != cudaSuccess;
}

int main(int argc, char** argv)
{
    int which = argc > 1 ? atoi(argv[1]) : 0;
    void* fn = NULL;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    encode_fn enc = (encode_fn)fn;
    switch (which) {
        case 0: return run<double, 132, 10>(enc, 144, 257, 257, 0, 0, 5, 0);
        case 1: return run<double, 132, 10>(enc, 144, 257, 257, -2, 0, 5, 0);
        case 2: return run<double, 132, 10>(enc, 144, 257, 257, -2, -1, 5, 1);
        case 3: return run<float, 132, 10>(enc, 160, 257, 257, 0, 0, 5, 0);
        case 4: return run<float, 136, 10>(enc, 160, 257, 257, -4, -1, 0, 0);
        case 5: return run<float, 136, 10>(enc, 160, 257, 257, 124, 250, 256, 1);
        case 6: return run<double, 128, 10>(enc, 144, 257, 257, 0, 0, 5, 0);
        case 7: return run<double, 32, 10>(enc, 144, 257, 257, 126, 255, 5, 0);
        case 8: return run<double, 132, 10>(enc, 16, 9, 9, -2, -1, 5, 0);
    }
    return 0;
}
