// tests/compat/drv2d_cuda.cpp -- test driver (not product code) for the CUDA_TESI faces of the 2D shim
// (-DMG_COMPAT_CUDA_TESI): one V(2,2) cycle on the finest two levels BY HAND through the (pointer, size, pitch) operators in the
// call sequence of the twin's MultiGrid2D::VCycle (CUDA_TESI/CUDA Lyapunov 2D/MultiGrid2D.cu:143-176), against the object's VCycle.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "MultiGrid2D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 65;
    float range[4] = {0, 1, 0, 1};
    float A[16] = {-1, -2, 0, -3};
    MultiGrid2D a(n, range, A, 2, 2), b(n, range, A, 2, 2);
    a.VCycle(0, 2, 2);

    Grid2D *F = b.grids2D[0], *C = b.grids2D[1];
    b.Relax(F, 2);
    float* d_res = b.CalculateResidual(F);
    b.Restrict(d_res, F->size, F->d_pitch, C->d_f, C->size, C->d_pitch);
    b.Set(C->d_v, C->size, C->d_pitch, 0.0f, true);
    b.VCycle(1, 2, 2);
    float* d_err = 0;
    size_t bytes = (size_t)F->d_pitch * F->size * sizeof(float);
    if (cudaMalloc((void**)&d_err, bytes) != cudaSuccess) return 2;
    cudaMemset(d_err, 0, bytes);
    b.Interpolate(d_err, F->size, F->d_pitch, C->d_v, C->size, C->d_pitch);
    b.ApplyCorrection(F->d_v, F->size, F->d_pitch, d_err, F->size, F->d_pitch);
    b.Relax(F, 2);
    cudaFree(d_res);
    cudaFree(d_err);

    size_t row = (size_t)n * sizeof(float);
    float* ha = (float*)malloc(row * n);
    float* hb = (float*)malloc(row * n);
    cudaMemcpy2D(ha, row, a.grids2D[0]->d_v, a.grids2D[0]->d_pitchByte, row, n, cudaMemcpyDeviceToHost);
    cudaMemcpy2D(hb, row, F->d_v, F->d_pitchByte, row, n, cudaMemcpyDeviceToHost);
    int same = memcmp(ha, hb, row * n) == 0;
    double nrm = 0;
    for (size_t i = 0; i < (size_t)n * n; i++) nrm += (double)ha[i] * ha[i];
    printf("CUDA_FACE %s hand-made V-cycle == VCycle: %d, |v|^2 = %.9e, d_matrixA %p sizeX_A %d\n", same && nrm > 0 ? "OK" : "FAILED", same, nrm,
           (void*)b.d_matrixA, b.sizeX_A);
    return same ? 0 : 1;
}
