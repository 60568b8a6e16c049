// tests/compat/drv1d_cuda.cpp -- test driver (not product code) for the CUDA_TESI faces of the 1D shim
// (-DMG_COMPAT_CUDA_TESI): one V(50,50) cycle on the finest two levels BY HAND through the device-pointer operators in the
// call sequence of the twin's MultiGrid1D::VCycle (CUDA_TESI/CUDA 1D/MultiGrid1D.cu:118-148), against the object's VCycle.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "MultiGrid1D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 257, nu = 50;
    float range[2] = {0, 1};
    MultiGrid1D a(n, range), b(n, range);
    a.VCycle(0, nu, nu);

    Grid1D *F = b.grids1D[0], *C = b.grids1D[1];
    b.Relax(F, nu);
    float* d_res = b.CalculateResidual(F);
    b.Restrict(d_res, F->sizeX, C->d_f, C->sizeX);
    b.Set(C->d_v, C->sizeX, 0.0f, true);
    b.VCycle(1, nu, nu);
    float* d_err = 0;
    if (cudaMalloc((void**)&d_err, (size_t)n * sizeof(float)) != cudaSuccess) return 2;
    cudaMemset(d_err, 0, (size_t)n * sizeof(float));
    b.Interpolate(d_err, F->sizeX, C->d_v, C->sizeX);
    b.ApplyCorrection(F->d_v, F->sizeX, d_err, F->sizeX);
    b.Relax(F, nu);
    cudaFree(d_res);
    cudaFree(d_err);

    float* ha = (float*)malloc((size_t)n * sizeof(float));
    float* hb = (float*)malloc((size_t)n * sizeof(float));
    cudaMemcpy(ha, a.grids1D[0]->d_v, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    cudaMemcpy(hb, F->d_v, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost);
    int same = memcmp(ha, hb, (size_t)n * sizeof(float)) == 0;
    double nrm = 0;
    for (int i = 0; i < n; i++) nrm += (double)ha[i] * ha[i];
    printf("CUDA_FACE %s hand-made V-cycle == VCycle: %d, |v|^2 = %.9e\n", same && nrm > 0 ? "OK" : "FAILED", same, nrm);
    return same ? 0 : 1;
}
