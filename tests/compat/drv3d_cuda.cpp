// tests/compat/drv3d_cuda.cpp -- test driver (not product code) for the CUDA_TESI faces of the shim (-DMG_COMPAT_CUDA_TESI):
// one V(2,2) cycle on the finest two levels done BY HAND through the operator methods with device pointers, exactly the call
// sequence of the twin's MultiGrid3D::VCycle (CUDA_TESI/CUDA Poisson 3D/MultiGrid3D.cu:270-300): Relax, CalculateResidual
// (device array, caller-owned), Restrict(d_residual, d_fsizeXYZ, coarse->d_f, d_csizeXYZ), Set(coarse->d_v, d_csizeXYZ, 0, true),
// VCycle(gridID+1), Interpolate into a cudaMalloc'ed error grid, ApplyCorrection, Relax -- against the object's own VCycle.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "MultiGrid3D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 33;
    int s[3] = {n, n, n};
    float range[6] = {0, 1, 0, 1, 0, 1};
    MultiGrid3D a(s, range), b(s, range);
    // interior v = 0 on the finest two levels of both (the twin leaves it uninitialised, C3/Grid3D.cu:220-224)
    for (int l = 0; l < 2; l++) {
        a.Set(a.grids3D[l]->d_v, a.grids3D[l]->d_sizeXYZ, 0.0f, true);
        b.Set(b.grids3D[l]->d_v, b.grids3D[l]->d_sizeXYZ, 0.0f, true);
    }
    a.VCycle(0, 2, 2);

    Grid3D *F = b.grids3D[0], *C = b.grids3D[1];
    b.Relax(F, 2);
    float* d_res = b.CalculateResidual(F);
    b.Restrict(d_res, F->d_sizeXYZ, C->d_f, C->d_sizeXYZ);
    b.Set(C->d_v, C->d_sizeXYZ, 0.0f, true);
    b.VCycle(1, 2, 2);
    float* d_err = 0;
    size_t bytes = (size_t)n * n * n * sizeof(float);
    if (cudaMalloc((void**)&d_err, bytes) != cudaSuccess) return 2;
    cudaMemset(d_err, 0, bytes);
    b.Interpolate(d_err, F->d_sizeXYZ, C->d_v, C->d_sizeXYZ);
    b.ApplyCorrection(F->d_v, F->d_sizeXYZ, d_err, F->d_sizeXYZ);
    b.Relax(F, 2);
    cudaFree(d_res);
    cudaFree(d_err);

    float* ha = (float*)malloc(bytes);
    float* hb = (float*)malloc(bytes);
    cudaMemcpy(ha, a.grids3D[0]->d_v, bytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb, F->d_v, bytes, cudaMemcpyDeviceToHost);
    int same = memcmp(ha, hb, bytes) == 0;
    // the index-map probe of the twin
    b.SetTESTTEST(F->d_v, F->d_sizeXYZ, 0.0f, true);
    cudaMemcpy(hb, F->d_v, bytes, cudaMemcpyDeviceToHost);
    int probe = hb[3 + 5 * n + 7 * n * n] == 15.0f;
    double nrm = 0;
    for (size_t i = 0; i < (size_t)n * n * n; i++) nrm += (double)ha[i] * ha[i];
    printf("CUDA_FACE %s hand-made V-cycle == VCycle: %d, index probe: %d, |v|^2 = %.9e\n", same && probe && nrm > 0 ? "OK" : "FAILED", same, probe, nrm);
    return same && probe ? 0 : 1;
}
