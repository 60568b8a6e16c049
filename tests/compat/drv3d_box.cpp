// tests/compat/drv3d_box.cpp -- test driver (not product code): the calls of the reference's 3D main
// (NOCUDA_TESI/POISSON_3D(TESI)/Poisson3DSolver.cpp:14-34) on a NON-CUBIC grid, with PrintDiff() switched on.  Built twice by
// tests/test_compat_equivalence.py: against the reference's own Grid3D/MultiGrid3D sources with -DNDEBUG (their asserts at
// N3/Grid3D.cpp:10-11 are all that forbids such a grid) and against the shim of include/compat/ -- log/diff.txt must come out
// byte for byte the same.  Also exercises the free-array operators the way VCycle uses them.
#include <stdio.h>
#include <stdlib.h>
#include "MultiGrid3D.h"

int main(int argc, char** argv)
{
    int nx = argc > 1 ? atoi(argv[1]) : 33, ny = argc > 2 ? atoi(argv[2]) : 17, nz = argc > 3 ? atoi(argv[3]) : 9;
    int v0 = argc > 4 ? atoi(argv[4]) : 2, nu = argc > 5 ? atoi(argv[5]) : 3;
    int finestGridSize[3] = {nx, ny, nz};
    float range[6] = {0, 1, 0, 1, 0, 1};
    MultiGrid3D multiGrid3D(finestGridSize, range);
    multiGrid3D.FullMultiGridVCycle(0, v0, nu, nu);
    // one more cycle by hand through the public operators, in VCycle's own order on the two finest levels
    Grid3D* fine = multiGrid3D.grids3D[0];
    Grid3D* coarse = multiGrid3D.grids3D[1];
    multiGrid3D.Relax(fine, nu);
    float* res = multiGrid3D.CalculateResidual(fine);
    multiGrid3D.Restrict(res, fine->sizeXYZ, coarse->h_f, coarse->sizeXYZ);
    multiGrid3D.setToValue(coarse->h_v, coarse->sizeXYZ, 0, true);
    multiGrid3D.Relax(coarse, 2 * nu);
    multiGrid3D.Interpolate(res, fine->sizeXYZ, coarse->h_v, coarse->sizeXYZ);
    multiGrid3D.ApplyCorrection(fine->h_v, fine->sizeXYZ, res, fine->sizeXYZ);
    multiGrid3D.Relax(fine, nu);
    multiGrid3D.PrintDiff();
    printf("finestGridSize: %d %d %d\n", nx, ny, nz);
    return 0;
}
