// tests/compat/drv3d.cpp -- test driver (not product code): the calls of the reference's 3D main
// (NOCUDA_TESI/POISSON_3D(TESI)/Poisson3DSolver.cpp:14-34) with PrintDiff() switched on, at a size given on the command
// line.  Built twice by tests/test_compat_equivalence.py: against the reference's own Grid3D/MultiGrid3D sources and against
// the shim of include/compat/ -- log/diff.txt must come out byte for byte the same.
#include <stdio.h>
#include <stdlib.h>
#include "MultiGrid3D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 17, v0 = argc > 2 ? atoi(argv[2]) : 2, nu = argc > 3 ? atoi(argv[3]) : 3;
    int finestGridSize[3] = {n, n, n};
    float range[6] = {0, 1, 0, 1, 0, 1};
    MultiGrid3D multiGrid3D(finestGridSize, range);
    multiGrid3D.FullMultiGridVCycle(0, v0, nu, nu);
    multiGrid3D.PrintDiff();
    printf("finestGridSize: %d\n", n);
    return 0;
}
