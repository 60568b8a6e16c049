// tests/compat/drv1d.cpp -- test driver (not product code): the calls of NOCUDA_TESI/EQUAZIONE 1D/Poisson1DSolver.cpp:13-25
// with PrintDiff() switched on, at a size given on the command line.  See drv3d.cpp.
#include <stdio.h>
#include <stdlib.h>
#include "MultiGrid1D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 129, v0 = argc > 2 ? atoi(argv[2]) : 2, nu = argc > 3 ? atoi(argv[3]) : 100;
    float range[2] = {0, 1};
    MultiGrid1D multiGrid1D(n, range);
    multiGrid1D.FullMultiGridVCycle(0, v0, nu, nu);
    multiGrid1D.PrintDiff();
    printf("finestGridSize: %d\n", n);
    return 0;
}
