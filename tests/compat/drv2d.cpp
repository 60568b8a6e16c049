// tests/compat/drv2d.cpp -- test driver (not product code): the calls of NOCUDA_TESI/PDE Lyapunov 2D/LyapunovSolver.cpp:13-44
// (A = [-1 -2; 0 -3], alfa = 2, FullMultiGridVCycle, PrintDiff) at a size given on the command line.  See drv3d.cpp.
#include <stdio.h>
#include <stdlib.h>
#include "MultiGrid2D.h"

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 33, v0 = argc > 2 ? atoi(argv[2]) : 1, nu = argc > 3 ? atoi(argv[3]) : 20;
    int finestGridSize[2] = {n, n};
    float range[4] = {0, 1, 0, 1};
    float* A = (float*)malloc(16 * sizeof(float));  // the reference's InitA writes 4 floats into a 2-float buffer of its own: ours is roomy
    A[0] = -1; A[1] = -2; A[2] = 0; A[3] = -3;
    MultiGrid2D multiGrid2D(finestGridSize, range, A, 2, 2);
    multiGrid2D.FullMultiGridVCycle(0, v0, nu, nu);
    multiGrid2D.PrintDiff();
    printf("finestGridSize: %d\n", n);
    return 0;
}
