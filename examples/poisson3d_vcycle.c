/* examples/poisson3d_vcycle.c -- the C ABI from plain C: what N3/Poisson3DSolver.cpp does through the reference's classes
 * (construct the hierarchy on [0,1]^3, run cycles), plus the residual history the reference cannot print.
 *
 *   gcc -O2 -I include examples/poisson3d_vcycle.c -o poisson3d -L pde_multigrid_b200 -lmg_b200 \
 *       -Wl,-rpath,'$ORIGIN/pde_multigrid_b200' -lm
 *   ./poisson3d [n] [cycles]        (n = 2^k + 1, default 129)
 */
#include <stdio.h>
#include <stdlib.h>

#include "mg_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        if ((call) != MG_OK) {                                                   \
            fprintf(stderr, "%s failed: %s\n", #call, mg_last_error());          \
            return 1;                                                            \
        }                                                                        \
    } while (0)

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 129, cycles = argc > 2 ? atoi(argv[2]) : 5;
    int size[3] = {n, n, n};
    double range[6] = {0, 1, 0, 1, 0, 1};
    mg3d_t* mg = NULL;
    double l2, linf;

    CHECK(mg3d_create(&mg, size, range, MG_F64, MG_CORRECTED)); /* MultiGrid3D(finestGridSizeXYZ, range): InitV/InitF on every level */
    CHECK(mg3d_residual_norm(mg, 0, &l2, &linf));
    printf("levels %d, ||r0||_2 = %.9e\n", mg3d_num_levels(mg), l2);
    for (int c = 1; c <= cycles; c++) {
        CHECK(mg3d_vcycle(mg, 0, 2, 2)); /* VCycle(0, 2, 2) */
        CHECK(mg3d_residual_norm(mg, 0, &l2, &linf));
        printf("after V(2,2) cycle %d: ||r||_2 = %.9e  ||r||_inf = %.3e\n", c, l2, linf);
    }
    /* the solution of the finest level in the reference's dense layout (x fastest), as grids3D[0]->h_v */
    double* v = (double*)malloc((size_t)n * n * n * sizeof(double));
    if (!v) return 1;
    CHECK(mg3d_get_field(mg, 0, MG_FIELD_V, v));
    printf("v(centre) = %.6f (exact solution sin(pi x) sin(pi y) sin(pi z) = 1)\n", v[((size_t)(n / 2) * n + n / 2) * n + n / 2]);
    free(v);
    CHECK(mg3d_destroy(mg));
    return 0;
}
