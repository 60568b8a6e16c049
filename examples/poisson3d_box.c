/* examples/poisson3d_box.c -- a NON-CUBIC grid through the C ABI (mg3b_*): the reference's constructor takes three sizes
 * (N3/MultiGrid3D.cpp:5-8) but Grid3D asserts them equal (N3/Grid3D.cpp:10-11); here FullMultiGridVCycle runs on e.g. 129 x 65 x 33.
 *
 *   gcc -O2 -I include examples/poisson3d_box.c -o poisson3d_box -L pde_multigrid_b200 -lmg_b200 \
 *       -Wl,-rpath,'$ORIGIN/pde_multigrid_b200' -lm
 *   ./poisson3d_box [nx ny nz] [v0]      (every size 2^k + 1; default 129 65 33, v0 = 4)
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
    int size[3] = {argc > 3 ? atoi(argv[1]) : 129, argc > 3 ? atoi(argv[2]) : 65, argc > 3 ? atoi(argv[3]) : 33};
    const int v0 = argc > 4 ? atoi(argv[4]) : 4;
    double range[6] = {0, 1, 0, 1, 0, 1};
    mg3b_t* mg = NULL;
    double l2, linf;

    CHECK(mg3b_create(&mg, size, range, MG_F64, MG_CORRECTED));
    const int levels = mg3b_num_levels(mg); /* (int)log2(min size - 1), N3/MultiGrid3D.cpp:24-34 */
    int coarsest[3];
    CHECK(mg3b_level_size(mg, levels - 1, coarsest));
    CHECK(mg3b_residual_norm(mg, 0, &l2, &linf));
    printf("%d x %d x %d: %d levels down to %d x %d x %d, ||r0||_2 = %.9e\n", size[0], size[1], size[2], levels, coarsest[0], coarsest[1],
           coarsest[2], l2);
    CHECK(mg3b_fmg(mg, 0, v0, 2, 2)); /* FullMultiGridVCycle(0, v0, 2, 2) */
    CHECK(mg3b_residual_norm(mg, 0, &l2, &linf));
    printf("after FMG(%d,2,2): ||r||_2 = %.9e  ||r||_inf = %.3e\n", v0, l2, linf);
    double* v = (double*)malloc((size_t)size[0] * size[1] * size[2] * sizeof(double));
    if (!v) return 1;
    CHECK(mg3b_get_field(mg, 0, MG_FIELD_V, v)); /* dense, x fastest: idx = x + y*sizeX + z*sizeX*sizeY */
    printf("v(centre) = %.6f (exact solution sin(pi x) sin(pi y) sin(pi z) = 1)\n",
           v[(size_t)(size[0] / 2) + (size_t)size[0] * ((size_t)(size[1] / 2) + (size_t)size[1] * (size_t)(size[2] / 2))]);
    free(v);
    CHECK(mg3b_destroy(mg));
    return 0;
}
