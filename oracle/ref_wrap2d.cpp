/*
 * oracle/ref_wrap2d.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the reference's CPU 2D Lyapunov multigrid solver
 * (/root/reference/NOCUDA_TESI/PDE Lyapunov 2D/{Grid2D,MultiGrid2D}.cpp) unmodified behind a
 * flat C interface.  Built as ref2d_f32 (as written) and ref2d_f64 (-DREF_F64).  The 2D
 * residual is consistent with the smoother (N2/MultiGrid2D.cpp:241 vs :403), so there is no
 * CORRECTED variant.
 */
#include "ref_wrap_common.h"

namespace REF_PREFIX {
#include "Grid2D.cpp"
#include "MultiGrid2D.cpp"
}

using REF_PREFIX::MultiGrid2D;
using REF_PREFIX::Grid2D;

extern "C" {

/* A4 = {A[0][0], A[0][1], A[1][0], A[1][1]}; A_size is passed as 2 exactly like
   N2/LyapunovSolver.cpp:15,37 */
void* REF_FN(create)(int n, const double* range4, const double* A4, int alfa)
{
    int sz[2] = {n, n};
    ref_real r[4], A[4];
    for (int i = 0; i < 4; i++) { r[i] = (ref_real)range4[i]; A[i] = (ref_real)A4[i]; }
    return new MultiGrid2D(sz, r, A, 2, alfa);
}

void REF_FN(destroy)(void* h)
{
    MultiGrid2D* mg = (MultiGrid2D*)h;
    for (int l = 0; l < mg->numGrids; l++) {
        free(mg->grids2D[l]->h_v);
        free(mg->grids2D[l]->h_f);
        free(mg->grids2D[l]->sizeXY);
    }
    free(mg->grids2D);
    free(mg->matrixA);
    ::operator delete((void*)mg);
}

int REF_FN(num_levels)(void* h) { return ((MultiGrid2D*)h)->numGrids; }
int REF_FN(level_size)(void* h, int l) { return ((MultiGrid2D*)h)->grids2D[l]->sizeX; }
ref_real* REF_FN(level_v)(void* h, int l) { return ((MultiGrid2D*)h)->grids2D[l]->h_v; }
ref_real* REF_FN(level_f)(void* h, int l) { return ((MultiGrid2D*)h)->grids2D[l]->h_f; }
double REF_FN(level_h)(void* h, int l) { return (double)((MultiGrid2D*)h)->grids2D[l]->h_x; }

void REF_FN(relax)(void* h, int l, int ncycles)
{
    MultiGrid2D* mg = (MultiGrid2D*)h;
    mg->Relax(mg->grids2D[l], ncycles);
}

void REF_FN(residual)(void* h, int l, ref_real* out)
{
    MultiGrid2D* mg = (MultiGrid2D*)h;
    Grid2D* g = mg->grids2D[l];
    ref_real* r = mg->CalculateResidual(g);
    memcpy(out, r, sizeof(ref_real) * (size_t)g->sizeX * g->sizeY);
    free(r);
}

void REF_FN(restrict_)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    int fs[2] = {fn, fn}, cs[2] = {cn, cn};
    ((MultiGrid2D*)h)->Restrict(fine, fs, coarse, cs);
}

void REF_FN(interpolate)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    int fs[2] = {fn, fn}, cs[2] = {cn, cn};
    ((MultiGrid2D*)h)->Interpolate(fine, fs, coarse, cs);
}

void REF_FN(apply_correction)(void* h, ref_real* fine, int fn, ref_real* err, int en)
{
    int fs[2] = {fn, fn}, es[2] = {en, en};
    ((MultiGrid2D*)h)->ApplyCorrection(fine, fs, err, es);
}

void REF_FN(set_to_value)(void* h, ref_real* grid, int n, double value, int modify_boundaries)
{
    int s[2] = {n, n};
    ((MultiGrid2D*)h)->setToValue(grid, s, (ref_real)value, modify_boundaries != 0);
}

void REF_FN(vcycle)(void* h, int l, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid2D*)h)->VCycle(l, v1, v2);
    refwrap::track_end_free();
}

void REF_FN(fmg)(void* h, int l, int v0, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid2D*)h)->FullMultiGridVCycle(l, v0, v1, v2);
    refwrap::track_end_free();
}

} // extern "C"
