/*
 * oracle/ref_wrap3d.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the reference's CPU 3D Poisson multigrid solver
 * (/root/reference/NOCUDA_TESI/POISSON_3D(TESI)/{Grid3D,MultiGrid3D}.cpp) unmodified
 * and exposes its methods through a flat C interface so that tests (ctypes) can call
 * each operator on caller-provided arrays.  Built four times by oracle/build_ref.py:
 *   REF_PREFIX=ref3d_f32   as written (float)
 *   REF_PREFIX=ref3d_f64   -DREF_F64 (`#define float double`)
 *   REF_PREFIX=ref3d_f32c / ref3d_f64c   same, but MultiGrid3D.cpp is taken from a
 *       build-time patched temp copy with the two residual signs of line 723 flipped
 *       (the CORRECTED mode of SURVEY.md section 0.5; the patch lives in build_ref.py).
 *   REF_PREFIX=ref3d_f32x / ref3d_f64x / ref3d_f32cx / ref3d_f64cx   the same four with -DNDEBUG: the only thing that
 *       keeps the reference from running a NON-CUBIC grid is the pair of asserts at N3/Grid3D.cpp:10-11 (its hierarchy,
 *       N3/MultiGrid3D.cpp:19-47, and every operator already work per dimension; the author's TODO, N3/TODO!!!), so with
 *       assertions compiled out -- no source change -- the reference itself is the oracle for sizeX != sizeY != sizeZ
 *       (create_xyz below).
 * Each build sits in its own namespace so the copies of the classes can share one .so.
 */
#include "ref_wrap_common.h"

namespace REF_PREFIX {
#include "Grid3D.cpp"
#include "MultiGrid3D.cpp"
}

using REF_PREFIX::MultiGrid3D;
using REF_PREFIX::Grid3D;

extern "C" {

void* REF_FN(create)(int n, const double* range6)
{
    int sz[3] = {n, n, n};
    ref_real r[6];
    for (int i = 0; i < 6; i++) r[i] = (ref_real)range6[i];
    MultiGrid3D* mg = new MultiGrid3D(sz, r);
    /* reference leaves the interior of v uninitialised (N3/Grid3D.cpp:61-76): zero it so
       that a bare VCycle starts from v = 0 */
    for (int l = 0; l < mg->numGrids; l++)
        mg->setToValue(mg->grids3D[l]->h_v, mg->grids3D[l]->sizeXYZ, 0.0f, false);
    return mg;
}

/* non-cubic finest grid (the *x variants only: the others abort on the reference's own assert) */
void* REF_FN(create_xyz)(int nx, int ny, int nz, const double* range6)
{
    int sz[3] = {nx, ny, nz};
    ref_real r[6];
    for (int i = 0; i < 6; i++) r[i] = (ref_real)range6[i];
    MultiGrid3D* mg = new MultiGrid3D(sz, r);
    for (int l = 0; l < mg->numGrids; l++)
        mg->setToValue(mg->grids3D[l]->h_v, mg->grids3D[l]->sizeXYZ, 0.0f, false);
    return mg;
}

void REF_FN(level_size_xyz)(void* h, int l, int* out3)
{
    Grid3D* g = ((MultiGrid3D*)h)->grids3D[l];
    out3[0] = g->sizeX; out3[1] = g->sizeY; out3[2] = g->sizeZ;
}

void REF_FN(destroy)(void* h)
{
    MultiGrid3D* mg = (MultiGrid3D*)h;
    for (int l = 0; l < mg->numGrids; l++) {
        free(mg->grids3D[l]->h_v);
        free(mg->grids3D[l]->h_f);
        free(mg->grids3D[l]->sizeXYZ);
    }
    free(mg->grids3D);
    /* objects themselves were new'ed by the reference and never deleted; operator delete
       without running the (double-freeing) destructors */
    ::operator delete((void*)mg);
}

int REF_FN(num_levels)(void* h) { return ((MultiGrid3D*)h)->numGrids; }
int REF_FN(level_size)(void* h, int l) { return ((MultiGrid3D*)h)->grids3D[l]->sizeX; }
ref_real* REF_FN(level_v)(void* h, int l) { return ((MultiGrid3D*)h)->grids3D[l]->h_v; }
ref_real* REF_FN(level_f)(void* h, int l) { return ((MultiGrid3D*)h)->grids3D[l]->h_f; }
double REF_FN(level_h)(void* h, int l) { return (double)((MultiGrid3D*)h)->grids3D[l]->h_x; }

void REF_FN(relax)(void* h, int l, int ncycles)
{
    MultiGrid3D* mg = (MultiGrid3D*)h;
    mg->Relax(mg->grids3D[l], ncycles);
}

/* out must hold n^3 values */
void REF_FN(residual)(void* h, int l, ref_real* out)
{
    MultiGrid3D* mg = (MultiGrid3D*)h;
    Grid3D* g = mg->grids3D[l];
    ref_real* r = mg->CalculateResidual(g);
    memcpy(out, r, sizeof(ref_real) * (size_t)g->sizeX * g->sizeY * g->sizeZ);
    free(r);
}

void REF_FN(restrict_)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    int fs[3] = {fn, fn, fn}, cs[3] = {cn, cn, cn};
    ((MultiGrid3D*)h)->Restrict(fine, fs, coarse, cs);
}

void REF_FN(interpolate)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    int fs[3] = {fn, fn, fn}, cs[3] = {cn, cn, cn};
    ((MultiGrid3D*)h)->Interpolate(fine, fs, coarse, cs);
}

void REF_FN(apply_correction)(void* h, ref_real* fine, int fn, ref_real* err, int en)
{
    int fs[3] = {fn, fn, fn}, es[3] = {en, en, en};
    ((MultiGrid3D*)h)->ApplyCorrection(fine, fs, err, es);
}

void REF_FN(set_to_value)(void* h, ref_real* grid, int n, double value, int modify_boundaries)
{
    int s[3] = {n, n, n};
    ((MultiGrid3D*)h)->setToValue(grid, s, (ref_real)value, modify_boundaries != 0);
}

void REF_FN(vcycle)(void* h, int l, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid3D*)h)->VCycle(l, v1, v2);
    refwrap::track_end_free();
}

void REF_FN(fmg)(void* h, int l, int v0, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid3D*)h)->FullMultiGridVCycle(l, v0, v1, v2);
    refwrap::track_end_free();
}

} // extern "C"
