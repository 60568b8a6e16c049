/*
 * oracle/ref_wrap1d.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the reference's CPU 1D first-order ODE multigrid solver
 * (/root/reference/NOCUDA_TESI/EQUAZIONE 1D/{Grid1D,MultiGrid1D}.cpp) unmodified behind a flat
 * C interface.  Built as ref1d_f32 / ref1d_f64 (-DREF_F64) and ref1d_f32c / ref1d_f64c (the
 * residual sign of N1/MultiGrid1D.cpp:210 flipped in a build-time temp copy, see build_ref.py).
 */
#include "ref_wrap_common.h"

namespace REF_PREFIX {
#include "Grid1D.cpp"
#include "MultiGrid1D.cpp"
}

using REF_PREFIX::MultiGrid1D;
using REF_PREFIX::Grid1D;

extern "C" {

void* REF_FN(create)(int n, const double* range2)
{
    ref_real r[2] = {(ref_real)range2[0], (ref_real)range2[1]};
    MultiGrid1D* mg = new MultiGrid1D(n, r);
    /* interior v is uninitialised in the reference (N1/Grid1D.cpp:30-34): zero it */
    for (int l = 0; l < mg->numGrids; l++)
        mg->setToValue(mg->grids1D[l]->h_v, mg->grids1D[l]->sizeX, 0.0f, false);
    return mg;
}

void REF_FN(destroy)(void* h)
{
    MultiGrid1D* mg = (MultiGrid1D*)h;
    for (int l = 0; l < mg->numGrids; l++) {
        free(mg->grids1D[l]->h_v);
        free(mg->grids1D[l]->h_f);
    }
    free(mg->grids1D);
    ::operator delete((void*)mg);
}

int REF_FN(num_levels)(void* h) { return ((MultiGrid1D*)h)->numGrids; }
int REF_FN(level_size)(void* h, int l) { return ((MultiGrid1D*)h)->grids1D[l]->sizeX; }
ref_real* REF_FN(level_v)(void* h, int l) { return ((MultiGrid1D*)h)->grids1D[l]->h_v; }
ref_real* REF_FN(level_f)(void* h, int l) { return ((MultiGrid1D*)h)->grids1D[l]->h_f; }
double REF_FN(level_h)(void* h, int l) { return (double)((MultiGrid1D*)h)->grids1D[l]->h_x; }

void REF_FN(relax)(void* h, int l, int ncycles)
{
    MultiGrid1D* mg = (MultiGrid1D*)h;
    mg->Relax(mg->grids1D[l], ncycles);
}

void REF_FN(residual)(void* h, int l, ref_real* out)
{
    MultiGrid1D* mg = (MultiGrid1D*)h;
    Grid1D* g = mg->grids1D[l];
    ref_real* r = mg->CalculateResidual(g);
    memcpy(out, r, sizeof(ref_real) * (size_t)g->sizeX);
    free(r);
}

void REF_FN(restrict_)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    ((MultiGrid1D*)h)->Restrict(fine, fn, coarse, cn);
}

void REF_FN(interpolate)(void* h, ref_real* fine, int fn, ref_real* coarse, int cn)
{
    ((MultiGrid1D*)h)->Interpolate(fine, fn, coarse, cn);
}

void REF_FN(apply_correction)(void* h, ref_real* fine, int fn, ref_real* err, int en)
{
    ((MultiGrid1D*)h)->ApplyCorrection(fine, fn, err, en);
}

void REF_FN(set_to_value)(void* h, ref_real* grid, int n, double value, int modify_boundaries)
{
    ((MultiGrid1D*)h)->setToValue(grid, n, (ref_real)value, modify_boundaries != 0);
}

void REF_FN(vcycle)(void* h, int l, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid1D*)h)->VCycle(l, v1, v2);
    refwrap::track_end_free();
}

void REF_FN(fmg)(void* h, int l, int v0, int v1, int v2)
{
    refwrap::track_begin();
    ((MultiGrid1D*)h)->FullMultiGridVCycle(l, v0, v1, v2);
    refwrap::track_end_free();
}

} // extern "C"
