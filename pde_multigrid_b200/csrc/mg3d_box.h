/* mg3d_box.h -- internal interface between mg3d_box_host.c and mg3d_box.cu (non-cubic 3D grids, mg3b_* of the ABI). */
#ifndef MG3D_BOX_H
#define MG3D_BOX_H

#include "mg_launch.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A level field is colour-split exactly like the cubic engine's (mg_launch.h): two arrays, one per red-black colour
   c = (x+y+z)&1, each compacted along x: element (x,y,z) at base[c*cstride + z*plane + y*hp + (x>>1)], hp = (nx+1)/2 rounded up
   to 128 bytes, plane = hp*ny, cstride = plane*nz.  A half-sweep then reads the other colour's array and its own colour of f and
   writes its own colour of v, all with unit stride. */
typedef struct {
    int nx, ny, nz;
    int hp;
    long long plane, cstride;
} mg_geom3b;

/* level operators on colour-split fields; every launcher returns the number of kernels launched or -1 */
int mgk3b_relax_colour(cudaStream_t s, int dtype, void* v, const void* f, mg_geom3b g, mg_coef3d c, int colour);
int mgk3b_residual_dense(cudaStream_t s, int dtype, const void* v, const void* f, void* r_dense, mg_geom3b g, mg_coef3d c, int corrected);
/* fv == NULL: Restrict(ff) -> cf; otherwise Restrict(CalculateResidual(fv, ff)) -> cf and cv = 0 */
int mgk3b_restrict(cudaStream_t s, int dtype, const void* fv, const void* ff, mg_geom3b g, mg_coef3d c, int corrected, void* cf, void* cv, mg_geom3b gc);
int mgk3b_interpolate(cudaStream_t s, int dtype, void* fv, mg_geom3b g, const void* cv, mg_geom3b gc, int add);
int mgk3b_set(cudaStream_t s, int dtype, void* a, mg_geom3b g, double value, int modify_boundaries);
int mgk3b_init_f(cudaStream_t s, int dtype, void* f, mg_geom3b g, const double* sx, const double* sy, const double* sz);
int mgk3b_repack(cudaStream_t s, int dtype, void* split, mg_geom3b g, void* dense, int to_split);
int mgk3b_residual_norm(cudaStream_t s, int dtype, const void* v, const void* f, mg_geom3b g, mg_coef3d c, int corrected, double* parts, int nparts,
                        double* out2);
/* the reference's operators on free DENSE arrays (x fastest), for the host-array entry points */
int mgk3b_dense_restrict(cudaStream_t s, int dtype, const void* fine, const int n[3], void* coarse, const int cn[3]);
int mgk3b_dense_interpolate(cudaStream_t s, int dtype, void* fine, const int n[3], const void* coarse, const int cn[3]);
int mgk3b_dense_apply_correction(cudaStream_t s, int dtype, void* fine, const void* err, const int n[3]);
int mgk3b_dense_set(cudaStream_t s, int dtype, void* a, const int n[3], double value, int modify_boundaries);

#ifdef __cplusplus
}
#endif
#endif
