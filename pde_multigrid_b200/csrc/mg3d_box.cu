// mg3d_box.cu -- the 3D Poisson operators on NON-CUBIC grids (sizeX != sizeY != sizeZ, each 2^k + 1).
//
// The reference asserts such grids away (N3/Grid3D.cpp:10-11; the author's own TODO) although its
// hierarchy (N3/MultiGrid3D.cpp:19-47) and operators are written per dimension.  This file is the device side of that
// generalisation: the same per-point formulas as the cubic engine (mg3d_device.cuh: relax_point, residual_point,
// restrict_point, interp_point -- expression order of the reference, no FMA contraction, IEEE division because the three
// mesh widths differ), on the reference's own dense layout (x fastest, idx = x + y*sx + z*sx*sy), so fields move between
// host and device without repacking.  One thread per point; the fine residual is evaluated on the fly inside the
// restriction and never written to memory, prolongation and correction are one kernel.
//
// Not tuned like the cubic path (no colour-split layout, no TMA staging, no temporal blocking): with both colours
// interleaved in every 32-byte sector a half-sweep touches all of v, which caps the smoother near half of the HBM
// roofline (DESIGN.md section 3 has the measurement that motivated the colour split).
#include "mg3d_device.cuh"

using namespace mgx;
using namespace mg3;

namespace {

struct BoxGeom {
    int nx, ny, nz;
};

__device__ __forceinline__ long long bidx(const BoxGeom& g, int x, int y, int z) { return (long long)x + (long long)g.nx * ((long long)y + (long long)g.ny * z); }

// MultiGrid3D::Relax, one colour (N3/MultiGrid3D.cpp:515 / :544): thread = one point of the colour
template <typename T>
__global__ void __launch_bounds__(128) k_box_relax(T* __restrict__ v, const T* __restrict__ f, BoxGeom g, Coef3<T> c, int colour)
{
    const int y = 1 + blockIdx.y, z = 1 + blockIdx.z;
    const int x = 1 + ((1 + y + z + colour) & 1) + 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (x > g.nx - 2) return;
    const long long i = bidx(g, x, y, z), sy = g.nx, sz = (long long)g.nx * g.ny;
    v[i] = relax_point<T, false>(v[i - 1], v[i + 1], v[i - sy], v[i + sy], v[i - sz], v[i + sz], f[i], c);
}

template <typename T>
__device__ __forceinline__ T box_residual_at(const T* __restrict__ v, const T* __restrict__ f, const BoxGeom& g, int x, int y, int z,
                                             const Coef3<T>& c, int corrected)
{
    if (x == 0 || x == g.nx - 1 || y == 0 || y == g.ny - 1 || z == 0 || z == g.nz - 1) return T(0);  // :704-705
    const long long i = bidx(g, x, y, z), sy = g.nx, sz = (long long)g.nx * g.ny;
    return residual_point<T, false>(v[i - 1], v[i + 1], v[i - sy], v[i + sy], v[i - sz], v[i + sz], v[i], f[i], c, corrected);
}

// MultiGrid3D::CalculateResidual (N3/MultiGrid3D.cpp:678-730) into a caller-visible array
template <typename T>
__global__ void __launch_bounds__(128) k_box_residual(const T* __restrict__ v, const T* __restrict__ f, T* __restrict__ r, BoxGeom g,
                                                      Coef3<T> c, int corrected)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= g.nx) return;
    r[bidx(g, x, y, z)] = box_residual_at(v, f, g, x, y, z, c, corrected);
}

// Restrict (N3/MultiGrid3D.cpp:50-184).  RESIDUAL: the fine operand is CalculateResidual(fine v, fine f), evaluated on
// the fly, and the coarse v is zeroed (setToValue(coarse v, 0, true), :634) -- the three steps of VCycle in one kernel.
template <typename T, bool RESIDUAL>
__global__ void __launch_bounds__(128) k_box_restrict(const T* __restrict__ fv, const T* __restrict__ ff, BoxGeom g, Coef3<T> c, int corrected,
                                                      T* __restrict__ cf, T* __restrict__ cv, BoxGeom gc)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, cy = blockIdx.y, cz = blockIdx.z;
    if (cx >= gc.nx) return;
    const int fx = 2 * cx, fy = 2 * cy, fz = 2 * cz;
    auto R = [&](int dx, int dy, int dz) -> T {
        if (RESIDUAL) return box_residual_at(fv, ff, g, fx + dx, fy + dy, fz + dz, c, corrected);
        return ff[bidx(g, fx + dx, fy + dy, fz + dz)];
    };
    const bool bnd = cx == 0 || cx == gc.nx - 1 || cy == 0 || cy == gc.ny - 1 || cz == 0 || cz == gc.nz - 1;
    const long long ci = bidx(gc, cx, cy, cz);
    cf[ci] = bnd ? R(0, 0, 0) : restrict_point<T>(R);  // :113-119 injection on the boundary
    if (RESIDUAL) cv[ci] = T(0);
}

// Interpolate (N3/MultiGrid3D.cpp:186-335), interior of the fine grid; ADD: + ApplyCorrection (:649-676)
template <typename T, bool ADD>
__global__ void __launch_bounds__(128) k_box_interpolate(T* __restrict__ fv, BoxGeom g, const T* __restrict__ cv, BoxGeom gc)
{
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, z = 1 + blockIdx.z;
    if (x > g.nx - 2) return;
    const int cx = x >> 1, cy = y >> 1, cz = z >> 1;
    auto C = [&](int dx, int dy, int dz) -> T { return cv[bidx(gc, cx + dx, cy + dy, cz + dz)]; };
    const T e = interp_point<T>(C, x & 1, y & 1, z & 1);
    const long long i = bidx(g, x, y, z);
    fv[i] = ADD ? add(fv[i], e) : e;
}

template <typename T>
__global__ void __launch_bounds__(128) k_box_apply_correction(T* __restrict__ fv, const T* __restrict__ err, BoxGeom g)
{
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x, y = 1 + blockIdx.y, z = 1 + blockIdx.z;
    if (x > g.nx - 2) return;
    const long long i = bidx(g, x, y, z);
    fv[i] = add(fv[i], err[i]);
}

// setToValue (N3/MultiGrid3D.cpp:587-621)
template <typename T>
__global__ void __launch_bounds__(128) k_box_set(T* __restrict__ a, BoxGeom g, T value, int modify_boundaries)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= g.nx) return;
    const bool bnd = x == 0 || x == g.nx - 1 || y == 0 || y == g.ny - 1 || z == 0 || z == g.nz - 1;
    if (modify_boundaries || !bnd) a[bidx(g, x, y, z)] = value;
}

// Grid3D::InitF (N3/Grid3D.cpp:78-96): (real)(-3*PI*PI*sin(PI x)*sin(PI y)*sin(PI z)), the product in double, left to
// right; the sines are host-libm tables (the reference's own values)
template <typename T>
__global__ void __launch_bounds__(128) k_box_init_f(T* __restrict__ f, BoxGeom g, const double* __restrict__ sx, const double* __restrict__ sy,
                                                    const double* __restrict__ sz)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= g.nx) return;
    const double PI = 3.141592653589793;
    const double k = __dmul_rn(__dmul_rn(-3.0, PI), PI);
    f[bidx(g, x, y, z)] = (T)__dmul_rn(__dmul_rn(__dmul_rn(k, sx[x]), sy[y]), sz[z]);
}

// ||r||_2^2 and ||r||_inf of CalculateResidual, two deterministic stages: one partial pair per block, then one block
template <typename T>
__global__ void __launch_bounds__(256) k_box_norm(const T* __restrict__ v, const T* __restrict__ f, BoxGeom g, Coef3<T> c, int corrected,
                                                  double* __restrict__ parts, int nparts)
{
    __shared__ double sh[64];
    const long long tot = (long long)g.nx * g.ny * g.nz, per = (tot + nparts - 1) / nparts;
    const long long lo = (long long)blockIdx.x * per, hi = lo + per < tot ? lo + per : tot;
    double s = 0.0, m = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int x = (int)(i % g.nx), y = (int)((i / g.nx) % g.ny), z = (int)(i / ((long long)g.nx * g.ny));
        const double r = (double)box_residual_at(v, f, g, x, y, z, c, corrected);
        s += r * r;
        m = fmax(m, fabs(r));
    }
    block_sum_max(s, m, sh);
    if (threadIdx.x == 0) { parts[blockIdx.x] = s; parts[nparts + blockIdx.x] = m; }
}

__global__ void k_box_norm_final(const double* __restrict__ parts, int nparts, double* __restrict__ out2)
{
    if (threadIdx.x || blockIdx.x) return;
    double s = 0.0, m = 0.0;
    for (int i = 0; i < nparts; i++) { s += parts[i]; m = fmax(m, parts[nparts + i]); }
    out2[0] = s;
    out2[1] = m;
}

inline dim3 grid_for(int w, int ny, int nz) { return dim3((unsigned)((w + 127) / 128), (unsigned)ny, (unsigned)nz); }
inline int ok() { return cudaPeekAtLastError() == cudaSuccess ? 1 : -1; }
inline BoxGeom bg(const int n[3]) { return BoxGeom{n[0], n[1], n[2]}; }

}  // namespace

#define BOX_DISPATCH(expr_f, expr_d) \
    do {                             \
        if (dtype == 0) { expr_f; }  \
        else { expr_d; }             \
    } while (0)

extern "C" {

int mgk3b_relax_colour(cudaStream_t s, int dtype, void* v, const void* f, const int n[3], mg_coef3d c, int colour)
{
    if (n[0] < 3 || n[1] < 3 || n[2] < 3) return 0;
    const dim3 grid = grid_for((n[0] - 2 + 1) / 2, n[1] - 2, n[2] - 2);
    BOX_DISPATCH((k_box_relax<float><<<grid, 128, 0, s>>>((float*)v, (const float*)f, bg(n), narrow<float>(c), colour)),
                 (k_box_relax<double><<<grid, 128, 0, s>>>((double*)v, (const double*)f, bg(n), narrow<double>(c), colour)));
    return ok();
}

int mgk3b_residual(cudaStream_t s, int dtype, const void* v, const void* f, void* r, const int n[3], mg_coef3d c, int corrected)
{
    const dim3 grid = grid_for(n[0], n[1], n[2]);
    BOX_DISPATCH((k_box_residual<float><<<grid, 128, 0, s>>>((const float*)v, (const float*)f, (float*)r, bg(n), narrow<float>(c), corrected)),
                 (k_box_residual<double><<<grid, 128, 0, s>>>((const double*)v, (const double*)f, (double*)r, bg(n), narrow<double>(c), corrected)));
    return ok();
}

/* fv == NULL: plain Restrict of the field ff; otherwise Restrict(CalculateResidual(fv, ff)) -> cf, and cv = 0 */
int mgk3b_restrict(cudaStream_t s, int dtype, const void* fv, const void* ff, const int n[3], mg_coef3d c, int corrected, void* cf, void* cv,
                   const int cn[3])
{
    const dim3 grid = grid_for(cn[0], cn[1], cn[2]);
    if (fv)
        BOX_DISPATCH((k_box_restrict<float, true><<<grid, 128, 0, s>>>((const float*)fv, (const float*)ff, bg(n), narrow<float>(c), corrected, (float*)cf, (float*)cv, bg(cn))),
                     (k_box_restrict<double, true><<<grid, 128, 0, s>>>((const double*)fv, (const double*)ff, bg(n), narrow<double>(c), corrected, (double*)cf, (double*)cv, bg(cn))));
    else
        BOX_DISPATCH((k_box_restrict<float, false><<<grid, 128, 0, s>>>(nullptr, (const float*)ff, bg(n), narrow<float>(c), corrected, (float*)cf, nullptr, bg(cn))),
                     (k_box_restrict<double, false><<<grid, 128, 0, s>>>(nullptr, (const double*)ff, bg(n), narrow<double>(c), corrected, (double*)cf, nullptr, bg(cn))));
    return ok();
}

int mgk3b_interpolate(cudaStream_t s, int dtype, void* fv, const int n[3], const void* cv, const int cn[3], int add)
{
    if (n[0] < 3 || n[1] < 3 || n[2] < 3) return 0;
    const dim3 grid = grid_for(n[0] - 2, n[1] - 2, n[2] - 2);
    if (add)
        BOX_DISPATCH((k_box_interpolate<float, true><<<grid, 128, 0, s>>>((float*)fv, bg(n), (const float*)cv, bg(cn))),
                     (k_box_interpolate<double, true><<<grid, 128, 0, s>>>((double*)fv, bg(n), (const double*)cv, bg(cn))));
    else
        BOX_DISPATCH((k_box_interpolate<float, false><<<grid, 128, 0, s>>>((float*)fv, bg(n), (const float*)cv, bg(cn))),
                     (k_box_interpolate<double, false><<<grid, 128, 0, s>>>((double*)fv, bg(n), (const double*)cv, bg(cn))));
    return ok();
}

int mgk3b_apply_correction(cudaStream_t s, int dtype, void* fv, const void* err, const int n[3])
{
    if (n[0] < 3 || n[1] < 3 || n[2] < 3) return 0;
    const dim3 grid = grid_for(n[0] - 2, n[1] - 2, n[2] - 2);
    BOX_DISPATCH((k_box_apply_correction<float><<<grid, 128, 0, s>>>((float*)fv, (const float*)err, bg(n))),
                 (k_box_apply_correction<double><<<grid, 128, 0, s>>>((double*)fv, (const double*)err, bg(n))));
    return ok();
}

int mgk3b_set(cudaStream_t s, int dtype, void* a, const int n[3], double value, int modify_boundaries)
{
    const dim3 grid = grid_for(n[0], n[1], n[2]);
    BOX_DISPATCH((k_box_set<float><<<grid, 128, 0, s>>>((float*)a, bg(n), (float)value, modify_boundaries)),
                 (k_box_set<double><<<grid, 128, 0, s>>>((double*)a, bg(n), value, modify_boundaries)));
    return ok();
}

int mgk3b_init_f(cudaStream_t s, int dtype, void* f, const int n[3], const double* sx, const double* sy, const double* sz)
{
    const dim3 grid = grid_for(n[0], n[1], n[2]);
    BOX_DISPATCH((k_box_init_f<float><<<grid, 128, 0, s>>>((float*)f, bg(n), sx, sy, sz)),
                 (k_box_init_f<double><<<grid, 128, 0, s>>>((double*)f, bg(n), sx, sy, sz)));
    return ok();
}

/* parts: 2 * nparts doubles of scratch; out2 = {sum r^2, max |r|} */
int mgk3b_residual_norm(cudaStream_t s, int dtype, const void* v, const void* f, const int n[3], mg_coef3d c, int corrected, double* parts,
                        int nparts, double* out2)
{
    BOX_DISPATCH((k_box_norm<float><<<nparts, 256, 0, s>>>((const float*)v, (const float*)f, bg(n), narrow<float>(c), corrected, parts, nparts)),
                 (k_box_norm<double><<<nparts, 256, 0, s>>>((const double*)v, (const double*)f, bg(n), narrow<double>(c), corrected, parts, nparts)));
    if (ok() < 0) return -1;
    k_box_norm_final<<<1, 32, 0, s>>>(parts, nparts, out2);
    return ok() < 0 ? -1 : 2;
}

}  // extern "C"
