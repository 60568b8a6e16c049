// mg3d_kernels.cu -- sm_100a kernels of the 3D Poisson multigrid hot path.
//
// Replaces the operators of the reference class MultiGrid3D (CPU: NOCUDA_TESI/POISSON_3D(TESI)/
// MultiGrid3D.cpp; GPU twin: CUDA_TESI/CUDA Poisson 3D/MultiGrid3D.cu:362-793).  Arithmetic follows
// the reference expression by expression through the non-contracting helpers of
// mg_exact.cuh, so results are bit-identical to the CPU solver for float and double.
//
// All stages are HBM-bound stencils (<= 2 flop/B): no tensor cores.  Layout: see mg_launch.h.
#include <stdint.h>

#include "mg_exact.cuh"
#include "mg_launch.h"

using namespace mgx;

namespace {

template <typename T>
struct Coef3 {
    T hx2, hy2, hz2, cx, cy, cz, den, rden;
};

template <typename T>
Coef3<T> narrow(const mg_coef3d& c)
{
    Coef3<T> r;
    r.hx2 = (T)c.hx2; r.hy2 = (T)c.hy2; r.hz2 = (T)c.hz2;
    r.cx = (T)c.cx; r.cy = (T)c.cy; r.cz = (T)c.cz;
    r.den = (T)c.den; r.rden = (T)c.rden;
    return r;
}

// N3/MultiGrid3D.cpp:532 -- left-to-right sum of the six weighted neighbours, minus f*hx2*hy2*hz2,
// true division by 2*(hy2*hz2 + hx2*hz2 + hx2*hy2).  O/E = x-1/x+1, N/S = y-1/y+1, D/U = z-1/z+1.
template <typename T>
__device__ __forceinline__ T relax_point(T O, T E, T N, T S, T D, T U, T f, const Coef3<T>& c)
{
    T s = add(mul(O, c.cx), mul(E, c.cx));
    s = add(s, mul(N, c.cy));
    s = add(s, mul(S, c.cy));
    s = add(s, mul(D, c.cz));
    s = add(s, mul(U, c.cz));
    s = sub(s, mul(mul(mul(f, c.hx2), c.hy2), c.hz2));
    return div(s, c.den);
}

// N3/MultiGrid3D.cpp:723 (REF_COMPAT, minus S / minus U) or the sign-corrected form.
template <typename T>
__device__ __forceinline__ T residual_point(T O, T E, T N, T S, T D, T U, T vc, T f, const Coef3<T>& c, int corrected)
{
    T v2 = mul(T(2), vc);
    T tx = div(add(sub(O, v2), E), c.hx2);
    T ty, tz;
    if (corrected) {
        ty = div(add(sub(N, v2), S), c.hy2);
        tz = div(add(sub(D, v2), U), c.hz2);
    } else {
        ty = div(sub(sub(N, v2), S), c.hy2);
        tz = div(sub(sub(D, v2), U), c.hz2);
    }
    return sub(sub(sub(f, tx), ty), tz);
}

// N3/MultiGrid3D.cpp:180 with the exact grouping.  R(dx,dy,dz) reads the fine value at offset
// (dx,dy,dz) from the fine centre; reference names: suffix _C dy=0, _N dy=-1, _S dy=+1;
// N dz=+1, S dz=-1, E dx=+1, O dx=-1.
template <typename T, typename Getter>
__device__ __forceinline__ T restrict_point(Getter R)
{
    T C_C = R(0, 0, 0), N_C = R(0, 0, 1), S_C = R(0, 0, -1), E_C = R(1, 0, 0), O_C = R(-1, 0, 0);
    T NE_C = R(1, 0, 1), NO_C = R(-1, 0, 1), SE_C = R(1, 0, -1), SO_C = R(-1, 0, -1);
    T C_N = R(0, -1, 0), N_N = R(0, -1, 1), S_N = R(0, -1, -1), E_N = R(1, -1, 0), O_N = R(-1, -1, 0);
    T NE_N = R(1, -1, 1), NO_N = R(-1, -1, 1), SE_N = R(1, -1, -1), SO_N = R(-1, -1, -1);
    T C_S = R(0, 1, 0), N_S = R(0, 1, 1), S_S = R(0, 1, -1), E_S = R(1, 1, 0), O_S = R(-1, 1, 0);
    T NE_S = R(1, 1, 1), NO_S = R(-1, 1, 1), SE_S = R(1, 1, -1), SO_S = R(-1, 1, -1);

    T t1 = mul(T(1 / 8.0f), C_C);
    T t2 = mul(T(1 / 16.0f), add(add(add(add(N_C, E_C), S_C), O_C), add(C_N, C_S)));
    T g1 = add(add(add(NE_C, SE_C), SO_C), NO_C);
    T g2 = add(add(add(N_N, E_N), S_N), O_N);
    T g3 = add(add(add(N_S, E_S), S_S), O_S);
    T t3 = mul(T(1 / 32.0f), add(add(g1, g2), g3));
    T h1 = add(add(add(NE_N, SE_N), SO_N), NO_N);
    T h2 = add(add(add(NE_S, SE_S), SO_S), NO_S);
    T t4 = mul(T(1 / 64.0f), add(h1, h2));
    return add(add(add(t1, t2), t3), t4);
}

// N3/MultiGrid3D.cpp:216-331: trilinear prolongation by parity of (y,x,z); c points at the coarse
// value (cx,cy,cz) = (fx/2, fy/2, fz/2); summation orders as written in the reference.
template <typename T>
__device__ __forceinline__ T interp_point(const T* __restrict__ c, int cp, long long cq, int ox, int oy, int oz)
{
    if (!oz) {
        if (!oy) {
            if (!ox) return c[0];                                    // PPP :216
            return mul(T(0.5f), add(c[0], c[1]));                    // PDP :222  O + E
        }
        if (!ox) return mul(T(0.5f), add(c[0], c[cp]));              // DPP :233  N + S
        return mul(T(0.25f), add(add(add(c[0], c[1]), c[cp]), c[cp + 1]));  // DDP :244  NO+NE+SO+SE
    }
    if (!oy) {
        if (!ox) return mul(T(0.5f), add(c[0], c[cq]));              // PPD :261  S + N
        return mul(T(0.25f), add(add(add(c[cq], c[cq + 1]), c[0]), c[1]));  // PDD :272
    }
    if (!ox) return mul(T(0.25f), add(add(add(c[0], c[cq]), c[cp]), c[cp + cq]));  // DPD :287
    // DDD :302  USO + UNO + UNE + USE + DSO + DNO + DNE + DSE
    T s = add(c[0], c[cq]);
    s = add(s, c[cq + 1]);
    s = add(s, c[1]);
    s = add(s, c[cp]);
    s = add(s, c[cp + cq]);
    s = add(s, c[cp + cq + 1]);
    s = add(s, c[cp + 1]);
    return mul(T(0.125f), s);
}

// ---------------------------------------------------------------------------------------------
// Smoother, one colour per launch, in place (MG_SMOOTHER_COLOUR).  Red points only read black
// neighbours and vice versa, so there is no race (unlike the reference's CUDARelax, which updates
// both colours in one launch behind a block-local barrier, C3/MultiGrid3D.cu:635-673).
// Thread t of a row owns the t-th point of the requested colour: x = 2t+1 or 2t+2.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_relax_colour(T* __restrict__ v, const T* __restrict__ f, mg_geom3d g, Coef3<T> c, int colour,
                               int zl_lo)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    const int z = g.z0 + zl;
    if (y > g.n - 2) return;
    const int x = 2 * t + 2 - ((y + z + colour) & 1);  // x = (y+z+colour) mod 2, x >= 1
    if (x > g.n - 2) return;
    const long long i = (long long)zl * g.plane + (long long)y * g.pitch + x;
    T O = v[i - 1], E = v[i + 1], N = v[i - g.pitch], S = v[i + g.pitch], D = v[i - g.plane], U = v[i + g.plane];
    v[i] = relax_point<T>(O, E, N, S, D, U, f[i], c);
}

// ---------------------------------------------------------------------------------------------
// CalculateResidual into a full array (only used when the caller asks for the residual itself).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_residual(const T* __restrict__ v, const T* __restrict__ f, T* __restrict__ r, mg_geom3d g,
                           Coef3<T> c, int corrected, int zl_lo)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    const int z = g.z0 + zl;
    if (x >= g.n || y >= g.n) return;
    const long long i = (long long)zl * g.plane + (long long)y * g.pitch + x;
    if (x == 0 || x == g.n - 1 || y == 0 || y == g.n - 1 || z == 0 || z == g.n - 1) {
        r[i] = T(0);
        return;
    }
    r[i] = residual_point<T>(v[i - 1], v[i + 1], v[i - g.pitch], v[i + g.pitch], v[i - g.plane], v[i + g.plane], v[i],
                             f[i], c, corrected);
}

// ---------------------------------------------------------------------------------------------
// Residual norms: sum r^2 (fp64) and max |r|, residual recomputed on the fly, warp-shuffle
// reduction, one partial per block (deterministic), finished by k_norm_final.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_reduce_sum_max(double& s, double& m, double* sh)
{
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) { sh[w] = s; sh[32 + w] = m; }
    __syncthreads();
    if (w == 0) {
        s = (l < nw) ? sh[l] : 0.0;
        m = (l < nw) ? sh[32 + l] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        }
    }
}

template <typename T>
__global__ void k_residual_norm(const T* __restrict__ v, const T* __restrict__ f, mg_geom3d g, Coef3<T> c,
                                int corrected, int zl_lo, int zl_hi, double* __restrict__ part)
{
    __shared__ double sh[64];
    double s = 0.0, m = 0.0;
    const int ni = g.n - 2;
    const long long rows = (long long)ni * (zl_hi - zl_lo);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int zl = zl_lo + (int)(row / ni);
        const int y = 1 + (int)(row % ni);
        const int z = g.z0 + zl;
        if (z == 0 || z == g.n - 1) continue;
        const long long base = (long long)zl * g.plane + (long long)y * g.pitch;
        for (int x = 1 + threadIdx.x; x <= g.n - 2; x += blockDim.x) {
            const long long i = base + x;
            T r = residual_point<T>(v[i - 1], v[i + 1], v[i - g.pitch], v[i + g.pitch], v[i - g.plane], v[i + g.plane],
                                    v[i], f[i], c, corrected);
            double rd = (double)r;
            s += rd * rd;
            m = fmax(m, fabs(rd));
        }
    }
    block_reduce_sum_max(s, m, sh);
    if (threadIdx.x == 0) { part[blockIdx.x] = s; part[gridDim.x + blockIdx.x] = m; }
}

__global__ void k_norm_final(const double* __restrict__ part, int nparts, double* __restrict__ out2)
{
    __shared__ double sh[64];
    double s = 0.0, m = 0.0;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        s += part[i];
        m = fmax(m, part[nparts + i]);
    }
    block_reduce_sum_max(s, m, sh);
    if (threadIdx.x == 0) { out2[0] = s; out2[1] = m; }
}

// ---------------------------------------------------------------------------------------------
// Restrict (27-point full weighting, boundary injection), one thread per coarse point.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_restrict(const T* __restrict__ fine, mg_geom3d gf, T* __restrict__ coarse, mg_geom3d gc, int czl_lo)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x;
    const int cy = blockIdx.y * blockDim.y + threadIdx.y;
    const int czl = czl_lo + blockIdx.z;
    const int cz = gc.z0 + czl;
    if (cx >= gc.n || cy >= gc.n) return;
    const long long ci = (long long)czl * gc.plane + (long long)cy * gc.pitch + cx;
    const long long fi = (long long)(2 * cz - gf.z0) * gf.plane + (long long)(2 * cy) * gf.pitch + 2 * cx;
    if (cx == 0 || cx == gc.n - 1 || cy == 0 || cy == gc.n - 1 || cz == 0 || cz == gc.n - 1) {
        coarse[ci] = fine[fi];  // N3/MultiGrid3D.cpp:113-119
        return;
    }
    const T* p = fine + fi;
    const int fp = gf.pitch;
    const long long fq = gf.plane;
    coarse[ci] = restrict_point<T>([&](int dx, int dy, int dz) { return p[dx + dy * fp + dz * fq]; });
}

// ---------------------------------------------------------------------------------------------
// Fused CalculateResidual -> Restrict -> coarse f, plus setToValue(coarse v, 0, true).
// A CTA owns a CX x CY x CZ tile of coarse points; it evaluates the fine residual on the
// (2CX+1)(2CY+1)(2CZ+1) fine points around the tile into shared memory (fine boundary = 0 as in
// CalculateResidual) and restricts from there.  The fine residual never goes to HBM.
// ---------------------------------------------------------------------------------------------
template <typename T, int CX, int CY, int CZ>
__global__ void __launch_bounds__(256)
k_residual_restrict(const T* __restrict__ v, const T* __restrict__ f, mg_geom3d gf, Coef3<T> c, int corrected,
                    T* __restrict__ cf, T* __restrict__ cv, mg_geom3d gc, int czl_lo, int czl_hi)
{
    constexpr int FX = 2 * CX + 1, FY = 2 * CY + 1, FZ = 2 * CZ + 1;
    __shared__ T r[FZ][FY][FX];

    const int cx0 = blockIdx.x * CX, cy0 = blockIdx.y * CY;
    const int czl0 = czl_lo + blockIdx.z * CZ;   // first coarse local plane of the tile
    const int cz0 = gc.z0 + czl0;                // global
    const int fx0 = 2 * cx0 - 1, fy0 = 2 * cy0 - 1, fz0 = 2 * cz0 - 1;  // global fine origin of the smem tile
    const int n = gf.n;

    for (int i = threadIdx.x; i < FX * FY * FZ; i += 256) {
        const int lx = i % FX, ly = (i / FX) % FY, lz = i / (FX * FY);
        const int fx = fx0 + lx, fy = fy0 + ly, fz = fz0 + lz;
        T val = T(0);
        if (fx >= 1 && fx <= n - 2 && fy >= 1 && fy <= n - 2 && fz >= 1 && fz <= n - 2) {
            const int fzl = fz - gf.z0;
            if (fzl >= 1 && fzl <= gf.nzl - 2) {
                const long long idx = (long long)fzl * gf.plane + (long long)fy * gf.pitch + fx;
                val = residual_point<T>(v[idx - 1], v[idx + 1], v[idx - gf.pitch], v[idx + gf.pitch], v[idx - gf.plane],
                                        v[idx + gf.plane], v[idx], f[idx], c, corrected);
            }
        }
        r[lz][ly][lx] = val;
    }
    __syncthreads();

    for (int i = threadIdx.x; i < CX * CY * CZ; i += 256) {
        const int tx = i % CX, ty = (i / CX) % CY, tz = i / (CX * CY);
        const int cx = cx0 + tx, cy = cy0 + ty, czl = czl0 + tz, cz = cz0 + tz;
        if (cx >= gc.n || cy >= gc.n || czl >= czl_hi) continue;
        const long long ci = (long long)czl * gc.plane + (long long)cy * gc.pitch + cx;
        T out = T(0);  // boundary: injection of the (zero) boundary residual, N3/MultiGrid3D.cpp:113-119 + :705
        if (!(cx == 0 || cx == gc.n - 1 || cy == 0 || cy == gc.n - 1 || cz == 0 || cz == gc.n - 1)) {
            const int lx = 2 * tx + 1, ly = 2 * ty + 1, lz = 2 * tz + 1;
            out = restrict_point<T>([&](int dx, int dy, int dz) { return r[lz + dz][ly + dy][lx + dx]; });
        }
        cf[ci] = out;
        cv[ci] = T(0);  // setToValue(coarse->h_v, 0, true), N3/MultiGrid3D.cpp:634
    }
}

// ---------------------------------------------------------------------------------------------
// Interpolate (+ ApplyCorrection when add != 0).  A thread owns the fine pair (2i, 2i+1) of a row,
// so x-parity is compile-time per statement and (y,z) parity is uniform per row: no divergence.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_interpolate(T* __restrict__ fine, mg_geom3d gf, const T* __restrict__ coarse, mg_geom3d gc, int add_,
                              int zl_lo)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // pair index: x = 2i, 2i+1
    const int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    const int z = gf.z0 + zl;
    if (y > gf.n - 2 || 2 * i > gf.n - 2) return;
    const int oy = y & 1, oz = z & 1;
    const int cy = y >> 1, czl = (z >> 1) - gc.z0;
    const T* c = coarse + (long long)czl * gc.plane + (long long)cy * gc.pitch + i;
    T* p = fine + (long long)zl * gf.plane + (long long)y * gf.pitch + 2 * i;
    if (i >= 1) {  // x = 2i even, interior
        T e = interp_point<T>(c, gc.pitch, gc.plane, 0, oy, oz);
        p[0] = add_ ? add(p[0], e) : e;
    }
    if (2 * i + 1 <= gf.n - 2) {
        T e = interp_point<T>(c, gc.pitch, gc.plane, 1, oy, oz);
        p[1] = add_ ? add(p[1], e) : e;
    }
}

template <typename T>
__global__ void k_apply_correction(T* __restrict__ fine, const T* __restrict__ err, mg_geom3d g, int zl_lo)
{
    const int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    if (x > g.n - 2 || y > g.n - 2) return;
    const long long i = (long long)zl * g.plane + (long long)y * g.pitch + x;
    fine[i] = add(fine[i], err[i]);
}

template <typename T>
__global__ void k_set(T* __restrict__ a, mg_geom3d g, T value, int modify_boundaries, int zl_lo)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    const int z = g.z0 + zl;
    if (x >= g.n || y >= g.n) return;
    if (!modify_boundaries && (x == 0 || x == g.n - 1 || y == 0 || y == g.n - 1 || z == 0 || z == g.n - 1)) return;
    a[(long long)zl * g.plane + (long long)y * g.pitch + x] = value;
}

// N3/Grid3D.cpp:92: h_f = -3*PI*PI*sin(PI*x)*sin(PI*y)*sin(PI*z) evaluated left to right in double,
// narrowed to T.  The sines come from host libm tables so they are the reference's own values.
template <typename T>
__global__ void k_init_f(T* __restrict__ f, mg_geom3d g, const double* __restrict__ sx, const double* __restrict__ sy,
                         const double* __restrict__ sz, int zl_lo)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int zl = zl_lo + blockIdx.z;
    const int z = g.z0 + zl;
    if (x >= g.n || y >= g.n) return;
    const double PI = 3.141592653589793;
    const double k = __dmul_rn(__dmul_rn(-3.0, PI), PI);
    f[(long long)zl * g.plane + (long long)y * g.pitch + x] = (T)__dmul_rn(__dmul_rn(__dmul_rn(k, sx[x]), sy[y]), sz[z]);
}

template <typename T>
__global__ void k_copy_rows(T* __restrict__ dst, long long dpitch, const T* __restrict__ src, long long spitch, int width,
                            long long rows)
{
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const T* s = src + r * spitch;
        T* d = dst + r * dpitch;
        for (int x = threadIdx.x; x < width; x += blockDim.x) d[x] = s[x];
    }
}

inline int launch_ok() { return cudaPeekAtLastError() == cudaSuccess ? 1 : -1; }

inline dim3 block2d(int nx) { int bx = nx >= 128 ? 128 : (nx >= 64 ? 64 : 32); return dim3(bx, 256 / bx, 1); }

template <typename T>
int relax_colour_t(cudaStream_t s, T* v, const T* f, mg_geom3d g, mg_coef3d c, int colour, int zl_lo, int zl_hi)
{
    if (zl_hi <= zl_lo || g.n < 3) return 0;
    const int halfw = (g.n - 1) / 2;
    int bx = halfw < 128 ? halfw : 128;
    int by = 256 / bx;
    if (by > g.n - 2) by = g.n - 2;
    if (by < 1) by = 1;
    dim3 block(bx, by, 1), grid((halfw + bx - 1) / bx, (g.n - 2 + by - 1) / by, zl_hi - zl_lo);
    k_relax_colour<T><<<grid, block, 0, s>>>(v, f, g, narrow<T>(c), colour, zl_lo);
    return launch_ok();
}

}  // namespace

#define DISPATCH(dtype, call_f32, call_f64) ((dtype) == 0 ? (call_f32) : (call_f64))

extern "C" {

int mgk3d_relax_colour(cudaStream_t s, int dtype, void* v, const void* f, mg_geom3d g, mg_coef3d c, int colour,
                       int zl_lo, int zl_hi)
{
    return DISPATCH(dtype, relax_colour_t<float>(s, (float*)v, (const float*)f, g, c, colour, zl_lo, zl_hi),
                    relax_colour_t<double>(s, (double*)v, (const double*)f, g, c, colour, zl_lo, zl_hi));
}

int mgk3d_residual(cudaStream_t s, int dtype, const void* v, const void* f, void* r, mg_geom3d g, mg_coef3d c,
                   int corrected, int zl_lo, int zl_hi)
{
    if (zl_hi <= zl_lo) return 0;
    dim3 block = block2d(g.n), grid((g.n + block.x - 1) / block.x, (g.n + block.y - 1) / block.y, zl_hi - zl_lo);
    if (dtype == 0)
        k_residual<float><<<grid, block, 0, s>>>((const float*)v, (const float*)f, (float*)r, g, narrow<float>(c), corrected, zl_lo);
    else
        k_residual<double><<<grid, block, 0, s>>>((const double*)v, (const double*)f, (double*)r, g, narrow<double>(c), corrected, zl_lo);
    return launch_ok();
}

int mgk3d_residual_norm(cudaStream_t s, int dtype, const void* v, const void* f, mg_geom3d g, mg_coef3d c,
                        int corrected, int zl_lo, int zl_hi, double* scratch, double* out2)
{
    const int nb = MGK_NORM_BLOCKS;
    if (dtype == 0)
        k_residual_norm<float><<<nb, 256, 0, s>>>((const float*)v, (const float*)f, g, narrow<float>(c), corrected, zl_lo, zl_hi, scratch);
    else
        k_residual_norm<double><<<nb, 256, 0, s>>>((const double*)v, (const double*)f, g, narrow<double>(c), corrected, zl_lo, zl_hi, scratch);
    k_norm_final<<<1, 256, 0, s>>>(scratch, nb, out2);
    return launch_ok() < 0 ? -1 : 2;
}

int mgk3d_restrict(cudaStream_t s, int dtype, const void* fine, mg_geom3d gf, void* coarse, mg_geom3d gc, int czl_lo,
                   int czl_hi)
{
    if (czl_hi <= czl_lo) return 0;
    dim3 block = block2d(gc.n), grid((gc.n + block.x - 1) / block.x, (gc.n + block.y - 1) / block.y, czl_hi - czl_lo);
    if (dtype == 0)
        k_restrict<float><<<grid, block, 0, s>>>((const float*)fine, gf, (float*)coarse, gc, czl_lo);
    else
        k_restrict<double><<<grid, block, 0, s>>>((const double*)fine, gf, (double*)coarse, gc, czl_lo);
    return launch_ok();
}

int mgk3d_residual_restrict(cudaStream_t s, int dtype, const void* v, const void* f, mg_geom3d gf, mg_coef3d c,
                            int corrected, void* coarse_f, void* coarse_v, mg_geom3d gc, int czl_lo, int czl_hi)
{
    if (czl_hi <= czl_lo) return 0;
    constexpr int CX = 16, CY = 8, CZ = 4;
    dim3 grid((gc.n + CX - 1) / CX, (gc.n + CY - 1) / CY, (czl_hi - czl_lo + CZ - 1) / CZ);
    if (dtype == 0)
        k_residual_restrict<float, CX, CY, CZ><<<grid, 256, 0, s>>>((const float*)v, (const float*)f, gf, narrow<float>(c), corrected,
                                                                    (float*)coarse_f, (float*)coarse_v, gc, czl_lo, czl_hi);
    else
        k_residual_restrict<double, CX, CY, CZ><<<grid, 256, 0, s>>>((const double*)v, (const double*)f, gf, narrow<double>(c), corrected,
                                                                     (double*)coarse_f, (double*)coarse_v, gc, czl_lo, czl_hi);
    return launch_ok();
}

int mgk3d_interpolate(cudaStream_t s, int dtype, void* fine, mg_geom3d gf, const void* coarse, mg_geom3d gc, int add,
                      int zl_lo, int zl_hi)
{
    if (zl_hi <= zl_lo || gf.n < 3) return 0;
    const int pairs = (gf.n - 1) / 2;  // pair i covers x = 2i, 2i+1 <= n-2
    int bx = pairs < 128 ? pairs : 128;
    int by = 256 / bx;
    if (by > gf.n - 2) by = gf.n - 2;
    if (by < 1) by = 1;
    dim3 block(bx, by, 1), grid((pairs + bx - 1) / bx, (gf.n - 2 + by - 1) / by, zl_hi - zl_lo);
    if (dtype == 0)
        k_interpolate<float><<<grid, block, 0, s>>>((float*)fine, gf, (const float*)coarse, gc, add, zl_lo);
    else
        k_interpolate<double><<<grid, block, 0, s>>>((double*)fine, gf, (const double*)coarse, gc, add, zl_lo);
    return launch_ok();
}

int mgk3d_apply_correction(cudaStream_t s, int dtype, void* fine, const void* err, mg_geom3d g, int zl_lo, int zl_hi)
{
    if (zl_hi <= zl_lo || g.n < 3) return 0;
    dim3 block = block2d(g.n), grid((g.n - 2 + block.x - 1) / block.x, (g.n - 2 + block.y - 1) / block.y, zl_hi - zl_lo);
    if (dtype == 0)
        k_apply_correction<float><<<grid, block, 0, s>>>((float*)fine, (const float*)err, g, zl_lo);
    else
        k_apply_correction<double><<<grid, block, 0, s>>>((double*)fine, (const double*)err, g, zl_lo);
    return launch_ok();
}

int mgk3d_set(cudaStream_t s, int dtype, void* a, mg_geom3d g, double value, int modify_boundaries, int zl_lo,
              int zl_hi)
{
    if (zl_hi <= zl_lo) return 0;
    dim3 block = block2d(g.n), grid((g.n + block.x - 1) / block.x, (g.n + block.y - 1) / block.y, zl_hi - zl_lo);
    if (dtype == 0)
        k_set<float><<<grid, block, 0, s>>>((float*)a, g, (float)value, modify_boundaries, zl_lo);
    else
        k_set<double><<<grid, block, 0, s>>>((double*)a, g, value, modify_boundaries, zl_lo);
    return launch_ok();
}

int mgk_copy_rows(cudaStream_t s, int dtype, void* dst, long long dpitch, const void* src, long long spitch, int width,
                  long long rows)
{
    if (rows <= 0 || width <= 0) return 0;
    const int threads = width >= 256 ? 256 : (width >= 64 ? 64 : 32);
    const long long want = rows < (1LL << 20) ? rows : (1LL << 20);
    if (dtype == 0)
        k_copy_rows<float><<<(unsigned)want, threads, 0, s>>>((float*)dst, dpitch, (const float*)src, spitch, width, rows);
    else
        k_copy_rows<double><<<(unsigned)want, threads, 0, s>>>((double*)dst, dpitch, (const double*)src, spitch, width, rows);
    return launch_ok();
}

int mgk3d_init_f(cudaStream_t s, int dtype, void* f, mg_geom3d g, const double* sx, const double* sy,
                 const double* sz, int zl_lo, int zl_hi)
{
    if (zl_hi <= zl_lo) return 0;
    dim3 block = block2d(g.n), grid((g.n + block.x - 1) / block.x, (g.n + block.y - 1) / block.y, zl_hi - zl_lo);
    if (dtype == 0)
        k_init_f<float><<<grid, block, 0, s>>>((float*)f, g, sx, sy, sz, zl_lo);
    else
        k_init_f<double><<<grid, block, 0, s>>>((double*)f, g, sx, sy, sz, zl_lo);
    return launch_ok();
}

}  // extern "C"
