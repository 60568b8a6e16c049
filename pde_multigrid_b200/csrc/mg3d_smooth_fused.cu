// mg3d_smooth_fused.cu -- temporally blocked smoother: TWO full red-black Gauss-Seidel sweeps (four
// half-sweeps) in ONE pass over HBM (MG_SMOOTHER_FUSED; replaces two calls' worth of
// MultiGrid3D::Relax, N3/MultiGrid3D.cpp:489-567, i.e. exactly the nu = 2 pre- or post-smoothing of a V(2,2)).
//
// The two-pass smoother (mg3d_smooth_tma.cu) is already at 0.96 of the HBM roofline of its own traffic
// (12 B/point per half-sweep in fp64 = 48 B/point for two sweeps).  This kernel moves v once in and once
// out and f once in: 24 B/point for the same two sweeps.
//
// A CTA owns an (TI half-indices x TY rows) column and marches along z.  Per plane, TMA brings both colour
// sub-tiles of v and of f (tile + 4 points of halo, zero-filled outside the grid) into an 8-slot shared-memory
// ring.  The four half-sweeps run as a software pipeline skewed along z, in place in shared memory:
//     phase A:  red,   sweep 1 on plane p-1   and   red,   sweep 2 on plane p-4      __syncthreads
//     phase B:  black, sweep 1 on plane p-2   and   black, sweep 2 on plane p-5      __syncthreads
// (each stage only reads the other colour, on planes that are in exactly the state the sequential algorithm
// would show it; stage k is restricted to the tile grown by 3-k points so that halo values computed
// redundantly are never consumed beyond their validity).  Plane p-5 is then final and is written to the
// OUTPUT field (out of place: neighbouring CTAs still need the old values of their halos).  Arithmetic per
// point is relax_point of mg3d_device.cuh, so the result is bit-identical to four colour launches.
//
// MEASURED (1025^3, one B200): 13.0 ms per two-sweep pass in fp64 and 11.1 ms in fp32, against 8.4 / 4.5 ms for the
// four colour launches it replaces -- NOT the default.  The pass moves 33 GB instead of 51.6 GB, but every stencil
// value is a shared-memory load (7 LDS + 1 STS per update, four stages per plane): the LSU data pipe saturates, the
// same limiter the residual+restrict kernel had before its register tiling (mg3d_rr_tma.cu).  Next step, worked
// out but not built: register tiling of this pipeline -- thread (i, y) computes all four stages of its own column,
// so D/U, the own-position O-or-E and (with two rows per thread) one of N/S are registers of the thread's own
// z-windows (raw colour 1: planes p-1..p+1, stage 1: p-3..p, stage 2: p-4..p-1, stage 3: p-5..p-3); 3 LDS + 1 STS
// per update through single-plane exchange buffers, raw v and f read by plain coalesced loads.  In fp64 the
// ceiling is modest (~3.6 ms of fp64 pipe + ~4.5 ms of issue slots per pass).
#include "mg3d_device.cuh"
#include "mg_tma.cuh"

using namespace mgx;
using namespace mg3;
using namespace mgtma;

namespace {

constexpr int TI = MGK3D_FU_TI;  // half-indices per tile (64 grid points in x)
constexpr int TY = MGK3D_FU_TY;  // rows per tile
constexpr int HALO = 4;          // grid points of halo: four half-sweeps
constexpr int BH = TY + 2 * HALO;
constexpr int NS = 8;            // ring slots: planes p+2, p+1 (in flight), p ... p-5
constexpr int NT = 1024;           // 32 warps, SPT = 1 slot per thread and stage (measured: 512x2 15.0 ms, 1024x1 14.0 ms, 256x4 17.6 ms per pass at 1025^3)
constexpr int SPT = ((TY + 6) * (TI + 4) + NT - 1) / NT;  // slots per thread and stage (4)

template <typename T> struct FBox {
    static constexpr int A = 16 / sizeof(T);   // half-indices loaded left of the tile (16-byte aligned TMA start)
    static constexpr int W = TI + 2 * A;       // 36 doubles / 40 floats; needed: half-indices i0-2 .. i0+33
    static constexpr int SUB = W * BH;         // elements of one colour sub-tile
    static constexpr int SUB_STRIDE = (SUB * (int)sizeof(T) + 127) / 128 * 128 / (int)sizeof(T);
    static constexpr int SLOT = 4 * SUB_STRIDE;  // v colour 0, v colour 1, f colour 0, f colour 1
};

struct FusedMaps {
    CUtensorMap v[2];
    CUtensorMap f[2];
};

template <typename T, bool FAST>
__global__ void __launch_bounds__(NT, 1)
k_relax_fused2(const __grid_constant__ FusedMaps maps, T* __restrict__ v_out, mg_geom3d g, Coef3<T> c, int zchunk, int zlo, int zhi,
               unsigned int* cond)
{
    // conditional launch (the exact fallback of mg3d_smooth_pipe.cu): nothing to do unless the flag was raised.  Nobody
    // changes the flag while this grid runs (the last CTA to finish clears it), so every CTA takes the same branch.
    if (cond && *reinterpret_cast<volatile unsigned int*>(cond) == 0u) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int A = FBox<T>::A, W = FBox<T>::W, SUBS = FBox<T>::SUB_STRIDE, SLOT = FBox<T>::SLOT;
    constexpr uint32_t SUB_BYTES = FBox<T>::SUB * sizeof(T);
    T* ring = reinterpret_cast<T*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NS * SLOT * sizeof(T));

    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * TI, X0 = 2 * i0;
    const int y0 = blockIdx.y * TY;
    const int zs = zlo + blockIdx.z * zchunk, ze = min(zs + zchunk, zhi);  // output planes [zs, ze), local indices
    const int n = g.n;

    if (tid == 0) {
        prefetch_tensormap(&maps.v[0]);
        prefetch_tensormap(&maps.v[1]);
        prefetch_tensormap(&maps.f[0]);
        prefetch_tensormap(&maps.f[1]);
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    const int pb = zs - HALO;        // first raw plane
    const int plast = ze + HALO - 1; // last raw plane that is loaded
    auto slot_of = [&](int p) { return ring + (size_t)((p - pb) & (NS - 1)) * SLOT; };
    auto issue = [&](int p) {
        const int s = (p - pb) & (NS - 1);
        T* dst = ring + (size_t)s * SLOT;
        mbar_arrive_expect_tx(&bars[s], 4 * SUB_BYTES);
        tma_load_3d(dst, &maps.v[0], &bars[s], i0 - A, y0 - HALO, p);
        tma_load_3d(dst + SUBS, &maps.v[1], &bars[s], i0 - A, y0 - HALO, p);
        tma_load_3d(dst + 2 * SUBS, &maps.f[0], &bars[s], i0 - A, y0 - HALO, p);
        tma_load_3d(dst + 3 * SUBS, &maps.f[1], &bars[s], i0 - A, y0 - HALO, p);
    };
    if (tid == 0) {
        issue(pb);
        if (pb + 1 <= plast) issue(pb + 1);
    }

    // Stage k (k = 0..3: red 1, black 1, red 2, black 2) covers the tile grown by e = 3-k points: (TY+2e) rows x 36
    // half-indices, at most one slot per thread.  Everything that does not depend on the plane is decoded once:
    // the shared-memory offset of the slot and three flag bits (bit 0: parity of colour+y; bit 1 / bit 2: the point
    // with x parity 0 / 1 exists, is interior and lies inside the stage's region).
    int cc[4][SPT];
    unsigned fl[4][SPT];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int e = 3 - k, rows = TY + 2 * e, col = k & 1;
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            const int idx = tid + s * NT;
            cc[k][s] = 0;
            fl[k][s] = 0;
            if (idx < rows * (TI + 4)) {
                const int rr = idx / (TI + 4), ii = idx - rr * (TI + 4) - 2;  // ii = half-index - i0 in [-2, TI+2)
                const int r = HALO - e + rr, y = y0 - HALO + r;
                cc[k][s] = r * W + ii + A;
                const bool yok = y >= 1 && y <= n - 2;
                unsigned f = (unsigned)((col + y + g.z0) & 1);
#pragma unroll
                for (int par = 0; par < 2; par++) {
                    const int x = X0 + 2 * ii + par;
                    if (yok && x >= 1 && x <= n - 2 && x >= X0 - e && x < X0 + 2 * TI + e) f |= 2u << par;
                }
                fl[k][s] = f;
            }
        }
    }
    // the thread's elements of the output tile (2 colours x TY rows x TI half-indices), decoded once
    constexpr int OPT = 2 * TY * TI / NT;
    static_assert(OPT * NT == 2 * TY * TI, "output tile must divide evenly");
    int st_src[OPT];
    long long st_dst[OPT];
    bool st_ok[OPT];
#pragma unroll
    for (int s = 0; s < OPT; s++) {
        const int idx = tid + s * NT;
        const int col = idx / (TY * TI), rem = idx - col * (TY * TI);
        const int rr = rem / TI, ii = rem - rr * TI;
        st_src[s] = col * SUBS + (HALO + rr) * W + ii + A;
        st_dst[s] = (long long)col * g.cstride + (long long)(y0 + rr) * g.hp + (i0 + ii);
        st_ok[s] = y0 + rr < n && i0 + ii <= (n - 1) / 2;
    }

    // One stage: colour `col` of plane p-d for the thread's SPT slots: all loads first, then SPT independent
    // arithmetic chains, then the stores.  sb[] holds the ring offsets of planes p..p-6.
    auto stage = [&](int k, int d, int col, bool plane_on, const int (&sb)[7], unsigned zpar) {
        if (!plane_on) return;
        T O[SPT], E[SPT], N[SPT], S[SPT], D[SPT], U[SPT], F[SPT];
        bool on[SPT];
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            const int par = (int)((fl[k][s] ^ zpar ^ (unsigned)d) & 1u);
            on[s] = (fl[k][s] & (2u << par)) != 0;
            O[s] = E[s] = N[s] = S[s] = D[s] = U[s] = F[s] = T(0);
            if (on[s]) {
                const T* oth = ring + sb[d] + (col ^ 1) * SUBS + cc[k][s];
                O[s] = oth[par - 1]; E[s] = oth[par]; N[s] = oth[-W]; S[s] = oth[W];
                D[s] = ring[sb[d + 1] + (col ^ 1) * SUBS + cc[k][s]];
                U[s] = ring[sb[d - 1] + (col ^ 1) * SUBS + cc[k][s]];
                F[s] = ring[sb[d] + (2 + col) * SUBS + cc[k][s]];
            }
        }
#pragma unroll
        for (int s = 0; s < SPT; s++)
            if (on[s]) ring[sb[d] + col * SUBS + cc[k][s]] = relax_point<T, FAST>(O[s], E[s], N[s], S[s], D[s], U[s], F[s], c);
    };

    // (Measured status and the planned register-tiled successor: file header.  Neither fewer instructions -- hoisted
    // ring offsets -- nor more ILP -- 4 slots per thread -- changed the time of this shared-memory version.)
    for (int p = pb; p <= plast + 1; p++) {
        if (p <= plast) mbar_wait(&bars[(p - pb) & (NS - 1)], ((p - pb) / NS) & 1);
        int sb[7];  // element offset of the slot of plane p-j
#pragma unroll
        for (int j = 0; j < 7; j++) sb[j] = ((p - j - pb) & (NS - 1)) * SLOT;
        const unsigned zpar = (unsigned)p & 1u;  // (the slab's z0 is folded into the per-slot parity bit above)
        auto plane_on = [&](int q, int e) { return g.z0 + q >= 1 && g.z0 + q <= n - 2 && q >= zs - e && q < ze + e; };
        // phase A: red of sweep 1 on plane p-1, red of sweep 2 on plane p-4
        stage(0, 1, 0, plane_on(p - 1, 3), sb, zpar);
        stage(2, 4, 0, plane_on(p - 4, 1), sb, zpar);
        __syncthreads();
        // phase B: black of sweep 1 on plane p-2, black of sweep 2 on plane p-5 (reads red of plane p-6)
        stage(1, 2, 1, plane_on(p - 2, 2), sb, zpar);
        stage(3, 5, 1, plane_on(p - 5, 0), sb, zpar);
        __syncthreads();
        if (tid == 0 && p + 2 <= plast) issue(p + 2);  // into the slot of plane p-6, which nobody reads any more
        // plane p-5 is final: write the tile (both colours) to the output field
        const int ps = p - 5;
        if (ps >= zs && ps < ze) {
#pragma unroll
            for (int s = 0; s < OPT; s++)
                if (st_ok[s]) {
                    MG_CHK(ps >= 0 && ps < g.nzl && st_dst[s] >= 0 && st_dst[s] < g.cstride + g.plane);
                    __stcs(v_out + st_dst[s] + (long long)ps * g.plane, ring[sb[5] + st_src[s]]);
                }
        }
    }
    if (cond) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned total = gridDim.x * gridDim.y * gridDim.z;
            if (atomicAdd(cond + 1, 1u) == total - 1) { cond[1] = 0u; __threadfence(); cond[0] = 0u; }
        }
    }
}

template <typename T>
size_t smem_bytes_t() { return (size_t)NS * FBox<T>::SLOT * sizeof(T) + NS * sizeof(uint64_t); }

template <typename T, bool FAST>
int launch_k(cudaStream_t s, const FusedMaps& m, T* v_out, mg_geom3d g, mg_coef3d c, dim3 grid, int zchunk, int zlo, int zhi, unsigned int* cond)
{
    MG_SET_SMEM_LIMIT((k_relax_fused2<T, FAST>), smem_bytes_t<T>());
    k_relax_fused2<T, FAST><<<grid, NT, smem_bytes_t<T>(), s>>>(m, v_out, g, narrow<T>(c), zchunk, zlo, zhi, cond);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

template <typename T>
int launch(cudaStream_t s, const void* const maps4[4], T* v_out, mg_geom3d g, mg_coef3d c, int zlo, int zhi, unsigned int* cond)
{
    if (zhi <= zlo) return 0;
    const int nz = zhi - zlo;
    FusedMaps m;
    memcpy(&m.v[0], maps4[0], sizeof(CUtensorMap));
    memcpy(&m.v[1], maps4[1], sizeof(CUtensorMap));
    memcpy(&m.f[0], maps4[2], sizeof(CUtensorMap));
    memcpy(&m.f[1], maps4[3], sizeof(CUtensorMap));
    const int tx = ((g.n + 1) / 2 + TI - 1) / TI, ty = (g.n + TY - 1) / TY;
    // z chunks: every chunk pays 2*HALO planes of warm-up; more chunks balance the waves of 148 one-CTA SMs
    int nchunk = 1;
    double best = 1e30;
    for (int k = 1; k <= 8; k *= 2) {
        const int zc = (nz + k - 1) / k;
        if (k > 1 && zc < 64) break;
        const long long ctas = (long long)tx * ty * k;
        const double cost = (double)((ctas + 147) / 148) * (zc + 2 * HALO + 2);
        if (cost < best) { best = cost; nchunk = k; }
    }
    const int zchunk = (nz + nchunk - 1) / nchunk;
    dim3 grid(tx, ty, (nz + zchunk - 1) / zchunk);
    if (c.fast_den) return launch_k<T, true>(s, m, v_out, g, c, grid, zchunk, zlo, zhi, cond);
    return launch_k<T, false>(s, m, v_out, g, c, grid, zchunk, zlo, zhi, cond);
}

}  // namespace

/* maps4: tensor maps of {v_in colour 0, v_in colour 1, f colour 0, f colour 1} with box
   (MGK3D_FU_BOX_I(esize), MGK3D_FU_BOX_Y, 1); v_out: the other v buffer of the level (all planes are written) */
extern "C" int mgk3d_relax_fused2(cudaStream_t s, int dtype, const void* const maps4[4], void* v_out, mg_geom3d g, mg_coef3d c,
                                  int zl_lo, int zl_hi, unsigned int* cond)
{
    if (dtype == 0) return launch<float>(s, maps4, (float*)v_out, g, c, zl_lo, zl_hi, cond);
    return launch<double>(s, maps4, (double*)v_out, g, c, zl_lo, zl_hi, cond);
}
