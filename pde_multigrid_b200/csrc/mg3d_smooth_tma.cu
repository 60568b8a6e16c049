// mg3d_smooth_tma.cu -- the hot kernel: red-black Gauss-Seidel half-sweep with TMA-staged, z-marching
// shared-memory tiles (replaces MultiGrid3D::Relax, N3/MultiGrid3D.cpp:489-567, on the large levels).
//
// One launch updates one colour in place, like k_relax_colour, but a CTA owns an (IT half-indices x YT
// rows) column of the grid and marches along z.  The OTHER colour's planes (the only v values a half-
// sweep reads) are streamed into a 4-slot shared-memory ring by TMA (cp.async.bulk.tensor.3d, one
// elected thread, mbarrier transaction barriers): every neighbour value enters the SM once per CTA
// instead of once per use, out-of-bounds halo elements are zero-filled by the hardware (no boundary
// branches in the load path), and the copy of plane z+2 overlaps the arithmetic of plane z.  The six
// neighbour reads of an update are conflict-free LDS; f and the updated colour are streamed straight
// from / to HBM with unit stride.  Traffic per updated point: 8 B (f) + 8 B (write) + 8 B x (YT+2)/YT
// (other colour incl. y-halo, x-halo negligible) -> HBM-bound; arithmetic identical to k_relax_colour.
#include "mg3d_device.cuh"
#include "mg_tma.cuh"

using namespace mgx;
using namespace mg3;
using namespace mgtma;

namespace {

constexpr int IT = MGK3D_TMA_IT;    // half-indices per tile
constexpr int YT = MGK3D_TMA_YT;    // rows per tile
constexpr int BH = MGK3D_TMA_BOX_Y; // box height = YT + 2 (one halo row on each side)
// Box width: the tile needs half-indices i0-1 .. i0+IT.  The innermost TMA coordinate must be 16-byte
// aligned (measured: scripts/probe/tma_probe.cu traps with "illegal instruction" otherwise), so the box
// starts A = 16/sizeof(T) elements before i0 and is IT + 2A wide (132 doubles / 136 floats).
template <typename T> struct Box { static constexpr int A = 16 / sizeof(T); static constexpr int W = IT + 2 * A; };
constexpr int RING = 4;
constexpr int NT = 256;

template <typename T, bool FAST>
__global__ void __launch_bounds__(NT)
k_relax_colour_tma(const __grid_constant__ CUtensorMap map_other, T* __restrict__ v_own, const T* __restrict__ f_own,
                   mg_geom3d g, Coef3<T> c, int colour, int zl_lo, int zl_hi, int zchunk)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int A = Box<T>::A, BW = Box<T>::W;
    constexpr int SLOT_ELEMS = BW * BH;
    constexpr uint32_t SLOT_BYTES = SLOT_ELEMS * sizeof(T);
    constexpr int SLOT_STRIDE = (SLOT_BYTES + 127) / 128 * 128 / sizeof(T);
    T* ring = reinterpret_cast<T*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)RING * SLOT_STRIDE * sizeof(T));

    const int i0 = blockIdx.x * IT;
    const int y0 = 1 + blockIdx.y * YT;
    const int zs = zl_lo + blockIdx.z * zchunk;                 // first local plane this CTA updates
    const int ze = min(zs + zchunk, zl_hi);                     // one past the last
    const int tid = threadIdx.x;

    if (tid == 0) {
        prefetch_tensormap(&map_other);
        for (int s = 0; s < RING; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // planes p = zs-1 .. ze are needed; plane p lives in slot (p - (zs-1)) % RING
    const int pbase = zs - 1;
    auto issue = [&](int p) {
        const int k = p - pbase, s = k % RING;
        mbar_arrive_expect_tx(&bars[s], SLOT_BYTES);
        tma_load_3d(ring + (size_t)s * SLOT_STRIDE, &map_other, &bars[s], i0 - A, y0 - 1, p);
    };
    if (tid == 0)
        for (int p = pbase; p <= min(pbase + RING - 1, ze); p++) issue(p);

    const int il = tid & (IT - 1);
    const int r0 = (tid / IT) * (YT / (NT / IT));  // first row of this thread; it owns YT/(NT/IT) consecutive rows
    constexpr int RPT = YT / (NT / IT);
    const int i = i0 + il;

    // f of the thread's RPT points is prefetched one plane ahead into registers: its HBM latency would
    // otherwise sit exposed between the barrier wait and the arithmetic (ncu: long_scoreboard dominant)
    auto load_f = [&](int z, T (&dst)[RPT]) {
        const long long zbase = (long long)z * g.plane;
#pragma unroll
        for (int rr = 0; rr < RPT; rr++) {
            const int y = y0 + r0 + rr;
            const int x = 2 * i + ((colour + y + g.z0 + z) & 1);
            if (y <= g.n - 2 && x >= 1 && x <= g.n - 2) MG_CHK_SITE(g, i, y, z);
            dst[rr] = (y <= g.n - 2 && x >= 1 && x <= g.n - 2) ? __ldg(f_own + zbase + (long long)y * g.hp + i) : T(0);
        }
    };
    T fnext[RPT];
    load_f(zs, fnext);

    for (int z = zs; z < ze; z++) {
        const int k = z - pbase;  // ring index of plane z; planes z-1, z, z+1 -> k-1, k, k+1
        T fcur[RPT];
#pragma unroll
        for (int rr = 0; rr < RPT; rr++) fcur[rr] = fnext[rr];
        if (z + 1 < ze) load_f(z + 1, fnext);
        // planes k-1 and k were waited for in earlier iterations (or right here on the first one)
        if (z == zs) {
            mbar_wait(&bars[(k - 1) % RING], ((k - 1) / RING) & 1);
            mbar_wait(&bars[k % RING], (k / RING) & 1);
        }
        mbar_wait(&bars[(k + 1) % RING], ((k + 1) / RING) & 1);

        const T* sD = ring + (size_t)((k - 1) % RING) * SLOT_STRIDE;
        const T* sC = ring + (size_t)(k % RING) * SLOT_STRIDE;
        const T* sU = ring + (size_t)((k + 1) % RING) * SLOT_STRIDE;
        const long long zbase = (long long)z * g.plane;
#pragma unroll
        for (int rr = 0; rr < RPT; rr++) {
            const int r = r0 + rr;
            const int y = y0 + r;
            const int q = (colour + y + g.z0 + z) & 1;
            const int x = 2 * i + q;
            if (y <= g.n - 2 && x >= 1 && x <= g.n - 2) {
                MG_CHK_SITE(g, i, y, z);
                const int cc = (r + 1) * BW + il + A;  // smem index of (i, y) in a slot
                const T O = sC[cc - 1 + q], E = sC[cc + q], N = sC[cc - BW], S = sC[cc + BW], D = sD[cc], U = sU[cc];
                __stcs(v_own + zbase + (long long)y * g.hp + i, relax_point<T, FAST>(O, E, N, S, D, U, fcur[rr], c));
            }
        }
        __syncthreads();  // every thread is done with slot k-1: it can be refilled
        if (tid == 0 && z + RING - 1 <= ze) issue(z + RING - 1);
    }
}

inline size_t smem_bytes(size_t esize)
{
    size_t slot = ((size_t)(IT + 2 * (16 / esize)) * BH * esize + 127) / 128 * 128;
    return RING * slot + RING * sizeof(uint64_t);
}

template <typename T>
int launch(cudaStream_t s, const void* map_other, T* v_own, const T* f_own, mg_geom3d g, mg_coef3d c, int colour, int zl_lo,
           int zl_hi)
{
    if (zl_hi <= zl_lo) return 0;
    const int cols = (g.n - 1) / 2;  // updated half-indices per row: 0 .. cols-1
    const int planes = zl_hi - zl_lo;
    int zchunk = 64;
    // enough CTAs for >= 2 waves of 148 SMs x 4 resident CTAs when the grid allows it
    const int tiles_xy = ((cols + IT - 1) / IT) * ((g.n - 2 + YT - 1) / YT);
    while (zchunk > 16 && (long long)tiles_xy * ((planes + zchunk - 1) / zchunk) < 148 * 8) zchunk /= 2;
    dim3 grid((cols + IT - 1) / IT, (g.n - 2 + YT - 1) / YT, (planes + zchunk - 1) / zchunk);
    const size_t smem = smem_bytes(sizeof(T));
    CUtensorMap map;
    memcpy(&map, map_other, sizeof map);
    if (c.fast_den) {
        MG_SET_SMEM_LIMIT((k_relax_colour_tma<T, true>), smem_bytes(sizeof(T)));
        k_relax_colour_tma<T, true><<<grid, NT, smem, s>>>(map, v_own, f_own, g, narrow<T>(c), colour, zl_lo, zl_hi, zchunk);
    } else {
        MG_SET_SMEM_LIMIT((k_relax_colour_tma<T, false>), smem_bytes(sizeof(T)));
        k_relax_colour_tma<T, false><<<grid, NT, smem, s>>>(map, v_own, f_own, g, narrow<T>(c), colour, zl_lo, zl_hi, zchunk);
    }
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace

extern "C" int mgk3d_relax_colour_tma(cudaStream_t s, int dtype, const void* tmap_other, void* v, const void* f, mg_geom3d g,
                                      mg_coef3d c, int colour, int zl_lo, int zl_hi)
{
    const size_t es = dtype == 0 ? 4 : 8;
    char* v_own = (char*)v + (size_t)colour * g.cstride * es;
    const char* f_own = (const char*)f + (size_t)colour * g.cstride * es;
    if (dtype == 0) return launch<float>(s, tmap_other, (float*)v_own, (const float*)f_own, g, c, colour, zl_lo, zl_hi);
    return launch<double>(s, tmap_other, (double*)v_own, (const double*)f_own, g, c, colour, zl_lo, zl_hi);
}
