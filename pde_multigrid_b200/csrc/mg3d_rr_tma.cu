// mg3d_rr_tma.cu -- fused CalculateResidual + Restrict (+ zero of the coarse v) for the large 3D levels,
// with TMA-staged z-marching shared-memory tiles.  Replaces, in one pass over HBM,
//     residual = CalculateResidual(fine)                       N3/MultiGrid3D.cpp:678-730
//     Restrict(residual, ..., coarse->h_f, ...)                N3/MultiGrid3D.cpp:50-184
//     setToValue(coarse->h_v, ..., 0, true)                    N3/MultiGrid3D.cpp:634
// The fine residual only ever exists in shared memory.
//
// A CTA owns a (CXT x CYT) column of coarse points and marches along z one FINE plane at a time:
//   * the two colour sub-tiles of fine v plane z+1 arrive by TMA (two cp.async.bulk.tensor.3d on one
//     mbarrier) into a 4-slot ring while plane z is being processed; halo / out-of-domain elements are
//     zero-filled by the hardware;
//   * the residual of fine plane z is evaluated from the ring (7 conflict-free LDS per point) into a
//     3-slot ring of residual planes stored parity-split in x, so that the 27 reads of the restriction
//     are unit-stride as well; f is prefetched one plane ahead into registers;
//   * after every odd fine plane 2k+1 the coarse plane k is produced from residual planes 2k-1, 2k, 2k+1
//     with the reference's exact grouping of the 27 weights, and written together with coarse v = 0.
// HBM traffic per fine point: v and f once (+ 19/16 x 36/32 halo re-reads that hit L2) and 2/8 coarse
// writes: the algorithmic 2*B*N_l + 2*B*N_{l+1} of SURVEY.md 8(d).
#include "mg3d_device.cuh"
#include "mg_tma.cuh"

using namespace mgx;
using namespace mg3;
using namespace mgtma;

namespace {

constexpr int CXT = MGK3D_RR_CXT;       // coarse points per tile in x
constexpr int CYT = MGK3D_RR_CYT;       // coarse points per tile in y
constexpr int VROWS = MGK3D_RR_BOX_Y;   // fine rows of a v tile: 2*CYT + 3
constexpr int RROWS = 2 * CYT + 1;      // fine rows of a residual plane tile
constexpr int RCOLS = CXT + 2;          // per-parity columns of a residual row (33 used, even pitch)
constexpr int HPR = CXT + 1;            // half-indices per residual row and colour (33)
constexpr int RING = 4;   // v planes: z-1, z, z+1 live + one in flight
constexpr int RRING = 3;  // residual planes: 2k-1, 2k, 2k+1
constexpr int NT = 256;
constexpr int NPT = (2 * RROWS * HPR + NT - 1) / NT;  // residual points per thread and plane (5)

template <typename T> struct VBox {
    static constexpr int A = 16 / sizeof(T);                      // TMA inner-coordinate alignment in elements
    static constexpr int W = (CXT + A + 1 + A - 1) / A * A;       // 36 doubles / 40 floats
};

template <typename T, bool FAST>
__global__ void __launch_bounds__(NT, 2)
k_residual_restrict_tma(const __grid_constant__ CUtensorMap map_c0, const __grid_constant__ CUtensorMap map_c1,
                        const T* __restrict__ f, mg_geom3d gf, Coef3<T> c, int corrected, T* __restrict__ cf,
                        T* __restrict__ cv, mg_geom3d gc, int czl_lo, int czl_hi, int zchunk)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int A = VBox<T>::A, W = VBox<T>::W;
    constexpr int VSUB = VROWS * W;                                // one colour sub-tile
    constexpr uint32_t VSUB_BYTES = VSUB * sizeof(T);
    constexpr int VSUB_STRIDE = (VSUB_BYTES + 127) / 128 * 128 / sizeof(T);
    constexpr int VSLOT = 2 * VSUB_STRIDE;                         // both colours of one fine plane
    constexpr int RSLOT = RROWS * 2 * RCOLS;                       // residual plane: [row][x parity][col]
    T* vring = reinterpret_cast<T*>(smem_raw);
    T* rring = vring + (size_t)RING * VSLOT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(rring + (size_t)RRING * RSLOT);

    const int tid = threadIdx.x;
    const int cx0 = blockIdx.x * CXT, cy0 = blockIdx.y * CYT;
    const int czl0 = czl_lo + blockIdx.z * zchunk;                 // coarse local planes [czl0, czl1)
    const int czl1 = min(czl0 + zchunk, czl_hi);
    const int n = gf.n;
    // fine GLOBAL planes whose residual is needed: 2*cz-1 .. 2*cz+1 for every coarse plane of the chunk
    const int zf0 = 2 * (gc.z0 + czl0) - 1, zf1 = 2 * (gc.z0 + czl1 - 1) + 1;
    const int fy0 = 2 * cy0 - 1;                                   // first fine row of the residual tile
    const int hi_org = cx0 - A;                                    // half-index of column 0 of a v sub-tile

    if (tid == 0) {
        prefetch_tensormap(&map_c0);
        prefetch_tensormap(&map_c1);
        for (int s = 0; s < RING; s++) mbar_init(&bars[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // v plane p (global) lives in slot (p - pbase) % RING; planes pbase = zf0-1 .. zf1+1 are streamed
    const int pbase = zf0 - 1;
    auto issue = [&](int p) {
        const int s = (p - pbase) % RING;
        T* dst = vring + (size_t)s * VSLOT;
        mbar_arrive_expect_tx(&bars[s], 2 * VSUB_BYTES);
        tma_load_3d(dst, &map_c0, &bars[s], hi_org, fy0 - 1, p - gf.z0);
        tma_load_3d(dst + VSUB_STRIDE, &map_c1, &bars[s], hi_org, fy0 - 1, p - gf.z0);
    };
    if (tid == 0)
        for (int p = pbase; p <= min(pbase + RING - 1, zf1 + 1); p++) issue(p);

    // The thread's residual points: slot s = tid + j*NT -> (colour, row ly, half-index hl), fixed over planes.
    // Everything that does not depend on z is folded into a few per-point constants here so that the
    // z loop is loads, arithmetic and one store per point (the first version of this kernel spent
    // ~128 instructions per fine point, mostly on index arithmetic, and was issue-bound: profiles/).
    int own_off[NPT], oth_off[NPT], r_even[NPT], flags[NPT];  // flags: 1 parity of (colour+y), 2 valid for q=0, 4 valid for q=1
    const T* fptr[NPT];
#pragma unroll
    for (int j = 0; j < NPT; j++) {
        // a warp covers half-indices 0..31 of ONE (colour, row): every shared-memory access of the residual
        // stage is unit-stride and conflict-free; the 33rd half-index of the 34 (colour, row) pairs is a
        // small tail handled by 34 threads of the last pass
        const int s = tid + j * NT;
        const bool main_part = s < 2 * RROWS * 32;
        const bool active = s < 2 * RROWS * HPR;
        const int cr = main_part ? (s >> 5) : (active ? s - 2 * RROWS * 32 : 0);
        const int hl = main_part ? (s & 31) : (active ? 32 : 0);
        const int col = cr / RROWS;
        const int ly = cr - col * RROWS;
        const int y = fy0 + ly, hi = cx0 - 1 + hl;
        const int cc = (ly + 1) * W + hl + A - 1;
        own_off[j] = col * VSUB_STRIDE + cc;
        oth_off[j] = (col ^ 1) * VSUB_STRIDE + cc;
        r_even[j] = (ly * 2) * RCOLS + hl;  // q = 1: even lx = 2*hl -> parity array 0; q = 0: odd lx -> array 1 at hl-1
        const bool yv = active && y >= 1 && y <= n - 2;
        const int x0 = 2 * hi;  // q = 0; q = 1 -> x0 + 1
        flags[j] = ((col + y) & 1) | ((yv && hl >= 1 && x0 >= 1 && x0 <= n - 2) ? 2 : 0) | ((yv && x0 + 1 >= 1 && x0 + 1 <= n - 2) ? 4 : 0) |
                   (active ? 8 : 0) | ((active && hl >= 1) ? 16 : 0);
        const bool fok = active && y >= 0 && y < n && hi >= 0 && hi < gf.hp;
        fptr[j] = fok ? f + (long long)col * gf.cstride + (long long)y * gf.hp + hi : nullptr;
    }
    const int nzl = gf.nzl;
    auto load_f = [&](int z, T (&dst)[NPT]) {
        const int zl = z - gf.z0;
        const bool zok = zl >= 0 && zl < nzl;
        const long long po = (long long)zl * gf.plane;
#pragma unroll
        for (int j = 0; j < NPT; j++) dst[j] = (zok && fptr[j]) ? __ldg(fptr[j] + po) : T(0);
    };
    T fnext[NPT];
    load_f(zf0, fnext);

    for (int z = zf0; z <= zf1; z++) {
        const int k = z - pbase;  // ring index of v plane z
        T fcur[NPT];
#pragma unroll
        for (int j = 0; j < NPT; j++) fcur[j] = fnext[j];
        if (z < zf1) load_f(z + 1, fnext);
        if (z == zf0) {
            mbar_wait(&bars[(k - 1) % RING], ((k - 1) / RING) & 1);
            mbar_wait(&bars[k % RING], (k / RING) & 1);
        }
        mbar_wait(&bars[(k + 1) % RING], ((k + 1) / RING) & 1);

        // ---- residual of fine plane z -> rring[(z - zf0) % RRING] ----
        const T* vD = vring + (size_t)((k - 1) % RING) * VSLOT;
        const T* vC = vring + (size_t)(k % RING) * VSLOT;
        const T* vU = vring + (size_t)((k + 1) % RING) * VSLOT;
        T* rz = rring + (size_t)((z - zf0) % RRING) * RSLOT;
        const bool zin = z >= 1 && z <= n - 2;
        const int zpar = z & 1;
#pragma unroll
        for (int j = 0; j < NPT; j++) {
            const int q = (flags[j] ^ zpar) & 1;  // x parity of this colour in this row of this plane
            const bool store = q ? (flags[j] & 8) : (flags[j] & 16);
            const bool valid = zin && (flags[j] & (q ? 4 : 2));
            T val = T(0);
            if (valid) {
                const int o = oth_off[j];
                val = residual_point<T, FAST>(vC[o - 1 + q], vC[o + q], vC[o - W], vC[o + W], vD[o], vU[o], vC[own_off[j]], fcur[j], c,
                                              corrected);
            }
            if (store) rz[r_even[j] + (q ? 0 : RCOLS - 1)] = val;
        }
        __syncthreads();  // residual plane z complete; v slot of plane z-1 free
        if (tid == 0 && z + RING - 1 <= zf1 + 1) issue(z + RING - 1);

        // ---- after an odd fine plane z = 2*cz+1: restrict planes z-2, z-1, z -> coarse plane cz ----
        if ((z & 1) && z > zf0) {
            const int cz = (z - 1) >> 1, czl = cz - gc.z0;
            const int tx = tid & (CXT - 1), ty = tid / CXT;
            const int cx = cx0 + tx, cy = cy0 + ty;
            if (cx < gc.n && cy < gc.n) {
                T out = T(0);  // boundary: injection of the zero boundary residual (N3/MultiGrid3D.cpp:113-119, :705)
                if (!(cx == 0 || cx == gc.n - 1 || cy == 0 || cy == gc.n - 1 || cz == 0 || cz == gc.n - 1)) {
                    const T* rm = rring + (size_t)((z - 2 - zf0) % RRING) * RSLOT + ((2 * ty + 1) * 2) * RCOLS + tx;
                    const T* rc = rring + (size_t)((z - 1 - zf0) % RRING) * RSLOT + ((2 * ty + 1) * 2) * RCOLS + tx;
                    const T* rp = rz + ((2 * ty + 1) * 2) * RCOLS + tx;
                    // centre lx = 2*tx+1 (odd: parity array 1 at tx); dx = -1/+1 -> even lx: array 0 at tx / tx+1
                    out = restrict_point<T>([&](int dx, int dy, int dz) {
                        const T* pl = dz < 0 ? rm : (dz == 0 ? rc : rp);
                        return dx == 0 ? pl[dy * 2 * RCOLS + RCOLS] : pl[dy * 2 * RCOLS + (dx > 0)];
                    });
                }
                const long long ci = off3(gc, cx, cy, czl);
                cf[ci] = out;
                cv[ci] = T(0);  // setToValue(coarse->h_v, 0, true), N3/MultiGrid3D.cpp:634
            }
            __syncthreads();  // the 3-slot residual ring: plane z-2 is overwritten by the next iteration
        }
    }
}

template <typename T>
size_t smem_bytes_t()
{
    constexpr size_t vsub = ((size_t)VROWS * VBox<T>::W * sizeof(T) + 127) / 128 * 128;
    return RING * 2 * vsub + (size_t)RRING * RROWS * 2 * RCOLS * sizeof(T) + RING * sizeof(uint64_t);
}

template <typename T, bool FAST>
int launch_k(cudaStream_t s, const CUtensorMap& m0, const CUtensorMap& m1, const T* f, mg_geom3d gf, mg_coef3d c, int corrected, T* cf,
             T* cv, mg_geom3d gc, int czl_lo, int czl_hi, dim3 grid, int zchunk)
{
    const size_t smem = smem_bytes_t<T>();
    static bool attr = (cudaFuncSetAttribute(k_residual_restrict_tma<T, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_t<T>()), true);
    (void)attr;
    k_residual_restrict_tma<T, FAST><<<grid, NT, smem, s>>>(m0, m1, f, gf, narrow<T>(c), corrected, cf, cv, gc, czl_lo, czl_hi, zchunk);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

template <typename T>
int launch(cudaStream_t s, const void* tmap_c0, const void* tmap_c1, const T* f, mg_geom3d gf, mg_coef3d c, int corrected, T* cf, T* cv,
           mg_geom3d gc, int czl_lo, int czl_hi)
{
    if (czl_hi <= czl_lo) return 0;
    static_assert(NT == CXT * CYT, "one coarse point per thread");
    const int planes = czl_hi - czl_lo;
    const int tiles_xy = ((gc.n + CXT - 1) / CXT) * ((gc.n + CYT - 1) / CYT);
    int zchunk = 32;
    while (zchunk > 8 && (long long)tiles_xy * ((planes + zchunk - 1) / zchunk) < 148 * 4) zchunk /= 2;
    dim3 grid((gc.n + CXT - 1) / CXT, (gc.n + CYT - 1) / CYT, (planes + zchunk - 1) / zchunk);
    CUtensorMap m0, m1;
    memcpy(&m0, tmap_c0, sizeof m0);
    memcpy(&m1, tmap_c1, sizeof m1);
    if (c.fast_h) return launch_k<T, true>(s, m0, m1, f, gf, c, corrected, cf, cv, gc, czl_lo, czl_hi, grid, zchunk);
    return launch_k<T, false>(s, m0, m1, f, gf, c, corrected, cf, cv, gc, czl_lo, czl_hi, grid, zchunk);
}

}  // namespace

extern "C" int mgk3d_residual_restrict_tma(cudaStream_t s, int dtype, const void* tmap_v_c0, const void* tmap_v_c1, const void* f,
                                           mg_geom3d gf, mg_coef3d c, int corrected, void* coarse_f, void* coarse_v, mg_geom3d gc,
                                           int czl_lo, int czl_hi)
{
    if (dtype == 0)
        return launch<float>(s, tmap_v_c0, tmap_v_c1, (const float*)f, gf, c, corrected, (float*)coarse_f, (float*)coarse_v, gc, czl_lo, czl_hi);
    return launch<double>(s, tmap_v_c0, tmap_v_c1, (const double*)f, gf, c, corrected, (double*)coarse_f, (double*)coarse_v, gc, czl_lo, czl_hi);
}
