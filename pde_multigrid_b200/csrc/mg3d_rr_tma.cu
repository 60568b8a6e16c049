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

    // Thread -> residual points, fixed over the planes.  Warp w, lane l.  Four regular passes: colour
    // (p >> 1), row ly = w + 8*(p & 1), half-index hl = l -- every index of these is ONE per-thread base plus
    // a compile-time constant, so the z loop is loads, arithmetic and one store per point.  (The first
    // versions kept per-point offset/flag arrays; under the register cap the compiler re-derived them from
    // threadIdx every plane: 72 instructions per residual point, 4.8e9 warp instructions per launch at
    // 1025^3, issue-bound at 0.53 of the HBM roofline -- profiles/r1_residual_restrict_tma_ncu_full.txt.)
    // The fifth pass is the irregular rest: row 16 of both colours (warps 0, 1) and the 33rd half-index
    // of the 34 (colour, row) pairs (warps 2, 3), with per-thread precomputed indices.
    static_assert(CXT == 32 && CYT == 8 && NT == 256, "thread mapping of the residual stage");
    const int w = tid >> 5, lane = tid & 31;
    const int hp = gf.hp;
    const int T0 = w * W + lane;           // + (8*jj + 1)*W + A - 1 -> element of a v colour sub-tile
    const int T1 = w * 2 * RCOLS + lane;   // + jj*16*RCOLS          -> element of a residual plane
    const int T2 = ((2 * w + 1) * 2) * RCOLS + lane;  // restriction stage: centre row of coarse point (lane, w)
    // Per-thread predicates live in one opaque bit mask (otherwise ptxas re-derives each of them from
    // blockIdx/threadIdx in every plane).  x parity q of a colour-c point in row y of plane z: (c + y + z) & 1.
    enum : unsigned {
        F_A0 = 1u << 0, F_A1 = 1u << 1,  // rows w:     residual defined for the q = 0 / q = 1 point
        F_B0 = 1u << 2, F_B1 = 1u << 3,  // rows w + 8
        F_FA = 1u << 4, F_FB = 1u << 5,  // f needed in rows w / w + 8
        F_L1 = 1u << 6,                  // lane >= 1: the q = 0 point belongs to the residual tile
        F_PW = 1u << 7,                  // y parity of rows w, w + 8 (fy0 is odd)
        F_TACT = 1u << 8, F_TST0 = 1u << 9, F_T0 = 1u << 10, F_T1 = 1u << 11, F_TF = 1u << 12, F_TP = 1u << 13,  // fifth pass
        F_CIN = 1u << 14, F_CBND = 1u << 15, F_CP = 1u << 16  // coarse point: inside the grid, on its boundary, (cx+cy)&1
    };
    unsigned fl = 0;
    int t_own, t_oth, t_r;
    const T *fp, *tfp;  // f of this thread's first point / fifth-pass point in plane zf0 (advanced by one plane per iteration)
    long long c_base;
    {
        const int x0 = 2 * (cx0 - 1 + lane);  // x of a q = 0 point; q = 1: x0 + 1
        const bool xv0 = lane >= 1 && x0 >= 1 && x0 <= n - 2;
        const bool xv1 = x0 + 1 >= 1 && x0 + 1 <= n - 2;
        const int ya = fy0 + w, yb = ya + 8;
        const bool yva = ya >= 1 && ya <= n - 2, yvb = yb >= 1 && yb <= n - 2;
        fl |= (yva && xv0 ? F_A0 : 0) | (yva && xv1 ? F_A1 : 0) | (yvb && xv0 ? F_B0 : 0) | (yvb && xv1 ? F_B1 : 0);
        fl |= (yva && (xv0 || xv1) ? F_FA : 0) | (yvb && (xv0 || xv1) ? F_FB : 0);  // f is only read where a residual is evaluated
        fl |= (lane >= 1 ? F_L1 : 0) | (((w + 1) & 1) ? F_PW : 0);
        const long long zoff = (long long)(zf0 - gf.z0) * gf.plane;
        fp = f + (zoff + (long long)ya * hp + (cx0 - 1 + lane));

        // fifth pass: row 16 of both colours (warps 0, 1) and the 33rd half-index of the 34 (colour, row)
        // pairs (warps 2, 3)
        int col, ly, hl;
        bool act;
        if (w < 2) { col = w; ly = RROWS - 1; hl = lane; act = true; }
        else {
            const int u = (w - 2) * 32 + lane;
            act = w < 4 && u < 2 * RROWS;
            col = (act && u >= RROWS) ? 1 : 0;
            ly = act ? u - col * RROWS : 0;
            hl = 32;
        }
        const int y = fy0 + ly, hi = cx0 - 1 + hl, cc = (ly + 1) * W + hl + A - 1;
        t_own = col * VSUB_STRIDE + cc;
        t_oth = (col ^ 1) * VSUB_STRIDE + cc;
        t_r = ly * 2 * RCOLS + hl;
        const bool tyv = act && y >= 1 && y <= n - 2;
        const bool tx0 = hl >= 1 && 2 * hi >= 1 && 2 * hi <= n - 2, tx1 = 2 * hi + 1 >= 1 && 2 * hi + 1 <= n - 2;
        fl |= (act ? F_TACT : 0) | (act && hl >= 1 ? F_TST0 : 0) | (tyv && tx0 ? F_T0 : 0) | (tyv && tx1 ? F_T1 : 0) |
              (tyv && (tx0 || tx1) ? F_TF : 0) | (((col + y) & 1) ? F_TP : 0);
        tfp = f + (zoff + (long long)col * gf.cstride + (long long)y * hp + hi);

        // the thread's coarse point (restriction stage): cx = cx0 + lane, cy = cy0 + w
        const int cx = cx0 + lane, cy = cy0 + w;
        fl |= (cx < gc.n && cy < gc.n ? F_CIN : 0) | ((cx == 0 || cx == gc.n - 1 || cy == 0 || cy == gc.n - 1) ? F_CBND : 0) |
              (((cx + cy) & 1) ? F_CP : 0);
        c_base = (long long)cy * gc.hp + (cx >> 1);
    }
    asm volatile("" : "+r"(fl));
    const long long f_b = 8ll * hp, f_c1 = gf.cstride;  // f offsets of rows w + 8 / of colour 1, in elements

    const int nzl = gf.nzl;
    // f of plane z (fp/tfp point into that plane) -> dst; addresses of points without a residual are never formed
    auto load_f = [&](int z, T (&dst)[5]) {
        const int zl = z - gf.z0;
        const unsigned m = (zl >= 0 && zl < nzl) ? fl : 0u;
        dst[0] = (m & F_FA) ? __ldg(fp) : T(0);
        dst[1] = (m & F_FB) ? __ldg(fp + f_b) : T(0);
        dst[2] = (m & F_FA) ? __ldg(fp + f_c1) : T(0);
        dst[3] = (m & F_FB) ? __ldg(fp + f_c1 + f_b) : T(0);
        dst[4] = (m & F_TF) ? __ldg(tfp) : T(0);
    };
    T fnext[5];
    load_f(zf0, fnext);

    unsigned rs = 0;  // residual-ring slot of plane z: (z - zf0) % RRING
    for (int z = zf0; z <= zf1; z++) {
        const unsigned k = (unsigned)(z - pbase);  // ring index of v plane z (>= 1)
        T fcur[5];
#pragma unroll
        for (int j = 0; j < 5; j++) fcur[j] = fnext[j];
        fp += gf.plane;
        tfp += gf.plane;
        if (z < zf1) load_f(z + 1, fnext);
        if (z == zf0) {
            mbar_wait(&bars[(k - 1) & (RING - 1)], ((k - 1) / RING) & 1);
            mbar_wait(&bars[k & (RING - 1)], (k / RING) & 1);
        }
        mbar_wait(&bars[(k + 1) & (RING - 1)], ((k + 1) / RING) & 1);

        // ---- residual of fine plane z -> rring[rs] ----
        const T* vD = vring + ((k - 1) & (RING - 1)) * VSLOT + T0;
        const T* vC = vring + (k & (RING - 1)) * VSLOT + T0;
        const T* vU = vring + ((k + 1) & (RING - 1)) * VSLOT + T0;
        T* rz = rring + rs * RSLOT;
        const unsigned vm = (z >= 1 && z <= n - 2) ? fl : 0u;  // residual is zero on the boundary planes
        const int q0 = ((fl / F_PW) ^ z) & 1;  // x parity of the colour-0 points of this thread's rows in this plane
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int col = p >> 1, jj = p & 1;  // compile-time after unrolling
            const int cp = (8 * jj + 1) * W + A - 1;
            const int q = q0 ^ col;
            const T* own = vC + col * VSUB_STRIDE + cp;
            const T* oc = vC + (col ^ 1) * VSUB_STRIDE + cp;
            const T* od = vD + (col ^ 1) * VSUB_STRIDE + cp;
            const T* ou = vU + (col ^ 1) * VSUB_STRIDE + cp;
            T val = T(0);
            if (vm & (jj ? (q ? F_B1 : F_B0) : (q ? F_A1 : F_A0)))
                val = residual_point<T, FAST>(oc[q - 1], oc[q], oc[-W], oc[W], od[0], ou[0], own[0], fcur[p], c, corrected);
            if (q || (fl & F_L1)) rz[T1 + jj * 16 * RCOLS + (q ? 0 : RCOLS - 1)] = val;
        }
        if (w < 4) {  // warp-uniform: warps 4..7 have no fifth-pass point
            const int q = ((fl / F_TP) ^ z) & 1;
            T val = T(0);
            if (vm & (q ? F_T1 : F_T0)) {
                const T *c0 = vC - T0, *d0 = vD - T0, *u0 = vU - T0;
                const int o = t_oth;
                val = residual_point<T, FAST>(c0[o - 1 + q], c0[o + q], c0[o - W], c0[o + W], d0[o], u0[o], c0[t_own], fcur[4], c, corrected);
            }
            if (fl & (q ? F_TACT : F_TST0)) rz[t_r + (q ? 0 : RCOLS - 1)] = val;
        }
        __syncthreads();  // residual plane z complete; v slot of plane z-1 free
        if (tid == 0 && z + RING - 1 <= zf1 + 1) issue(z + RING - 1);

        // ---- after an odd fine plane z = 2*cz+1: restrict planes z-2, z-1, z -> coarse plane cz ----
        if ((z & 1) && z > zf0) {
            const int cz = (z - 1) >> 1, czl = cz - gc.z0;
            if (fl & F_CIN) {
                T out = T(0);  // boundary: injection of the zero boundary residual (N3/MultiGrid3D.cpp:113-119, :705)
                if (!((fl & F_CBND) || cz == 0 || cz == gc.n - 1)) {
                    const T* rm = rring + (rs == 2 ? 0u : rs + 1) * RSLOT + T2;  // plane z-2
                    const T* rc = rring + (rs == 0 ? 2u : rs - 1) * RSLOT + T2;  // plane z-1
                    const T* rp = rz + T2;
                    // centre lx = 2*lane+1 (odd: parity array 1 at lane); dx = -1/+1 -> even lx: array 0 at lane / lane+1
                    out = restrict_point<T>([&](int dx, int dy, int dz) {
                        const T* pl = dz < 0 ? rm : (dz == 0 ? rc : rp);
                        return dx == 0 ? pl[dy * 2 * RCOLS + RCOLS] : pl[dy * 2 * RCOLS + (dx > 0)];
                    });
                }
                const long long ci = ((((fl / F_CP) ^ cz) & 1) ? gc.cstride : 0ll) + (long long)czl * gc.plane + c_base;
                cf[ci] = out;
                cv[ci] = T(0);  // setToValue(coarse->h_v, 0, true), N3/MultiGrid3D.cpp:634
            }
            __syncthreads();  // the 3-slot residual ring: plane z-2 is overwritten by the next iteration
        }
        rs = rs == RRING - 1 ? 0 : rs + 1;
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
