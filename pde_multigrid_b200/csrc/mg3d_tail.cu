// mg3d_tail.cu -- the coarse tail of the 3D V-cycle in ONE launch.
//
// Levels with n <= 17 (17^3, 9^3, 5^3, 3^3: 5 794 points in all) are pure launch latency when every operator
// is its own kernel: ~10 launches per level and cycle, ~0.08 ms per level against microseconds of work; at
// 257^3, and on 8-GPU slabs where every GPU runs the agglomerated levels, that is >10 % of the cycle.  Here a
// single persistent CTA runs the whole recursion of MultiGrid3D::VCycle (N3/MultiGrid3D.cpp:623-647) from the
// first such level down to the coarsest and back, with every level's v and f resident in shared memory and
// __syncthreads() between colours and operators.  The per-point formulas are the shared device functions of
// mg3d_device.cuh (relax_point, residual_point, restrict_point, interp_point), so results stay bit-identical.
// All levels' v and f are written back to their (colour-split) HBM arrays at the end: the reference keeps
// them observable (grids3D[l]->h_v / h_f) and the parity tests compare them.
#include "mg3d_device.cuh"

using namespace mgx;
using namespace mg3;

namespace {

constexpr int NT = 1024;
constexpr int MAXL = MGK3D_TAIL_MAX_LEVELS;

template <typename T>
struct TailArgs {
    int nlev, v1, v2, corrected;
    int n[MAXL];
    T* v[MAXL];
    T* f[MAXL];
    mg_geom3d g[MAXL];
    Coef3<T> c[MAXL];
};

// A thread's interior points are the same in every half-sweep of a call: their array index and colour are worked out once
// (the integer divisions), the sweeps then only test a colour bit.  At most MAXP points per thread (15^3 interior points of
// the 17^3 level over 1024 threads).
constexpr int MAXP = ((MGK3D_TAIL_N - 2) * (MGK3D_TAIL_N - 2) * (MGK3D_TAIL_N - 2) + NT - 1) / NT;

template <typename T, bool FAST_DEN>
__device__ void tail_relax(T* v, const T* f, int n, const Coef3<T>& c, int ncycles)
{
    if (ncycles <= 0) return;  // (uniform over the block)
    const int ni = n - 2, tot = ni * ni * ni, n2 = n * n;
    int pi[MAXP];
    unsigned valid = 0, odd = 0;
#pragma unroll
    for (int j = 0; j < MAXP; j++) {
        const int idx = threadIdx.x + j * NT;
        pi[j] = 0;
        if (idx < tot) {
            const int x = 1 + idx % ni, y = 1 + (idx / ni) % ni, z = 1 + idx / (ni * ni);
            pi[j] = (z * n + y) * n + x;
            valid |= 1u << j;
            odd |= (unsigned)((x + y + z) & 1) << j;
        }
    }
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++) {
            const unsigned mine = valid & (colour ? odd : ~odd);
#pragma unroll
            for (int j = 0; j < MAXP; j++)
                if ((mine >> j) & 1u) {
                    const int i = pi[j];
                    v[i] = relax_point<T, FAST_DEN>(v[i - 1], v[i + 1], v[i - n], v[i + n], v[i - n2], v[i + n2], f[i], c);
                }
            __syncthreads();
        }
}

template <typename T, bool FAST_DEN, bool FAST_H>
__global__ void __launch_bounds__(NT) k_vcycle_tail(TailArgs<T> a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* base = reinterpret_cast<T*>(smem_raw);
    T *sv[MAXL], *sf[MAXL];
    {
        T* p = base;
        for (int l = 0; l < a.nlev; l++) {
            const int n3 = a.n[l] * a.n[l] * a.n[l];
            sv[l] = p; p += n3;
            sf[l] = p; p += n3;
        }
    }
    T* sr = sv[a.nlev - 1] + 2 * a.n[a.nlev - 1] * a.n[a.nlev - 1] * a.n[a.nlev - 1];  // residual of the current fine level
    const int tid = threadIdx.x;

    {   // load the entry level (coarser levels are produced below)
        const int n = a.n[0];
        for (int idx = tid; idx < n * n * n; idx += NT) {
            const int x = idx % n, y = (idx / n) % n, z = idx / (n * n);
            const long long o = off3(a.g[0], x, y, z);
            sv[0][idx] = a.v[0][o];
            sf[0][idx] = a.f[0][o];
        }
    }
    __syncthreads();

    const int last = a.nlev - 1;
    for (int l = 0; l < last; l++) {
        const int n = a.n[l], cn = a.n[l + 1];
        tail_relax<T, FAST_DEN>(sv[l], sf[l], n, a.c[l], a.v1);
        // CalculateResidual (boundary zero, N3/MultiGrid3D.cpp:704-705)
        for (int idx = tid; idx < n * n * n; idx += NT) {
            const int x = idx % n, y = (idx / n) % n, z = idx / (n * n);
            T r = T(0);
            if (x > 0 && x < n - 1 && y > 0 && y < n - 1 && z > 0 && z < n - 1) {
                const T* v = sv[l] + idx;
                r = residual_point<T, FAST_H>(v[-1], v[1], v[-n], v[n], v[-n * n], v[n * n], v[0], sf[l][idx], a.c[l], a.corrected);
            }
            sr[idx] = r;
        }
        __syncthreads();
        // Restrict -> coarse f (boundary: injection), coarse v = 0 everywhere (N3/MultiGrid3D.cpp:632-634)
        for (int idx = tid; idx < cn * cn * cn; idx += NT) {
            const int cx = idx % cn, cy = (idx / cn) % cn, cz = idx / (cn * cn);
            const T* p = sr + ((2 * cz) * n + 2 * cy) * n + 2 * cx;
            T out;
            if (cx == 0 || cx == cn - 1 || cy == 0 || cy == cn - 1 || cz == 0 || cz == cn - 1) out = p[0];
            else out = restrict_point<T>([&](int dx, int dy, int dz) { return p[dx + dy * n + dz * n * n]; });
            sf[l + 1][idx] = out;
            sv[l + 1][idx] = T(0);
        }
        __syncthreads();
    }
    tail_relax<T, FAST_DEN>(sv[last], sf[last], a.n[last], a.c[last], a.v1);
    tail_relax<T, FAST_DEN>(sv[last], sf[last], a.n[last], a.c[last], a.v2);
    for (int l = last - 1; l >= 0; l--) {
        const int n = a.n[l], cn = a.n[l + 1], ni = n - 2;
        // Interpolate + ApplyCorrection (N3/MultiGrid3D.cpp:638-642), interior of the fine level
        for (int idx = tid; idx < ni * ni * ni; idx += NT) {
            const int x = 1 + idx % ni, y = 1 + (idx / ni) % ni, z = 1 + idx / (ni * ni);
            const T* cptr = sv[l + 1] + ((z >> 1) * cn + (y >> 1)) * cn + (x >> 1);
            const T e = interp_point<T>([&](int dx, int dy, int dz) { return cptr[dx + dy * cn + dz * cn * cn]; }, x & 1, y & 1, z & 1);
            T* p = sv[l] + (z * n + y) * n + x;
            *p = add(*p, e);
        }
        __syncthreads();
        tail_relax<T, FAST_DEN>(sv[l], sf[l], n, a.c[l], a.v2);
    }

    // write everything back: v of every level, f of the levels below the entry level
    for (int l = 0; l < a.nlev; l++) {
        const int n = a.n[l];
        for (int idx = tid; idx < n * n * n; idx += NT) {
            const int x = idx % n, y = (idx / n) % n, z = idx / (n * n);
            const long long o = off3(a.g[l], x, y, z);
            a.v[l][o] = sv[l][idx];
            if (l > 0) a.f[l][o] = sf[l][idx];
        }
    }
}

template <typename T>
int launch(cudaStream_t s, int nlev, const mg_geom3d* g, const mg_coef3d* c, void* const* v, void* const* f, int v1, int v2, int corrected)
{
    TailArgs<T> a;
    a.nlev = nlev; a.v1 = v1; a.v2 = v2; a.corrected = corrected;
    size_t elems = 0;
    bool fast_den = true, fast_h = true;
    for (int l = 0; l < nlev; l++) {
        a.n[l] = g[l].n;
        a.v[l] = (T*)v[l];
        a.f[l] = (T*)f[l];
        a.g[l] = g[l];
        a.c[l] = narrow<T>(c[l]);
        elems += 2 * (size_t)g[l].n * g[l].n * g[l].n;
        fast_den = fast_den && c[l].fast_den;
        fast_h = fast_h && c[l].fast_h;
    }
    elems += (size_t)g[0].n * g[0].n * g[0].n;  // residual scratch of the largest level
    const size_t smem = elems * sizeof(T);
#define MG_TAIL_LAUNCH(FD, FH)                                                                                          \
    do {                                                                                                                \
        MG_SET_SMEM_LIMIT((k_vcycle_tail<T, FD, FH>), 200 * 1024);                                                      \
        k_vcycle_tail<T, FD, FH><<<1, NT, smem, s>>>(a);                                                                \
    } while (0)
    if (fast_den && fast_h) MG_TAIL_LAUNCH(true, true);
    else MG_TAIL_LAUNCH(false, false);
#undef MG_TAIL_LAUNCH
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace

/* V(v1,v2) on levels 0..nlev-1 of the given sub-hierarchy (g[0].n <= MGK3D_TAIL_N, all whole-level, z0 = 0) */
extern "C" int mgk3d_vcycle_tail(cudaStream_t s, int dtype, int nlev, const mg_geom3d* g, const mg_coef3d* c, void* const* v,
                                 void* const* f, int v1, int v2, int corrected)
{
    if (nlev < 1 || nlev > MAXL || g[0].n > MGK3D_TAIL_N) return -1;
    if (dtype == 0) return launch<float>(s, nlev, g, c, v, f, v1, v2, corrected);
    return launch<double>(s, nlev, g, c, v, f, v1, v2, corrected);
}
