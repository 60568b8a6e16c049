/*
 * mg3d_host.c -- C host driver of the 3D Poisson multigrid: hierarchy, V-cycle, FMG, field I/O.
 *
 * Mirrors the control flow of the reference class MultiGrid3D
 * (NOCUDA_TESI/POISSON_3D(TESI)/MultiGrid3D.cpp: InitGrids :19-47, VCycle :623-647,
 * FullMultiGridVCycle :569-585) over the sm_100a kernels of mg3d_kernels.cu.  Host code is C; all
 * device work goes through the launchers declared in mg_launch.h.  No CPU compute path exists here:
 * the host only computes per-level scalars (h, h^2 products, sin tables for InitF).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mg_host_common.h"
#include "mg_profile.h"

typedef struct {
    mg_geom3d g;
    mg_coef3d c;
    void* v;
    void* f;
    double h[3];    /* h_x, h_y, h_z (values of the level's dtype) */
    int own_lo;     /* local plane range [own_lo, own_hi) owned by this rank */
    int own_hi;
    int has_tma;    /* tensor maps of the two colour arrays of v are valid */
    unsigned char tmap_v[2][128] __attribute__((aligned(64)));  /* smoother boxes */
    unsigned char tmap_rr[2][128] __attribute__((aligned(64))); /* residual+restrict boxes */
} mg_level3d;

struct mg3d_s {
    int dtype, mode, nlevels;
    int rank, nranks;
    int smoother, sweeps_per_pass;
    double range[6];
    cudaStream_t stream;
    mg_level3d* lv;
    void* arena;
    double* d_scratch; /* 2*MGK_NORM_BLOCKS partials + 2 outputs */
    double* d_tables;  /* 3*n0 doubles: sin tables of InitF */
    double* h_out2;    /* pinned */
    void* staging;     /* dense device staging buffer for large host<->device field copies */
    size_t staging_bytes;
    long long launches;
    mg_prof prof;
};

#define PROF_BEGIN(mg, level, op) mg_prof_begin(&(mg)->prof, (mg)->stream, (level), (op), (mg)->launches)
#define PROF_END(mg) mg_prof_end(&(mg)->prof, (mg)->stream, (mg)->launches)

/* h = range/(real)(n-1) (N3/Grid3D.cpp:31-45) and the products of N3/MultiGrid3D.cpp:498-500,532,
   computed in the level's own precision exactly like the reference does on the host */
static void level_coefs(int dtype, int n, const double* range, double h[3], mg_coef3d* c)
{
    if (dtype == MG_F32) {
        float xr = (float)range[1] - (float)range[0];
        float yr = (float)range[3] - (float)range[2];
        float zr = (float)range[5] - (float)range[4];
        float hx = xr / (float)(n - 1), hy = yr / (float)(n - 1), hz = zr / (float)(n - 1);
        float hx2 = hx * hx, hy2 = hy * hy, hz2 = hz * hz;
        float cx = hy2 * hz2, cy = hx2 * hz2, cz = hx2 * hy2;
        float den = 2 * (cx + cy + cz);
        float rden = 1.0f / den;
        h[0] = hx; h[1] = hy; h[2] = hz;
        c->hx2 = hx2; c->hy2 = hy2; c->hz2 = hz2;
        c->cx = cx; c->cy = cy; c->cz = cz;
        c->den = den; c->rden = rden;
        c->ihx2 = 1.0f / hx2; c->ihy2 = 1.0f / hy2; c->ihz2 = 1.0f / hz2;
    } else {
        double xr = range[1] - range[0], yr = range[3] - range[2], zr = range[5] - range[4];
        double hx = xr / (double)(n - 1), hy = yr / (double)(n - 1), hz = zr / (double)(n - 1);
        double hx2 = hx * hx, hy2 = hy * hy, hz2 = hz * hz;
        double cx = hy2 * hz2, cy = hx2 * hz2, cz = hx2 * hy2;
        double den = 2 * (cx + cy + cz);
        h[0] = hx; h[1] = hy; h[2] = hz;
        c->hx2 = hx2; c->hy2 = hy2; c->hz2 = hz2;
        c->cx = cx; c->cy = cy; c->cz = cz;
        c->den = den; c->rden = 1.0 / den;
        c->ihx2 = 1.0 / hx2; c->ihy2 = 1.0 / hy2; c->ihz2 = 1.0 / hz2;
    }
    /* exact-arithmetic shortcuts (mg_exact.cuh): a power of two has frexp mantissa 0.5; den = 6*2^e has 0.75 */
    int e;
    c->fast_h = frexp(c->hx2, &e) == 0.5 && frexp(c->hy2, &e) == 0.5 && frexp(c->hz2, &e) == 0.5;
    c->fast_den = c->fast_h && frexp(c->den, &e) == 0.75;
    if (getenv("MG_B200_IEEE_DIV")) c->fast_h = c->fast_den = 0;
}

static void set_geom(mg_geom3d* g, int n, int dtype, int z0, int nzl)
{
    g->n = n;
    g->hp = mg_pitch((n + 1) / 2, dtype);
    g->plane = (long long)g->hp * n;
    g->cstride = g->plane * nzl;
    g->z0 = z0;
    g->nzl = nzl;
}

static size_t field_bytes(const mg_level3d* L, int dtype)
{
    return mg_align256(2 * (size_t)L->g.cstride * mg_esize(dtype));
}

static int check_level(const mg3d_t* mg, int level)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    if (level < 0 || level >= mg->nlevels) return mg_fail(MG_ERR_ARG, "level %d out of range [0,%d)", level, mg->nlevels);
    return MG_OK;
}

static void* field_ptr(mg_level3d* L, int field) { return field == MG_FIELD_V ? L->v : L->f; }

int mg3d_create(mg3d_t** out, const int sz[3], const double range[6], int dtype, int residual_mode)
{
    if (!out || !sz || !range) return mg_fail(MG_ERR_ARG, "null argument");
    *out = NULL;
    /* N3/Grid3D.cpp:10-29 asserts */
    if (sz[0] != sz[1] || sz[0] != sz[2]) return mg_fail(MG_ERR_ARG, "sizeX == sizeY == sizeZ required (got %d,%d,%d)", sz[0], sz[1], sz[2]);
    const int n = sz[0];
    if (n < 3 || ((n - 1) & (n - 2)) != 0) return mg_fail(MG_ERR_ARG, "size must be 2^k+1 with k >= 1 (got %d)", n);
    if (!(range[1] > range[0]) || !(range[3] > range[2]) || !(range[5] > range[4])) return mg_fail(MG_ERR_ARG, "range must satisfy b > a on every axis");
    if (dtype != MG_F32 && dtype != MG_F64) return mg_fail(MG_ERR_ARG, "dtype must be MG_F32 or MG_F64");
    if (residual_mode != MG_REF_COMPAT && residual_mode != MG_CORRECTED) return mg_fail(MG_ERR_ARG, "bad residual_mode");
    int st = mg_require_device();
    if (st) return st;

    mg3d_t* mg = (mg3d_t*)calloc(1, sizeof *mg);
    if (!mg) return mg_fail(MG_ERR_NOMEM, "host allocation failed");
    mg->dtype = dtype;
    mg->mode = residual_mode;
    mg->rank = 0;
    mg->nranks = 1;
    mg->smoother = MG_SMOOTHER_AUTO;
    mg->sweeps_per_pass = 1;
    memcpy(mg->range, range, sizeof mg->range);
    mg->nlevels = mg_num_levels_for(n);
    mg->lv = (mg_level3d*)calloc((size_t)mg->nlevels, sizeof(mg_level3d));
    if (!mg->lv) { free(mg); return mg_fail(MG_ERR_NOMEM, "host allocation failed"); }

    size_t total = 0;
    int nl = n;
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level3d* L = &mg->lv[l];
        set_geom(&L->g, nl, dtype, 0, nl);
        L->own_lo = 0;
        L->own_hi = nl;
        level_coefs(dtype, nl, range, L->h, &L->c);
        total += 2 * field_bytes(L, dtype);
        nl = (nl - 1) / 2 + 1; /* N3/MultiGrid3D.cpp:40-42 */
    }
    cudaError_t e = cudaMalloc(&mg->arena, total);
    if (e != cudaSuccess) {
        free(mg->lv); free(mg);
        return mg_fail(e == cudaErrorMemoryAllocation ? MG_ERR_NOMEM : MG_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", total, cudaGetErrorString(e));
    }
    char* p = (char*)mg->arena;
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level3d* L = &mg->lv[l];
        L->v = p; p += field_bytes(L, dtype);
        L->f = p; p += field_bytes(L, dtype);
    }
    if (cudaStreamCreateWithFlags(&mg->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&mg->d_scratch, (2 * MGK_NORM_BLOCKS + 2) * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&mg->d_tables, 3 * (size_t)n * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&mg->h_out2, 2 * sizeof(double)) != cudaSuccess) {
        int code = mg_fail(MG_ERR_CUDA, "stream/scratch setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        mg3d_destroy(mg);
        return code;
    }
    /* TMA tensor maps of v for the levels large enough to fill the z-marching tiles */
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level3d* L = &mg->lv[l];
        if ((L->g.n - 1) / 2 < MGK3D_TMA_IT || getenv("MG_B200_NO_TMA")) continue;
        for (int col = 0; col < 2; col++) {
            st = mg_tma_make_colour_map(L->tmap_v[col], dtype, (char*)L->v + (size_t)col * L->g.cstride * mg_esize(dtype), &L->g,
                                        MGK3D_TMA_BOX_I(mg_esize(dtype)), MGK3D_TMA_BOX_Y);
            if (!st) st = mg_tma_make_colour_map(L->tmap_rr[col], dtype, (char*)L->v + (size_t)col * L->g.cstride * mg_esize(dtype), &L->g,
                                                 MGK3D_RR_BOX_I(mg_esize(dtype)), MGK3D_RR_BOX_Y);
            if (st) { mg3d_destroy(mg); return st; }
        }
        L->has_tma = 1;
    }
    /* pad elements of the layout are never used by a kernel, but keep them defined */
    MG_CUDA(cudaMemsetAsync(mg->arena, 0, total, mg->stream));
    st = mg3d_init_problem(mg);
    if (st) { mg3d_destroy(mg); return st; }
    *out = mg;
    return MG_OK;
}

int mg3d_destroy(mg3d_t* mg)
{
    if (!mg) return MG_OK;
    if (mg->stream) { cudaStreamSynchronize(mg->stream); cudaStreamDestroy(mg->stream); }
    if (mg->arena) cudaFree(mg->arena);
    if (mg->d_scratch) cudaFree(mg->d_scratch);
    if (mg->d_tables) cudaFree(mg->d_tables);
    if (mg->h_out2) cudaFreeHost(mg->h_out2);
    if (mg->staging) cudaFree(mg->staging);
    mg_prof_free(&mg->prof);
    free(mg->lv);
    free(mg);
    return MG_OK;
}

int mg3d_num_levels(const mg3d_t* mg) { return mg ? mg->nlevels : 0; }
int mg3d_level_size(const mg3d_t* mg, int level) { return (mg && level >= 0 && level < mg->nlevels) ? mg->lv[level].g.n : 0; }
double mg3d_level_h(const mg3d_t* mg, int level) { return (mg && level >= 0 && level < mg->nlevels) ? mg->lv[level].h[0] : 0.0; }
void* mg3d_stream(mg3d_t* mg) { return mg ? (void*)mg->stream : NULL; }
long long mg3d_kernel_launches(const mg3d_t* mg) { return mg ? mg->launches : 0; }

int mg3d_set_smoother(mg3d_t* mg, int smoother, int sweeps_per_pass)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    if (smoother < MG_SMOOTHER_AUTO || smoother > MG_SMOOTHER_FUSED) return mg_fail(MG_ERR_ARG, "bad smoother %d", smoother);
    if (sweeps_per_pass < 1 || sweeps_per_pass > 4) return mg_fail(MG_ERR_ARG, "sweeps_per_pass must be 1..4");
    mg->smoother = smoother;
    mg->sweeps_per_pass = sweeps_per_pass;
    return MG_OK;
}

int mg3d_profile(mg3d_t* mg, int enable)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    return mg_prof_enable(&mg->prof, mg->stream, enable);
}

int mg3d_profile_read(mg3d_t* mg, int level, int op, double* ms_total, long long* kernel_launches, long long* calls)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (op < 0 || op >= MG_OP_COUNT) return mg_fail(MG_ERR_ARG, "bad op %d", op);
    st = mg_prof_collect(&mg->prof, mg->stream);
    if (st) return st;
    if (ms_total) *ms_total = mg->prof.ms[level][op];
    if (kernel_launches) *kernel_launches = mg->prof.kl[level][op];
    if (calls) *calls = mg->prof.calls[level][op];
    return MG_OK;
}

int mg3d_sync(mg3d_t* mg)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

/* dense host array (x fastest, idx = x + y*n + z*n*n) <-> colour-split device field: one linear copy
   between the host array and a dense device staging buffer, plus a repack kernel. */
static int staging_reserve(mg3d_t* mg, size_t bytes)
{
    if (mg->staging_bytes >= bytes) return MG_OK;
    if (mg->staging) { MG_CUDA(cudaStreamSynchronize(mg->stream)); cudaFree(mg->staging); mg->staging = NULL; mg->staging_bytes = 0; }
    MG_CUDA(cudaMalloc(&mg->staging, bytes));
    mg->staging_bytes = bytes;
    return MG_OK;
}

static int copy_in(mg3d_t* mg, void* dev, const mg_geom3d* g, const void* host)
{
    size_t dense = (size_t)g->n * g->n * (size_t)g->nzl * mg_esize(mg->dtype);
    int st = staging_reserve(mg, dense);
    if (st) return st;
    MG_CUDA(cudaMemcpyAsync(mg->staging, host, dense, cudaMemcpyHostToDevice, mg->stream));
    MG_LAUNCH(mg->launches, mgk3d_repack(mg->stream, mg->dtype, dev, *g, mg->staging, 1, 0, g->nzl));
    return MG_OK;
}

static int copy_out(mg3d_t* mg, void* host, const void* dev, const mg_geom3d* g)
{
    size_t dense = (size_t)g->n * g->n * (size_t)g->nzl * mg_esize(mg->dtype);
    int st = staging_reserve(mg, dense);
    if (st) return st;
    MG_LAUNCH(mg->launches, mgk3d_repack(mg->stream, mg->dtype, (void*)dev, *g, mg->staging, 0, 0, g->nzl));
    MG_CUDA(cudaMemcpyAsync(host, mg->staging, dense, cudaMemcpyDeviceToHost, mg->stream));
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg3d_set_field(mg3d_t* mg, int level, int field, const void* host_dense)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_dense || (field != MG_FIELD_V && field != MG_FIELD_F)) return mg_fail(MG_ERR_ARG, "bad field/pointer");
    if (mg->nranks != 1) return mg_fail(MG_ERR_STATE, "set_field takes the whole grid: single-GPU handles only");
    st = copy_in(mg, field_ptr(&mg->lv[level], field), &mg->lv[level].g, host_dense);
    if (st) return st;
    MG_CUDA(cudaStreamSynchronize(mg->stream)); /* host buffer may be reused by the caller */
    return MG_OK;
}

int mg3d_get_field(mg3d_t* mg, int level, int field, void* host_dense)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_dense || (field != MG_FIELD_V && field != MG_FIELD_F)) return mg_fail(MG_ERR_ARG, "bad field/pointer");
    if (mg->nranks != 1) return mg_fail(MG_ERR_STATE, "get_field returns the whole grid: single-GPU handles only");
    return copy_out(mg, host_dense, field_ptr(&mg->lv[level], field), &mg->lv[level].g);
}

/* Grid3D::InitV / InitF on every level (N3/Grid3D.cpp:61-96) */
int mg3d_init_problem(mg3d_t* mg)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    const double PI = 3.141592653589793; /* N3/inclusion.h:9 */
    const int n0 = mg->lv[0].g.n;
    double* tab = (double*)malloc(3 * (size_t)n0 * sizeof(double));
    if (!tab) return mg_fail(MG_ERR_NOMEM, "host allocation failed");
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level3d* L = &mg->lv[l];
        const int n = L->g.n;
        for (int a = 0; a < 3; a++)
            for (int i = 0; i < n; i++) {
                /* float x = x_a + posX*h_x;  sin(PI*x) in double (N3/Grid3D.cpp:88-92) */
                double x;
                if (mg->dtype == MG_F32) {
                    float xf = (float)mg->range[2 * a] + i * (float)L->h[a];
                    x = xf;
                } else {
                    x = mg->range[2 * a] + i * L->h[a];
                }
                tab[(size_t)a * n + i] = sin(PI * x);
            }
        cudaError_t e = cudaMemcpyAsync(mg->d_tables, tab, 3 * (size_t)n * sizeof(double), cudaMemcpyHostToDevice, mg->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(mg->stream);
        if (e != cudaSuccess) { free(tab); return mg_fail(MG_ERR_CUDA, "table upload failed: %s", cudaGetErrorString(e)); }
        int k1 = mgk3d_set(mg->stream, mg->dtype, L->v, L->g, 0.0, 1, 0, L->g.nzl);
        int k2 = mgk3d_init_f(mg->stream, mg->dtype, L->f, L->g, mg->d_tables, mg->d_tables + n, mg->d_tables + 2 * n, 0, L->g.nzl);
        if (k1 < 0 || k2 < 0) { free(tab); return mg_fail(MG_ERR_CUDA, "init launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
        mg->launches += k1 + k2;
        e = cudaStreamSynchronize(mg->stream); /* d_tables is reused by the next level */
        if (e != cudaSuccess) { free(tab); return mg_fail(MG_ERR_CUDA, "init failed: %s", cudaGetErrorString(e)); }
    }
    free(tab);
    return MG_OK;
}

/* Relax: ncycles x (red half-sweep, black half-sweep), N3/MultiGrid3D.cpp:489-567 */
static int relax_level(mg3d_t* mg, int level, int ncycles)
{
    mg_level3d* L = &mg->lv[level];
    const int lo = L->own_lo > 1 ? L->own_lo : 1;
    const int hi = L->own_hi < L->g.nzl - 1 ? L->own_hi : L->g.nzl - 1;
    if (ncycles <= 0) return MG_OK;
    PROF_BEGIN(mg, level, MG_OP_RELAX);
    const int use_tma = L->has_tma && mg->smoother != MG_SMOOTHER_COLOUR;
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++) {
            if (use_tma)
                MG_LAUNCH(mg->launches, mgk3d_relax_colour_tma(mg->stream, mg->dtype, L->tmap_v[colour ^ 1], L->v, L->f, L->g, L->c, colour, lo, hi));
            else
                MG_LAUNCH(mg->launches, mgk3d_relax_colour(mg->stream, mg->dtype, L->v, L->f, L->g, L->c, colour, lo, hi));
        }
    PROF_END(mg);
    return MG_OK;
}

int mg3d_relax(mg3d_t* mg, int level, int ncycles)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (ncycles < 0) return mg_fail(MG_ERR_ARG, "ncycles < 0");
    return relax_level(mg, level, ncycles);
}

int mg3d_residual(mg3d_t* mg, int level, void* host_out)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_out) return mg_fail(MG_ERR_ARG, "null output");
    mg_level3d* L = &mg->lv[level];
    void* r = NULL;
    MG_CUDA(cudaMalloc(&r, field_bytes(L, mg->dtype)));
    int k = mgk3d_residual(mg->stream, mg->dtype, L->v, L->f, r, L->g, L->c, mg->mode == MG_CORRECTED, 0, L->g.nzl);
    if (k < 0) { cudaFree(r); return mg_fail(MG_ERR_CUDA, "residual launch failed: %s", cudaGetErrorString(cudaGetLastError())); }
    mg->launches += k;
    st = copy_out(mg, host_out, r, &L->g);
    cudaFree(r);
    return st;
}

int mg3d_residual_norm(mg3d_t* mg, int level, double* l2, double* linf)
{
    int st = check_level(mg, level);
    if (st) return st;
    mg_level3d* L = &mg->lv[level];
    double* out2 = mg->d_scratch + 2 * MGK_NORM_BLOCKS;
    MG_LAUNCH(mg->launches, mgk3d_residual_norm(mg->stream, mg->dtype, L->v, L->f, L->g, L->c, mg->mode == MG_CORRECTED,
                                                L->own_lo, L->own_hi, mg->d_scratch, out2));
    MG_CUDA(cudaMemcpyAsync(mg->h_out2, out2, 2 * sizeof(double), cudaMemcpyDeviceToHost, mg->stream));
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    if (l2) *l2 = sqrt(mg->h_out2[0]);
    if (linf) *linf = mg->h_out2[1];
    return MG_OK;
}

int mg3d_restrict(mg3d_t* mg, int fine_level, int field)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    if (field != MG_FIELD_V && field != MG_FIELD_F) return mg_fail(MG_ERR_ARG, "bad field");
    mg_level3d *F = &mg->lv[fine_level], *C = &mg->lv[fine_level + 1];
    MG_LAUNCH(mg->launches, mgk3d_restrict(mg->stream, mg->dtype, field_ptr(F, field), F->g, field_ptr(C, field), C->g, C->own_lo, C->own_hi));
    return MG_OK;
}

static int residual_restrict_level(mg3d_t* mg, int fine_level)
{
    mg_level3d *F = &mg->lv[fine_level], *C = &mg->lv[fine_level + 1];
    PROF_BEGIN(mg, fine_level, MG_OP_RESIDUAL_RESTRICT);
    if (F->has_tma && mg->smoother != MG_SMOOTHER_COLOUR)
        MG_LAUNCH(mg->launches, mgk3d_residual_restrict_tma(mg->stream, mg->dtype, F->tmap_rr[0], F->tmap_rr[1], F->f, F->g, F->c,
                                                            mg->mode == MG_CORRECTED, C->f, C->v, C->g, C->own_lo, C->own_hi));
    else
        MG_LAUNCH(mg->launches, mgk3d_residual_restrict(mg->stream, mg->dtype, F->v, F->f, F->g, F->c, mg->mode == MG_CORRECTED,
                                                        C->f, C->v, C->g, C->own_lo, C->own_hi));
    PROF_END(mg);
    return MG_OK;
}

int mg3d_residual_restrict(mg3d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return residual_restrict_level(mg, fine_level);
}

static int interpolate_level(mg3d_t* mg, int fine_level, int add)
{
    mg_level3d *F = &mg->lv[fine_level], *C = &mg->lv[fine_level + 1];
    const int lo = F->own_lo > 1 ? F->own_lo : 1;
    const int hi = F->own_hi < F->g.nzl - 1 ? F->own_hi : F->g.nzl - 1;
    PROF_BEGIN(mg, fine_level, MG_OP_INTERPOLATE);
    MG_LAUNCH(mg->launches, mgk3d_interpolate(mg->stream, mg->dtype, F->v, F->g, C->v, C->g, add, lo, hi));
    PROF_END(mg);
    return MG_OK;
}

int mg3d_interpolate(mg3d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return interpolate_level(mg, fine_level, 0);
}

int mg3d_interpolate_correct(mg3d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return interpolate_level(mg, fine_level, 1);
}

int mg3d_set_to_value(mg3d_t* mg, int level, int field, double value, int modify_boundaries)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (field != MG_FIELD_V && field != MG_FIELD_F) return mg_fail(MG_ERR_ARG, "bad field");
    mg_level3d* L = &mg->lv[level];
    MG_LAUNCH(mg->launches, mgk3d_set(mg->stream, mg->dtype, field_ptr(L, field), L->g, value, modify_boundaries, L->own_lo, L->own_hi));
    return MG_OK;
}

/* VCycle, N3/MultiGrid3D.cpp:623-647.  CalculateResidual + Restrict + setToValue(coarse v, 0, true)
   are one kernel; Interpolate + ApplyCorrection are one kernel: no residual / error grid exists. */
static int vcycle_rec(mg3d_t* mg, int level, int v1, int v2)
{
    int st = relax_level(mg, level, v1);
    if (st) return st;
    if (level != mg->nlevels - 1) {
        st = residual_restrict_level(mg, level);
        if (st) return st;
        st = vcycle_rec(mg, level + 1, v1, v2);
        if (st) return st;
        st = interpolate_level(mg, level, 1);
        if (st) return st;
    }
    return relax_level(mg, level, v2);
}

int mg3d_vcycle(mg3d_t* mg, int level, int v1, int v2)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (v1 < 0 || v2 < 0) return mg_fail(MG_ERR_ARG, "negative sweep count");
    return vcycle_rec(mg, level, v1, v2);
}

/* FullMultiGridVCycle, N3/MultiGrid3D.cpp:569-585 */
static int fmg_rec(mg3d_t* mg, int level, int v0, int v1, int v2)
{
    int st;
    if (level != mg->nlevels - 1) {
        mg_level3d *F = &mg->lv[level], *C = &mg->lv[level + 1];
        PROF_BEGIN(mg, level, MG_OP_OTHER);
        MG_LAUNCH(mg->launches, mgk3d_restrict(mg->stream, mg->dtype, F->f, F->g, C->f, C->g, C->own_lo, C->own_hi));
        PROF_END(mg);
        st = fmg_rec(mg, level + 1, v0, v1, v2);
        if (st) return st;
        st = interpolate_level(mg, level, 0);
        if (st) return st;
    } else {
        mg_level3d* L = &mg->lv[level];
        MG_LAUNCH(mg->launches, mgk3d_set(mg->stream, mg->dtype, L->v, L->g, 0.0, 0, L->own_lo, L->own_hi));
    }
    for (int i = 0; i < v0; i++) {
        st = vcycle_rec(mg, level, v1, v2);
        if (st) return st;
    }
    return MG_OK;
}

int mg3d_fmg(mg3d_t* mg, int level, int v0, int v1, int v2)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (v0 < 0 || v1 < 0 || v2 < 0) return mg_fail(MG_ERR_ARG, "negative cycle/sweep count");
    return fmg_rec(mg, level, v0, v1, v2);
}

/* ---- reference-facing operators on HOST arrays (N3/MultiGrid3D.h:16-27) ---------------------- */

static int cubic(const int s[3], int* n)
{
    if (!s || s[0] != s[1] || s[0] != s[2] || s[0] < 3) return mg_fail(MG_ERR_ARG, "cubic size >= 3 required");
    *n = s[0];
    return MG_OK;
}

static void temp_geom(int n, int dtype, mg_geom3d* g) { set_geom(g, n, dtype, 0, n); }

static int temp_alloc(mg3d_t* mg, const mg_geom3d* g, void** p)
{
    MG_CUDA(cudaMalloc(p, 2 * (size_t)g->cstride * mg_esize(mg->dtype)));
    return MG_OK;
}

int mg3d_restrict_host(mg3d_t* mg, const void* fine, const int fs[3], void* coarse, const int cs[3])
{
    int fn = 0, cn = 0, st;
    if (!mg || !fine || !coarse) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = cubic(fs, &fn)) || (st = cubic(cs, &cn))) return st;
    if (cn != (fn - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1"); /* N3/MultiGrid3D.cpp:60-62 */
    mg_geom3d gf, gc;
    temp_geom(fn, mg->dtype, &gf);
    temp_geom(cn, mg->dtype, &gc);
    void *df = NULL, *dc = NULL;
    if ((st = temp_alloc(mg, &gf, &df))) return st;
    if ((st = temp_alloc(mg, &gc, &dc))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &gf, fine);
    if (!st) {
        int k = mgk3d_restrict(mg->stream, mg->dtype, df, gf, dc, gc, 0, cn);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "restrict launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, coarse, dc, &gc);
    cudaFree(df); cudaFree(dc);
    return st;
}

int mg3d_interpolate_host(mg3d_t* mg, void* fine, const int fs[3], const void* coarse, const int cs[3])
{
    int fn = 0, cn = 0, st;
    if (!mg || !fine || !coarse) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = cubic(fs, &fn)) || (st = cubic(cs, &cn))) return st;
    if (cn != (fn - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1"); /* N3/MultiGrid3D.cpp:196-198 */
    mg_geom3d gf, gc;
    temp_geom(fn, mg->dtype, &gf);
    temp_geom(cn, mg->dtype, &gc);
    void *df = NULL, *dc = NULL;
    if ((st = temp_alloc(mg, &gf, &df))) return st;
    if ((st = temp_alloc(mg, &gc, &dc))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &gf, fine); /* boundary of fine is kept */
    if (!st) st = copy_in(mg, dc, &gc, coarse);
    if (!st) {
        int k = mgk3d_interpolate(mg->stream, mg->dtype, df, gf, dc, gc, 0, 1, fn - 1);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "interpolate launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, fine, df, &gf);
    cudaFree(df); cudaFree(dc);
    return st;
}

int mg3d_apply_correction_host(mg3d_t* mg, void* fine, const int fs[3], const void* error, const int es[3])
{
    int fn = 0, en = 0, st;
    if (!mg || !fine || !error) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = cubic(fs, &fn)) || (st = cubic(es, &en))) return st;
    if (fn != en) return mg_fail(MG_ERR_ARG, "fsize != esize"); /* N3/MultiGrid3D.cpp:660-662 */
    mg_geom3d g;
    temp_geom(fn, mg->dtype, &g);
    void *df = NULL, *de = NULL;
    if ((st = temp_alloc(mg, &g, &df))) return st;
    if ((st = temp_alloc(mg, &g, &de))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &g, fine);
    if (!st) st = copy_in(mg, de, &g, error);
    if (!st) {
        int k = mgk3d_apply_correction(mg->stream, mg->dtype, df, de, g, 1, fn - 1);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "apply_correction launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, fine, df, &g);
    cudaFree(df); cudaFree(de);
    return st;
}

int mg3d_set_to_value_host(mg3d_t* mg, void* grid, const int s[3], double value, int modify_boundaries)
{
    int n = 0, st;
    if (!mg || !grid) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = cubic(s, &n))) return st;
    mg_geom3d g;
    temp_geom(n, mg->dtype, &g);
    void* d = NULL;
    if ((st = temp_alloc(mg, &g, &d))) return st;
    st = copy_in(mg, d, &g, grid);
    if (!st) {
        int k = mgk3d_set(mg->stream, mg->dtype, d, g, value, modify_boundaries, 0, n);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "set launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, grid, d, &g);
    cudaFree(d);
    return st;
}

int mg3d_vcycle_host(mg3d_t* mg, void* v_host, const void* f_host, int v1, int v2, int cycles)
{
    if (!mg || !v_host || !f_host) return mg_fail(MG_ERR_ARG, "null argument");
    if (mg->nranks != 1) return mg_fail(MG_ERR_STATE, "vcycle_host: single-GPU handles only");
    if (v1 < 0 || v2 < 0 || cycles < 0) return mg_fail(MG_ERR_ARG, "negative count");
    mg_level3d* L = &mg->lv[0];
    int st = copy_in(mg, L->v, &L->g, v_host);
    if (!st) st = copy_in(mg, L->f, &L->g, f_host);
    for (int i = 0; i < cycles && !st; i++) st = vcycle_rec(mg, 0, v1, v2);
    if (!st) st = copy_out(mg, v_host, L->v, &L->g);
    return st;
}

/* multi-GPU entry points: implemented in mg3d_dist.c */
