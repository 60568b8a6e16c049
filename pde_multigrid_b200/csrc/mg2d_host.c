/*
 * mg2d_host.c -- C host driver of the 2D Lyapunov multigrid (first-order upwind PDE
 * K1 V_x + K2 V_y + alfa V = f, K = A x): hierarchy, V-cycle, FMG, field I/O.
 *
 * Mirrors the reference class MultiGrid2D (NOCUDA_TESI/PDE Lyapunov 2D/MultiGrid2D.cpp: InitGrids
 * :20-42, InitA :45-60, VCycle :314-340, FullMultiGridVCycle :296-312) over the kernels of
 * mg2d_kernels.cu.  No CPU compute path: the host only derives h per level.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "mg_host_common.h"
#include "mg_profile.h"

typedef struct {
    mg_geom2d g;
    mg_coef2d c;
    void* v;
    void* f;
} mg_level2d;

/* CUDA-graph cache of whole V-cycles (key = level, v1, v2): at 1025^2 a V(2,2) is ~60 dependent launches on
   L2-resident fields, i.e. pure launch latency; the second call with a key captures the cycle, later calls replay it */
#define MG2_GRAPH_SLOTS 4
typedef struct {
    int used, level, v1, v2, calls;
    cudaGraphExec_t exec;
    long long launches;
} mg2_graph_slot;

struct mg2d_s {
    int dtype, nlevels;
    int use_graphs;
    mg2_graph_slot graphs[MG2_GRAPH_SLOTS];
    cudaStream_t stream;
    mg_level2d* lv;
    void* arena;
    double* d_scratch; /* 2*1184 partials + 2 outputs */
    double* h_out2;
    long long launches;
    mg_prof prof;
};

#define PROF_BEGIN(mg, level, op) mg_prof_begin(&(mg)->prof, (mg)->stream, (level), (op), (mg)->launches)
#define PROF_END(mg) mg_prof_end(&(mg)->prof, (mg)->stream, (mg)->launches)
#define SCRATCH_PARTS 1184

static size_t field_bytes2(const mg_geom2d* g, int dtype) { return mg_align256((size_t)g->pitch * g->n * mg_esize(dtype)); }

static int check_level(const mg2d_t* mg, int level)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    if (level < 0 || level >= mg->nlevels) return mg_fail(MG_ERR_ARG, "level %d out of range [0,%d)", level, mg->nlevels);
    return MG_OK;
}

static void* field_ptr(mg_level2d* L, int field) { return field == MG_FIELD_V ? L->v : L->f; }

/* h = range/(real)(n-1) in the level's precision, N2/Grid2D.cpp:24-35 */
static void level_coefs(int dtype, int n, const double* range, const double* A4, int alfa, mg_coef2d* c)
{
    if (dtype == MG_F32) {
        float xr = (float)range[1] - (float)range[0], yr = (float)range[3] - (float)range[2];
        c->hx = xr / (float)(n - 1);
        c->hy = yr / (float)(n - 1);
        c->xa = (float)range[0];
        c->ya = (float)range[2];
        for (int i = 0; i < 4; i++) c->A[i] = (float)A4[i];
    } else {
        double xr = range[1] - range[0], yr = range[3] - range[2];
        c->hx = xr / (double)(n - 1);
        c->hy = yr / (double)(n - 1);
        c->xa = range[0];
        c->ya = range[2];
        for (int i = 0; i < 4; i++) c->A[i] = A4[i];
    }
    c->alfa = alfa;
}

int mg2d_create(mg2d_t** out, const int sz[2], const double range[4], const double A4[4], int alfa, int dtype)
{
    if (!out || !sz || !range || !A4) return mg_fail(MG_ERR_ARG, "null argument");
    *out = NULL;
    if (sz[0] != sz[1]) return mg_fail(MG_ERR_ARG, "sizeX == sizeY required (got %d,%d)", sz[0], sz[1]); /* N2/Grid2D.cpp:9 */
    const int n = sz[0];
    if (n < 3 || ((n - 1) & (n - 2)) != 0) return mg_fail(MG_ERR_ARG, "size must be 2^k+1 with k >= 1 (got %d)", n);
    if (!(range[1] > range[0]) || !(range[3] > range[2])) return mg_fail(MG_ERR_ARG, "range must satisfy b > a on every axis");
    if (dtype != MG_F32 && dtype != MG_F64) return mg_fail(MG_ERR_ARG, "dtype must be MG_F32 or MG_F64");
    int st = mg_require_device();
    if (st) return st;
    mg2d_t* mg = (mg2d_t*)calloc(1, sizeof *mg);
    if (!mg) return mg_fail(MG_ERR_NOMEM, "host allocation failed");
    mg->dtype = dtype;
    mg->use_graphs = getenv("MG_B200_NO_GRAPH") ? 0 : 1;
    mg->nlevels = mg_num_levels_for(n);
    mg->lv = (mg_level2d*)calloc((size_t)mg->nlevels, sizeof(mg_level2d));
    if (!mg->lv) { free(mg); return mg_fail(MG_ERR_NOMEM, "host allocation failed"); }
    size_t total = 0;
    int nl = n;
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level2d* L = &mg->lv[l];
        L->g.n = nl;
        L->g.pitch = mg_pitch(nl, dtype);
        level_coefs(dtype, nl, range, A4, alfa, &L->c);
        total += 2 * field_bytes2(&L->g, dtype);
        nl = (nl - 1) / 2 + 1;
    }
    cudaError_t e = cudaMalloc(&mg->arena, total);
    if (e != cudaSuccess) {
        free(mg->lv); free(mg);
        return mg_fail(e == cudaErrorMemoryAllocation ? MG_ERR_NOMEM : MG_ERR_CUDA, "cudaMalloc of %zu bytes failed: %s", total, cudaGetErrorString(e));
    }
    char* p = (char*)mg->arena;
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level2d* L = &mg->lv[l];
        L->v = p; p += field_bytes2(&L->g, dtype);
        L->f = p; p += field_bytes2(&L->g, dtype);
    }
    if (cudaStreamCreateWithFlags(&mg->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&mg->d_scratch, (2 * SCRATCH_PARTS + 2) * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&mg->h_out2, 2 * sizeof(double)) != cudaSuccess ||
        cudaMemsetAsync(mg->arena, 0, total, mg->stream) != cudaSuccess) {
        int code = mg_fail(MG_ERR_CUDA, "stream/scratch setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        mg2d_destroy(mg);
        return code;
    }
    st = mg2d_init_problem(mg);
    if (st) { mg2d_destroy(mg); return st; }
    *out = mg;
    return MG_OK;
}

int mg2d_destroy(mg2d_t* mg)
{
    if (!mg) return MG_OK;
    if (mg->stream) cudaStreamSynchronize(mg->stream);
    for (int i = 0; i < MG2_GRAPH_SLOTS; i++)
        if (mg->graphs[i].used && mg->graphs[i].exec) cudaGraphExecDestroy(mg->graphs[i].exec);
    if (mg->stream) cudaStreamDestroy(mg->stream);
    if (mg->arena) cudaFree(mg->arena);
    if (mg->d_scratch) cudaFree(mg->d_scratch);
    if (mg->h_out2) cudaFreeHost(mg->h_out2);
    mg_prof_free(&mg->prof);
    free(mg->lv);
    free(mg);
    return MG_OK;
}

int mg2d_num_levels(const mg2d_t* mg) { return mg ? mg->nlevels : 0; }
int mg2d_level_size(const mg2d_t* mg, int level) { return (mg && level >= 0 && level < mg->nlevels) ? mg->lv[level].g.n : 0; }
double mg2d_level_h(const mg2d_t* mg, int level) { return (mg && level >= 0 && level < mg->nlevels) ? mg->lv[level].c.hx : 0.0; }
void* mg2d_stream(mg2d_t* mg) { return mg ? (void*)mg->stream : NULL; }
long long mg2d_kernel_launches(const mg2d_t* mg) { return mg ? mg->launches : 0; }

int mg2d_sync(mg2d_t* mg)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg2d_profile(mg2d_t* mg, int enable)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    return mg_prof_enable(&mg->prof, mg->stream, enable);
}

int mg2d_profile_read(mg2d_t* mg, int level, int op, double* ms_total, long long* kernel_launches, long long* calls)
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

static int copy_in(mg2d_t* mg, void* dev, const mg_geom2d* g, const void* host)
{
    size_t es = mg_esize(mg->dtype);
    MG_CUDA(cudaMemcpy2DAsync(dev, (size_t)g->pitch * es, host, (size_t)g->n * es, (size_t)g->n * es, (size_t)g->n,
                              cudaMemcpyHostToDevice, mg->stream));
    return MG_OK;
}

static int copy_out(mg2d_t* mg, void* host, const void* dev, const mg_geom2d* g)
{
    size_t es = mg_esize(mg->dtype);
    MG_CUDA(cudaMemcpy2DAsync(host, (size_t)g->n * es, dev, (size_t)g->pitch * es, (size_t)g->n * es, (size_t)g->n,
                              cudaMemcpyDeviceToHost, mg->stream));
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg2d_set_field(mg2d_t* mg, int level, int field, const void* host_dense)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_dense || (field != MG_FIELD_V && field != MG_FIELD_F)) return mg_fail(MG_ERR_ARG, "bad field/pointer");
    st = copy_in(mg, field_ptr(&mg->lv[level], field), &mg->lv[level].g, host_dense);
    if (st) return st;
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg2d_get_field(mg2d_t* mg, int level, int field, void* host_dense)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_dense || (field != MG_FIELD_V && field != MG_FIELD_F)) return mg_fail(MG_ERR_ARG, "bad field/pointer");
    return copy_out(mg, host_dense, field_ptr(&mg->lv[level], field), &mg->lv[level].g);
}

/* Grid2D::InitV / InitF on every level, N2/Grid2D.cpp:50-80 */
int mg2d_init_problem(mg2d_t* mg)
{
    if (!mg) return mg_fail(MG_ERR_ARG, "null handle");
    for (int l = 0; l < mg->nlevels; l++) {
        mg_level2d* L = &mg->lv[l];
        MG_LAUNCH(mg->launches, mgk2d_init_v(mg->stream, mg->dtype, L->v, L->g, L->c));
        MG_LAUNCH(mg->launches, mgk2d_set(mg->stream, mg->dtype, L->f, L->g, 0.0, 1));
    }
    return MG_OK;
}

/* Relax, N2/MultiGrid2D.cpp:199-273 */
static int relax_level(mg2d_t* mg, int level, int ncycles)
{
    mg_level2d* L = &mg->lv[level];
    if (ncycles <= 0) return MG_OK;
    PROF_BEGIN(mg, level, MG_OP_RELAX);
    if (L->g.n <= MGK2D_SMALL_N) { /* the whole level fits in one CTA: every sweep in a single launch */
        MG_LAUNCH(mg->launches, mgk2d_relax_small(mg->stream, mg->dtype, L->v, L->f, L->g, L->c, ncycles));
    } else {
        for (int k = 0; k < ncycles; k++)
            for (int colour = 0; colour < 2; colour++)
                MG_LAUNCH(mg->launches, mgk2d_relax_colour(mg->stream, mg->dtype, L->v, L->f, L->g, L->c, colour));
    }
    PROF_END(mg);
    return MG_OK;
}

int mg2d_relax(mg2d_t* mg, int level, int ncycles)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (ncycles < 0) return mg_fail(MG_ERR_ARG, "ncycles < 0");
    return relax_level(mg, level, ncycles);
}

int mg2d_residual(mg2d_t* mg, int level, void* host_out)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!host_out) return mg_fail(MG_ERR_ARG, "null output");
    mg_level2d* L = &mg->lv[level];
    void* r = NULL;
    MG_CUDA(cudaMalloc(&r, field_bytes2(&L->g, mg->dtype)));
    int k = mgk2d_residual(mg->stream, mg->dtype, L->v, L->f, r, L->g, L->c);
    if (k < 0) { cudaFree(r); return mg_fail(MG_ERR_CUDA, "residual launch failed"); }
    mg->launches += k;
    st = copy_out(mg, host_out, r, &L->g);
    cudaFree(r);
    return st;
}

static int read_out2(mg2d_t* mg)
{
    MG_CUDA(cudaMemcpyAsync(mg->h_out2, mg->d_scratch + 2 * SCRATCH_PARTS, 2 * sizeof(double), cudaMemcpyDeviceToHost, mg->stream));
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg2d_residual_norm(mg2d_t* mg, int level, double* l2, double* linf)
{
    int st = check_level(mg, level);
    if (st) return st;
    mg_level2d* L = &mg->lv[level];
    MG_LAUNCH(mg->launches, mgk2d_residual_norm(mg->stream, mg->dtype, L->v, L->f, L->g, L->c, mg->d_scratch, mg->d_scratch + 2 * SCRATCH_PARTS));
    st = read_out2(mg);
    if (st) return st;
    if (l2) *l2 = sqrt(mg->h_out2[0]);
    if (linf) *linf = mg->h_out2[1];
    return MG_OK;
}

int mg2d_mean_abs_error(mg2d_t* mg, double* mae)
{
    if (!mg || !mae) return mg_fail(MG_ERR_ARG, "null argument");
    mg_level2d* L = &mg->lv[0];
    MG_LAUNCH(mg->launches, mgk2d_abs_error_sum(mg->stream, mg->dtype, L->v, L->g, L->c, mg->d_scratch, mg->d_scratch + 2 * SCRATCH_PARTS));
    int st = read_out2(mg);
    if (st) return st;
    long long ni = (long long)(L->g.n - 2) * (L->g.n - 2); /* interior points only, C2/Grid2D.cu:130-148 */
    *mae = ni > 0 ? mg->h_out2[0] / (double)ni : 0.0;
    return MG_OK;
}

int mg2d_restrict(mg2d_t* mg, int fine_level, int field)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    if (field != MG_FIELD_V && field != MG_FIELD_F) return mg_fail(MG_ERR_ARG, "bad field");
    mg_level2d *F = &mg->lv[fine_level], *C = &mg->lv[fine_level + 1];
    MG_LAUNCH(mg->launches, mgk2d_restrict(mg->stream, mg->dtype, field_ptr(F, field), F->g, field_ptr(C, field), C->g));
    return MG_OK;
}

static int residual_restrict_level(mg2d_t* mg, int level)
{
    mg_level2d *F = &mg->lv[level], *C = &mg->lv[level + 1];
    PROF_BEGIN(mg, level, MG_OP_RESIDUAL_RESTRICT);
    MG_LAUNCH(mg->launches, mgk2d_residual_restrict(mg->stream, mg->dtype, F->v, F->f, F->g, F->c, C->f, C->v, C->g));
    PROF_END(mg);
    return MG_OK;
}

int mg2d_residual_restrict(mg2d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return residual_restrict_level(mg, fine_level);
}

static int interpolate_level(mg2d_t* mg, int fine_level, int add)
{
    mg_level2d *F = &mg->lv[fine_level], *C = &mg->lv[fine_level + 1];
    PROF_BEGIN(mg, fine_level, MG_OP_INTERPOLATE);
    MG_LAUNCH(mg->launches, mgk2d_interpolate(mg->stream, mg->dtype, F->v, F->g, C->v, C->g, add));
    PROF_END(mg);
    return MG_OK;
}

int mg2d_interpolate(mg2d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return interpolate_level(mg, fine_level, 0);
}

int mg2d_interpolate_correct(mg2d_t* mg, int fine_level)
{
    int st = check_level(mg, fine_level);
    if (st) return st;
    if (fine_level == mg->nlevels - 1) return mg_fail(MG_ERR_ARG, "level %d is the coarsest", fine_level);
    return interpolate_level(mg, fine_level, 1);
}

int mg2d_set_to_value(mg2d_t* mg, int level, int field, double value, int modify_boundaries)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (field != MG_FIELD_V && field != MG_FIELD_F) return mg_fail(MG_ERR_ARG, "bad field");
    mg_level2d* L = &mg->lv[level];
    MG_LAUNCH(mg->launches, mgk2d_set(mg->stream, mg->dtype, field_ptr(L, field), L->g, value, modify_boundaries));
    return MG_OK;
}

/* VCycle, N2/MultiGrid2D.cpp:314-340 */
static int vcycle_rec(mg2d_t* mg, int level, int v1, int v2)
{
    int st = relax_level(mg, level, v1);
    if (st) return st;
    if (level != mg->nlevels - 1) {
        if ((st = residual_restrict_level(mg, level))) return st;
        if ((st = vcycle_rec(mg, level + 1, v1, v2))) return st;
        if ((st = interpolate_level(mg, level, 1))) return st;
    }
    return relax_level(mg, level, v2);
}

int mg2d_vcycle(mg2d_t* mg, int level, int v1, int v2)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (v1 < 0 || v2 < 0) return mg_fail(MG_ERR_ARG, "negative sweep count");
    const long long per_cycle = 2LL * (v1 + v2) * (mg->nlevels - level);
    if (!mg->use_graphs || mg->prof.enabled || per_cycle > 4096) return vcycle_rec(mg, level, v1, v2);
    mg2_graph_slot* g = NULL;
    for (int i = 0; i < MG2_GRAPH_SLOTS && !g; i++)
        if (mg->graphs[i].used && mg->graphs[i].level == level && mg->graphs[i].v1 == v1 && mg->graphs[i].v2 == v2) g = &mg->graphs[i];
    if (!g) {
        for (int i = 0; i < MG2_GRAPH_SLOTS && !g; i++)
            if (!mg->graphs[i].used) g = &mg->graphs[i];
        if (!g) return vcycle_rec(mg, level, v1, v2); /* cache full: run eagerly */
        memset(g, 0, sizeof *g);
        g->used = 1; g->level = level; g->v1 = v1; g->v2 = v2;
    }
    if (++g->calls == 1) return vcycle_rec(mg, level, v1, v2); /* first call eager: kernel attributes, lazy loading */
    if (!g->exec) {
        const long long l0 = mg->launches;
        cudaGraph_t graph = NULL;
        MG_CUDA(cudaStreamBeginCapture(mg->stream, cudaStreamCaptureModeThreadLocal));
        st = vcycle_rec(mg, level, v1, v2);
        cudaError_t e = cudaStreamEndCapture(mg->stream, &graph);
        g->launches = mg->launches - l0;
        mg->launches = l0; /* nothing ran during the capture */
        if (!st && e == cudaSuccess && graph) e = cudaGraphInstantiate(&g->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (st || e != cudaSuccess || !g->exec) {
            g->exec = NULL;
            cudaGetLastError();
            mg->use_graphs = 0; /* capture is not possible here: stay eager */
            return st ? st : vcycle_rec(mg, level, v1, v2);
        }
    }
    MG_CUDA(cudaGraphLaunch(g->exec, mg->stream));
    mg->launches += g->launches;
    return MG_OK;
}

/* FullMultiGridVCycle, N2/MultiGrid2D.cpp:296-312 */
static int fmg_rec(mg2d_t* mg, int level, int v0, int v1, int v2)
{
    int st;
    if (level != mg->nlevels - 1) {
        mg_level2d *F = &mg->lv[level], *C = &mg->lv[level + 1];
        MG_LAUNCH(mg->launches, mgk2d_restrict(mg->stream, mg->dtype, F->f, F->g, C->f, C->g));
        if ((st = fmg_rec(mg, level + 1, v0, v1, v2))) return st;
        if ((st = interpolate_level(mg, level, 0))) return st;
    } else {
        mg_level2d* L = &mg->lv[level];
        MG_LAUNCH(mg->launches, mgk2d_set(mg->stream, mg->dtype, L->v, L->g, 0.0, 0));
    }
    for (int i = 0; i < v0; i++)
        if ((st = vcycle_rec(mg, level, v1, v2))) return st;
    return MG_OK;
}

int mg2d_fmg(mg2d_t* mg, int level, int v0, int v1, int v2)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (v0 < 0 || v1 < 0 || v2 < 0) return mg_fail(MG_ERR_ARG, "negative cycle/sweep count");
    return fmg_rec(mg, level, v0, v1, v2);
}

/* ---- reference-facing operators on HOST arrays (N2/MultiGrid2D.h:21-27) ---------------------- */

static int square(const int s[2], int* n)
{
    if (!s || s[0] != s[1] || s[0] < 3) return mg_fail(MG_ERR_ARG, "square size >= 3 required");
    *n = s[0];
    return MG_OK;
}

static int temp2(mg2d_t* mg, int n, mg_geom2d* g, void** p)
{
    g->n = n;
    g->pitch = mg_pitch(n, mg->dtype);
    MG_CUDA(cudaMalloc(p, (size_t)g->pitch * n * mg_esize(mg->dtype)));
    return MG_OK;
}

int mg2d_restrict_host(mg2d_t* mg, const void* fine, const int fs[2], void* coarse, const int cs[2])
{
    int fn = 0, cn = 0, st;
    if (!mg || !fine || !coarse) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = square(fs, &fn)) || (st = square(cs, &cn))) return st;
    if (cn != (fn - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1"); /* N2/MultiGrid2D.cpp:71-72 */
    mg_geom2d gf, gc;
    void *df = NULL, *dc = NULL;
    if ((st = temp2(mg, fn, &gf, &df))) return st;
    if ((st = temp2(mg, cn, &gc, &dc))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &gf, fine);
    if (!st) {
        int k = mgk2d_restrict(mg->stream, mg->dtype, df, gf, dc, gc);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "restrict launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, coarse, dc, &gc);
    cudaFree(df); cudaFree(dc);
    return st;
}

int mg2d_interpolate_host(mg2d_t* mg, void* fine, const int fs[2], const void* coarse, const int cs[2])
{
    int fn = 0, cn = 0, st;
    if (!mg || !fine || !coarse) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = square(fs, &fn)) || (st = square(cs, &cn))) return st;
    if (cn != (fn - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1"); /* N2/MultiGrid2D.cpp:136-137 */
    mg_geom2d gf, gc;
    void *df = NULL, *dc = NULL;
    if ((st = temp2(mg, fn, &gf, &df))) return st;
    if ((st = temp2(mg, cn, &gc, &dc))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &gf, fine);
    if (!st) st = copy_in(mg, dc, &gc, coarse);
    if (!st) {
        int k = mgk2d_interpolate(mg->stream, mg->dtype, df, gf, dc, gc, 0);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "interpolate launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, fine, df, &gf);
    cudaFree(df); cudaFree(dc);
    return st;
}

int mg2d_apply_correction_host(mg2d_t* mg, void* fine, const int fs[2], const void* error, const int es[2])
{
    int fn = 0, en = 0, st;
    if (!mg || !fine || !error) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = square(fs, &fn)) || (st = square(es, &en))) return st;
    if (fn != en) return mg_fail(MG_ERR_ARG, "fsize != esize"); /* N2/MultiGrid2D.cpp:351-352 */
    mg_geom2d g, g2;
    void *df = NULL, *de = NULL;
    if ((st = temp2(mg, fn, &g, &df))) return st;
    if ((st = temp2(mg, fn, &g2, &de))) { cudaFree(df); return st; }
    st = copy_in(mg, df, &g, fine);
    if (!st) st = copy_in(mg, de, &g, error);
    if (!st) {
        int k = mgk2d_apply_correction(mg->stream, mg->dtype, df, de, g);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "apply_correction launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, fine, df, &g);
    cudaFree(df); cudaFree(de);
    return st;
}

int mg2d_set_to_value_host(mg2d_t* mg, void* grid, const int s[2], double value, int modify_boundaries)
{
    int n = 0, st;
    if (!mg || !grid) return mg_fail(MG_ERR_ARG, "null argument");
    if ((st = square(s, &n))) return st;
    mg_geom2d g;
    void* d = NULL;
    if ((st = temp2(mg, n, &g, &d))) return st;
    st = copy_in(mg, d, &g, grid);
    if (!st) {
        int k = mgk2d_set(mg->stream, mg->dtype, d, g, value, modify_boundaries);
        if (k < 0) st = mg_fail(MG_ERR_CUDA, "set launch failed"); else mg->launches += k;
    }
    if (!st) st = copy_out(mg, grid, d, &g);
    cudaFree(d);
    return st;
}

/* ---- the CUDA_TESI faces (C2/MultiGrid2D.h:20-25, C2/Grid2D.h:19-22): operators on pitched DEVICE arrays given as
        (pointer, size, pitch in elements).  The engine's own 2D layout is pitched as well, so the kernels run on the
        caller's arrays in place, and a level's fields can be handed out without a copy. ---------------------------------- */
int mg2d_level_device_ptr(mg2d_t* mg, int level, int field, void** ptr, int* pitch_elems)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!ptr || (field != MG_FIELD_V && field != MG_FIELD_F)) return mg_fail(MG_ERR_ARG, "bad field/pointer");
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    *ptr = field == MG_FIELD_V ? mg->lv[level].v : mg->lv[level].f;
    if (pitch_elems) *pitch_elems = mg->lv[level].g.pitch;
    return MG_OK;
}

static int dev2_args(mg2d_t* mg, const void* a, int an, int ap, const void* b, int bn, int bp, int need_b)
{
    if (!mg || !a || (need_b && !b)) return mg_fail(MG_ERR_ARG, "null argument");
    if (an < 3 || ap < an || (need_b && (bn < 3 || bp < bn))) return mg_fail(MG_ERR_ARG, "size >= 3 and pitch >= size required");
    MG_CUDA(cudaDeviceSynchronize()); /* the caller's arrays may have been produced on any stream */
    return MG_OK;
}

static int dev2_done(mg2d_t* mg, int k)
{
    if (k < 0) return mg_fail(MG_ERR_CUDA, "operator launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    mg->launches += k;
    MG_CUDA(cudaStreamSynchronize(mg->stream));
    return MG_OK;
}

int mg2d_restrict_device(mg2d_t* mg, const void* fine, int fsize, int f_pitch, void* coarse, int csize, int c_pitch)
{
    int st = dev2_args(mg, fine, fsize, f_pitch, coarse, csize, c_pitch, 1);
    if (st) return st;
    if (csize != (fsize - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1");
    mg_geom2d gf = {fsize, f_pitch}, gc = {csize, c_pitch};
    return dev2_done(mg, mgk2d_restrict(mg->stream, mg->dtype, fine, gf, coarse, gc));
}

int mg2d_interpolate_device(mg2d_t* mg, void* fine, int fsize, int f_pitch, const void* coarse, int csize, int c_pitch)
{
    int st = dev2_args(mg, fine, fsize, f_pitch, coarse, csize, c_pitch, 1);
    if (st) return st;
    if (csize != (fsize - 1) / 2 + 1) return mg_fail(MG_ERR_ARG, "csize != (fsize-1)/2+1");
    mg_geom2d gf = {fsize, f_pitch}, gc = {csize, c_pitch};
    return dev2_done(mg, mgk2d_interpolate(mg->stream, mg->dtype, fine, gf, coarse, gc, 0));
}

int mg2d_apply_correction_device(mg2d_t* mg, void* fine, int fsize, int f_pitch, const void* error, int esize, int e_pitch)
{
    int st = dev2_args(mg, fine, fsize, f_pitch, error, esize, e_pitch, 1);
    if (st) return st;
    if (fsize != esize || f_pitch != e_pitch) return mg_fail(MG_ERR_ARG, "fine and error must have the same size and pitch");
    mg_geom2d g = {fsize, f_pitch};
    return dev2_done(mg, mgk2d_apply_correction(mg->stream, mg->dtype, fine, error, g));
}

int mg2d_set_device(mg2d_t* mg, void* v, int size, int pitch, double value, int modify_border)
{
    int st = dev2_args(mg, v, size, pitch, NULL, 0, 0, 0);
    if (st) return st;
    mg_geom2d g = {size, pitch};
    return dev2_done(mg, mgk2d_set(mg->stream, mg->dtype, v, g, value, modify_border));
}

/* CalculateResidual(grid) into a caller-owned pitched DEVICE array with the level's own pitch (C2/MultiGrid2D.cu:105-127) */
int mg2d_residual_device(mg2d_t* mg, int level, void* dev_out)
{
    int st = check_level(mg, level);
    if (st) return st;
    if (!dev_out) return mg_fail(MG_ERR_ARG, "null output");
    mg_level2d* L = &mg->lv[level];
    MG_CUDA(cudaDeviceSynchronize());
    return dev2_done(mg, mgk2d_residual(mg->stream, mg->dtype, L->v, L->f, dev_out, L->g, L->c));
}

int mg2d_vcycle_host(mg2d_t* mg, void* v_host, const void* f_host, int v1, int v2, int cycles)
{
    if (!mg || !v_host || !f_host) return mg_fail(MG_ERR_ARG, "null argument");
    if (v1 < 0 || v2 < 0 || cycles < 0) return mg_fail(MG_ERR_ARG, "negative count");
    mg_level2d* L = &mg->lv[0];
    int st = copy_in(mg, L->v, &L->g, v_host);
    if (!st) st = copy_in(mg, L->f, &L->g, f_host);
    for (int i = 0; i < cycles && !st; i++) st = vcycle_rec(mg, 0, v1, v2);
    if (!st) st = copy_out(mg, v_host, L->v, &L->g);
    return st;
}
