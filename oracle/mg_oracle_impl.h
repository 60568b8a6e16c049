/*
 * oracle/mg_oracle_impl.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's CPU multigrid algorithm (NOCUDA_TESI), included
 * twice by mg_oracle.c with REAL = float / double and SFX = _f32 / _f64.
 *
 * Differences from the reference that do NOT change any result bit:
 *   - loops run z-outer / x-inner (the reference runs y/x/z with z the slowest memory index,
 *     N3/MultiGrid3D.cpp:509-518); every operator is either pointwise on old data or ordered
 *     only by colour, so the visiting order inside one pass is irrelevant;
 *   - 64-bit indices (the reference's `int idx` overflows at 2049^3);
 *   - no per-cycle leaks: residual/error temporaries are allocated and freed per level.
 * Expression order, operand types and division are kept exactly as written in the reference
 * (SURVEY.md Appendix A); build with -ffp-contract=off and never with -ffast-math.
 *
 * Pinned by tests/test_oracle.py against oracle/_ref (the reference itself, compiled) and the
 * golden vectors under tests/golden/ generated from it.
 */

#define ORC_CAT2(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT2(a, b)
#define FN(name) ORC_CAT(name, SFX)

/* ------------------------------------------------------------------------------------------ */
/* 3D Poisson                                                                                  */
/* ------------------------------------------------------------------------------------------ */

#define IDX3(x, y, z) ((size_t)(x) + (size_t)(y) * (size_t)n + (size_t)(z) * (size_t)n * (size_t)n)

/* h = range/(real)(n-1), N3/Grid3D.cpp:31-45 */
static void FN(orc3d_h)(int n, const double* range, REAL* hx, REAL* hy, REAL* hz)
{
    REAL xr = (REAL)range[1] - (REAL)range[0];
    REAL yr = (REAL)range[3] - (REAL)range[2];
    REAL zr = (REAL)range[5] - (REAL)range[4];
    *hx = xr / (REAL)(n - 1);
    *hy = yr / (REAL)(n - 1);
    *hz = zr / (REAL)(n - 1);
}

/* Grid3D::InitV, N3/Grid3D.cpp:61-76 (boundary v = 0); interior additionally zeroed (App. B8) */
void FN(orc3d_init_v)(REAL* v, int n)
{
    size_t tot = (size_t)n * n * n;
    for (size_t i = 0; i < tot; i++) v[i] = 0.0f;
}

/* Grid3D::InitF, N3/Grid3D.cpp:78-96: x in REAL, product in double with PI double, narrowed */
void FN(orc3d_init_f)(REAL* f, int n, const double* range)
{
    const double PI = 3.141592653589793; /* N3/inclusion.h:9 */
    REAL hx, hy, hz;
    FN(orc3d_h)(n, range, &hx, &hy, &hz);
    REAL xa = (REAL)range[0], ya = (REAL)range[2], za = (REAL)range[4];
    for (int pz = 0; pz < n; pz++)
        for (int py = 0; py < n; py++)
            for (int px = 0; px < n; px++) {
                REAL x = xa + px * hx;
                REAL y = ya + py * hy;
                REAL z = za + pz * hz;
                f[IDX3(px, py, pz)] = (REAL)(-3 * PI * PI * sin(PI * x) * sin(PI * y) * sin(PI * z));
            }
}

/* MultiGrid3D::Relax, N3/MultiGrid3D.cpp:489-567 */
void FN(orc3d_relax)(REAL* v, const REAL* f, int n, const double* range, int ncycles)
{
    REAL h_x, h_y, h_z;
    FN(orc3d_h)(n, range, &h_x, &h_y, &h_z);
    REAL h_x2 = h_x * h_x, h_y2 = h_y * h_y, h_z2 = h_z * h_z;
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++) /* :515 even first, :544 odd second */
            for (int pz = 1; pz < n - 1; pz++)
                for (int py = 1; py < n - 1; py++)
                    for (int px = 1; px < n - 1; px++) {
                        if ((py + px + pz) % 2 != colour) continue;
                        REAL O = v[IDX3(px - 1, py, pz)];
                        REAL E = v[IDX3(px + 1, py, pz)];
                        REAL N = v[IDX3(px, py - 1, pz)];
                        REAL S = v[IDX3(px, py + 1, pz)];
                        REAL D = v[IDX3(px, py, pz - 1)];
                        REAL U = v[IDX3(px, py, pz + 1)];
                        size_t idx = IDX3(px, py, pz);
                        /* :532 verbatim */
                        v[idx] = (O * (h_y2 * h_z2) + E * (h_y2 * h_z2) + N * (h_x2 * h_z2) + S * (h_x2 * h_z2) +
                                  D * (h_x2 * h_y2) + U * (h_x2 * h_y2) - f[idx] * h_x2 * h_y2 * h_z2) /
                                 (2 * (h_y2 * h_z2 + h_x2 * h_z2 + h_x2 * h_y2));
                    }
}

/* Weighted Jacobi.  NOT in the reference (its only smoother is the red-black Gauss-Seidel above; SURVEY.md 8f
   rank 4): this function is the definition the engine's MG_SMOOTHER_JACOBI is tested against -- parity with the
   reference is unpinned by construction.  One sweep: every interior point, from the OLD values of its six
   neighbours,  v_new = v_old + omega * (gs - v_old)  with gs the expression of N3/MultiGrid3D.cpp:532. */
void FN(orc3d_relax_jacobi)(REAL* v, const REAL* f, int n, const double* range, double omega_d, int ncycles)
{
    REAL h_x, h_y, h_z;
    FN(orc3d_h)(n, range, &h_x, &h_y, &h_z);
    REAL h_x2 = h_x * h_x, h_y2 = h_y * h_y, h_z2 = h_z * h_z;
    const REAL omega = (REAL)omega_d;
    size_t tot = (size_t)n * n * n;
    REAL* nv = (REAL*)malloc(tot * sizeof(REAL));
    for (int k = 0; k < ncycles; k++) {
        for (size_t i = 0; i < tot; i++) nv[i] = v[i];
        for (int pz = 1; pz < n - 1; pz++)
            for (int py = 1; py < n - 1; py++)
                for (int px = 1; px < n - 1; px++) {
                    REAL O = v[IDX3(px - 1, py, pz)];
                    REAL E = v[IDX3(px + 1, py, pz)];
                    REAL N = v[IDX3(px, py - 1, pz)];
                    REAL S = v[IDX3(px, py + 1, pz)];
                    REAL D = v[IDX3(px, py, pz - 1)];
                    REAL U = v[IDX3(px, py, pz + 1)];
                    size_t idx = IDX3(px, py, pz);
                    REAL gs = (O * (h_y2 * h_z2) + E * (h_y2 * h_z2) + N * (h_x2 * h_z2) + S * (h_x2 * h_z2) +
                               D * (h_x2 * h_y2) + U * (h_x2 * h_y2) - f[idx] * h_x2 * h_y2 * h_z2) /
                              (2 * (h_y2 * h_z2 + h_x2 * h_z2 + h_x2 * h_y2));
                    REAL d = gs - v[idx];
                    REAL u = omega * d;
                    nv[idx] = v[idx] + u;
                }
        for (size_t i = 0; i < tot; i++) v[i] = nv[i];
    }
    free(nv);
}

/* MultiGrid3D::CalculateResidual, N3/MultiGrid3D.cpp:678-730; corrected != 0 flips the two
   wrong signs of :723 */
void FN(orc3d_residual)(const REAL* v, const REAL* f, REAL* r, int n, const double* range, int corrected)
{
    REAL h_x, h_y, h_z;
    FN(orc3d_h)(n, range, &h_x, &h_y, &h_z);
    REAL h_x2 = h_x * h_x, h_y2 = h_y * h_y, h_z2 = h_z * h_z;
    for (int pz = 0; pz < n; pz++)
        for (int py = 0; py < n; py++)
            for (int px = 0; px < n; px++) {
                size_t idx = IDX3(px, py, pz);
                if (px == 0 || px == n - 1 || py == 0 || py == n - 1 || pz == 0 || pz == n - 1) {
                    r[idx] = 0.0f;
                    continue;
                }
                REAL O = v[IDX3(px - 1, py, pz)];
                REAL E = v[IDX3(px + 1, py, pz)];
                REAL N = v[IDX3(px, py - 1, pz)];
                REAL S = v[IDX3(px, py + 1, pz)];
                REAL D = v[IDX3(px, py, pz - 1)];
                REAL U = v[IDX3(px, py, pz + 1)];
                if (corrected)
                    r[idx] = f[idx] - ((O - 2 * v[idx] + E) / h_x2) - ((N - 2 * v[idx] + S) / h_y2) -
                             ((D - 2 * v[idx] + U) / h_z2);
                else /* :723 verbatim */
                    r[idx] = f[idx] - ((O - 2 * v[idx] + E) / h_x2) - ((N - 2 * v[idx] - S) / h_y2) -
                             ((D - 2 * v[idx] - U) / h_z2);
            }
}

/* MultiGrid3D::Restrict, N3/MultiGrid3D.cpp:50-184.  Reference names: suffix _C = y, _N = y-1,
   _S = y+1; N = z+1, S = z-1, E = x+1, O = x-1. */
void FN(orc3d_restrict)(const REAL* fine, int n, REAL* coarse)
{
    int cn = (n - 1) / 2 + 1;
#define F3(dx, dy, dz) fine[IDX3(fx + (dx), fy + (dy), fz + (dz))]
    for (int cz = 0; cz < cn; cz++)
        for (int cy = 0; cy < cn; cy++)
            for (int cx = 0; cx < cn; cx++) {
                int fx = 2 * cx, fy = 2 * cy, fz = 2 * cz;
                size_t cidx = (size_t)cx + (size_t)cy * cn + (size_t)cz * cn * cn;
                if (cx == 0 || cx == cn - 1 || cy == 0 || cy == cn - 1 || cz == 0 || cz == cn - 1) {
                    coarse[cidx] = F3(0, 0, 0); /* :113-119 injection on the boundary */
                    continue;
                }
                REAL C_C = F3(0, 0, 0), N_C = F3(0, 0, 1), S_C = F3(0, 0, -1), E_C = F3(1, 0, 0), O_C = F3(-1, 0, 0);
                REAL NE_C = F3(1, 0, 1), NO_C = F3(-1, 0, 1), SE_C = F3(1, 0, -1), SO_C = F3(-1, 0, -1);
                REAL C_N = F3(0, -1, 0), N_N = F3(0, -1, 1), S_N = F3(0, -1, -1), E_N = F3(1, -1, 0), O_N = F3(-1, -1, 0);
                REAL NE_N = F3(1, -1, 1), NO_N = F3(-1, -1, 1), SE_N = F3(1, -1, -1), SO_N = F3(-1, -1, -1);
                REAL C_S = F3(0, 1, 0), N_S = F3(0, 1, 1), S_S = F3(0, 1, -1), E_S = F3(1, 1, 0), O_S = F3(-1, 1, 0);
                REAL NE_S = F3(1, 1, 1), NO_S = F3(-1, 1, 1), SE_S = F3(1, 1, -1), SO_S = F3(-1, 1, -1);
                /* :180 verbatim */
                coarse[cidx] = (1 / 8.0f) * (C_C) + (1 / 16.0f) * ((N_C + E_C + S_C + O_C) + (C_N + C_S)) +
                               (1 / 32.0f) * ((NE_C + SE_C + SO_C + NO_C) + (N_N + E_N + S_N + O_N) +
                                              (N_S + E_S + S_S + O_S)) +
                               (1 / 64.0f) * ((NE_N + SE_N + SO_N + NO_N) + (NE_S + SE_S + SO_S + NO_S));
            }
#undef F3
}

/* MultiGrid3D::Interpolate, N3/MultiGrid3D.cpp:186-335; interior of fine only */
void FN(orc3d_interpolate)(REAL* fine, int n, const REAL* coarse)
{
    int cn = (n - 1) / 2 + 1;
#define C3(dx, dy, dz) coarse[(size_t)(cx + (dx)) + (size_t)(cy + (dy)) * cn + (size_t)(cz + (dz)) * cn * cn]
    for (int fz = 1; fz < n - 1; fz++)
        for (int fy = 1; fy < n - 1; fy++)
            for (int fx = 1; fx < n - 1; fx++) {
                int cx = fx / 2, cy = fy / 2, cz = fz / 2;
                size_t fidx = IDX3(fx, fy, fz);
                int oy = fy % 2, ox = fx % 2, oz = fz % 2;
                if (!oy && !ox && !oz) fine[fidx] = C3(0, 0, 0);                                  /* PPP :216 */
                else if (!oy && ox && !oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(1, 0, 0)); /* PDP :222 */
                else if (oy && !ox && !oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(0, 1, 0)); /* DPP :233 */
                else if (oy && ox && !oz)                                                        /* DDP :244 */
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 0) + C3(1, 0, 0) + C3(0, 1, 0) + C3(1, 1, 0));
                else if (!oy && !ox && oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(0, 0, 1)); /* PPD :261 */
                else if (!oy && ox && oz)                                                        /* PDD :272 */
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 1) + C3(1, 0, 1) + C3(0, 0, 0) + C3(1, 0, 0));
                else if (oy && !ox && oz)                                                        /* DPD :287 */
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 0) + C3(0, 0, 1) + C3(0, 1, 0) + C3(0, 1, 1));
                else                                                                             /* DDD :302 */
                    fine[fidx] = (1 / 8.0f) * (C3(0, 0, 0) + C3(0, 0, 1) + C3(1, 0, 1) + C3(1, 0, 0) + C3(0, 1, 0) +
                                               C3(0, 1, 1) + C3(1, 1, 1) + C3(1, 1, 0));
            }
#undef C3
}

/* MultiGrid3D::ApplyCorrection, N3/MultiGrid3D.cpp:649-676 */
void FN(orc3d_apply_correction)(REAL* fine, const REAL* err, int n)
{
    for (int pz = 1; pz < n - 1; pz++)
        for (int py = 1; py < n - 1; py++)
            for (int px = 1; px < n - 1; px++) {
                size_t idx = IDX3(px, py, pz);
                fine[idx] = fine[idx] + err[idx];
            }
}

/* MultiGrid3D::setToValue, N3/MultiGrid3D.cpp:587-621 */
void FN(orc3d_set)(REAL* g, int n, double value, int modify_boundaries)
{
    int lo = modify_boundaries ? 0 : 1, hi = modify_boundaries ? n : n - 1;
    for (int pz = lo; pz < hi; pz++)
        for (int py = lo; py < hi; py++)
            for (int px = lo; px < hi; px++) g[IDX3(px, py, pz)] = (REAL)value;
}
#undef IDX3

/* level arrays v[l], f[l] of size n_l^3, n_{l+1} = (n_l-1)/2+1.
   MultiGrid3D::VCycle, N3/MultiGrid3D.cpp:623-647 */
void FN(orc3d_vcycle)(REAL** v, REAL** f, int n0, int nlevels, const double* range, int level, int v1, int v2,
                      int corrected)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    FN(orc3d_relax)(v[level], f[level], n, range, v1);
    if (level != nlevels - 1) {
        size_t tot = (size_t)n * n * n;
        REAL* tmp = (REAL*)malloc(tot * sizeof(REAL));
        FN(orc3d_residual)(v[level], f[level], tmp, n, range, corrected);
        FN(orc3d_restrict)(tmp, n, f[level + 1]);
        int cn = (n - 1) / 2 + 1;
        FN(orc3d_set)(v[level + 1], cn, 0.0, 1); /* :634 boundary included */
        FN(orc3d_vcycle)(v, f, n0, nlevels, range, level + 1, v1, v2, corrected);
        /* :638 the reference's fine_error is an uninitialised malloc whose boundary is never read */
        FN(orc3d_interpolate)(tmp, n, v[level + 1]);
        FN(orc3d_apply_correction)(v[level], tmp, n);
        free(tmp);
    }
    FN(orc3d_relax)(v[level], f[level], n, range, v2);
}

/* the same V-cycle with the weighted-Jacobi smoother in place of Relax (no reference counterpart, see above) */
void FN(orc3d_vcycle_jacobi)(REAL** v, REAL** f, int n0, int nlevels, const double* range, int level, int v1, int v2,
                             int corrected, double omega)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    FN(orc3d_relax_jacobi)(v[level], f[level], n, range, omega, v1);
    if (level != nlevels - 1) {
        size_t tot = (size_t)n * n * n;
        REAL* tmp = (REAL*)malloc(tot * sizeof(REAL));
        FN(orc3d_residual)(v[level], f[level], tmp, n, range, corrected);
        FN(orc3d_restrict)(tmp, n, f[level + 1]);
        int cn = (n - 1) / 2 + 1;
        FN(orc3d_set)(v[level + 1], cn, 0.0, 1);
        FN(orc3d_vcycle_jacobi)(v, f, n0, nlevels, range, level + 1, v1, v2, corrected, omega);
        FN(orc3d_interpolate)(tmp, n, v[level + 1]);
        FN(orc3d_apply_correction)(v[level], tmp, n);
        free(tmp);
    }
    FN(orc3d_relax_jacobi)(v[level], f[level], n, range, omega, v2);
}

/* MultiGrid3D::FullMultiGridVCycle, N3/MultiGrid3D.cpp:569-585 */
void FN(orc3d_fmg)(REAL** v, REAL** f, int n0, int nlevels, const double* range, int level, int v0, int v1, int v2,
                   int corrected)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    if (level != nlevels - 1) {
        FN(orc3d_restrict)(f[level], n, f[level + 1]);
        FN(orc3d_fmg)(v, f, n0, nlevels, range, level + 1, v0, v1, v2, corrected);
        FN(orc3d_interpolate)(v[level], n, v[level + 1]);
    } else {
        FN(orc3d_set)(v[level], n, 0.0, 0);
    }
    for (int i = 0; i < v0; i++) FN(orc3d_vcycle)(v, f, n0, nlevels, range, level, v1, v2, corrected);
}

/* ------------------------------------------------------------------------------------------ */
/* 3D Poisson on a NON-CUBIC grid (sizeX != sizeY != sizeZ, every one 2^k + 1)                  */
/*                                                                                            */
/* The reference asserts such grids away (N3/Grid3D.cpp:10-11, the author's TODO) although its */
/* hierarchy (N3/MultiGrid3D.cpp:19-47: numGrids from the SMALLEST dimension, every dimension  */
/* halved per level) and all its operators are written per dimension.  These functions restate */
/* the same operators with (nx, ny, nz); pinned against the reference compiled with -DNDEBUG   */
/* (oracle/_ref, the ref3d_*x variants) by tests/test_oracle.py.                               */
/* ------------------------------------------------------------------------------------------ */

#define IDXB(x, y, z) ((size_t)(x) + (size_t)(y) * (size_t)nx + (size_t)(z) * (size_t)nx * (size_t)ny)

static void FN(orc3b_h)(int nx, int ny, int nz, const double* range, REAL* hx, REAL* hy, REAL* hz)
{
    REAL xr = (REAL)range[1] - (REAL)range[0];
    REAL yr = (REAL)range[3] - (REAL)range[2];
    REAL zr = (REAL)range[5] - (REAL)range[4];
    *hx = xr / (REAL)(nx - 1);
    *hy = yr / (REAL)(ny - 1);
    *hz = zr / (REAL)(nz - 1);
}

void FN(orc3b_init_v)(REAL* v, int nx, int ny, int nz)
{
    size_t tot = (size_t)nx * ny * nz;
    for (size_t i = 0; i < tot; i++) v[i] = 0.0f;
}

void FN(orc3b_init_f)(REAL* f, int nx, int ny, int nz, const double* range)
{
    const double PI = 3.141592653589793;
    REAL hx, hy, hz;
    FN(orc3b_h)(nx, ny, nz, range, &hx, &hy, &hz);
    REAL xa = (REAL)range[0], ya = (REAL)range[2], za = (REAL)range[4];
    for (int pz = 0; pz < nz; pz++)
        for (int py = 0; py < ny; py++)
            for (int px = 0; px < nx; px++) {
                REAL x = xa + px * hx;
                REAL y = ya + py * hy;
                REAL z = za + pz * hz;
                f[IDXB(px, py, pz)] = (REAL)(-3 * PI * PI * sin(PI * x) * sin(PI * y) * sin(PI * z));
            }
}

void FN(orc3b_relax)(REAL* v, const REAL* f, int nx, int ny, int nz, const double* range, int ncycles)
{
    REAL h_x, h_y, h_z;
    FN(orc3b_h)(nx, ny, nz, range, &h_x, &h_y, &h_z);
    REAL h_x2 = h_x * h_x, h_y2 = h_y * h_y, h_z2 = h_z * h_z;
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++)
            for (int pz = 1; pz < nz - 1; pz++)
                for (int py = 1; py < ny - 1; py++)
                    for (int px = 1; px < nx - 1; px++) {
                        if ((py + px + pz) % 2 != colour) continue;
                        REAL O = v[IDXB(px - 1, py, pz)];
                        REAL E = v[IDXB(px + 1, py, pz)];
                        REAL N = v[IDXB(px, py - 1, pz)];
                        REAL S = v[IDXB(px, py + 1, pz)];
                        REAL D = v[IDXB(px, py, pz - 1)];
                        REAL U = v[IDXB(px, py, pz + 1)];
                        size_t idx = IDXB(px, py, pz);
                        v[idx] = (O * (h_y2 * h_z2) + E * (h_y2 * h_z2) + N * (h_x2 * h_z2) + S * (h_x2 * h_z2) +
                                  D * (h_x2 * h_y2) + U * (h_x2 * h_y2) - f[idx] * h_x2 * h_y2 * h_z2) /
                                 (2 * (h_y2 * h_z2 + h_x2 * h_z2 + h_x2 * h_y2));
                    }
}

void FN(orc3b_residual)(const REAL* v, const REAL* f, REAL* r, int nx, int ny, int nz, const double* range, int corrected)
{
    REAL h_x, h_y, h_z;
    FN(orc3b_h)(nx, ny, nz, range, &h_x, &h_y, &h_z);
    REAL h_x2 = h_x * h_x, h_y2 = h_y * h_y, h_z2 = h_z * h_z;
    for (int pz = 0; pz < nz; pz++)
        for (int py = 0; py < ny; py++)
            for (int px = 0; px < nx; px++) {
                size_t idx = IDXB(px, py, pz);
                if (px == 0 || px == nx - 1 || py == 0 || py == ny - 1 || pz == 0 || pz == nz - 1) {
                    r[idx] = 0.0f;
                    continue;
                }
                REAL O = v[IDXB(px - 1, py, pz)];
                REAL E = v[IDXB(px + 1, py, pz)];
                REAL N = v[IDXB(px, py - 1, pz)];
                REAL S = v[IDXB(px, py + 1, pz)];
                REAL D = v[IDXB(px, py, pz - 1)];
                REAL U = v[IDXB(px, py, pz + 1)];
                if (corrected)
                    r[idx] = f[idx] - ((O - 2 * v[idx] + E) / h_x2) - ((N - 2 * v[idx] + S) / h_y2) -
                             ((D - 2 * v[idx] + U) / h_z2);
                else
                    r[idx] = f[idx] - ((O - 2 * v[idx] + E) / h_x2) - ((N - 2 * v[idx] - S) / h_y2) -
                             ((D - 2 * v[idx] - U) / h_z2);
            }
}

void FN(orc3b_restrict)(const REAL* fine, int nx, int ny, int nz, REAL* coarse)
{
    int cnx = (nx - 1) / 2 + 1, cny = (ny - 1) / 2 + 1, cnz = (nz - 1) / 2 + 1;
#define F3(dx, dy, dz) fine[IDXB(fx + (dx), fy + (dy), fz + (dz))]
    for (int cz = 0; cz < cnz; cz++)
        for (int cy = 0; cy < cny; cy++)
            for (int cx = 0; cx < cnx; cx++) {
                int fx = 2 * cx, fy = 2 * cy, fz = 2 * cz;
                size_t cidx = (size_t)cx + (size_t)cy * cnx + (size_t)cz * cnx * cny;
                if (cx == 0 || cx == cnx - 1 || cy == 0 || cy == cny - 1 || cz == 0 || cz == cnz - 1) {
                    coarse[cidx] = F3(0, 0, 0);
                    continue;
                }
                REAL C_C = F3(0, 0, 0), N_C = F3(0, 0, 1), S_C = F3(0, 0, -1), E_C = F3(1, 0, 0), O_C = F3(-1, 0, 0);
                REAL NE_C = F3(1, 0, 1), NO_C = F3(-1, 0, 1), SE_C = F3(1, 0, -1), SO_C = F3(-1, 0, -1);
                REAL C_N = F3(0, -1, 0), N_N = F3(0, -1, 1), S_N = F3(0, -1, -1), E_N = F3(1, -1, 0), O_N = F3(-1, -1, 0);
                REAL NE_N = F3(1, -1, 1), NO_N = F3(-1, -1, 1), SE_N = F3(1, -1, -1), SO_N = F3(-1, -1, -1);
                REAL C_S = F3(0, 1, 0), N_S = F3(0, 1, 1), S_S = F3(0, 1, -1), E_S = F3(1, 1, 0), O_S = F3(-1, 1, 0);
                REAL NE_S = F3(1, 1, 1), NO_S = F3(-1, 1, 1), SE_S = F3(1, 1, -1), SO_S = F3(-1, 1, -1);
                coarse[cidx] = (1 / 8.0f) * (C_C) + (1 / 16.0f) * ((N_C + E_C + S_C + O_C) + (C_N + C_S)) +
                               (1 / 32.0f) * ((NE_C + SE_C + SO_C + NO_C) + (N_N + E_N + S_N + O_N) +
                                              (N_S + E_S + S_S + O_S)) +
                               (1 / 64.0f) * ((NE_N + SE_N + SO_N + NO_N) + (NE_S + SE_S + SO_S + NO_S));
            }
#undef F3
}

void FN(orc3b_interpolate)(REAL* fine, int nx, int ny, int nz, const REAL* coarse)
{
    int cnx = (nx - 1) / 2 + 1, cny = (ny - 1) / 2 + 1;
#define C3(dx, dy, dz) coarse[(size_t)(cx + (dx)) + (size_t)(cy + (dy)) * cnx + (size_t)(cz + (dz)) * cnx * cny]
    for (int fz = 1; fz < nz - 1; fz++)
        for (int fy = 1; fy < ny - 1; fy++)
            for (int fx = 1; fx < nx - 1; fx++) {
                int cx = fx / 2, cy = fy / 2, cz = fz / 2;
                size_t fidx = IDXB(fx, fy, fz);
                int oy = fy % 2, ox = fx % 2, oz = fz % 2;
                if (!oy && !ox && !oz) fine[fidx] = C3(0, 0, 0);
                else if (!oy && ox && !oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(1, 0, 0));
                else if (oy && !ox && !oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(0, 1, 0));
                else if (oy && ox && !oz)
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 0) + C3(1, 0, 0) + C3(0, 1, 0) + C3(1, 1, 0));
                else if (!oy && !ox && oz) fine[fidx] = (1 / 2.0f) * (C3(0, 0, 0) + C3(0, 0, 1));
                else if (!oy && ox && oz)
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 1) + C3(1, 0, 1) + C3(0, 0, 0) + C3(1, 0, 0));
                else if (oy && !ox && oz)
                    fine[fidx] = (1 / 4.0f) * (C3(0, 0, 0) + C3(0, 0, 1) + C3(0, 1, 0) + C3(0, 1, 1));
                else
                    fine[fidx] = (1 / 8.0f) * (C3(0, 0, 0) + C3(0, 0, 1) + C3(1, 0, 1) + C3(1, 0, 0) + C3(0, 1, 0) +
                                               C3(0, 1, 1) + C3(1, 1, 1) + C3(1, 1, 0));
            }
#undef C3
}

void FN(orc3b_apply_correction)(REAL* fine, const REAL* err, int nx, int ny, int nz)
{
    for (int pz = 1; pz < nz - 1; pz++)
        for (int py = 1; py < ny - 1; py++)
            for (int px = 1; px < nx - 1; px++) {
                size_t idx = IDXB(px, py, pz);
                fine[idx] = fine[idx] + err[idx];
            }
}

void FN(orc3b_set)(REAL* g, int nx, int ny, int nz, double value, int modify_boundaries)
{
    int lo = modify_boundaries ? 0 : 1, e = modify_boundaries ? 0 : 1;
    for (int pz = lo; pz < nz - e; pz++)
        for (int py = lo; py < ny - e; py++)
            for (int px = lo; px < nx - e; px++) g[IDXB(px, py, pz)] = (REAL)value;
}
#undef IDXB

static void FN(orc3b_level)(const int* n0, int level, int* nx, int* ny, int* nz)
{
    *nx = n0[0]; *ny = n0[1]; *nz = n0[2];
    for (int l = 0; l < level; l++) {
        *nx = (*nx - 1) / 2 + 1; *ny = (*ny - 1) / 2 + 1; *nz = (*nz - 1) / 2 + 1;
    }
}

/* MultiGrid3D::VCycle, N3/MultiGrid3D.cpp:623-647, on the per-dimension hierarchy of :19-47 */
void FN(orc3b_vcycle)(REAL** v, REAL** f, const int* n0, int nlevels, const double* range, int level, int v1, int v2,
                      int corrected)
{
    int nx, ny, nz;
    FN(orc3b_level)(n0, level, &nx, &ny, &nz);
    FN(orc3b_relax)(v[level], f[level], nx, ny, nz, range, v1);
    if (level != nlevels - 1) {
        size_t tot = (size_t)nx * ny * nz;
        REAL* tmp = (REAL*)malloc(tot * sizeof(REAL));
        FN(orc3b_residual)(v[level], f[level], tmp, nx, ny, nz, range, corrected);
        FN(orc3b_restrict)(tmp, nx, ny, nz, f[level + 1]);
        FN(orc3b_set)(v[level + 1], (nx - 1) / 2 + 1, (ny - 1) / 2 + 1, (nz - 1) / 2 + 1, 0.0, 1);
        FN(orc3b_vcycle)(v, f, n0, nlevels, range, level + 1, v1, v2, corrected);
        FN(orc3b_interpolate)(tmp, nx, ny, nz, v[level + 1]);
        FN(orc3b_apply_correction)(v[level], tmp, nx, ny, nz);
        free(tmp);
    }
    FN(orc3b_relax)(v[level], f[level], nx, ny, nz, range, v2);
}

/* MultiGrid3D::FullMultiGridVCycle, N3/MultiGrid3D.cpp:569-585 */
void FN(orc3b_fmg)(REAL** v, REAL** f, const int* n0, int nlevels, const double* range, int level, int v0, int v1, int v2,
                   int corrected)
{
    int nx, ny, nz;
    FN(orc3b_level)(n0, level, &nx, &ny, &nz);
    if (level != nlevels - 1) {
        FN(orc3b_restrict)(f[level], nx, ny, nz, f[level + 1]);
        FN(orc3b_fmg)(v, f, n0, nlevels, range, level + 1, v0, v1, v2, corrected);
        FN(orc3b_interpolate)(v[level], nx, ny, nz, v[level + 1]);
    } else {
        FN(orc3b_set)(v[level], nx, ny, nz, 0.0, 0);
    }
    for (int i = 0; i < v0; i++) FN(orc3b_vcycle)(v, f, n0, nlevels, range, level, v1, v2, corrected);
}

/* ------------------------------------------------------------------------------------------ */
/* 2D Lyapunov                                                                                 */
/* ------------------------------------------------------------------------------------------ */

#define IDX2(x, y) ((size_t)(x) + (size_t)(y) * (size_t)n)

static void FN(orc2d_h)(int n, const double* range, REAL* hx, REAL* hy)
{
    REAL xr = (REAL)range[1] - (REAL)range[0];
    REAL yr = (REAL)range[3] - (REAL)range[2];
    *hx = xr / (REAL)(n - 1); /* N2/Grid2D.cpp:34-35 */
    *hy = yr / (REAL)(n - 1);
}

/* Grid2D::InitV, N2/Grid2D.cpp:50-68 */
void FN(orc2d_init_v)(REAL* v, int n, const double* range)
{
    REAL h_x, h_y;
    FN(orc2d_h)(n, range, &h_x, &h_y);
    REAL x_a = (REAL)range[0], y_a = (REAL)range[2];
    for (int py = 0; py < n; py++)
        for (int px = 0; px < n; px++) {
            if (px == 0 || px == n - 1 || py == 0 || py == n - 1) {
                REAL yi = y_a + py * h_y;
                REAL xj = x_a + px * h_x;
                REAL sol = 2 * xj * xj - 4 * xj * yi + 2 * yi * yi;
                v[IDX2(px, py)] = sol;
            } else
                v[IDX2(px, py)] = 0.0f;
        }
}

/* Grid2D::InitF, N2/Grid2D.cpp:70-80 */
void FN(orc2d_init_f)(REAL* f, int n)
{
    for (size_t i = 0; i < (size_t)n * n; i++) f[i] = 0.0f;
}

/* MultiGrid2D::Relax, N2/MultiGrid2D.cpp:199-273; A = matrixA[0..3], alfa is int */
void FN(orc2d_relax)(REAL* h_v, const REAL* f, int n, const double* range, const double* A4, int alfa, int ncycles)
{
    REAL h_x, h_y;
    FN(orc2d_h)(n, range, &h_x, &h_y);
    REAL x_a = (REAL)range[0], y_a = (REAL)range[2];
    REAL matrixA[4] = {(REAL)A4[0], (REAL)A4[1], (REAL)A4[2], (REAL)A4[3]};
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++)
            for (int py = 1; py < n - 1; py++)
                for (int px = 1; px < n - 1; px++) {
                    if ((py + px) % 2 != colour) continue;
                    size_t idx = IDX2(px, py);
                    REAL xj = x_a + px * h_x;
                    REAL yi = y_a + py * h_y;
                    REAL K1 = matrixA[0] * xj + matrixA[1] * yi;
                    REAL K2 = matrixA[2] * xj + matrixA[3] * yi;
                    REAL den = K1 * h_y + K2 * h_x - alfa * h_x * h_y;
                    size_t idxVarX = IDX2(px + 1, py);
                    size_t idxVarY = IDX2(px, py + 1);
                    /* :241 verbatim */
                    h_v[idx] = (h_y * K1 * h_v[idxVarX] + h_x * K2 * h_v[idxVarY] - f[idx] * h_x * h_y) / (den);
                }
}

/* MultiGrid2D::CalculateResidual, N2/MultiGrid2D.cpp:367-408 */
void FN(orc2d_residual)(const REAL* h_v, const REAL* h_f, REAL* r, int n, const double* range, const double* A4,
                        int alfa)
{
    REAL h_x, h_y;
    FN(orc2d_h)(n, range, &h_x, &h_y);
    REAL x_a = (REAL)range[0], y_a = (REAL)range[2];
    REAL matrixA[4] = {(REAL)A4[0], (REAL)A4[1], (REAL)A4[2], (REAL)A4[3]};
    for (int py = 0; py < n; py++)
        for (int px = 0; px < n; px++) {
            size_t idx = IDX2(px, py);
            if (px == 0 || px == n - 1 || py == 0 || py == n - 1) {
                r[idx] = 0.0f;
                continue;
            }
            REAL xj = x_a + px * h_x;
            REAL yi = y_a + py * h_y;
            REAL K1 = matrixA[0] * xj + matrixA[1] * yi;
            REAL K2 = matrixA[2] * xj + matrixA[3] * yi;
            size_t idxVarX = IDX2(px + 1, py);
            size_t idxVarY = IDX2(px, py + 1);
            /* :403 verbatim */
            r[idx] = h_f[idx] - (h_y * K1 * h_v[idxVarX] + h_x * K2 * h_v[idxVarY] -
                                 h_v[idx] * (h_y * K1 + h_x * K2 - alfa * h_x * h_y)) /
                                    (h_x * h_y);
        }
}

/* MultiGrid2D::Restrict, N2/MultiGrid2D.cpp:63-126; N = y-1, S = y+1, E = x+1, O = x-1 */
void FN(orc2d_restrict)(const REAL* fine, int n, REAL* coarse)
{
    int cn = (n - 1) / 2 + 1;
    for (int cy = 0; cy < cn; cy++)
        for (int cx = 0; cx < cn; cx++) {
            int fx = 2 * cx, fy = 2 * cy;
            size_t cidx = (size_t)cx + (size_t)cy * cn;
            if (cx == 0 || cx == cn - 1 || cy == 0 || cy == cn - 1) {
                coarse[cidx] = fine[IDX2(fx, fy)];
                continue;
            }
            REAL C = fine[IDX2(fx, fy)], N = fine[IDX2(fx, fy - 1)], S = fine[IDX2(fx, fy + 1)];
            REAL E = fine[IDX2(fx + 1, fy)], O = fine[IDX2(fx - 1, fy)];
            REAL NE = fine[IDX2(fx + 1, fy - 1)], NO = fine[IDX2(fx - 1, fy - 1)];
            REAL SE = fine[IDX2(fx + 1, fy + 1)], SO = fine[IDX2(fx - 1, fy + 1)];
            coarse[cidx] = (1 / 16.0f) * (NO + NE + SO + SE + 2 * (O + E + N + S) + 4 * C); /* :123 */
        }
}

/* MultiGrid2D::Interpolate, N2/MultiGrid2D.cpp:128-196 */
void FN(orc2d_interpolate)(REAL* fine, int n, const REAL* coarse)
{
    int cn = (n - 1) / 2 + 1;
    for (int fy = 1; fy < n - 1; fy++)
        for (int fx = 1; fx < n - 1; fx++) {
            int cx = fx / 2, cy = fy / 2;
            size_t fidx = IDX2(fx, fy);
            size_t c00 = (size_t)cx + (size_t)cy * cn;
            if (fy % 2 == 0 && fx % 2 == 0) fine[fidx] = coarse[c00];
            else if (fy % 2 != 0 && fx % 2 == 0) fine[fidx] = (1 / 2.0f) * (coarse[c00] + coarse[c00 + cn]);
            else if (fy % 2 == 0 && fx % 2 != 0) fine[fidx] = (1 / 2.0f) * (coarse[c00] + coarse[c00 + 1]);
            else fine[fidx] = (1 / 4.0f) * (coarse[c00] + coarse[c00 + 1] + coarse[c00 + cn] + coarse[c00 + cn + 1]);
        }
}

/* MultiGrid2D::ApplyCorrection, N2/MultiGrid2D.cpp:343-366 */
void FN(orc2d_apply_correction)(REAL* fine, const REAL* err, int n)
{
    for (int py = 1; py < n - 1; py++)
        for (int px = 1; px < n - 1; px++) fine[IDX2(px, py)] = fine[IDX2(px, py)] + err[IDX2(px, py)];
}

/* MultiGrid2D::setToValue, N2/MultiGrid2D.cpp:275-292 */
void FN(orc2d_set)(REAL* g, int n, double value, int modify_boundaries)
{
    int lo = modify_boundaries ? 0 : 1, hi = modify_boundaries ? n : n - 1;
    for (int py = lo; py < hi; py++)
        for (int px = lo; px < hi; px++) g[IDX2(px, py)] = (REAL)value;
}
#undef IDX2

/* MultiGrid2D::VCycle, N2/MultiGrid2D.cpp:314-340 */
void FN(orc2d_vcycle)(REAL** v, REAL** f, int n0, int nlevels, const double* range, const double* A4, int alfa,
                      int level, int v1, int v2)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    FN(orc2d_relax)(v[level], f[level], n, range, A4, alfa, v1);
    if (level != nlevels - 1) {
        REAL* tmp = (REAL*)malloc((size_t)n * n * sizeof(REAL));
        FN(orc2d_residual)(v[level], f[level], tmp, n, range, A4, alfa);
        FN(orc2d_restrict)(tmp, n, f[level + 1]);
        FN(orc2d_set)(v[level + 1], (n - 1) / 2 + 1, 0.0, 1);
        FN(orc2d_vcycle)(v, f, n0, nlevels, range, A4, alfa, level + 1, v1, v2);
        FN(orc2d_interpolate)(tmp, n, v[level + 1]);
        FN(orc2d_apply_correction)(v[level], tmp, n);
        free(tmp);
    }
    FN(orc2d_relax)(v[level], f[level], n, range, A4, alfa, v2);
}

/* MultiGrid2D::FullMultiGridVCycle, N2/MultiGrid2D.cpp:296-312 */
void FN(orc2d_fmg)(REAL** v, REAL** f, int n0, int nlevels, const double* range, const double* A4, int alfa,
                   int level, int v0, int v1, int v2)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    if (level != nlevels - 1) {
        FN(orc2d_restrict)(f[level], n, f[level + 1]);
        FN(orc2d_fmg)(v, f, n0, nlevels, range, A4, alfa, level + 1, v0, v1, v2);
        FN(orc2d_interpolate)(v[level], n, v[level + 1]);
    } else {
        FN(orc2d_set)(v[level], n, 0.0, 0);
    }
    for (int i = 0; i < v0; i++) FN(orc2d_vcycle)(v, f, n0, nlevels, range, A4, alfa, level, v1, v2);
}

/* ------------------------------------------------------------------------------------------ */
/* 1D first-order ODE                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* exp on REAL: the reference calls exp(float) from C++ <math.h>, which resolves to the float
   overload (expf) in the fp32 build and to exp(double) under the fp64 wrapper */
#define ORC_EXP(x) FN(orc_exp)(x)

static REAL FN(orc1d_h)(int n, const double* range)
{
    REAL xr = (REAL)range[1] - (REAL)range[0];
    return xr / (REAL)(n - 1); /* N1/Grid1D.cpp:15 */
}

/* Grid1D::InitV, N1/Grid1D.cpp:30-34 (+ zero interior, App. B8) */
void FN(orc1d_init_v)(REAL* v, int n, const double* range)
{
    REAL x_a = (REAL)range[0], x_b = (REAL)range[1];
    for (int i = 0; i < n; i++) v[i] = 0.0f;
    v[0] = (ORC_EXP(x_a) + x_a - 3) / (1 + ORC_EXP(-x_a));
    v[n - 1] = (ORC_EXP(x_b) + x_b - 3) / (1 + ORC_EXP(-x_b));
}

/* Grid1D::InitF, N1/Grid1D.cpp:36-43 */
void FN(orc1d_init_f)(REAL* f, int n, const double* range)
{
    REAL h_x = FN(orc1d_h)(n, range), x_a = (REAL)range[0];
    for (int px = 0; px < n; px++) {
        REAL xj = x_a + px * h_x;
        f[px] = ORC_EXP(xj);
    }
}

/* MultiGrid1D::Relax, N1/MultiGrid1D.cpp:79-118 */
void FN(orc1d_relax)(REAL* h_v, const REAL* h_f, int n, const double* range, int ncycles)
{
    REAL h_x = FN(orc1d_h)(n, range), x_a = (REAL)range[0];
    for (int k = 0; k < ncycles; k++)
        for (int colour = 0; colour < 2; colour++)
            for (int px = 1; px < n - 1; px++) {
                if (px % 2 != colour) continue;
                REAL xj = x_a + px * h_x;
                /* :101 verbatim */
                h_v[px] = (h_v[px + 1] * (ORC_EXP(xj) + 1) - h_f[px] * h_x * (ORC_EXP(xj) + 1)) / (ORC_EXP(xj) + 1 + h_x);
            }
}

/* MultiGrid1D::CalculateResidual, N1/MultiGrid1D.cpp:190-214; corrected flips the sign of :210 */
void FN(orc1d_residual)(const REAL* h_v, const REAL* h_f, REAL* r, int n, const double* range, int corrected)
{
    REAL h_x = FN(orc1d_h)(n, range), x_a = (REAL)range[0];
    for (int px = 0; px < n; px++) {
        if (px == 0 || px == n - 1) {
            r[px] = 0;
            continue;
        }
        REAL xj = x_a + px * h_x;
        if (corrected)
            r[px] = h_f[px] - (h_v[px + 1] - h_v[px]) / h_x + h_v[px] / (ORC_EXP(xj) + 1);
        else
            r[px] = h_f[px] - (h_v[px + 1] - h_v[px]) / h_x - h_v[px] / (ORC_EXP(xj) + 1);
    }
}

/* MultiGrid1D::Restrict, N1/MultiGrid1D.cpp:34-58 */
void FN(orc1d_restrict)(const REAL* fine, int n, REAL* coarse)
{
    int cn = (n - 1) / 2 + 1;
    for (int cx = 0; cx < cn; cx++) {
        if (cx == 0 || cx == cn - 1) {
            coarse[cx] = fine[2 * cx];
            continue;
        }
        REAL C = fine[2 * cx], E = fine[2 * cx + 1], O = fine[2 * cx - 1];
        coarse[cx] = (1 / 4.0f) * (O + 2 * C + E);
    }
}

/* MultiGrid1D::Interpolate, N1/MultiGrid1D.cpp:60-77 */
void FN(orc1d_interpolate)(REAL* fine, int n, const REAL* coarse)
{
    for (int fx = 1; fx < n - 1; fx++) {
        int cx = fx / 2;
        if (fx % 2 == 0) fine[fx] = coarse[cx];
        else fine[fx] = (1 / 2.0f) * (coarse[cx] + coarse[cx + 1]);
    }
}

/* MultiGrid1D::ApplyCorrection, N1/MultiGrid1D.cpp:177-188 */
void FN(orc1d_apply_correction)(REAL* fine, const REAL* err, int n)
{
    for (int px = 1; px < n - 1; px++) fine[px] = fine[px] + err[px];
}

/* MultiGrid1D::setToValue, N1/MultiGrid1D.cpp:120-130 */
void FN(orc1d_set)(REAL* g, int n, double value, int modify_boundaries)
{
    int lo = modify_boundaries ? 0 : 1, hi = modify_boundaries ? n : n - 1;
    for (int px = lo; px < hi; px++) g[px] = (REAL)value;
}

/* MultiGrid1D::VCycle, N1/MultiGrid1D.cpp:150-175 */
void FN(orc1d_vcycle)(REAL** v, REAL** f, int n0, int nlevels, const double* range, int level, int v1, int v2,
                      int corrected)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    FN(orc1d_relax)(v[level], f[level], n, range, v1);
    if (level != nlevels - 1) {
        REAL* tmp = (REAL*)malloc((size_t)n * sizeof(REAL));
        FN(orc1d_residual)(v[level], f[level], tmp, n, range, corrected);
        FN(orc1d_restrict)(tmp, n, f[level + 1]);
        FN(orc1d_set)(v[level + 1], (n - 1) / 2 + 1, 0.0, 1);
        FN(orc1d_vcycle)(v, f, n0, nlevels, range, level + 1, v1, v2, corrected);
        FN(orc1d_interpolate)(tmp, n, v[level + 1]);
        FN(orc1d_apply_correction)(v[level], tmp, n);
        free(tmp);
    }
    FN(orc1d_relax)(v[level], f[level], n, range, v2);
}

/* MultiGrid1D::FullMultiGridVCycle, N1/MultiGrid1D.cpp:132-148 */
void FN(orc1d_fmg)(REAL** v, REAL** f, int n0, int nlevels, const double* range, int level, int v0, int v1, int v2,
                   int corrected)
{
    int n = n0;
    for (int l = 0; l < level; l++) n = (n - 1) / 2 + 1;
    if (level != nlevels - 1) {
        FN(orc1d_restrict)(f[level], n, f[level + 1]);
        FN(orc1d_fmg)(v, f, n0, nlevels, range, level + 1, v0, v1, v2, corrected);
        FN(orc1d_interpolate)(v[level], n, v[level + 1]);
    } else {
        FN(orc1d_set)(v[level], n, 0.0, 0);
    }
    for (int i = 0; i < v0; i++) FN(orc1d_vcycle)(v, f, n0, nlevels, range, level, v1, v2, corrected);
}

#undef ORC_EXP
#undef FN
#undef ORC_CAT
#undef ORC_CAT2
