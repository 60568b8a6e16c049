/*
 * oracle/ref_wrap_common.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Shared prologue of the wrapper translation units that compile the UNMODIFIED
 * reference CPU solver (NOCUDA_TESI) straight from /root/reference into
 * oracle/_ref/libmg_ref.so.  No reference source is copied into this repo: the
 * wrappers #include the reference .cpp files by name through -I search paths
 * given by oracle/build_ref.py.
 *
 * What this header does, in order:
 *   1. pulls in every libc header the reference's inclusion.h pulls in, BEFORE
 *      any macro games, so `#define float double` cannot mangle libc;
 *   2. interposes malloc with a tracking allocator: the reference leaks the
 *      residual and the interpolated error on every level of every V-cycle
 *      (N3/MultiGrid3D.cpp:629,638; N2/MultiGrid2D.cpp:320,331;
 *      N1/MultiGrid1D.cpp:156,167) -- the harness frees them after each cycle;
 *      every block is over-allocated by 64 bytes, which also neutralises the
 *      2-float/4-write heap overflow of InitA (N2/MultiGrid2D.cpp:50-58).
 *   3. optionally re-defines `float` to `double` (REF_F64) -- the fp64 oracle
 *      is the same source, as verified in SURVEY.md section 8c.
 */
#ifndef ORACLE_REF_WRAP_COMMON_H
#define ORACLE_REF_WRAP_COMMON_H

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <fcntl.h>
#include <string.h>
#include <unistd.h>
#include <time.h>
#include <stdint.h>

namespace refwrap {

struct Tracker {
    void** ptrs;
    size_t n, cap;
    int on;
};

static Tracker g_trk = {0, 0, 0, 0};

static inline void* tracked_malloc(size_t nbytes)
{
    void* p = ::malloc(nbytes + 64);
    if (g_trk.on && p) {
        if (g_trk.n == g_trk.cap) {
            g_trk.cap = g_trk.cap ? 2 * g_trk.cap : 64;
            g_trk.ptrs = (void**)::realloc(g_trk.ptrs, g_trk.cap * sizeof(void*));
        }
        g_trk.ptrs[g_trk.n++] = p;
    }
    return p;
}

static inline void track_begin() { g_trk.on = 1; g_trk.n = 0; }

/* free everything the reference allocated (and leaked) since track_begin() */
static inline void track_end_free()
{
    for (size_t i = 0; i < g_trk.n; i++) ::free(g_trk.ptrs[i]);
    g_trk.n = 0;
    g_trk.on = 0;
}

} // namespace refwrap

#define malloc(n) refwrap::tracked_malloc(n)

#ifdef REF_F64
typedef double ref_real;
#define float double
#else
typedef float ref_real;
#endif

#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
/* exported symbol name: e.g. ref3d_f64c_relax */
#define REF_FN(name) REF_CAT(REF_CAT(REF_PREFIX, _), name)

#endif
