/*
 * mg_host_common.h -- shared helpers of the C host drivers: error reporting, CUDA call checking,
 * pitched-layout arithmetic.  Internal; the public ABI is include/mg_b200.h.
 */
#ifndef MG_HOST_COMMON_H
#define MG_HOST_COMMON_H

#include <cuda_runtime_api.h>
#include <stddef.h>

#include "../../include/mg_b200.h"
#include "mg_launch.h"

#ifdef __cplusplus
extern "C" {
#endif

int mg_fail(int code, const char* fmt, ...); /* records the message for mg_last_error(), returns code */

#define MG_CUDA(call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return mg_fail(MG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

/* launcher return value -> status; adds to the handle's launch counter */
#define MG_LAUNCH(counter, call)                                                                     \
    do {                                                                                             \
        int k_ = (call);                                                                             \
        if (k_ < 0)                                                                                  \
            return mg_fail(MG_ERR_CUDA, "%s: launch failed: %s", #call, cudaGetErrorString(cudaGetLastError())); \
        (counter) += k_;                                                                             \
    } while (0)

static inline size_t mg_esize(int dtype) { return dtype == MG_F32 ? 4 : 8; }
/* row pitch in elements: n rounded up to 128 bytes */
static inline int mg_pitch(int n, int dtype)
{
    int per = (int)(128 / mg_esize(dtype));
    return (n + per - 1) / per * per;
}
static inline size_t mg_align256(size_t b) { return (b + 255) & ~(size_t)255; }
/* numGrids = (int)log2(minSize-1), N3/MultiGrid3D.cpp:33-34 */
static inline int mg_num_levels_for(int n)
{
    int k = 0, m = n - 1;
    while (m > 1) { m >>= 1; k++; }
    return k;
}
/* 128-byte CUtensorMap over one colour array of a colour-split 3D field (mg_tma_host.c) */
int mg_tma_make_colour_map(void* out128, int dtype, void* base, const mg_geom3d* g, int box_i, int box_y);
int mg_require_device(void); /* MG_OK, or MG_ERR_CUDA with a message: there is no CPU fallback */

#ifdef __cplusplus
}
#endif
#endif
