/* mg_host_common.c -- error state and device probing shared by the host drivers. */
#include "mg_host_common.h"

#include <stdarg.h>
#include <stdio.h>

static __thread char g_err[512] = "";

int mg_fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

const char* mg_last_error(void) { return g_err; }

const char* mg_version(void) { return "pde_multigrid_b200 0.1 (sm_100a)"; }

int mg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int mg_require_device(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1) {
        cudaGetLastError();
        return mg_fail(MG_ERR_CUDA, "no CUDA device available (%s): this engine has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    return MG_OK;
}
