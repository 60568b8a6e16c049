/*
 * include/compat/mg_compat_common.h -- helpers shared by the header-compatible C++ shim classes.
 *
 * The shims give the reference's own class interface (NOCUDA_TESI variant: host arrays h_v / h_f as
 * public members, host int[] sizes) on top of the C ABI of libmg_b200.so, so that the reference's
 * main() files build unchanged:   g++ -I include/compat -I include wrapper.cpp -lmg_b200
 * They use the reference's include-guard names, so a wrapper TU that includes them first makes the
 * reference's own headers no-ops.  Error convention of the reference is assert()
 * -> abort(): MG_CHECK keeps it.  All numerics run on the GPU; nothing here computes on the CPU
 * except the diagnostic dumps (PrintDiff & co), which are not on the hot path.
 */
#ifndef MG_COMPAT_COMMON_H
#define MG_COMPAT_COMMON_H

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "mg_b200.h"

/* -DMG_COMPAT_CUDA_TESI selects the faces of the reference's GPU twin (CUDA_TESI/...): grids carry DEVICE arrays
   (d_v, d_f, 3D: d_sizeXYZ; 2D: d_pitch) and the operators take device pointers -- the methods have the same C++
   signatures as the host ones, so the interpretation is a compile-time choice, exactly as it is between the reference's
   two source trees.  Needs the CUDA runtime headers and -lcudart. */
#ifdef MG_COMPAT_CUDA_TESI
#include <cuda_runtime_api.h>
#define MG_CUDA_CHECK(call)                                                                     \
    do {                                                                                        \
        cudaError_t mg_e_ = (call);                                                             \
        if (mg_e_ != cudaSuccess) {                                                             \
            fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(mg_e_));               \
            abort();                                                                            \
        }                                                                                       \
    } while (0)
#endif

#define MG_CHECK(call)                                                                          \
    do {                                                                                        \
        int mg_st_ = (call);                                                                    \
        if (mg_st_ != MG_OK) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #call, mg_st_, mg_last_error());            \
            abort();                                                                            \
        }                                                                                       \
    } while (0)

static inline int mg_compat_open_log(const char* path) { return open(path, O_RDWR | O_CREAT | O_TRUNC, S_IRWXU); }

static inline void mg_compat_write(int fd, const char* line)
{
    if (fd >= 0) {
        ssize_t r = write(fd, line, strlen(line));
        (void)r;
    }
}

#endif
