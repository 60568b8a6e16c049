// mg_halo_p2p.cu -- halo exchange by direct NVLink stores into the neighbour's ghost planes.
//
// The z-slab neighbours map each other's field arena with CUDA IPC.  One "push" kernel copies this
// rank's boundary planes straight into the ghost planes of rank-1 / rank+1 (plain st.global on peer
// pointers: NVLink 5 through NVSwitch), fences system-wide and then raises a sequence flag in the
// neighbour's memory; one "wait" kernel spins (single thread, acquire loads, bounded) until both
// neighbours' flags have reached the expected sequence number.  Two tiny launches per exchange and no
// host round trip or collective-library protocol: ~5 us instead of the ~20-90 us of a grouped
// ncclSend/ncclRecv (measured on 8 B200: profiles/), which is what the latency-bound coarse distributed
// levels need.  Write-after-read safety comes from the SPMD schedule: every rank alternates compute and
// bidirectional exchanges, so a neighbour cannot run ahead by more than one exchange.
#include <stdint.h>

#include "mg_launch.h"

namespace {

struct PushArgs {
    const void* src[4];
    void* dst[4];
    unsigned long long bytes[4];  // multiples of 16
    unsigned int* flag[2];        // peer flags to raise (nullptr = none)
    unsigned int value[2];
    unsigned int* done_counter;   // local, zero on entry, reset by the last CTA
};

__global__ void __launch_bounds__(256) k_halo_push(PushArgs a)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nth = (size_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (!a.bytes[s]) continue;
        const uint4* src = reinterpret_cast<const uint4*>(a.src[s]);
        uint4* dst = reinterpret_cast<uint4*>(a.dst[s]);
        const size_t n16 = a.bytes[s] / 16;
        for (size_t i = tid; i < n16; i += nth) dst[i] = src[i];
    }
    __threadfence_system();  // my stores are visible to the peer before the flag
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(a.done_counter, 1u);
        if (prev == gridDim.x - 1) {  // last CTA: everything has been stored and fenced
            *a.done_counter = 0;
            __threadfence_system();
            for (int f = 0; f < 2; f++)
                if (a.flag[f]) {
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag[f]), "r"(a.value[f]) : "memory");
                }
        }
    }
}

__global__ void k_halo_wait(const unsigned int* f0, unsigned int v0, const unsigned int* f1, unsigned int v1, unsigned int* error_flag,
                            long long max_cycles)
{
    const long long t0 = clock64();
    const unsigned int* fl[2] = {f0, f1};
    const unsigned int want[2] = {v0, v1};
    for (int k = 0; k < 2; k++) {
        if (!fl[k]) continue;
        while (true) {
            unsigned int cur;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(fl[k]) : "memory");
            if ((int)(cur - want[k]) >= 0) break;
            if (clock64() - t0 > max_cycles) {  // never hang the GPU: report and carry on
                *error_flag = 1;
                return;
            }
            __nanosleep(100);
        }
    }
}

}  // namespace

extern "C" {

/* up to 4 (src, dst, bytes) segments; flags[k] (may be NULL) receives values[k] once all segments are stored */
int mgk_halo_push(cudaStream_t s, const void* const src[4], void* const dst[4], const unsigned long long bytes[4],
                  unsigned int* const flags[2], const unsigned int values[2], unsigned int* done_counter)
{
    PushArgs a;
    unsigned long long total = 0;
    for (int i = 0; i < 4; i++) {
        a.src[i] = src[i];
        a.dst[i] = dst[i];
        a.bytes[i] = bytes[i];
        total += bytes[i];
    }
    for (int k = 0; k < 2; k++) {
        a.flag[k] = flags[k];
        a.value[k] = values[k];
    }
    a.done_counter = done_counter;
    if (!total && !flags[0] && !flags[1]) return 0;
    unsigned long long want = (total / 16 + 256 * 8 - 1) / (256 * 8);  // ~8 x 16 B per thread
    int grid = (int)(want < 1 ? 1 : (want > 296 ? 296 : want));
    k_halo_push<<<grid, 256, 0, s>>>(a);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

int mgk_halo_wait(cudaStream_t s, const unsigned int* f0, unsigned int v0, const unsigned int* f1, unsigned int v1,
                  unsigned int* error_flag)
{
    if (!f0 && !f1) return 0;
    k_halo_wait<<<1, 1, 0, s>>>(f0, v0, f1, v1, error_flag, 6000000000LL /* ~3 s */);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // extern "C"
