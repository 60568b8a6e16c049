// mg_halo_p2p.cu -- halo exchange by direct NVLink stores into the neighbour's ghost planes.
//
// The z-slab neighbours map each other's field arena with CUDA IPC.  One "push" kernel copies this
// rank's boundary planes straight into the ghost planes of rank-1 / rank+1 (plain st.global on peer
// pointers: NVLink 5 through NVSwitch), fences system-wide, raises a sequence flag in the neighbour's
// memory and finally (last CTA, one thread, acquire loads, bounded) waits until both neighbours' flags
// have reached this exchange's sequence number.  One tiny launch per exchange and no
// host round trip or collective-library protocol: ~5 us instead of the ~20-90 us of a grouped
// ncclSend/ncclRecv (measured on 8 B200: profiles/), which is what the latency-bound coarse distributed
// levels need.  Write-after-read safety comes from the SPMD schedule: every rank alternates compute and
// bidirectional exchanges, so a neighbour cannot run ahead by more than one exchange.
#include <stdint.h>
#include <stdlib.h>

#include "mg_launch.h"

namespace {

// flag block layout (unsigned words, one per 128-byte line): [0] raised by rank-1, [32] raised by rank+1,
// [64] push-completion counter, [96] wait-timeout error, [128] sent-to-below count, [160] sent-to-above count,
// [192] expected-from-below count, [224] expected-from-above count.  The sequence counters live in device
// memory so that the kernel arguments never change: the exchange can be captured into a CUDA graph.
struct XArgs {
    const void* src[4];
    void* dst[4];
    unsigned long long bytes[4];  // multiples of 16
    unsigned int* peer_flag[2];   // [0]: word in rank-1's block to raise, [1]: word in rank+1's block (nullptr = no send)
    int wait_below, wait_above;   // expect a push from rank-1 / rank+1
    unsigned int* block;          // my flag block
    long long max_cycles;
};

__global__ void __launch_bounds__(256) k_halo_exchange(XArgs a)
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
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence_system();  // (cumulative over the CTA barrier) this CTA's stores are visible to the peer before the flag
    unsigned int* blk = a.block;
    const unsigned int prev = atomicAdd(blk + 64, 1u);
    if (prev != gridDim.x - 1) return;
    // last CTA: every segment has been stored and fenced
    blk[64] = 0;
    __threadfence_system();
    for (int k = 0; k < 2; k++)
        if (a.peer_flag[k]) {
            const unsigned int v = ++blk[128 + 32 * k];
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_flag[k]), "r"(v) : "memory");
        }
    // then wait for the neighbours' pushes of this same exchange.  Bounded (a GPU must never hang): on expiry the sticky
    // error word is set -- every API call that hands data back checks it -- and BOTH expected counters have already
    // advanced, so later exchanges stay in step.  max_cycles <= 0: wait for ever.
    const long long t0 = clock64();
    const int waits[2] = {a.wait_below, a.wait_above};
    unsigned int want[2] = {0, 0};
    for (int k = 0; k < 2; k++)
        if (waits[k]) want[k] = ++blk[192 + 32 * k];
    for (int k = 0; k < 2; k++) {
        if (!waits[k]) continue;
        const unsigned int* fl = blk + 32 * k;
        while (true) {
            unsigned int cur;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(fl) : "memory");
            if ((int)(cur - want[k]) >= 0) break;
            if (a.max_cycles > 0 && clock64() - t0 > a.max_cycles) {
                blk[96] = 1;
                break;
            }
        }
    }
}

// ---- agglomeration gather by direct stores (replaces ncclAllGather on the first agglomerated level) ----
// Every rank stores its share of the restricted right-hand side into ALL peers' copies of the level.  Unlike a halo exchange
// the receivers are not only the schedule-coupled neighbours, so the write-after-read safety is explicit, two phases per gather:
//   ready: rank r tells every peer "I am at gather #seq: nothing of mine reads the previous contents any more", and waits for
//          the same word from every peer -- only then does anybody overwrite anybody's array;
//   done:  after the stores (system fence), rank r tells every peer "my share of #seq is in your memory" and waits for all.
// Flag words of the block (one 128-byte line of 32 words each): [256 + p] ready from rank p, [288 + p] done from rank p,
// [320] this rank's gather sequence number, [352] push-completion counter.  nranks <= 32.
struct GArgs {
    const void* src[2];            // my share: colour 0, colour 1
    unsigned long long bytes;      // per colour, multiple of 16
    void* dst[MG_GATHER_MAX_PEERS][2];
    unsigned int* peer_block[MG_GATHER_MAX_PEERS];
    int npeers, me;
    unsigned int* block;
    long long max_cycles;
};

__device__ __forceinline__ void wait_word(const unsigned int* w, unsigned int want, long long t0, long long max_cycles, unsigned int* err)
{
    while (true) {
        unsigned int cur;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(w) : "memory");
        if ((int)(cur - want) >= 0) return;
        if (max_cycles > 0 && clock64() - t0 > max_cycles) { *err = 1; return; }
    }
}

__global__ void k_gather_ready(GArgs a)
{
    if (threadIdx.x) return;
    unsigned int* blk = a.block;
    const unsigned int seq = ++blk[320];
    __threadfence_system();
    for (int p = 0; p < a.npeers; p++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_block[p] + 256 + a.me), "r"(seq) : "memory");
    const long long t0 = clock64();
    for (int p = 0; p < a.npeers; p++) {
        const int pr = p < a.me ? p : p + 1;  // rank of peer slot p
        wait_word(blk + 256 + pr, seq, t0, a.max_cycles, blk + 96);
    }
}

__global__ void __launch_bounds__(256) k_gather_push(GArgs a)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    const size_t n16 = a.bytes / 16;
    for (int c = 0; c < 2; c++) {
        const uint4* src = reinterpret_cast<const uint4*>(a.src[c]);
        for (size_t i = tid; i < n16; i += nth) {
            const uint4 v = src[i];
            for (int p = 0; p < a.npeers; p++) reinterpret_cast<uint4*>(a.dst[p][c])[i] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence_system();
    unsigned int* blk = a.block;
    const unsigned int prev = atomicAdd(blk + 352, 1u);
    if (prev != gridDim.x - 1) return;
    blk[352] = 0;
    __threadfence_system();
    const unsigned int seq = blk[320];
    for (int p = 0; p < a.npeers; p++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.peer_block[p] + 288 + a.me), "r"(seq) : "memory");
    const long long t0 = clock64();
    for (int p = 0; p < a.npeers; p++) {
        const int pr = p < a.me ? p : p + 1;
        wait_word(blk + 288 + pr, seq, t0, a.max_cycles, blk + 96);
    }
}

long long halo_max_cycles()
{
    static long long max_cycles = -2;
    if (max_cycles == -2) {
        const char* env = getenv("MG_B200_HALO_TIMEOUT_S");
        const double sec = env ? atof(env) : 300.0;
        max_cycles = sec > 0 ? (long long)(sec * 2.0e9) : 0;
    }
    return max_cycles;
}

}  // namespace

extern "C" {

/* All-gather of one level field by direct stores: src2 = this rank's share of the two colour arrays (`bytes` each), dst[p] = the
   same two places inside peer slot p's copy, peer_block[p] = that peer's flag block.  Two launches (ready handshake, push + done). */
int mgk_gather_push(cudaStream_t s, const void* const src2[2], unsigned long long bytes, void* const dst[][2], unsigned int* const peer_block[],
                    int npeers, int me, unsigned int* flag_block)
{
    if (npeers < 1 || npeers > MG_GATHER_MAX_PEERS) return -1;
    GArgs a;
    a.src[0] = src2[0]; a.src[1] = src2[1];
    a.bytes = bytes;
    for (int p = 0; p < npeers; p++) { a.dst[p][0] = dst[p][0]; a.dst[p][1] = dst[p][1]; a.peer_block[p] = peer_block[p]; }
    a.npeers = npeers; a.me = me;
    a.block = flag_block;
    a.max_cycles = halo_max_cycles();
    k_gather_ready<<<1, 32, 0, s>>>(a);
    if (cudaPeekAtLastError() != cudaSuccess) return -1;
    unsigned long long want = (bytes / 16 + 256 * 4 - 1) / (256 * 4);
    int grid = (int)(want < 1 ? 1 : (want > 296 ? 296 : want));
    k_gather_push<<<grid, 256, 0, s>>>(a);
    return cudaPeekAtLastError() == cudaSuccess ? 2 : -1;
}


/* One launch per halo exchange: push up to 4 (src, dst, bytes) segments into the neighbours' ghost planes,
   raise their sequence flags, then wait for the neighbours' own pushes.  peer_flag[k] == NULL: nothing is sent
   in that direction. */
int mgk_halo_exchange(cudaStream_t s, const void* const src[4], void* const dst[4], const unsigned long long bytes[4],
                      unsigned int* const peer_flag[2], int wait_below, int wait_above, unsigned int* flag_block)
{
    XArgs a;
    unsigned long long total = 0;
    for (int i = 0; i < 4; i++) {
        a.src[i] = src[i];
        a.dst[i] = dst[i];
        a.bytes[i] = bytes[i];
        total += bytes[i];
    }
    a.peer_flag[0] = peer_flag[0];
    a.peer_flag[1] = peer_flag[1];
    a.wait_below = wait_below;
    a.wait_above = wait_above;
    a.block = flag_block;
    /* ordinary rank skew (a neighbour busy with host work between two calls) must not become an error: five minutes by
       default, MG_B200_HALO_TIMEOUT_S seconds if set (0 = no limit) */
    const long long max_cycles = halo_max_cycles();
    a.max_cycles = max_cycles;
    if (!total && !peer_flag[0] && !peer_flag[1] && !wait_below && !wait_above) return 0;
    unsigned long long want = (total / 16 + 256 * 8 - 1) / (256 * 8);  // ~8 x 16 B per thread
    int grid = (int)(want < 1 ? 1 : (want > 296 ? 296 : want));
    k_halo_exchange<<<grid, 256, 0, s>>>(a);
    return cudaPeekAtLastError() == cudaSuccess ? 1 : -1;
}

}  // extern "C"
