// Halo exchange over NVLink peer memory — one kernel per exchange, no NCCL on the
// critical path (SURVEY.md §7 hard part 7: ~10 exchanges per step, each a few
// hundred KB, are latency-bound; a grouped ncclSend/ncclRecv costs 20-40 us plus
// host dispatch, a peer-store kernel a few us).
//
// Every rank keeps its field windows in one cudaMalloc'ed arena whose IPC handle
// the neighbours have opened, so a neighbour's ghost region is an ordinary device
// pointer here.  halo_exchange_kernel
//   1. copies up to HALO_MAX_COPIES boundary strips (2-D, pitched, 4-byte words)
//      from this rank's windows into the neighbours' ghost regions — plain global
//      stores that travel over NVLink/NVSwitch;
//   2. last block out: __threadfence_system(), then writes the exchange's sequence
//      number into each neighbour's flag slot (st.release.sys);
//   3. waits until every neighbour's sequence number has arrived in this rank's
//      own flag slots (ld.acquire.sys), i.e. until this rank's ghosts are filled.
// Kernels after it in the stream may read the ghosts.  Every rank pushes before it
// waits, so there is no circular wait; ping-pong field buffers guarantee a
// neighbour never overwrites ghosts that are still being read (see dist.py).
#include "kernels.h"

namespace fs {

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256)
halo_exchange_kernel(const HaloArgs a, unsigned int *done_counter, int *status)
{
    // ---- 1. push ---------------------------------------------------------------------------
    // blocks are dealt to copies in proportion to their size (prefix sums in a.block_end)
    int c = 0;
    while (c < a.n_copies && (int)blockIdx.x >= a.block_end[c]) c++;
    if (c < a.n_copies) {
        const HaloCopy &d = a.copies[c];
        const int b0 = c == 0 ? 0 : a.block_end[c - 1];
        const int nb = a.block_end[c] - b0, b = blockIdx.x - b0;
        if (d.vec16) {
            // 16-byte path (rows, pitches and bases are 16-byte multiples: the usual case), four independent
            // load/store pairs in flight per thread — peer stores over NVLink are latency-bound otherwise
            // (round 2: 160 GB/s with one pair per thread)
            const unsigned row_q = (unsigned)(d.row_words >> 2);
            const unsigned quads = row_q * (unsigned)d.rows;
            const uint4 *src = reinterpret_cast<const uint4 *>(d.src);
            uint4 *dst = reinterpret_cast<uint4 *>(d.dst);
            const size_t sp = d.src_pitch_words >> 2, dp = d.dst_pitch_words >> 2;
            const unsigned stride = (unsigned)nb * blockDim.x;
            for (unsigned k0 = (unsigned)b * blockDim.x + threadIdx.x; k0 < quads; k0 += 4 * stride) {
                uint4 val[4];
                unsigned r[4], w[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned k = k0 + u * stride;
                    r[u] = k / row_q;
                    w[u] = k - r[u] * row_q;
                    if (k < quads) val[u] = __ldg(src + (size_t)r[u] * sp + w[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; u++)
                    if (k0 + u * stride < quads) dst[(size_t)r[u] * dp + w[u]] = val[u];
            }
        } else {
            const long long words = (long long)d.row_words * d.rows;
            for (long long k = (long long)b * blockDim.x + threadIdx.x; k < words; k += (long long)nb * blockDim.x) {
                const int r = (int)(k / d.row_words), w = (int)(k - (long long)r * d.row_words);
                d.dst[(size_t)r * d.dst_pitch_words + w] = d.src[(size_t)r * d.src_pitch_words + w];
            }
        }
    }
    // ---- 2. signal (last block out) -----------------------------------------------------------
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(done_counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done_counter = 0;   // ready for the next exchange (stream-ordered)
    __threadfence_system();
    if (threadIdx.x < a.n_peers) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.signal[threadIdx.x]), "l"(a.seq) : "memory");
        // ---- 3. wait -----------------------------------------------------------------------------
        // A neighbour that never signals (a crashed rank) must not hang this GPU: give up after
        // a.timeout_ns and raise the context's status flag (read by fs_tile_check).
        unsigned long long seen;
        const unsigned long long t0 = global_ns();
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.wait[threadIdx.x]) : "memory");
            if (seen >= a.seq) break;
            if (a.timeout_ns && global_ns() - t0 > a.timeout_ns) {
                if (status) atomicExch(status, FS_ERR_HALO_TIMEOUT);
                break;
            }
            __nanosleep(200);
        }
    }
}

int launch_halo_exchange(const Launch &L, HaloArgs &a, unsigned int *done_counter, int *status)
{
    if (a.n_copies < 0 || a.n_copies > HALO_MAX_COPIES || a.n_peers < 0 || a.n_peers > HALO_MAX_PEERS)
        return (int)cudaErrorInvalidValue;
    // ~16 KB of payload per block, at least one block per copy, at most 8 per SM
    long long total = 0;
    for (int c = 0; c < a.n_copies; c++) total += (long long)a.copies[c].row_words * a.copies[c].rows * 4;
    const int budget = L.num_sms * 8;
    int blocks = 0;
    for (int c = 0; c < a.n_copies; c++) {
        HaloCopy &h = a.copies[c];
        h.vec16 = ((h.row_words | h.src_pitch_words | h.dst_pitch_words) & 3) == 0 && (uintptr_t)h.src % 16 == 0 &&
                  (uintptr_t)h.dst % 16 == 0;
        const long long bytes = (long long)a.copies[c].row_words * a.copies[c].rows * 4;
        long long nb = (bytes + 16383) / 16384;
        if (total > 0 && nb > 1) {
            const long long cap = (long long)budget * bytes / total + 1;
            if (nb > cap) nb = cap;
        }
        if (nb < 1) nb = 1;
        blocks += (int)nb;
        a.block_end[c] = blocks;
    }
    if (blocks == 0) blocks = 1;   // nothing to copy: still signal + wait
    halo_exchange_kernel<<<blocks, 256, 0, L.stream>>>(a, done_counter, status);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int preload_halo_kernels()
{
    FS_PRELOAD(halo_exchange_kernel);
    return 0;
}

}  // namespace fs
