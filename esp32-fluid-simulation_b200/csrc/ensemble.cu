// Batched ensemble of small grids (BASELINE.json configs[1]: 65,536 independent
// 80x60-class sims): one CTA per grid, the WHOLE loop() body (ino:249-289) and
// any number of consecutive steps run out of shared memory; HBM is touched only
// to load the state once and to store it once per call (40 B/node).
//
// Shared-memory plan for N = dim_x*dim_y nodes (40*N bytes; 61x81 -> 197.6 KB):
//     A, B   : two velocity buffers (8N each).  advect reads A, writes B; A is
//              then dead and is reused as d (4N) | p (4N) for the projection;
//              the projected velocity ends up in B, and A/B swap roles.
//     C1, C2 : two dye buffers (12N each), ping-pong for the dye advect.
// (The reference's six separate arrays would need 48N = 237 KB > 227 KB.)
//
// Work is organised by PAIRS of consecutive nodes (2q, 2q+1): a pair always holds
// one node of each red/black colour, so in every half-sweep each thread updates
// exactly one node of each of its pairs.  (i,j) of a thread's pairs are computed
// once per grid and kept in registers.
#include "advect.cuh"
#include "kernels.h"
#include "sor.cuh"

namespace fs {

constexpr int ENS_MAX_THREADS = 1024;
constexpr int ENS_MAX_ROUNDS = 3;                                       // pairs per thread
constexpr int ENS_MAX_NODES = 2 * ENS_MAX_THREADS * ENS_MAX_ROUNDS;     // also bounded by smem (40 B/node)

template <class P>
struct SmemFetch {
    const typename P::raw_t *base;
    int dim_x;
    __device__ __forceinline__ void operator()(int gi, int gj, typename P::raw_t (&o)[P::NC]) const
    {
        const typename P::raw_t *q = base + (gj * dim_x + gi) * P::NC;
#pragma unroll
        for (int ch = 0; ch < P::NC; ch++) o[ch] = q[ch];
    }
};

struct EnsArgs {
    float2 *v;
    uint32_t *c;
    const fs_drag *drags;   // device: [n_steps][batch][max_drags]
    const int *counts;      // device: [n_steps][batch]
    int max_drags, batch, dim_x, dim_y, iters, n_steps;
    float dt, two_dx_inv;
    SorCoef k;
};

__global__ void __launch_bounds__(ENS_MAX_THREADS, 1) ensemble_kernel(const EnsArgs a)
{
    // the CTA is sized by the launcher so that the node pairs divide (almost) evenly over the threads:
    // every phase ends in a barrier, and idle threads in the last round were the top stall (ncu)
    const int ENS_THREADS = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.dim_x * a.dim_y, dim_x = a.dim_x, dim_y = a.dim_y;
    float2 *A = reinterpret_cast<float2 *>(smem_raw);
    float2 *B = A + N;
    uint32_t *C1 = reinterpret_cast<uint32_t *>(B + N);
    uint32_t *C2 = C1 + 3 * (size_t)N;
    const int tid = threadIdx.x;

    // this thread's pairs: node n0 = 2q at (i0, j0); n0+1 is the next node in index order
    int n0[ENS_MAX_ROUNDS], ij0[ENS_MAX_ROUNDS];
#pragma unroll
    for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
        const int n = 2 * (tid + r * ENS_THREADS);
        n0[r] = n < N ? n : -1;
        const int j = n / dim_x;
        ij0[r] = (j << 16) | (n - j * dim_x);
    }

    for (int g = blockIdx.x; g < a.batch; g += gridDim.x) {
        // ---- load the grid's state -------------------------------------------------------
        const float2 *gv = a.v + (size_t)g * N;
        const uint32_t *gc = a.c + (size_t)g * N * 3;
        for (int n = tid; n < N; n += ENS_THREADS) A[n] = __ldg(gv + n);
        for (int n = tid; n < 3 * N; n += ENS_THREADS) C1[n] = __ldg(gc + n);
        __syncthreads();

        for (int step = 0; step < a.n_steps; step++) {
            // ---- advect velocity, no-slip (ino:253): A -> B ---------------------------------
            {
                SmemFetch<Vec2Payload> fetch{reinterpret_cast<const float *>(A), dim_x};
#pragma unroll
                for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                    if (n0[r] < 0) continue;
                    int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const int n = n0[r] + cc;
                        if (n < N) {
                            float si, sj, out[2];
                            backtrace(si, sj, i, j, A[n], a.dt);
                            sample<Vec2Payload>(out, fetch, si, sj, dim_x, dim_y, true);
                            B[n] = make_float2(out[0], out[1]);
                        }
                        if (++i == dim_x) { i = 0; j++; }
                    }
                }
            }
            __syncthreads();
            // ---- drags (ino:264-269): in order, one thread ------------------------------------
            if (a.max_drags > 0 && tid == 0) {
                const size_t slot = (size_t)step * a.batch + g;
                const int cnt = min(a.counts[slot], a.max_drags);
                const fs_drag *dr = a.drags + slot * a.max_drags;
                for (int q = 0; q < cnt; q++) {
                    const fs_drag m = dr[q];
                    if (m.cy < dim_x && m.cx < dim_y) B[m.cx * dim_x + m.cy] = make_float2(m.vy, m.vx);
                }
            }
            __syncthreads();
            // ---- divergence (ino:274) into d, zero p (poisson.cpp:117-119); A is dead now ------
            float *d = reinterpret_cast<float *>(A);
            float *p = d + N;
#pragma unroll
            for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                if (n0[r] < 0) continue;
                int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int n = n0[r] + cc;
                    if (n < N) {
                        const bool wall = i == 0 || i == dim_x - 1 || j == 0 || j == dim_y - 1;
                        float s;
                        if (!wall) {
                            s = __fadd_rn(__fadd_rn(-B[n - 1].x, B[n + 1].x),
                                          __fadd_rn(-B[n - dim_x].y, B[n + dim_x].y));
                        } else {
                            const float2 c0 = B[n];
                            s = 0.0f;
                            s = __fadd_rn(s, i > 0 ? -B[n - 1].x : c0.x);
                            s = __fadd_rn(s, i < dim_x - 1 ? B[n + 1].x : -c0.x);
                            s = __fadd_rn(s, j > 0 ? -B[n - dim_x].y : c0.y);
                            s = __fadd_rn(s, j < dim_y - 1 ? B[n + dim_x].y : -c0.y);
                        }
                        // store dx*d: the same product every iteration (poisson.cpp:88,109)
                        d[n] = __fmul_rn(a.k.dx, __fmul_rn(s, a.two_dx_inv));
                        p[n] = 0.0f;
                    }
                    if (++i == dim_x) { i = 0; j++; }
                }
            }
            __syncthreads();
            // ---- red-black SOR (ino:275): colour 0 = (i+j) even first ---------------------------
            for (int hs = 0; hs < 2 * a.iters; hs++) {
                const int parity = hs & 1;
#pragma unroll
                for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                    if (n0[r] < 0) continue;
                    int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
                    int n = n0[r];
                    if (((i + j) & 1) != parity) {      // the pair's other node has this colour
                        n++;
                        if (++i == dim_x) { i = 0; j++; }
                    }
                    if (n >= N) continue;
                    const bool hl = i > 0, hr = i < dim_x - 1, hd = j > 0, hu = j < dim_y - 1;
                    const float pc = p[n];
                    float out;
                    if (hl && hr && hd && hu)
                        out = sor_update_interior(pc, p[n - 1], p[n + 1], p[n - dim_x], p[n + dim_x], d[n], a.k);
                    else
                        out = sor_update_wall(pc, hl ? p[n - 1] : 0.f, hr ? p[n + 1] : 0.f,
                                              hd ? p[n - dim_x] : 0.f, hu ? p[n + dim_x] : 0.f, hl, hr,
                                              hd, hu, d[n], a.k);
                    p[n] = out;
                }
                __syncthreads();
            }
            // ---- subtract gradient (ino:276), in place on B --------------------------------------
#pragma unroll
            for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                if (n0[r] < 0) continue;
                int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    const int n = n0[r] + cc;
                    if (n < N) {
                        const float pc = p[n];
                        const float pl = i > 0 ? p[n - 1] : pc, pr = i < dim_x - 1 ? p[n + 1] : pc;
                        const float pd = j > 0 ? p[n - dim_x] : pc, pu = j < dim_y - 1 ? p[n + dim_x] : pc;
                        float2 c0 = B[n];
                        c0.x = __fsub_rn(c0.x, __fmul_rn(__fsub_rn(pr, pl), a.two_dx_inv));
                        c0.y = __fsub_rn(c0.y, __fmul_rn(__fsub_rn(pu, pd), a.two_dx_inv));
                        B[n] = c0;
                    }
                    if (++i == dim_x) { i = 0; j++; }
                }
            }
            __syncthreads();
            // ---- advect dye, free-slip sampling (ino:282): C1 -> C2 with the projected velocity -----
            {
                SmemFetch<RgbPayload> fetch{C1, dim_x};
#pragma unroll
                for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                    if (n0[r] < 0) continue;
                    int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        const int n = n0[r] + cc;
                        if (n < N) {
                            float si, sj;
                            uint32_t out[3];
                            backtrace(si, sj, i, j, B[n], a.dt);
                            sample<RgbPayload>(out, fetch, si, sj, dim_x, dim_y, false);
                            C2[3 * n + 0] = out[0];
                            C2[3 * n + 1] = out[1];
                            C2[3 * n + 2] = out[2];
                        }
                        if (++i == dim_x) { i = 0; j++; }
                    }
                }
            }
            __syncthreads();
            // pointer swaps of ino:255 and ino:286
            float2 *tv = A; A = B; B = tv;
            uint32_t *tc = C1; C1 = C2; C2 = tc;
        }

        // ---- store the grid's state ----------------------------------------------------------
        float2 *ov = a.v + (size_t)g * N;
        uint32_t *oc = a.c + (size_t)g * N * 3;
        for (int n = tid; n < N; n += ENS_THREADS) ov[n] = A[n];
        for (int n = tid; n < 3 * N; n += ENS_THREADS) oc[n] = C1[n];
        __syncthreads();
    }
}

size_t ensemble_smem_bytes(int dim_x, int dim_y) { return (size_t)40 * dim_x * dim_y; }

bool ensemble_supported(int dim_x, int dim_y, size_t max_smem_optin)
{
    const long long n = (long long)dim_x * dim_y;
    return dim_x < 65536 && dim_y < 32768 && n <= ENS_MAX_NODES && (size_t)(40 * n) <= max_smem_optin;
}

int launch_ensemble(const Launch &L, float2 *v, uint32_t *c, const fs_drag *drags_dev,
                    const int *counts_dev, int max_drags, int batch, int dim_x, int dim_y, float dt,
                    float dx, int iters, float omega, int n_steps)
{
    if (batch <= 0 || n_steps <= 0) return 0;
    EnsArgs a;
    a.v = v; a.c = c; a.drags = drags_dev; a.counts = counts_dev;
    a.max_drags = max_drags; a.batch = batch; a.dim_x = dim_x; a.dim_y = dim_y;
    a.iters = iters; a.n_steps = n_steps; a.dt = dt;
    a.two_dx_inv = 1.0f / (2.0f * dx);
    a.k = make_sor_coef(dx, omega);
    const size_t smem = ensemble_smem_bytes(dim_x, dim_y);
    cudaError_t e = cudaFuncSetAttribute(ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = batch < L.num_sms ? batch : L.num_sms;   // persistent: one CTA per SM walks the batch
    // threads: ENS_MAX_ROUNDS pairs per thread, rounded up to whole warps (80x60 -> 800, 61x81 -> 832)
    const int pairs = (dim_x * dim_y + 1) / 2;
    int threads = ((pairs + ENS_MAX_ROUNDS - 1) / ENS_MAX_ROUNDS + 31) / 32 * 32;
    if (threads < 128) threads = 128;
    if (threads > ENS_MAX_THREADS) threads = ENS_MAX_THREADS;
    ensemble_kernel<<<grid, threads, smem, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

}  // namespace fs
