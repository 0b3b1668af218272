// Batched ensemble of small grids (BASELINE.json configs[1]: 65,536 independent
// 80x60-class sims): one CTA per grid, the WHOLE loop() body (ino:249-289) and
// any number of consecutive steps run with the velocity, divergence and pressure
// resident in shared memory.
//
// Shared-memory plan for N = dim_x*dim_y nodes (16*N bytes; 61x81 -> 79 KB, so TWO
// grids are resident per SM and one grid's barriers overlap the other's work):
//     A, B : two velocity buffers (8N each).  advect reads A, writes B; A is then
//            dead and is reused as d (4N) | p (4N) for the projection; the
//            projected velocity ends up in B, and A/B swap roles.
//     dye  : NOT held in shared memory (24 of round 1's 40 B/node): the dye advect
//            gathers it from global memory through L2 and writes the result back;
//            between the steps of one call it ping-pongs between the caller's array
//            and a per-CTA scratch slot that stays L2-resident.
//     d, p : stored COLOUR-SEPARATED — node n lives at [colour(n)][n >> 1] — so the
//            stride-2 red/black accesses of a half-sweep become unit-stride (round 1:
//            43 % of the shared-memory wavefronts were bank conflicts).
//
// Work is organised by PAIRS of consecutive nodes (2q, 2q+1): a pair always holds
// one node of each red/black colour, so in every half-sweep each thread updates
// exactly one node of each of its pairs.  (i,j) of a thread's pairs are computed
// once per grid and kept in registers.
#include "ensemble_reg.cuh"
#include "kernels.h"

namespace fs {

constexpr int ENS_MAX_NODES = 6144;     // 2 nodes x threads x pairs per thread of every variant below

// ENS_MAX_THREADS x MINB = CTA size limit and CTAs per SM the kernel is compiled for; ENS_MAX_ROUNDS =
// node pairs per thread
// DYE_SMEM: both dye buffers live in shared memory too (40 B/node, one grid per SM) — faster than going
// through L1/L2 for every corner while it fits; otherwise the dye is streamed (16 B/node).
template <int ENS_MAX_THREADS, int MINB, int ENS_MAX_ROUNDS, bool DYE_SMEM>
__global__ void __launch_bounds__(ENS_MAX_THREADS, MINB) ensemble_kernel(const EnsArgs a)
{
    // the CTA is sized by the launcher so that the node pairs divide (almost) evenly over the threads:
    // every phase ends in a barrier, and idle threads in the last round were the top stall (ncu)
    const int ENS_THREADS = blockDim.x;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = a.dim_x * a.dim_y, dim_x = a.dim_x, dim_y = a.dim_y;
    const int HQ = (N + 1) >> 1;                       // nodes per colour plane
    float2 *A = reinterpret_cast<float2 *>(smem_raw);
    float2 *B = A + N + 1;                             // + 8 bytes: the two colour planes of d and p need 4*(N+1) each
    const int tid = threadIdx.x;
    // dye ping-pong: two shared-memory buffers, or the caller's array and this CTA's global scratch slot
    uint32_t *C1 = DYE_SMEM ? reinterpret_cast<uint32_t *>(B + N) : nullptr;
    uint32_t *C2 = DYE_SMEM ? C1 + 3 * (size_t)N : a.scratch + (size_t)blockIdx.x * N * 3;

    // this thread's pairs: node n0 = 2q at (i0, j0); n0+1 is the next node in index order
    int n0[ENS_MAX_ROUNDS], ij0[ENS_MAX_ROUNDS];
#pragma unroll
    for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
        const int n = 2 * (tid + r * ENS_THREADS);
        n0[r] = n < N ? n : -1;
        const int j = n / dim_x;
        ij0[r] = (j << 16) | (n - j * dim_x);
    }
    // colour-separated address of node n: plane (i+j)&1, slot n>>1
    auto cs = [&](int n, int colour) { return colour * HQ + (n >> 1); };
    // SOR descriptors: for pair r and colour c, the node n of that colour: slot q = n>>1 (bits 0-12),
    // b = n&1 (bit 13: the horizontal neighbours sit at slots q+b-1, q+b of the other plane), bo = b &
    // (dim_x odd) (bit 14: the vertical neighbours at slots q+bo-hd0, q+bo+hu0), one bit per EXISTING
    // neighbour (15 left, 16 right, 17 down, 18 up), and the index of neg_a_ii_inv[#neighbours] (19-20).
    const int hu0 = dim_x >> 1, hd0 = (dim_x + 1) >> 1;
    unsigned node_desc[ENS_MAX_ROUNDS][2];
#pragma unroll
    for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
        node_desc[r][0] = node_desc[r][1] = 0xffffffffu;
        if (n0[r] < 0) continue;
        int i = ij0[r] & 0xffff, j = ij0[r] >> 16;
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
            const int n = n0[r] + cc;
            if (n < N) {
                const unsigned hl = i > 0, hr = i < dim_x - 1, hd = j > 0, hu = j < dim_y - 1;
                node_desc[r][(i + j) & 1] = (unsigned)(n >> 1) | ((unsigned)(n & 1) << 13) | ((unsigned)(n & dim_x & 1) << 14) |
                                            (hl << 15) | (hr << 16) | (hd << 17) | (hu << 18) | ((4u - hl - hr - hd - hu) << 19);
            }
            if (++i == dim_x) { i = 0; j++; }
        }
    }

    for (int g = blockIdx.x; g < a.batch; g += gridDim.x) {
        // ---- load the grid's velocity -------------------------------------------------------
        const float2 *gv = a.v + (size_t)g * N;
        uint32_t *user_c = a.c + (size_t)g * N * 3;
        for (int n = tid; n < N; n += ENS_THREADS) A[n] = __ldg(gv + n);
        if constexpr (DYE_SMEM)
            for (int n = tid; n < 3 * N; n += ENS_THREADS) C1[n] = __ldg(user_c + n);
        else
            C1 = user_c;
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
            if (a.max_drags > 0) {
                if (tid == 0) {
                    const size_t slot = (size_t)step * a.batch + g;
                    const int cnt = min(a.counts[slot], a.max_drags);
                    const fs_drag *dr = a.drags + slot * a.max_drags;
                    for (int q = 0; q < cnt; q++) {
                        const fs_drag m = dr[q];
                        if (m.cy < dim_x && m.cx < dim_y) B[m.cx * dim_x + m.cy] = make_float2(m.vy, m.vx);
                    }
                }
                __syncthreads();
            }
            // ---- divergence (ino:274) into d, zero p (poisson.cpp:117-119); A is dead now ------
            float *d = reinterpret_cast<float *>(A);    // [2][HQ] colour planes
            float *p = d + 2 * HQ;                      // [2][HQ]
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
                        const int at = cs(n, (i + j) & 1);
                        d[at] = __fmul_rn(a.k.dx, __fmul_rn(s, a.two_dx_inv));
                        p[at] = 0.0f;
                    }
                    if (++i == dim_x) { i = 0; j++; }
                }
            }
            __syncthreads();
            // ---- red-black SOR (ino:275): colour 0 = (i+j) even first ---------------------------
            // Every thread updates the colour-`parity` node of each of its pairs; everything about
            // that node that does not change between half-sweeps (slot, which neighbours exist, the
            // Gauss-Seidel coefficient) was packed into one register per (pair, colour) before the
            // loop.  ONE branch-free path for interior and wall nodes: a missing neighbour enters the
            // interior sum ((L+R)+D)+U as +0.0f, which is bit-identical to pois_gs_safe's running sum
            // (sor.cuh: sor_update_coef) — a warp of 64 consecutive nodes nearly always holds a wall
            // node, so the two-path version executed both paths for almost every warp (ncu: the SOR
            // phase was 64 % of 1,257 thread-instructions per node-step).
            for (int hs = 0; hs < 2 * a.iters; hs += 2) {
#pragma unroll
                for (int parity = 0; parity < 2; parity++) {
                    float *pm = p + parity * HQ;               // the plane being updated
                    const float *po = p + (parity ^ 1) * HQ;   // its neighbours' plane
                    const float *dm = d + parity * HQ;
#pragma unroll
                    for (int r = 0; r < ENS_MAX_ROUNDS; r++) {
                        const unsigned desc = node_desc[r][parity];
                        if (desc == 0xffffffffu) continue;     // no such node
                        const int q = desc & 0x1fff;
                        const float *ctr = po + q + ((desc >> 13) & 1);    // left neighbour at ctr[-1], right at ctr[0]
                        const float *ver = po + q + ((desc >> 14) & 1);    // down at ver[-hd0], up at ver[hu0]
                        // unconditional loads (always inside the CTA's shared memory), then selects
                        const float l = (desc & (1u << 15)) ? ctr[-1] : 0.0f, rr = (desc & (1u << 16)) ? ctr[0] : 0.0f;
                        const float dn = (desc & (1u << 17)) ? ver[-hd0] : 0.0f, up = (desc & (1u << 18)) ? ver[hu0] : 0.0f;
                        const unsigned ci = desc >> 19;        // 0: 4 neighbours, 1: 3, 2: 2
                        const float coef = ci == 0 ? a.k.neg_quarter : ci == 1 ? a.k.neg_third : a.k.neg_half;
                        pm[q] = sor_update_coef(pm[q], l, rr, dn, up, dm[q], coef, a.k);
                    }
                    __syncthreads();
                }
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
                        const int col = (i + j) & 1;
                        const float *po = p + (col ^ 1) * HQ;
                        const float pc = p[cs(n, col)];
                        const float pl = i > 0 ? po[(n - 1) >> 1] : pc, pr = i < dim_x - 1 ? po[(n + 1) >> 1] : pc;
                        const float pd = j > 0 ? po[(n - dim_x) >> 1] : pc, pu = j < dim_y - 1 ? po[(n + dim_x) >> 1] : pc;
                        float2 c0 = B[n];
                        c0.x = __fsub_rn(c0.x, __fmul_rn(__fsub_rn(pr, pl), a.two_dx_inv));
                        c0.y = __fsub_rn(c0.y, __fmul_rn(__fsub_rn(pu, pd), a.two_dx_inv));
                        B[n] = c0;
                    }
                    if (++i == dim_x) { i = 0; j++; }
                }
            }
            __syncthreads();
            // ---- advect dye, free-slip sampling (ino:282) with the projected velocity: C1 -> C2 ---------
            {
                DyeFetch fetch{C1, dim_x};
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
            __syncthreads();                // (CTA-scope ordering of the dye stores before the next step's reads)
            // pointer swaps of ino:255 and ino:286
            uint32_t *tc = C1; C1 = C2; C2 = tc;
            float2 *tv = A; A = B; B = tv;
        }

        // ---- store the grid's state ----------------------------------------------------------
        float2 *ov = a.v + (size_t)g * N;
        for (int n = tid; n < N; n += ENS_THREADS) ov[n] = A[n];
        if (C1 != user_c)                   // the final dye sits in shared memory / in the scratch slot
            for (int n = tid; n < 3 * N; n += ENS_THREADS) user_c[n] = C1[n];
        __syncthreads();
        if constexpr (!DYE_SMEM) C2 = a.scratch + (size_t)blockIdx.x * N * 3;   // next grid: C1 = its own array again
    }
}

// ---- second generation: projection in registers (ensemble_reg.cuh) -------------------------------------
struct EnsEnvDevice {
    static constexpr bool kAsync = true;     // bulk-copy pipeline available
    static constexpr bool kEmulatePipe = false;
    int tid, nthreads, block, nblocks;
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};

template <int R, int MAXT, bool DYE_SMEM>
__global__ void __launch_bounds__(MAXT, 1) ensemble_reg_kernel(const EnsArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const EnsEnvDevice env{(int)threadIdx.x, (int)blockDim.x, (int)blockIdx.x, (int)gridDim.x};
    // the pipelined flow keeps 8 dye results per thread in registers: only the 96-register build of R = 2 has room
    ens_reg_body<R, DYE_SMEM, EnsEnvDevice, (R == 2)>(a, smem_raw, env);
}

size_t ensemble_scratch_bytes(int dim_x, int dim_y, int grid) { return (size_t)grid * dim_x * dim_y * 12; }

size_t ensemble_smem_bytes(int dim_x, int dim_y, bool dye_smem)
{
    return (size_t)(dye_smem ? 40 : 16) * dim_x * dim_y + 16;
}

// first-generation kernel — variant 0: dye in shared memory while 40 B/node fit one CTA, else streamed; 1/2/3:
// streamed dye, two CTAs per SM; 4: streamed dye, one CTA per SM
static bool ens_dye_smem(int dim_x, int dim_y, int variant)
{
    return variant == 0 && ensemble_smem_bytes(dim_x, dim_y, true) <= 227 * 1024 &&
           (dim_x * dim_y + 1) / 2 <= 1024 * 3;
}

bool ensemble_supported(int dim_x, int dim_y, size_t max_smem_optin)
{
    const long long n = (long long)dim_x * dim_y;
    return dim_x < 65536 && dim_y < 32768 && n <= ENS_MAX_NODES && (size_t)(16 * n + 16) <= max_smem_optin;
}

// What a call runs.  Option "ensemble" (ctx->opt_ens): 0 = automatic (the register-tiled kernel, falling back
// to the first generation for shapes it does not take); 1-4 = first generation with streamed dye (see above),
// 5 = first generation, automatic; 6/7/8/9 = register-tiled with R = 2/4/6/8 rows per thread; 12/14/16/18 =
// the same with the dye streamed through L1/L2 instead of held in shared memory; 20 / 21 = automatic with the
// pipelined flow of the dye-resident R = 2 kernel forced on / off (default: on for calls of up to 3 steps).
struct EnsPlan {
    bool reg;          // register-tiled kernel
    int R;             // rows per thread
    bool dye_smem;
    int threads;
    size_t smem;
    int old_variant;   // first generation: its variant number
};
constexpr int ENS_REG_DEFAULT_R = 2;
constexpr int ENS_PIPE_MAX_STEPS = 3;    // measured crossover (80x60, K=10: 5.85 n + 0.47 ms pipelined against 5.41 n + 2.07 ms per call of n steps; 61x81: 3)
constexpr size_t ENS_SMEM_LIMIT = 227 * 1024;

// R = 2 is compiled twice: up to 20 warps (5 per SM sub-partition: 96 registers) and up to 21 (6 on one
// sub-partition: 80 registers) — 61x81, the reference's own grid, needs 656 threads
static int ens_reg_maxt(int R) { return R == 2 ? 672 : R == 4 ? 512 : R == 6 ? 384 : 256; }

static bool ens_reg_plan(EnsPlan &pl, int dim_x, int dim_y, int R, bool want_dye_smem)
{
    if (dim_x < 2 || dim_y < 2) return false;          // sample_interior needs a cell to clamp to
    const int ns = ens_reg_blocks(dim_x, dim_y, R);
    int threads = (ns + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    if (threads > ens_reg_maxt(R)) return false;
    pl.reg = true;
    pl.R = R;
    pl.threads = threads;
    pl.dye_smem = want_dye_smem && ens_reg_smem_bytes(dim_x, dim_y, R, true) <= ENS_SMEM_LIMIT;
    pl.smem = ens_reg_smem_bytes(dim_x, dim_y, R, pl.dye_smem);
    pl.old_variant = 0;
    return pl.smem <= ENS_SMEM_LIMIT;
}

static EnsPlan ens_plan(int dim_x, int dim_y, int variant)
{
    EnsPlan pl{};
    if (variant == 0) {
        // fewest rows per thread that fit a CTA starting from the default: more threads hide the advects' latency
        for (int R = ENS_REG_DEFAULT_R; R <= 8; R += 2)
            if (ens_reg_plan(pl, dim_x, dim_y, R, true)) return pl;
    } else if (variant >= 6 && variant <= 9) {
        if (ens_reg_plan(pl, dim_x, dim_y, 2 * (variant - 5), true)) return pl;
    } else if (variant >= 12 && variant <= 18 && variant % 2 == 0) {
        if (ens_reg_plan(pl, dim_x, dim_y, variant - 10, false)) return pl;
    }
    pl = EnsPlan{};
    pl.old_variant = (variant >= 1 && variant <= 4) ? variant : 0;
    return pl;
}

template <int R, int MAXT, bool DYE_SMEM>
static int ens_reg_occupancy(const EnsPlan &pl, int *per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(ensemble_reg_kernel<R, MAXT, DYE_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, ensemble_reg_kernel<R, MAXT, DYE_SMEM>, pl.threads, pl.smem);
}

template <bool DYE_SMEM>
static int ens_reg_occupancy_r(const EnsPlan &pl, int *per_sm)
{
    switch (pl.R) {
        case 2: return pl.threads <= 640 ? ens_reg_occupancy<2, 640, DYE_SMEM>(pl, per_sm) : ens_reg_occupancy<2, 672, DYE_SMEM>(pl, per_sm);
        case 4: return ens_reg_occupancy<4, 512, DYE_SMEM>(pl, per_sm);
        case 6: return ens_reg_occupancy<6, 384, DYE_SMEM>(pl, per_sm);
        default: return ens_reg_occupancy<8, 256, DYE_SMEM>(pl, per_sm);
    }
}

int ensemble_grid(int batch, int dim_x, int dim_y, int num_sms, int variant)
{
    if (variant == 20 || variant == 21) variant = 0;
    const EnsPlan pl = ens_plan(dim_x, dim_y, variant);
    int per_sm = 1;
    if (pl.reg) {
        // persistent CTAs walking the batch, as many per SM as registers and shared memory allow
        const int e = pl.dye_smem ? ens_reg_occupancy_r<true>(pl, &per_sm) : ens_reg_occupancy_r<false>(pl, &per_sm);
        if (e != 0 || per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
    } else if (!ens_dye_smem(dim_x, dim_y, pl.old_variant) && pl.old_variant != 4 &&
               ensemble_smem_bytes(dim_x, dim_y, false) * 2 + 2048 <= 227 * 1024) {
        // first generation: one per SM, or two when the variant streams the dye and 2 x 16N bytes fit
        per_sm = 2;
    }
    return batch < per_sm * num_sms ? batch : per_sm * num_sms;
}

template <int MAXT, int MINB, int ROUNDS, bool DYE_SMEM>
static int launch_ens(const Launch &L, const EnsArgs &a, int grid, int threads_override)
{
    const size_t smem = ensemble_smem_bytes(a.dim_x, a.dim_y, DYE_SMEM);
    cudaError_t e = cudaFuncSetAttribute(ensemble_kernel<MAXT, MINB, ROUNDS, DYE_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    // threads: ROUNDS pairs per thread, rounded up to whole warps
    const int pairs = (a.dim_x * a.dim_y + 1) / 2;
    int threads = ((pairs + ROUNDS - 1) / ROUNDS + 31) / 32 * 32;
    if (threads_override > 0 && threads_override * ROUNDS >= pairs) threads = threads_override;
    if (threads < 128) threads = 128;
    if (threads > MAXT || threads * ROUNDS < pairs) return (int)cudaErrorInvalidValue;
    ensemble_kernel<MAXT, MINB, ROUNDS, DYE_SMEM><<<grid, threads, smem, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

template <int R, int MAXT, bool DYE_SMEM>
static int launch_ens_reg(const Launch &L, const EnsArgs &a, const EnsPlan &pl, int grid)
{
    cudaError_t e = cudaFuncSetAttribute(ensemble_reg_kernel<R, MAXT, DYE_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return (int)e;
    ensemble_reg_kernel<R, MAXT, DYE_SMEM><<<grid, pl.threads, pl.smem, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

template <bool DYE_SMEM>
static int launch_ens_reg_r(const Launch &L, const EnsArgs &a, const EnsPlan &pl, int grid)
{
    switch (pl.R) {
        case 2: return pl.threads <= 640 ? launch_ens_reg<2, 640, DYE_SMEM>(L, a, pl, grid) : launch_ens_reg<2, 672, DYE_SMEM>(L, a, pl, grid);
        case 4: return launch_ens_reg<4, 512, DYE_SMEM>(L, a, pl, grid);
        case 6: return launch_ens_reg<6, 384, DYE_SMEM>(L, a, pl, grid);
        default: return launch_ens_reg<8, 256, DYE_SMEM>(L, a, pl, grid);
    }
}

int launch_ensemble(const Launch &L, float2 *v, uint32_t *c, uint32_t *scratch, const fs_drag *drags_dev,
                    const int *counts_dev, int max_drags, int batch, int dim_x, int dim_y, float dt,
                    float dx, int iters, float omega, int n_steps, int variant)
{
    if (batch <= 0 || n_steps <= 0) return 0;
    EnsArgs a;
    a.v = v; a.c = c; a.scratch = scratch; a.drags = drags_dev; a.counts = counts_dev;
    a.max_drags = max_drags; a.batch = batch; a.dim_x = dim_x; a.dim_y = dim_y;
    a.iters = iters; a.n_steps = n_steps; a.dt = dt;
    // variants 20 / 21: automatic plan with the pipelined flow forced on / off (measurement)
    a.pipe_max_steps = variant == 20 ? 0x7fffffff : variant == 21 ? 0 : ENS_PIPE_MAX_STEPS;
    if (variant == 20 || variant == 21) variant = 0;
    a.two_dx_inv = 1.0f / (2.0f * dx);
    a.k = make_sor_coef(dx, omega);
    const int grid = ensemble_grid(batch, dim_x, dim_y, L.num_sms, variant);
    const EnsPlan pl = ens_plan(dim_x, dim_y, variant);
    if (pl.reg) return pl.dye_smem ? launch_ens_reg_r<true>(L, a, pl, grid) : launch_ens_reg_r<false>(L, a, pl, grid);
    if (ens_dye_smem(dim_x, dim_y, pl.old_variant)) return launch_ens<1024, 1, 3, true>(L, a, grid, 0);   // everything resident
    switch (pl.old_variant) {
        case 1: return launch_ens<512, 2, 6, false>(L, a, grid, 0);       // two grids per SM, 6 pairs per thread
        case 2: return launch_ens<512, 2, 6, false>(L, a, grid, 512);     // two grids per SM, 16 warps each
        case 3: return launch_ens<512, 2, 6, false>(L, a, grid, 384);     // two grids per SM, 12 warps each
        case 4: return launch_ens<1024, 1, 3, false>(L, a, grid, 0);      // one grid per SM, dye streamed
        default: return launch_ens<512, 2, 6, false>(L, a, grid, 0);      // (too large for 40 B/node)
    }
}

}  // namespace fs
