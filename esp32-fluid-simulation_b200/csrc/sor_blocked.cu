// Temporally-blocked red-black SOR (poisson.cpp:14-125): several colour half-
// sweeps per HBM round trip.
//
// One CTA owns a 128 x (R*NW)-node REGION of the grid = its output tile plus a
// halo of H nodes (H = number of fused half-sweeps).  The region lives in
// REGISTERS for the whole pass:
//
//   * a warp owns a 128-column x R-row strip; lane t owns columns 4t..4t+3 of
//     every row of the strip (p and dx*d: 8R registers);
//   * vertical neighbours of a node are in the same thread's registers;
//   * of the two horizontal neighbours, one is in the same thread and the other
//     one comes from the adjacent lane by ONE warp shuffle per row;
//   * only the strips' first/last rows cross warps — they are exchanged through
//     a 16 KB double-buffered shared-memory mailbox, one __syncthreads per
//     half-sweep.
//
// Each half-sweep updates every node of one colour in the region from the other
// colour, exactly as the sequential sweep would; nodes closer than s to a region
// edge that is not a domain wall are stale after s half-sweeps, which is why only
// the tile interior (>= H from those edges) is written back.  Redundant halo work
// aside, every stored value is computed by the same operations in the same order
// as poisson.cpp — the result is bit-identical to the reference for any H.
//
// Two loaders feed the same sweep code:
//   sor_blocked_kernel      one tile per CTA, region read straight from global into registers
//                           (any pitch/alignment; the fallback);
//   sor_blocked_tma_kernel  persistent CTAs; while tile k is swept out of registers, the bulk-
//                           tensor loads (TMA, SASS UTMALDG) of tile k+1's p and d regions land in
//                           shared memory, hiding the DRAM latency behind 2T half-sweeps.
//
// (A packed-f32x2 version of the sweep — FADD2/FMUL2, two nodes per instruction — was built and
// measured: bit-exact but 15 % SLOWER on B200 (packed ops issue at half rate and the pack/unpack
// moves serialise the row chains); profiles/r01_kernel_sweep_packed_f32x2_rejected.json.  Also note
// that ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false.)
//
// Traffic per pass: read p + d over the region, write p over the tile, i.e.
// ~(8*redundancy + 4) B/node for H/2 full iterations instead of 12 B per
// iteration per colour sector.
#include "kernels.h"
#include "sor.cuh"
#include "tma.cuh"

namespace fs {

constexpr int BLK_RW = 128;  // region width in nodes (32 lanes x 4 columns)

struct BlockedArgs {
    float *p_out;
    const float *p_in;   // nullptr = all zero (poisson.cpp:117-119)
    const float *div;
    Geo g;
    SorCoef k;
    int first_parity;    // global colour of the first half-sweep
    int n_half;          // fused half-sweeps H
    int hpx, hpy;        // halo in x (multiple of 4, >= H) and y (= H)
    int lax, lay;        // local coordinate of tile (0,0)'s first output node
    int tw_out, th_out;  // output tile size
    int vec_ok;          // rows are 16-byte aligned: float4 loads/stores allowed
    int early_trigger;   // persistent kernels: let the next pass's CTAs take freed SMs early (PDL)
};

// One colour over a warp's strip.  Q = colour offset inside the strip: row r
// updates columns {0,2} when (r+Q) is even and {1,3} when it is odd.
//
// WALL: 0 = the region lies strictly inside the domain: pois_sor_fast everywhere.
// Regions that touch a domain wall or stick out of the domain are handled WITHOUT per-node
// branches: every cell outside the domain holds +0.0f for the whole pass (zero-filled by the
// loader, never updated), so the interior sum ((L + R) + D) + U of a wall node is bit-identical to
// pois_gs_safe's sum over its existing neighbours (sor.cuh: sor_update_coef), and the only per-node
// difference left is the coefficient neg_a_ii_inv[#neighbours]:
//   1 = only y-walls in reach (CTA-uniform): the coefficient is a row-uniform value, rows outside
//       the domain are skipped — costs a compare and a branch per row;
//   2 = x-walls in reach: additionally a per-column select of the coefficient and of "is a node".
// (The first version branched per node into sor_update_wall: wall tiles cost 3x an interior tile
// and 24 % of the kernel's instructions.)
template <int R, int Q, int WALL>
__device__ __forceinline__ void strip_half_sweep(float (&p)[R][4], const float (&dxd)[R][4],
                                                 const float (&dn)[4], const float (&up)[4],
                                                 const SorCoef &k, int gi0, int gj0, int GX, int GY)
{
    bool node[4], wallc[4];
    if constexpr (WALL == 2) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            node[c] = (unsigned)(gi0 + c) < (unsigned)GX;
            wallc[c] = gi0 + c == 0 || gi0 + c == GX - 1;
        }
    }
    auto upd = [&](int r, int c, float l, float rr, float d, float u, float c_in, float c_wall) {
        if constexpr (WALL == 2) {
            const float n = sor_update_coef(p[r][c], l, rr, d, u, dxd[r][c], wallc[c] ? c_wall : c_in, k);
            p[r][c] = node[c] ? n : p[r][c];
        } else if constexpr (WALL == 1) {
            p[r][c] = sor_update_coef(p[r][c], l, rr, d, u, dxd[r][c], c_in, k);
        } else {
            p[r][c] = sor_update_interior(p[r][c], l, rr, d, u, dxd[r][c], k);
        }
    };
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int gj = gj0 + r;
        float c_in = k.neg_quarter, c_wall = k.neg_third;
        if constexpr (WALL != 0) {
            if ((unsigned)gj >= (unsigned)GY) continue;          // not a row of the domain (warp-uniform)
            const bool wallr = gj == 0 || gj == GY - 1;
            c_in = wallr ? k.neg_third : k.neg_quarter;          // 3 or 4 neighbours
            c_wall = wallr ? k.neg_half : k.neg_third;           // 2 (corner) or 3
        }
        if (((r + Q) & 1) == 0) {
            const float lft = __shfl_up_sync(0xffffffffu, p[r][3], 1);
            const float d0 = r > 0 ? p[r > 0 ? r - 1 : 0][0] : dn[0], u0 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][0] : up[0];
            const float d2 = r > 0 ? p[r > 0 ? r - 1 : 0][2] : dn[2], u2 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][2] : up[2];
            const float c1 = p[r][1], c3 = p[r][3];
            upd(r, 0, lft, c1, d0, u0, c_in, c_wall);
            upd(r, 2, c1, c3, d2, u2, c_in, c_wall);
        } else {
            const float rgt = __shfl_down_sync(0xffffffffu, p[r][0], 1);
            const float d1 = r > 0 ? p[r > 0 ? r - 1 : 0][1] : dn[1], u1 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][1] : up[1];
            const float d3 = r > 0 ? p[r > 0 ? r - 1 : 0][3] : dn[3], u3 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][3] : up[3];
            const float c0 = p[r][0], c2 = p[r][2];
            upd(r, 1, c0, c2, d1, u1, c_in, c_wall);
            upd(r, 3, c2, rgt, d3, u3, c_in, c_wall);
        }
    }
}

// barrier over the NW compute warps only (named barrier 1): the persistent kernels run one more
// warp — the producer — that must not take part in the sweeps' hand-offs
template <int NW>
__device__ __forceinline__ void compute_barrier()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
}

// all H half-sweeps of one pass over a warp's strip, with the inter-warp row mailbox
template <int R, int NW, int WALL>
__device__ __forceinline__ void sweep_pass(float (&p)[R][4], const float (&dxd)[R][4], const BlockedArgs &a,
                                           float4 (*mail)[NW][2][32], int gi0, int gj0, int n_half)
{
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    // colour bookkeeping: node (column c, row r) of this strip has global parity (c + r + pb) & 1
    const int pb = (gi0 + gj0) & 1;
    float dn[4] = {0.f, 0.f, 0.f, 0.f}, up[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < n_half; s++) {
        const int buf = s & 1;
        mail[buf][w][0][t] = make_float4(p[0][0], p[0][1], p[0][2], p[0][3]);
        mail[buf][w][1][t] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
        compute_barrier<NW>();
        if (w > 0) {
            const float4 q = mail[buf][w - 1][1][t];
            dn[0] = q.x; dn[1] = q.y; dn[2] = q.z; dn[3] = q.w;
        }
        if (w < NW - 1) {
            const float4 q = mail[buf][w + 1][0][t];
            up[0] = q.x; up[1] = q.y; up[2] = q.z; up[3] = q.w;
        }
        const int q_eff = (a.first_parity + s + pb) & 1;
        if (q_eff == 0) strip_half_sweep<R, 0, WALL>(p, dxd, dn, up, a.k, gi0, gj0, a.g.GX, a.g.GY);
        else            strip_half_sweep<R, 1, WALL>(p, dxd, dn, up, a.k, gi0, gj0, a.g.GX, a.g.GY);
    }
}

// write back the tile interior, clipped to the compute rectangle; with PUSH, also straight into the
// ghost regions of the neighbouring ranks that need it (peer memory, 16-byte stores)
template <int R, bool PUSH = false>
__device__ __forceinline__ void store_tile(const float (&p)[R][4], const BlockedArgs &a, float *p_out, int rlx0,
                                           int rly0, int lx0, int ly0, const SorPushArgs *push = nullptr)
{
    const Geo &g = a.g;
    const int ox0 = max(rlx0 + a.hpx, g.x0), ox1 = min(rlx0 + a.hpx + a.tw_out, g.x1);
    const int oy0 = max(rly0 + a.hpy, g.y0), oy1 = min(rly0 + a.hpy + a.th_out, g.y1);
    const bool cols_full = lx0 >= ox0 && lx0 + 3 < ox1;
    unsigned peers = 0;                                     // CTA-uniform: strips this tile's output meets
    if constexpr (PUSH) {
        for (int k = 0; k < push->n_peers; k++) {
            const SorPushPeer &q = push->peer[k];
            if (ox0 < q.sx1 && ox1 > q.sx0 && oy0 < q.sy1 && oy1 > q.sy0) peers |= 1u << k;
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ly = ly0 + r;
        if (ly < oy0 || ly >= oy1) continue;
        float *row = p_out + (size_t)ly * g.nx;
        if (cols_full && a.vec_ok) {
            const float4 val = make_float4(p[r][0], p[r][1], p[r][2], p[r][3]);
            *reinterpret_cast<float4 *>(row + lx0) = val;
            if constexpr (PUSH) {
                for (unsigned m = peers; m; m &= m - 1) {
                    const SorPushPeer &q = push->peer[__ffs(m) - 1];
                    if (ly >= q.sy0 && ly < q.sy1 && lx0 >= q.sx0 && lx0 < q.sx1)
                        *reinterpret_cast<float4 *>(q.base + (size_t)(ly + q.dy) * q.pitch + (lx0 + q.dx)) = val;
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (lx0 + c >= ox0 && lx0 + c < ox1) row[lx0 + c] = p[r][c];
        }
    }
}

// a tile is a RIM tile when its region crosses a side of the rectangle that faces another rank:
// it reads ghosts and/or produces values a neighbour needs (CTA-uniform; same test on the host)
template <int R, int NW>
__host__ __device__ __forceinline__ bool region_is_rim(const Geo &g, const SorPushArgs &q, int rlx0, int rly0)
{
    return (q.has_l && rlx0 < g.x0) || (q.has_r && rlx0 + BLK_RW > g.x1) || (q.has_d && rly0 < g.y0) ||
           (q.has_u && rly0 + R * NW > g.y1);
}

__device__ __forceinline__ unsigned long long halo_global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// warp 0: wait until every neighbour's flag carries `seq` (their strips are in our ghosts).  Lane k polls
// neighbour k's flag, so the system-scope acquire loads (~1-2 us each) overlap instead of queueing up
// behind each other on one thread.
__device__ __forceinline__ void wait_for_neighbours(const SorPushArgs &q)
{
    const int lane = threadIdx.x & 31;
    if (lane < q.n_wait) {
        const unsigned long long t0 = halo_global_ns();
        for (;;) {
            if (q.status && *reinterpret_cast<volatile int *>(q.status) == FS_ERR_HALO_TIMEOUT) break;   // already given up
            unsigned long long seen;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(q.wait[lane]) : "memory");
            if (seen >= q.seq_wait) break;
            if (q.timeout_ns && halo_global_ns() - t0 > q.timeout_ns) {
                if (q.status) atomicExch(q.status, FS_ERR_HALO_TIMEOUT);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");   // the neighbours' generic stores -> this CTA's TMA reads
}

// Work order of the persistent kernel: the frame of tiles along the window's edge first (they
// are the ones that can touch a wall and take the slower path), then the interior, row by row.
__device__ __forceinline__ void tile_coords(int idx, int ntx, int nty, int &tx, int &ty)
{
    if (ntx < 3 || nty < 3) { tx = idx % ntx; ty = idx / ntx; return; }
    const int frame = 2 * ntx + 2 * (nty - 2);
    if (idx < ntx) { tx = idx; ty = 0; }
    else if (idx < 2 * ntx) { tx = idx - ntx; ty = nty - 1; }
    else if (idx < frame) { const int k = idx - 2 * ntx; tx = (k & 1) ? ntx - 1 : 0; ty = 1 + (k >> 1); }
    else { const int k = idx - frame; tx = 1 + k % (ntx - 2); ty = 1 + k / (ntx - 2); }
}

__device__ __forceinline__ void rowmajor_coords(int idx, int ntx, int &tx, int &ty)
{
    ty = idx / ntx;
    tx = idx - ty * ntx;
}

// does a region touch a domain wall or stick out of the domain?  (CTA-uniform)  0 = no,
// 1 = y-walls only, 2 = x-walls (and possibly y-walls): the WALL mode of strip_half_sweep
template <int R, int NW>
__device__ __forceinline__ int region_wall_mode(const BlockedArgs &a, int rlx0, int rly0)
{
    const int rgx0 = a.g.ox + rlx0, rgy0 = a.g.oy + rly0;
    if (rgx0 <= 0 || rgx0 + BLK_RW >= a.g.GX) return 2;
    return (rgy0 <= 0 || rgy0 + R * NW >= a.g.GY) ? 1 : 0;
}

template <int R, int NW>
__device__ __forceinline__ void sweep_region(float (&p)[R][4], const float (&dxd)[R][4], const BlockedArgs &a,
                                             float4 (*mail)[NW][2][32], int rlx0, int rly0, int gi0, int gj0,
                                             int n_half)
{
    const int mode = region_wall_mode<R, NW>(a, rlx0, rly0);
    if (mode == 0)      sweep_pass<R, NW, 0>(p, dxd, a, mail, gi0, gj0, n_half);
    else if (mode == 1) sweep_pass<R, NW, 1>(p, dxd, a, mail, gi0, gj0, n_half);
    else                sweep_pass<R, NW, 2>(p, dxd, a, mail, gi0, gj0, n_half);
}

// ---- loader 1: one tile per CTA, region read straight from global into registers ---------------
template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) sor_blocked_kernel(const BlockedArgs a)
{
    __shared__ float4 mail[2][NW][2][32];
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const Geo &g = a.g;
    const int rlx0 = a.lax + (int)blockIdx.x * a.tw_out - a.hpx;
    const int rly0 = a.lay + (int)blockIdx.y * a.th_out - a.hpy;
    const int lx0 = rlx0 + 4 * t, ly0 = rly0 + w * R;
    const int gi0 = g.ox + lx0, gj0 = g.oy + ly0;

    float p[R][4], dxd[R][4];
    const bool cols_in = lx0 >= 0 && lx0 + 3 < g.nx;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ly = ly0 + r;
        const bool row_in = ly >= 0 && ly < g.ny;
        const size_t base = (size_t)(row_in ? ly : 0) * g.nx;
        if (row_in && cols_in && a.vec_ok) {
            const float4 dv = __ldg(reinterpret_cast<const float4 *>(a.div + base + lx0));
            dxd[r][0] = __fmul_rn(a.k.dx, dv.x);
            dxd[r][1] = __fmul_rn(a.k.dx, dv.y);
            dxd[r][2] = __fmul_rn(a.k.dx, dv.z);
            dxd[r][3] = __fmul_rn(a.k.dx, dv.w);
            if (a.p_in) {
                const float4 pv = __ldg(reinterpret_cast<const float4 *>(a.p_in + base + lx0));
                p[r][0] = pv.x; p[r][1] = pv.y; p[r][2] = pv.z; p[r][3] = pv.w;
            } else {
                p[r][0] = p[r][1] = p[r][2] = p[r][3] = 0.0f;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int lx = lx0 + c;
                const bool in = row_in && lx >= 0 && lx < g.nx;
                dxd[r][c] = in ? __fmul_rn(a.k.dx, __ldg(a.div + base + lx)) : 0.0f;
                p[r][c] = (in && a.p_in) ? __ldg(a.p_in + base + lx) : 0.0f;
            }
        }
    }
    sweep_region<R, NW>(p, dxd, a, mail, rlx0, rly0, gi0, gj0, a.n_half);
    store_tile<R>(p, a, a.p_out, rlx0, rly0, lx0, ly0);
}

// ---- loader 2: persistent CTAs, next tile prefetched into shared memory by TMA -----------------
// The hardware zero-fills the parts of a region outside the window — exactly what loader 1 does by
// hand.
#ifdef FS_SOR_PROF
// development aid (build with FS_NVCC_EXTRA=-DFS_SOR_PROF): per-CTA cycle sums of the tile phases
__device__ unsigned long long g_sor_prof[1024][8];
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(slot, a, b) do { if (threadIdx.x == 0) g_sor_prof[blockIdx.x][slot] += (unsigned long long)((b) - (a)); } while (0)
#else
#define PROF_T(v)
#define PROF_ADD(slot, a, b)
#endif

template <int R, int NW, int MINB, bool PUSH>
__device__ __forceinline__ void sor_blocked_tma_body(const CUtensorMap &p_map, const CUtensorMap &d_map,
                                                     const BlockedArgs &a, int ntx, int nty, int *work_counter,
                                                     const SorPushArgs *push)
{
    constexpr int RH = R * NW;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *sd = reinterpret_cast<float4 *>(smem);                        // [RH][32] float4 = d region
    float4 *sp = sd + RH * 32;                                            // p region
    float4 (*mail)[NW][2][32] = reinterpret_cast<float4 (*)[NW][2][32]>(sp + RH * 32);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_tile[2];
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const Geo &g = a.g;
    const bool has_p = a.p_in != nullptr;
    const uint32_t tx_bytes = (has_p ? 2u : 1u) * RH * BLK_RW * 4u;
    const int n_tiles = ntx * nty;

    // thread 0: tiles are claimed dynamically (wall tiles cost more than interior ones) TWO ahead:
    // the ticket used for a prefetch was drawn one tile earlier, so the atomic's L2 round trip
    // (~1 us) is never waited for in front of the tile's first barrier.
    // (A dedicated producer warp was built and measured in round 2: 17 warps cap the kernel at 120
    // registers per thread, the sweeps spill, and the solve got 20-60 % SLOWER.)
    auto prefetch = [&](int idx) {
        if (idx < n_tiles) {
            int tx, ty;
            tile_coords(idx, ntx, nty, tx, ty);
            const int x = a.lax + tx * a.tw_out - a.hpx, y = a.lay + ty * a.th_out - a.hpy;
            mbar_expect_tx(&bar, tx_bytes);
            tma_load_2d(sd, &d_map, x, y, &bar);
            if (has_p) tma_load_2d(sp, &p_map, x, y, &bar);
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // Programmatic dependent launch, both directions.  (1) Let the NEXT pass's CTAs take this SM the
    // moment this CTA exits: their prologue and their first d load then overlap this pass's tail.
    if (a.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int ahead = n_tiles;                       // thread 0 only: the ticket drawn one tile early
    if (w == 0) {
        // (the tile counters of all passes were zeroed before the solve's first pass)
        int first = n_tiles, fx = 0, fy = 0;
        if (t == 0) {
            first = atomicAdd(work_counter, 1);
            s_tile[0] = first;
            if (first < n_tiles) {
                int tx, ty;
                tile_coords(first, ntx, nty, tx, ty);
                fx = a.lax + tx * a.tw_out - a.hpx;
                fy = a.lay + ty * a.th_out - a.hpy;
                mbar_expect_tx(&bar, tx_bytes);
                // (2) this grid may have been scheduled while the previous pass is still draining: d does
                // not depend on it (pass >= 2) — load it now; p_in is the previous pass's output and must wait
                if (has_p) tma_load_2d(sd, &d_map, fx, fy, &bar);
            }
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if constexpr (PUSH) {
            if (push->n_wait > 0) wait_for_neighbours(*push);       // p_in's ghosts are the neighbours' previous pass
        }
        if (t == 0) {
            if (first < n_tiles) {
                if (has_p) tma_load_2d(sp, &p_map, fx, fy, &bar);
                else       tma_load_2d(sd, &d_map, fx, fy, &bar);   // pass 1: d is the previous KERNEL's output
            }
            ahead = atomicAdd(work_counter, 1);
        }
    }
    __syncthreads();
    uint32_t phase = 0;
    PROF_T(t_begin);
    for (int it = 0;; it++) {
        const int idx = s_tile[it & 1];
        if (idx >= n_tiles) break;
        int tx, ty;
        tile_coords(idx, ntx, nty, tx, ty);
        const int rlx0 = a.lax + tx * a.tw_out - a.hpx;
        const int rly0 = a.lay + ty * a.th_out - a.hpy;
        const int lx0 = rlx0 + 4 * t, ly0 = rly0 + w * R;
        const int gi0 = g.ox + lx0, gj0 = g.oy + ly0;

        PROF_T(t0);
        mbar_wait(&bar, phase);
        PROF_T(t1);
        phase ^= 1;
        float p[R][4], dxd[R][4];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 dv = sd[(w * R + r) * 32 + t];
            dxd[r][0] = __fmul_rn(a.k.dx, dv.x);
            dxd[r][1] = __fmul_rn(a.k.dx, dv.y);
            dxd[r][2] = __fmul_rn(a.k.dx, dv.z);
            dxd[r][3] = __fmul_rn(a.k.dx, dv.w);
            if (has_p) {
                const float4 pv = sp[(w * R + r) * 32 + t];
                p[r][0] = pv.x; p[r][1] = pv.y; p[r][2] = pv.z; p[r][3] = pv.w;
            } else {
                p[r][0] = p[r][1] = p[r][2] = p[r][3] = 0.0f;
            }
        }
        __syncthreads();                       // every warp has drained the staging buffers
        PROF_T(t2);
        if (threadIdx.x == 0) {
            s_tile[(it + 1) & 1] = ahead;      // visible after sweep_pass's first barrier
            prefetch(ahead);
            ahead = atomicAdd(work_counter, 1);   // consumed one tile later
        }
        PROF_T(t3);
        sweep_region<R, NW>(p, dxd, a, mail, rlx0, rly0, gi0, gj0, a.n_half);
        PROF_T(t4);
        store_tile<R, PUSH>(p, a, a.p_out, rlx0, rly0, lx0, ly0, push);
        PROF_T(t5);
        PROF_ADD(0, t0, t1);   // wait for the TMA
        PROF_ADD(1, t1, t2);   // staging -> registers + barrier
        PROF_ADD(2, t2, t3);   // issue the next prefetch
        PROF_ADD(3, t3, t4);   // sweeps
        PROF_ADD(4, t4, t5);   // store
        PROF_ADD(5, 0, 1);     // tiles
        PROF_ADD(6, t_begin, t5);
        if constexpr (PUSH) {
            if (push->n_peers > 0 && region_is_rim<R, NW>(g, *push, rlx0, rly0)) {
                // publish: the CTA that completes the LAST rim tile tells every neighbour that this
                // rank's strips are in their ghosts — interior tiles keep running underneath.
                // Ordering (PTX memory model, causality order is cumulative): this CTA's peer stores
                // happen-before thread 0's acq_rel increment (bar.sync, then a gpu-scope release); the
                // last incrementer acquires every earlier increment and then RELEASES the flags at
                // system scope, so a neighbour that acquires a flag sees all rim tiles' stores.  No
                // per-thread system fence: it stalled all 16 warps for microseconds on every rim tile.
                __syncthreads();
                if (threadIdx.x == 0) {
                    int prev;
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(prev) : "l"(push->rim_done) : "memory");
                    if (prev == push->rim_total - 1) {
                        for (int k = 0; k < push->n_peers; k++)
                            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(push->signal[k]), "l"(push->seq_signal)
                                         : "memory");
                    }
                }
            }
        }
    }
}

#ifdef FS_SOR_PROF
extern "C" int fs_debug_sor_prof(unsigned long long *out, int reset)
{
    if (out) cudaMemcpyFromSymbol(out, g_sor_prof, sizeof(g_sor_prof));
    if (reset) {
        static unsigned long long zero[1024][8];
        cudaMemcpyToSymbol(g_sor_prof, zero, sizeof(zero));
    }
    return 0;
}
#endif

template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
sor_blocked_tma_kernel(const __grid_constant__ CUtensorMap p_map, const __grid_constant__ CUtensorMap d_map,
                       const BlockedArgs a, int ntx, int nty, int *work_counter)
{
    sor_blocked_tma_body<R, NW, MINB, false>(p_map, d_map, a, ntx, nty, work_counter, nullptr);
}

// the same pass fused with its halo exchange (decomposed grids; see SorPushArgs in kernels.h)
template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
sor_blocked_push_kernel(const __grid_constant__ CUtensorMap p_map, const __grid_constant__ CUtensorMap d_map,
                        const BlockedArgs a, int ntx, int nty, int *work_counter,
                        const __grid_constant__ SorPushArgs push)
{
    sor_blocked_tma_body<R, NW, MINB, true>(p_map, d_map, a, ntx, nty, work_counter, &push);
}

// ---- loader 3: the WHOLE solve in one persistent launch ------------------------------------------
// Work item w = (pass k, tile): claimed in order from one atomic counter; tiles in ROW-MAJOR order
// inside a pass, so the items a tile depends on (rows <= ty+1 of the previous pass) were claimed
// about a whole pass earlier and are long complete — only the very first launch-wide wave can wait.
// (Wall-tiles-first, the order of the per-pass kernel, made the bottom frame row of pass k wait for
// the LAST interior row of pass k-1 and idled dozens of SMs at every pass boundary: 1.08 ms.)  A tile of pass k >= 1
// reads p over its region from the buffer pass k-1 wrote, so it depends on the 3x3 neighbourhood of
// tiles of pass k-1 (region = tile + halo <= one tile on every side); the same wait also covers the
// write-after-read hazard of the ping-pong (pass k writes the buffer pass k-1 read: its readers of
// this tile's area are exactly those 3x3 neighbours).  Completion is published per tile with a
// release store of the solve's generation number and observed with acquire loads by the thread
// that issues the TMA prefetch.  The prefetch of the NEXT item is issued early if its dependencies
// are already met, otherwise after this CTA has published its own tile (never blocking before its
// own work is done: every dependency has a smaller work index, so waits cannot form a cycle).
// No global barrier between passes: launch gaps, pipeline ramps and per-pass tails disappear.
// MEASURED (4096^2, K=50): 0.98-1.11 ms against 0.87-0.88 ms for one launch per pass — the nine
// ld.acquire.gpu probes (~3 us, serialised by their acquire semantics) and the release fence run on
// thread 0, which is also a compute lane, so every warp waits for them at the tile's first barrier;
// that costs more than the gaps and tails it removes.  Kept as an option ("sor_one_launch"), not
// the default; a dedicated producer warp would be the next thing to try.
//   profiles/r01_kernel_sweep_sor_one_launch.json
struct SolveArgs {
    BlockedArgs a;            // geometry + coefficients; n_half = half-sweeps of a regular pass
    float *buf[2];            // buf[k & 1] is written by pass k
    int passes, n_half_last;
    int ntx, nty;
    int *work_counter;        // zeroed before the launch
    unsigned int *done;       // [passes][ntx*nty] generation stamps
    unsigned int gen;
};

template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
sor_solve_tma_kernel(const __grid_constant__ CUtensorMap d_map, const __grid_constant__ CUtensorMap p0_map,
                     const __grid_constant__ CUtensorMap p1_map, const SolveArgs sa)
{
    constexpr int RH = R * NW;
    const BlockedArgs &a = sa.a;
    extern __shared__ __align__(128) unsigned char smem[];
    float4 *sd = reinterpret_cast<float4 *>(smem);
    float4 *sp = sd + RH * 32;
    float4 (*mail)[NW][2][32] = reinterpret_cast<float4 (*)[NW][2][32]>(sp + RH * 32);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_item[2];
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const Geo &g = a.g;
    const int n_tiles = sa.ntx * sa.nty, total = sa.passes * n_tiles;
    const uint32_t region_bytes = RH * BLK_RW * 4u;

    auto deps_ready = [&](int item) -> bool {
        const int k = item / n_tiles;
        if (k == 0) return true;
        int tx, ty;
        rowmajor_coords(item - k * n_tiles, sa.ntx, tx, ty);
        const unsigned int *flags = sa.done + (size_t)(k - 1) * n_tiles;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                const int x = tx + dx, y = ty + dy;
                if (x < 0 || x >= sa.ntx || y < 0 || y >= sa.nty) continue;
                unsigned int seen;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + y * sa.ntx + x) : "memory");
                if (seen != sa.gen) return false;
            }
        return true;
    };
    auto issue = [&](int item) {
        const int k = item / n_tiles;
        int tx, ty;
        rowmajor_coords(item - k * n_tiles, sa.ntx, tx, ty);
        const int x = a.lax + tx * a.tw_out - a.hpx, y = a.lay + ty * a.th_out - a.hpy;
        asm volatile("fence.proxy.async;" ::: "memory");   // other CTAs' generic stores -> this CTA's async-proxy reads
        mbar_expect_tx(&bar, k > 0 ? 2u * region_bytes : region_bytes);
        tma_load_2d(sd, &d_map, x, y, &bar);
        if (k > 0) tma_load_2d(sp, ((k - 1) & 1) ? &p1_map : &p0_map, x, y, &bar);
    };

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const int first = atomicAdd(sa.work_counter, 1);
        s_item[0] = first;
        if (first < total) {
            while (!deps_ready(first)) __nanosleep(100);
            issue(first);
        }
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int it = 0;; it++) {
        const int item = s_item[it & 1];
        if (item >= total) break;
        const int k = item / n_tiles, idx = item - k * n_tiles;
        int tx, ty;
        rowmajor_coords(idx, sa.ntx, tx, ty);
        const int rlx0 = a.lax + tx * a.tw_out - a.hpx;
        const int rly0 = a.lay + ty * a.th_out - a.hpy;
        const int lx0 = rlx0 + 4 * t, ly0 = rly0 + w * R;
        const int gi0 = g.ox + lx0, gj0 = g.oy + ly0;

        mbar_wait(&bar, phase);
        phase ^= 1;
        float p[R][4], dxd[R][4];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 dv = sd[(w * R + r) * 32 + t];
            dxd[r][0] = __fmul_rn(a.k.dx, dv.x);
            dxd[r][1] = __fmul_rn(a.k.dx, dv.y);
            dxd[r][2] = __fmul_rn(a.k.dx, dv.z);
            dxd[r][3] = __fmul_rn(a.k.dx, dv.w);
            if (k > 0) {
                const float4 pv = sp[(w * R + r) * 32 + t];
                p[r][0] = pv.x; p[r][1] = pv.y; p[r][2] = pv.z; p[r][3] = pv.w;
            } else {
                p[r][0] = p[r][1] = p[r][2] = p[r][3] = 0.0f;
            }
        }
        __syncthreads();                       // every warp has drained the staging buffers
        int next = total;
        bool issued = true;
        if (threadIdx.x == 0) {
            next = atomicAdd(sa.work_counter, 1);
            s_item[(it + 1) & 1] = next;       // visible after sweep_pass's first barrier
            issued = next >= total;
            if (!issued && deps_ready(next)) {
                issue(next);
                issued = true;
            }
        }
        const int n_half = k == sa.passes - 1 ? sa.n_half_last : a.n_half;
        sweep_region<R, NW>(p, dxd, a, mail, rlx0, rly0, gi0, gj0, n_half);
        store_tile<R>(p, a, sa.buf[k & 1], rlx0, rly0, lx0, ly0);
        __syncthreads();                       // the whole tile has been stored
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sa.done + (size_t)k * n_tiles + ty * sa.ntx + tx),
                         "r"(sa.gen)
                         : "memory");
            if (!issued) {                     // the next item waits on tiles still in flight (pass boundary)
                while (!deps_ready(next)) __nanosleep(100);
                issue(next);
            }
        }
    }
}

template <int R, int NW>
static bool tile_cfg(BlockedArgs &a, int &ntx, int &nty)
{
    const int H = a.n_half;
    a.hpx = (H + 3) & ~3;
    a.hpy = H;
    a.tw_out = BLK_RW - 2 * a.hpx;
    a.th_out = R * NW - 2 * a.hpy;
    a.lax = a.g.x0 & ~3;  // keep every region origin a multiple of 4 columns (float4 alignment)
    a.lay = a.g.y0;
    if (a.tw_out <= 0 || a.th_out <= 0) return false;
    ntx = (a.g.x1 - a.lax + a.tw_out - 1) / a.tw_out;
    nty = (a.g.y1 - a.lay + a.th_out - 1) / a.th_out;
    return true;
}

template <int R, int NW, int MINB>
static int launch_cfg(const Launch &L, BlockedArgs &a)
{
    int ntx, nty;
    if (!tile_cfg<R, NW>(a, ntx, nty)) return (int)cudaErrorInvalidValue;
    if (ntx <= 0 || nty <= 0) return 0;
    sor_blocked_kernel<R, NW, MINB><<<dim3(ntx, nty), 32 * NW, 0, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

template <int R, int NW, int MINB>
static int launch_cfg_tma(const Launch &L, BlockedArgs &a, int *work_counter)
{
    const Geo &g = a.g;
    int ntx, nty;
    if (!tile_cfg<R, NW>(a, ntx, nty)) return (int)cudaErrorInvalidValue;
    if (ntx <= 0 || nty <= 0) return 0;
    CUtensorMap p_map, d_map;
    if (!tma_make_map_2d(&d_map, a.div, g.nx, g.ny, g.nx, BLK_RW, R * NW)) return (int)cudaErrorInvalidValue;
    p_map = d_map;
    if (a.p_in && !tma_make_map_2d(&p_map, a.p_in, g.nx, g.ny, g.nx, BLK_RW, R * NW))
        return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * R * NW * BLK_RW * 4 + sizeof(float4) * 2 * NW * 2 * 32;
    cudaError_t e = cudaFuncSetAttribute(sor_blocked_tma_kernel<R, NW, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = ntx * nty;
    const int grid = n_tiles < MINB * L.num_sms ? n_tiles : MINB * L.num_sms;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * NW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: overlap this launch with the previous pass's tail
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    a.early_trigger = 1;
    e = cudaLaunchKernelEx(&cfg, sor_blocked_tma_kernel<R, NW, MINB>, p_map, d_map, a, ntx, nty, work_counter);
    ++*L.launches;
    return (int)e;
}

template <int R, int NW, int MINB>
static int launch_cfg_push(const Launch &L, BlockedArgs &a, int *work_counter, SorPushArgs &push, int grid_limit)
{
    const Geo &g = a.g;
    int ntx, nty;
    if (!tile_cfg<R, NW>(a, ntx, nty)) return (int)cudaErrorInvalidValue;
    if (ntx <= 0 || nty <= 0) return (int)cudaErrorInvalidValue;   // an empty rectangle cannot take part in an exchange
    CUtensorMap p_map, d_map;
    if (!tma_make_map_2d(&d_map, a.div, g.nx, g.ny, g.nx, BLK_RW, R * NW)) return (int)cudaErrorInvalidValue;
    p_map = d_map;
    if (a.p_in && !tma_make_map_2d(&p_map, a.p_in, g.nx, g.ny, g.nx, BLK_RW, R * NW))
        return (int)cudaErrorInvalidValue;
    // rim tiles: same test as the kernel's
    push.rim_total = 0;
    for (int ty = 0; ty < nty; ty++)
        for (int tx = 0; tx < ntx; tx++)
            push.rim_total += region_is_rim<R, NW>(g, push, a.lax + tx * a.tw_out - a.hpx, a.lay + ty * a.th_out - a.hpy);
    if (push.n_peers > 0 && push.rim_total == 0) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * R * NW * BLK_RW * 4 + sizeof(float4) * 2 * NW * 2 * 32;
    cudaError_t e = cudaFuncSetAttribute(sor_blocked_push_kernel<R, NW, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = ntx * nty;
    int grid = n_tiles < MINB * L.num_sms ? n_tiles : MINB * L.num_sms;
    if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * NW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = L.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // (not when several emulated ranks share the device: early CTAs would sit on SMs the other ranks need)
    a.early_trigger = grid_limit > 0 ? 0 : 1;
    e = cudaLaunchKernelEx(&cfg, sor_blocked_push_kernel<R, NW, MINB>, p_map, d_map, a, ntx, nty, work_counter, push);
    ++*L.launches;
    return (int)e;
}

template <int R, int NW, int MINB>
static int launch_solve_cfg(const Launch &L, SolveArgs &sa)
{
    BlockedArgs &a = sa.a;
    const Geo &g = a.g;
    int ntx, nty;
    if (!tile_cfg<R, NW>(a, ntx, nty)) return -1;
    if (ntx <= 0 || nty <= 0 || a.hpx > a.tw_out || a.hpy > a.th_out) return -1;   // halo must stay inside the 3x3
    if ((long long)ntx * nty * sa.passes > 0x3fffffff) return -1;
    sa.ntx = ntx;
    sa.nty = nty;
    CUtensorMap d_map, p0_map, p1_map;
    if (!tma_make_map_2d(&d_map, a.div, g.nx, g.ny, g.nx, BLK_RW, R * NW) ||
        !tma_make_map_2d(&p0_map, sa.buf[0], g.nx, g.ny, g.nx, BLK_RW, R * NW) ||
        !tma_make_map_2d(&p1_map, sa.buf[1], g.nx, g.ny, g.nx, BLK_RW, R * NW))
        return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * R * NW * BLK_RW * 4 + sizeof(float4) * 2 * NW * 2 * 32;
    cudaError_t e = cudaFuncSetAttribute(sor_solve_tma_kernel<R, NW, MINB>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long total = (long long)ntx * nty * sa.passes;
    const int grid = total < MINB * L.num_sms ? (int)total : MINB * L.num_sms;
    sor_solve_tma_kernel<R, NW, MINB><<<grid, 32 * NW, smem, L.stream>>>(d_map, p0_map, p1_map, sa);
    ++*L.launches;
    return (int)cudaGetLastError();
}

// Flags needed by launch_sor_solve for a grid: passes * tiles (upper bound over the shapes).
size_t sor_solve_flag_count(const Geo &g, int iters, int t_block)
{
    const int passes = (iters + t_block - 1) / t_block;
    const long long tiles = (long long)((g.x1 - g.x0) / 64 + 3) * ((g.y1 - g.y0) / 32 + 3);   // generous
    return (size_t)(passes * tiles);
}

// The whole solve (iters full iterations, t_block per pass) in ONE launch.  p is the caller's
// buffer (receives the result), scratch the ping-pong partner.  Returns -1 when the configuration
// is not eligible (the caller then uses one launch per pass).
int launch_sor_solve(const Launch &L, float *p, float *scratch, const float *div, const Geo &g, float dx,
                     float omega, int iters, int t_block, int shape, int *work_counter, unsigned int *done,
                     size_t done_capacity, unsigned int gen)
{
    if (iters <= 0 || t_block <= 0 || g.x1 <= g.x0 || g.y1 <= g.y0) return -1;
    if (2 * t_block > SOR_BLOCKED_MAX_HALF) return -1;
    const bool vec_ok = (g.nx % 4 == 0) && ((uintptr_t)p % 16 == 0) && ((uintptr_t)scratch % 16 == 0) &&
                        ((uintptr_t)div % 16 == 0);
    if (!vec_ok || !work_counter || !done || tma_encode_fn() == nullptr) return -1;
    SolveArgs sa;
    BlockedArgs &a = sa.a;
    a.p_out = nullptr;
    a.p_in = nullptr;
    a.div = div;
    a.g = g;
    a.k = make_sor_coef(dx, omega);
    a.first_parity = 0;
    a.vec_ok = 1;
    a.early_trigger = 0;
    sa.passes = (iters + t_block - 1) / t_block;
    const int t_last = iters - (sa.passes - 1) * t_block;
    a.n_half = sa.passes > 1 ? 2 * t_block : 2 * t_last;   // tile geometry follows the regular pass
    sa.n_half_last = 2 * t_last;
    // the LAST pass must land in the caller's buffer
    sa.buf[(sa.passes - 1) & 1] = p;
    sa.buf[sa.passes & 1] = scratch;
    sa.work_counter = work_counter;
    sa.done = done;
    sa.gen = gen;
    if (sor_solve_flag_count(g, iters, t_block) > done_capacity) return -1;
    switch (shape) {
        case 2: return launch_solve_cfg<12, 8, 2>(L, sa);
        case 3: return launch_solve_cfg<12, 16, 1>(L, sa);
        case 5: return launch_solve_cfg<10, 16, 1>(L, sa);
        default: return -1;
    }
}

int launch_sor_blocked(const Launch &L, float *p_out, const float *p_in, const float *div, const Geo &g,
                       float dx, float omega, int first_parity, int n_half, int shape, int *work_counter)
{
    if (g.x1 <= g.x0 || g.y1 <= g.y0 || n_half <= 0) return 0;
    if (n_half > SOR_BLOCKED_MAX_HALF) return (int)cudaErrorInvalidValue;
    BlockedArgs a;
    a.p_out = p_out;
    a.p_in = p_in;
    a.div = div;
    a.g = g;
    a.k = make_sor_coef(dx, omega);
    a.first_parity = first_parity & 1;
    a.n_half = n_half;
    a.vec_ok = (g.nx % 4 == 0) && ((uintptr_t)p_out % 16 == 0) && ((uintptr_t)div % 16 == 0) &&
               (!p_in || (uintptr_t)p_in % 16 == 0);
    const bool tma_ok = a.vec_ok && work_counter && tma_encode_fn() != nullptr;
    switch (shape) {
        case 1: return launch_cfg<12, 16, 1>(L, a);  // 128 x 192 region, one CTA per SM
        case 2:                                      // 128 x 96 region, persistent, TMA-prefetched
            return tma_ok ? launch_cfg_tma<12, 8, 2>(L, a, work_counter) : launch_cfg<12, 8, 2>(L, a);
        case 3:                                      // 128 x 192 region, persistent, TMA-prefetched
            return tma_ok ? launch_cfg_tma<12, 16, 1>(L, a, work_counter) : launch_cfg<12, 16, 1>(L, a);
        case 4:                                      // 128 x 144 region, 12 warps (up to 168 registers)
            return tma_ok ? launch_cfg_tma<12, 12, 1>(L, a, work_counter) : launch_cfg<12, 16, 1>(L, a);
        case 5:                                      // 128 x 160 region, 16 warps x 10 rows
            return tma_ok ? launch_cfg_tma<10, 16, 1>(L, a, work_counter) : launch_cfg<12, 16, 1>(L, a);
        case 6:                                      // 128 x 192 region, 12 warps x 16 rows
            return tma_ok ? launch_cfg_tma<16, 12, 1>(L, a, work_counter) : launch_cfg<12, 16, 1>(L, a);
        // (20 x 10 and 24 x 8 strips would hide more latency, but their row mailboxes push the CTA past
        // 227 KB of shared memory next to the two staged regions)
        case 7:                                      // 128 x 176 region, 16 warps x 11 rows
            return tma_ok ? launch_cfg_tma<11, 16, 1>(L, a, work_counter) : launch_cfg<12, 16, 1>(L, a);
        default: return launch_cfg<12, 8, 2>(L, a);  // 128 x 96 region, two CTAs per SM
    }
}

// One blocked pass fused with its halo exchange (see SorPushArgs).  Needs the TMA loader: 16-byte
// aligned rows and pointers (decomposed windows always are); shapes 3, 5 and 7 (default).
int launch_sor_blocked_push(const Launch &L, float *p_out, const float *p_in, const float *div, const Geo &g,
                            float dx, float omega, int first_parity, int n_half, int shape, int *work_counter,
                            SorPushArgs &push, int grid_limit)
{
    if (g.x1 <= g.x0 || g.y1 <= g.y0 || n_half <= 0 || n_half > SOR_BLOCKED_MAX_HALF) return (int)cudaErrorInvalidValue;
    BlockedArgs a;
    a.p_out = p_out;
    a.p_in = p_in;
    a.div = div;
    a.g = g;
    a.k = make_sor_coef(dx, omega);
    a.first_parity = first_parity & 1;
    a.n_half = n_half;
    a.vec_ok = (g.nx % 4 == 0) && ((uintptr_t)p_out % 16 == 0) && ((uintptr_t)div % 16 == 0) &&
               (!p_in || (uintptr_t)p_in % 16 == 0);
    if (!a.vec_ok || !work_counter || tma_encode_fn() == nullptr) return (int)cudaErrorInvalidValue;
    for (int k = 0; k < push.n_peers; k++) {
        const SorPushPeer &q = push.peer[k];
        if ((q.sx0 | q.sx1 | q.dx | q.pitch) & 3 || (uintptr_t)q.base % 16) return (int)cudaErrorInvalidValue;
    }
    if (push.n_peers > 0 && ((g.x0 | g.x1) & 3)) return (int)cudaErrorInvalidValue;
    switch (shape) {
        case 3: return launch_cfg_push<12, 16, 1>(L, a, work_counter, push, grid_limit);
        case 5: return launch_cfg_push<10, 16, 1>(L, a, work_counter, push, grid_limit);
        default: return launch_cfg_push<11, 16, 1>(L, a, work_counter, push, grid_limit);
    }
}

int preload_sor_blocked_kernels()
{
    FS_PRELOAD((sor_blocked_push_kernel<12, 16, 1>));
    FS_PRELOAD((sor_blocked_push_kernel<10, 16, 1>));
    FS_PRELOAD((sor_blocked_push_kernel<11, 16, 1>));
    FS_PRELOAD((sor_blocked_tma_kernel<12, 16, 1>));
    FS_PRELOAD((sor_blocked_tma_kernel<11, 16, 1>));
    FS_PRELOAD((sor_blocked_kernel<12, 16, 1>));
    return 0;
}

}  // namespace fs
