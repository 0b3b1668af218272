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
// Traffic per pass: read p + d over the region, write p over the tile, i.e.
// ~(8*redundancy + 4) B/node for H/2 full iterations instead of 12 B per
// iteration per colour sector.
#include "kernels.h"
#include "sor.cuh"

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
};

template <bool WALL>
__device__ __forceinline__ float update_node(float pc, float l, float r, float d, float u, float dxd,
                                             const SorCoef &k, int gi, int gj, int GX, int GY)
{
    if constexpr (!WALL) {
        return sor_update_interior(pc, l, r, d, u, dxd, k);
    } else {
        if ((unsigned)gi >= (unsigned)GX || (unsigned)gj >= (unsigned)GY) return pc;  // not a node
        const bool hl = gi > 0, hr = gi < GX - 1, hd = gj > 0, hu = gj < GY - 1;
        if (hl && hr && hd && hu) return sor_update_interior(pc, l, r, d, u, dxd, k);
        return sor_update_wall(pc, l, r, d, u, hl, hr, hd, hu, dxd, k);
    }
}

// One colour over a warp's strip.  Q = colour offset inside the strip: row r
// updates columns {0,2} when (r+Q) is even and {1,3} when it is odd.
template <int R, int Q, bool WALL>
__device__ __forceinline__ void strip_half_sweep(float (&p)[R][4], const float (&dxd)[R][4],
                                                 const float (&dn)[4], const float (&up)[4],
                                                 const SorCoef &k, int gi0, int gj0, int GX, int GY)
{
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int gj = gj0 + r;
        if (((r + Q) & 1) == 0) {
            const float lft = __shfl_up_sync(0xffffffffu, p[r][3], 1);
            const float d0 = r > 0 ? p[r > 0 ? r - 1 : 0][0] : dn[0], u0 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][0] : up[0];
            const float d2 = r > 0 ? p[r > 0 ? r - 1 : 0][2] : dn[2], u2 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][2] : up[2];
            const float n0 = update_node<WALL>(p[r][0], lft, p[r][1], d0, u0, dxd[r][0], k, gi0 + 0, gj, GX, GY);
            const float n2 = update_node<WALL>(p[r][2], p[r][1], p[r][3], d2, u2, dxd[r][2], k, gi0 + 2, gj, GX, GY);
            p[r][0] = n0;
            p[r][2] = n2;
        } else {
            const float rgt = __shfl_down_sync(0xffffffffu, p[r][0], 1);
            const float d1 = r > 0 ? p[r > 0 ? r - 1 : 0][1] : dn[1], u1 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][1] : up[1];
            const float d3 = r > 0 ? p[r > 0 ? r - 1 : 0][3] : dn[3], u3 = r < R - 1 ? p[r < R - 1 ? r + 1 : 0][3] : up[3];
            const float n1 = update_node<WALL>(p[r][1], p[r][0], p[r][2], d1, u1, dxd[r][1], k, gi0 + 1, gj, GX, GY);
            const float n3 = update_node<WALL>(p[r][3], p[r][2], rgt, d3, u3, dxd[r][3], k, gi0 + 3, gj, GX, GY);
            p[r][1] = n1;
            p[r][3] = n3;
        }
    }
}

template <int R, int NW, bool WALL>
__device__ __forceinline__ void run_pass(const BlockedArgs &a, float4 (*mail)[NW][2][32])
{
    const int w = threadIdx.x >> 5, t = threadIdx.x & 31;
    const Geo &g = a.g;
    // region origin (local) and this thread's first column / this warp's first row
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

    // colour bookkeeping: node (column c, row r) of this strip has global parity (c + r + pb) & 1
    const int pb = (gi0 + gj0) & 1;
    float dn[4] = {0.f, 0.f, 0.f, 0.f}, up[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < a.n_half; s++) {
        const int buf = s & 1;
        mail[buf][w][0][t] = make_float4(p[0][0], p[0][1], p[0][2], p[0][3]);
        mail[buf][w][1][t] = make_float4(p[R - 1][0], p[R - 1][1], p[R - 1][2], p[R - 1][3]);
        __syncthreads();
        if (w > 0) {
            const float4 q = mail[buf][w - 1][1][t];
            dn[0] = q.x; dn[1] = q.y; dn[2] = q.z; dn[3] = q.w;
        }
        if (w < NW - 1) {
            const float4 q = mail[buf][w + 1][0][t];
            up[0] = q.x; up[1] = q.y; up[2] = q.z; up[3] = q.w;
        }
        const int q_eff = (a.first_parity + s + pb) & 1;
        if (q_eff == 0) strip_half_sweep<R, 0, WALL>(p, dxd, dn, up, a.k, gi0, gj0, g.GX, g.GY);
        else            strip_half_sweep<R, 1, WALL>(p, dxd, dn, up, a.k, gi0, gj0, g.GX, g.GY);
    }

    // write back the tile interior, clipped to the compute rectangle
    const int ox0 = max(rlx0 + a.hpx, g.x0), ox1 = min(rlx0 + a.hpx + a.tw_out, g.x1);
    const int oy0 = max(rly0 + a.hpy, g.y0), oy1 = min(rly0 + a.hpy + a.th_out, g.y1);
    const bool cols_full = lx0 >= ox0 && lx0 + 3 < ox1;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int ly = ly0 + r;
        if (ly < oy0 || ly >= oy1) continue;
        float *row = a.p_out + (size_t)ly * g.nx;
        if (cols_full && a.vec_ok) {
            *reinterpret_cast<float4 *>(row + lx0) = make_float4(p[r][0], p[r][1], p[r][2], p[r][3]);
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++)
                if (lx0 + c >= ox0 && lx0 + c < ox1) row[lx0 + c] = p[r][c];
        }
    }
}

template <int R, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB) sor_blocked_kernel(const BlockedArgs a)
{
    __shared__ float4 mail[2][NW][2][32];
    // does the region touch a domain wall or stick out of the domain?  (CTA-uniform)
    const int rgx0 = a.g.ox + a.lax + (int)blockIdx.x * a.tw_out - a.hpx;
    const int rgy0 = a.g.oy + a.lay + (int)blockIdx.y * a.th_out - a.hpy;
    const bool wall = rgx0 <= 0 || rgy0 <= 0 || rgx0 + BLK_RW >= a.g.GX || rgy0 + R * NW >= a.g.GY;
    if (!wall) run_pass<R, NW, false>(a, mail);
    else       run_pass<R, NW, true>(a, mail);
}

template <int R, int NW, int MINB>
static int launch_cfg(const Launch &L, BlockedArgs &a)
{
    const Geo &g = a.g;
    const int H = a.n_half;
    a.hpx = (H + 3) & ~3;
    a.hpy = H;
    a.tw_out = BLK_RW - 2 * a.hpx;
    a.th_out = R * NW - 2 * a.hpy;
    if (a.tw_out <= 0 || a.th_out <= 0) return (int)cudaErrorInvalidValue;
    a.lax = g.x0 & ~3;  // keep every region origin a multiple of 4 columns (float4 alignment)
    a.lay = g.y0;
    const int ntx = (g.x1 - a.lax + a.tw_out - 1) / a.tw_out;
    const int nty = (g.y1 - a.lay + a.th_out - 1) / a.th_out;
    if (ntx <= 0 || nty <= 0) return 0;
    sor_blocked_kernel<R, NW, MINB><<<dim3(ntx, nty), 32 * NW, 0, L.stream>>>(a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int launch_sor_blocked(const Launch &L, float *p_out, const float *p_in, const float *div, const Geo &g,
                       float dx, float omega, int first_parity, int n_half, int shape)
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
    switch (shape) {
        case 1: return launch_cfg<12, 16, 1>(L, a);  // 128 x 192 region, one CTA per SM
        default: return launch_cfg<12, 8, 2>(L, a);  // 128 x 96 region, two CTAs per SM
    }
}

}  // namespace fs
