// TMA-tiled advection (advect.h:74-85).
//
// One CTA produces a 64x32-node output tile.  The source field's tile plus a
// CFL-bounded halo (HALO nodes each way, +1 for the upper bilinear corner) is
// staged into shared memory by ONE bulk-tensor copy (cp.async.bulk.tensor.2d,
// SASS UTMALDG) while the threads read their velocities and form the backtrace;
// the four corner reads of every node then hit shared memory.  A backtrace that
// leaves the staged tile (a touch impulse can move a node dozens of cells,
// SURVEY.md §7 hard part 6) falls back, per fetch, to an L2 gather from global
// memory — so the result never depends on the halo width.
//
// The field is described to the TMA unit as a 2-D tensor of 32-bit words,
// NC*nx words per row (NC = 2 for velocity, 3 for UQ32 dye): rows must be
// 16-byte multiples, which holds for the big grids (the small / odd ones take
// advect_gather).  Out-of-tensor parts of a box are zero-filled by the hardware
// and are never sampled: `sample` clamps to the domain first.
//
// The dye result (12 B per node) is staged in shared memory and written with a
// bulk-tensor store (UTMASTG) so the 12-byte AoS elements leave the SM as full
// lines; velocity (8 B) is stored directly, coalesced.
#include "advect.cuh"
#include "kernels.h"
#include "stencil.cuh"
#include "tma.cuh"
#include "upscale.cuh"

namespace fs {

constexpr int AT_TX = 64, AT_TY = 32, AT_HALO = 4, AT_THREADS = 256;

template <class P>
struct TileShape {
    static constexpr int ALIGN = P::NC == 2 ? 2 : 4;   // box rows must be 16-byte multiples
    static constexpr int W = ((AT_TX + 2 * AT_HALO + 1 + ALIGN - 1) / ALIGN) * ALIGN;
    static constexpr int H = AT_TY + 2 * AT_HALO + 1;
    static constexpr int ROW_WORDS = W * P::NC;
    static constexpr int IN_BYTES = ROW_WORDS * H * 4;
    static constexpr int OUT_BYTES = P::NC == 3 ? AT_TX * AT_TY * 12 : 0;
    static_assert(ROW_WORDS <= 256, "TMA box dimension limit");
};

// Fetch a source node: shared-memory tile if staged, else global (L2) gather.
template <class P, int W = TileShape<P>::W, int H = TileShape<P>::H>
struct TileFetch {
    static constexpr int ROW_WORDS = W * P::NC;
    const typename P::raw_t *tile;   // smem, ROW_WORDS words per row
    const typename P::raw_t *base;   // global window
    int bx0, by0;                    // local coordinate of the tile's first staged node
    int ox, oy, nx;
    int vx0, vy0, vw, vh;            // valid part of the window (Geo::vx0..)
    int *status;
    __device__ __forceinline__ void operator()(int gi, int gj, typename P::raw_t (&o)[P::NC]) const
    {
        const int lx = gi - ox, ly = gj - oy;
        if ((unsigned)(lx - vx0) >= (unsigned)vw || (unsigned)(ly - vy0) >= (unsigned)vh) {  // not valid here
            if (status) atomicCAS(status, 0, FS_ERR_HALO_OVERRUN);   // the first error wins
#pragma unroll
            for (int ch = 0; ch < P::NC; ch++) o[ch] = 0;
            return;
        }
        const int tx = lx - bx0, ty = ly - by0;
        if ((unsigned)tx < (unsigned)W && (unsigned)ty < (unsigned)H) {
            const typename P::raw_t *q = tile + ty * ROW_WORDS + tx * P::NC;
            if constexpr (P::NC == 2) {
                const float2 t = *reinterpret_cast<const float2 *>(q);
                o[0] = t.x;
                o[1] = t.y;
            } else {
#pragma unroll
                for (int ch = 0; ch < P::NC; ch++) o[ch] = q[ch];
            }
        } else {
            const typename P::raw_t *q = base + ((size_t)ly * nx + lx) * P::NC;
#pragma unroll
            for (int ch = 0; ch < P::NC; ch++) o[ch] = __ldg(q + ch);
        }
    }
};

// The general `sample` (walls, corners, backtraces that leave the staged tile) is kept out of line:
// the unrolled per-row loop then only carries the short interior path, and the kernel stays inside
// the instruction cache (ncu: `no_instruction` was the top stall with it inlined 8 times).  The result
// comes back by value, in registers: an out-parameter put every node's result through local memory.
template <class P>
struct Raw {
    typename P::raw_t v[P::NC];
};
template <class P, class F>
__device__ __noinline__ Raw<P> sample_slow(const F &fetch, float si, float sj, int GX, int GY, bool no_slip)
{
    Raw<P> r;
    sample<P>(r.v, fetch, si, sj, GX, GY, no_slip);
    return r;
}

// Where the straight-line interior path applies (CTA-uniform): the bilinear cell (tx..tx+1, ty..ty+1)
// must be staged AND inside the valid part of this rank's window (outside the window the hardware
// zero-filled the tile).  Two unsigned compares on the floored source replace advect.h:26-29's four
// float compares: the valid rectangle lies inside the domain, so floor(si) in [vx0, vx1-1) already
// gives 0 <= si < GX-1 (advect.h:38, "!x_oob && !y_oob"); a source beyond +-2^31 saturates the
// conversion and fails the range test.
struct FastWindow {
    int cx, cy;              // global node whose cell is the first one on the fast path
    unsigned x_span, y_span; // cells on the fast path each way
    int toff;                // word offset of that node in the staged tile
};
template <int NC>
__device__ __forceinline__ FastWindow fast_window(const Geo &g, int bx0, int by0, int W, int H)
{
    const int tx_lo = max(0, g.vx0 - bx0), tx_hi = min(W - 1, g.vx1 - 1 - bx0);   // tx in [tx_lo, tx_hi)
    const int ty_lo = max(0, g.vy0 - by0), ty_hi = min(H - 1, g.vy1 - 1 - by0);
    FastWindow f;
    f.cx = g.ox + bx0 + tx_lo;
    f.cy = g.oy + by0 + ty_lo;
    f.x_span = (unsigned)max(0, tx_hi - tx_lo);
    f.y_span = (unsigned)max(0, ty_hi - ty_lo);
    f.toff = (ty_lo * W + tx_lo) * NC;
    return f;
}
// true: (tx, ty) = the cell relative to FastWindow's first node, (di, dj) = advect.h:34-35's fractions
// (float(int(floor(s))) == floorf(s) exactly for every s that passes the range test)
__device__ __forceinline__ bool fast_cell(const FastWindow &f, float si, float sj, int &tx, int &ty, float &di,
                                          float &dj)
{
    const int gi = __float2int_rd(si), gj = __float2int_rd(sj);
    tx = gi - f.cx;
    ty = gj - f.cy;
    if ((unsigned)tx >= f.x_span || (unsigned)ty >= f.y_span) return false;
    di = __fsub_rn(si, (float)gi);
    dj = __fsub_rn(sj, (float)gj);
    return true;
}

struct TmaAdvectArgs {
    void *next_p;
    const void *p;
    const float2 *vel;
    Geo g;
    float dt;
    int no_slip;
    int vel_is_p;        // velocity advect: the velocity of a node is already in the staged tile
    int store_tma;       // dye: write the tile with a bulk-tensor store
    int *status;
    // fused gradient-subtract (fs_step, ino:276+282): when grad_p is set, `vel` holds the UNPROJECTED
    // velocity; each thread forms v - grad p for its node, stores it to v_out and advects with it
    // (the dye backtrace only needs the projected velocity at the node itself).
    const float *grad_p;
    float2 *v_out;
    float two_dx_inv;
};

// (Round 2 built the persistent variant the round-1 review asked for — one CTA per SM slot walking the
// tiles with a two-stage TMA ring and the next tile's velocities prefetched into registers — and
// measured it SLOWER: velocity 0.083 vs 0.075 ms, dye 0.151 vs 0.126 ms at 4096^2.  The second stage
// and the prefetch registers cut the resident CTAs from 3-4 to 2-3 per SM, and this kernel hides its
// shared-memory and ALU latencies with resident warps, not with a deeper pipeline; one tile per CTA
// stays.)
// (launch bounds: the dye's 62 KB of tiles allow 3 CTAs per SM, so its build may use up to 85 registers — 60 instead of
// 48 measured 1 % faster; squeezing the velocity build into 40 registers for 6 CTAs per SM instead of 5 measured 3 %
// SLOWER, it keeps the default)
template <class P, bool STORE_TMA>
__global__ void __launch_bounds__(AT_THREADS, P::NC == 2 ? 5 : 3)
advect_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ CUtensorMap out_map,
                  const TmaAdvectArgs a)
{
    using raw_t = typename P::raw_t;
    using TS = TileShape<P>;
    static_assert(!STORE_TMA || P::NC == 3, "only the dye leaves through a bulk-tensor store");
    extern __shared__ __align__(128) unsigned char smem[];
    raw_t *tile = reinterpret_cast<raw_t *>(smem);
    raw_t *otile = reinterpret_cast<raw_t *>(smem + ((TS::IN_BYTES + 127) & ~127));
    __shared__ __align__(8) uint64_t bar;

    const Geo &g = a.g;
    const int tx0 = g.x0 + blockIdx.x * AT_TX, ty0 = g.y0 + blockIdx.y * AT_TY;   // tile origin (local)
    const int bx0 = tx0 - AT_HALO, by0 = ty0 - AT_HALO;                           // staged box origin

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, TS::IN_BYTES);
        tma_load_2d(tile, &in_map, bx0 * P::NC, by0, &bar);
    }

    // while the tile is in flight: velocities (dye advect) for this thread's nodes
    constexpr int ROWS_PER_IT = AT_THREADS / AT_TX, ITERS = AT_TY / ROWS_PER_IT;
    const int cx = threadIdx.x % AT_TX, cy = threadIdx.x / AT_TX;
    // per-thread invariants of the row loop (ncu: per-iteration rectangle tests, int -> float conversions of the
    // node coordinates and 64-bit index arithmetic were ~10 of ~76 instructions per node)
    const int lx = tx0 + cx;
    const bool live_x = lx < g.x1;
    const size_t l0 = (size_t)(ty0 + cy) * g.nx + lx, lstep = (size_t)ROWS_PER_IT * g.nx;
    const float fgi = (float)(g.ox + lx);                   // node coordinates as floats: exact
    float2 vel[ITERS];
    if (!a.vel_is_p) {
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            const int ly = ty0 + cy + it * ROWS_PER_IT;
            const bool live = live_x && ly < g.y1;
            vel[it] = live ? __ldg(a.vel + l0 + it * lstep) : make_float2(0.f, 0.f);
            if (a.grad_p && live) {
                vel[it] = grad_sub_value(vel[it], a.grad_p, g, lx, ly, a.two_dx_inv);
                a.v_out[l0 + it * lstep] = vel[it];
            }
        }
    }
    mbar_wait(&bar, 0);

    TileFetch<P> fetch{tile, reinterpret_cast<const raw_t *>(a.p), bx0, by0, g.ox, g.oy, g.nx,
                       g.vx0, g.vy0, g.vx1 - g.vx0, g.vy1 - g.vy0, a.status};
    const FastWindow fw = fast_window<P::NC>(g, bx0, by0, TS::W, TS::H);
    const raw_t *ftile = tile + fw.toff;
    const float fgj0 = (float)(g.oy + ty0 + cy);
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const int ry = cy + it * ROWS_PER_IT;
        const int ly = ty0 + ry;
        const bool live = live_x && ly < g.y1;
        Raw<P> out;
#pragma unroll
        for (int ch = 0; ch < P::NC; ch++) out.v[ch] = 0;
        if (live) {
            float2 vv;
            if (a.vel_is_p) {
                if constexpr (P::NC == 2)
                    vv = *reinterpret_cast<const float2 *>(tile + (ry + AT_HALO) * TS::ROW_WORDS + (cx + AT_HALO) * 2);
                else
                    vv = make_float2(0.f, 0.f);
            } else {
                vv = vel[it];
            }
            float di, dj;
            int tx, ty;
            // advect.h:81 (backtrace()) with the node coordinates already in float
            const float si = __fsub_rn(fgi, __fmul_rn(vv.x, a.dt));
            const float sj = __fsub_rn(fgj0 + (float)(it * ROWS_PER_IT), __fmul_rn(vv.y, a.dt));
            // interior (advect.h:38: !x_oob && !y_oob) with all four corners staged: straight-line path
            if (fast_cell(fw, si, sj, tx, ty, di, dj)) {
                const float wi = __fsub_rn(1.0f, di), wj = __fsub_rn(1.0f, dj);
                const raw_t *q = ftile + ty * TS::ROW_WORDS + tx * P::NC;
                if constexpr (P::NC == 2) {
                    const float2 p11 = *reinterpret_cast<const float2 *>(q);
                    const float2 p21 = *reinterpret_cast<const float2 *>(q + 2);
                    const float2 p12 = *reinterpret_cast<const float2 *>(q + TS::ROW_WORDS);
                    const float2 p22 = *reinterpret_cast<const float2 *>(q + TS::ROW_WORDS + 2);
                    out.v[0] = mixf(wi, di, mixf(wj, dj, p11.x, p12.x), mixf(wj, dj, p21.x, p22.x));
                    out.v[1] = mixf(wi, di, mixf(wj, dj, p11.y, p12.y), mixf(wj, dj, p21.y, p22.y));
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const float a11 = uq32_to_float(q[ch]), a21 = uq32_to_float(q[3 + ch]);
                        const float a12 = uq32_to_float(q[TS::ROW_WORDS + ch]);
                        const float a22 = uq32_to_float(q[TS::ROW_WORDS + 3 + ch]);
                        out.v[ch] = uq32_from_float(mixf(wi, di, mixf(wj, dj, a11, a12), mixf(wj, dj, a21, a22)));
                    }
                }
            } else {
                out = sample_slow<P>(fetch, si, sj, g.GX, g.GY, a.no_slip != 0);
            }
        }
        if constexpr (P::NC == 2) {
            if (live) reinterpret_cast<float2 *>(a.next_p)[l0 + it * lstep] = make_float2(out.v[0], out.v[1]);
        } else if constexpr (STORE_TMA) {
            raw_t *q = otile + (ry * AT_TX + cx) * 3;
            q[0] = out.v[0]; q[1] = out.v[1]; q[2] = out.v[2];
        } else {
            if (live) {
                raw_t *q = reinterpret_cast<raw_t *>(a.next_p) + (l0 + it * lstep) * 3;
                q[0] = out.v[0]; q[1] = out.v[1]; q[2] = out.v[2];
            }
        }
    }
    if constexpr (STORE_TMA) {
        // out_map describes the compute rectangle only, so the hardware clips partial tiles
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store_2d(&out_map, blockIdx.x * AT_TX * 3, blockIdx.y * AT_TY, otile);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

// ---- fused: advect velocity (no-slip) + drag overwrite + divergence (ino:253, 264-269, 274) -------------
// The CTA advects its 64x32 tile PLUS a one-node ring into shared memory, applies the drag records
// that land there (in queue order, later wins), then forms the divergence of the tile from shared
// memory and writes both the forced velocity and the divergence.  Saves the divergence kernel's
// read of the velocity (8 of its 12 B/node) and two launches; the ring costs 9 % more backtraces.
constexpr int AD_LEFT = AT_HALO + 2;   // ring + halo, rounded up so the box starts on a 16-byte boundary (TMA)
constexpr int AD_W = ((AD_LEFT + AT_TX + 1 + AT_HALO + 1 + 1) / 2) * 2;   // staged source tile (nodes)
constexpr int AD_H = AT_TY + 2 + 2 * AT_HALO + 1;
constexpr int AD_AW = AT_TX + 2, AD_AH = AT_TY + 2;               // advected tile + ring
constexpr int AD_MAX_DRAGS = 128;

struct AdvDivArgs {
    float2 *v_out;
    const float2 *v_in;
    float *div;
    Geo g;                    // compute rectangle = where the DIVERGENCE is written
    int sx0, sy0, sx1, sy1;   // where the forced VELOCITY is stored (decomposed grids: the owned rectangle only —
                              // the ring around it is recomputed by every rank instead of being exchanged)
    int *status;
    float dt, two_dx_inv;
    int n_drags;
    fs_drag drags[AD_MAX_DRAGS];
};

__global__ void __launch_bounds__(AT_THREADS)
advect_div_tma_kernel(const __grid_constant__ CUtensorMap in_map, const __grid_constant__ AdvDivArgs a)
{
    using P = Vec2Payload;
    constexpr int ROW_WORDS = AD_W * 2;
    extern __shared__ __align__(128) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    float2 *adv = reinterpret_cast<float2 *>(smem + ((ROW_WORDS * AD_H * 4 + 127) & ~127));
    __shared__ __align__(8) uint64_t bar;

    const Geo &g = a.g;
    const int tx0 = g.x0 + blockIdx.x * AT_TX, ty0 = g.y0 + blockIdx.y * AT_TY;
    const int bx0 = tx0 - AD_LEFT, by0 = ty0 - 1 - AT_HALO;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, ROW_WORDS * AD_H * 4);
        tma_load_2d(tile, &in_map, bx0 * 2, by0, &bar);
    }
    TileFetch<P, AD_W, AD_H> fetch{tile, reinterpret_cast<const float *>(a.v_in), bx0, by0, g.ox, g.oy, g.nx,
                                   g.vx0, g.vy0, g.vx1 - g.vx0, g.vy1 - g.vy0, a.status};
    const FastWindow fw = fast_window<2>(g, bx0, by0, AD_W, AD_H);
    const float *ftile = tile + fw.toff;
    // only the compute rectangle and its one-node ring are advected (partial tiles stick out of it; on a
    // decomposed grid nodes further out would backtrace into ghosts that were never refreshed):
    // in (ax, ay) = position in the advected tile + ring
    const int ax_lo = max(g.x0 - 1, 0) - (tx0 - 1), ay_lo = max(g.y0 - 1, 0) - (ty0 - 1);
    const unsigned ax_span = (unsigned)max(0, min(g.x1 + 1, g.nx) - (tx0 - 1) - ax_lo);
    const unsigned ay_span = (unsigned)max(0, min(g.y1 + 1, g.ny) - (ty0 - 1) - ay_lo);
    mbar_wait(&bar, 0);

    // ---- advect tile + ring into shared memory ----
    int ay = threadIdx.x / AD_AW, ax = threadIdx.x - ay * AD_AW;
#pragma unroll 1
    for (int k = threadIdx.x; k < AD_AW * AD_AH; k += AT_THREADS) {
        float2 out = make_float2(0.f, 0.f);
        if ((unsigned)(ax - ax_lo) < ax_span && (unsigned)(ay - ay_lo) < ay_span) {
            const float2 vv = *reinterpret_cast<const float2 *>(tile + (ay + AT_HALO) * ROW_WORDS + (ax + AD_LEFT - 1) * 2);
            float si, sj, di, dj;
            int tx, ty;
            backtrace(si, sj, g.ox + tx0 - 1 + ax, g.oy + ty0 - 1 + ay, vv, a.dt);
            if (fast_cell(fw, si, sj, tx, ty, di, dj)) {
                const float wi = __fsub_rn(1.0f, di), wj = __fsub_rn(1.0f, dj);
                const float *q = ftile + ty * ROW_WORDS + tx * 2;
                const float2 p11 = *reinterpret_cast<const float2 *>(q);
                const float2 p21 = *reinterpret_cast<const float2 *>(q + 2);
                const float2 p12 = *reinterpret_cast<const float2 *>(q + ROW_WORDS);
                const float2 p22 = *reinterpret_cast<const float2 *>(q + ROW_WORDS + 2);
                out.x = mixf(wi, di, mixf(wj, dj, p11.x, p12.x), mixf(wj, dj, p21.x, p22.x));
                out.y = mixf(wi, di, mixf(wj, dj, p11.y, p12.y), mixf(wj, dj, p21.y, p22.y));
            } else {
                const Raw<P> r = sample_slow<P>(fetch, si, sj, g.GX, g.GY, true);
                out = make_float2(r.v[0], r.v[1]);
            }
        }
        adv[k] = out;
        ax += AT_THREADS % AD_AW;                       // k += AT_THREADS without the division
        ay += AT_THREADS / AD_AW;
        if (ax >= AD_AW) {
            ax -= AD_AW;
            ay++;
        }
    }
    __syncthreads();

    // ---- drag overwrite (ino:264-269) on the tile + ring: parallel hit test, in-order replay ----
    if (a.n_drags > 0) {
        int hit = 0;
        for (int k = threadIdx.x; k < a.n_drags; k += AT_THREADS) {
            const int lx = (int)a.drags[k].cy - g.ox, ly = (int)a.drags[k].cx - g.oy;
            hit |= (lx >= tx0 - 1 && lx <= tx0 + AT_TX && ly >= ty0 - 1 && ly <= ty0 + AT_TY);
        }
        if (__syncthreads_or(hit)) {
            if (threadIdx.x == 0) {
                for (int k = 0; k < a.n_drags; k++) {
                    const fs_drag m = a.drags[k];
                    if (m.cy >= g.GX || m.cx >= g.GY) continue;
                    const int ax = (int)m.cy - g.ox - (tx0 - 1), ay = (int)m.cx - g.oy - (ty0 - 1);
                    if (ax >= 0 && ax < AD_AW && ay >= 0 && ay < AD_AH) adv[ay * AD_AW + ax] = make_float2(m.vy, m.vx);
                }
            }
            __syncthreads();
        }
    }

    // ---- forced velocity + its divergence for the tile ----
    constexpr int ROWS_PER_IT = AT_THREADS / AT_TX, ITERS = AT_TY / ROWS_PER_IT;
    const int cx = threadIdx.x % AT_TX, cy = threadIdx.x / AT_TX;
    // CTA-uniform: a whole tile inside the compute and store rectangles that touches no global wall takes the
    // interior expression (div_expr_fast, finitediff.cpp:9-17) with no per-node tests and running pointers
    // (ncu: the general loop below spent ~56 instructions per node on rectangle / wall tests and 64-bit
    // address arithmetic, a third of the kernel)
    const bool whole = tx0 + AT_TX <= g.x1 && ty0 + AT_TY <= g.y1 && tx0 >= a.sx0 && tx0 + AT_TX <= a.sx1 &&
                       ty0 >= a.sy0 && ty0 + AT_TY <= a.sy1 && g.ox + tx0 > 0 && g.ox + tx0 + AT_TX < g.GX &&
                       g.oy + ty0 > 0 && g.oy + ty0 + AT_TY < g.GY;
    if (whole) {
        const float2 *ap = adv + (cy + 1) * AD_AW + cx + 1;
        const size_t l0 = (size_t)(ty0 + cy) * g.nx + tx0 + cx, lstep = (size_t)ROWS_PER_IT * g.nx;
        float2 *vo = a.v_out + l0;
        float *dv = a.div + l0;
#pragma unroll
        for (int it = 0; it < ITERS; it++) {
            const float2 c = ap[0];
            const float s = __fadd_rn(__fadd_rn(-ap[-1].x, ap[1].x), __fadd_rn(-ap[-AD_AW].y, ap[AD_AW].y));
            *vo = c;
            *dv = __fmul_rn(s, a.two_dx_inv);
            ap += ROWS_PER_IT * AD_AW;
            vo += lstep;
            dv += lstep;
        }
        return;
    }
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
        const int ry = cy + it * ROWS_PER_IT;
        const int lx = tx0 + cx, ly = ty0 + ry;
        if (lx >= g.x1 || ly >= g.y1) continue;
        const float2 c = adv[(ry + 1) * AD_AW + cx + 1];
        const float d = div_value(c, adv[(ry + 1) * AD_AW + cx].x, adv[(ry + 1) * AD_AW + cx + 2].x,
                                  adv[ry * AD_AW + cx + 1].y, adv[(ry + 2) * AD_AW + cx + 1].y, g.ox + lx, g.oy + ly,
                                  g.GX, g.GY, a.two_dx_inv);
        const size_t l = (size_t)ly * g.nx + lx;
        if (lx >= a.sx0 && lx < a.sx1 && ly >= a.sy0 && ly < a.sy1) a.v_out[l] = c;
        a.div[l] = d;
    }
}

// ---- fused: advect dye (ino:282) + the 4x RGB565 frame of the NEW dye (draw_routine, ino:116-177) --------
// The CTA advects its 64x32 tile PLUS the next row and column (the far corners of its last cells) into
// shared memory, stores the tile's dye, and renders its 64x32 cells straight from shared memory: the
// frame never re-reads the dye (12 of the stand-alone upscale's 44 B/node) and costs no extra launch.
// On a decomposed grid the extra row/column lie in the neighbour's rectangle: they are recomputed here
// (same inputs, same bits), which needs the dye ghosts one node wider than the advect halo.
constexpr int FR_NX = AT_TX + 1, FR_NY = AT_TY + 1;                         // advected nodes per CTA
constexpr int FR_W = ((FR_NX + 2 * AT_HALO + 1 + 3) / 4) * 4;              // staged source tile (nodes)
constexpr int FR_H = FR_NY + 2 * AT_HALO + 1;
constexpr int FR_ROW_WORDS = FR_W * 3;
constexpr int FR_IN_BYTES = FR_ROW_WORDS * FR_H * 4;
constexpr int FR_EXT_PITCH = FR_NX * 3;                                    // 195 words: odd multiple of 3, conflict-free columns
constexpr int FR_ITERS = (FR_NX * FR_NY + AT_THREADS - 1) / AT_THREADS;
static_assert(FR_ROW_WORDS <= 256, "TMA box dimension limit");

struct FrameAdvectArgs {
    uint32_t *next_c;
    const uint32_t *c;
    const float2 *vel;
    uint16_t *frame;        // first pixel of the cell at the compute rectangle's first node
    size_t frame_pitch;     // pixels per image row
    Geo g;
    float dt;
    int no_slip;
    int frame_aligned8;
    int *status;
};

__global__ void __launch_bounds__(AT_THREADS)
advect_rgb_frame_kernel(const __grid_constant__ CUtensorMap in_map, const FrameAdvectArgs a)
{
    using P = RgbPayload;
    extern __shared__ __align__(128) unsigned char smem[];
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem);
    uint32_t *ext = reinterpret_cast<uint32_t *>(smem + ((FR_IN_BYTES + 127) & ~127));   // [FR_NY][FR_EXT_PITCH]
    __shared__ __align__(8) uint64_t bar;

    const Geo &g = a.g;
    const int tx0 = g.x0 + blockIdx.x * AT_TX, ty0 = g.y0 + blockIdx.y * AT_TY;
    const int bx0 = tx0 - AT_HALO, by0 = ty0 - AT_HALO;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, FR_IN_BYTES);
        tma_load_2d(tile, &in_map, bx0 * 3, by0, &bar);
    }
    // nodes this CTA advects: its tile + the next row/column, inside the compute rectangle grown by one
    // node (and inside the window)
    const int lim_x = min(g.x1 + 1, g.nx), lim_y = min(g.y1 + 1, g.ny);
    float2 vel[FR_ITERS];
#pragma unroll
    for (int it = 0; it < FR_ITERS; it++) {
        const int k = threadIdx.x + it * AT_THREADS;
        const int ny = k / FR_NX, nx = k - ny * FR_NX;
        const int lx = tx0 + nx, ly = ty0 + ny;
        vel[it] = (k < FR_NX * FR_NY && lx < lim_x && ly < lim_y) ? __ldg(a.vel + (size_t)ly * g.nx + lx) : make_float2(0.f, 0.f);
    }
    mbar_wait(&bar, 0);

    TileFetch<P, FR_W, FR_H> fetch{tile, a.c, bx0, by0, g.ox, g.oy, g.nx, g.vx0, g.vy0, g.vx1 - g.vx0, g.vy1 - g.vy0, a.status};
    const FastWindow fw = fast_window<3>(g, bx0, by0, FR_W, FR_H);
    const uint32_t *ftile = tile + fw.toff;
#pragma unroll
    for (int it = 0; it < FR_ITERS; it++) {
        const int k = threadIdx.x + it * AT_THREADS;
        if (k >= FR_NX * FR_NY) break;
        const int ny = k / FR_NX, nx = k - ny * FR_NX;
        const int lx = tx0 + nx, ly = ty0 + ny;
        Raw<P> out;
        out.v[0] = out.v[1] = out.v[2] = 0u;
        if (lx < lim_x && ly < lim_y) {
            float si, sj, di, dj;
            int tx, ty;
            backtrace(si, sj, g.ox + lx, g.oy + ly, vel[it], a.dt);
            if (fast_cell(fw, si, sj, tx, ty, di, dj)) {
                const float wi = __fsub_rn(1.0f, di), wj = __fsub_rn(1.0f, dj);
                const uint32_t *q = ftile + ty * FR_ROW_WORDS + tx * 3;
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float a11 = uq32_to_float(q[ch]), a21 = uq32_to_float(q[3 + ch]);
                    const float a12 = uq32_to_float(q[FR_ROW_WORDS + ch]);
                    const float a22 = uq32_to_float(q[FR_ROW_WORDS + 3 + ch]);
                    out.v[ch] = uq32_from_float(mixf(wi, di, mixf(wj, dj, a11, a12), mixf(wj, dj, a21, a22)));
                }
            } else {
                out = sample_slow<P>(fetch, si, sj, g.GX, g.GY, a.no_slip != 0);
            }
            if (nx < AT_TX && ny < AT_TY && lx < g.x1 && ly < g.y1) {      // the tile itself: the advected dye
                uint32_t *q = a.next_c + ((size_t)ly * g.nx + lx) * 3;
                q[0] = out.v[0]; q[1] = out.v[1]; q[2] = out.v[2];
            }
        }
        uint32_t *e = ext + ny * FR_EXT_PITCH + nx * 3;
        e[0] = out.v[0]; e[1] = out.v[1]; e[2] = out.v[2];
    }
    __syncthreads();

    // ---- frame: lane = cell column (j), warp = cell rows (i); a warp writes 256 contiguous bytes per image row ----
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ly = ty0 + lane, gj = g.oy + ly;
    if (ly >= g.y1 || gj >= g.GY - 1) return;                              // no cell starts at the last node column
    for (int ci = w; ci < AT_TX; ci += AT_THREADS / 32) {
        const int lx = tx0 + ci, gi = g.ox + lx;
        if (lx >= g.x1 || gi >= g.GX - 1) break;
        uint32_t c11[3], c12[3], c21[3], c22[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            c11[ch] = ext[lane * FR_EXT_PITCH + ci * 3 + ch];              // (i,   j)
            c12[ch] = ext[(lane + 1) * FR_EXT_PITCH + ci * 3 + ch];        // (i,   j+1)
            c21[ch] = ext[lane * FR_EXT_PITCH + (ci + 1) * 3 + ch];        // (i+1, j)
            c22[ch] = ext[(lane + 1) * FR_EXT_PITCH + (ci + 1) * 3 + ch];  // (i+1, j+1)
        }
        upscale_cell_rgb565(a.frame + 4 * (size_t)(lx - g.x0) * a.frame_pitch + 4 * (size_t)(ly - g.y0), a.frame_pitch,
                            a.frame_aligned8 != 0, c11, c12, c21, c22);
    }
}

// ---- host side ---------------------------------------------------------------------------

template <class P>
bool advect_tma_legal(const void *p, const Geo &g)
{
    // TMA: 16-byte aligned base, row pitch AND box start (tile origin - halo = x0 - 4 + 64k nodes);
    // boxes of at most 256 words
    return ((uintptr_t)p % 16 == 0) && (((size_t)g.nx * P::NC * 4) % 16 == 0) && (g.x0 % 4 == 0) &&
           g.nx >= AT_TX && g.ny >= AT_TY &&
           tma_encode_fn() != nullptr;
}

template <class P>
static int launch_tma(const Launch &L, void *next_p, const void *p, const float2 *vel, const Geo &g, float dt,
                      bool no_slip, int *status, const float *grad_p = nullptr, float2 *v_out = nullptr,
                      float two_dx_inv = 0.0f)
{
    using TS = TileShape<P>;
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    if (w <= 0 || h <= 0) return 0;
    CUtensorMap in_map, out_map;
    if (!tma_make_map_2d(&in_map, p, (uint64_t)g.nx * P::NC, g.ny, (uint64_t)g.nx * P::NC, TS::ROW_WORDS, TS::H))
        return (int)cudaErrorInvalidValue;
    TmaAdvectArgs a;
    a.next_p = next_p; a.p = p; a.vel = vel; a.g = g; a.dt = dt; a.no_slip = no_slip ? 1 : 0;
    a.vel_is_p = (P::NC == 2 && (const void *)vel == p) ? 1 : 0;
    a.status = status;
    a.grad_p = grad_p;
    a.v_out = v_out;
    a.two_dx_inv = two_dx_inv;
    a.store_tma = 0;
    out_map = in_map;
    if (P::NC == 3) {
        // the store map covers exactly the compute rectangle: base = its first node
        const char *obase = (const char *)next_p + ((size_t)g.y0 * g.nx + g.x0) * 12;
        if ((uintptr_t)obase % 16 == 0 &&
            tma_make_map_2d(&out_map, obase, (uint64_t)w * 3, h, (uint64_t)g.nx * 3, AT_TX * 3, AT_TY))
            a.store_tma = 1;
    }
    const size_t smem = ((TS::IN_BYTES + 127) & ~127) + TS::OUT_BYTES;
    auto kern = advect_tma_kernel<P, false>;
    if constexpr (P::NC == 3)
        if (a.store_tma) kern = advect_tma_kernel<P, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((w + AT_TX - 1) / AT_TX, (h + AT_TY - 1) / AT_TY);
    kern<<<grid, AT_THREADS, smem, L.stream>>>(in_map, out_map, a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

bool advect_vec2f_tma_legal(const float2 *p, const Geo &g) { return advect_tma_legal<Vec2Payload>(p, g); }
bool advect_rgb_tma_legal(const uint32_t *c, const Geo &g) { return advect_tma_legal<RgbPayload>(c, g); }

int launch_advect_vec2f_tma(const Launch &L, float2 *next_p, const float2 *p, const float2 *vel, const Geo &g,
                            float dt, bool no_slip, int *status)
{
    return launch_tma<Vec2Payload>(L, next_p, p, vel, g, dt, no_slip, status);
}

int launch_advect_rgb_tma(const Launch &L, uint32_t *next_c, const uint32_t *c, const float2 *vel, const Geo &g,
                          float dt, bool no_slip, int *status)
{
    return launch_tma<RgbPayload>(L, next_c, c, vel, g, dt, no_slip, status);
}

// everything a launch of advect_div_tma_kernel passes: also what a CUDA-graph kernel node of it is re-armed with
struct AdvDivLaunchParams {
    CUtensorMap map;
    AdvDivArgs args;
    void *ptrs[2];
};

static int fill_advect_div(AdvDivLaunchParams &q, float2 *v_out, const float2 *v_in, float *div, const fs_drag *drags_host,
                           int n_drags, const Geo &g, float dt, float dx, const int *store_rect, int *status)
{
    if (n_drags > AD_MAX_DRAGS) return (int)cudaErrorInvalidValue;
    if (!tma_make_map_2d(&q.map, v_in, (uint64_t)g.nx * 2, g.ny, (uint64_t)g.nx * 2, AD_W * 2, AD_H))
        return (int)cudaErrorInvalidValue;
    AdvDivArgs &a = q.args;
    a.v_out = v_out; a.v_in = v_in; a.div = div; a.g = g; a.dt = dt;
    a.sx0 = store_rect ? store_rect[0] : g.x0; a.sy0 = store_rect ? store_rect[1] : g.y0;
    a.sx1 = store_rect ? store_rect[2] : g.x1; a.sy1 = store_rect ? store_rect[3] : g.y1;
    a.status = status;
    a.two_dx_inv = 1.0f / (2.0f * dx);
    a.n_drags = n_drags;
    for (int k = 0; k < n_drags; k++) a.drags[k] = drags_host[k];
    q.ptrs[0] = &q.map;
    q.ptrs[1] = &q.args;
    return 0;
}

int launch_advect_div_tma(const Launch &L, float2 *v_out, const float2 *v_in, float *div, const fs_drag *drags_host,
                          int n_drags, const Geo &g, float dt, float dx, const int *store_rect, int *status)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    if (w <= 0 || h <= 0) return 0;
    AdvDivLaunchParams q;
    int e0 = fill_advect_div(q, v_out, v_in, div, drags_host, n_drags, g, dt, dx, store_rect, status);
    if (e0) return e0;
    const size_t smem = ((AD_W * 2 * AD_H * 4 + 127) & ~127) + AD_AW * AD_AH * 8;
    cudaError_t e = cudaFuncSetAttribute(advect_div_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((w + AT_TX - 1) / AT_TX, (h + AT_TY - 1) / AT_TY);
    advect_div_tma_kernel<<<grid, AT_THREADS, smem, L.stream>>>(q.map, q.args);
    ++*L.launches;
    return (int)cudaGetLastError();
}

// ---- CUDA-graph support: the step's only per-step kernel arguments are the drag records of this kernel ----
const void *advect_div_kernel_func() { return (const void *)advect_div_tma_kernel; }
size_t advect_div_params_bytes() { return sizeof(AdvDivLaunchParams); }
// fills `storage` (advect_div_params_bytes() bytes, must stay alive until the graph launch has been enqueued)
// and returns the kernelParams array for cudaGraphExecKernelNodeSetParams
int advect_div_graph_params(void *storage, void ***kernel_params, float2 *v_out, const float2 *v_in, float *div,
                            const fs_drag *drags_host, int n_drags, const Geo &g, float dt, float dx)
{
    AdvDivLaunchParams *q = reinterpret_cast<AdvDivLaunchParams *>(storage);
    int e = fill_advect_div(*q, v_out, v_in, div, drags_host, n_drags, g, dt, dx, nullptr, nullptr);
    if (e) return e;
    *kernel_params = q->ptrs;
    return 0;
}

int advect_div_max_drags() { return AD_MAX_DRAGS; }

// frame_cells_y = number of cell columns (j) of the frame = its row pitch / 4
int launch_advect_rgb_frame(const Launch &L, uint32_t *next_c, uint16_t *frame, int frame_cells_y, const uint32_t *c,
                            const float2 *vel, const Geo &g, float dt, bool no_slip, int *status)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    if (w <= 0 || h <= 0) return 0;
    CUtensorMap in_map;
    if (!tma_make_map_2d(&in_map, c, (uint64_t)g.nx * 3, g.ny, (uint64_t)g.nx * 3, FR_ROW_WORDS, FR_H))
        return (int)cudaErrorInvalidValue;
    FrameAdvectArgs a;
    a.next_c = next_c; a.c = c; a.vel = vel; a.frame = frame; a.frame_pitch = 4 * (size_t)frame_cells_y;
    a.g = g; a.dt = dt; a.no_slip = no_slip ? 1 : 0; a.status = status;
    a.frame_aligned8 = (uintptr_t)frame % 8 == 0;
    const size_t smem = ((FR_IN_BYTES + 127) & ~127) + (size_t)FR_NY * FR_EXT_PITCH * 4;
    cudaError_t e = cudaFuncSetAttribute(advect_rgb_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((w + AT_TX - 1) / AT_TX, (h + AT_TY - 1) / AT_TY);
    advect_rgb_frame_kernel<<<grid, AT_THREADS, smem, L.stream>>>(in_map, a);
    ++*L.launches;
    return (int)cudaGetLastError();
}

// dye advect with the gradient-subtract of the projection folded in: v_out = v_tmp - grad p, and the
// dye is advected with v_out (ino:276 + ino:282 in one pass over the grid)
int launch_advect_rgb_tma_grad(const Launch &L, uint32_t *next_c, const uint32_t *c, float2 *v_out,
                               const float2 *v_tmp, const float *p, const Geo &g, float dt, float dx,
                               bool no_slip, int *status)
{
    return launch_tma<RgbPayload>(L, next_c, c, v_tmp, g, dt, no_slip, status, p, v_out, 1.0f / (2.0f * dx));
}

int preload_advect_tma_kernels()
{
    FS_PRELOAD((advect_tma_kernel<Vec2Payload, false>));
    FS_PRELOAD((advect_tma_kernel<RgbPayload, false>));
    FS_PRELOAD((advect_tma_kernel<RgbPayload, true>));
    FS_PRELOAD(advect_div_tma_kernel);
    FS_PRELOAD(advect_rgb_frame_kernel);
    return 0;
}

}  // namespace fs
