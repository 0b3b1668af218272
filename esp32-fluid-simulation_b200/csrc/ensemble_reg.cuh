// Ensemble kernel, second generation: the whole loop() body (ino:249-289) of one small grid per CTA with
// the PROJECTION (divergence -> red-black SOR -> gradient subtract, ino:274-276) held in REGISTERS.
//
// ncu of the first-generation kernel (ensemble.cu; profiles/r02_ncu_ensemble_*.json) showed it
// instruction-bound at ~1,000 thread-instructions per node-step, half of them in the SOR phase: every
// update loaded p, d and four neighbours from shared memory and decoded a packed node descriptor.  Here a
// thread owns a block of 4 columns x R rows (R even) of the grid for the whole projection:
//   * p, dx*d and the Gauss-Seidel coefficient neg_a_ii_inv[#neighbours] (poisson.cpp:67) of its 4R nodes
//     live in registers; with R even the colour of node (r, c) of the block is (r + c) & 1 for every
//     thread, so after unrolling each half-sweep is straight-line code on named registers;
//   * 3 of 4 neighbours of an update are the thread's own registers; the others come from the adjacent
//     threads through a per-thread MAILBOX in shared memory: after a half-sweep a thread publishes the
//     values it just updated on the rim of its block (one per row towards the left or right thread, two
//     per top/bottom row), the next half-sweep reads the neighbours' — R + 4 loads and stores per 2R
//     updates.  The slots written in a half-sweep (colour q) are never the ones read in it (colour q^1),
//     so ONE barrier per half-sweep suffices;
//   * cells beyond a wall read as +0.0f (a block of zeros stands in for the missing neighbour thread; nodes
//     of the block outside the grid are forced to +0 with an AND mask), which makes ((L+R)+D)+U
//     bit-identical to pois_gs_safe's running sum (sor.cuh: sor_update_coef) — one branch-free path for
//     interior and wall nodes;
//   * the mailboxes alias the velocity buffer that is dead during the projection.
// The two advects keep a node-strided mapping over ALL threads of the CTA, two (R = 2) or four nodes per thread
// at a time: the interior case of sample<T> (advect.h:38-42) is evaluated unconditionally on a clamped cell so
// that the corner loads of several nodes are in flight together, and only nodes whose backtrace left the grid
// are redone through the general sample().
// State I/O (dye resident in shared memory, ens_reg_body_resident): copies between the caller's arrays and shared
// memory are CONGRUENT modulo 16 bytes for any N (the data pointers of a grid are shifted by its address
// modulo 16), so they move in 16-byte pieces; calls of few steps pipeline them as bulk copies (cp.async.bulk +
// mbarriers) under the neighbouring grids' compute — ens_resident_pipelined.
//
// The body is a template over an execution environment (thread id, CTA size, barrier) so that
// tests/emu/ compiles THIS source for the host and runs it on CPU threads against the oracle (also under
// AddressSanitizer).
#pragma once

#include "ensemble_common.cuh"
#ifdef __CUDACC__
#include "tma.cuh"
#endif

namespace fs {

template <int R>
struct EnsRegLayout {
    static_assert(R >= 2 && R % 2 == 0, "R must be even: the colour pattern of a block must not depend on the thread");
    static constexpr int MAIL_WORDS = 2 * R + 8;         // H(r, side) = 2r + side, V(side, c) = 2R + 4 side + c
    static constexpr int MAIL_STRIDE = MAIL_WORDS | 1;   // odd => lanes hit distinct banks
};

// threads that own a block of the grid
__host__ __device__ static inline int ens_reg_blocks(int dim_x, int dim_y, int R) { return ((dim_x + 3) / 4) * ((dim_y + R - 1) / R); }
// one velocity buffer: N nodes + 6 of padding (a block's clamped loads may run 3 nodes past the end, and the data
// may start 8 bytes into the buffer, see ens_reg_body_resident), or the mailboxes of all block threads + the block
// of zeros (+ the same shift), whichever is larger
__host__ __device__ static inline size_t ens_reg_vbuf_bytes(int dim_x, int dim_y, int R)
{
    const size_t v = 8 * ((size_t)dim_x * dim_y + 6);
    const size_t m = 4 * (size_t)((2 * R + 8) | 1) * ((size_t)ens_reg_blocks(dim_x, dim_y, R) + 1) + 16;
    return ((v > m ? v : m) + 127) & ~(size_t)127;      // (buffers start on 128-byte lines: bulk copies measured 7 % slower
}                                                       //  end to end from a buffer that was only 16-byte aligned)
// one dye buffer: 12N bytes + up to 12 bytes of shift + the 16-byte granule a load may over-read
__host__ __device__ static inline size_t ens_reg_cbuf_bytes(int dim_x, int dim_y)
{
    return ((size_t)12 * dim_x * dim_y + 32 + 127) & ~(size_t)127;
}
__host__ __device__ static inline size_t ens_reg_smem_bytes(int dim_x, int dim_y, int R, bool dye_smem)
{
    // dye resident: two dye buffers and two mbarriers (the bulk-copy pipeline of ens_reg_body_resident)
    return 2 * ens_reg_vbuf_bytes(dim_x, dim_y, R) + (dye_smem ? 2 * ens_reg_cbuf_bytes(dim_x, dim_y) + 16 : 0);
}

// Interior case of sample<T> (advect.h:38-42) on a cell clamped into the grid: no branches, so the loads of
// several nodes overlap.  Returns true when the backtrace left [0, GX-1) x [0, GY-1) (equivalent to the tests
// of sample(), advect.h:26-29) — the caller then redoes the node through sample().  Needs GX, GY >= 2.
template <class P, class Fetch>
__device__ __forceinline__ bool sample_interior(typename P::raw_t (&out)[P::NC], const Fetch &fetch, float i, float j,
                                                float gx2, float gy2)     // (float)(GX - 2), (float)(GY - 2)
{
    using raw_t = typename P::raw_t;
    constexpr int NC = P::NC;
    const float i_floor = floorf(i), j_floor = floorf(j);
    const float di = __fsub_rn(i, i_floor), dj = __fsub_rn(j, j_floor);
    const float wi = __fsub_rn(1.0f, di), wj = __fsub_rn(1.0f, dj);
    // clamp in float first: the float -> int conversion of a huge value is only defined on the device
    const float ci = fminf(fmaxf(i_floor, 0.0f), gx2), cj = fminf(fmaxf(j_floor, 0.0f), gy2);
    // 0 <= i < GX-1  <=>  0 <= floor(i) <= GX-2  <=>  the clamp left floor(i) alone.  (A NaN passes fmaxf as 0 and
    // compares unequal: it is redone by sample(), which takes the same cell (0, .) through cvt.rzi.)
    const bool oob = ci != i_floor || cj != j_floor;
    const int gi = (int)ci, gj = (int)cj;
    raw_t p11[NC], p12[NC], p21[NC], p22[NC];
    fetch(gi, gj, p11);
    fetch(gi, gj + 1, p12);
    fetch(gi + 1, gj, p21);
    fetch(gi + 1, gj + 1, p22);
#pragma unroll
    for (int ch = 0; ch < NC; ch++) {
        const float a = mixf(wj, dj, P::to_float(p11[ch]), P::to_float(p12[ch]));
        const float b = mixf(wj, dj, P::to_float(p21[ch]), P::to_float(p22[ch]));
        out[ch] = P::from_float(mixf(wi, di, a, b));
    }
    return oob;
}

// One colour half-sweep (poisson.cpp:14-61: colour 0 = (i+j) even first) over a thread's block.
// FIRST: every p is still +0 (poisson.cpp:117-119), nothing to read.
// MASKED: the block may hold nodes outside the grid (dim_x % 4 or dim_y % R nonzero), kept at +0 by an AND.
template <int R, int Q, bool FIRST, bool MASKED>
__device__ __forceinline__ void ens_half_sweep(float (&p)[R][4], const float (&d)[R][4], const float (&coef)[R][4],
                                               const unsigned (&cm)[4], const unsigned (&rm)[R], float *mine,
                                               const float *left, const float *right, const float *down,
                                               const float *up, const SorCoef &k)
{
    // the neighbours' rim values first, so that all loads are in flight together
    float hv[R], vd[2], vu[2];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int c0 = (Q + r) & 1;                     // this row updates columns c0 and c0 + 2
        // c0 == 0: column 0 needs the left thread's column 3; c0 == 1: column 3 needs the right thread's column 0
        hv[r] = FIRST ? 0.0f : (c0 == 0 ? left[2 * r + 1] : right[2 * r + 0]);
    }
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        vd[kk] = FIRST ? 0.0f : down[2 * R + 4 + Q + 2 * kk];       // row 0 updates columns Q, Q+2: row R-1 of the thread below
        vu[kk] = FIRST ? 0.0f : up[2 * R + (Q ^ 1) + 2 * kk];       // row R-1 (odd) updates columns Q^1, (Q^1)+2: row 0 of the thread above
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int c0 = (Q + r) & 1;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {
            const int c = c0 + 2 * kk;
            const float l = c > 0 ? p[r][c > 0 ? c - 1 : 0] : hv[r];
            const float rt = c < 3 ? p[r][c < 3 ? c + 1 : 3] : hv[r];
            const float dn = r > 0 ? p[r > 0 ? r - 1 : 0][c] : vd[kk];
            const float u = r < R - 1 ? p[r < R - 1 ? r + 1 : R - 1][c] : vu[kk];
            const float nv = sor_update_coef(p[r][c], l, rt, dn, u, d[r][c], coef[r][c], k);
            p[r][c] = MASKED ? __uint_as_float(__float_as_uint(nv) & cm[c] & rm[r]) : nv;   // nodes outside the grid stay +0
        }
        mine[2 * r + c0] = c0 == 0 ? p[r][0] : p[r][3];
    }
    mine[2 * R + Q] = p[0][Q];
    mine[2 * R + Q + 2] = p[0][Q + 2];
    mine[2 * R + 4 + (Q ^ 1)] = p[R - 1][Q ^ 1];
    mine[2 * R + 4 + (Q ^ 1) + 2] = p[R - 1][(Q ^ 1) + 2];
}

// poisson_solve's iteration loop (poisson.cpp:114-125) for one thread; every thread of the CTA calls it (barriers)
template <int R, bool MASKED, class Env>
__device__ __forceinline__ void ens_sor(float (&p)[R][4], const float (&d)[R][4], const float (&coef)[R][4],
                                        const unsigned (&cm)[4], const unsigned (&rm)[R], float *mine, const float *left,
                                        const float *right, const float *down, const float *up, const SorCoef &k,
                                        int iters, bool act, const Env &env)
{
    if (act) ens_half_sweep<R, 0, true, MASKED>(p, d, coef, cm, rm, mine, left, right, down, up, k);
    env.sync();
    if (act) ens_half_sweep<R, 1, false, MASKED>(p, d, coef, cm, rm, mine, left, right, down, up, k);
    env.sync();
    for (int it = 1; it < iters; it++) {
        if (act) ens_half_sweep<R, 0, false, MASKED>(p, d, coef, cm, rm, mine, left, right, down, up, k);
        env.sync();
        if (act) ens_half_sweep<R, 1, false, MASKED>(p, d, coef, cm, rm, mine, left, right, down, up, k);
        env.sync();
    }
}

// keeps the compiler from turning `x & mask` back into the comparison the mask came from (two selects and a
// compare per use instead of one 3-input logic op), and from rematerialising an offset in every half-sweep
#define FS_OPAQUE_REG(x) asm volatile("" : "+r"(x))

// Everything a thread knows about its place in the CTA: its 4 x R block for the projection, its mailbox
// neighbours, its node-strided walk for the advects.
template <int R>
struct EnsMap {
    int tid, NT, dim_x, dim_y, N, CG, RS, NS;
    bool ragged;          // some blocks reach beyond the grid
    bool even_x;          // block rows start 16-byte aligned
    bool act;             // this thread owns a block
    int s, g, i0, j0;
    unsigned cm[4], rm[R];
    float coef[R][4];
    int row[R];           // node index of (i0, j0 + r), row clamped into the grid
    int o_zero, o_mine, o_left, o_right, o_down, o_up;   // mailboxes (word offsets into the dead velocity buffer)
    int row_dn, row_up, off_l;                           // the block's ring in the velocity array, clamped into the buffer
    float adv_di, adv_dj, adv_i0, adv_j0, fdim_x, gx2, gy2;   // node-strided mapping of the advects, as floats (exact)
};

template <int R>
__device__ __forceinline__ void ens_map_init(EnsMap<R> &m, const EnsArgs &a, int tid, int NT)
{
    constexpr int MS = EnsRegLayout<R>::MAIL_STRIDE;
    const int dim_x = a.dim_x, dim_y = a.dim_y;
    m.tid = tid; m.NT = NT; m.dim_x = dim_x; m.dim_y = dim_y; m.N = dim_x * dim_y;
    m.CG = (dim_x + 3) >> 2; m.RS = (dim_y + R - 1) / R; m.NS = m.CG * m.RS;
    m.ragged = (dim_x & 3) != 0 || dim_y % R != 0;
    m.even_x = (dim_x & 1) == 0;
    m.act = tid < m.NS;
    m.s = m.act ? tid / m.CG : 0;
    m.g = m.act ? tid - m.s * m.CG : 0;
    m.i0 = 4 * m.g; m.j0 = R * m.s;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        m.cm[c] = m.i0 + c < dim_x ? 0xffffffffu : 0u;
        FS_OPAQUE_REG(m.cm[c]);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        m.rm[r] = m.j0 + r < dim_y ? 0xffffffffu : 0u;
        FS_OPAQUE_REG(m.rm[r]);
        m.row[r] = min(m.j0 + r, dim_y - 1) * dim_x + m.i0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = m.i0 + c, j = m.j0 + r;
            const int nb = 4 - (i == 0) - (i == dim_x - 1) - (j == 0) - (j == dim_y - 1);
            m.coef[r][c] = nb == 4 ? a.k.neg_quarter : nb == 3 ? a.k.neg_third : a.k.neg_half;
        }
    }
    // a missing neighbour thread is the block of zeros
    m.o_zero = m.NS * MS;
    m.o_mine = tid * MS;
    m.o_left = m.g > 0 ? m.o_mine - MS : m.o_zero;
    m.o_right = m.g < m.CG - 1 ? m.o_mine + MS : m.o_zero;
    m.o_down = m.s > 0 ? m.o_mine - m.CG * MS : m.o_zero;
    m.o_up = m.s < m.RS - 1 ? m.o_mine + m.CG * MS : m.o_zero;
    // (kept in registers: recomputing them cost ~25 integer instructions in every half-sweep)
    FS_OPAQUE_REG(m.o_left); FS_OPAQUE_REG(m.o_right); FS_OPAQUE_REG(m.o_down); FS_OPAQUE_REG(m.o_up);
    m.row_dn = m.s > 0 ? m.row[0] - dim_x : m.row[0];
    m.row_up = min(m.j0 + R, dim_y - 1) * dim_x + m.i0;
    m.off_l = m.g > 0 ? -1 : 0;
    m.adv_di = (float)(NT % dim_x); m.adv_dj = (float)(NT / dim_x);
    m.adv_i0 = (float)(tid % dim_x); m.adv_j0 = (float)(tid / dim_x);
    m.fdim_x = (float)dim_x; m.gx2 = (float)(dim_x - 2); m.gy2 = (float)(dim_y - 2);
}

// ---- advect velocity, no-slip (ino:253): A -> B ------------------------------------------------------------
template <int R, int U>
__device__ __forceinline__ void ens_advect_velocity(const EnsMap<R> &m, const EnsArgs &a, const float2 *A, float2 *B)
{
    const int N = m.N, NT = m.NT;
    SmemFetch<Vec2Payload> fetch{reinterpret_cast<const float *>(A), m.dim_x};
    float fi = m.adv_i0, fj = m.adv_j0;             // node coordinates as floats (exact), advect.h:81
    for (int nb = m.tid; nb < N; nb += U * NT) {
        float res[U][2];
        float si[U], sj[U];
        bool oob[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int n = min(nb + u * NT, N - 1);      // (clamped: the surplus lanes recompute the last node)
            const float2 vel = A[n];
            si[u] = __fsub_rn(fi, __fmul_rn(vel.x, a.dt));
            sj[u] = __fsub_rn(fj, __fmul_rn(vel.y, a.dt));
            oob[u] = sample_interior<Vec2Payload>(res[u], fetch, si[u], sj[u], m.gx2, m.gy2);
            fi += m.adv_di;
            fj += m.adv_dj;
            if (fi >= m.fdim_x) { fi -= m.fdim_x; fj += 1.0f; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (nb + u * NT < N) {
                if (oob[u]) sample<Vec2Payload>(res[u], fetch, si[u], sj[u], m.dim_x, m.dim_y, true);
                B[nb + u * NT] = make_float2(res[u][0], res[u][1]);
            }
        }
    }
}

// ---- drags (ino:264-269): in order, one thread; ends in a barrier when there are records -------------------
template <int R, class Env>
__device__ __forceinline__ void ens_drags(const EnsMap<R> &m, const EnsArgs &a, float2 *B, int step, int grid, const Env &env)
{
    if (a.max_drags > 0) {
        if (m.tid == 0) {
            const size_t slot = (size_t)step * a.batch + grid;
            const int cnt = min(a.counts[slot], a.max_drags);
            const fs_drag *dr = a.drags + slot * a.max_drags;
            for (int q = 0; q < cnt; q++) {
                const fs_drag d = dr[q];
                if (d.cy < m.dim_x && d.cx < m.dim_y) B[d.cx * m.dim_x + d.cy] = make_float2(d.vy, d.vx);
            }
        }
        env.sync();
    }
}

// ---- projection in registers (ino:274-276): divergence of B, red-black SOR, gradient subtracted from B in place.
// `mail` = the velocity buffer that is dead now.  Ends BEFORE the barrier that publishes B.
// rows16: dim_x is even AND B itself is 16-byte aligned, so every row of a block starts on a 16-byte boundary.
template <int R, class Env>
__device__ __forceinline__ void ens_project(const EnsMap<R> &m, const EnsArgs &a, float *mail, float2 *B, bool rows16, const Env &env)
{
    constexpr int MS = EnsRegLayout<R>::MAIL_STRIDE;
    const int dim_x = m.dim_x, dim_y = m.dim_y, i0 = m.i0, j0 = m.j0;
    float p[R][4], d[R][4];
    if (m.tid < MS) mail[m.o_zero + m.tid] = 0.0f;       // (first read after the first half-sweep's barrier)
    if (m.act) {
        // divergence (ino:274, finitediff.cpp:9-39) of the block from B and its one-node ring
        float vx[R][4], vy[R][4], xl[R], xr[R], yd[4], yu[4];
        if (rows16) {
            // every row of the block starts 16-byte aligned, two nodes per load
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const float4 t = *reinterpret_cast<const float4 *>(B + m.row[r] + 2 * h);
                    vx[r][2 * h] = t.x; vy[r][2 * h] = t.y; vx[r][2 * h + 1] = t.z; vy[r][2 * h + 1] = t.w;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float4 td = *reinterpret_cast<const float4 *>(B + m.row_dn + 2 * h);
                const float4 tu = *reinterpret_cast<const float4 *>(B + m.row_up + 2 * h);
                yd[2 * h] = td.y; yd[2 * h + 1] = td.w; yu[2 * h] = tu.y; yu[2 * h + 1] = tu.w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float2 t = B[m.row[r] + c];
                    vx[r][c] = t.x;
                    vy[r][c] = t.y;
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                yd[c] = B[m.row_dn + c].y;
                yu[c] = B[m.row_up + c].y;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            xl[r] = B[m.row[r] + m.off_l].x;
            xr[r] = B[m.row[r] + 4].x;
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int i = i0 + c, j = j0 + r;
                const float xm = c > 0 ? vx[r][c > 0 ? c - 1 : 0] : xl[r], xp = c < 3 ? vx[r][c < 3 ? c + 1 : 3] : xr[r];
                const float ym = r > 0 ? vy[r > 0 ? r - 1 : 0][c] : yd[c], yp = r < R - 1 ? vy[r < R - 1 ? r + 1 : R - 1][c] : yu[c];
                const bool wall = i == 0 || i == dim_x - 1 || j == 0 || j == dim_y - 1;
                float sd;
                if (!wall) {
                    sd = __fadd_rn(__fadd_rn(-xm, xp), __fadd_rn(-ym, yp));
                } else {
                    sd = 0.0f;
                    sd = __fadd_rn(sd, i > 0 ? -xm : vx[r][c]);
                    sd = __fadd_rn(sd, i < dim_x - 1 ? xp : -vx[r][c]);
                    sd = __fadd_rn(sd, j > 0 ? -ym : vy[r][c]);
                    sd = __fadd_rn(sd, j < dim_y - 1 ? yp : -vy[r][c]);
                }
                // dx*d: the same product every iteration (poisson.cpp:88,109)
                const float dd = __fmul_rn(a.k.dx, __fmul_rn(sd, a.two_dx_inv));
                d[r][c] = __uint_as_float(__float_as_uint(dd) & m.cm[c] & m.rm[r]);
                p[r][c] = 0.0f;                    // poisson.cpp:117-119
            }
        }
    }
    // ---- red-black SOR (ino:275) ---------------------------------------------------------------------
    if (a.iters > 0) {
        if (m.ragged)
            ens_sor<R, true>(p, d, m.coef, m.cm, m.rm, mail + m.o_mine, mail + m.o_left, mail + m.o_right, mail + m.o_down, mail + m.o_up, a.k, a.iters, m.act, env);
        else
            ens_sor<R, false>(p, d, m.coef, m.cm, m.rm, mail + m.o_mine, mail + m.o_left, mail + m.o_right, mail + m.o_down, mail + m.o_up, a.k, a.iters, m.act, env);
    } else {
        if (m.act)
            for (int w = 0; w < MS; w++) mail[m.o_mine + w] = 0.0f;
        env.sync();
    }
    // ---- subtract gradient (ino:276, finitediff.cpp:41-82), in place on B ----------------------------
    if (m.act) {
        float hl[R], hr[R], vd[4], vu[4];
#pragma unroll
        for (int r = 0; r < R; r++) {
            hl[r] = mail[m.o_left + 2 * r + 1];
            hr[r] = mail[m.o_right + 2 * r + 0];
        }
#pragma unroll
        for (int c = 0; c < 4; c++) {
            vd[c] = mail[m.o_down + 2 * R + 4 + c];
            vu[c] = mail[m.o_up + 2 * R + c];
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            float gx[4], gy[4];                    // (p_right - p_left) / 2dx, (p_up - p_down) / 2dx
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int i = i0 + c, j = j0 + r;
                const float pc = p[r][c];
                const float pl = i > 0 ? (c > 0 ? p[r][c > 0 ? c - 1 : 0] : hl[r]) : pc;
                const float pr = i < dim_x - 1 ? (c < 3 ? p[r][c < 3 ? c + 1 : 3] : hr[r]) : pc;
                const float pd = j > 0 ? (r > 0 ? p[r > 0 ? r - 1 : 0][c] : vd[c]) : pc;
                const float pu = j < dim_y - 1 ? (r < R - 1 ? p[r < R - 1 ? r + 1 : R - 1][c] : vu[c]) : pc;
                gx[c] = __fmul_rn(__fsub_rn(pr, pl), a.two_dx_inv);
                gy[c] = __fmul_rn(__fsub_rn(pu, pd), a.two_dx_inv);
            }
            if (rows16) {
                // dim_x even: nodes are inside the grid in pairs, 16 bytes per access
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (m.cm[2 * h] & m.rm[r]) {
                        float4 t = *reinterpret_cast<const float4 *>(B + m.row[r] + 2 * h);
                        t.x = __fsub_rn(t.x, gx[2 * h]);
                        t.y = __fsub_rn(t.y, gy[2 * h]);
                        t.z = __fsub_rn(t.z, gx[2 * h + 1]);
                        t.w = __fsub_rn(t.w, gy[2 * h + 1]);
                        *reinterpret_cast<float4 *>(B + m.row[r] + 2 * h) = t;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    if (m.cm[c] & m.rm[r]) {
                        float2 c0 = B[m.row[r] + c];
                        c0.x = __fsub_rn(c0.x, gx[c]);
                        c0.y = __fsub_rn(c0.y, gy[c]);
                        B[m.row[r] + c] = c0;
                    }
                }
            }
        }
    }
}

// ---- advect dye, free-slip sampling (ino:282) with the projected velocity B: C1 -> C2 (distinct buffers) ----
template <int R, int U, class Fetch>
__device__ __forceinline__ void ens_advect_dye(const EnsMap<R> &m, const EnsArgs &a, const float2 *B, const Fetch &fetch, uint32_t *C2)
{
    const int N = m.N, NT = m.NT;
    float fi = m.adv_i0, fj = m.adv_j0;
    for (int nb = m.tid; nb < N; nb += U * NT) {
        uint32_t res[U][3];
        float si[U], sj[U];
        bool oob[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int n = min(nb + u * NT, N - 1);
            const float2 vel = B[n];
            si[u] = __fsub_rn(fi, __fmul_rn(vel.x, a.dt));
            sj[u] = __fsub_rn(fj, __fmul_rn(vel.y, a.dt));
            oob[u] = sample_interior<RgbPayload>(res[u], fetch, si[u], sj[u], m.gx2, m.gy2);
            fi += m.adv_di;
            fj += m.adv_dj;
            if (fi >= m.fdim_x) { fi -= m.fdim_x; fj += 1.0f; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int n = nb + u * NT;
            if (n < N) {
                if (oob[u]) sample<RgbPayload>(res[u], fetch, si[u], sj[u], m.dim_x, m.dim_y, false);
                C2[3 * n + 0] = res[u][0];
                C2[3 * n + 1] = res[u][1];
                C2[3 * n + 2] = res[u][2];
            }
        }
    }
}

// The same advect IN PLACE: every thread keeps the results of all its nodes (at most 4R: the CTA has at least one
// thread per 4 x R block) in registers until the whole CTA has finished reading C, then writes them back — so the
// second dye buffer is free to receive the NEXT grid's dye while this one is being stepped.
template <int R, int U, class Env>
__device__ __forceinline__ void ens_advect_dye_in_place(const EnsMap<R> &m, const EnsArgs &a, const float2 *B, uint32_t *C, const Env &env)
{
    constexpr int KN = 4 * R;
    static_assert(KN % U == 0, "");
    const int N = m.N, NT = m.NT;
    SmemFetch<RgbPayload> fetch{C, m.dim_x};
    uint32_t res[KN][3];
    float fi = m.adv_i0, fj = m.adv_j0;
#pragma unroll
    for (int k0 = 0; k0 < KN; k0 += U) {
        if (m.tid + k0 * NT < N) {
            float si[U], sj[U];
            bool oob[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int n = min(m.tid + (k0 + u) * NT, N - 1);
                const float2 vel = B[n];
                si[u] = __fsub_rn(fi, __fmul_rn(vel.x, a.dt));
                sj[u] = __fsub_rn(fj, __fmul_rn(vel.y, a.dt));
                oob[u] = sample_interior<RgbPayload>(res[k0 + u], fetch, si[u], sj[u], m.gx2, m.gy2);
                fi += m.adv_di;
                fj += m.adv_dj;
                if (fi >= m.fdim_x) { fi -= m.fdim_x; fj += 1.0f; }
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (oob[u] && m.tid + (k0 + u) * NT < N) sample<RgbPayload>(res[k0 + u], fetch, si[u], sj[u], m.dim_x, m.dim_y, false);
        }
    }
    env.sync();                                 // every corner has been read
#pragma unroll
    for (int k = 0; k < KN; k++) {
        const int n = m.tid + k * NT;
        if (n < N) {
            C[3 * n + 0] = res[k][0];
            C[3 * n + 1] = res[k][1];
            C[3 * n + 2] = res[k][2];
        }
    }
}

// cooperative copies between global and shared memory: 16 bytes per access when everything is aligned, else 8 (the
// velocity) or 4 (the dye).  BATCH loads are issued before the first store: the phase is latency-bound, what counts
// is the bytes in flight per thread (a plain loop kept ONE load in flight).
template <class T, int BATCH>
__device__ __forceinline__ void ens_copy_batched(T *dst, const T *src, int cnt, int tid, int NT, bool from_global)
{
    for (int n = tid; n < cnt; n += BATCH * NT) {
        T t[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; u++) {       // (index clamped, load unconditional: keeps t[] in registers)
            const int k = min(n + u * NT, cnt - 1);
            t[u] = from_global ? __ldg(src + k) : src[k];
        }
#pragma unroll
        for (int u = 0; u < BATCH; u++)
            if (n + u * NT < cnt) dst[n + u * NT] = t[u];
    }
}
// words = 32-bit words to copy; wide8 = the data is a whole number of 8-byte elements, 8-byte aligned (velocity)
__device__ __forceinline__ void ens_copy_in(void *dst_smem, const void *src_gmem, int words, bool vec16, bool wide8, int tid, int NT)
{
    if (vec16)
        ens_copy_batched<uint4, 4>(reinterpret_cast<uint4 *>(dst_smem), reinterpret_cast<const uint4 *>(src_gmem), words / 4, tid, NT, true);
    else if (wide8)
        ens_copy_batched<uint2, 8>(reinterpret_cast<uint2 *>(dst_smem), reinterpret_cast<const uint2 *>(src_gmem), words / 2, tid, NT, true);
    else
        ens_copy_batched<uint32_t, 8>(reinterpret_cast<uint32_t *>(dst_smem), reinterpret_cast<const uint32_t *>(src_gmem), words, tid, NT, true);
}
__device__ __forceinline__ void ens_copy_out(void *dst_gmem, const void *src_smem, int words, bool vec16, bool wide8, int tid, int NT)
{
    if (vec16)
        ens_copy_batched<uint4, 4>(reinterpret_cast<uint4 *>(dst_gmem), reinterpret_cast<const uint4 *>(src_smem), words / 4, tid, NT, false);
    else if (wide8)
        ens_copy_batched<uint2, 8>(reinterpret_cast<uint2 *>(dst_gmem), reinterpret_cast<const uint2 *>(src_smem), words / 2, tid, NT, false);
    else
        ens_copy_batched<uint32_t, 8>(reinterpret_cast<uint32_t *>(dst_gmem), reinterpret_cast<const uint32_t *>(src_smem), words, tid, NT, false);
}

// CONGRUENT copies: source and destination have the same address modulo 16, so the part of the data that lies between
// 16-byte boundaries moves in 16-byte pieces whatever N is (for odd N a grid's arrays start 4, 8 or 12 bytes off such a
// boundary; the kernel shifts its shared-memory data pointers by the same amount).
// Loads take the whole 16-byte granules the data touches (the few foreign bytes land in the buffers' slack) — except a
// last granule that would reach beyond the end of the caller's ARRAY, which is copied by words.
struct EnsCongLoad {
    const unsigned char *s0;   // first granule
    unsigned char *d0;
    uint32_t vec_bytes;        // granules copied whole
    uint32_t tail_words;       // words after them (last grid of the array only)
};
__device__ __forceinline__ EnsCongLoad ens_cong_load_plan(void *dst_data, const void *src, uint32_t bytes, const void *arr_end)
{
    const unsigned char *sp = reinterpret_cast<const unsigned char *>(src);
    if (((reinterpret_cast<uintptr_t>(sp) | bytes) & 15) == 0) {     // everything on 16-byte boundaries (N % 4 == 0): no head,
        EnsCongLoad q;                                               // no foreign bytes — the short way (one thread makes
        q.s0 = sp;                                                   // these plans between two barriers of the whole CTA)
        q.d0 = reinterpret_cast<unsigned char *>(dst_data);
        q.vec_bytes = bytes;
        q.tail_words = 0;
        return q;
    }
    const uint32_t h = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 15);
    const unsigned char *end = sp + bytes;
    const uint32_t over = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(end) & 15)) & 15);
    const bool whole = end + over <= reinterpret_cast<const unsigned char *>(arr_end);
    EnsCongLoad p;
    p.s0 = sp - h;
    p.d0 = reinterpret_cast<unsigned char *>(dst_data) - h;
    const unsigned char *vend = whole ? end + over : end - (reinterpret_cast<uintptr_t>(end) & 15);
    p.vec_bytes = (uint32_t)(vend - p.s0);
    p.tail_words = whole ? 0u : (uint32_t)(end - vend) / 4;
    return p;
}
__device__ __forceinline__ void ens_copy_cong_in(void *dst_data, const void *src, uint32_t bytes, const void *arr_end, int tid, int NT)
{
    const EnsCongLoad p = ens_cong_load_plan(dst_data, src, bytes, arr_end);
    ens_copy_batched<uint4, 4>(reinterpret_cast<uint4 *>(p.d0), reinterpret_cast<const uint4 *>(p.s0), (int)(p.vec_bytes / 16), tid, NT, true);
    if (tid < (int)p.tail_words)
        reinterpret_cast<uint32_t *>(p.d0 + p.vec_bytes)[tid] = __ldg(reinterpret_cast<const uint32_t *>(p.s0 + p.vec_bytes) + tid);
}
// Stores must not touch foreign bytes: words up to the first boundary, 16-byte pieces, words after the last boundary.
struct EnsCongStore {
    uint32_t head_words, body_bytes, tail_words;
};
__device__ __forceinline__ EnsCongStore ens_cong_store_plan(const void *dst, uint32_t bytes)
{
    EnsCongStore p;
    if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0) {
        p.head_words = 0;
        p.body_bytes = bytes;
        p.tail_words = 0;
        return p;
    }
    uint32_t h = (uint32_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
    if (h > bytes) h = bytes;
    p.head_words = h / 4;
    p.body_bytes = (bytes - h) & ~15u;
    p.tail_words = (bytes - h - p.body_bytes) / 4;
    return p;
}
__device__ __forceinline__ void ens_copy_cong_out(void *dst, const void *src_data, uint32_t bytes, int tid, int NT)
{
    const EnsCongStore p = ens_cong_store_plan(dst, bytes);
    unsigned char *dp = reinterpret_cast<unsigned char *>(dst);
    const unsigned char *sp = reinterpret_cast<const unsigned char *>(src_data);
    const uint32_t h = p.head_words * 4;
    if (tid < (int)p.head_words) reinterpret_cast<uint32_t *>(dp)[tid] = reinterpret_cast<const uint32_t *>(sp)[tid];
    ens_copy_batched<uint4, 4>(reinterpret_cast<uint4 *>(dp + h), reinterpret_cast<const uint4 *>(sp + h), (int)(p.body_bytes / 16), tid, NT, false);
    if (tid < (int)p.tail_words)
        reinterpret_cast<uint32_t *>(dp + h + p.body_bytes)[tid] = reinterpret_cast<const uint32_t *>(sp + h + p.body_bytes)[tid];
}

#ifdef __CUDACC__
// 1-D bulk copies (TMA) and their completion mechanisms
__device__ __forceinline__ void ens_bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    mbar_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void ens_bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void ens_bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void ens_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// the congruent copies as bulk copies, issued by ONE thread from plans it made earlier (off the critical path)
__device__ __forceinline__ void ens_bulk_load_cong(const EnsCongLoad &p, uint64_t *bar)
{
    ens_bulk_load(p.d0, p.s0, p.vec_bytes, bar);
    for (uint32_t w = 0; w < p.tail_words; w++)          // (the last grid of the caller's array only)
        reinterpret_cast<uint32_t *>(p.d0 + p.vec_bytes)[w] = __ldg(reinterpret_cast<const uint32_t *>(p.s0 + p.vec_bytes) + w);
}
__device__ __forceinline__ void ens_bulk_store_cong(const EnsCongStore &p, void *dst, const void *src_data)
{
    unsigned char *dp = reinterpret_cast<unsigned char *>(dst);
    const unsigned char *sp = reinterpret_cast<const unsigned char *>(src_data);
    const uint32_t h = p.head_words * 4;
    if (p.body_bytes)
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dp + h), "r"(smem_u32(sp + h)), "r"(p.body_bytes)
                     : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");   // (always one group per store: the wait_group counts rely on it)
    for (uint32_t w = 0; w < p.head_words; w++) reinterpret_cast<uint32_t *>(dp)[w] = reinterpret_cast<const uint32_t *>(sp)[w];
    for (uint32_t w = 0; w < p.tail_words; w++)
        reinterpret_cast<uint32_t *>(dp + h + p.body_bytes)[w] = reinterpret_cast<const uint32_t *>(sp + h + p.body_bytes)[w];
}
__device__ __forceinline__ void ens_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// Env: { int tid, nthreads, block, nblocks; void sync() const; static constexpr bool kAsync, kEmulatePipe; }
//
// Dye RESIDENT in shared memory (40 B/node: velocity x2, dye x2).  `cong`: the caller's arrays are 16-byte aligned, so
// every copy between global and shared memory is CONGRUENT (see above): the data pointers of a grid sit `hv` / `hc`
// bytes into their buffers, hv / hc = the grid's velocity / dye address modulo 16 (0 unless N is odd / not a
// multiple of 4).  Otherwise plain 8- and 4-byte copies.
struct EnsResident {
    unsigned char *va, *vb;        // velocity buffers: advect source (then mailboxes), advect destination / projection
    unsigned char *ca, *cb;        // dye buffers: current, other
    const unsigned char *v_end, *c_end;   // ends of the caller's arrays
    bool cong;
};

// State I/O at the grid boundaries: load, step n_steps times (dye ping-pong between the two buffers, ino:286), store.
// The next grid's state is pulled into L2 during the last step so that its load pays the L2 latency instead of HBM's.
template <int R, class Env>
__device__ __forceinline__ void ens_resident_sync(const EnsMap<R> &m, const EnsArgs &a, EnsResident rs, const Env &env)
{
    constexpr int U = R == 2 ? 2 : 4;                      // nodes per thread in flight in the advects
    const int tid = m.tid, NT = m.NT, N = m.N;
    for (int grid = env.block; grid < a.batch; grid += env.nblocks) {
        float2 *gv = a.v + (size_t)grid * N;
        uint32_t *gc = a.c + (size_t)grid * N * 3;
        const uint32_t hv = rs.cong ? (uint32_t)(reinterpret_cast<uintptr_t>(gv) & 15) : 0;
        const uint32_t hc = rs.cong ? (uint32_t)(reinterpret_cast<uintptr_t>(gc) & 15) : 0;
        float2 *A = reinterpret_cast<float2 *>(rs.va + hv), *B = reinterpret_cast<float2 *>(rs.vb + hv);
        uint32_t *C1 = reinterpret_cast<uint32_t *>(rs.ca + hc), *C2 = reinterpret_cast<uint32_t *>(rs.cb + hc);
        const bool rows16 = m.even_x && hv == 0;
        if (rs.cong) {
            ens_copy_cong_in(A, gv, 8u * N, rs.v_end, tid, NT);
            ens_copy_cong_in(C1, gc, 12u * N, rs.c_end, tid, NT);
        } else {
            ens_copy_in(A, gv, 2 * N, false, true, tid, NT);
            ens_copy_in(C1, gc, 3 * N, false, false, tid, NT);
        }
        env.sync();
        for (int step = 0; step < a.n_steps; step++) {
            if (step == a.n_steps - 1 && grid + env.nblocks < a.batch) {
                const char *nv = reinterpret_cast<const char *>(a.v + (size_t)(grid + env.nblocks) * N);
                const char *nc = reinterpret_cast<const char *>(a.c + (size_t)(grid + env.nblocks) * N * 3);
                for (int l = tid * 128; l < N * 8; l += NT * 128) prefetch_l2(nv + l);
                for (int l = tid * 128; l < N * 12; l += NT * 128) prefetch_l2(nc + l);
            }
            ens_advect_velocity<R, U>(m, a, A, B);
            env.sync();
            ens_drags(m, a, B, step, grid, env);
            ens_project(m, a, reinterpret_cast<float *>(A), B, rows16, env);
            env.sync();
            ens_advect_dye<R, U>(m, a, B, SmemFetch<RgbPayload>{C1, m.dim_x}, C2);
            env.sync();
            // pointer swaps of ino:255 and ino:286
            uint32_t *tc = C1; C1 = C2; C2 = tc;
            float2 *tv = A; A = B; B = tv;
        }
        if (rs.cong) {
            ens_copy_cong_out(gv, A, 8u * N, tid, NT);
            ens_copy_cong_out(gc, C1, 12u * N, tid, NT);
        } else {
            ens_copy_out(gv, A, 2 * N, false, true, tid, NT);
            ens_copy_out(gc, C1, 3 * N, false, false, tid, NT);
        }
        env.sync();
    }
}

// The same with the state I/O OFF the critical path (needs `cong`).  In a grid's LAST step the dye is advected in place,
// so the other dye buffer is free from the start of that step: it receives the next grid's dye.  The next grid's velocity
// lands in the dead velocity buffer during that step's dye advect, while this grid's projected velocity is already on its
// way out; this grid's dye leaves while the next grid's first advect runs.  On the device (`async`) the copies are 1-D
// bulk copies (cp.async.bulk, SASS UBLKCP) issued by thread 0 and tracked by two mbarriers / bulk groups; in the host
// emulation (Env::kEmulatePipe) the same copies are made cooperatively at the same program points.
template <int R, class Env>
__device__ __forceinline__ void ens_resident_pipelined(const EnsMap<R> &m, const EnsArgs &a, EnsResident rs, unsigned char *bars,
                                                       bool async, const Env &env)
{
    constexpr int U = R == 2 ? 2 : 4;
    const int tid = m.tid, NT = m.NT, N = m.N;
    const uint32_t v_bytes = 8u * N, c_bytes = 12u * N;
#ifdef __CUDACC__
    uint64_t *bar_v = reinterpret_cast<uint64_t *>(bars), *bar_c = bar_v + 1;
    uint32_t phase_v = 0, phase_c = 0;
#else
    async = false;
    (void)bars;
#endif
    auto head = [](const void *p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) & 15); };
#ifdef __CUDACC__
    // N % 4 == 0: every grid's arrays start and end on 16-byte boundaries — the bulk copies need no plan (thread 0 issues
    // them between two barriers of the whole CTA, so its path is kept short; measured: no difference, 6.37 ms either way —
    // the 4 % that the 80x60 one-step call lost when the copies became congruent, 6.08 -> 6.37 ms, are elsewhere).
    const bool trivial = (N & 3) == 0;
    auto bulk_in = [&](unsigned char *buf, const void *src, uint32_t bytes, const unsigned char *end, uint64_t *bar) {
        if (trivial)
            ens_bulk_load(buf, src, bytes, bar);
        else
            ens_bulk_load_cong(ens_cong_load_plan(buf + head(src), src, bytes, end), bar);
    };
    auto bulk_out = [&](void *dst, const void *src_data, uint32_t bytes) {
        if (trivial)
            ens_bulk_store(dst, src_data, bytes);
        else
            ens_bulk_store_cong(ens_cong_store_plan(dst, bytes), dst, src_data);
    };
#endif
    // (Tried: issuing from a lane that owns no block and idles through the projection, with the copies' plans made there
    // and parked in shared memory — the extra live state pushed the kernel from 64 to 330 bytes of spills: 6.3 -> 7.7 ms.)

    int grid = env.block;
    if (grid >= a.batch) return;
    // ---- prologue: the first grid's state ------------------------------------------------------------------
    {
        const float2 *gv = a.v + (size_t)grid * N;
        const uint32_t *gc = a.c + (size_t)grid * N * 3;
#ifdef __CUDACC__
        if (async) {
            if (tid == 0) {
                mbar_init(bar_v, 1);
                mbar_init(bar_c, 1);
                ens_fence_proxy_async();
            }
            env.sync();
            if (tid == 0) {
                bulk_in(rs.va, gv, v_bytes, rs.v_end, bar_v);
                bulk_in(rs.ca, gc, c_bytes, rs.c_end, bar_c);
            }
            mbar_wait(bar_v, phase_v); phase_v ^= 1;
            mbar_wait(bar_c, phase_c); phase_c ^= 1;
        } else
#endif
        {
            ens_copy_cong_in(rs.va + head(gv), gv, v_bytes, rs.v_end, tid, NT);
            ens_copy_cong_in(rs.ca + head(gc), gc, c_bytes, rs.c_end, tid, NT);
        }
        env.sync();
    }

    for (; grid < a.batch; grid += env.nblocks) {
        const int next = grid + env.nblocks;
        const bool has_next = next < a.batch;
        float2 *gv = a.v + (size_t)grid * N, *gv_next = a.v + (size_t)next * N;
        uint32_t *gc = a.c + (size_t)grid * N * 3, *gc_next = a.c + (size_t)next * N * 3;
        const uint32_t hv = head(gv), hc = head(gc);
        float2 *A = reinterpret_cast<float2 *>(rs.va + hv), *B = reinterpret_cast<float2 *>(rs.vb + hv);
        uint32_t *C_cur = reinterpret_cast<uint32_t *>(rs.ca + hc), *C_oth = reinterpret_cast<uint32_t *>(rs.cb + hc);
        const bool rows16 = m.even_x && hv == 0;
        for (int step = 0; step < a.n_steps; step++) {
            const bool last = step == a.n_steps - 1;
            ens_advect_velocity<R, U>(m, a, A, B);
            env.sync();
#ifdef __CUDACC__
            // the previous grid's dye must have left the other dye buffer before anything is written there (this step's
            // dye advect, or the bulk load below); thread 0 reaches the barriers of the projection only after this wait
            if (async && step == 0 && tid == 0) ens_bulk_wait_read<0>();
#endif
            if (last && has_next) {         // the next grid's dye -> the dye buffer this step does not use
#ifdef __CUDACC__
                if (async) {
                    if (tid == 0) bulk_in(rs.cb, gc_next, c_bytes, rs.c_end, bar_c);
                } else
#endif
                    ens_copy_cong_in(rs.cb + head(gc_next), gc_next, c_bytes, rs.c_end, tid, NT);
            }

            ens_drags(m, a, B, step, grid, env);
            ens_project(m, a, reinterpret_cast<float *>(A), B, rows16, env);
            if (!last) {
                env.sync();
                ens_advect_dye<R, U>(m, a, B, SmemFetch<RgbPayload>{C_cur, m.dim_x}, C_oth);
                env.sync();
                // pointer swaps of ino:255 and ino:286 (buffers and data pointers alike)
                uint32_t *tc = C_cur; C_cur = C_oth; C_oth = tc;
                float2 *tv = A; A = B; B = tv;
                unsigned char *tb = rs.ca; rs.ca = rs.cb; rs.cb = tb;
                tb = rs.va; rs.va = rs.vb; rs.vb = tb;
                continue;
            }
#ifdef __CUDACC__
            if (async) ens_fence_proxy_async();         // B (and the mailboxes in A) before the bulk copies below
#endif
            env.sync();
            // the projected velocity is final: out it goes; A's buffer (advect source, then mailboxes) is dead: in comes the
            // next grid's velocity — both under the dye advect.  (No buffer swap: the next grid's velocity is in `va`.)
#ifdef __CUDACC__
            if (async) {
                if (tid == 0) {
                    bulk_out(gv, B, v_bytes);
                    if (has_next) bulk_in(rs.va, gv_next, v_bytes, rs.v_end, bar_v);
                }
            } else
#endif
            {
                ens_copy_cong_out(gv, B, v_bytes, tid, NT);
                if (has_next) ens_copy_cong_in(rs.va + head(gv_next), gv_next, v_bytes, rs.v_end, tid, NT);
            }
            ens_advect_dye_in_place<R, U>(m, a, B, C_cur, env);
#ifdef __CUDACC__
            if (async) ens_fence_proxy_async();         // C_cur before its bulk store
#endif
            env.sync();
        }
        // ---- the grid's dye leaves; wait for the next grid's state ------------------------------------------------
#ifdef __CUDACC__
        if (async) {
            if (tid == 0) {
                bulk_out(gc, C_cur, c_bytes);
                ens_bulk_wait_read<1>();    // the velocity store (the older group) no longer reads B: the next advect may write it
            }
            if (has_next) {
                mbar_wait(bar_v, phase_v); phase_v ^= 1;
                mbar_wait(bar_c, phase_c); phase_c ^= 1;
            }
        } else
#endif
            ens_copy_cong_out(gc, C_cur, c_bytes, tid, NT);
        unsigned char *tb = rs.ca; rs.ca = rs.cb; rs.cb = tb;      // the next grid's dye is in the other buffer
        env.sync();
    }
#ifdef __CUDACC__
    if (async && tid == 0) ens_bulk_wait_all();     // shared memory must outlive the last bulk stores
#endif
}

// PIPE: compile the pipelined flow too (it keeps up to 4R dye results per thread in registers: R = 2 only)
template <int R, bool PIPE, class Env>
__device__ __forceinline__ void ens_reg_body_resident(const EnsArgs &a, unsigned char *smem_raw, const Env &env)
{
    EnsMap<R> m;
    ens_map_init(m, a, env.tid, env.nthreads);
    const size_t vbuf = ens_reg_vbuf_bytes(m.dim_x, m.dim_y, R), cbuf = ens_reg_cbuf_bytes(m.dim_x, m.dim_y);
    EnsResident rs;
    rs.va = smem_raw;
    rs.vb = smem_raw + vbuf;
    rs.ca = smem_raw + 2 * vbuf;
    rs.cb = rs.ca + cbuf;
    rs.v_end = reinterpret_cast<const unsigned char *>(a.v + (size_t)a.batch * m.N);
    rs.c_end = reinterpret_cast<const unsigned char *>(a.c + (size_t)a.batch * m.N * 3);
    rs.cong = ((reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.c)) & 15) == 0;
    if constexpr (PIPE) {
        // (a call of many steps amortises the state I/O anyway, and the plain flow's steps are a few % faster)
        const bool async = Env::kAsync && rs.cong && a.n_steps <= a.pipe_max_steps;
        if (async || (Env::kEmulatePipe && rs.cong)) {
            ens_resident_pipelined<R>(m, a, rs, rs.cb + cbuf, async, env);
            return;
        }
    }
    ens_resident_sync<R>(m, a, rs, env);
}

// Dye STREAMED through L1/L2 (16 B/node of shared memory: grids too large for 40 B/node): the dye advect gathers it
// from global memory and writes the result back; between the steps of one call it ping-pongs between the caller's
// array and a per-CTA scratch slot that stays L2-resident.
template <int R, class Env>
__device__ __forceinline__ void ens_reg_body_streamed(const EnsArgs &a, unsigned char *smem_raw, const Env &env)
{
    constexpr int U = R == 2 ? 2 : 4;
    EnsMap<R> m;
    ens_map_init(m, a, env.tid, env.nthreads);
    const int tid = m.tid, NT = m.NT, N = m.N;
    const size_t vbuf = ens_reg_vbuf_bytes(m.dim_x, m.dim_y, R);
    float2 *A = reinterpret_cast<float2 *>(smem_raw);
    float2 *B = reinterpret_cast<float2 *>(smem_raw + vbuf);
    const bool vec16 = (N & 1) == 0 && (reinterpret_cast<uintptr_t>(a.v) & 15) == 0;
    uint32_t *scratch = a.scratch + (size_t)env.block * N * 3;

    for (int grid = env.block; grid < a.batch; grid += env.nblocks) {
        uint32_t *user_c = a.c + (size_t)grid * N * 3;
        uint32_t *C1 = user_c, *C2 = scratch;
        ens_copy_in(A, a.v + (size_t)grid * N, 2 * N, vec16, true, tid, NT);
        env.sync();
        for (int step = 0; step < a.n_steps; step++) {
            if (step == a.n_steps - 1 && grid + env.nblocks < a.batch) {
                const char *nv = reinterpret_cast<const char *>(a.v + (size_t)(grid + env.nblocks) * N);
                for (int l = tid * 128; l < N * 8; l += NT * 128) prefetch_l2(nv + l);
            }
            ens_advect_velocity<R, U>(m, a, A, B);
            env.sync();
            ens_drags(m, a, B, step, grid, env);
            ens_project(m, a, reinterpret_cast<float *>(A), B, m.even_x, env);
            env.sync();
            ens_advect_dye<R, U>(m, a, B, DyeFetch{C1, m.dim_x}, C2);
            env.sync();                     // (CTA-scope ordering of the dye stores before the next step's reads)
            // pointer swaps of ino:255 and ino:286
            uint32_t *tc = C1; C1 = C2; C2 = tc;
            float2 *tv = A; A = B; B = tv;
        }
        ens_copy_out(a.v + (size_t)grid * N, A, 2 * N, vec16, true, tid, NT);
        if (C1 != user_c)                   // the final dye sits in the scratch slot
            for (int n = tid; n < 3 * N; n += NT) user_c[n] = C1[n];
        env.sync();
    }
}

template <int R, bool DYE_SMEM, class Env, bool PIPE = (R == 2)>
__device__ __forceinline__ void ens_reg_body(const EnsArgs &a, unsigned char *smem_raw, const Env &env)
{
    if constexpr (DYE_SMEM)
        ens_reg_body_resident<R, PIPE>(a, smem_raw, env);
    else
        ens_reg_body_streamed<R>(a, smem_raw, env);
}

}  // namespace fs
