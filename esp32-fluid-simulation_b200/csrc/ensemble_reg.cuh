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
// The two advects keep a node-strided mapping over ALL threads of the CTA, four nodes per thread at a time:
// the interior case of sample<T> (advect.h:38-42) is evaluated unconditionally on a clamped cell so that
// the 16 corner loads of four nodes are in flight together, and only nodes whose backtrace left the grid
// are redone through the general sample().
//
// The body is a template over an execution environment (thread id, CTA size, barrier) so that
// tests/emu/ compiles THIS source for the host and runs it on CPU threads against the oracle.
#pragma once

#include "ensemble_common.cuh"

namespace fs {

template <int R>
struct EnsRegLayout {
    static_assert(R >= 2 && R % 2 == 0, "R must be even: the colour pattern of a block must not depend on the thread");
    static constexpr int MAIL_WORDS = 2 * R + 8;         // H(r, side) = 2r + side, V(side, c) = 2R + 4 side + c
    static constexpr int MAIL_STRIDE = MAIL_WORDS | 1;   // odd => lanes hit distinct banks
};

// threads that own a block of the grid
__host__ __device__ static inline int ens_reg_blocks(int dim_x, int dim_y, int R) { return ((dim_x + 3) / 4) * ((dim_y + R - 1) / R); }
// one velocity buffer: N nodes + 4 of padding (a block's clamped loads may run 3 nodes past the end), or
// the mailboxes of all block threads + the block of zeros, whichever is larger
__host__ __device__ static inline size_t ens_reg_vbuf_bytes(int dim_x, int dim_y, int R)
{
    const size_t v = 8 * ((size_t)dim_x * dim_y + 4);
    const size_t m = 4 * (size_t)((2 * R + 8) | 1) * ((size_t)ens_reg_blocks(dim_x, dim_y, R) + 1);
    return ((v > m ? v : m) + 15) & ~(size_t)15;
}
__host__ __device__ static inline size_t ens_reg_smem_bytes(int dim_x, int dim_y, int R, bool dye_smem)
{
    return 2 * ens_reg_vbuf_bytes(dim_x, dim_y, R) + (dye_smem ? (size_t)24 * dim_x * dim_y : 0);
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
// compare per use instead of one 3-input logic op)
#define FS_OPAQUE_REG(x) asm volatile("" : "+r"(x))

// Env: { int tid, nthreads, block, nblocks; void sync() const; }
template <int R, bool DYE_SMEM, class Env>
__device__ __forceinline__ void ens_reg_body(const EnsArgs &a, unsigned char *smem_raw, const Env &env)
{
    constexpr int MS = EnsRegLayout<R>::MAIL_STRIDE;
    constexpr int U = R == 2 ? 2 : 4;                      // nodes per thread in flight in the advects (R = 2: 80 registers)
    const int tid = env.tid, NT = env.nthreads;
    const int dim_x = a.dim_x, dim_y = a.dim_y, N = dim_x * dim_y;
    const int CG = (dim_x + 3) >> 2, RS = (dim_y + R - 1) / R, NS = CG * RS;
    const size_t vbuf = ens_reg_vbuf_bytes(dim_x, dim_y, R);
    float2 *A = reinterpret_cast<float2 *>(smem_raw);
    float2 *B = reinterpret_cast<float2 *>(smem_raw + vbuf);
    uint32_t *C1 = DYE_SMEM ? reinterpret_cast<uint32_t *>(smem_raw + 2 * vbuf) : nullptr;
    uint32_t *C2 = DYE_SMEM ? C1 + 3 * (size_t)N : a.scratch + (size_t)env.block * N * 3;

    const bool ragged = (dim_x & 3) != 0 || dim_y % R != 0;   // some blocks reach beyond the grid
    const bool even_x = (dim_x & 1) == 0;                     // block rows start 16-byte aligned
    // ---- this thread's block: columns i0..i0+3, rows j0..j0+R-1 ------------------------------------
    const bool act = tid < NS;
    const int s = act ? tid / CG : 0, g = act ? tid - s * CG : 0;
    const int i0 = 4 * g, j0 = R * s;
    unsigned cm[4], rm[R];
    float coef[R][4];
    int row[R];                                            // node index of (i0, j0 + r), row clamped into the grid
#pragma unroll
    for (int c = 0; c < 4; c++) {
        cm[c] = i0 + c < dim_x ? 0xffffffffu : 0u;
        FS_OPAQUE_REG(cm[c]);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        rm[r] = j0 + r < dim_y ? 0xffffffffu : 0u;
        FS_OPAQUE_REG(rm[r]);
        row[r] = min(j0 + r, dim_y - 1) * dim_x + i0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = i0 + c, j = j0 + r;
            const int nb = 4 - (i == 0) - (i == dim_x - 1) - (j == 0) - (j == dim_y - 1);
            coef[r][c] = nb == 4 ? a.k.neg_quarter : nb == 3 ? a.k.neg_third : a.k.neg_half;
        }
    }
    // mailboxes (word offsets into the dead velocity buffer); a missing neighbour is the block of zeros
    const int o_zero = NS * MS;
    const int o_mine = tid * MS;
    int o_left = g > 0 ? o_mine - MS : o_zero, o_right = g < CG - 1 ? o_mine + MS : o_zero;
    int o_down = s > 0 ? o_mine - CG * MS : o_zero, o_up = s < RS - 1 ? o_mine + CG * MS : o_zero;
    // (kept in registers: recomputing them cost ~25 integer instructions in every half-sweep)
    FS_OPAQUE_REG(o_left); FS_OPAQUE_REG(o_right); FS_OPAQUE_REG(o_down); FS_OPAQUE_REG(o_up);
    // neighbours of the block in the velocity array, clamped so that every address is inside the buffer
    const int row_dn = s > 0 ? row[0] - dim_x : row[0];
    const int row_up = min(j0 + R, dim_y - 1) * dim_x + i0;
    const int off_l = g > 0 ? -1 : 0;

    // 16-byte copies between global and shared memory need N % 4 == 0 (12N and 8N multiples of 16) and aligned arrays
    const bool vec16 = (N & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.v) | reinterpret_cast<uintptr_t>(a.c)) & 15) == 0;
    // ---- node-strided mapping of the advects ----------------------------------------------------------
    // (coordinates kept as floats: exact for these sizes, and the backtrace wants them as floats)
    const float adv_di = (float)(NT % dim_x), adv_dj = (float)(NT / dim_x), adv_i0 = (float)(tid % dim_x), adv_j0 = (float)(tid / dim_x);
    const float fdim_x = (float)dim_x, gx2 = (float)(dim_x - 2), gy2 = (float)(dim_y - 2);

    for (int grid = env.block; grid < a.batch; grid += env.nblocks) {
        // ---- load the grid's state ---------------------------------------------------------------------
        const float2 *gv = a.v + (size_t)grid * N;
        uint32_t *user_c = a.c + (size_t)grid * N * 3;
        if (vec16) {
            // 16 bytes per load: four times the bytes in flight per thread (the phase is latency-bound)
            const uint4 *src = reinterpret_cast<const uint4 *>(gv);
            uint4 *dst = reinterpret_cast<uint4 *>(A);
            for (int n = tid; n < N / 2; n += NT) dst[n] = __ldg(src + n);
        } else {
            for (int n = tid; n < N; n += NT) A[n] = __ldg(gv + n);
        }
        if constexpr (DYE_SMEM) {
            if (vec16) {
                const uint4 *src = reinterpret_cast<const uint4 *>(user_c);
                uint4 *dst = reinterpret_cast<uint4 *>(C1);
                for (int n = tid; n < 3 * N / 4; n += NT) dst[n] = __ldg(src + n);
            } else {
                for (int n = tid; n < 3 * N; n += NT) C1[n] = __ldg(user_c + n);
            }
        } else {
            C1 = user_c;
        }
        env.sync();

        for (int step = 0; step < a.n_steps; step++) {
            // during the last step, pull the next grid's state into L2: its load phase then pays the L2
            // latency instead of HBM's
            if (step == a.n_steps - 1 && grid + env.nblocks < a.batch) {
                const char *nv = reinterpret_cast<const char *>(a.v + (size_t)(grid + env.nblocks) * N);
                const char *nc = reinterpret_cast<const char *>(a.c + (size_t)(grid + env.nblocks) * N * 3);
                for (int l = tid * 128; l < N * 8; l += NT * 128) prefetch_l2(nv + l);
                if constexpr (DYE_SMEM)
                    for (int l = tid * 128; l < N * 12; l += NT * 128) prefetch_l2(nc + l);
            }
            // ---- advect velocity, no-slip (ino:253): A -> B -----------------------------------------------
            {
                SmemFetch<Vec2Payload> fetch{reinterpret_cast<const float *>(A), dim_x};
                float fi = adv_i0, fj = adv_j0;             // node coordinates as floats (exact), advect.h:81
                for (int nb = tid; nb < N; nb += U * NT) {
                    float res[U][2];
                    float si[U], sj[U];
                    bool oob[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int n = min(nb + u * NT, N - 1);      // (clamped: the surplus lanes recompute the last node)
                        const float2 vel = A[n];
                        si[u] = __fsub_rn(fi, __fmul_rn(vel.x, a.dt));
                        sj[u] = __fsub_rn(fj, __fmul_rn(vel.y, a.dt));
                        oob[u] = sample_interior<Vec2Payload>(res[u], fetch, si[u], sj[u], gx2, gy2);
                        fi += adv_di;
                        fj += adv_dj;
                        if (fi >= fdim_x) { fi -= fdim_x; fj += 1.0f; }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (nb + u * NT < N) {
                            if (oob[u]) sample<Vec2Payload>(res[u], fetch, si[u], sj[u], dim_x, dim_y, true);
                            B[nb + u * NT] = make_float2(res[u][0], res[u][1]);
                        }
                    }
                }
            }
            env.sync();
            // ---- drags (ino:264-269): in order, one thread -----------------------------------------------
            if (a.max_drags > 0) {
                if (tid == 0) {
                    const size_t slot = (size_t)step * a.batch + grid;
                    const int cnt = min(a.counts[slot], a.max_drags);
                    const fs_drag *dr = a.drags + slot * a.max_drags;
                    for (int q = 0; q < cnt; q++) {
                        const fs_drag m = dr[q];
                        if (m.cy < dim_x && m.cx < dim_y) B[m.cx * dim_x + m.cy] = make_float2(m.vy, m.vx);
                    }
                }
                env.sync();
            }
            // ---- projection in registers; the mailboxes live in A, which is dead now -----------------------
            float *mail = reinterpret_cast<float *>(A);
            float p[R][4], d[R][4];
            if (tid < MS) mail[o_zero + tid] = 0.0f;       // (first read after the first half-sweep's barrier)
            if (act) {
                // divergence (ino:274, finitediff.cpp:9-39) of the block from B and its one-node ring
                float vx[R][4], vy[R][4], xl[R], xr[R], yd[4], yu[4];
                if (even_x) {
                    // dim_x even: every row of the block starts 16-byte aligned, two nodes per load
#pragma unroll
                    for (int r = 0; r < R; r++) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const float4 t = *reinterpret_cast<const float4 *>(B + row[r] + 2 * h);
                            vx[r][2 * h] = t.x; vy[r][2 * h] = t.y; vx[r][2 * h + 1] = t.z; vy[r][2 * h + 1] = t.w;
                        }
                    }
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const float4 td = *reinterpret_cast<const float4 *>(B + row_dn + 2 * h);
                        const float4 tu = *reinterpret_cast<const float4 *>(B + row_up + 2 * h);
                        yd[2 * h] = td.y; yd[2 * h + 1] = td.w; yu[2 * h] = tu.y; yu[2 * h + 1] = tu.w;
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < R; r++) {
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const float2 t = B[row[r] + c];
                            vx[r][c] = t.x;
                            vy[r][c] = t.y;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        yd[c] = B[row_dn + c].y;
                        yu[c] = B[row_up + c].y;
                    }
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    xl[r] = B[row[r] + off_l].x;
                    xr[r] = B[row[r] + 4].x;
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
                        d[r][c] = __uint_as_float(__float_as_uint(dd) & cm[c] & rm[r]);
                        p[r][c] = 0.0f;                    // poisson.cpp:117-119
                    }
                }
            }
            // ---- red-black SOR (ino:275) ---------------------------------------------------------------------
            if (a.iters > 0) {
                if (ragged)
                    ens_sor<R, true>(p, d, coef, cm, rm, mail + o_mine, mail + o_left, mail + o_right, mail + o_down, mail + o_up, a.k, a.iters, act, env);
                else
                    ens_sor<R, false>(p, d, coef, cm, rm, mail + o_mine, mail + o_left, mail + o_right, mail + o_down, mail + o_up, a.k, a.iters, act, env);
            } else {
                if (act)
                    for (int w = 0; w < MS; w++) mail[o_mine + w] = 0.0f;
                env.sync();
            }
            // ---- subtract gradient (ino:276, finitediff.cpp:41-82), in place on B ----------------------------
            if (act) {
                float hl[R], hr[R], vd[4], vu[4];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    hl[r] = mail[o_left + 2 * r + 1];
                    hr[r] = mail[o_right + 2 * r + 0];
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    vd[c] = mail[o_down + 2 * R + 4 + c];
                    vu[c] = mail[o_up + 2 * R + c];
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
                    if (even_x) {
                        // dim_x even: nodes are inside the grid in pairs, 16 bytes per access
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            if (cm[2 * h] & rm[r]) {
                                float4 t = *reinterpret_cast<const float4 *>(B + row[r] + 2 * h);
                                t.x = __fsub_rn(t.x, gx[2 * h]);
                                t.y = __fsub_rn(t.y, gy[2 * h]);
                                t.z = __fsub_rn(t.z, gx[2 * h + 1]);
                                t.w = __fsub_rn(t.w, gy[2 * h + 1]);
                                *reinterpret_cast<float4 *>(B + row[r] + 2 * h) = t;
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            if (cm[c] & rm[r]) {
                                float2 c0 = B[row[r] + c];
                                c0.x = __fsub_rn(c0.x, gx[c]);
                                c0.y = __fsub_rn(c0.y, gy[c]);
                                B[row[r] + c] = c0;
                            }
                        }
                    }
                }
            }
            env.sync();
            // ---- advect dye, free-slip sampling (ino:282) with the projected velocity: C1 -> C2 ---------------
            {
                DyeFetch fetch{C1, dim_x};
                float fi = adv_i0, fj = adv_j0;
                for (int nb = tid; nb < N; nb += U * NT) {
                    uint32_t res[U][3];
                    float si[U], sj[U];
                    bool oob[U];
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int n = min(nb + u * NT, N - 1);
                        const float2 vel = B[n];
                        si[u] = __fsub_rn(fi, __fmul_rn(vel.x, a.dt));
                        sj[u] = __fsub_rn(fj, __fmul_rn(vel.y, a.dt));
                        oob[u] = sample_interior<RgbPayload>(res[u], fetch, si[u], sj[u], gx2, gy2);
                        fi += adv_di;
                        fj += adv_dj;
                        if (fi >= fdim_x) { fi -= fdim_x; fj += 1.0f; }
                    }
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        const int n = nb + u * NT;
                        if (n < N) {
                            if (oob[u]) sample<RgbPayload>(res[u], fetch, si[u], sj[u], dim_x, dim_y, false);
                            C2[3 * n + 0] = res[u][0];
                            C2[3 * n + 1] = res[u][1];
                            C2[3 * n + 2] = res[u][2];
                        }
                    }
                }
            }
            env.sync();                     // (CTA-scope ordering of the dye stores before the next step's reads)
            // pointer swaps of ino:255 and ino:286
            uint32_t *tc = C1; C1 = C2; C2 = tc;
            float2 *tv = A; A = B; B = tv;
        }

        // ---- store the grid's state ----------------------------------------------------------------------
        float2 *ov = a.v + (size_t)grid * N;
        if (vec16) {
            const uint4 *src = reinterpret_cast<const uint4 *>(A);
            uint4 *dst = reinterpret_cast<uint4 *>(ov);
            for (int n = tid; n < N / 2; n += NT) dst[n] = src[n];
        } else {
            for (int n = tid; n < N; n += NT) ov[n] = A[n];
        }
        if (C1 != user_c) {                 // the final dye sits in shared memory / in the scratch slot
            if (DYE_SMEM && vec16) {
                const uint4 *src = reinterpret_cast<const uint4 *>(C1);
                uint4 *dst = reinterpret_cast<uint4 *>(user_c);
                for (int n = tid; n < 3 * N / 4; n += NT) dst[n] = src[n];
            } else {
                for (int n = tid; n < 3 * N; n += NT) user_c[n] = C1[n];
            }
        }
        env.sync();
        if constexpr (!DYE_SMEM) C2 = a.scratch + (size_t)env.block * N * 3;   // next grid: C1 = its own array again
    }
}

}  // namespace fs
