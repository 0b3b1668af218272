// Finite-difference stencils with the wall handling fused in
// (finitediff.cpp:9-82), the drag overwrite (ino:264-269) and the max-
// displacement reduction that sizes advect halos on decomposed grids.
//
// domain_iter (operations.h:11-38) splits a grid into an interior that gets
// `expr_fast` and four edges that get `expr_safe`.  The GPU analogue is a
// per-thread wall test on GLOBAL coordinates: interior threads take the fast
// expression, wall threads the safe one — the two associate the sum
// differently, and both orders are kept.
#include "kernels.h"

namespace fs {

constexpr int ST_BX = 64, ST_BY = 4;

// one node of calculate_divergence, finitediff.cpp:9-31
__device__ __forceinline__ float div_node(const float2 *__restrict__ v, const Geo &g, int lx, int ly,
                                          float two_dx_inv)
{
    const size_t l = (size_t)ly * g.nx + lx;
    const int gi = g.ox + lx, gj = g.oy + ly;
    const int i_max = g.GX - 1, j_max = g.GY - 1;
    const bool wall = gi == 0 || gi == i_max || gj == 0 || gj == j_max;
    float s;
    if (!wall) {  // div_expr_fast, finitediff.cpp:29: (-L.x + R.x) + (-D.y + U.y)
        const float lxv = __ldg(&v[l - 1].x), rxv = __ldg(&v[l + 1].x);
        const float dyv = __ldg(&v[l - g.nx].y), uyv = __ldg(&v[l + g.nx].y);
        s = __fadd_rn(__fadd_rn(-lxv, rxv), __fadd_rn(-dyv, uyv));
    } else {      // div_expr_safe, finitediff.cpp:16-20: ghost velocity is the negated wall node
        const float2 c = __ldg(&v[l]);
        s = 0.0f;
        s = __fadd_rn(s, gi > 0 ? -__ldg(&v[l - 1].x) : c.x);
        s = __fadd_rn(s, gi < i_max ? __ldg(&v[l + 1].x) : -c.x);
        s = __fadd_rn(s, gj > 0 ? -__ldg(&v[l - g.nx].y) : c.y);
        s = __fadd_rn(s, gj < j_max ? __ldg(&v[l + g.nx].y) : -c.y);
    }
    return __fmul_rn(s, two_dx_inv);
}

// calculate_divergence, finitediff.cpp:33-39 — one node per thread (any pitch / alignment)
__global__ void __launch_bounds__(ST_BX *ST_BY)
divergence_kernel(float *__restrict__ div, const float2 *__restrict__ v, Geo g, float two_dx_inv)
{
    const int lx = g.x0 + blockIdx.x * ST_BX + threadIdx.x;
    const int ly = g.y0 + blockIdx.y * ST_BY + threadIdx.y;
    if (lx >= g.x1 || ly >= g.y1) return;
    div[(size_t)ly * g.nx + lx] = div_node(v, g, lx, ly, two_dx_inv);
}

// Four consecutive nodes per thread: 6 independent 16-byte loads in flight per thread (the one-node
// kernel was latency-bound with ~12 KB in flight per SM, ncu long_scoreboard 19.8 — profiles/
// r01_ncu_stencils_v1.json), the two horizontal neighbours that belong to the adjacent lanes come
// by warp shuffle, the result leaves as one 16-byte store.  Needs nx % 4 == 0 and 16-byte bases.
constexpr int X4_ROWS = 8;
// The divergence kernel walks X4_RPT consecutive rows per thread with a rolling three-row window: a row of
// velocities is loaded once and serves as the upper neighbour, the centre and the lower neighbour of three
// consecutive output rows ((RPT + 2) / RPT = 1.5 loads per output row instead of 3: the one-row version ran at
// 0.54 of the copy bandwidth against the gradient kernel's 0.99, its L1 traffic three times the algorithmic read).
constexpr int X4_RPT = 4;
__global__ void __launch_bounds__(32 * X4_ROWS)
divergence_x4_kernel(float *__restrict__ div, const float2 *__restrict__ v, Geo g, float two_dx_inv)
{
    const int lane = threadIdx.x;
    const int lx4 = (g.x0 & ~3) + 4 * (blockIdx.x * 32 + lane);
    const int ly0 = g.y0 + (blockIdx.y * X4_ROWS + threadIdx.y) * X4_RPT;
    if (ly0 >= g.y1) return;                                 // warp-uniform
    const int gi0 = g.ox + lx4;
    const bool can_load = lx4 + 3 < g.nx;
    // all four nodes inside the rectangle and strictly interior to the domain in x
    const bool fast_x = can_load && lx4 >= g.x0 && lx4 + 3 < g.x1 && gi0 > 0 && gi0 + 3 < g.GX - 1;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_row = [&](int y, float4 &a, float4 &b) {
        a = zero;
        b = zero;
        if (can_load && y >= 0 && y < g.ny) {
            const float2 *q = v + (size_t)y * g.nx + lx4;
            a = __ldg(reinterpret_cast<const float4 *>(q));
            b = __ldg(reinterpret_cast<const float4 *>(q + 2));
        }
    };
    // all RPT + 2 rows up front: 12 independent 16-byte loads in flight per thread
    float4 r01[X4_RPT + 2], r23[X4_RPT + 2];
#pragma unroll
    for (int r = 0; r < X4_RPT + 2; r++) load_row(ly0 - 1 + r, r01[r], r23[r]);
#pragma unroll
    for (int r = 0; r < X4_RPT; r++) {
        const int ly = ly0 + r;
        if (ly >= g.y1) break;                               // warp-uniform
        const float4 d01 = r01[r], d23 = r23[r], c01 = r01[r + 1], c23 = r23[r + 1], u01 = r01[r + 2], u23 = r23[r + 2];
        const int gj = g.oy + ly;
        const bool fast = fast_x && gj > 0 && gj < g.GY - 1;
        const size_t l = (size_t)ly * g.nx + lx4;
        float left_x = __shfl_up_sync(0xffffffffu, c23.z, 1);    // node 3 of the lane to the left
        float right_x = __shfl_down_sync(0xffffffffu, c01.x, 1); // node 0 of the lane to the right
        if (fast) {
            if (lane == 0) left_x = __ldg(&v[l - 1].x);
            if (lane == 31) right_x = __ldg(&v[l + 4].x);
            // div_expr_fast, finitediff.cpp:29: (-L.x + R.x) + (-D.y + U.y)
            float4 o;
            o.x = __fmul_rn(__fadd_rn(__fadd_rn(-left_x, c01.z), __fadd_rn(-d01.y, u01.y)), two_dx_inv);
            o.y = __fmul_rn(__fadd_rn(__fadd_rn(-c01.x, c23.x), __fadd_rn(-d01.w, u01.w)), two_dx_inv);
            o.z = __fmul_rn(__fadd_rn(__fadd_rn(-c01.z, c23.z), __fadd_rn(-d23.y, u23.y)), two_dx_inv);
            o.w = __fmul_rn(__fadd_rn(__fadd_rn(-c23.x, right_x), __fadd_rn(-d23.w, u23.w)), two_dx_inv);
            *reinterpret_cast<float4 *>(div + l) = o;
        } else {
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                const int lx = lx4 + c;
                if (lx >= g.x0 && lx < g.x1) div[(size_t)ly * g.nx + lx] = div_node(v, g, lx, ly, two_dx_inv);
            }
        }
    }
}

// one node of subtract_gradient, finitediff.cpp:41-73 (v_out may alias v_in: only v[ij] itself is read)
__device__ __forceinline__ void grad_node(float2 *v_out, const float2 *v_in, const float *__restrict__ p,
                                          const Geo &g, int lx, int ly, float two_dx_inv)
{
    const size_t l = (size_t)ly * g.nx + lx;
    const int gi = g.ox + lx, gj = g.oy + ly;
    const float pc = __ldg(&p[l]);
    // a missing neighbour is replaced by the node's own pressure (finitediff.cpp:51-54)
    const float pl = gi > 0 ? __ldg(&p[l - 1]) : pc;
    const float pr = gi < g.GX - 1 ? __ldg(&p[l + 1]) : pc;
    const float pd = gj > 0 ? __ldg(&p[l - g.nx]) : pc;
    const float pu = gj < g.GY - 1 ? __ldg(&p[l + g.nx]) : pc;
    const float gx = __fmul_rn(__fsub_rn(pr, pl), two_dx_inv);
    const float gy = __fmul_rn(__fsub_rn(pu, pd), two_dx_inv);
    float2 c = v_in[l];  // v_out may alias v_in (the reference runs in place, finitediff.cpp:80)
    c.x = __fsub_rn(c.x, gx);
    c.y = __fsub_rn(c.y, gy);
    v_out[l] = c;
}

// subtract_gradient, finitediff.cpp:75-82 — one node per thread (any pitch / alignment)
__global__ void __launch_bounds__(ST_BX *ST_BY)
subtract_gradient_kernel(float2 *v_out, const float2 *v_in, const float *__restrict__ p, Geo g,
                         float two_dx_inv)
{
    const int lx = g.x0 + blockIdx.x * ST_BX + threadIdx.x;
    const int ly = g.y0 + blockIdx.y * ST_BY + threadIdx.y;
    if (lx >= g.x1 || ly >= g.y1) return;
    grad_node(v_out, v_in, p, g, lx, ly, two_dx_inv);
}

// four consecutive nodes per thread (see divergence_x4_kernel)
__global__ void __launch_bounds__(32 * X4_ROWS)
subtract_gradient_x4_kernel(float2 *v_out, const float2 *v_in, const float *__restrict__ p, Geo g,
                            float two_dx_inv)
{
    const int lane = threadIdx.x;
    const int lx4 = (g.x0 & ~3) + 4 * (blockIdx.x * 32 + lane);
    const int ly = g.y0 + blockIdx.y * X4_ROWS + threadIdx.y;
    if (ly >= g.y1) return;                                  // warp-uniform
    const size_t l = (size_t)ly * g.nx + lx4;
    const int gi0 = g.ox + lx4, gj = g.oy + ly;
    const bool can_load = lx4 + 3 < g.nx;
    const bool fast = can_load && lx4 >= g.x0 && lx4 + 3 < g.x1 && gi0 > 0 && gi0 + 3 < g.GX - 1 && gj > 0 &&
                      gj < g.GY - 1;
    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f), pd = pc, pu = pc, v01 = pc, v23 = pc;
    if (can_load) pc = __ldg(reinterpret_cast<const float4 *>(p + l));
    if (fast) {
        pd = __ldg(reinterpret_cast<const float4 *>(p + l - g.nx));
        pu = __ldg(reinterpret_cast<const float4 *>(p + l + g.nx));
        v01 = *reinterpret_cast<const float4 *>(v_in + l);
        v23 = *reinterpret_cast<const float4 *>(v_in + l + 2);
    }
    float left = __shfl_up_sync(0xffffffffu, pc.w, 1), right = __shfl_down_sync(0xffffffffu, pc.x, 1);
    if (fast) {
        if (lane == 0) left = __ldg(p + l - 1);
        if (lane == 31) right = __ldg(p + l + 4);
        // grad_sub_expr_fast, finitediff.cpp:70-72
        v01.x = __fsub_rn(v01.x, __fmul_rn(__fsub_rn(pc.y, left), two_dx_inv));
        v01.y = __fsub_rn(v01.y, __fmul_rn(__fsub_rn(pu.x, pd.x), two_dx_inv));
        v01.z = __fsub_rn(v01.z, __fmul_rn(__fsub_rn(pc.z, pc.x), two_dx_inv));
        v01.w = __fsub_rn(v01.w, __fmul_rn(__fsub_rn(pu.y, pd.y), two_dx_inv));
        v23.x = __fsub_rn(v23.x, __fmul_rn(__fsub_rn(pc.w, pc.y), two_dx_inv));
        v23.y = __fsub_rn(v23.y, __fmul_rn(__fsub_rn(pu.z, pd.z), two_dx_inv));
        v23.z = __fsub_rn(v23.z, __fmul_rn(__fsub_rn(right, pc.z), two_dx_inv));
        v23.w = __fsub_rn(v23.w, __fmul_rn(__fsub_rn(pu.w, pd.w), two_dx_inv));
        *reinterpret_cast<float4 *>(v_out + l) = v01;
        *reinterpret_cast<float4 *>(v_out + l + 2) = v23;
    } else {
#pragma unroll 1
        for (int c = 0; c < 4; c++) {
            const int lx = lx4 + c;
            if (lx >= g.x0 && lx < g.x1) grad_node(v_out, v_in, p, g, lx, ly, two_dx_inv);
        }
    }
}

// Drag overwrite, ino:264-269.  The queue is drained IN ORDER and each record
// SETS one node, so a later record wins: one thread replays the list.  Records
// travel as kernel parameters (no staging copy; at most DRAG_CHUNK per launch).
constexpr int DRAG_CHUNK = 128;
struct DragChunk {
    fs_drag d[DRAG_CHUNK];
};

__global__ void apply_drags_kernel(float2 *__restrict__ v, DragChunk chunk, int n, Geo g)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int k = 0; k < n; k++) {
        const fs_drag m = chunk.d[k];
        // index(coords.y, coords.x, N_ROWS): coords.y runs along the fast axis
        const int gi = m.cy, gj = m.cx;
        if (gi >= g.GX || gj >= g.GY) continue;  // reference: out-of-bounds write
        const int lx = gi - g.ox, ly = gj - g.oy;
        if (lx < g.x0 || lx >= g.x1 || ly < g.y0 || ly >= g.y1) continue;  // another rank's node
        v[(size_t)ly * g.nx + lx] = make_float2(m.vy, m.vx);  // swapped, ino:267
    }
}

// max over the rectangle of max(|v.x|,|v.y|) as float bits (non-negative floats
// order like unsigned ints); NaN is ignored.
__global__ void __launch_bounds__(256)
max_displacement_kernel(unsigned int *__restrict__ out_bits, const float2 *__restrict__ vel, Geo g)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    const size_t n = (size_t)w * h;
    float m = 0.0f;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (size_t)gridDim.x * blockDim.x) {
        const int lx = g.x0 + (int)(k % w), ly = g.y0 + (int)(k / w);
        const float2 c = __ldg(&vel[(size_t)ly * g.nx + lx]);
        m = fmaxf(m, fmaxf(fabsf(c.x), fabsf(c.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));
}

static inline dim3 st_grid(const Geo &g)
{
    return dim3((g.x1 - g.x0 + ST_BX - 1) / ST_BX, (g.y1 - g.y0 + ST_BY - 1) / ST_BY);
}

static inline dim3 x4_grid(const Geo &g, int rows_per_thread = 1)
{
    const int cols = g.x1 - (g.x0 & ~3), rows_per_cta = X4_ROWS * rows_per_thread;
    return dim3((cols + 127) / 128, (g.y1 - g.y0 + rows_per_cta - 1) / rows_per_cta);
}

int launch_divergence(const Launch &L, float *div, const float2 *v, const Geo &g, float dx)
{
    if (g.x1 <= g.x0 || g.y1 <= g.y0) return 0;
    const float two_dx_inv = 1.0f / (2.0f * dx);  // finitediff.cpp:36, formed on the host in float
    if (g.nx % 4 == 0 && (uintptr_t)div % 16 == 0 && (uintptr_t)v % 16 == 0)
        divergence_x4_kernel<<<x4_grid(g, X4_RPT), dim3(32, X4_ROWS), 0, L.stream>>>(div, v, g, two_dx_inv);
    else
        divergence_kernel<<<st_grid(g), dim3(ST_BX, ST_BY), 0, L.stream>>>(div, v, g, two_dx_inv);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int launch_subtract_gradient(const Launch &L, float2 *v_out, const float2 *v_in, const float *p,
                             const Geo &g, float dx)
{
    if (g.x1 <= g.x0 || g.y1 <= g.y0) return 0;
    const float two_dx_inv = 1.0f / (2.0f * dx);  // finitediff.cpp:79
    if (g.nx % 4 == 0 && (uintptr_t)p % 16 == 0 && (uintptr_t)v_in % 16 == 0 && (uintptr_t)v_out % 16 == 0)
        subtract_gradient_x4_kernel<<<x4_grid(g), dim3(32, X4_ROWS), 0, L.stream>>>(v_out, v_in, p, g,
                                                                                   two_dx_inv);
    else
        subtract_gradient_kernel<<<st_grid(g), dim3(ST_BX, ST_BY), 0, L.stream>>>(v_out, v_in, p, g,
                                                                                  two_dx_inv);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int launch_apply_drags(const Launch &L, float2 *v, const fs_drag *drags_host, int n, const Geo &g)
{
    for (int base = 0; base < n; base += DRAG_CHUNK) {
        DragChunk chunk;
        const int m = n - base < DRAG_CHUNK ? n - base : DRAG_CHUNK;
        for (int k = 0; k < m; k++) chunk.d[k] = drags_host[base + k];
        apply_drags_kernel<<<1, 32, 0, L.stream>>>(v, chunk, m, g);
        ++*L.launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int launch_max_displacement(const Launch &L, unsigned int *out_bits, const float2 *vel, const Geo &g)
{
    cudaError_t e = cudaMemsetAsync(out_bits, 0, sizeof(unsigned int), L.stream);
    if (e != cudaSuccess) return (int)e;
    if (g.x1 <= g.x0 || g.y1 <= g.y0) return 0;
    max_displacement_kernel<<<L.num_sms * 4, 256, 0, L.stream>>>(out_bits, vel, g);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int preload_stencil_kernels()
{
    FS_PRELOAD(divergence_kernel);
    FS_PRELOAD(divergence_x4_kernel);
    FS_PRELOAD(subtract_gradient_kernel);
    FS_PRELOAD(subtract_gradient_x4_kernel);
    FS_PRELOAD(apply_drags_kernel);
    FS_PRELOAD(max_displacement_kernel);
    return 0;
}

}  // namespace fs
