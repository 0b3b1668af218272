// Red-black SOR pressure solve (poisson.cpp:14-125).
//
//   sor_half_sweep_kernel : one colour per launch, in place in global memory.
//                           The simple, always-legal variant (any shape, any
//                           window); 2*iters launches per solve.
//
// Within one colour every update reads only the OTHER colour, so the visiting
// order inside a half-sweep is free: one thread per
// updated node is bit-identical to the reference's sequential walk as long as
// colour 0 ((i+j) even, the reference's on_red=false pass) goes first.
#include "kernels.h"
#include "sor.cuh"

namespace fs {

constexpr int HS_BX = 64, HS_BY = 4;

__global__ void __launch_bounds__(HS_BX *HS_BY)
sor_half_sweep_kernel(float *__restrict__ p, const float *__restrict__ div, Geo g, SorCoef k,
                      int parity)
{
    const int ly = g.y0 + blockIdx.y * HS_BY + threadIdx.y;
    if (ly >= g.y1) return;
    const int gj = g.oy + ly;
    // first node of this colour in the row: global (gi + gj) & 1 == parity
    const int lx = g.x0 + ((g.ox + g.x0 + gj + parity) & 1) + 2 * (blockIdx.x * HS_BX + threadIdx.x);
    if (lx >= g.x1) return;
    const int gi = g.ox + lx;
    const size_t l = (size_t)ly * g.nx + lx;
    const float pc = p[l], d = __ldg(&div[l]);
    float out;
    if (gi > 0 && gi < g.GX - 1 && gj > 0 && gj < g.GY - 1) {
        out = sor_update_interior(pc, p[l - 1], p[l + 1], p[l - g.nx], p[l + g.nx],
                                  __fmul_rn(k.dx, d), k);
    } else {
        const bool hl = gi > 0, hr = gi < g.GX - 1, hd = gj > 0, hu = gj < g.GY - 1;
        out = sor_update_wall(pc, hl ? p[l - 1] : 0.0f, hr ? p[l + 1] : 0.0f,
                              hd ? p[l - g.nx] : 0.0f, hu ? p[l + g.nx] : 0.0f, hl, hr, hd, hu,
                              __fmul_rn(k.dx, d), k);
    }
    p[l] = out;
}

// Residual of the system the sweeps relax: r_ij = gs_ij(p) - p_ij with gs the Gauss-Seidel value of
// poisson.cpp:63-90 / :101-109 (the SOR update moves p_ij by omega * r_ij).  Per-thread partials are
// reduced with warp shuffles; one atomicMax + one atomicAdd per warp.  out[0] = max |r| as float
// bits (non-negative floats order like unsigned ints), out64 = sum of r^2 in double.
__global__ void __launch_bounds__(256)
sor_residual_kernel(const float *__restrict__ p, const float *__restrict__ div, Geo g, SorCoef k,
                    unsigned int *out_max_bits, double *out_sumsq)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    const size_t n = (size_t)w * h;
    float m = 0.0f;
    double ss = 0.0;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int lx = g.x0 + (int)(q % w), ly = g.y0 + (int)(q / w);
        const int gi = g.ox + lx, gj = g.oy + ly;
        const size_t l = (size_t)ly * g.nx + lx;
        const bool hl = gi > 0, hr = gi < g.GX - 1, hd = gj > 0, hu = gj < g.GY - 1;
        const float dxd = __fmul_rn(k.dx, __ldg(&div[l]));
        float gs;
        if (hl && hr && hd && hu) {
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(p[l - 1], p[l + 1]), p[l - g.nx]), p[l + g.nx]);
            gs = __fmul_rn(-0.25f, __fsub_rn(dxd, sum));
        } else {
            float sum = 0.0f;
            int a = 0;
            if (hl) { sum = __fadd_rn(sum, p[l - 1]); a++; }
            if (hr) { sum = __fadd_rn(sum, p[l + 1]); a++; }
            if (hd) { sum = __fadd_rn(sum, p[l - g.nx]); a++; }
            if (hu) { sum = __fadd_rn(sum, p[l + g.nx]); a++; }
            const float coef = a == 4 ? k.neg_quarter : a == 3 ? k.neg_third : a == 2 ? k.neg_half : 0.0f;
            gs = __fmul_rn(coef, __fsub_rn(dxd, sum));
        }
        const float r = __fsub_rn(gs, p[l]);
        m = fmaxf(m, fabsf(r));
        ss += (double)r * (double)r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out_max_bits, __float_as_uint(m));
        atomicAdd(out_sumsq, ss);
    }
}

int launch_sor_residual(const Launch &L, const float *p, const float *div, const Geo &g, float dx,
                        unsigned int *out_max_bits, double *out_sumsq)
{
    cudaError_t e = cudaMemsetAsync(out_max_bits, 0, sizeof(unsigned int), L.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(out_sumsq, 0, sizeof(double), L.stream);
    if (e != cudaSuccess) return (int)e;
    if (g.x1 <= g.x0 || g.y1 <= g.y0) return 0;
    sor_residual_kernel<<<L.num_sms * 8, 256, 0, L.stream>>>(p, div, g, make_sor_coef(dx, 1.0f), out_max_bits,
                                                              out_sumsq);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int launch_sor_half_sweep(const Launch &L, float *p, const float *div, const Geo &g, float dx,
                          float omega, int parity)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    if (w <= 0 || h <= 0) return 0;
    const int per_row = (w + 1) / 2;
    dim3 block(HS_BX, HS_BY), grid((per_row + HS_BX - 1) / HS_BX, (h + HS_BY - 1) / HS_BY);
    sor_half_sweep_kernel<<<grid, block, 0, L.stream>>>(p, div, g, make_sor_coef(dx, omega),
                                                        parity & 1);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int preload_sor_kernels()
{
    FS_PRELOAD(sor_half_sweep_kernel);
    FS_PRELOAD(sor_residual_kernel);
    return 0;
}

}  // namespace fs
