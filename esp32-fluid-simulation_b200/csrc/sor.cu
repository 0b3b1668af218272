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

}  // namespace fs
