// Advection kernels (advect.h:74-85): one thread per node, coalesced velocity
// read, backtrace, bilinear sample, coalesced write.
//
//   advect_gather<P> : the four corner reads go straight to L1/L2.  Correct for
//                      any displacement and any grid shape; used for odd row
//                      pitches (TMA needs 16-byte row strides) and as the
//                      fallback of the tiled kernel.
#include "advect.cuh"
#include "kernels.h"

namespace fs {

constexpr int ADV_BX = 64, ADV_BY = 4;

template <class P>
__global__ void __launch_bounds__(ADV_BX *ADV_BY)
advect_gather(typename P::raw_t *__restrict__ next_p, const typename P::raw_t *__restrict__ p,
              const float2 *__restrict__ vel, Geo g, float dt, bool no_slip, int *status)
{
    const int lx = g.x0 + blockIdx.x * ADV_BX + threadIdx.x;
    const int ly = g.y0 + blockIdx.y * ADV_BY + threadIdx.y;
    if (lx >= g.x1 || ly >= g.y1) return;
    const size_t l = (size_t)ly * g.nx + lx;
    const int gi = g.ox + lx, gj = g.oy + ly;

    float si, sj;
    backtrace(si, sj, gi, gj, __ldg(vel + l), dt);

    GlobalFetch<P> fetch{p, g.ox, g.oy, g.nx, g.vx0, g.vy0, g.vx1 - g.vx0, g.vy1 - g.vy0, status};
    typename P::raw_t out[P::NC];
    sample<P>(out, fetch, si, sj, g.GX, g.GY, no_slip);

    if constexpr (P::NC == 2) {
        reinterpret_cast<float2 *>(next_p)[l] = make_float2(out[0], out[1]);
    } else {
#pragma unroll
        for (int ch = 0; ch < P::NC; ch++) next_p[l * P::NC + ch] = out[ch];
    }
}

template <class P>
static int launch_gather(const Launch &L, typename P::raw_t *next_p, const typename P::raw_t *p,
                         const float2 *vel, const Geo &g, float dt, bool no_slip, int *status)
{
    const int w = g.x1 - g.x0, h = g.y1 - g.y0;
    if (w <= 0 || h <= 0) return 0;
    dim3 block(ADV_BX, ADV_BY), grid((w + ADV_BX - 1) / ADV_BX, (h + ADV_BY - 1) / ADV_BY);
    advect_gather<P><<<grid, block, 0, L.stream>>>(next_p, p, vel, g, dt, no_slip, status);
    ++*L.launches;
    return (int)cudaGetLastError();
}

int launch_advect_vec2f_gather(const Launch &L, float2 *next_p, const float2 *p, const float2 *vel,
                               const Geo &g, float dt, bool no_slip, int *status)
{
    return launch_gather<Vec2Payload>(L, reinterpret_cast<float *>(next_p),
                                      reinterpret_cast<const float *>(p), vel, g, dt, no_slip,
                                      status);
}

int launch_advect_rgb_gather(const Launch &L, uint32_t *next_c, const uint32_t *c, const float2 *vel,
                             const Geo &g, float dt, bool no_slip, int *status)
{
    return launch_gather<RgbPayload>(L, next_c, c, vel, g, dt, no_slip, status);
}

int preload_advect_kernels()
{
    FS_PRELOAD(advect_gather<Vec2Payload>);
    FS_PRELOAD(advect_gather<RgbPayload>);
    return 0;
}

}  // namespace fs
