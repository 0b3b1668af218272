// Pieces shared by the two ensemble kernels (ensemble.cu: node-pair mapping, p/d in shared
// memory; ensemble_reg.cuh: register-tiled projection).
#pragma once

#include "advect.cuh"
#include "sor.cuh"

namespace fs {

template <class P>
struct SmemFetch {
    const typename P::raw_t *base;
    int dim_x;
    __device__ __forceinline__ void operator()(int gi, int gj, typename P::raw_t (&o)[P::NC]) const
    {
        const typename P::raw_t *q = base + (gj * dim_x + gi) * P::NC;
        if constexpr (P::NC == 2) {
            // one 8-byte load, stated explicitly: with a run-time offset in `base` the compiler can no longer prove the
            // alignment and splits it into two 4-byte loads (ncu: the velocity advect's LDS count doubled)
            const float2 t = *reinterpret_cast<const float2 *>(q);
            o[0] = t.x;
            o[1] = t.y;
        } else {
#pragma unroll
            for (int ch = 0; ch < P::NC; ch++) o[ch] = q[ch];
        }
    }
};

// dye of one grid in global memory.  Plain (L1-cached) loads: every dye word is a corner of ~4
// backtraces, and going to L2 for each of them cost 10x the grid's bytes in L2 traffic (measured:
// the whole step 55 % slower).  The previous step's result was written by this very CTA — ordered
// by the __syncthreads between the steps (CTA scope) — so only the read-only (ld.global.nc) path
// must not be used.
struct DyeFetch {
    const uint32_t *base;
    int dim_x;
    __device__ __forceinline__ void operator()(int gi, int gj, uint32_t (&o)[3]) const
    {
        const uint32_t *q = base + (size_t)(gj * dim_x + gi) * 3;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) o[ch] = q[ch];
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

struct EnsArgs {
    float2 *v;
    uint32_t *c;
    uint32_t *scratch;      // [gridDim.x][3N]: per-CTA dye ping-pong slot
    const fs_drag *drags;   // device: [n_steps][batch][max_drags]
    const int *counts;      // device: [n_steps][batch]
    int max_drags, batch, dim_x, dim_y, iters, n_steps;
    int pipe_max_steps;     // register-tiled, dye-resident kernel: calls of up to this many steps take the pipelined flow
    float dt, two_dx_inv;
    SorCoef k;
};

}  // namespace fs
