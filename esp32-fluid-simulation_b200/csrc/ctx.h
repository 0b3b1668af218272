// Private to the library: the context behind the opaque fs_ctx handle and the small helpers the
// C-ABI translation units (api.cu, dist.cu) share.
#pragma once

#include <cstring>

#include "kernels.h"

using namespace fs;

constexpr int WORK_SLOTS = 64;
constexpr int E2E_MAX_BANDS = 16;
enum Scratch { S_VTMP = 0, S_CTMP, S_DIV, S_P, S_P2, S_HV, S_HV2, S_HC, S_HC2, S_HP, S_HD, S_HIMG, S_EDRAG, S_ECNT, S_ECTMP, S_SOLVE_FLAGS, S_COUNT };

struct fs_ctx {
    int device;
    cudaStream_t stream;
    int owns_stream;            // created by fs_ctx_create(..., FS_STREAM_NEW)
    int num_sms;
    uint64_t launches;
    void *scratch[S_COUNT];
    size_t scratch_bytes[S_COUNT];
    int *status_dev;            // device flag raised by tile advects (FS_ERR_HALO_OVERRUN)
    unsigned int *maxdisp_dev;  // max-displacement reduction cell
    double *resid_dev;            // residual reduction cells: [0] sum of squares, then max bits
    unsigned int *halo_done_dev;  // block counter of the halo-exchange kernel
    int *work_dev;              // WORK_SLOTS tile counters of the persistent SOR kernel (one per pass)
    int *rim_dev;               // WORK_SLOTS rim-tile counters of the fused SOR + halo-exchange pass
    cudaStream_t copy_in, copy_out;   // side streams of fsh_step: PCIe copies overlap the compute
    cudaEvent_t ev_start, ev_c_in, ev_v_done;
    cudaEvent_t ev_band_in[E2E_MAX_BANDS], ev_band_done[E2E_MAX_BANDS];   // fsh_step's dye bands: arrived / advected
    int *band_flag_dev;         // raised when a dye backtrace of fsh_step reached a band that had not arrived yet
    int opt_e2e_bands;          // row bands the dye of fsh_step travels in (1 = one piece)
    int stat_e2e_redos;         // fsh_step calls whose banded dye advect had to be redone in one piece
    size_t max_smem_optin;
    unsigned int solve_gen;     // generation stamp of the single-launch solve's completion flags
    int opt_sor_one_launch;
    int opt_halo_timeout_ms;
    int opt_ens;                // ensemble kernel variant (CTA sizing), see fs_ctx_set_option
    int opt_sor_grid_limit;     // cap on the persistent SOR grid (0 = one CTA per SM): lets several emulated
                                // ranks share one device without starving each other
    int opt_sor, opt_sor_t, opt_sor_shape, opt_advect, opt_fuse;
};

struct DeviceGuard {
    int prev;
    bool changed;
    explicit DeviceGuard(int dev) : prev(-1), changed(false)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
            cudaSetDevice(dev);
            changed = true;
        }
    }
    ~DeviceGuard()
    {
        if (changed) cudaSetDevice(prev);
    }
};

inline int ensure(fs_ctx *ctx, Scratch slot, size_t bytes, void **out)
{
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            // the old buffer may still be in use by enqueued work
            FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            FS_CUDA_TRY(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        FS_CUDA_TRY(cudaMalloc(&ctx->scratch[slot], bytes));
        ctx->scratch_bytes[slot] = bytes;
    }
    *out = ctx->scratch[slot];
    return FS_OK;
}

inline Launch mk(fs_ctx *ctx) { return Launch{ctx->stream, &ctx->launches, ctx->num_sms}; }

inline bool bad_dims(int dim_x, int dim_y)
{
    // the reference's domain_iter visits nodes twice when a dimension is 1
    // (operations.h:26-37); 2 is the smallest well-defined grid.  Indices are
    // 32-bit in the reference (operations.h:7-9).
    return dim_x < 2 || dim_y < 2 || (long long)dim_x * dim_y > 0x7fffffffLL;
}

inline bool bad_tile(const fs_tile *t, int need_ring)
{
    if (!t) return true;
    if (bad_dims(t->gdim_x, t->gdim_y) || t->nx < 1 || t->ny < 1) return true;
    if (t->x0 < 0 || t->y0 < 0 || t->x1 > t->nx || t->y1 > t->ny || t->x0 > t->x1 || t->y0 > t->y1)
        return true;
    // the window must lie inside the global grid
    if (t->ox < 0 || t->oy < 0 || t->ox + t->nx > t->gdim_x || t->oy + t->ny > t->gdim_y) return true;
    // every computed node needs `need_ring` neighbours inside the window unless
    // the global wall cuts them off
    if (need_ring > 0 && t->x1 > t->x0 && t->y1 > t->y0) {
        if (t->x0 - need_ring < 0 && t->ox + t->x0 - need_ring >= 0) return true;
        if (t->y0 - need_ring < 0 && t->oy + t->y0 - need_ring >= 0) return true;
        if (t->x1 + need_ring > t->nx && t->ox + t->x1 + need_ring <= t->gdim_x) return true;
        if (t->y1 + need_ring > t->ny && t->oy + t->y1 + need_ring <= t->gdim_y) return true;
    }
    return false;
}

// rectangle grown by r, clipped to the window (== clipped to the global grid
// when bad_tile(t, r) passed)
inline Geo grown(const Geo &g, int r)
{
    Geo o = g;
    o.x0 = g.x0 - r < 0 ? 0 : g.x0 - r;
    o.y0 = g.y0 - r < 0 ? 0 : g.y0 - r;
    o.x1 = g.x1 + r > g.nx ? g.nx : g.x1 + r;
    o.y1 = g.y1 + r > g.ny ? g.ny : g.y1 + r;
    return o;
}

// operator cores shared with dist.cu (defined in api.cu)
int core_advect_vec2f(fs_ctx *ctx, fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel, const Geo &g, float dt,
                      int no_slip, int *status);
int core_advect_rgb(fs_ctx *ctx, fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel, const Geo &g,
                    float dt, int no_slip, int *status);
