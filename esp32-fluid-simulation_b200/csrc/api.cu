// C ABI (include/fluid_b200.h): context, argument validation, operator
// orchestration.  No arithmetic of the sim lives here — only which kernels run
// in which order on which buffers.
#include <cstdio>
#include <cmath>
#include <cstring>
#include <new>

#include "ctx.h"

// ---- operator cores over a Geo (shared with dist.cu) ----------------------------------------

int core_advect_vec2f(fs_ctx *ctx, fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel,
                      const Geo &g, float dt, int no_slip, int *status)
{
    if (ctx->opt_advect == 1 && advect_vec2f_tma_legal((const float2 *)p, g))
        return launch_advect_vec2f_tma(mk(ctx), (float2 *)next_p, (const float2 *)p, (const float2 *)vel, g,
                                       dt, no_slip != 0, status);
    return launch_advect_vec2f_gather(mk(ctx), (float2 *)next_p, (const float2 *)p,
                                      (const float2 *)vel, g, dt, no_slip != 0, status);
}

int core_advect_rgb(fs_ctx *ctx, fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                    const Geo &g, float dt, int no_slip, int *status)
{
    if (ctx->opt_advect == 1 && advect_rgb_tma_legal((const uint32_t *)c, g))
        return launch_advect_rgb_tma(mk(ctx), (uint32_t *)next_c, (const uint32_t *)c, (const float2 *)vel, g,
                                     dt, no_slip != 0, status);
    return launch_advect_rgb_gather(mk(ctx), (uint32_t *)next_c, (const uint32_t *)c,
                                    (const float2 *)vel, g, dt, no_slip != 0, status);
}

namespace {

// ---- operator cores over a Geo ------------------------------------------------

// poisson_solve over a whole grid: zero, then iters x (colour 0, colour 1)
int core_poisson_solve(fs_ctx *ctx, float *p, const float *div, const Geo &g, float dx, int iters,
                       float omega)
{
    if (ctx->opt_sor == 1 && iters > 0) {
        // temporally blocked: T full iterations per HBM round trip, ping-ponging between p and a
        // scratch buffer so that the LAST pass lands in the caller's p; pass 1 starts from zero
        // without reading anything (poisson.cpp:117-119)
        const int T = ctx->opt_sor_t;
        int passes = (iters + T - 1) / T;
        // A remainder of 1 or 2 iterations is folded into the FIRST pass (T + r iterations) instead of getting a pass of
        // its own: measured at 4096^2 (tools_sor_pass_cost.py) a pass costs ~0.02 ms + ~0.011 ms per iteration, so
        // K = 50 as 8 + 7 x 6 beats 8 x 6 + 2 by 0.008 ms.  (The first pass starts from zero and reads no p.)
        int first_extra = 0;
        if (passes > 1 && ctx->opt_sor_one_launch != 1) {
            const int r = iters - (passes - 1) * T;
            if (r <= 2 && 2 * (T + r) <= SOR_BLOCKED_MAX_HALF) {
                first_extra = r;
                passes--;
            }
        }
        void *scratch;
        int e = ensure(ctx, S_P2, sizeof(float) * (size_t)g.nx * g.ny, &scratch);
        if (e) return e;
        if (ctx->opt_sor_one_launch == 1) {
            // all passes in ONE persistent launch (tile-level dependencies instead of kernel boundaries)
            const size_t need = sor_solve_flag_count(g, iters, T);
            void *flags;
            const bool fresh = ctx->scratch_bytes[S_SOLVE_FLAGS] < need * sizeof(unsigned int);
            if ((e = ensure(ctx, S_SOLVE_FLAGS, need * sizeof(unsigned int), &flags))) return e;
            if (fresh)
                FS_CUDA_TRY(cudaMemsetAsync(flags, 0, ctx->scratch_bytes[S_SOLVE_FLAGS], ctx->stream));
            if (++ctx->solve_gen == 0) {   // stamp wrapped: start over from clean flags
                FS_CUDA_TRY(cudaMemsetAsync(flags, 0, ctx->scratch_bytes[S_SOLVE_FLAGS], ctx->stream));
                ctx->solve_gen = 1;
            }
            FS_CUDA_TRY(cudaMemsetAsync(ctx->work_dev, 0, sizeof(int), ctx->stream));
            e = launch_sor_solve(mk(ctx), p, (float *)scratch, div, g, dx, omega, iters, T, ctx->opt_sor_shape,
                                 ctx->work_dev, (unsigned int *)flags,
                                 ctx->scratch_bytes[S_SOLVE_FLAGS] / sizeof(unsigned int), ctx->solve_gen);
            if (e >= 0) return e;          // launched (or a CUDA error); -1 = not eligible, fall through
        }
        float *bufs[2] = {(passes & 1) ? p : (float *)scratch, (passes & 1) ? (float *)scratch : p};
        const float *src = nullptr;
        int done = 0;
        for (int k = 0; k < passes; k++) {
            const int want = k == 0 ? T + first_extra : T;
            const int t = iters - done < want ? iters - done : want;
            float *dst = bufs[k & 1];
            if (k % WORK_SLOTS == 0)  // one zeroed tile counter per pass of the persistent kernel
                FS_CUDA_TRY(cudaMemsetAsync(ctx->work_dev, 0, WORK_SLOTS * sizeof(int), ctx->stream));
            if ((e = launch_sor_blocked(mk(ctx), dst, src, div, g, dx, omega, 0, 2 * t,
                                        ctx->opt_sor_shape, ctx->work_dev + k % WORK_SLOTS)))
                return e;
            src = dst;
            done += t;
        }
        return FS_OK;
    }
    FS_CUDA_TRY(cudaMemsetAsync(p, 0, sizeof(float) * (size_t)g.nx * g.ny, ctx->stream));
    for (int k = 0; k < iters; k++) {
        int e = launch_sor_half_sweep(mk(ctx), p, div, g, dx, omega, 0);
        if (e) return e;
        e = launch_sor_half_sweep(mk(ctx), p, div, g, dx, omega, 1);
        if (e) return e;
    }
    return FS_OK;
}

// loop() body, ino:249-289.  v is advected into v_tmp, forced and projected
// there, and the projection's last operator writes back into v (out of place
// gradient-subtract), so the reference's pointer swap (ino:255) costs no copy.
// The dye goes c_in -> c_out.
int core_step_velocity(fs_ctx *ctx, fs_vec2f *v, fs_vec2f *v_tmp, const fs_drag *drags, int n_drags,
                       const Geo &g, float dt, float dx, int iters, float omega, float *p, float *div,
                       bool with_gradient = true)
{
    int e;
    const bool whole = g.ox == 0 && g.oy == 0 && g.x0 == 0 && g.y0 == 0 && g.x1 == g.nx && g.y1 == g.ny &&
                       g.nx == g.GX && g.ny == g.GY;
    if ((ctx->opt_fuse & 1) && ctx->opt_advect == 1 && whole && n_drags <= advect_div_max_drags() &&
        advect_vec2f_tma_legal((const float2 *)v, g)) {
        // ino:253 + 264-269 + 274 in one pass over the grid
        if ((e = launch_advect_div_tma(mk(ctx), (float2 *)v_tmp, (const float2 *)v, div, drags, n_drags, g, dt, dx)))
            return e;
    } else {
        if ((e = core_advect_vec2f(ctx, v_tmp, v, v, g, dt, 1, nullptr))) return e;           // ino:253
        if (n_drags > 0 && (e = launch_apply_drags(mk(ctx), (float2 *)v_tmp, drags, n_drags, g)))
            return e;                                                                        // ino:264-269
        if ((e = launch_divergence(mk(ctx), div, (const float2 *)v_tmp, g, dx))) return e;    // ino:274
    }
    if ((e = core_poisson_solve(ctx, p, div, g, dx, iters, omega))) return e;             // ino:275
    if (!with_gradient) return FS_OK;   // folded into the dye advect by the caller
    return launch_subtract_gradient(mk(ctx), (float2 *)v, (const float2 *)v_tmp, p, g, dx);  // ino:276
}

int core_step(fs_ctx *ctx, fs_vec2f *v, fs_vec2f *v_tmp, const fs_rgb_uq32 *c_in,
              fs_rgb_uq32 *c_out, const fs_drag *drags, int n_drags, int dim_x, int dim_y, float dt,
              float dx, int iters, float omega, float *p, float *div)
{
    const Geo g = geo_full(dim_x, dim_y);
    int e;
    const bool fuse_grad = (ctx->opt_fuse & 2) && ctx->opt_advect == 1 && advect_rgb_tma_legal((const uint32_t *)c_in, g);
    if ((e = core_step_velocity(ctx, v, v_tmp, drags, n_drags, g, dt, dx, iters, omega, p, div, !fuse_grad)))
        return e;
    if (fuse_grad)   // ino:276 + ino:282 in one pass: v = v_tmp - grad p, dye advected with v
        return launch_advect_rgb_tma_grad(mk(ctx), (uint32_t *)c_out, (const uint32_t *)c_in, (float2 *)v,
                                          (const float2 *)v_tmp, p, g, dt, dx, false, nullptr);
    return core_advect_rgb(ctx, c_out, c_in, v, g, dt, 0, nullptr);                       // ino:282
}

int ensure_copy_streams(fs_ctx *ctx)
{
    if (ctx->copy_in) return FS_OK;
    FS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    FS_CUDA_TRY(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    FS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    FS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_c_in, cudaEventDisableTiming));
    FS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_v_done, cudaEventDisableTiming));
    for (int b = 0; b < E2E_MAX_BANDS; b++) {
        FS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_band_in[b], cudaEventDisableTiming));
        FS_CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_band_done[b], cudaEventDisableTiming));
    }
    FS_CUDA_TRY(cudaMalloc(&ctx->band_flag_dev, sizeof(int)));
    return FS_OK;
}

}  // namespace

extern "C" {

const char *fs_version(void) { return "fluid_b200 0.1 (sm_100a)"; }

const char *fs_error_string(int code)
{
    switch (code) {
        case FS_OK: return "ok";
        case FS_ERR_INVALID_ARG: return "invalid argument";
        case FS_ERR_NO_CONTEXT: return "no context";
        case FS_ERR_UNSUPPORTED: return "unsupported";
        case FS_ERR_HALO_OVERRUN: return "advect backtrace left the local window";
        case FS_ERR_HALO_TIMEOUT: return "halo exchange: a neighbour never signalled";
        case FS_ERR_WOULD_BLOCK: return "would block: frame slots full / no frame produced";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int fs_ctx_create(fs_ctx **out, int device, void *stream)
{
    if (!out) return FS_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    FS_CUDA_TRY(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) return (int)cudaErrorInvalidDevice;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    FS_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return (int)cudaErrorNoKernelImageForDevice;  // sm_100a only, no fallback
    fs_ctx *ctx = new (std::nothrow) fs_ctx();
    if (!ctx) return (int)cudaErrorMemoryAllocation;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->stream = (cudaStream_t)stream;
    if (stream == FS_STREAM_NEW) {          // a non-blocking stream of the context's own
        cudaStream_t own;
        cudaError_t se = cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking);
        if (se != cudaSuccess) {
            delete ctx;
            return (int)se;
        }
        ctx->stream = own;
        ctx->owns_stream = 1;
    }
    ctx->num_sms = prop.multiProcessorCount;
    ctx->max_smem_optin = prop.sharedMemPerBlockOptin;
    ctx->opt_sor = 1;
    ctx->opt_sor_t = 6;
    ctx->opt_sor_shape = 7;
    ctx->opt_sor_one_launch = 0;
    ctx->opt_halo_timeout_ms = 10000;
    ctx->opt_ens = 0;
    ctx->opt_advect = 1;
    ctx->opt_fuse = 5;
    ctx->opt_e2e_bands = 8;
    cudaError_t e = cudaMalloc(&ctx->status_dev, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->maxdisp_dev, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->work_dev, WORK_SLOTS * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->rim_dev, WORK_SLOTS * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->resid_dev, 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->halo_done_dev, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ctx->halo_done_dev, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ctx->status_dev, 0, sizeof(int));
    if (e != cudaSuccess) {
        delete ctx;
        return (int)e;
    }
    *out = ctx;
    return FS_OK;
}

int fs_ctx_destroy(fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int s = 0; s < S_COUNT; s++)
        if (ctx->scratch[s]) cudaFree(ctx->scratch[s]);
    cudaFree(ctx->status_dev);
    cudaFree(ctx->maxdisp_dev);
    cudaFree(ctx->work_dev);
    cudaFree(ctx->rim_dev);
    cudaFree(ctx->halo_done_dev);
    cudaFree(ctx->resid_dev);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_c_in) cudaEventDestroy(ctx->ev_c_in);
    if (ctx->ev_v_done) cudaEventDestroy(ctx->ev_v_done);
    for (int b = 0; b < E2E_MAX_BANDS; b++) {
        if (ctx->ev_band_in[b]) cudaEventDestroy(ctx->ev_band_in[b]);
        if (ctx->ev_band_done[b]) cudaEventDestroy(ctx->ev_band_done[b]);
    }
    if (ctx->band_flag_dev) cudaFree(ctx->band_flag_dev);
    delete ctx;
    return FS_OK;
}

int fs_ctx_synchronize(fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    DeviceGuard guard(ctx->device);
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

static int *opt_slot(fs_ctx *ctx, const char *name)
{
    if (!strcmp(name, "sor")) return &ctx->opt_sor;
    if (!strcmp(name, "sor_t")) return &ctx->opt_sor_t;
    if (!strcmp(name, "sor_shape")) return &ctx->opt_sor_shape;
    if (!strcmp(name, "sor_one_launch")) return &ctx->opt_sor_one_launch;
    if (!strcmp(name, "halo_timeout_ms")) return &ctx->opt_halo_timeout_ms;
    if (!strcmp(name, "sor_grid_limit")) return &ctx->opt_sor_grid_limit;
    if (!strcmp(name, "ensemble")) return &ctx->opt_ens;
    if (!strcmp(name, "num_sms")) return &ctx->num_sms;
    if (!strcmp(name, "advect")) return &ctx->opt_advect;
    if (!strcmp(name, "fuse")) return &ctx->opt_fuse;
    if (!strcmp(name, "e2e_bands")) return &ctx->opt_e2e_bands;
    if (!strcmp(name, "e2e_redos")) return &ctx->stat_e2e_redos;
    return nullptr;
}

int fs_ctx_set_option(fs_ctx *ctx, const char *name, int value)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!name) return FS_ERR_INVALID_ARG;
    int *slot = opt_slot(ctx, name);
    if (!slot) return FS_ERR_INVALID_ARG;
    if (slot == &ctx->opt_sor_t && (value < 1 || value > 8)) return FS_ERR_INVALID_ARG;
    if (slot == &ctx->num_sms || slot == &ctx->stat_e2e_redos) return FS_ERR_INVALID_ARG;   // read-only
    if (slot == &ctx->opt_e2e_bands && (value < 1 || value > E2E_MAX_BANDS)) return FS_ERR_INVALID_ARG;
    *slot = value;
    return FS_OK;
}

int fs_ctx_get_option(fs_ctx *ctx, const char *name, int *value)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!name || !value) return FS_ERR_INVALID_ARG;
    int *slot = opt_slot(ctx, name);
    if (!slot) return FS_ERR_INVALID_ARG;
    *value = *slot;
    return FS_OK;
}

uint64_t fs_ctx_launch_count(fs_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fs_host_alloc(void **out, size_t bytes)
{
    if (!out) return FS_ERR_INVALID_ARG;
    FS_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return FS_OK;
}

int fs_host_free(void *p)
{
    FS_CUDA_TRY(cudaFreeHost(p));
    return FS_OK;
}

// ---- device-pointer operators ---------------------------------------------------

int fs_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel, int dim_x, int dim_y,
                    float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_p || !p || !vel || bad_dims(dim_x, dim_y) || next_p == p || next_p == vel)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_advect_vec2f(ctx, next_p, p, vel, geo_full(dim_x, dim_y), dt, no_slip, nullptr);
}

int fs_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel, int dim_x,
                       int dim_y, float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_c || !c || !vel || bad_dims(dim_x, dim_y) || next_c == c) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_advect_rgb(ctx, next_c, c, vel, geo_full(dim_x, dim_y), dt, no_slip, nullptr);
}

int fs_calculate_divergence(float *div, const fs_vec2f *v, int dim_x, int dim_y, float dx,
                            fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!div || !v || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_divergence(mk(ctx), div, (const float2 *)v, geo_full(dim_x, dim_y), dx);
}

int fs_subtract_gradient(fs_vec2f *v, const float *p, int dim_x, int dim_y, float dx, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !p || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_subtract_gradient(mk(ctx), (float2 *)v, (const float2 *)v, p,
                                    geo_full(dim_x, dim_y), dx);
}

int fs_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx, int iters,
                     float omega, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!p || !div || p == div || bad_dims(dim_x, dim_y) || iters < 0) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_poisson_solve(ctx, p, div, geo_full(dim_x, dim_y), dx, iters, omega);
}

int fs_sor_half_sweep(float *p, const float *div, int dim_x, int dim_y, float dx, float omega,
                      int parity, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!p || !div || p == div || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_sor_half_sweep(mk(ctx), p, div, geo_full(dim_x, dim_y), dx, omega, parity);
}

int fs_poisson_residual(float *max_abs, double *l2, const float *p, const float *div, int dim_x, int dim_y,
                        float dx, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!p || !div || bad_dims(dim_x, dim_y) || (!max_abs && !l2)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    unsigned int *max_bits = (unsigned int *)(ctx->resid_dev + 1);
    int e = launch_sor_residual(mk(ctx), p, div, geo_full(dim_x, dim_y), dx, max_bits, ctx->resid_dev);
    if (e) return e;
    double host[2];
    FS_CUDA_TRY(cudaMemcpyAsync(host, ctx->resid_dev, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (l2) *l2 = sqrt(host[0]);
    if (max_abs) memcpy(max_abs, &host[1], sizeof(float));
    return FS_OK;
}

int fs_apply_drags(fs_vec2f *v, const fs_drag *drags, int n, int dim_x, int dim_y, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || n < 0 || (n > 0 && !drags) || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_apply_drags(mk(ctx), (float2 *)v, drags, n, geo_full(dim_x, dim_y));
}

int fs_advect_drags_divergence(fs_vec2f *v_out, float *div, const fs_vec2f *v_in, const fs_drag *drags, int n_drags,
                               int dim_x, int dim_y, float dt, float dx, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v_out || !div || !v_in || v_out == v_in || n_drags < 0 || (n_drags > 0 && !drags) || bad_dims(dim_x, dim_y))
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const Geo g = geo_full(dim_x, dim_y);
    int e;
    if (ctx->opt_advect == 1 && n_drags <= advect_div_max_drags() && advect_vec2f_tma_legal((const float2 *)v_in, g))
        return launch_advect_div_tma(mk(ctx), (float2 *)v_out, (const float2 *)v_in, div, drags, n_drags, g, dt, dx);
    if ((e = core_advect_vec2f(ctx, v_out, v_in, v_in, g, dt, 1, nullptr))) return e;
    if (n_drags > 0 && (e = launch_apply_drags(mk(ctx), (float2 *)v_out, drags, n_drags, g))) return e;
    return launch_divergence(mk(ctx), div, (const float2 *)v_out, g, dx);
}

// dye advect + frame: fused where the TMA kernel is legal, else advect followed by the stand-alone upscale
static int core_advect_rgb_frame(fs_ctx *ctx, fs_rgb_uq32 *next_c, uint16_t *frame, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                                 int dim_x, int dim_y, float dt, int no_slip)
{
    const Geo g = geo_full(dim_x, dim_y);
    if (ctx->opt_advect == 1 && (ctx->opt_fuse & 4) && advect_rgb_tma_legal((const uint32_t *)c, g))
        return launch_advect_rgb_frame(mk(ctx), (uint32_t *)next_c, frame, dim_y - 1, (const uint32_t *)c, (const float2 *)vel,
                                       g, dt, no_slip != 0, nullptr);
    int e = core_advect_rgb(ctx, next_c, c, vel, g, dt, no_slip, nullptr);
    if (e) return e;
    return launch_upscale4_rgb565(mk(ctx), frame, (const uint32_t *)next_c, dim_x, dim_y);
}

int fs_advect_rgb_frame(fs_rgb_uq32 *next_c, uint16_t *frame, const fs_rgb_uq32 *c, const fs_vec2f *vel, int dim_x,
                        int dim_y, float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_c || !frame || !c || !vel || bad_dims(dim_x, dim_y) || next_c == c) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_advect_rgb_frame(ctx, next_c, frame, c, vel, dim_x, dim_y, dt, no_slip);
}

int fs_step_frame(fs_vec2f *v, const fs_rgb_uq32 *c_in, fs_rgb_uq32 *c_out, uint16_t *frame, const fs_drag *drags,
                  int n_drags, int dim_x, int dim_y, float dt, float dx, int iters, float omega, float *p_out,
                  float *div_out, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c_in || !c_out || !frame || c_in == c_out || n_drags < 0 || (n_drags > 0 && !drags) ||
        bad_dims(dim_x, dim_y) || iters < 0)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *v_tmp, *p = p_out, *d = div_out;
    int e;
    if ((e = ensure(ctx, S_VTMP, n * sizeof(fs_vec2f), &v_tmp))) return e;
    if (!p && (e = ensure(ctx, S_P, n * sizeof(float), &p))) return e;
    if (!d && (e = ensure(ctx, S_DIV, n * sizeof(float), &d))) return e;
    if ((e = core_step_velocity(ctx, v, (fs_vec2f *)v_tmp, drags, n_drags, geo_full(dim_x, dim_y), dt, dx, iters, omega,
                                (float *)p, (float *)d)))
        return e;
    return core_advect_rgb_frame(ctx, c_out, frame, c_in, v, dim_x, dim_y, dt, 0);   // ino:282 + ino:116-177
}

int fs_step_pingpong(fs_vec2f *v, const fs_rgb_uq32 *c_in, fs_rgb_uq32 *c_out, const fs_drag *drags,
                     int n_drags, int dim_x, int dim_y, float dt, float dx, int iters, float omega,
                     float *p_out, float *div_out, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c_in || !c_out || c_in == c_out || n_drags < 0 || (n_drags > 0 && !drags) ||
        bad_dims(dim_x, dim_y) || iters < 0)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *v_tmp, *p = p_out, *d = div_out;
    int e;
    if ((e = ensure(ctx, S_VTMP, n * sizeof(fs_vec2f), &v_tmp))) return e;
    if (!p && (e = ensure(ctx, S_P, n * sizeof(float), &p))) return e;
    if (!d && (e = ensure(ctx, S_DIV, n * sizeof(float), &d))) return e;
    return core_step(ctx, v, (fs_vec2f *)v_tmp, c_in, c_out, drags, n_drags, dim_x, dim_y, dt, dx, iters,
                     omega, (float *)p, (float *)d);
}

int fs_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags, int n_drags, int dim_x, int dim_y,
            float dt, float dx, int iters, float omega, float *p_out, float *div_out, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!c || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *c_tmp;
    int e;
    if ((e = ensure(ctx, S_CTMP, n * sizeof(fs_rgb_uq32), &c_tmp))) return e;
    if ((e = fs_step_pingpong(v, c, (fs_rgb_uq32 *)c_tmp, drags, n_drags, dim_x, dim_y, dt, dx, iters, omega,
                              p_out, div_out, ctx)))
        return e;
    // the reference swaps pointers (ino:286); an in-place raw-pointer ABI has to copy back —
    // fs_step_pingpong is the entry point without this copy
    FS_CUDA_TRY(cudaMemcpyAsync(c, c_tmp, n * sizeof(fs_rgb_uq32), cudaMemcpyDeviceToDevice,
                                ctx->stream));
    return FS_OK;
}

int fs_upscale4_rgb565(uint16_t *out, const fs_rgb_uq32 *c, int dim_x, int dim_y, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out || !c || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_upscale4_rgb565(mk(ctx), out, (const uint32_t *)c, dim_x, dim_y);
}

int fs_ensemble_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags, const int *drag_counts,
                     int max_drags, int batch, int dim_x, int dim_y, float dt, float dx, int iters,
                     float omega, int n_steps, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c || batch < 0 || n_steps < 0 || iters < 0 || max_drags < 0 || bad_dims(dim_x, dim_y) ||
        (max_drags > 0 && (!drags || !drag_counts)))
        return FS_ERR_INVALID_ARG;
    if (!ensemble_supported(dim_x, dim_y, ctx->max_smem_optin)) return FS_ERR_UNSUPPORTED;
    if (batch == 0 || n_steps == 0) return FS_OK;
    DeviceGuard guard(ctx->device);
    void *d_drags = nullptr, *d_counts = nullptr;
    if (max_drags > 0) {
        const size_t slots = (size_t)n_steps * batch;
        int e;
        if ((e = ensure(ctx, S_EDRAG, slots * max_drags * sizeof(fs_drag), &d_drags))) return e;
        if ((e = ensure(ctx, S_ECNT, slots * sizeof(int), &d_counts))) return e;
        FS_CUDA_TRY(cudaMemcpyAsync(d_drags, drags, slots * max_drags * sizeof(fs_drag),
                                    cudaMemcpyHostToDevice, ctx->stream));
        FS_CUDA_TRY(cudaMemcpyAsync(d_counts, drag_counts, slots * sizeof(int), cudaMemcpyHostToDevice,
                                    ctx->stream));
    }
    void *d_scratch;
    {
        const int grid = ensemble_grid(batch, dim_x, dim_y, ctx->num_sms, ctx->opt_ens);
        int e = ensure(ctx, S_ECTMP, ensemble_scratch_bytes(dim_x, dim_y, grid), &d_scratch);
        if (e) return e;
    }
    return launch_ensemble(mk(ctx), (float2 *)v, (uint32_t *)c, (uint32_t *)d_scratch, (const fs_drag *)d_drags,
                           (const int *)d_counts, max_drags, batch, dim_x, dim_y, dt, dx, iters, omega,
                           n_steps, ctx->opt_ens);
}

int fs_ensemble_step_dev(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags_dev, const int *counts_dev, int max_drags,
                         int batch, int dim_x, int dim_y, float dt, float dx, int iters, float omega, int n_steps,
                         fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c || batch < 0 || n_steps < 0 || iters < 0 || max_drags < 0 || bad_dims(dim_x, dim_y) ||
        (max_drags > 0 && (!drags_dev || !counts_dev)))
        return FS_ERR_INVALID_ARG;
    if (!ensemble_supported(dim_x, dim_y, ctx->max_smem_optin)) return FS_ERR_UNSUPPORTED;
    if (batch == 0 || n_steps == 0) return FS_OK;
    DeviceGuard guard(ctx->device);
    void *d_scratch;
    const int grid = ensemble_grid(batch, dim_x, dim_y, ctx->num_sms, ctx->opt_ens);
    int e = ensure(ctx, S_ECTMP, ensemble_scratch_bytes(dim_x, dim_y, grid), &d_scratch);
    if (e) return e;
    return launch_ensemble(mk(ctx), (float2 *)v, (uint32_t *)c, (uint32_t *)d_scratch, drags_dev, counts_dev, max_drags,
                           batch, dim_x, dim_y, dt, dx, iters, omega, n_steps, ctx->opt_ens);
}

int fs_init_color_wheel(fs_vec2f *v, fs_rgb_uq32 *c, int batch, int dim_x, int dim_y, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c || batch < 0 || bad_dims(dim_x, dim_y) || (long long)batch * dim_x * dim_y > 0x7fffffffLL)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_init_color_wheel(mk(ctx), (float2 *)v, (uint32_t *)c, batch, dim_x, dim_y);
}

int fs_touch_to_drags(fs_drag *drags_out, int *counts_out, const int *samples, int n_samples, int batch, int max_drags,
                      int n_rows, int n_cols, const fs_touch_cal *cal, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!drags_out || !counts_out || !samples || n_samples < 0 || batch < 0 || max_drags < 1 || n_rows < 1 || n_cols < 1)
        return FS_ERR_INVALID_ARG;
    const fs_touch_cal def = {200, 3700, 240, 3800, 10};      // ino:17-21
    if (!cal) cal = &def;
    if (cal->polling_ms <= 0) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const int c4[4] = {cal->min_x, cal->max_x, cal->min_y, cal->max_y};
    return launch_touch_to_drags(mk(ctx), drags_out, counts_out, samples, n_samples, batch, max_drags, n_rows, n_cols, c4,
                                 cal->polling_ms);
}

// ---- host-pointer drop-ins --------------------------------------------------------

#define H2D(dst, src, bytes) FS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, bytes) FS_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream))

int fsh_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel, int dim_x, int dim_y,
                     float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_p || !p || !vel || bad_dims(dim_x, dim_y) || next_p == p || next_p == vel)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t bytes = (size_t)dim_x * dim_y * sizeof(fs_vec2f);
    void *d_p, *d_vel, *d_out;
    int e;
    if ((e = ensure(ctx, S_HV, bytes, &d_p))) return e;
    if ((e = ensure(ctx, S_VTMP, bytes, &d_out))) return e;
    H2D(d_p, p, bytes);
    d_vel = d_p;
    if (vel != p) {
        if ((e = ensure(ctx, S_HV2, bytes, &d_vel))) return e;
        H2D(d_vel, vel, bytes);
    }
    if ((e = core_advect_vec2f(ctx, (fs_vec2f *)d_out, (fs_vec2f *)d_p, (fs_vec2f *)d_vel,
                               geo_full(dim_x, dim_y), dt, no_slip, nullptr)))
        return e;
    D2H(next_p, d_out, bytes);
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

int fsh_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel, int dim_x,
                        int dim_y, float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_c || !c || !vel || bad_dims(dim_x, dim_y) || next_c == c) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *d_c, *d_vel, *d_out;
    int e;
    if ((e = ensure(ctx, S_HC, n * sizeof(fs_rgb_uq32), &d_c))) return e;
    if ((e = ensure(ctx, S_HV, n * sizeof(fs_vec2f), &d_vel))) return e;
    if ((e = ensure(ctx, S_CTMP, n * sizeof(fs_rgb_uq32), &d_out))) return e;
    H2D(d_c, c, n * sizeof(fs_rgb_uq32));
    H2D(d_vel, vel, n * sizeof(fs_vec2f));
    if ((e = core_advect_rgb(ctx, (fs_rgb_uq32 *)d_out, (fs_rgb_uq32 *)d_c, (fs_vec2f *)d_vel,
                             geo_full(dim_x, dim_y), dt, no_slip, nullptr)))
        return e;
    D2H(next_c, d_out, n * sizeof(fs_rgb_uq32));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

int fsh_calculate_divergence(float *div, const fs_vec2f *v, int dim_x, int dim_y, float dx,
                             fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!div || !v || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *d_v, *d_div;
    int e;
    if ((e = ensure(ctx, S_HV, n * sizeof(fs_vec2f), &d_v))) return e;
    if ((e = ensure(ctx, S_HD, n * sizeof(float), &d_div))) return e;
    H2D(d_v, v, n * sizeof(fs_vec2f));
    if ((e = launch_divergence(mk(ctx), (float *)d_div, (const float2 *)d_v, geo_full(dim_x, dim_y), dx)))
        return e;
    D2H(div, d_div, n * sizeof(float));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

int fsh_subtract_gradient(fs_vec2f *v, const float *p, int dim_x, int dim_y, float dx, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !p || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *d_v, *d_p;
    int e;
    if ((e = ensure(ctx, S_HV, n * sizeof(fs_vec2f), &d_v))) return e;
    if ((e = ensure(ctx, S_HP, n * sizeof(float), &d_p))) return e;
    H2D(d_v, v, n * sizeof(fs_vec2f));
    H2D(d_p, p, n * sizeof(float));
    if ((e = launch_subtract_gradient(mk(ctx), (float2 *)d_v, (const float2 *)d_v, (const float *)d_p,
                                      geo_full(dim_x, dim_y), dx)))
        return e;
    D2H(v, d_v, n * sizeof(fs_vec2f));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

int fsh_poisson_solve(float *p, const float *div, int dim_x, int dim_y, float dx, int iters,
                      float omega, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!p || !div || p == div || bad_dims(dim_x, dim_y) || iters < 0) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *d_p, *d_div;
    int e;
    if ((e = ensure(ctx, S_HP, n * sizeof(float), &d_p))) return e;
    if ((e = ensure(ctx, S_HD, n * sizeof(float), &d_div))) return e;
    H2D(d_div, div, n * sizeof(float));
    if ((e = core_poisson_solve(ctx, (float *)d_p, (const float *)d_div, geo_full(dim_x, dim_y), dx,
                                iters, omega)))
        return e;
    D2H(p, d_p, n * sizeof(float));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

int fsh_step(fs_vec2f *v, fs_rgb_uq32 *c, const fs_drag *drags, int n_drags, int dim_x, int dim_y,
             float dt, float dx, int iters, float omega, float *p_out, float *div_out, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !c || n_drags < 0 || (n_drags > 0 && !drags) || bad_dims(dim_x, dim_y) || iters < 0)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    void *d_v, *d_vtmp, *d_c, *d_c2, *d_p, *d_div;
    int e;
    if ((e = ensure(ctx, S_HV, n * sizeof(fs_vec2f), &d_v))) return e;
    if ((e = ensure(ctx, S_VTMP, n * sizeof(fs_vec2f), &d_vtmp))) return e;
    if ((e = ensure(ctx, S_HC, n * sizeof(fs_rgb_uq32), &d_c))) return e;
    if ((e = ensure(ctx, S_HC2, n * sizeof(fs_rgb_uq32), &d_c2))) return e;
    if ((e = ensure(ctx, S_P, n * sizeof(float), &d_p))) return e;
    if ((e = ensure(ctx, S_DIV, n * sizeof(float), &d_div))) return e;
    if ((e = ensure_copy_streams(ctx))) return e;
    // The step is PCIe-bound (20 B/node each way), so the schedule is built around the two copy engines.
    // Up: velocity first (the only thing the projection needs), then the dye in row bands on a side
    // stream under the velocity phase.  Down: the projected velocity leaves on a second side stream as
    // soon as it exists, and dye band b follows as soon as it is advected — which needs the projected
    // velocity and dye bands 0..b+1 only, so the last bands are still arriving while the first ones go
    // home.  A backtrace longer than a band (>= 128 rows) would read dye that has not arrived: the
    // advect's valid-rectangle check raises a flag instead of using it, and the dye is then advected
    // again in one piece once everything is here.
    const Geo g = geo_full(dim_x, dim_y);
    int bands = ctx->opt_e2e_bands;
    if (bands > dim_y / 128) bands = dim_y / 128;
    if (bands < 1 || !advect_rgb_tma_legal((const uint32_t *)d_c, g)) bands = 1;
    int row0[E2E_MAX_BANDS + 1];
    for (int b = 0; b <= bands; b++) row0[b] = b == bands ? dim_y : (int)((long long)dim_y * b / bands) / 32 * 32;
    const size_t row_c = (size_t)dim_x * sizeof(fs_rgb_uq32);
    FS_CUDA_TRY(cudaEventRecord(ctx->ev_start, ctx->stream));          // earlier work on the scratch buffers
    FS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_start, 0));
    FS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_start, 0));
    H2D(d_v, v, n * sizeof(fs_vec2f));
    FS_CUDA_TRY(cudaEventRecord(ctx->ev_c_in, ctx->stream));           // the velocity owns the link first
    FS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_c_in, 0));
    for (int b = 0; b < bands; b++) {
        FS_CUDA_TRY(cudaMemcpyAsync((char *)d_c + row0[b] * row_c, (const char *)c + row0[b] * row_c,
                                    (row0[b + 1] - row0[b]) * row_c, cudaMemcpyHostToDevice, ctx->copy_in));
        FS_CUDA_TRY(cudaEventRecord(ctx->ev_band_in[b], ctx->copy_in));
    }
    if ((e = core_step_velocity(ctx, (fs_vec2f *)d_v, (fs_vec2f *)d_vtmp, drags, n_drags, g, dt, dx, iters,
                                omega, (float *)d_p, (float *)d_div)))
        return e;
    FS_CUDA_TRY(cudaEventRecord(ctx->ev_v_done, ctx->stream));
    FS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_v_done, 0));
    FS_CUDA_TRY(cudaMemcpyAsync(v, d_v, n * sizeof(fs_vec2f), cudaMemcpyDeviceToHost, ctx->copy_out));
    if (p_out)
        FS_CUDA_TRY(cudaMemcpyAsync(p_out, d_p, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
    if (div_out)
        FS_CUDA_TRY(cudaMemcpyAsync(div_out, d_div, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
    if (bands == 1) {
        FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_band_in[0], 0));
        if ((e = core_advect_rgb(ctx, (fs_rgb_uq32 *)d_c2, (fs_rgb_uq32 *)d_c, (fs_vec2f *)d_v, g, dt, 0, nullptr)))
            return e;
        D2H(c, d_c2, n * sizeof(fs_rgb_uq32));
        FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        FS_CUDA_TRY(cudaStreamSynchronize(ctx->copy_out));
        return FS_OK;
    }
    FS_CUDA_TRY(cudaMemsetAsync(ctx->band_flag_dev, 0, sizeof(int), ctx->stream));
    for (int b = 0; b < bands; b++) {
        const int have = b + 1 < bands ? b + 1 : bands - 1;            // last band that must have arrived
        FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_band_in[have], 0));
        Geo gb = g;
        gb.y0 = row0[b];
        gb.y1 = row0[b + 1];
        gb.vy1 = row0[have + 1];
        if ((e = core_advect_rgb(ctx, (fs_rgb_uq32 *)d_c2, (fs_rgb_uq32 *)d_c, (fs_vec2f *)d_v, gb, dt, 0,
                                 ctx->band_flag_dev)))
            return e;
        FS_CUDA_TRY(cudaEventRecord(ctx->ev_band_done[b], ctx->stream));
        FS_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_band_done[b], 0));
        FS_CUDA_TRY(cudaMemcpyAsync((char *)c + row0[b] * row_c, (const char *)d_c2 + row0[b] * row_c,
                                    (row0[b + 1] - row0[b]) * row_c, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    int flag = 0;
    FS_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->band_flag_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->copy_out));
    if (flag) {
        ctx->stat_e2e_redos++;
        if ((e = core_advect_rgb(ctx, (fs_rgb_uq32 *)d_c2, (fs_rgb_uq32 *)d_c, (fs_vec2f *)d_v, g, dt, 0, nullptr)))
            return e;
        D2H(c, d_c2, n * sizeof(fs_rgb_uq32));
        FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return FS_OK;
}

int fsh_upscale4_rgb565(uint16_t *out, const fs_rgb_uq32 *c, int dim_x, int dim_y, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out || !c || bad_dims(dim_x, dim_y)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)dim_x * dim_y;
    const size_t img = 16 * (size_t)(dim_x - 1) * (dim_y - 1) * sizeof(uint16_t);
    void *d_c, *d_img;
    int e;
    if ((e = ensure(ctx, S_HC, n * sizeof(fs_rgb_uq32), &d_c))) return e;
    if ((e = ensure(ctx, S_HIMG, img ? img : 16, &d_img))) return e;
    H2D(d_c, c, n * sizeof(fs_rgb_uq32));
    if ((e = launch_upscale4_rgb565(mk(ctx), (uint16_t *)d_img, (const uint32_t *)d_c, dim_x, dim_y)))
        return e;
    if (img) D2H(out, d_img, img);
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FS_OK;
}

// ---- decomposed grids ---------------------------------------------------------------

int fs_tile_advect_vec2f(fs_vec2f *next_p, const fs_vec2f *p, const fs_vec2f *vel, const fs_tile *t,
                         float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_p || !p || !vel || bad_tile(t, 0) || next_p == p || next_p == vel)
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_advect_vec2f(ctx, next_p, p, vel, geo_tile(*t), dt, no_slip, ctx->status_dev);
}

int fs_tile_advect_rgb_uq32(fs_rgb_uq32 *next_c, const fs_rgb_uq32 *c, const fs_vec2f *vel,
                            const fs_tile *t, float dt, int no_slip, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!next_c || !c || !vel || bad_tile(t, 0) || next_c == c) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return core_advect_rgb(ctx, next_c, c, vel, geo_tile(*t), dt, no_slip, ctx->status_dev);
}

int fs_tile_calculate_divergence(float *div, const fs_vec2f *v, const fs_tile *t, float dx,
                                 fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!div || !v || bad_tile(t, 1)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_divergence(mk(ctx), div, (const float2 *)v, geo_tile(*t), dx);
}

int fs_tile_subtract_gradient(fs_vec2f *v, const float *p, const fs_tile *t, float dx, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || !p || bad_tile(t, 1)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_subtract_gradient(mk(ctx), (float2 *)v, (const float2 *)v, p, geo_tile(*t), dx);
}

int fs_tile_sor_sweeps(float *p_out, const float *p_in, const float *div, const fs_tile *t, float dx,
                       float omega, int first_parity, int n_half, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!p_out || !div || p_out == p_in || p_out == div || n_half < 0 || bad_tile(t, n_half))
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    const Geo g = geo_tile(*t);
    if (ctx->opt_sor == 1 && n_half > 0 && n_half <= SOR_BLOCKED_MAX_HALF) {
        FS_CUDA_TRY(cudaMemsetAsync(ctx->work_dev, 0, sizeof(int), ctx->stream));
        return launch_sor_blocked(mk(ctx), p_out, p_in, div, g, dx, omega, first_parity, n_half,
                                  ctx->opt_sor_shape, ctx->work_dev);
    }
    // seed p_out on the rectangle grown by n_half, then sweep in place on
    // rectangles that shrink by one node per half-sweep
    const Geo seed = grown(g, n_half);
    const size_t w = (size_t)(seed.x1 - seed.x0) * sizeof(float), h = seed.y1 - seed.y0;
    if (w == 0 || h == 0) return FS_OK;
    const size_t off = (size_t)seed.y0 * g.nx + seed.x0, pitch = (size_t)g.nx * sizeof(float);
    if (p_in) {
        FS_CUDA_TRY(cudaMemcpy2DAsync(p_out + off, pitch, p_in + off, pitch, w, h,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        FS_CUDA_TRY(cudaMemset2DAsync(p_out + off, pitch, 0, w, h, ctx->stream));
    }
    for (int s = 0; s < n_half; s++) {
        int e = launch_sor_half_sweep(mk(ctx), p_out, div, grown(g, n_half - 1 - s), dx, omega,
                                      (first_parity + s) & 1);
        if (e) return e;
    }
    return FS_OK;
}

int fs_tile_apply_drags(fs_vec2f *v, const fs_drag *drags, int n, const fs_tile *t, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!v || n < 0 || (n > 0 && !drags) || bad_tile(t, 0)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    return launch_apply_drags(mk(ctx), (float2 *)v, drags, n, geo_tile(*t));
}

int fs_tile_check(fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    DeviceGuard guard(ctx->device);
    int flag = 0;
    FS_CUDA_TRY(cudaMemcpyAsync(&flag, ctx->status_dev, sizeof(int), cudaMemcpyDeviceToHost,
                                ctx->stream));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (flag) {
        FS_CUDA_TRY(cudaMemsetAsync(ctx->status_dev, 0, sizeof(int), ctx->stream));
        return flag;
    }
    return FS_OK;
}

int fs_tile_max_displacement(int *out_nodes, const fs_vec2f *vel, const fs_tile *t, float dt,
                             fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out_nodes || !vel || bad_tile(t, 0)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    int e = launch_max_displacement(mk(ctx), ctx->maxdisp_dev, (const float2 *)vel, geo_tile(*t));
    if (e) return e;
    unsigned int bits = 0;
    FS_CUDA_TRY(cudaMemcpyAsync(&bits, ctx->maxdisp_dev, sizeof(bits), cudaMemcpyDeviceToHost,
                                ctx->stream));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float m;
    memcpy(&m, &bits, sizeof(m));
    // |v*dt| <= m*|dt| up to one rounding; +1 node covers it and the +1 corner of the bilinear cell
    double d = (double)m * (dt < 0 ? -(double)dt : (double)dt);
    long long nodes = (long long)d + 2;
    *out_nodes = nodes > 0x3fffffff ? 0x3fffffff : (int)nodes;
    return FS_OK;
}

// ---- halo exchange over peer memory ------------------------------------------------------------

int fs_ctx_set_stream(fs_ctx *ctx, void *stream)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    ctx->stream = (cudaStream_t)stream;
    return FS_OK;
}

int fs_arena_alloc(void **out, size_t bytes, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out || bytes == 0) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    FS_CUDA_TRY(cudaMalloc(out, bytes));   // cudaMalloc (not a pool): exportable through CUDA IPC
    FS_CUDA_TRY(cudaMemset(*out, 0, bytes));
    return FS_OK;
}

int fs_arena_free(void *arena, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    DeviceGuard guard(ctx->device);
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    FS_CUDA_TRY(cudaFree(arena));
    return FS_OK;
}

int fs_ipc_export(void *arena, unsigned char handle[64], fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!arena || !handle) return FS_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    FS_CUDA_TRY(cudaIpcGetMemHandle(&h, arena));
    memcpy(handle, &h, 64);
    return FS_OK;
}

int fs_ipc_open(void **peer_arena, const unsigned char handle[64], fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!peer_arena || !handle) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    FS_CUDA_TRY(cudaIpcOpenMemHandle(peer_arena, h, cudaIpcMemLazyEnablePeerAccess));
    return FS_OK;
}

int fs_ipc_close(void *peer_arena, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    DeviceGuard guard(ctx->device);
    FS_CUDA_TRY(cudaIpcCloseMemHandle(peer_arena));
    return FS_OK;
}

int fs_halo_exchange(const fs_halo_copy *copies, int n_copies, void *const *signal_flags,
                     void *const *wait_flags, int n_peers, unsigned long long seq, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (n_copies < 0 || n_copies > FS_HALO_MAX_COPIES || n_peers < 0 || n_peers > FS_HALO_MAX_PEERS ||
        (n_copies > 0 && !copies) || (n_peers > 0 && (!signal_flags || !wait_flags)))
        return FS_ERR_INVALID_ARG;
    DeviceGuard guard(ctx->device);
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < n_copies; c++) {
        const fs_halo_copy &s = copies[c];
        if (!s.src || !s.dst || s.row_bytes < 0 || s.rows < 0 || (s.row_bytes & 3) || (s.src_pitch & 3) ||
            (s.dst_pitch & 3) || ((uintptr_t)s.src & 3) || ((uintptr_t)s.dst & 3))
            return FS_ERR_INVALID_ARG;
        a.copies[c].src = (const uint32_t *)s.src;
        a.copies[c].dst = (uint32_t *)s.dst;
        a.copies[c].src_pitch_words = (int)(s.src_pitch / 4);
        a.copies[c].dst_pitch_words = (int)(s.dst_pitch / 4);
        a.copies[c].row_words = s.row_bytes / 4;
        a.copies[c].rows = s.rows;
    }
    for (int k = 0; k < n_peers; k++) {
        a.signal[k] = (unsigned long long *)signal_flags[k];
        a.wait[k] = (unsigned long long *)wait_flags[k];
    }
    a.seq = seq;
    a.timeout_ns = ctx->opt_halo_timeout_ms > 0 ? (unsigned long long)ctx->opt_halo_timeout_ms * 1000000ull : 0ull;
    a.n_copies = n_copies;
    a.n_peers = n_peers;
    return launch_halo_exchange(mk(ctx), a, ctx->halo_done_dev, ctx->status_dev);
}

}  // extern "C"
