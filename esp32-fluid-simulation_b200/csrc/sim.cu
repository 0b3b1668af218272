// fs_sim_* — a device-resident simulation stepped by ONE CUDA-graph launch per loop() body, with the
// sketch's colour hand-off (ino:285-288) as an asynchronous frame stream.
//
// The sketch's loop() task produces a dye field per step and hands it to the draw task through a
// pair of binary semaphores (color_consumed / color_produced) — a double buffer.  Here:
//   * the state (v, two dye buffers, scratch) stays on the device; the caller's arrays are touched
//     only by fs_sim_upload / fs_sim_download;
//   * a step is a captured CUDA graph (fused advect+drags+divergence, the SOR passes with their
//     programmatic-dependent-launch edges, gradient-subtract, dye advect with the RGB565 frame
//     rendered in the same kernel) — one cudaGraphLaunch; the only per-step kernel arguments, the
//     drag records, are written into the captured kernel node with cudaGraphExecKernelNodeSetParams;
//     two graphs (even / odd steps) stand for the dye pointer swap of ino:286;
//   * the frame of step s goes to one of TWO device frame buffers and from there to one of two
//     pinned host buffers on a copy stream, overlapping step s+1 ("color_produced"); the producer may
//     run at most two frames ahead of the consumer, who takes a frame with fs_sim_acquire_frame and
//     gives the slot back with fs_sim_release_frame ("color_consumed").
#include <cstdio>
#include <new>

#include "ctx.h"

struct fs_sim {
    fs_ctx *ctx;
    int dim_x, dim_y, iters, want_frame;
    float dt, dx, omega;
    size_t n, frame_px;
    fs_vec2f *v;
    fs_rgb_uq32 *c[2];
    float *p, *div;
    uint16_t *frame_dev[2];
    uint16_t *frame_host[2];            // pinned
    cudaStream_t copy_stream;
    cudaEvent_t ev_step[2], ev_copied[2];
    unsigned long long produced, consumed;   // frames
    unsigned long long steps;
    // graphs: [parity]
    cudaGraph_t graph[2];               // kept alive: drag_node[] are handles into them
    cudaGraphExec_t exec[2];
    cudaGraphNode_t drag_node[2];
    cudaKernelNodeParams drag_node_params[2];
    void *drag_storage;                 // advect_div_params_bytes()
    void *captured_vtmp;                // the context's velocity scratch the graphs were captured with
    uint64_t kernels_per_graph;
    int graph_state;                    // 0 = not captured yet, 1 = captured, -1 = graphs unavailable (eager stepping)
    unsigned long long graph_launches, eager_steps;
};

namespace {

int eager_step(fs_sim *s, int parity, const fs_drag *drags, int n_drags)
{
    if (s->want_frame)
        return fs_step_frame(s->v, s->c[parity], s->c[parity ^ 1], s->frame_dev[parity], drags, n_drags, s->dim_x, s->dim_y,
                             s->dt, s->dx, s->iters, s->omega, s->p, s->div, s->ctx);
    return fs_step_pingpong(s->v, s->c[parity], s->c[parity ^ 1], drags, n_drags, s->dim_x, s->dim_y, s->dt, s->dx, s->iters,
                            s->omega, s->p, s->div, s->ctx);
}

// capture one step per parity; find the kernel node that carries the drag records
int capture_graphs(fs_sim *s)
{
    fs_ctx *ctx = s->ctx;
    const Geo g = geo_full(s->dim_x, s->dim_y);
    if (!((ctx->opt_fuse & 1) && ctx->opt_advect == 1 && advect_vec2f_tma_legal((const float2 *)s->v, g))) return -1;
    cudaStream_t user = ctx->stream, cap;
    if (cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) return -1;
    int ok = 1;
    for (int parity = 0; parity < 2 && ok; parity++) {
        cudaGraph_t graph = nullptr;
        ctx->stream = cap;
        const uint64_t launches0 = ctx->launches;
        ok = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int e = eager_step(s, parity, nullptr, 0);
            const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
            ok = e == FS_OK && ce == cudaSuccess && graph != nullptr;
        }
        ctx->stream = user;
        s->kernels_per_graph = ctx->launches - launches0;
        ctx->launches = launches0;          // capturing is not launching
        if (ok) {
            size_t nn = 0;
            ok = cudaGraphGetNodes(graph, nullptr, &nn) == cudaSuccess && nn > 0;
            cudaGraphNode_t *nodes = ok ? new (std::nothrow) cudaGraphNode_t[nn] : nullptr;
            ok = ok && nodes && cudaGraphGetNodes(graph, nodes, &nn) == cudaSuccess;
            bool found = false;
            for (size_t k = 0; ok && k < nn && !found; k++) {
                cudaGraphNodeType ty;
                if (cudaGraphNodeGetType(nodes[k], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
                cudaKernelNodeParams kp;
                if (cudaGraphKernelNodeGetParams(nodes[k], &kp) != cudaSuccess) continue;
                if (kp.func == advect_div_kernel_func()) {
                    s->drag_node[parity] = nodes[k];
                    s->drag_node_params[parity] = kp;
                    found = true;
                }
            }
            delete[] nodes;
            ok = ok && found && cudaGraphInstantiate(&s->exec[parity], graph, 0) == cudaSuccess;
        }
        s->graph[parity] = graph;
    }
    cudaStreamDestroy(cap);
    s->captured_vtmp = ctx->scratch[S_VTMP];
    if (!ok) {
        cudaGetLastError();                 // clear a capture failure: eager stepping still works
        for (int k = 0; k < 2; k++) {
            if (s->exec[k]) { cudaGraphExecDestroy(s->exec[k]); s->exec[k] = nullptr; }
            if (s->graph[k]) { cudaGraphDestroy(s->graph[k]); s->graph[k] = nullptr; }
        }
        return -1;
    }
    return 1;
}

}  // namespace

extern "C" {

int fs_sim_create(fs_sim **out, int dim_x, int dim_y, float dt, float dx, int iters, float omega, int frame, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out || bad_dims(dim_x, dim_y) || iters < 0) return FS_ERR_INVALID_ARG;
    *out = nullptr;
    DeviceGuard guard(ctx->device);
    fs_sim *s = new (std::nothrow) fs_sim();
    if (!s) return (int)cudaErrorMemoryAllocation;
    memset(s, 0, sizeof(*s));
    s->ctx = ctx;
    s->dim_x = dim_x; s->dim_y = dim_y; s->iters = iters; s->want_frame = frame ? 1 : 0;
    s->dt = dt; s->dx = dx; s->omega = omega;
    s->n = (size_t)dim_x * dim_y;
    s->frame_px = (size_t)16 * (dim_x - 1) * (dim_y - 1);
    cudaError_t e = cudaMalloc(&s->v, s->n * sizeof(fs_vec2f));
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaMalloc(&s->c[k], s->n * sizeof(fs_rgb_uq32));
    if (e == cudaSuccess) e = cudaMalloc(&s->p, s->n * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->div, s->n * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&s->ev_step[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_copied[k], cudaEventDisableTiming);
        if (e == cudaSuccess && s->want_frame) e = cudaMalloc(&s->frame_dev[k], s->frame_px * sizeof(uint16_t));
        if (e == cudaSuccess && s->want_frame) e = cudaHostAlloc(&s->frame_host[k], s->frame_px * sizeof(uint16_t), cudaHostAllocDefault);
    }
    s->drag_storage = ::operator new(advect_div_params_bytes(), std::nothrow);
    if (e != cudaSuccess || !s->drag_storage) {
        fs_sim_destroy(s);
        return e != cudaSuccess ? (int)e : (int)cudaErrorMemoryAllocation;
    }
    *out = s;
    return FS_OK;
}

int fs_sim_destroy(fs_sim *s)
{
    if (!s) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->copy_stream) cudaStreamSynchronize(s->copy_stream);
    for (int k = 0; k < 2; k++) {
        if (s->exec[k]) cudaGraphExecDestroy(s->exec[k]);
        if (s->graph[k]) cudaGraphDestroy(s->graph[k]);
        cudaFree(s->c[k]);
        cudaFree(s->frame_dev[k]);
        if (s->frame_host[k]) cudaFreeHost(s->frame_host[k]);
        if (s->ev_step[k]) cudaEventDestroy(s->ev_step[k]);
        if (s->ev_copied[k]) cudaEventDestroy(s->ev_copied[k]);
    }
    cudaFree(s->v);
    cudaFree(s->p);
    cudaFree(s->div);
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    ::operator delete(s->drag_storage);
    delete s;
    return FS_OK;
}

int fs_sim_upload(fs_sim *s, const fs_vec2f *v, const fs_rgb_uq32 *c)
{
    if (!s || !v || !c) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(s->ctx->device);
    FS_CUDA_TRY(cudaMemcpyAsync(s->v, v, s->n * sizeof(fs_vec2f), cudaMemcpyDefault, s->ctx->stream));
    FS_CUDA_TRY(cudaMemcpyAsync(s->c[s->steps & 1], c, s->n * sizeof(fs_rgb_uq32), cudaMemcpyDefault, s->ctx->stream));
    return FS_OK;
}

int fs_sim_download(fs_sim *s, fs_vec2f *v, fs_rgb_uq32 *c, float *p, float *div)
{
    if (!s) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(s->ctx->device);
    cudaStream_t st = s->ctx->stream;
    if (v) FS_CUDA_TRY(cudaMemcpyAsync(v, s->v, s->n * sizeof(fs_vec2f), cudaMemcpyDefault, st));
    if (c) FS_CUDA_TRY(cudaMemcpyAsync(c, s->c[s->steps & 1], s->n * sizeof(fs_rgb_uq32), cudaMemcpyDefault, st));
    if (p) FS_CUDA_TRY(cudaMemcpyAsync(p, s->p, s->n * sizeof(float), cudaMemcpyDefault, st));
    if (div) FS_CUDA_TRY(cudaMemcpyAsync(div, s->div, s->n * sizeof(float), cudaMemcpyDefault, st));
    FS_CUDA_TRY(cudaStreamSynchronize(st));
    return FS_OK;
}

int fs_sim_step(fs_sim *s, const fs_drag *drags, int n_drags)
{
    if (!s || n_drags < 0 || (n_drags > 0 && !drags)) return FS_ERR_INVALID_ARG;
    fs_ctx *ctx = s->ctx;
    DeviceGuard guard(ctx->device);
    const int parity = (int)(s->steps & 1);
    if (s->want_frame) {
        if (s->produced - s->consumed >= 2) return FS_ERR_WOULD_BLOCK;   // both frame slots are waiting for the consumer
        // frame slot `parity` was copied out two steps ago: that copy must have finished before it is overwritten
        if (s->produced >= 2) FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, s->ev_copied[parity], 0));
    }
    int e;
    if (s->graph_state == 1 && s->captured_vtmp != ctx->scratch[S_VTMP]) {   // the context re-sized its scratch: capture again
        for (int k = 0; k < 2; k++) {
            cudaGraphExecDestroy(s->exec[k]); s->exec[k] = nullptr;
            cudaGraphDestroy(s->graph[k]); s->graph[k] = nullptr;
        }
        s->graph_state = 0;
    }
    if (s->graph_state == 0 && s->steps >= 2) s->graph_state = capture_graphs(s);   // after two eager steps sized every scratch buffer
    if (s->graph_state == 1 && n_drags <= advect_div_max_drags()) {
        void **kp;
        if ((e = advect_div_graph_params(s->drag_storage, &kp, (float2 *)ctx->scratch[S_VTMP], (const float2 *)s->v, s->div, drags,
                                         n_drags, geo_full(s->dim_x, s->dim_y), s->dt, s->dx)))
            return e;
        cudaKernelNodeParams np = s->drag_node_params[parity];
        np.kernelParams = kp;
        np.extra = nullptr;
        cudaError_t ue = cudaGraphExecKernelNodeSetParams(s->exec[parity], s->drag_node[parity], &np);
        if (ue != cudaSuccess) {
            fprintf(stderr, "[fs_sim] cudaGraphExecKernelNodeSetParams: %s — stepping without the graph\n", cudaGetErrorString(ue));
            cudaGetLastError();
            s->graph_state = -1;
            if ((e = eager_step(s, parity, drags, n_drags))) return e;
            s->eager_steps++;
            goto stepped;
        }
        FS_CUDA_TRY(cudaGraphLaunch(s->exec[parity], ctx->stream));
        s->graph_launches++;
        ctx->launches += s->kernels_per_graph;   // kernels executed (fs_ctx_launch_count), in one graph launch
    } else {
        if ((e = eager_step(s, parity, drags, n_drags))) return e;
        s->eager_steps++;
    }
stepped:
    s->steps++;
    if (s->want_frame) {
        // "color_produced": the frame leaves on the copy stream while the next step computes
        FS_CUDA_TRY(cudaEventRecord(s->ev_step[parity], ctx->stream));
        FS_CUDA_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_step[parity], 0));
        FS_CUDA_TRY(cudaMemcpyAsync(s->frame_host[parity], s->frame_dev[parity], s->frame_px * sizeof(uint16_t),
                                    cudaMemcpyDeviceToHost, s->copy_stream));
        FS_CUDA_TRY(cudaEventRecord(s->ev_copied[parity], s->copy_stream));
        s->produced++;
    }
    return FS_OK;
}

int fs_sim_acquire_frame(fs_sim *s, const uint16_t **frame, int *rows, int *cols)
{
    if (!s || !s->want_frame || !frame) return FS_ERR_INVALID_ARG;
    if (s->consumed == s->produced) return FS_ERR_WOULD_BLOCK;           // nothing produced yet
    DeviceGuard guard(s->ctx->device);
    const int slot = (int)(s->consumed & 1);
    FS_CUDA_TRY(cudaEventSynchronize(s->ev_copied[slot]));
    *frame = s->frame_host[slot];
    if (rows) *rows = 4 * (s->dim_x - 1);
    if (cols) *cols = 4 * (s->dim_y - 1);
    return FS_OK;
}

int fs_sim_release_frame(fs_sim *s)
{
    if (!s || !s->want_frame || s->consumed == s->produced) return FS_ERR_INVALID_ARG;
    s->consumed++;                                                       // "color_consumed"
    return FS_OK;
}

int fs_sim_stats(const fs_sim *s, unsigned long long *steps, unsigned long long *graph_launches, unsigned long long *eager_steps)
{
    if (!s) return FS_ERR_INVALID_ARG;
    if (steps) *steps = s->steps;
    if (graph_launches) *graph_launches = s->graph_launches;
    if (eager_steps) *eager_steps = s->eager_steps;
    return FS_OK;
}

}  // extern "C"
