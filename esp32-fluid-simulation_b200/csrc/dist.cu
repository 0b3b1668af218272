// fs_dist_* — one rank of a block-decomposed grid, the whole decomposed loop()
// body behind the C ABI so that a host harness (C++, not only Python) can drive a multi-GPU step.
//
// Rank r owns a rectangle of the global grid and keeps a padded WINDOW of every field (rectangle +
// `ghost` nodes towards every neighbouring rank, none towards a domain wall) in ONE cudaMalloc'ed
// arena that its neighbours map through CUDA IPC.  All arithmetic runs in the window kernels, which
// evaluate walls, colour parity and advect coordinates in GLOBAL coordinates — the decomposed step
// is bit-identical to the single-GPU step (and to the reference).
//
// One step (ino:249-289), P = number of blocked SOR passes, H = half-sweeps per pass:
//   1. advect v + drags + divergence in ONE kernel on the rectangle grown by D (what the SOR passes
//      recompute redundantly): every rank recomputes the forced velocity of that ring itself instead
//      of exchanging it; the velocity is STORED on the owned rectangle only;
//   2. P blocked SOR passes, each FUSED WITH ITS HALO EXCHANGE (sor_blocked_push_kernel): rim tiles
//      run first and store their results straight into the neighbours' ghosts over NVLink, the last
//      rim tile publishes the exchange's sequence number, interior tiles run on underneath; the next
//      pass waits for the neighbours' numbers before its first load.  The last pass computes the
//      rectangle grown by one node (the gradient's ring), so no exchange follows it;
//   3. gradient-subtract on the rectangle;
//   4. ONE exchange kernel (halo.cu) for the projected velocity (ghost width A + D + 1) and the dye
//      (width A; A = the static advect halo) — the velocity halo also serves the next step's advect;
//   5. dye advect.
// => P exchanges per step (P-1 of them overlapped with SOR tiles) instead of P+2 kernel-separated
// ones, and no NCCL call on the data path (torch.distributed only ships the 64-byte IPC handles).
//
// Why this is race-free (the one-sided protocol, emulated under adversarial schedules in
// tests/test_dist_cpu.py): no kernel of the sequence writes ghost cells of a buffer that a
// neighbour pushes into during the same phase; p ping-pongs between two buffers and a rank enters
// pass k only after every neighbour finished the rim tiles of pass k-1 — the only tiles that read
// the ghosts pass k's pushes overwrite; the velocity/dye pushes of step 4 go into the buffers that
// were last read before the step's SOR hand-shakes.  A neighbour that never signals raises
// FS_ERR_HALO_TIMEOUT (option "halo_timeout_ms") instead of hanging the GPU.
#include <new>
#include <vector>

#include "ctx.h"

namespace {

constexpr size_t FLAG_BYTES = 256;   // 3 sets of 9 uint64 flag slots (one per direction): SOR passes + main-stream exchanges,
                                     // the velocity exchange's side stream, the dye exchange's side stream; padded

struct RankGeo {
    int gx0, gx1, gy0, gy1;   // owned rectangle, global coordinates
    int ox, oy, nx, ny;       // window origin (global) and extents
    int x0, y0, x1, y1;       // owned rectangle, window-local coordinates
};

struct Neighbour {
    int rank, dx, dy;
    char *base;               // its arena as seen from this rank
    bool opened;              // mapped with cudaIpcOpenMemHandle (must be closed)
};

inline int round_up4(int v) { return (v + 3) & ~3; }

// [lo, hi) of part k when n nodes are cut into `parts` pieces whose boundaries are multiples of 4
// (keeps window rows 16-byte aligned) — the same rule as dist.py's split()
inline void split4(int n, int parts, int k, int &lo, int &hi)
{
    auto cut = [&](int i) -> int {
        if (i <= 0) return 0;
        if (i >= parts) return n;
        return (int)(((long long)n * i / parts) / 4 * 4);
    };
    lo = cut(k);
    hi = cut(k + 1);
}

}  // namespace

struct fs_dist {
    fs_ctx *ctx;
    fs_dist_config cfg;
    int px, py, rx, ry;
    std::vector<RankGeo> ranks;
    RankGeo me;
    size_t max_nodes, arena_bytes;
    char *arena;
    size_t off_v[2], off_c[2], off_p[2], off_div;
    int cur_v, cur_c;                  // which buffer holds the current velocity / dye
    int n_nb;
    Neighbour nb[8];
    bool connected;
    unsigned long long seq;            // sequence number of the last exchange (identical on all ranks)
    unsigned long long exchanges;
    // SOR plan
    int T, n_pass;
    int sched[WORK_SLOTS];             // iterations of each SOR pass
    int D, A, vw, cw;                  // div ring, advect halo, exchanged ghost widths of v and dye
    bool v_halo_ok;                    // the current velocity's ghosts are valid to width vw
    float *p_last;
    int has_l, has_r, has_d, has_u;
    uint16_t *frame;                   // this rank's part of the RGB565 frame (cfg.frame), or nullptr
    int cells_x, cells_y;
    cudaEvent_t ev[6];                 // phase boundaries of the LAST step (fs_dist_info: phase_ms)
    bool ev_valid;
    // the two per-step field exchanges run on side streams, off the critical path (see fs_dist_step):
    bool c_halo_ok;                    // the current dye's ghosts are valid to width cw
    cudaStream_t xs_v, xs_c;           // velocity / dye exchange streams (high priority)
    cudaEvent_t ev_grad, ev_vx, ev_dye, ev_cx;
    bool vx_pending, cx_pending;       // an exchange is in flight whose completion the main stream has not waited for yet
    unsigned long long seq_side[3];    // sequence numbers of flag sets 1 and 2 (index 0 unused: set 0 uses `seq`)
    unsigned int *done_side;           // device: last-block counters of the side-stream exchange kernels [2]
};

namespace {

Geo window_geo(const fs_dist *d)
{
    const RankGeo &m = d->me;
    Geo g;
    g.GX = d->cfg.gdim_x; g.GY = d->cfg.gdim_y;
    g.ox = m.ox; g.oy = m.oy; g.nx = m.nx; g.ny = m.ny;
    g.x0 = m.x0; g.y0 = m.y0; g.x1 = m.x1; g.y1 = m.y1;
    g.vx0 = 0; g.vy0 = 0; g.vx1 = m.nx; g.vy1 = m.ny;
    return g;
}

// the owned rectangle grown by r towards neighbouring ranks (clipped to the window; towards a wall
// the window ends at the rectangle)
void grow_rect(const Geo &g, int r, int &x0, int &y0, int &x1, int &y1)
{
    x0 = g.x0 - r < 0 ? 0 : g.x0 - r;
    y0 = g.y0 - r < 0 ? 0 : g.y0 - r;
    x1 = g.x1 + r > g.nx ? g.nx : g.x1 + r;
    y1 = g.y1 + r > g.ny ? g.ny : g.y1 + r;
}

template <class T>
T *field(const fs_dist *d, size_t off) { return reinterpret_cast<T *>(d->arena + off); }

// strip of MY rectangle that neighbour (dx,dy) needs, wx / wy nodes deep (window-local coordinates)
void send_strip(const RankGeo &m, int dx, int dy, int wx, int wy, int &sx0, int &sy0, int &sx1, int &sy1)
{
    sx0 = dx > 0 ? m.x1 - wx : m.x0;
    sx1 = dx < 0 ? m.x0 + wx : m.x1;
    sy0 = dy > 0 ? m.y1 - wy : m.y0;
    sy1 = dy < 0 ? m.y0 + wy : m.y1;
}

unsigned long long *flag_slot(char *arena_base, int dx, int dy, int set = 0)
{
    return reinterpret_cast<unsigned long long *>(arena_base) + set * 9 + ((dy + 1) * 3 + (dx + 1));
}

unsigned long long timeout_ns(const fs_ctx *ctx)
{
    return ctx->opt_halo_timeout_ms > 0 ? (unsigned long long)ctx->opt_halo_timeout_ms * 1000000ull : 0ull;
}

// one stand-alone exchange kernel for up to 2 fields (velocity: 8 B/node, dye: 12 B/node)
// set 0: on the context's stream, flag set / sequence shared with the SOR passes; 1 / 2: on the velocity / dye side
// stream with a flag set, sequence number and last-block counter of its own (their hand-shakes interleave with the
// SOR passes' in an order that differs from rank to rank)
int exchange_fields(fs_dist *d, const size_t *offs, const int *elem_bytes, const int *widths, int n_fields, int set = 0)
{
    if (d->n_nb == 0) return FS_OK;
    fs_ctx *ctx = d->ctx;
    HaloArgs a;
    memset(&a, 0, sizeof(a));
    int nc = 0;
    for (int k = 0; k < d->n_nb; k++) {
        const Neighbour &n = d->nb[k];
        const RankGeo &pr = d->ranks[n.rank];
        for (int f = 0; f < n_fields; f++) {
            int sx0, sy0, sx1, sy1;
            send_strip(d->me, n.dx, n.dy, widths[f], widths[f], sx0, sy0, sx1, sy1);
            const int es = elem_bytes[f];
            const int dxo = d->me.ox - pr.ox, dyo = d->me.oy - pr.oy;
            HaloCopy &c = a.copies[nc++];
            c.src = reinterpret_cast<const uint32_t *>(d->arena + offs[f] + ((size_t)sy0 * d->me.nx + sx0) * es);
            c.dst = reinterpret_cast<uint32_t *>(n.base + offs[f] + ((size_t)(sy0 + dyo) * pr.nx + (sx0 + dxo)) * es);
            c.src_pitch_words = d->me.nx * es / 4;
            c.dst_pitch_words = pr.nx * es / 4;
            c.row_words = (sx1 - sx0) * es / 4;
            c.rows = sy1 - sy0;
        }
        a.signal[k] = flag_slot(n.base, -n.dx, -n.dy, set);
        a.wait[k] = flag_slot(d->arena, n.dx, n.dy, set);
    }
    a.n_copies = nc;
    a.n_peers = d->n_nb;
    a.seq = set == 0 ? ++d->seq : ++d->seq_side[set];
    a.timeout_ns = timeout_ns(ctx);
    d->exchanges++;
    Launch L = mk(ctx);
    if (set == 1) L.stream = d->xs_v;
    if (set == 2) L.stream = d->xs_c;
    return launch_halo_exchange(L, a, set == 0 ? ctx->halo_done_dev : d->done_side + (set - 1), ctx->status_dev);
}

// everything the side streams still have in flight (before the fields are overwritten, read back or freed)
void drain_side_streams(fs_dist *d)
{
    if (d->xs_v) cudaStreamSynchronize(d->xs_v);
    if (d->xs_c) cudaStreamSynchronize(d->xs_c);
    d->vx_pending = d->cx_pending = false;
}

}  // namespace

extern "C" {

int fs_dist_create(fs_dist **out, const fs_dist_config *cfg, fs_ctx *ctx)
{
    if (!ctx) return FS_ERR_NO_CONTEXT;
    if (!out || !cfg) return FS_ERR_INVALID_ARG;
    *out = nullptr;
    if (bad_dims(cfg->gdim_x, cfg->gdim_y) || cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world ||
        cfg->iters < 0 || cfg->ghost < 0 || (cfg->ghost & 3) || cfg->advect_halo < 0)
        return FS_ERR_INVALID_ARG;
    int px = cfg->px, py = cfg->py;
    if (px <= 0 || py <= 0) {   // 1 -> 1x1, 2 -> 1x2, 4 -> 2x2, 8 -> 2x4 (same rule as dist.py)
        px = 1;
        while ((px * 2) * (px * 2) <= cfg->world && cfg->world % (px * 2) == 0) px *= 2;
        py = cfg->world / px;
    }
    if (px * py != cfg->world) return FS_ERR_INVALID_ARG;
    fs_dist *d = new (std::nothrow) fs_dist();
    if (!d) return (int)cudaErrorMemoryAllocation;
    d->ctx = ctx;
    d->cfg = *cfg;
    d->px = px; d->py = py;
    d->cfg.px = px; d->cfg.py = py;
    d->rx = cfg->rank % px; d->ry = cfg->rank / px;
    const int ghost = cfg->world > 1 ? cfg->ghost : 0;
    d->ranks.resize(cfg->world);
    d->max_nodes = 0;
    int min_extent = 0x7fffffff;
    for (int r = 0; r < cfg->world; r++) {
        RankGeo &m = d->ranks[r];
        const int rx = r % px, ry = r / px;
        split4(cfg->gdim_x, px, rx, m.gx0, m.gx1);
        split4(cfg->gdim_y, py, ry, m.gy0, m.gy1);
        const int gl = rx > 0 ? ghost : 0, gr = rx < px - 1 ? ghost : 0;
        const int gd = ry > 0 ? ghost : 0, gu = ry < py - 1 ? ghost : 0;
        m.ox = m.gx0 - gl; m.oy = m.gy0 - gd;
        m.nx = (m.gx1 - m.gx0) + gl + gr; m.ny = (m.gy1 - m.gy0) + gd + gu;
        m.x0 = gl; m.y0 = gd; m.x1 = gl + (m.gx1 - m.gx0); m.y1 = gd + (m.gy1 - m.gy0);
        if ((size_t)m.nx * m.ny > d->max_nodes) d->max_nodes = (size_t)m.nx * m.ny;
        if (px > 1 && m.gx1 - m.gx0 < min_extent) min_extent = m.gx1 - m.gx0;
        if (py > 1 && m.gy1 - m.gy0 < min_extent) min_extent = m.gy1 - m.gy0;
    }
    d->me = d->ranks[cfg->rank];
    // the SOR plan: T iterations per pass; the LAST pass also computes the gradient's ring
    d->T = ctx->opt_sor_t < 1 ? 1 : ctx->opt_sor_t;
    d->A = cfg->world > 1 ? cfg->advect_halo : 0;
    // Pass schedule: T iterations per pass.  A remainder of 1 or 2 iterations is folded into the FIRST pass when the
    // ghosts allow it (that pass starts from p = 0: it needs no pressure ghosts, only the divergence on a wider
    // ring) — one pass and one hand-shake less per step (K = 50, T = 6: 8 + 7 x 6 instead of 8 x 6 + 2; same rule
    // as core_poisson_solve, api.cu).
    for (int fold = 1; fold >= 0; fold--) {
        int n = cfg->iters > 0 ? (cfg->iters + d->T - 1) / d->T : 0;
        if (n > WORK_SLOTS) {
            delete d;
            return FS_ERR_INVALID_ARG;
        }
        for (int k = 0; k < n; k++) d->sched[k] = cfg->iters - k * d->T < d->T ? cfg->iters - k * d->T : d->T;
        if (fold) {
            const int r = n > 1 ? d->sched[n - 1] : 0;
            if (n <= 1 || r > 2 || 2 * (d->T + r) > SOR_BLOCKED_MAX_HALF) continue;
            d->sched[0] = d->T + r;
            n--;
        }
        d->n_pass = n;
        int need_d = 0;
        for (int k = 0; k < n; k++) {
            const int need = 2 * d->sched[k] + (k == n - 1 ? 1 : 0);
            if (need > need_d) need_d = need;
        }
        if (cfg->world == 1) need_d = 0;
        d->D = round_up4(need_d);
        d->vw = round_up4(d->D + 1 + d->A);
        if (!fold || cfg->world == 1 || d->vw <= ghost) break;   // folded plan does not fit the ghosts: plain plan
    }
    d->cw = round_up4(d->A + (cfg->frame && cfg->world > 1 ? 1 : 0));   // the frame's far corners lie one node beyond
    // identical verdict on every rank: the widest exchanged strip must fit the ghosts AND the
    // narrowest rectangle of the decomposition (a strip is cut out of the sender's rectangle)
    if (cfg->world > 1 && (d->vw > ghost || d->cw > ghost || ghost > min_extent || d->n_pass > WORK_SLOTS)) {
        delete d;
        return FS_ERR_INVALID_ARG;
    }
    d->has_l = d->rx > 0; d->has_r = d->rx < px - 1; d->has_d = d->ry > 0; d->has_u = d->ry < py - 1;
    d->n_nb = 0;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            if (!dx && !dy) continue;
            const int nx_ = d->rx + dx, ny_ = d->ry + dy;
            if (nx_ < 0 || nx_ >= px || ny_ < 0 || ny_ >= py) continue;
            Neighbour &n = d->nb[d->n_nb++];
            n.rank = ny_ * px + nx_; n.dx = dx; n.dy = dy; n.base = nullptr; n.opened = false;
        }
    // arena: flags, then every field sized for the LARGEST window so offsets agree on all ranks
    auto per = [&](size_t b) { return (d->max_nodes * b + 255) / 256 * 256; };
    size_t cur = FLAG_BYTES;
    for (int i = 0; i < 2; i++) { d->off_v[i] = cur; cur += per(8); }
    for (int i = 0; i < 2; i++) { d->off_c[i] = cur; cur += per(12); }
    for (int i = 0; i < 2; i++) { d->off_p[i] = cur; cur += per(4); }
    d->off_div = cur; cur += per(4);
    d->arena_bytes = cur;
    DeviceGuard guard(ctx->device);
    cudaError_t e = cudaMalloc(&d->arena, d->arena_bytes);   // cudaMalloc (not a pool): exportable through CUDA IPC
    if (e == cudaSuccess) e = cudaMemset(d->arena, 0, d->arena_bytes);
    if (e != cudaSuccess) {
        delete d;
        return (int)e;
    }
    // Load every kernel of the step NOW: the first launch of a lazily loaded kernel may wait for the
    // device to go idle, which never happens while a resident kernel spins on a neighbour's flag.
    {
        int pe;
        if ((pe = preload_advect_kernels()) || (pe = preload_advect_tma_kernels()) || (pe = preload_stencil_kernels()) ||
            (pe = preload_sor_kernels()) || (pe = preload_sor_blocked_kernels()) || (pe = preload_halo_kernels()) ||
            (pe = preload_upscale_kernels())) {
            cudaFree(d->arena);
            delete d;
            return pe;
        }
    }
    if (!(ctx->opt_fuse & 1)) {
        // the one-kernel-per-operator path advects into a scratch window: allocate it now — a
        // cudaMalloc inside a step would wait for kernels that may themselves be waiting for this rank
        void *scratch;
        int e2 = ensure(ctx, S_VTMP, (size_t)d->me.nx * d->me.ny * sizeof(fs_vec2f), &scratch);
        if (e2) {
            cudaFree(d->arena);
            delete d;
            return e2;
        }
    }
    d->frame = nullptr;
    d->cells_x = (d->me.gx1 < cfg->gdim_x - 1 ? d->me.gx1 : cfg->gdim_x - 1) - d->me.gx0;
    d->cells_y = (d->me.gy1 < cfg->gdim_y - 1 ? d->me.gy1 : cfg->gdim_y - 1) - d->me.gy0;
    if (cfg->frame) {
        if (d->me.nx < 64 || d->me.ny < 32 || d->cells_x < 1 || d->cells_y < 1) {   // the fused kernel's tile
            cudaFree(d->arena);
            delete d;
            return FS_ERR_UNSUPPORTED;
        }
        e = cudaMalloc(&d->frame, (size_t)16 * d->cells_x * d->cells_y * sizeof(uint16_t));
        if (e != cudaSuccess) {
            cudaFree(d->arena);
            delete d;
            return (int)e;
        }
    }
    for (int k = 0; k < 6; k++) cudaEventCreate(&d->ev[k]);
    d->ev_valid = false;
    d->xs_v = d->xs_c = nullptr;
    d->done_side = nullptr;
    d->vx_pending = d->cx_pending = false;
    d->c_halo_ok = false;
    d->seq_side[0] = d->seq_side[1] = d->seq_side[2] = 0;
    if (d->n_nb > 0) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);   // (their few blocks should not queue behind a step's kernels)
        cudaError_t se = cudaStreamCreateWithPriority(&d->xs_v, cudaStreamNonBlocking, prio_hi);
        if (se == cudaSuccess) se = cudaStreamCreateWithPriority(&d->xs_c, cudaStreamNonBlocking, prio_hi);
        if (se == cudaSuccess) se = cudaMalloc(&d->done_side, 2 * sizeof(unsigned int));
        if (se == cudaSuccess) se = cudaMemset(d->done_side, 0, 2 * sizeof(unsigned int));
        if (se == cudaSuccess) se = cudaEventCreateWithFlags(&d->ev_grad, cudaEventDisableTiming);
        if (se == cudaSuccess) se = cudaEventCreateWithFlags(&d->ev_vx, cudaEventDisableTiming);
        if (se == cudaSuccess) se = cudaEventCreateWithFlags(&d->ev_dye, cudaEventDisableTiming);
        if (se == cudaSuccess) se = cudaEventCreateWithFlags(&d->ev_cx, cudaEventDisableTiming);
        if (se != cudaSuccess) {
            for (int k = 0; k < 6; k++) cudaEventDestroy(d->ev[k]);
            if (d->xs_v) cudaStreamDestroy(d->xs_v);
            if (d->xs_c) cudaStreamDestroy(d->xs_c);
            cudaFree(d->done_side);
            cudaFree(d->arena);
            cudaFree(d->frame);
            delete d;
            return (int)se;
        }
    }
    d->cur_v = d->cur_c = 0;
    d->connected = d->n_nb == 0;
    d->seq = 0;
    d->exchanges = 0;
    d->v_halo_ok = false;
    d->p_last = field<float>(d, d->off_p[0]);
    *out = d;
    return FS_OK;
}

int fs_dist_destroy(fs_dist *d)
{
    if (!d) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    drain_side_streams(d);
    for (int k = 0; k < d->n_nb; k++)
        if (d->nb[k].opened) cudaIpcCloseMemHandle(d->nb[k].base);
    cudaFree(d->arena);
    cudaFree(d->frame);
    for (int k = 0; k < 6; k++) cudaEventDestroy(d->ev[k]);
    if (d->xs_v) {
        cudaStreamDestroy(d->xs_v);
        cudaStreamDestroy(d->xs_c);
        cudaFree(d->done_side);
        cudaEventDestroy(d->ev_grad); cudaEventDestroy(d->ev_vx); cudaEventDestroy(d->ev_dye); cudaEventDestroy(d->ev_cx);
    }
    delete d;
    return FS_OK;
}

int fs_dist_frame(fs_dist *d, uint16_t **frame, int *rows, int *cols)
{
    if (!d || !d->frame) return FS_ERR_INVALID_ARG;
    if (frame) *frame = d->frame;
    if (rows) *rows = 4 * d->cells_x;
    if (cols) *cols = 4 * d->cells_y;
    return FS_OK;
}

int fs_dist_window(const fs_dist *d, fs_tile *t)
{
    if (!d || !t) return FS_ERR_INVALID_ARG;
    const RankGeo &m = d->me;
    t->gdim_x = d->cfg.gdim_x; t->gdim_y = d->cfg.gdim_y;
    t->ox = m.ox; t->oy = m.oy; t->nx = m.nx; t->ny = m.ny;
    t->x0 = m.x0; t->y0 = m.y0; t->x1 = m.x1; t->y1 = m.y1;
    return FS_OK;
}

int fs_dist_info(const fs_dist *d, fs_dist_info_t *info)
{
    if (!d || !info) return FS_ERR_INVALID_ARG;
    info->px = d->px; info->py = d->py;
    info->n_neighbours = d->n_nb;
    info->sor_passes = d->n_pass;
    info->sor_t = d->T;
    info->div_ring = d->D;
    info->velocity_halo = d->vw;
    info->dye_halo = d->cw;
    // hand-shakes per step: one per SOR pass but the last (fused into the passes), the velocity exchange and the dye
    // exchange (side streams)
    info->exchanges_per_step = d->n_nb ? (d->n_pass > 0 ? d->n_pass - 1 : 0) + 1 + (d->cw > 0 ? 1 : 0) : 0;
    info->exchanges = d->exchanges;
    info->arena_bytes = d->arena_bytes;
    // phases of the last step (CUDA events on the compute stream): advect+drags+div | SOR passes (with their fused
    // exchanges) | gradient | velocity+dye exchange | dye advect (+frame)
    for (int k = 0; k < 5; k++) info->phase_ms[k] = -1.0f;
    if (d->ev_valid && cudaEventSynchronize(d->ev[5]) == cudaSuccess)
        for (int k = 0; k < 5; k++) cudaEventElapsedTime(&info->phase_ms[k], d->ev[k], d->ev[k + 1]);
    return FS_OK;
}

int fs_dist_ipc_handle(fs_dist *d, unsigned char handle[64])
{
    if (!d || !handle) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(d->ctx->device);
    cudaIpcMemHandle_t h;
    FS_CUDA_TRY(cudaIpcGetMemHandle(&h, d->arena));
    memcpy(handle, &h, 64);
    return FS_OK;
}

int fs_dist_connect(fs_dist *d, const unsigned char *handles)
{
    if (!d || (!handles && d->n_nb > 0)) return FS_ERR_INVALID_ARG;
    DeviceGuard guard(d->ctx->device);
    for (int k = 0; k < d->n_nb; k++) {
        Neighbour &n = d->nb[k];
        if (n.base) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)n.rank * 64, 64);
        void *ptr = nullptr;
        FS_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        n.base = (char *)ptr;
        n.opened = true;
    }
    d->connected = true;
    return FS_OK;
}

int fs_dist_connect_local(fs_dist *d, fs_dist *const *all)
{
    if (!d || (!all && d->n_nb > 0)) return FS_ERR_INVALID_ARG;
    for (int k = 0; k < d->n_nb; k++) {
        fs_dist *o = all[d->nb[k].rank];
        if (!o || o->arena_bytes != d->arena_bytes || o->cfg.rank != d->nb[k].rank) return FS_ERR_INVALID_ARG;
        d->nb[k].base = o->arena;   // same process: the pointer itself (peer access is the caller's business)
        d->nb[k].opened = false;
    }
    d->connected = true;
    return FS_OK;
}

int fs_dist_upload(fs_dist *d, const fs_vec2f *v_window, const fs_rgb_uq32 *c_window)
{
    if (!d || !v_window || !c_window) return FS_ERR_INVALID_ARG;
    fs_ctx *ctx = d->ctx;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)d->me.nx * d->me.ny;
    drain_side_streams(d);   // (this rank's pushes of an earlier run; the caller keeps the ranks in step around an upload)
    FS_CUDA_TRY(cudaMemcpyAsync(d->arena + d->off_v[d->cur_v], v_window, n * sizeof(fs_vec2f), cudaMemcpyDefault, ctx->stream));
    FS_CUDA_TRY(cudaMemcpyAsync(d->arena + d->off_c[d->cur_c], c_window, n * sizeof(fs_rgb_uq32), cudaMemcpyDefault, ctx->stream));
    d->v_halo_ok = d->c_halo_ok = false;   // ghosts are refreshed by an exchange before they are read
    return FS_OK;
}

int fs_dist_download(fs_dist *d, fs_vec2f *v_rect, fs_rgb_uq32 *c_rect, float *p_rect, float *div_rect)
{
    if (!d) return FS_ERR_INVALID_ARG;
    fs_ctx *ctx = d->ctx;
    DeviceGuard guard(ctx->device);
    const RankGeo &m = d->me;
    const size_t w = m.x1 - m.x0, h = m.y1 - m.y0, first = (size_t)m.y0 * m.nx + m.x0;
    auto pull = [&](void *dst, const char *src, size_t es) -> cudaError_t {
        return cudaMemcpy2DAsync(dst, w * es, src + first * es, (size_t)m.nx * es, w * es, h, cudaMemcpyDefault, ctx->stream);
    };
    if (v_rect) FS_CUDA_TRY(pull(v_rect, d->arena + d->off_v[d->cur_v], 8));
    if (c_rect) FS_CUDA_TRY(pull(c_rect, d->arena + d->off_c[d->cur_c], 12));
    if (p_rect) FS_CUDA_TRY(pull(p_rect, (const char *)d->p_last, 4));
    if (div_rect) FS_CUDA_TRY(pull(div_rect, d->arena + d->off_div, 4));
    FS_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    drain_side_streams(d);   // after a download on every rank, nobody pushes into anybody's arena any more
    return FS_OK;
}

int fs_dist_device_fields(fs_dist *d, fs_vec2f **v, fs_rgb_uq32 **c, float **p, float **div)
{
    if (!d) return FS_ERR_INVALID_ARG;
    if (v) *v = field<fs_vec2f>(d, d->off_v[d->cur_v]);
    if (c) *c = field<fs_rgb_uq32>(d, d->off_c[d->cur_c]);
    if (p) *p = d->p_last;
    if (div) *div = field<float>(d, d->off_div);
    return FS_OK;
}

int fs_dist_step(fs_dist *d, const fs_drag *drags, int n_drags)
{
    if (!d) return FS_ERR_INVALID_ARG;
    if (n_drags < 0 || (n_drags > 0 && !drags)) return FS_ERR_INVALID_ARG;
    if (!d->connected) return FS_ERR_INVALID_ARG;
    fs_ctx *ctx = d->ctx;
    DeviceGuard guard(ctx->device);
    const fs_dist_config &cfg = d->cfg;
    const Geo gw = window_geo(d);
    const bool multi = d->n_nb > 0;
    int e;

    fs_vec2f *v = field<fs_vec2f>(d, d->off_v[d->cur_v]), *v2 = field<fs_vec2f>(d, d->off_v[d->cur_v ^ 1]);
    fs_rgb_uq32 *c = field<fs_rgb_uq32>(d, d->off_c[d->cur_c]), *c2 = field<fs_rgb_uq32>(d, d->off_c[d->cur_c ^ 1]);
    float *div = field<float>(d, d->off_div);

    // the current velocity's and dye's ghosts (fresh state only: afterwards the previous step's side-stream exchanges
    // filled them)
    if (multi && (!d->v_halo_ok || !d->c_halo_ok)) {
        const size_t offs[2] = {d->off_v[d->cur_v], d->off_c[d->cur_c]};
        const int es[2] = {8, 12}, ws[2] = {d->vw, d->cw};
        if ((e = exchange_fields(d, offs, es, ws, d->cw > 0 ? 2 : 1))) return e;
    }
    d->v_halo_ok = d->c_halo_ok = true;
    // the velocity ghosts this step's advect gathers from: pushed by the neighbours under the previous step's dye advect
    if (d->vx_pending) {
        FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, d->ev_vx, 0));
        d->vx_pending = false;
    }
    cudaEventRecord(d->ev[0], ctx->stream);

    // ---- 1. advect v (no-slip) + drags + divergence (ino:253, 264-269, 274) -------------------------
    Geo gd = gw;                                   // divergence rectangle: owned, grown by D
    grow_rect(gw, d->D, gd.x0, gd.y0, gd.x1, gd.y1);
    grow_rect(gw, d->vw, gd.vx0, gd.vy0, gd.vx1, gd.vy1);   // velocity ghosts are valid to width vw
    const int own[4] = {gw.x0, gw.y0, gw.x1, gw.y1};
    const fs_vec2f *v_forced;                      // where the forced velocity of the rectangle lives
    if ((ctx->opt_fuse & 1) && ctx->opt_advect == 1 && n_drags <= advect_div_max_drags() &&
        advect_vec2f_tma_legal((const float2 *)v, gd)) {
        if ((e = launch_advect_div_tma(mk(ctx), (float2 *)v2, (const float2 *)v, div, drags, n_drags, gd, cfg.dt, cfg.dx,
                                       own, ctx->status_dev)))
            return e;
        v_forced = v2;
    } else {
        // one kernel per operator, through a scratch window: no compute kernel may write ghost cells
        // of v2 (the neighbours push into them in step 4)
        void *scratch;
        if ((e = ensure(ctx, S_VTMP, (size_t)gw.nx * gw.ny * sizeof(fs_vec2f), &scratch))) return e;
        Geo ga = gd;                               // advect on the divergence rectangle + its 1-node ring
        grow_rect(gw, d->D + (multi ? 1 : 0), ga.x0, ga.y0, ga.x1, ga.y1);
        if ((e = core_advect_vec2f(ctx, (fs_vec2f *)scratch, v, v, ga, cfg.dt, 1, ctx->status_dev))) return e;
        if (n_drags > 0 && (e = launch_apply_drags(mk(ctx), (float2 *)scratch, drags, n_drags, ga))) return e;
        if ((e = launch_divergence(mk(ctx), div, (const float2 *)scratch, gd, cfg.dx))) return e;
        v_forced = (const fs_vec2f *)scratch;
    }

    cudaEventRecord(d->ev[1], ctx->stream);
    // ---- 2. SOR (ino:275): P blocked passes, each fused with its halo exchange ---------------------------
    float *bufs[2] = {field<float>(d, d->off_p[0]), field<float>(d, d->off_p[1])};
    if (d->n_pass == 0) {
        FS_CUDA_TRY(cudaMemsetAsync(bufs[0], 0, (size_t)gw.nx * gw.ny * sizeof(float), ctx->stream));
        d->p_last = bufs[0];
    } else {
        FS_CUDA_TRY(cudaMemsetAsync(ctx->work_dev, 0, WORK_SLOTS * sizeof(int), ctx->stream));
        FS_CUDA_TRY(cudaMemsetAsync(ctx->rim_dev, 0, WORK_SLOTS * sizeof(int), ctx->stream));
        const unsigned long long base = d->seq;
        for (int k = 0; k < d->n_pass; k++) {
            const bool last = k == d->n_pass - 1;
            const int t = d->sched[k];
            Geo gp = gw;
            if (last && multi) grow_rect(gw, 1, gp.x0, gp.y0, gp.x1, gp.y1);   // + the gradient's ring
            SorPushArgs push;
            memset(&push, 0, sizeof(push));
            push.has_l = d->has_l; push.has_r = d->has_r; push.has_d = d->has_d; push.has_u = d->has_u;
            push.timeout_ns = timeout_ns(ctx);
            push.status = ctx->status_dev;
            push.rim_done = ctx->rim_dev + k;
            if (multi && k >= 1) {                 // p_in's ghosts come from the neighbours' pass k-1
                push.n_wait = d->n_nb;
                push.seq_wait = base + k;
                for (int q = 0; q < d->n_nb; q++) push.wait[q] = flag_slot(d->arena, d->nb[q].dx, d->nb[q].dy);
            }
            if (multi && !last) {                  // the next pass needs this one's rim in the neighbours' ghosts
                const int t_next = d->sched[k + 1];
                const int need = 2 * t_next + (k + 1 == d->n_pass - 1 ? 1 : 0);
                push.n_peers = d->n_nb;
                push.seq_signal = base + k + 1;
                for (int q = 0; q < d->n_nb; q++) {
                    const Neighbour &n = d->nb[q];
                    const RankGeo &pr = d->ranks[n.rank];
                    SorPushPeer &pp = push.peer[q];
                    send_strip(d->me, n.dx, n.dy, round_up4(need), need, pp.sx0, pp.sy0, pp.sx1, pp.sy1);
                    pp.base = reinterpret_cast<float *>(n.base + d->off_p[k & 1]);
                    pp.pitch = pr.nx;
                    pp.dx = d->me.ox - pr.ox;
                    pp.dy = d->me.oy - pr.oy;
                    push.signal[q] = flag_slot(n.base, -n.dx, -n.dy);
                }
            }
            const int shape = (ctx->opt_sor_shape == 3 || ctx->opt_sor_shape == 5) ? ctx->opt_sor_shape : 7;
            if ((e = launch_sor_blocked_push(mk(ctx), bufs[k & 1], k ? bufs[(k - 1) & 1] : nullptr, div, gp, cfg.dx,
                                             cfg.omega, 0, 2 * t, shape, ctx->work_dev + k, push, ctx->opt_sor_grid_limit)))
                return e > 0 ? e : FS_ERR_UNSUPPORTED;
        }
        if (multi) {
            d->seq = base + (d->n_pass - 1);
            d->exchanges += d->n_pass - 1;
        }
        d->p_last = bufs[(d->n_pass - 1) & 1];
    }

    cudaEventRecord(d->ev[2], ctx->stream);
    // ---- 3. gradient-subtract (ino:276) on the rectangle: v2 = v_forced - grad p ---------------------------------
    if ((e = launch_subtract_gradient(mk(ctx), (float2 *)v2, (const float2 *)v_forced, d->p_last, gw, cfg.dx))) return e;

    cudaEventRecord(d->ev[3], ctx->stream);
    // ---- 4. halos of the projected velocity and of the dye, OFF the critical path ---------------------------------
    // The dye advect gathers DYE from the ghosts but (without a frame) starts its backtraces from owned nodes only, and the next step's
    // velocity advect gathers VELOCITY: so the projected velocity's strips travel on a side stream under this step's
    // dye advect (the next step waits for them), and the new dye's strips under the NEXT step's advect + SOR (its dye
    // advect waits for them).  Each side stream has a flag set of its own.  Race freedom as in DESIGN.md 5: a neighbour
    // reads the ghosts of buffer b only after its own exchange kernel of the same step saw this rank's flag, and read
    // them last two steps earlier — before the SOR hand-shakes that this rank's push comes after.
    if (multi) {
        FS_CUDA_TRY(cudaEventRecord(d->ev_grad, ctx->stream));
        FS_CUDA_TRY(cudaStreamWaitEvent(d->xs_v, d->ev_grad, 0));
        const size_t offs[1] = {d->off_v[d->cur_v ^ 1]};
        const int es[1] = {8}, ws[1] = {d->vw};
        if ((e = exchange_fields(d, offs, es, ws, 1, 1))) return e;
        FS_CUDA_TRY(cudaEventRecord(d->ev_vx, d->xs_v));
        d->vx_pending = true;
        if (d->frame) {                // the frame's far corners are advected from nodes one beyond the rectangle: their
            FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, d->ev_vx, 0));   // backtraces start from velocity GHOSTS
            d->vx_pending = false;
        }
        if (d->cx_pending) {           // this step's dye ghosts: pushed under this step's advect + SOR
            FS_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, d->ev_cx, 0));
            d->cx_pending = false;
        }
    }

    cudaEventRecord(d->ev[4], ctx->stream);
    // ---- 5. advect dye (ino:282, free-slip sampling) with the projected velocity -------------------------------
    Geo gc = gw;
    grow_rect(gw, d->cw, gc.vx0, gc.vy0, gc.vx1, gc.vy1);   // dye ghosts are valid to width cw
    if (d->frame) {   // + this rank's part of the 4x RGB565 frame, rendered from the tile in shared memory (ino:116-177)
        if (!advect_rgb_tma_legal((const uint32_t *)c, gc)) return FS_ERR_UNSUPPORTED;
        if ((e = launch_advect_rgb_frame(mk(ctx), (uint32_t *)c2, d->frame, d->cells_y, (const uint32_t *)c, (const float2 *)v2,
                                         gc, cfg.dt, false, ctx->status_dev)))
            return e;
    } else if ((e = core_advect_rgb(ctx, c2, c, v2, gc, cfg.dt, 0, ctx->status_dev))) {
        return e;
    }

    cudaEventRecord(d->ev[5], ctx->stream);
    d->ev_valid = true;
    if (multi && d->cw > 0) {          // the new dye's strips: under the next step's advect + SOR
        FS_CUDA_TRY(cudaEventRecord(d->ev_dye, ctx->stream));
        FS_CUDA_TRY(cudaStreamWaitEvent(d->xs_c, d->ev_dye, 0));
        const size_t offs[1] = {d->off_c[d->cur_c ^ 1]};
        const int es[1] = {12}, ws[1] = {d->cw};
        if ((e = exchange_fields(d, offs, es, ws, 1, 2))) return e;
        FS_CUDA_TRY(cudaEventRecord(d->ev_cx, d->xs_c));
        d->cx_pending = true;
    }
    d->cur_v ^= 1;    // ino:255 / ino:286: the pointer swaps
    d->cur_c ^= 1;
    return FS_OK;
}

int fs_dist_check(fs_dist *d)
{
    if (!d) return FS_ERR_INVALID_ARG;
    {
        DeviceGuard guard(d->ctx->device);
        drain_side_streams(d);     // a time-out of a side-stream hand-shake is reported too
    }
    return fs_tile_check(d->ctx);
}

}  // extern "C"
