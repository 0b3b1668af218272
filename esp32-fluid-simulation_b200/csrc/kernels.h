// Host-side launchers of the sm_100a kernels (one per operator variant).  All
// take a Geo (fs_common.cuh) so the same kernels serve a whole grid on one GPU
// and a rank's padded window of a decomposed grid.  Each returns a cudaError_t
// cast to int (0 = ok) and bumps *launches by the number of kernels enqueued.
#pragma once

#include "fs_common.cuh"

namespace fs {

struct Launch {
    cudaStream_t stream;
    uint64_t *launches;  // counter owned by the context
    int num_sms;
};

// advect.cu — advect.h:74-85
int launch_advect_vec2f_gather(const Launch &L, float2 *next_p, const float2 *p, const float2 *vel,
                               const Geo &g, float dt, bool no_slip, int *status);
int launch_advect_rgb_gather(const Launch &L, uint32_t *next_c, const uint32_t *c, const float2 *vel,
                             const Geo &g, float dt, bool no_slip, int *status);

// advect_tma.cu — same operator, source tile staged in shared memory by TMA (needs 16-byte row pitch)
bool advect_vec2f_tma_legal(const float2 *p, const Geo &g);
bool advect_rgb_tma_legal(const uint32_t *c, const Geo &g);
int launch_advect_vec2f_tma(const Launch &L, float2 *next_p, const float2 *p, const float2 *vel,
                            const Geo &g, float dt, bool no_slip, int *status);
int launch_advect_rgb_tma(const Launch &L, uint32_t *next_c, const uint32_t *c, const float2 *vel,
                          const Geo &g, float dt, bool no_slip, int *status);

// fused advect velocity (no-slip) + drags + divergence: the divergence is written on g's compute
// rectangle (every node of it needs its 4 neighbours inside the window or beyond a global wall), the
// forced velocity on store_rect = {x0,y0,x1,y1} in local coordinates (nullptr = the compute rectangle)
int launch_advect_div_tma(const Launch &L, float2 *v_out, const float2 *v_in, float *div, const fs_drag *drags_host,
                          int n_drags, const Geo &g, float dt, float dx, const int *store_rect = nullptr,
                          int *status = nullptr);
int advect_div_max_drags();
// CUDA-graph support (sim.cu): re-arm a captured advect_div_tma_kernel node with a step's drag records
const void *advect_div_kernel_func();
size_t advect_div_params_bytes();
int advect_div_graph_params(void *storage, void ***kernel_params, float2 *v_out, const float2 *v_in, float *div,
                            const fs_drag *drags_host, int n_drags, const Geo &g, float dt, float dx);
// fused dye advect + 4x RGB565 frame of the advected dye (ino:282 + ino:116-177): the frame covers the
// cells that start at the nodes of g's compute rectangle (needs the dye valid one node beyond the
// advect halo); frame = first pixel of the cell at (x0, y0), frame_cells_y = its cell columns
int launch_advect_rgb_frame(const Launch &L, uint32_t *next_c, uint16_t *frame, int frame_cells_y, const uint32_t *c,
                            const float2 *vel, const Geo &g, float dt, bool no_slip, int *status);
int launch_advect_rgb_tma_grad(const Launch &L, uint32_t *next_c, const uint32_t *c, float2 *v_out,
                               const float2 *v_tmp, const float *p, const Geo &g, float dt, float dx,
                               bool no_slip, int *status);

// stencil.cu — finitediff.cpp:9-82, ino:264-269
int launch_divergence(const Launch &L, float *div, const float2 *v, const Geo &g, float dx);
int launch_subtract_gradient(const Launch &L, float2 *v_out, const float2 *v_in, const float *p,
                             const Geo &g, float dx);
int launch_apply_drags(const Launch &L, float2 *v, const fs_drag *drags_host, int n, const Geo &g);
int launch_max_displacement(const Launch &L, unsigned int *out_bits, const float2 *vel, const Geo &g);

// sor.cu — poisson.cpp:14-125
int launch_sor_half_sweep(const Launch &L, float *p, const float *div, const Geo &g, float dx,
                          float omega, int parity);

int launch_sor_residual(const Launch &L, const float *p, const float *div, const Geo &g, float dx,
                        unsigned int *out_max_bits, double *out_sumsq);

// sor_blocked.cu — `n_half` colour half-sweeps per HBM round trip, p_in -> p_out (distinct
// buffers; p_in == nullptr means all zero).  shape 0 = 128x96 region, 2 CTAs/SM; 1 = 128x192, 1 CTA/SM;
// 2/3 = the same regions with persistent CTAs and TMA prefetch (need a ZEROED device work counter).
constexpr int SOR_BLOCKED_MAX_HALF = 16;
int launch_sor_blocked(const Launch &L, float *p_out, const float *p_in, const float *div, const Geo &g,
                       float dx, float omega, int first_parity, int n_half, int shape, int *work_counter);

// The blocked SOR pass FUSED WITH ITS HALO EXCHANGE (decomposed grids): while a tile is written
// back, the parts of it that neighbouring ranks need as ghosts are stored straight into their
// windows over NVLink peer memory; the CTA that finishes the last such ("rim") tile publishes the
// exchange's sequence number in the neighbours' flag slots, and the interior tiles run on
// underneath.  The NEXT pass waits for the neighbours' numbers before its first load.
struct SorPushPeer {
    float *base;             // the neighbour's p_out window (peer-mapped pointer, element (0,0))
    int pitch;               // its window pitch in nodes
    int dx, dy;              // local coordinate here + (dx, dy) = local coordinate there
    int sx0, sy0, sx1, sy1;  // strip of THIS rank's rectangle it needs (local coordinates, x multiples of 4)
};
struct SorPushArgs {
    int n_peers;                            // strips to push (0 = none)
    int n_wait;                             // flags to wait for before the first load (0 = none)
    SorPushPeer peer[8];
    unsigned long long *signal[8];          // flag slots in the neighbours' arenas, one per peer
    unsigned long long *wait[8];            // this rank's flag slots
    unsigned long long seq_signal, seq_wait, timeout_ns;
    int has_l, has_r, has_d, has_u;         // sides of the rectangle that face another rank
    int rim_total;                          // tiles whose region crosses such a side (filled by the launcher)
    int *rim_done;                          // device counter, zeroed before the launch
    int *status;                            // raised to FS_ERR_HALO_TIMEOUT when a neighbour never signals
};
int launch_sor_blocked_push(const Launch &L, float *p_out, const float *p_in, const float *div, const Geo &g,
                            float dx, float omega, int first_parity, int n_half, int shape, int *work_counter,
                            SorPushArgs &push, int grid_limit);

// ensemble.cu — whole loop() body per grid, resident in shared memory
size_t ensemble_smem_bytes(int dim_x, int dim_y, bool dye_smem);
bool ensemble_supported(int dim_x, int dim_y, size_t max_smem_optin);
int ensemble_grid(int batch, int dim_x, int dim_y, int num_sms, int variant);
size_t ensemble_scratch_bytes(int dim_x, int dim_y, int grid);   // per-CTA dye ping-pong slots
int launch_ensemble(const Launch &L, float2 *v, uint32_t *c, uint32_t *scratch, const fs_drag *drags_dev,
                    const int *counts_dev, int max_drags, int batch, int dim_x, int dim_y, float dt,
                    float dx, int iters, float omega, int n_steps, int variant);

// the whole solve in ONE persistent launch with tile-level dependencies between passes; returns -1
// when not eligible (caller falls back to one launch per pass)
size_t sor_solve_flag_count(const Geo &g, int iters, int t_block);
int launch_sor_solve(const Launch &L, float *p, float *scratch, const float *div, const Geo &g, float dx,
                     float omega, int iters, int t_block, int shape, int *work_counter, unsigned int *done,
                     size_t done_capacity, unsigned int gen);

// halo.cu — one-kernel halo exchange over peer (NVLink) memory
constexpr int HALO_MAX_COPIES = 32, HALO_MAX_PEERS = 8;
struct HaloCopy {
    const uint32_t *src;      // this rank's window
    uint32_t *dst;            // a neighbour's ghost region (peer pointer)
    int src_pitch_words, dst_pitch_words, row_words, rows;
    int vec16;                // filled by the launcher: everything is 16-byte aligned
};
struct HaloArgs {
    HaloCopy copies[HALO_MAX_COPIES];
    int block_end[HALO_MAX_COPIES];           // filled by the launcher
    unsigned long long *signal[HALO_MAX_PEERS];  // the neighbours' flag slots for this rank
    unsigned long long *wait[HALO_MAX_PEERS];    // this rank's flag slots, one per neighbour
    unsigned long long seq;
    unsigned long long timeout_ns;               // 0 = wait forever
    int n_copies, n_peers;
};
int launch_halo_exchange(const Launch &L, HaloArgs &a, unsigned int *done_counter, int *status);

// inputs.cu — setup() ino:196-241 and touch_routine() ino:63-96 on the device
int launch_init_color_wheel(const Launch &L, float2 *v, uint32_t *c, int batch, int dim_x, int dim_y);
int launch_touch_to_drags(const Launch &L, fs_drag *drags, int *counts, const int *samples, int n_samples, int batch,
                          int max_drags, int n_rows, int n_cols, const int cal[4], int polling_ms);

// Load every kernel a decomposed step can launch (see FS_PRELOAD below): returns a cudaError_t.
int preload_advect_kernels();
int preload_advect_tma_kernels();
int preload_stencil_kernels();
int preload_sor_kernels();
int preload_sor_blocked_kernels();
int preload_halo_kernels();
int preload_upscale_kernels();
#define FS_PRELOAD(k)                                                        \
    do {                                                                     \
        cudaFuncAttributes fa;                                               \
        cudaError_t pe = cudaFuncGetAttributes(&fa, k);                      \
        if (pe != cudaSuccess) return (int)pe;                               \
    } while (0)

// upscale.cu — ino:116-177
int launch_upscale4_rgb565(const Launch &L, uint16_t *out, const uint32_t *c, int dim_x, int dim_y);

}  // namespace fs
