// CPU execution of the register-tiled ensemble kernel's OWN source (ensemble_reg.cuh): one pthread per
// CUDA thread, pthread_barrier for __syncthreads, a heap block for the shared memory.  Lets the `not gpu`
// test tier check the kernel's thread mapping, mailbox protocol and arithmetic bit for bit against the
// oracle before any GPU time is spent.  TEST INFRASTRUCTURE ONLY.
#include "cuda_host_shim.h"

#include <pthread.h>

#include <vector>

#include "../../esp32-fluid-simulation_b200/csrc/ensemble_reg.cuh"

namespace {

template <bool PIPE>
struct EnvHost {
    static constexpr bool kAsync = false;       // no bulk copies on the host:
    static constexpr bool kEmulatePipe = PIPE;  // ... the pipelined flow makes them cooperatively at the same program points
    int tid, nthreads, block, nblocks;
    pthread_barrier_t *bar;
    void sync() const { pthread_barrier_wait(bar); }
};

struct Job {
    const fs::EnsArgs *a;
    unsigned char *smem;
    int tid, nthreads, block, nblocks;
    pthread_barrier_t *bar;
    int R, dye_smem, pipe;
};

template <int R, bool PIPE>
void run_rp(const Job &j)
{
    const EnvHost<PIPE> env{j.tid, j.nthreads, j.block, j.nblocks, j.bar};
    if (j.dye_smem)
        fs::ens_reg_body<R, true>(*j.a, j.smem, env);
    else
        fs::ens_reg_body<R, false>(*j.a, j.smem, env);
}

template <int R>
void run_r(const Job &j)
{
    if (j.pipe)
        run_rp<R, true>(j);
    else
        run_rp<R, false>(j);
}

void *thread_main(void *arg)
{
    const Job &j = *static_cast<const Job *>(arg);
    switch (j.R) {
        case 2: run_r<2>(j); break;
        case 4: run_r<4>(j); break;
        case 6: run_r<6>(j); break;
        default: run_r<8>(j); break;
    }
    return nullptr;
}

}  // namespace

// v: [batch][dim_y][dim_x] float2, c: [batch][dim_y][dim_x][3] uint32 — updated in place like fs_ensemble_step.
// nblocks "CTAs" walk the batch (run one after the other); threads = 0 picks the launcher's CTA size; pipe = 1 runs
// the pipelined flow of the dye-resident R = 2 kernel (ens_resident_pipelined) with cooperative copies.
extern "C" int ens_emu_step(float *v, uint32_t *c, const fs_drag *drags, const int *counts, int max_drags, int batch,
                            int dim_x, int dim_y, float dt, float dx, int iters, float omega, int n_steps, int R,
                            int dye_smem, int nblocks, int threads, int pipe)
{
    if (R < 2 || R > 8 || R % 2 || dim_x < 2 || dim_y < 2 || batch <= 0 || nblocks <= 0) return -1;
    const int n = dim_x * dim_y;
    fs::EnsArgs a;
    std::vector<uint32_t> scratch((size_t)nblocks * n * 3);
    a.v = reinterpret_cast<float2 *>(v);
    a.c = c;
    a.scratch = scratch.data();
    a.drags = drags;
    a.counts = counts;
    a.max_drags = max_drags; a.batch = batch; a.dim_x = dim_x; a.dim_y = dim_y; a.iters = iters; a.n_steps = n_steps;
    a.dt = dt;
    a.pipe_max_steps = 0x7fffffff;
    a.two_dx_inv = 1.0f / (2.0f * dx);
    a.k = fs::make_sor_coef(dx, omega);
    if (threads <= 0) {
        threads = (fs::ens_reg_blocks(dim_x, dim_y, R) + 31) / 32 * 32;
        if (threads < 64) threads = 64;
    }
    if (threads < fs::ens_reg_blocks(dim_x, dim_y, R)) return -2;
    const size_t smem_bytes = fs::ens_reg_smem_bytes(dim_x, dim_y, R, dye_smem != 0);
    for (int b = 0; b < nblocks && b < batch; b++) {
        // poison the shared memory: nothing may depend on its initial contents
        std::vector<unsigned char> smem(smem_bytes + 64, 0xA5);
        unsigned char *base = smem.data() + ((16 - reinterpret_cast<uintptr_t>(smem.data()) % 16) % 16);
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, threads);
        std::vector<Job> jobs(threads);
        std::vector<pthread_t> th(threads);
        pthread_attr_t attr;
        pthread_attr_init(&attr);
        pthread_attr_setstacksize(&attr, 256 * 1024);
        for (int t = 0; t < threads; t++) {
            jobs[t] = Job{&a, base, t, threads, b, nblocks < batch ? nblocks : batch, &bar, R, dye_smem, pipe};
            if (pthread_create(&th[t], &attr, thread_main, &jobs[t]) != 0) return -3;
        }
        for (int t = 0; t < threads; t++) pthread_join(th[t], nullptr);
        pthread_attr_destroy(&attr);
        pthread_barrier_destroy(&bar);
    }
    return 0;
}
