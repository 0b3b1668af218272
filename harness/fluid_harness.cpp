// fluid_harness — host C++ harness that runs the reference's step order (loop(),
// ESP32-fluid-simulation.ino:249-289) through a table of operator function pointers and
// lets every operator be bound, per call site, either to a CPU library (the reference's
// own sources behind oracle/_ref, or the oracle restatement) or to the CUDA library
// (fsh_* entry points of include/fluid_b200.h).  TEST / MEASUREMENT TOOL: it loads the
// CPU library as a checker, never as part of the product.
//
//   fluid_harness --gpu-lib <libfluid_b200.so> --cpu-lib <libfluid_ref.so> --cpu-prefix ref_
//                 [--dim-x 61 --dim-y 81 --steps 20 --iters 10 --drags 4]
//                 [--gpu-ops advect_v,drags,divergence,poisson,gradient,advect_c | all | none]
//                 [--decomposed WORLD]
//
// --decomposed WORLD adds a third run: the same steps through fs_dist_* (SURVEY.md 8b), WORLD ranks of
// a block decomposition emulated on device 0 (one context + stream per rank, fs_dist_connect_local),
// i.e. the C++-sequenced multi-GPU step with its fused halo exchanges, driven from C++.
//
// Runs the sequence twice — once all-CPU, once with the selected operators on the GPU —
// on the same seeded inputs, prints per-operator wall times and FNV-1a hashes of the
// final fields, and exits 0 only if the two runs agree bit for bit.
#include <dlfcn.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fluid_b200.h"

namespace {

struct Ops {  // the reference's operator surface as plain function pointers (host pointers)
    void (*advect_v)(float *, float *, float *, int, int, float, int) = nullptr;
    void (*advect_c)(uint32_t *, uint32_t *, float *, int, int, float, int) = nullptr;
    void (*divergence)(float *, float *, int, int, float) = nullptr;
    void (*gradient)(float *, float *, int, int, float) = nullptr;
    void (*poisson)(float *, float *, int, int, float, int, float) = nullptr;
    void (*drags)(float *, const void *, int, int, int) = nullptr;
};

void *must_sym(void *lib, const std::string &name)
{
    void *s = dlsym(lib, name.c_str());
    if (!s) {
        std::fprintf(stderr, "missing symbol %s\n", name.c_str());
        std::exit(2);
    }
    return s;
}

// --- GPU binding: adapters from the reference-shaped pointers to the C ABI --------------------
struct Gpu {
    fs_ctx *ctx = nullptr;
    decltype(&fsh_advect_vec2f) advect_vec2f;
    decltype(&fsh_advect_rgb_uq32) advect_rgb;
    decltype(&fsh_calculate_divergence) divergence;
    decltype(&fsh_subtract_gradient) gradient;
    decltype(&fsh_poisson_solve) poisson;
    decltype(&fs_apply_drags) apply_drags_dev;
    decltype(&fs_error_string) err;
} G;

void gcheck(int code, const char *what)
{
    if (code != FS_OK) {
        std::fprintf(stderr, "%s failed: %s\n", what, G.err(code));
        std::exit(3);
    }
}
void g_advect_v(float *n, float *p, float *v, int x, int y, float dt, int ns)
{
    gcheck(G.advect_vec2f((fs_vec2f *)n, (fs_vec2f *)p, (fs_vec2f *)v, x, y, dt, ns, G.ctx), "fsh_advect_vec2f");
}
void g_advect_c(uint32_t *n, uint32_t *c, float *v, int x, int y, float dt, int ns)
{
    gcheck(G.advect_rgb((fs_rgb_uq32 *)n, (fs_rgb_uq32 *)c, (fs_vec2f *)v, x, y, dt, ns, G.ctx), "fsh_advect_rgb_uq32");
}
void g_div(float *d, float *v, int x, int y, float dx) { gcheck(G.divergence(d, (fs_vec2f *)v, x, y, dx, G.ctx), "fsh_calculate_divergence"); }
void g_grad(float *v, float *p, int x, int y, float dx) { gcheck(G.gradient((fs_vec2f *)v, p, x, y, dx, G.ctx), "fsh_subtract_gradient"); }
void g_pois(float *p, float *d, int x, int y, float dx, int k, float w) { gcheck(G.poisson(p, d, x, y, dx, k, w, G.ctx), "fsh_poisson_solve"); }

uint64_t splitmix(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
uint64_t fnv(const void *d, size_t n)
{
    const unsigned char *b = (const unsigned char *)d;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
    return h;
}

struct Drag { uint16_t cx, cy; float vx, vy; };

std::vector<Drag> step_drags(int st, int n_drags, int dim_x, int dim_y)
{
    std::vector<Drag> dr(n_drags);
    for (int k = 0; k < n_drags; k++) {
        uint64_t h = splitmix(0xD4A6 + (uint64_t)st * 1000 + k);
        dr[k].cx = (uint16_t)(h % dim_y);
        dr[k].cy = (uint16_t)((h >> 20) % dim_x);
        dr[k].vx = (float)((int)((h >> 40) & 0x3ff) - 512);
        dr[k].vy = (float)((int)((h >> 50) & 0x3ff) - 512);
    }
    return dr;
}

struct State {
    std::vector<float> v, p, d;
    std::vector<uint32_t> c;
};

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// loop(), ino:249-289, through the operator table
void run(const Ops &o, State &s, int dim_x, int dim_y, int steps, int iters, int n_drags, double *t_op)
{
    const size_t n = (size_t)dim_x * dim_y;
    std::vector<float> v_tmp(2 * n);
    std::vector<uint32_t> c_tmp(3 * n);
    const float dt = 1 / 30.0f;
    for (int st = 0; st < steps; st++) {
        double t0 = now();
        o.advect_v(v_tmp.data(), s.v.data(), s.v.data(), dim_x, dim_y, dt, 1);             // ino:253
        s.v.swap(v_tmp);                                                                  // ino:255
        double t1 = now();
        std::vector<Drag> dr = step_drags(st, n_drags, dim_x, dim_y);
        o.drags(s.v.data(), dr.data(), n_drags, dim_x, dim_y);                            // ino:264-269
        double t2 = now();
        o.divergence(s.d.data(), s.v.data(), dim_x, dim_y, 1.0f);                          // ino:274
        double t3 = now();
        o.poisson(s.p.data(), s.d.data(), dim_x, dim_y, 1.0f, iters, 1.96f);               // ino:275
        double t4 = now();
        o.gradient(s.v.data(), s.p.data(), dim_x, dim_y, 1.0f);                            // ino:276
        double t5 = now();
        o.advect_c(c_tmp.data(), s.c.data(), s.v.data(), dim_x, dim_y, dt, 0);             // ino:282
        s.c.swap(c_tmp);                                                                  // ino:286
        double t6 = now();
        t_op[0] += t1 - t0; t_op[1] += t2 - t1; t_op[2] += t3 - t2; t_op[3] += t4 - t3;
        t_op[4] += t5 - t4; t_op[5] += t6 - t5;
    }
}

// the same steps through fs_dist_*: `world` ranks of one process on device 0
int run_decomposed(void *gl, State &s, int dim_x, int dim_y, int steps, int iters, int n_drags, int world)
{
    auto ctx_create = (decltype(&fs_ctx_create))must_sym(gl, "fs_ctx_create");
    auto ctx_destroy = (decltype(&fs_ctx_destroy))must_sym(gl, "fs_ctx_destroy");
    auto set_opt = (decltype(&fs_ctx_set_option))must_sym(gl, "fs_ctx_set_option");
    auto get_opt = (decltype(&fs_ctx_get_option))must_sym(gl, "fs_ctx_get_option");
    auto d_create = (decltype(&fs_dist_create))must_sym(gl, "fs_dist_create");
    auto d_destroy = (decltype(&fs_dist_destroy))must_sym(gl, "fs_dist_destroy");
    auto d_window = (decltype(&fs_dist_window))must_sym(gl, "fs_dist_window");
    auto d_connect = (decltype(&fs_dist_connect_local))must_sym(gl, "fs_dist_connect_local");
    auto d_upload = (decltype(&fs_dist_upload))must_sym(gl, "fs_dist_upload");
    auto d_download = (decltype(&fs_dist_download))must_sym(gl, "fs_dist_download");
    auto d_step = (decltype(&fs_dist_step))must_sym(gl, "fs_dist_step");
    auto d_check = (decltype(&fs_dist_check))must_sym(gl, "fs_dist_check");
    G.err = (decltype(G.err))must_sym(gl, "fs_error_string");
    std::vector<fs_ctx *> ctx(world);
    std::vector<fs_dist *> rank(world);
    for (int r = 0; r < world; r++) {
        gcheck(ctx_create(&ctx[r], 0, FS_STREAM_NEW), "fs_ctx_create");
        int sms = 0;
        gcheck(get_opt(ctx[r], "num_sms", &sms), "num_sms");
        gcheck(set_opt(ctx[r], "sor_grid_limit", sms / world > 0 ? sms / world : 1), "sor_grid_limit");   // all ranks stay resident
        fs_dist_config cfg = {};
        cfg.gdim_x = dim_x; cfg.gdim_y = dim_y; cfg.world = world; cfg.rank = r;
        cfg.ghost = 64; cfg.advect_halo = 24; cfg.iters = iters;
        cfg.dt = 1 / 30.0f; cfg.dx = 1.0f; cfg.omega = 1.96f;
        gcheck(d_create(&rank[r], &cfg, ctx[r]), "fs_dist_create");
    }
    for (int r = 0; r < world; r++) gcheck(d_connect(rank[r], rank.data()), "fs_dist_connect_local");
    std::vector<fs_tile> win(world);
    for (int r = 0; r < world; r++) {
        gcheck(d_window(rank[r], &win[r]), "fs_dist_window");
        const fs_tile &w = win[r];
        std::vector<float> v((size_t)2 * w.nx * w.ny);
        std::vector<uint32_t> c((size_t)3 * w.nx * w.ny);
        for (int y = 0; y < w.ny; y++) {
            std::memcpy(&v[(size_t)2 * y * w.nx], &s.v[2 * ((size_t)(w.oy + y) * dim_x + w.ox)], sizeof(float) * 2 * w.nx);
            std::memcpy(&c[(size_t)3 * y * w.nx], &s.c[3 * ((size_t)(w.oy + y) * dim_x + w.ox)], sizeof(uint32_t) * 3 * w.nx);
        }
        gcheck(d_upload(rank[r], (const fs_vec2f *)v.data(), (const fs_rgb_uq32 *)c.data()), "fs_dist_upload");
    }
    for (int st = 0; st < steps; st++) {
        std::vector<Drag> dr = step_drags(st, n_drags, dim_x, dim_y);
        for (int r = 0; r < world; r++)      // asynchronous: every rank's step is only enqueued here
            gcheck(d_step(rank[r], (const fs_drag *)dr.data(), n_drags), "fs_dist_step");
    }
    for (int r = 0; r < world; r++) {
        gcheck(d_check(rank[r]), "fs_dist_check");
        const fs_tile &w = win[r];
        const int rw = w.x1 - w.x0, rh = w.y1 - w.y0;
        std::vector<float> v((size_t)2 * rw * rh), p((size_t)rw * rh), d((size_t)rw * rh);
        std::vector<uint32_t> c((size_t)3 * rw * rh);
        gcheck(d_download(rank[r], (fs_vec2f *)v.data(), (fs_rgb_uq32 *)c.data(), p.data(), d.data()), "fs_dist_download");
        for (int y = 0; y < rh; y++) {
            const size_t g = (size_t)(w.oy + w.y0 + y) * dim_x + (w.ox + w.x0);
            std::memcpy(&s.v[2 * g], &v[(size_t)2 * y * rw], sizeof(float) * 2 * rw);
            std::memcpy(&s.c[3 * g], &c[(size_t)3 * y * rw], sizeof(uint32_t) * 3 * rw);
            std::memcpy(&s.p[g], &p[(size_t)y * rw], sizeof(float) * rw);
            std::memcpy(&s.d[g], &d[(size_t)y * rw], sizeof(float) * rw);
        }
    }
    for (int r = 0; r < world; r++) {
        d_destroy(rank[r]);
        ctx_destroy(ctx[r]);
    }
    return 0;
}

}  // namespace

int main(int argc, char **argv)
{
    // --decomposed N keeps N ranks x 3 streams of mutually waiting kernels on ONE device: each needs a hardware work
    // queue of its own (include/fluid_b200.h, fs_dist_create); must be set before the CUDA library initialises
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    std::string gpu_lib, cpu_lib, prefix = "ref_", gpu_ops = "all";
    int dim_x = 61, dim_y = 81, steps = 20, iters = 10, n_drags = 4, decomposed = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--gpu-lib") gpu_lib = next();
        else if (a == "--cpu-lib") cpu_lib = next();
        else if (a == "--cpu-prefix") prefix = next();
        else if (a == "--gpu-ops") gpu_ops = next();
        else if (a == "--dim-x") dim_x = std::atoi(next());
        else if (a == "--dim-y") dim_y = std::atoi(next());
        else if (a == "--steps") steps = std::atoi(next());
        else if (a == "--iters") iters = std::atoi(next());
        else if (a == "--drags") n_drags = std::atoi(next());
        else if (a == "--decomposed") decomposed = std::atoi(next());
        else {
            std::printf("usage: %s --cpu-lib LIB [--cpu-prefix ref_|oracle_] [--gpu-lib LIB] [--gpu-ops list|all|none]\n"
                        "          [--dim-x N --dim-y N --steps N --iters N --drags N] [--decomposed WORLD]\n", argv[0]);
            return a == "--help" ? 0 : 2;
        }
    }
    if (cpu_lib.empty()) { std::fprintf(stderr, "--cpu-lib is required\n"); return 2; }
    void *cl = dlopen(cpu_lib.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!cl) { std::fprintf(stderr, "dlopen %s: %s\n", cpu_lib.c_str(), dlerror()); return 2; }
    Ops cpu;
    cpu.advect_v = (decltype(cpu.advect_v))must_sym(cl, prefix + "advect_vec2f");
    cpu.advect_c = (decltype(cpu.advect_c))must_sym(cl, prefix + "advect_rgb_uq32");
    cpu.divergence = (decltype(cpu.divergence))must_sym(cl, prefix + "calculate_divergence");
    cpu.gradient = (decltype(cpu.gradient))must_sym(cl, prefix + "subtract_gradient");
    cpu.poisson = (decltype(cpu.poisson))must_sym(cl, prefix + "poisson_solve");
    cpu.drags = (decltype(cpu.drags))must_sym(cl, prefix + "apply_drags");

    Ops mixed = cpu;
    bool any_gpu = !gpu_lib.empty() && gpu_ops != "none";
    void *gl = nullptr;
    if (any_gpu || (decomposed > 0 && !gpu_lib.empty())) {
        gl = dlopen(gpu_lib.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!gl) { std::fprintf(stderr, "dlopen %s: %s\n", gpu_lib.c_str(), dlerror()); return 2; }
    }
    if (any_gpu) {
        auto ctx_create = (decltype(&fs_ctx_create))must_sym(gl, "fs_ctx_create");
        G.err = (decltype(G.err))must_sym(gl, "fs_error_string");
        G.advect_vec2f = (decltype(G.advect_vec2f))must_sym(gl, "fsh_advect_vec2f");
        G.advect_rgb = (decltype(G.advect_rgb))must_sym(gl, "fsh_advect_rgb_uq32");
        G.divergence = (decltype(G.divergence))must_sym(gl, "fsh_calculate_divergence");
        G.gradient = (decltype(G.gradient))must_sym(gl, "fsh_subtract_gradient");
        G.poisson = (decltype(G.poisson))must_sym(gl, "fsh_poisson_solve");
        gcheck(ctx_create(&G.ctx, 0, nullptr), "fs_ctx_create");
        auto has = [&](const char *name) { return gpu_ops == "all" || ("," + gpu_ops + ",").find(std::string(",") + name + ",") != std::string::npos; };
        if (has("advect_v")) mixed.advect_v = g_advect_v;
        if (has("advect_c")) mixed.advect_c = g_advect_c;
        if (has("divergence")) mixed.divergence = g_div;
        if (has("gradient")) mixed.gradient = g_grad;
        if (has("poisson")) mixed.poisson = g_pois;
        // the drag overwrite is a handful of host stores on host-resident state: it stays on the CPU side
    }

    const size_t n = (size_t)dim_x * dim_y;
    State a, b;
    a.v.resize(2 * n); a.c.resize(3 * n); a.p.assign(n, 0.f); a.d.assign(n, 0.f);
    for (size_t k = 0; k < 2 * n; k++) a.v[k] = ((float)(splitmix(0xF1D0 + k) >> 40) / 16777216.0f * 2 - 1) * 90.0f;
    for (size_t k = 0; k < 3 * n; k++) a.c[k] = (uint32_t)(splitmix(0xD1E + k) >> 33) * 2u;
    b = a;
    State dstate = a;

    double t_cpu[6] = {0}, t_mix[6] = {0};
    run(cpu, a, dim_x, dim_y, steps, iters, n_drags, t_cpu);
    run(mixed, b, dim_x, dim_y, steps, iters, n_drags, t_mix);
    const char *names[6] = {"advect_v", "drags", "divergence", "poisson", "gradient", "advect_c"};
    std::printf("grid %dx%d, %d steps, K=%d, gpu-ops=%s\n", dim_x, dim_y, steps, iters, any_gpu ? gpu_ops.c_str() : "none");
    for (int k = 0; k < 6; k++)
        std::printf("  %-10s cpu %9.3f ms/step   selected %9.3f ms/step\n", names[k], 1e3 * t_cpu[k] / steps, 1e3 * t_mix[k] / steps);
    uint64_t ha[4] = {fnv(a.v.data(), 8 * n), fnv(a.c.data(), 12 * n), fnv(a.p.data(), 4 * n), fnv(a.d.data(), 4 * n)};
    uint64_t hb[4] = {fnv(b.v.data(), 8 * n), fnv(b.c.data(), 12 * n), fnv(b.p.data(), 4 * n), fnv(b.d.data(), 4 * n)};
    std::printf("  fnv1a64 v/c/p/d  cpu      %016llx %016llx %016llx %016llx\n", (unsigned long long)ha[0], (unsigned long long)ha[1], (unsigned long long)ha[2], (unsigned long long)ha[3]);
    std::printf("  fnv1a64 v/c/p/d  selected %016llx %016llx %016llx %016llx\n", (unsigned long long)hb[0], (unsigned long long)hb[1], (unsigned long long)hb[2], (unsigned long long)hb[3]);
    bool same = !std::memcmp(ha, hb, sizeof(ha));
    if (decomposed > 0 && gl) {
        run_decomposed(gl, dstate, dim_x, dim_y, steps, iters, n_drags, decomposed);
        uint64_t hd[4] = {fnv(dstate.v.data(), 8 * n), fnv(dstate.c.data(), 12 * n), fnv(dstate.p.data(), 4 * n), fnv(dstate.d.data(), 4 * n)};
        std::printf("  fnv1a64 v/c/p/d  fs_dist  %016llx %016llx %016llx %016llx  (%d ranks)\n", (unsigned long long)hd[0],
                    (unsigned long long)hd[1], (unsigned long long)hd[2], (unsigned long long)hd[3], decomposed);
        same = same && !std::memcmp(ha, hd, sizeof(ha));
    }
    std::printf("%s\n", same ? "MATCH: bit-identical" : "MISMATCH");
    return same ? 0 : 1;
}
