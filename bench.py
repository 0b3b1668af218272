#!/usr/bin/env python
"""bench.py — the reference's headline metric on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): Mcell-steps/s of the full stable-fluids step (advect v +
drags + divergence + 50 red-black SOR iterations + gradient-subtract + advect
dye) on a 4096x4096 grid; one "step" = one loop() body (ino:249-289) over the
whole grid.  N>1 (torchrun, one rank per GPU): the grid is block-decomposed with
4096x4096 nodes per GPU (weak scaling) and halos exchanged over NCCL.

Prints ONE JSON line (rank 0).  `value` is device-resident throughput; `e2e` is
the same step through the host-pointer drop-in fsh_step() with pinned host
buffers (H2D + D2H inside the timed region); `roofline` is the SOR solve against
the measured HBM copy peak; `cpu_baseline` is the reference's own C++ timed on
this box's host.
"""
from __future__ import annotations

import argparse
import json
import os

# fs_dist keeps three streams per rank busy with kernels that wait for other ranks' flags (the SOR passes and the two
# side-stream halo exchanges): each needs a hardware work queue of its own, or a waiting kernel can sit in front of the
# one it waits for.  The default is 8 queues shared with torch's and NCCL's streams; must be set before CUDA starts.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mcell-steps/s (advect+project, 50 SOR iters) at 4096^2"   # N>1: 4096^2 nodes PER GPU (config.grid)
UNIT = "Mcell-steps/s"
GRID = 4096          # per-GPU tile edge
ITERS = 50
N_DRAGS = 16
SOR_BYTES_PER_NODE_ITER = 12.0     # read p + read d + write p per full iteration
STEP_BYTES_PER_NODE = 80.0 + 12.0 * ITERS


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C++ (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_impl():
    import oracle
    oracle.build()
    if oracle.have_ref():
        r = oracle.Ref()
        if r.saturates():
            return r, "reference"
    return oracle.Oracle(), "port"


def cpu_step_rate(dim_x, dim_y, steps, warmup):
    """Mcell-steps/s of the single-threaded reference on a dim_x x dim_y grid."""
    from esp32_fluid_simulation_b200 import synth
    impl, kind = cpu_impl()
    v = synth.velocity(GRID, GRID, window=(0, 0, dim_x, dim_y))
    c = synth.dye(GRID, GRID, window=(0, 0, dim_x, dim_y))
    times = []
    for s in range(warmup + steps):
        dr = synth.drags(dim_x, dim_y, s, n=N_DRAGS)
        t0 = time.perf_counter()
        impl.step(v, c, dr, synth.DT, synth.DX, ITERS, synth.OMEGA)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return dim_x * dim_y * len(times) / total / 1e6, total / len(times), kind


def pin_to_one_core():
    """The reference is single-threaded; keep it on one core so the scheduler does not migrate it."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        os.sched_setaffinity(0, {cores[len(cores) // 2]})
        return cores[len(cores) // 2]
    except Exception:  # noqa: BLE001
        return None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the SAME config as the B200 arm: every step is one full loop() body on the whole 4096^2 grid
    # (~4 s of CPU each; the driver's --steps 20 --warmup 3 is ~90 s).  Bounded: at most 25 timed steps.
    dim_x, dim_y = GRID, GRID
    core = pin_to_one_core()
    steps, warmup = max(1, min(args.steps, 25)), max(0, min(args.warmup, 3))
    rate, sec, kind = cpu_step_rate(dim_x, dim_y, steps, warmup)
    sample = (f"{steps} timed + {warmup} warm-up full loop() bodies on the whole {dim_x}x{dim_y} grid, "
              f"K={ITERS}, {N_DRAGS} drags; single thread pinned to core {core} (the reference is single-threaded)")
    workload = f"single {GRID}x{GRID} grid, {ITERS} SOR iterations, velocity + dye advection"
    same = True
    if args.gpus > 1:
        # the B200 arm's N>1 workload is one grid of 4096^2 nodes PER GPU; the single-threaded reference is timed
        # on one rank's share of it per step (the whole grid would take N x 4 s per step); the metric is a rate
        from esp32_fluid_simulation_b200.dist import process_grid
        px, py = process_grid(args.gpus)
        workload = (f"{GRID * px}x{GRID * py} grid ({GRID}x{GRID} nodes per GPU on the B200 arm), {ITERS} SOR iterations, "
                    "velocity + dye advection")
        sample += f"; bounded sample: one {GRID}x{GRID} rectangle (one rank's share of the {GRID * px}x{GRID * py} grid) per step"
        same = False
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+uq32", "data": "synthetic",
        "config": {"workload": workload,
                   "grid": [dim_x, dim_y], "sor_iters": ITERS, "drags_per_step": N_DRAGS, "same_config": same,
                   "sample": sample},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

def start_watchdog(seconds: float):
    """Multi-rank runs only: if a rank dies, its neighbours spin in halo_exchange_kernel waiting for its
    flags.  Exit hard (which tears the CUDA context down and kills the kernel) instead of hanging."""
    def fire():
        print(f"[bench] watchdog: no result after {seconds:.0f} s, aborting", file=sys.stderr, flush=True)
        os._exit(3)
    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


def verify_decomposed(fb, torch, dist, local_rank, world, rank, gx, gy, iters, ghost, halo, n_drags) -> dict:
    """UNTIMED parity leg of bench.py's N>1 arm (runs before the timed region):
      oracle_small      a 1024x768, K=20, 3-step decomposed run (same code path, same process grid),
                        gathered on rank 0 and bit-compared with the CPU checker (the reference's own
                        compiled code when oracle/_ref travelled, else the pinned port);
      one_gpu_equals_n  one step of THE BENCH GRID: every rank also steps the whole global grid on
                        its own GPU through fs_step and compares its rectangle bit for bit (the
                        single-GPU path is oracle-checked at 4096^2 by tests/test_gpu_parity.py)."""
    from esp32_fluid_simulation_b200 import synth
    from esp32_fluid_simulation_b200.dist import NativeDist, _device_view
    dev = torch.device("cuda", local_rank)
    out = {}

    def connect(sim):
        handles = [None] * world
        dist.all_gather_object(handles, sim.ipc_handle())
        sim.connect(handles)

    # (a) small run against the CPU checker
    sx, sy, sk, steps = 1024, 768, 20, 3
    ctx = fb.Context(local_rank, torch.cuda.current_stream(dev))
    sim = NativeDist(ctx, sx, sy, world, rank, sk, ghost=32, advect_halo=12, dt=synth.DT, dx=synth.DX, omega=synth.OMEGA)
    connect(sim)
    w = sim.window
    sim.upload(synth.velocity(sx, sy, vmax=150.0, window=(w.ox, w.oy, w.nx, w.ny)),
               synth.dye(sx, sy, window=(w.ox, w.oy, w.nx, w.ny)))
    drs = [synth.drags(sx, sy, s, n=8, vmax=300.0) for s in range(steps)]
    for s in range(steps):
        sim.step(drs[s])
    sim.check()
    got = sim.download("vcp")
    parts = [None] * world
    dist.all_gather_object(parts, (w.ox + w.x0, w.oy + w.y0, got["v"], got["c"], got["p"]))
    ok = 1
    if rank == 0:
        import oracle as _oracle            # test infrastructure: the checker, never the thing measured
        _oracle.build()
        chk = _oracle.Checker()
        ov, oc = synth.velocity(sx, sy, vmax=150.0), synth.dye(sx, sy)
        for s in range(steps):
            ov, oc, op, _ = chk.step(ov, oc, drs[s], synth.DT, synth.DX, sk, synth.OMEGA, want_fields=True)
        for x0, y0, pv, pc, pp in parts:
            h, wd = pv.shape[:2]
            same = (np.array_equal(pv.view(np.uint32), ov[y0:y0 + h, x0:x0 + wd].view(np.uint32)) and
                    np.array_equal(pc, oc[y0:y0 + h, x0:x0 + wd]) and
                    np.array_equal(pp.view(np.uint32), op[y0:y0 + h, x0:x0 + wd].view(np.uint32)))
            ok &= int(same)
        out["checker"] = chk.kind
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.broadcast(flag, 0)
    out["oracle_small"] = bool(flag.item())
    sim.close()

    # (b) the bench grid itself: decomposed step == single-GPU step of the whole grid
    g = torch.Generator(device=dev).manual_seed(4321)          # same seed, same generator: same field everywhere
    v = (torch.rand(gy, gx, 2, device=dev, generator=g) - 0.5) * 120.0
    c = torch.randint(0, 2 ** 31 - 1, (gy, gx, 3), device=dev, dtype=torch.int32, generator=g)
    sim = NativeDist(ctx, gx, gy, world, rank, iters, ghost=ghost, advect_halo=halo, dt=synth.DT, dx=synth.DX,
                     omega=synth.OMEGA)
    connect(sim)
    w = sim.window
    sim.upload(v[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx].contiguous(), c[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx].contiguous())
    dr = synth.drags(gx, gy, 0, n=n_drags)
    sim.step(dr)
    sim.check()
    c2 = torch.empty_like(c)
    ctx.step_pingpong(v, c, c2, dr, gx, gy, synth.DT, synth.DX, iters, synth.OMEGA)
    ctx.synchronize()
    pv, pc, _, _ = sim.device_fields()
    dv = torch.as_tensor(_device_view(pv, (w.ny, w.nx, 2), "<f4"), device=dev)
    dc = torch.as_tensor(_device_view(pc, (w.ny, w.nx, 3), "<i4"), device=dev)
    ys, xs = slice(w.oy + w.y0, w.oy + w.y1), slice(w.ox + w.x0, w.ox + w.x1)
    same = (torch.equal(dv[w.y0:w.y1, w.x0:w.x1].view(torch.int32), v[ys, xs].view(torch.int32)) and
            torch.equal(dc[w.y0:w.y1, w.x0:w.x1], c2[ys, xs]))
    flag = torch.tensor([int(same)], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["one_gpu_equals_n"] = bool(flag.item())
    out["one_gpu_equals_n_grid"] = [gx, gy]
    sim.close()
    del v, c, c2
    torch.cuda.empty_cache()
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        from esp32_fluid_simulation_b200 import dist as fdist
        args.clock_sampler = ClockSampler
        args.verify_fn = verify_decomposed
        watchdog = start_watchdog(float(os.environ.get("FS_BENCH_WATCHDOG_S", "900")))
        result = fdist.bench_decomposed(args, GRID, args.iters, N_DRAGS)
    else:
        result = bench_single(args, fb, synth, torch)
    if rank == 0:
        print(json.dumps(result), flush=True)
    if world > 1:
        dist.barrier()
        watchdog.cancel()
        dist.destroy_process_group()


def bench_single(args, fb, synth, torch):
    n, ny = GRID, GRID
    if getattr(args, "grid", ""):        # denominators for the decomposed configs: one rank's rectangle on one GPU
        n, ny = (int(t) for t in args.grid.lower().split("x"))
    nodes = n * ny
    stream = torch.cuda.Stream()
    ctx = fb.Context(0, stream)
    v0, c0 = synth.velocity(n, ny), synth.dye(n, ny)
    drags = [synth.drags(n, ny, s, n=N_DRAGS) for s in range(args.warmup + args.steps)]
    with torch.cuda.stream(stream):
        dv = torch.from_numpy(v0).cuda()
        dc = torch.from_numpy(c0.view(np.int32)).cuda()
    stream.synchronize()

    with torch.cuda.stream(stream):
        dc2 = torch.empty_like(dc)
    dyes = [dc, dc2]

    def step(s):
        # loop() swaps the dye pointers (ino:286): c_in -> c_out, then the roles alternate
        ctx.step_pingpong(dv, dyes[s & 1], dyes[(s & 1) ^ 1], drags[s], n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)

    for s in range(args.warmup):
        step(s)
    stream.synchronize()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ctx.launch_count
    with ClockSampler(0) as clk:
        time.sleep(0.15)
        ev0.record(stream)
        for s in range(args.warmup, args.warmup + args.steps):
            step(s)
        ev1.record(stream)
        launches = ctx.launch_count - launches0
        stream.synchronize()
        torch.cuda.synchronize()
        # keep the device busy a little longer so the sampler sees clocks under load
        t_end = time.time() + 0.3
        while time.time() < t_end:
            step(args.warmup)
        stream.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    value = nodes / (ms * 1e-3) / 1e6

    # --- dominant kernel: the SOR solve, timed alone on the same stream ---
    with torch.cuda.stream(stream):
        dd = torch.randn(ny, n, device="cuda") * 10
        dp = torch.empty_like(dd)
    for _ in range(2):
        ctx.poisson_solve(dp, dd, n, ny, synth.DX, ITERS, synth.OMEGA)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    l0 = ctx.launch_count
    s0.record(stream)
    for _ in range(reps):
        ctx.poisson_solve(dp, dd, n, ny, synth.DX, ITERS, synth.OMEGA)
    s1.record(stream)
    stream.synchronize()
    sor_launches = (ctx.launch_count - l0) // reps
    sor_ms = s0.elapsed_time(s1) / reps
    peak, peak_src = measured_peaks()
    achieved = SOR_BYTES_PER_NODE_ITER * nodes * ITERS / (sor_ms * 1e-3) / 1e9
    # `traffic`: ncu dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per LAUNCH, from the
    # ncu capture recorded in profiles/sor_traffic.json — used only if that capture was taken on the kernel
    # configuration this run uses (strip shape + iterations per pass); otherwise null, never a stale constant
    sor_t, sor_shape = ctx.get_option("sor_t"), ctx.get_option("sor_shape")
    traffic = traffic_solve = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "sor_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("sor_t") == sor_t and tj.get("sor_shape") == sor_shape and tj.get("grid") == [n, ny]:
                traffic, traffic_solve = tj.get("dram_bytes_per_pass"), tj.get("dram_bytes_per_solve")
                traffic_src = tj.get("source")
        except Exception:
            traffic = traffic_solve = None
    alg_launch = SOR_BYTES_PER_NODE_ITER * nodes * ITERS / max(sor_launches, 1)
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel": f"sor_blocked_tma_kernel (strip shape {sor_shape}): one launch = one pass of the SOR solve = up to "
                  f"{sor_t} fused red-black iterations over the whole grid (fs_poisson_solve, K={ITERS} => "
                  f"{sor_launches} launches)",
        "algorithmic_bytes_per_launch": alg_launch, "avg_launch_ms": sor_ms / max(sor_launches, 1),
        "algorithmic_bytes_per_solve": SOR_BYTES_PER_NODE_ITER * nodes * ITERS, "traffic_per_solve": traffic_solve,
        "note": "achieved = 12 B/node-iteration (SURVEY 8d) x nodes x iterations per launch / average launch time; "
                f"frac > 1 because temporal blocking moves ~1/{sor_t} of those bytes (see traffic); the kernel is "
                "instruction-issue bound, not HBM bound (profiles/)",
        "ms": sor_ms, "launches_per_solve": sor_launches,
        "gnode_iters_per_s": nodes * ITERS / (sor_ms * 1e-3) / 1e9,
        "step_frac": (STEP_BYTES_PER_NODE * nodes / (ms * 1e-3) / 1e9) / peak,
        "sor_share_of_step": sor_ms / ms,
    }

    # --- the advection kernels (north_star names them too): CUDA-event time of each, alone, on the same stream ---
    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            fn()
        a1.record(stream)
        stream.synchronize()
        return a0.elapsed_time(a1) / reps

    with torch.cuda.stream(stream):
        dv2 = torch.empty_like(dv)
    adv = {}
    for name, bpn, fn in (
            ("advect velocity + drags + divergence (advect_div_tma_kernel, fused in fs_step)", 28.0,
             lambda: ctx.advect_drags_divergence(dv2, dd, dv, drags[0], n, ny, synth.DT, synth.DX)),
            ("advect velocity (advect_tma_kernel<Vec2Payload>)", 16.0,
             lambda: ctx.advect(dv2, dv, dv, n, ny, synth.DT, True)),
            ("advect dye (advect_tma_kernel<RgbPayload>)", 32.0,
             lambda: ctx.advect(dc2, dc, dv, n, ny, synth.DT, False))):
        t_ms = timed(fn)
        gbs = bpn * nodes / (t_ms * 1e-3) / 1e9
        adv[name] = {"ms": t_ms, "algorithmic_bytes_per_node": bpn, "achieved": gbs, "unit": "GB/s", "frac": gbs / peak}
    roofline_advect = {"bound": "hbm", "peak": peak, "peak_source": peak_src, "kernels": adv,
                       "traffic": "profiles/ (ncu dram__bytes per launch of each kernel)",
                       "note": "algorithmic bytes per node (SURVEY 8d): velocity 8 read + 8 written (+ 4 divergence "
                               "written + the divergence kernel's 8 read saved = 28 for the fused kernel), dye 8 + 12 "
                               "read + 12 written"}

    # --- BASELINE.json configs[1]: 65,536 independent 80x60 grids, one CTA per grid (fs_ensemble_step) ---
    extra = {}
    try:
        eb, ex, ey, ek = 65536, 80, 60, 10
        with torch.cuda.stream(stream):
            g = torch.Generator(device="cuda").manual_seed(7)
            ev = (torch.rand(eb, ey, ex, 2, device="cuda", generator=g) - 0.5) * 120.0
            ec = torch.randint(0, 2 ** 31 - 1, (eb, ey, ex, 3), device="cuda", dtype=torch.int32, generator=g)
        ens = {}
        for n_steps in (1, 16):
            t_ms = timed(lambda: ctx.ensemble_step(ev, ec, eb, ex, ey, synth.DT, synth.DX, ek, synth.OMEGA, n_steps), reps=3)
            cells = eb * ex * ey * n_steps
            ens[f"n_steps={n_steps}"] = {"ms": t_ms, "gcell_steps_per_s": cells / (t_ms * 1e-3) / 1e9,
                                         "grid_steps_per_s": eb * n_steps / (t_ms * 1e-3),
                                         "state_io_GBps": eb * ex * ey * 40 / (t_ms * 1e-3) / 1e9}
        # the reference's own node grid (61 x 81 nodes = 60 x 80 cells, ino:37-38): odd N, every copy congruent-shifted
        ens_ref = {}
        rb, rx, ry = 16384, 61, 81
        with torch.cuda.stream(stream):
            rv = (torch.rand(rb, ry, rx, 2, device="cuda", generator=g) - 0.5) * 120.0
            rc = torch.randint(0, 2 ** 31 - 1, (rb, ry, rx, 3), device="cuda", dtype=torch.int32, generator=g)
        for n_steps in (1, 16):
            t_ms = timed(lambda: ctx.ensemble_step(rv, rc, rb, rx, ry, synth.DT, synth.DX, ek, synth.OMEGA, n_steps), reps=3)
            ens_ref[f"n_steps={n_steps}"] = {"ms": t_ms, "gcell_steps_per_s": rb * rx * ry * n_steps / (t_ms * 1e-3) / 1e9,
                                             "grid_steps_per_s": rb * n_steps / (t_ms * 1e-3)}
        del rv, rc
        extra["ensemble"] = {"workload": f"{eb} independent {ex}x{ey} grids, K={ek}, one CTA per grid, state resident in "
                                         "shared memory (BASELINE.json configs[1])", "results": ens,
                             "reference_node_grid_61x81_x16384": ens_ref,
                             "kernel": "ensemble_reg_kernel<2,640,1> (csrc/ensemble_reg.cuh): projection in registers, "
                                       "4x2-node block per thread, rim values through per-thread mailboxes; calls of "
                                       "<= 3 steps move the state with cp.async.bulk copies under the neighbouring "
                                       "grids' compute",
                             "bound": "instruction issue 62 % / shared-memory wavefronts 65 % (profiles/"
                                      "r02_ncu_ensemble_reg_r2*.json), not HBM: state I/O is 40 B/node per CALL"}
        del ev, ec
    except Exception as e:  # noqa: BLE001 — never let an extra take the headline down
        extra["ensemble"] = {"error": f"{type(e).__name__}: {e}"}

    # --- fs_sim: the same step as ONE CUDA-graph launch, and the frame-only end-to-end path ---
    try:
        sim = fb.Sim(ctx, n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)
        sim.upload(dv, dc)
        for s in range(4):
            sim.step(drags[s % len(drags)])
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_steps = max(args.steps, 10)
        l0 = ctx.launch_count
        g0.record(stream)
        for s in range(k_steps):
            sim.step(drags[s % len(drags)])
        g1.record(stream)
        stream.synchronize()
        st = sim.stats
        extra["graph_step"] = {"ms_per_step": g0.elapsed_time(g1) / k_steps, "graph_launches_per_step": 1,
                               "kernels_per_graph": (ctx.launch_count - l0) // k_steps, "stats": st,
                               "api": "fs_sim_step: one cudaGraphLaunch per loop() body, drag records re-armed with "
                                      "cudaGraphExecKernelNodeSetParams"}
        sim.close()
        # frames only leave the device: state resident, every step's 4x RGB565 frame streamed to pinned host memory
        simf = fb.Sim(ctx, n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA, frame=True)
        simf.upload(dv, dc)
        f_steps, got = 6, 0
        for s in range(2):                       # warm-up: two eager steps, then the graphs exist
            simf.step(drags[s])
        while simf.acquire_frame() is not None:
            simf.release_frame()
        t0 = time.perf_counter()
        for s in range(f_steps):
            while not simf.step(drags[s % len(drags)]):
                simf.acquire_frame()
                simf.release_frame()
                got += 1
        while got < f_steps:
            simf.acquire_frame()
            simf.release_frame()
            got += 1
        f_s = (time.perf_counter() - t0) / f_steps
        frame_bytes = 16 * (n - 1) * (ny - 1) * 2
        extra["e2e_frames_only"] = {"value": nodes / f_s / 1e6, "unit": UNIT, "ms_per_step": f_s * 1e3, "steps": f_steps,
                                    "h2d_bytes_per_step": 12 * N_DRAGS, "d2h_bytes_per_step": frame_bytes,
                                    "api": "fs_sim_step + fs_sim_acquire_frame/release_frame: state stays on the device, "
                                           "the step's RGB565 frame (rendered inside the dye advect) is copied to pinned "
                                           "host memory on a side stream under the next step (PCIe-bound: 32 B/node)"}
        simf.close()
    except Exception as e:  # noqa: BLE001
        extra["graph_step"] = {"error": f"{type(e).__name__}: {e}"}

    # --- e2e: the host-pointer drop-in fsh_step with pinned host buffers ---
    hv = torch.from_numpy(v0.copy()).pin_memory()
    hc = torch.from_numpy(c0.view(np.int32).copy()).pin_memory()
    hv_np, hc_np = hv.numpy(), hc.numpy().view(np.uint32)
    e2e_steps = max(3, min(args.steps, 10))
    for s in range(2):
        ctx.step(hv_np, hc_np, drags[s % len(drags)], n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ctx.step(hv_np, hc_np, drags[s % len(drags)], n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    state_bytes = nodes * (8 + 12)
    e2e = {"value": nodes / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": state_bytes + 12 * N_DRAGS,
           "d2h_bytes_per_step": state_bytes, "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "api": "fsh_step (host pointers in, host pointers out; pinned buffers)",
           "dye_bands": ctx.get_option("e2e_bands"), "dye_redos": ctx.get_option("e2e_redos")}
    # the same call with the dye travelling in fewer / more row bands (PCIe-bound either way: 20 B/node each direction)
    band_ms = {}
    for bands in (1, 4, 16):
        ctx.set_option("e2e_bands", bands)
        ctx.step(hv_np, hc_np, drags[0], n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)
        t0 = time.perf_counter()
        for s in range(3):
            ctx.step(hv_np, hc_np, drags[s % len(drags)], n, ny, synth.DT, synth.DX, ITERS, synth.OMEGA)
        band_ms[str(bands)] = round((time.perf_counter() - t0) / 3 * 1e3, 3)
    ctx.set_option("e2e_bands", e2e["dye_bands"])
    extra["e2e_ms_by_dye_bands"] = band_ms

    # --- cpu baseline beside it (bounded: 2 full-size steps, ~12 s) ---
    cpu = None
    if not args.no_cpu_baseline and n == ny == GRID:
        rate, sec, kind = cpu_step_rate(n, n, 2, 0)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": f"2 full steps of the same {n}x{n} K={ITERS} workload, single thread "
                         f"({sec:.2f} s/step)"}

    return {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+uq32", "data": "synthetic",
        "config": {"workload": f"single {n}x{ny} grid, {ITERS} SOR iterations, velocity + dye advection",
                   "grid": [n, ny], "sor_iters": ITERS, "drags_per_step": N_DRAGS,
                   "l2": "state (v 134 MB + dye 201 MB + p/div 134 MB) exceeds the 126 MB L2; no flush needed",
                   "options": {k: ctx.get_option(k) for k in ("sor", "sor_t", "sor_shape", "advect", "fuse")}},
        "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_advect": roofline_advect, "cpu_baseline": cpu, "extra": extra,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--grid", default="", help="N=1 only: WxH grid instead of 4096x4096 (the single-GPU denominator of a "
                    "decomposed config, e.g. 8192x4096 = one rank's rectangle of 16384^2 on 8 GPUs)")
    ap.add_argument("--global-grid", default="", help="N>1 only: GXxGY global grid instead of 4096^2 per GPU "
                    "(BASELINE.json configs[3] = 16384x16384, configs[4] = 24576x32768)")
    ap.add_argument("--iters", type=int, default=ITERS, help="SOR iterations per step (N>1 extra configs)")
    ap.add_argument("--upscale", action="store_true", help="N>1: also produce the 4x RGB565 frame every step")
    ap.add_argument("--no-extra", action="store_true", help="N>1: skip the extra BASELINE configs measured beside the headline")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the untimed parity leg (oracle_small, "
                    "one_gpu_equals_n) that runs before the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
