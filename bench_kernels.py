#!/usr/bin/env python
"""Per-kernel timings of the hot path at one grid size, each against its
algorithmic-bytes roofline, plus variant sweeps.  Development
tool: `python bench_kernels.py [--n 4096] [--sweep]`; the headline bench is
bench.py."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import esp32_fluid_simulation_b200 as fb  # noqa: E402
from esp32_fluid_simulation_b200 import synth  # noqa: E402


def timeit(stream, fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--sor-sweep", default="", help="comma-separated SOR shapes: time poisson_solve for T = 4..8 each")
    ap.add_argument("--out", default="")
    ap.add_argument("--ensemble", type=int, default=0, help="batch size: bench the smem-resident ensemble kernel instead")
    ap.add_argument("--ens-shape", default="80x60")
    ap.add_argument("--ens-variant", default="0", help="comma-separated values of option 'ensemble' to time")
    ap.add_argument("--ens-steps", default="1,16", help="comma-separated steps per call")
    args = ap.parse_args()
    nx, ny = args.n, args.ny or args.n
    nodes = nx * ny
    peak = 6456.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    stream = torch.cuda.Stream()
    ctx = fb.Context(0, stream)
    if args.ensemble:
        return bench_ensemble(args, ctx, stream, peak)
    with torch.cuda.stream(stream):
        v = torch.from_numpy(synth.velocity(nx, ny)).cuda()
        c = torch.from_numpy(synth.dye(nx, ny).view(np.int32)).cuda()
        v2, c2 = torch.empty_like(v), torch.empty_like(c)
        d = torch.empty(ny, nx, device="cuda")
        p = torch.empty(ny, nx, device="cuda")
    stream.synchronize()
    rows = []

    def rec(name, ms, bytes_per_node, extra=None):
        gbs = bytes_per_node * nodes / (ms * 1e-3) / 1e9
        r = {"kernel": name, "ms": round(ms, 4), "alg_B_per_node": bytes_per_node,
             "alg_GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3)}
        if extra:
            r.update(extra)
        rows.append(r)
        print(json.dumps(r), flush=True)

    def run_ops(tag=""):
        rec("advect_vec2f" + tag, timeit(stream, lambda: ctx.advect(v2, v, v, nx, ny, synth.DT, True)), 16)
        rec("calculate_divergence" + tag, timeit(stream, lambda: ctx.calculate_divergence(d, v, nx, ny, 1.0)), 12)
        rec("subtract_gradient" + tag, timeit(stream, lambda: ctx.subtract_gradient(v2, p, nx, ny, 1.0)), 20)
        rec("advect_rgb_uq32" + tag, timeit(stream, lambda: ctx.advect(c2, c, v, nx, ny, synth.DT, False)), 32)
        rec("advect_drags_divergence" + tag, timeit(stream, lambda: ctx.advect_drags_divergence(v2, d, v, synth.drags(nx, ny, 0), nx, ny, synth.DT, 1.0)), 28)
        frame = torch.empty((nx - 1) * 4, (ny - 1) * 4, dtype=torch.int16, device="cuda")
        rec("upscale4_rgb565" + tag, timeit(stream, lambda: ctx.upscale4_rgb565(frame, c, nx, ny)), 44)
        rec("advect_rgb_frame (advect + frame in one kernel)" + tag,
            timeit(stream, lambda: ctx.advect_rgb_frame(c2, frame, c, v, nx, ny, synth.DT, False)), 32 + 32)

    def run_sor(tag=""):
        ms = timeit(stream, lambda: ctx.poisson_solve(p, d, nx, ny, 1.0, args.iters, 1.96), reps=5, warm=2)
        rec("poisson_solve" + tag, ms, 12 * args.iters,
            {"gnode_iters_per_s": round(nodes * args.iters / (ms * 1e-3) / 1e9, 1)})

    def run_step(tag=""):
        dr = synth.drags(nx, ny, 0)
        ms = timeit(stream, lambda: ctx.step(v, c, dr, nx, ny, synth.DT, 1.0, args.iters, 1.96), reps=5, warm=2)
        rec("step" + tag, ms, 80 + 12 * args.iters, {"mcell_steps_per_s": round(nodes / (ms * 1e-3) / 1e6, 1)})

    ctx.calculate_divergence(d, v, nx, ny, 1.0)
    if args.sor_sweep:
        for shape in (int(x) for x in args.sor_sweep.split(",")):
            ctx.set_option("sor_shape", shape)
            for t in (4, 5, 6, 7, 8):
                ctx.set_option("sor_t", t)
                run_sor(f"[blocked shape={shape} T={t}]")
    elif args.sweep:
        for adv in (0, 1):
            ctx.set_option("advect", adv)
            run_ops(f"[advect={adv}]")
        ctx.set_option("advect", 1)
        ctx.set_option("sor", 0)
        run_sor("[half-sweeps]")
        ctx.set_option("sor", 1)
        for shape in (2, 3, 5):
            for t in (4, 6, 8):
                ctx.set_option("sor_shape", shape)
                ctx.set_option("sor_t", t)
                ctx.set_option("sor_one_launch", 1)
                run_sor(f"[one-launch shape={shape} T={t}]")
        ctx.set_option("sor_one_launch", 0)
        for shape in (0, 2, 3, 4, 5, 6):
            ctx.set_option("sor_shape", shape)
            for t in (2, 4, 6, 8):
                ctx.set_option("sor_t", t)
                run_sor(f"[blocked shape={shape} T={t}]")
        ctx.set_option("sor_shape", 7)
        ctx.set_option("sor_t", 6)
        ctx.set_option("sor_one_launch", 0)
        for fuse in (0, 1, 2, 3):
            ctx.set_option("fuse", fuse)
            run_step(f"[fuse={fuse}]")
    else:
        run_ops()
        run_sor()
        run_step()
    if args.out:
        json.dump({"grid": [nx, ny], "iters": args.iters, "peak_gbs": peak, "rows": rows}, open(args.out, "w"), indent=1)


def bench_ensemble(args, ctx, stream, peak):
    """BASELINE.json configs[1]: `batch` independent CYD-sized grids, K=10, one CTA per grid."""
    dim_x, dim_y = (int(x) for x in args.ens_shape.split("x"))
    batch, iters = args.ensemble, 10
    n = dim_x * dim_y
    with torch.cuda.stream(stream):
        g = torch.Generator(device="cuda").manual_seed(7)
        v = (torch.rand(batch, dim_y, dim_x, 2, device="cuda", generator=g) - 0.5) * 120.0
        c = torch.randint(0, 2 ** 31 - 1, (batch, dim_y, dim_x, 3), device="cuda", dtype=torch.int32, generator=g)
    stream.synchronize()
    rows = []
    for variant in (int(x) for x in str(args.ens_variant).split(",")):
        ctx.set_option("ensemble", variant)
        for n_steps in (int(x) for x in args.ens_steps.split(",")):
            ms = timeit(stream, lambda: ctx.ensemble_step(v, c, batch, dim_x, dim_y, synth.DT, 1.0, iters, 1.96, n_steps),
                        reps=3, warm=1)
            cells = batch * n * n_steps
            r = {"kernel": f"ensemble_step[{dim_x}x{dim_y} x{batch}, K={iters}, n_steps={n_steps}, variant={variant}]",
                 "ms": round(ms, 3),
                 "mcell_steps_per_s": round(cells / (ms * 1e-3) / 1e6, 1),
                 "grid_steps_per_s": round(batch * n_steps / (ms * 1e-3), 1),
                 "hbm_GBps_state_io": round(batch * n * 40 / (ms * 1e-3) / 1e9, 1)}
            rows.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        json.dump({"rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
