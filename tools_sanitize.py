#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools_sanitize.py"""
import os
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # emulated ranks x 3 streams wait for each other (tests/conftest.py)
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import esp32_fluid_simulation_b200 as fb  # noqa: E402
from esp32_fluid_simulation_b200 import synth  # noqa: E402

ctx = fb.Context(0)
for dim_x, dim_y, iters in ((256, 224, 10), (61, 81, 6)):
    v = torch.from_numpy(synth.velocity(dim_x, dim_y, vmax=90.0)).cuda()
    c = torch.from_numpy(synth.dye(dim_x, dim_y).view(np.int32)).cuda()
    dr = synth.drags(dim_x, dim_y, 0, n=8)
    for sor_shape, one in ((3, 0), (2, 0), (0, 0), (3, 1)):
        ctx.set_option("sor_shape", sor_shape)
        ctx.set_option("sor_one_launch", one)
        for fuse in (1, 2, 0):
            ctx.set_option("fuse", fuse)
            ctx.step(v, c, dr, dim_x, dim_y, synth.DT, 1.0, iters, 1.96)
    ctx.set_option("sor", 0)
    ctx.step(v, c, dr, dim_x, dim_y, synth.DT, 1.0, 3, 1.96)
    ctx.set_option("sor", 1)
    ctx.set_option("advect", 0)
    ctx.step(v, c, dr, dim_x, dim_y, synth.DT, 1.0, 3, 1.96)
    ctx.set_option("advect", 1)
    out = torch.empty((dim_x - 1) * 4, (dim_y - 1) * 4, dtype=torch.int16, device="cuda")
    ctx.upscale4_rgb565(out, c, dim_x, dim_y)
    p = torch.zeros(dim_y, dim_x, device="cuda")
    d = torch.randn(dim_y, dim_x, device="cuda")
    ctx.poisson_residual(p, d, dim_x, dim_y, 1.0)
bv = torch.from_numpy(np.stack([synth.velocity(80, 60, seed=b) for b in range(3)])).cuda()
bc = torch.from_numpy(np.stack([synth.dye(80, 60, seed=b, n_splats=4) for b in range(3)]).view(np.int32)).cuda()
ctx.ensemble_step(bv, bc, 3, 80, 60, synth.DT, 1.0, 4, 1.96, 2)
for variant in (1, 4, 5, 6, 7, 8, 9, 12, 14):           # first-generation variants, register-tiled R = 2..8, streamed dye
    ctx.set_option("ensemble", variant)
    ctx.ensemble_step(bv, bc, 3, 80, 60, synth.DT, 1.0, 4, 1.96, 3)
ctx.set_option("ensemble", 0)

# round 2: frame-in-advect, device-side inputs, graph-stepped sim, strip shapes, decomposed step with fused exchanges
ctx.set_option("sor_one_launch", 0)
ctx.set_option("fuse", 5)
for sor_shape in (7, 5, 3):
    ctx.set_option("sor_shape", sor_shape)
    dim_x, dim_y = 256, 224
    v = torch.from_numpy(synth.velocity(dim_x, dim_y, vmax=90.0)).cuda()
    c = torch.from_numpy(synth.dye(dim_x, dim_y).view(np.int32)).cuda()
    c2 = torch.empty_like(c)
    frame = torch.empty((dim_x - 1) * 4, (dim_y - 1) * 4, dtype=torch.int16, device="cuda")
    ctx.step_frame(v, c, c2, frame, synth.drags(dim_x, dim_y, 0, n=8), dim_x, dim_y, synth.DT, 1.0, 14, 1.96)
ctx.set_option("sor_shape", 7)
wv = torch.empty(2, 81, 61, 2, device="cuda")
wc = torch.empty(2, 81, 61, 3, dtype=torch.int32, device="cuda")
ctx.init_color_wheel(wv, wc, 2, 61, 81)
samples = torch.from_numpy(np.random.default_rng(0).integers(0, 4000, (2, 40, 3)).astype(np.int32)).cuda()
dd, dcnt = torch.zeros(2, 10, 3, dtype=torch.int32, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda")
ctx.touch_to_drags(dd, dcnt, samples, 40, 2, 10, 61, 81)
sim = fb.Sim(ctx, 256, 192, synth.DT, 1.0, 12, 1.96, frame=True)
sim.upload(synth.velocity(256, 192), synth.dye(256, 192))
for s in range(5):
    while not sim.step(synth.drags(256, 192, s, n=4)):
        sim.acquire_frame()
        sim.release_frame()
sim.close()

from esp32_fluid_simulation_b200.dist import NativeDist  # noqa: E402
world, gx, gy = 4, 512, 384
ranks = []
for r in range(world):
    rc = fb.Context(0, torch.cuda.Stream())
    rc.set_option("sor_grid_limit", max(1, torch.cuda.get_device_properties(0).multi_processor_count // world))
    rc.set_option("halo_timeout_ms", 60000)             # everything is slow under the sanitizer
    ranks.append(NativeDist(rc, gx, gy, world, r, 13, ghost=32, advect_halo=12, frame=True))
for d in ranks:
    d.connect_local(ranks)
v0, c0 = synth.velocity(gx, gy, vmax=120.0), synth.dye(gx, gy)
for d in ranks:
    w = d.window
    d.upload(np.ascontiguousarray(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx]), np.ascontiguousarray(c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx]))
torch.cuda.synchronize()
for s in range(2):
    for d in ranks:
        d.step(synth.drags(gx, gy, s, n=6, vmax=150.0))
for d in ranks:
    d.check()
    d.close()
ctx.synchronize()
print("sanitize run complete, launches:", ctx.launch_count)
