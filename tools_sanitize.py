#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools_sanitize.py"""
import os
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
ctx.synchronize()
print("sanitize run complete, launches:", ctx.launch_count)
