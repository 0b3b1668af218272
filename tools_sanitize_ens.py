#!/usr/bin/env python
"""compute-sanitizer target for the ensemble kernels only (memcheck / racecheck of the register-tiled
projection's mailbox protocol): `compute-sanitizer --tool racecheck python tools_sanitize_ens.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import esp32_fluid_simulation_b200 as fb  # noqa: E402
from esp32_fluid_simulation_b200 import synth  # noqa: E402

ctx = fb.Context(0)
for dim_x, dim_y in ((80, 60), (61, 81), (13, 11)):
    nb = 3 if dim_x < 80 else 300     # 80x60: more grids than CTAs, so the pipelined flow prefetches
    bv = torch.from_numpy(np.stack([synth.velocity(dim_x, dim_y, seed=b, vmax=90.0) for b in range(nb)])).cuda()
    bc = torch.from_numpy(np.stack([synth.dye(dim_x, dim_y, seed=b, n_splats=4) for b in range(nb)]).view(np.int32)).cuda()
    for variant in (0, 5, 6, 7, 8, 9, 12, 14, 20, 21):
        ctx.set_option("ensemble", variant)
        ctx.ensemble_step(bv, bc, nb, dim_x, dim_y, synth.DT, 1.0, 3, 1.96, 2)
ctx.synchronize()
print("launches:", ctx.launch_count)
