#!/usr/bin/env python
"""Development aid: cycle breakdown of the persistent SOR kernel's tile phases (needs a build with
FS_NVCC_EXTRA=-DFS_SOR_PROF; run on a GPU box:  FS_NVCC_EXTRA=-DFS_SOR_PROF python tools_sor_phase_profile.py)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import esp32_fluid_simulation_b200 as fb  # noqa: E402

fb.build_cuda(force=True)
from esp32_fluid_simulation_b200 import _lib  # noqa: E402

L = _lib.lib()
n, iters = 4096, 50
stream = torch.cuda.Stream()
ctx = fb.Context(0, stream)
with torch.cuda.stream(stream):
    d = torch.randn(n, n, device="cuda") * 10
    p = torch.empty_like(d)
out = {}
for shape, t in ((3, 8), (5, 6), (5, 8), (7, 6)):
    ctx.set_option("sor_shape", shape)
    ctx.set_option("sor_t", t)
    for _ in range(2):
        ctx.poisson_solve(p, d, n, n, 1.0, iters, 1.96)
    ctx.synchronize()
    L.fs_debug_sor_prof(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.poisson_solve(p, d, n, n, 1.0, iters, 1.96)
    e1.record(stream)
    ctx.synchronize()
    buf = np.zeros((1024, 8), np.uint64)
    L.fs_debug_sor_prof(buf.ctypes.data_as(C.POINTER(C.c_ulonglong)), 0)
    used = buf[buf[:, 5] > 0]
    tiles = used[:, 5].sum()
    names = ["wait_tma", "stage_to_regs", "issue_prefetch", "sweeps", "store"]
    per_tile = {nm: float(used[:, i].sum() / tiles) for i, nm in enumerate(names)}
    per_tile["sum"] = sum(per_tile.values())
    row = {"shape": shape, "T": t, "solve_ms": e0.elapsed_time(e1), "ctas": int(len(used)), "tiles": int(tiles),
           "cycles_per_tile": per_tile,
           "cta_busy_cycles_mean": float(used[:, :5].sum(axis=1).mean()), "cta_busy_cycles_max": float(used[:, :5].sum(axis=1).max())}
    out[f"shape{shape}_T{t}"] = row
    print(json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_sor_phase_profile.json"), "w"), indent=1)
