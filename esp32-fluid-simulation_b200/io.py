"""State dump / restore.

The reference's .gitignore (lines 4-8) shows the author's uncommitted desktop harness wrote
`sim_velocity.arr`, `sim_color.arr`, `sim_pressure.arr`, `sim_divergence.arr` and `sim_params.json`;
the format itself is not in the repository, so it is defined here: each `.arr` is the raw
little-endian array in the reference's dense layout (ij = dim_x*j + i; velocity float32 x,y; colour
uint32 r,g,b; scalars float32) and `sim_params.json` carries the shape and step constants.  A
checkpoint is just velocity + colour: the pressure is not warm-started (poisson.cpp:117-119).
"""
from __future__ import annotations

import json
import os

import numpy as np

FILES = {"velocity": ("sim_velocity.arr", "<f4", 2), "color": ("sim_color.arr", "<u4", 3),
         "pressure": ("sim_pressure.arr", "<f4", 0), "divergence": ("sim_divergence.arr", "<f4", 0)}


def dump_state(directory: str, velocity, color, pressure=None, divergence=None, *, dt=1 / 30.0, dx=1.0,
               iters=10, omega=1.96, step=0) -> None:
    os.makedirs(directory, exist_ok=True)
    dim_y, dim_x = velocity.shape[:2]
    fields = {"velocity": velocity, "color": color, "pressure": pressure, "divergence": divergence}
    written = []
    for name, a in fields.items():
        if a is None:
            continue
        fname, dtype, ch = FILES[name]
        a = np.ascontiguousarray(a)
        want = (dim_y, dim_x, ch) if ch else (dim_y, dim_x)
        if a.shape != want:
            raise ValueError(f"{name}: shape {a.shape}, expected {want}")
        a.astype(dtype, copy=False).tofile(os.path.join(directory, fname))
        written.append(name)
    with open(os.path.join(directory, "sim_params.json"), "w") as f:
        json.dump({"dim_x": int(dim_x), "dim_y": int(dim_y), "dt": float(dt), "dx": float(dx),
                   "iters": int(iters), "omega": float(omega), "step": int(step), "fields": written,
                   "layout": "ij = dim_x*j + i; velocity float32[2], color uint32[3] (UQ32 raw), scalars float32; "
                             "little-endian"}, f, indent=1)


def load_state(directory: str) -> dict:
    with open(os.path.join(directory, "sim_params.json")) as f:
        params = json.load(f)
    dim_x, dim_y = params["dim_x"], params["dim_y"]
    out = {"params": params}
    for name in params["fields"]:
        fname, dtype, ch = FILES[name]
        a = np.fromfile(os.path.join(directory, fname), dtype=dtype)
        out[name] = a.reshape((dim_y, dim_x, ch) if ch else (dim_y, dim_x))
    return out
