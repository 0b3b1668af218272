"""Helpers for the -m gpu tier: move reference-layout numpy arrays to/from cuda:0."""
import numpy as np
import torch


def to_dev(a: np.ndarray) -> torch.Tensor:
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint32:
        a = a.view(np.int32)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    return torch.from_numpy(a).cuda()


def to_host(t: torch.Tensor, dtype=None) -> np.ndarray:
    a = t.cpu().numpy()
    if dtype is not None:
        a = a.view(dtype)
    return a


def rand_fields(seed, dim_x, dim_y, vmax):
    rng = np.random.default_rng(seed)
    v = ((rng.random((dim_y, dim_x, 2), np.float32) - np.float32(0.5)) * np.float32(2 * vmax))
    c = rng.integers(0, 2 ** 32, (dim_y, dim_x, 3), dtype=np.uint32)
    return v.astype(np.float32), c


def rand_drags(seed, dim_x, dim_y, n, oob=False):
    from esp32_fluid_simulation_b200.synth import DRAG_DTYPE
    rng = np.random.default_rng(seed)
    d = np.zeros(n, DRAG_DTYPE)
    d["cx"] = rng.integers(0, dim_y + (3 if oob else 0), n)
    d["cy"] = rng.integers(0, dim_x + (3 if oob else 0), n)
    d["vx"] = rng.normal(0, 400, n)
    d["vy"] = rng.normal(0, 400, n)
    return d
