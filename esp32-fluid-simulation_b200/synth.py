"""Seeded synthetic inputs for the step: random velocity, dye
splats, touch-force impulses.  Counter-based (splitmix64 of seed + element
index), so any sub-window of a big grid can be generated independently and every
rank of a decomposed run sees the same global field.  numpy only."""
from __future__ import annotations

import numpy as np

DRAG_DTYPE = np.dtype([("cx", "<u2"), ("cy", "<u2"), ("vx", "<f4"), ("vy", "<f4")])  # ino:45-48

DT = np.float32(1 / 30.0)      # ino:16
DX = np.float32(1.0)           # ino:274-276
OMEGA = np.float32(1.96)       # ino:275
DYE_CAP = 0xFFFFF800           # two lerps of values <= this cannot reach 2^32


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def _node_ids(dim_x, x0, y0, nx, ny):
    j = np.arange(y0, y0 + ny, dtype=np.uint64)[:, None]
    i = np.arange(x0, x0 + nx, dtype=np.uint64)[None, :]
    return j * np.uint64(dim_x) + i


def _unit(h: np.ndarray) -> np.ndarray:
    """uint64 hash -> float32 in [0,1) with 24 random bits."""
    return ((h >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / (1 << 24))


def velocity(dim_x, dim_y, seed=0xF1D0, vmax=60.0, window=None) -> np.ndarray:
    """float32[ny, nx, 2], i.i.d. uniform in [-vmax, vmax] nodes/s per component
    (|displacement| <= 2 nodes/step at vmax=60, dt=1/30: the CFL-bounded regime)."""
    x0, y0, nx, ny = window or (0, 0, dim_x, dim_y)
    ids = _node_ids(dim_x, x0, y0, nx, ny)
    out = np.empty((ny, nx, 2), np.float32)
    with np.errstate(over="ignore"):
        for ch in range(2):
            h = splitmix64(ids * np.uint64(2) + np.uint64(ch) + np.uint64(seed) * np.uint64(0x100000001))
            out[..., ch] = (_unit(h) * np.float32(2.0) - np.float32(1.0)) * np.float32(vmax)
    return out


def dye(dim_x, dim_y, seed=0xD1E, n_splats=64, window=None, saturate=False) -> np.ndarray:
    """uint32[ny, nx, 3]: background hash(ij) & 0xFFFF0000 scaled into the lower
    half of the range, plus Gaussian splats (sigma = min(dim)/32, peak 0xFFFF0000,
    random RGB), capped at DYE_CAP.  saturate=True instead fills with 0xFFFFFFFF
    speckles for the conversion-overflow known-answer case."""
    x0, y0, nx, ny = window or (0, 0, dim_x, dim_y)
    ids = _node_ids(dim_x, x0, y0, nx, ny)
    acc = np.empty((ny, nx, 3), np.float64)
    with np.errstate(over="ignore"):
        for ch in range(3):
            h = splitmix64(ids * np.uint64(3) + np.uint64(ch) + np.uint64(seed) * np.uint64(0x100000001))
            acc[..., ch] = ((h >> np.uint64(33)) & np.uint64(0x7FFF0000)).astype(np.float64)
    sigma = max(1.0, min(dim_x, dim_y) / 32.0)
    meta = splitmix64(np.arange(n_splats * 5, dtype=np.uint64) + np.uint64(seed) * np.uint64(7919))
    u = _unit(meta).astype(np.float64).reshape(n_splats, 5)
    r = int(np.ceil(4 * sigma))
    for k in range(n_splats):
        cx, cy = u[k, 0] * (dim_x - 1), u[k, 1] * (dim_y - 1)
        ax0, ax1 = max(x0, int(cx) - r), min(x0 + nx, int(cx) + r + 1)
        ay0, ay1 = max(y0, int(cy) - r), min(y0 + ny, int(cy) + r + 1)
        if ax0 >= ax1 or ay0 >= ay1:
            continue
        gx = np.exp(-0.5 * ((np.arange(ax0, ax1) - cx) / sigma) ** 2)[None, :]
        gy = np.exp(-0.5 * ((np.arange(ay0, ay1) - cy) / sigma) ** 2)[:, None]
        g = gx * gy * float(0xFFFF0000)
        for ch in range(3):
            acc[ay0 - y0:ay1 - y0, ax0 - x0:ax1 - x0, ch] += g * u[k, 2 + ch]
    out = np.minimum(acc, float(DYE_CAP)).astype(np.uint32)
    if saturate:
        mask = (splitmix64(ids + np.uint64(seed)) & np.uint64(3)) == 0
        out[mask] = np.uint32(0xFFFFFFFF)
    return out


def drags(dim_x, dim_y, step, n=16, seed=0xD4A6, vmax=1000.0) -> np.ndarray:
    """`n` touch impulses for step `step`: uniform in-range nodes, velocity uniform
    in [-vmax, vmax] nodes/s (far beyond any halo: exercises the gather fallback)."""
    base = np.arange(n * 4, dtype=np.uint64) + np.uint64(step) * np.uint64(1 << 20) + \
        np.uint64(seed) * np.uint64(0x100000001)
    with np.errstate(over="ignore"):
        u = _unit(splitmix64(base)).reshape(n, 4)
    out = np.zeros(n, DRAG_DTYPE)
    out["cx"] = np.minimum((u[:, 0] * dim_y).astype(np.int64), dim_y - 1)  # column = slow axis j
    out["cy"] = np.minimum((u[:, 1] * dim_x).astype(np.int64), dim_x - 1)  # row = fast axis i
    out["vx"] = (u[:, 2] * 2 - 1) * np.float32(vmax)
    out["vy"] = (u[:, 3] * 2 - 1) * np.float32(vmax)
    return out


# ---------------------------------------------------------------------------------------------
# The reference's own inputs, restated host-side: they run on the MCU's CPU in the
# reference too and are not part of the hot path.
# ---------------------------------------------------------------------------------------------

def _uq32(x: np.ndarray) -> np.ndarray:
    """UQ32(float), uq32.h:13, saturating like the CUDA path."""
    y = x.astype(np.float32) + np.float32(0.5)
    return np.where(y >= np.float32(4294967296.0), np.uint64(0xFFFFFFFF),
                    np.maximum(y, 0).astype(np.uint64)).astype(np.uint32)


def color_wheel(dim_x: int, dim_y: int):
    """setup(), ino:196-241: zero velocity; three-sector colour wheel (atan2f of the node's offset from
    the centre, sectors split at +-pi/3); then two IN-PLACE 1-2-1 smoothing passes, first along j, then
    along i, each walked in index order so the lower-index neighbour is already smoothed.  Returns
    (velocity float32[dim_y, dim_x, 2], dye uint32[dim_y, dim_x, 3]) in the reference layout
    (dim_x = N_ROWS is the fast axis)."""
    f32 = np.float32
    i = np.arange(dim_x, dtype=np.int64)[None, :]
    j = np.arange(dim_y, dtype=np.int64)[:, None]
    ang = np.arctan2((-(i - dim_x // 2)).astype(f32), (j - dim_y // 2).astype(f32)).astype(np.float64)
    third = 3.1415926535897932384626433832795 / 3
    full = _uq32(np.array([4294967295.0], f32))[0]               # Vector3<float>(UINT32_MAX, ...) -> UQ32
    c = np.zeros((dim_y, dim_x, 3), np.uint32)
    red, green = ang < -third, (ang >= -third) & (ang < third)
    c[..., 0][red] = full
    c[..., 1][green] = full
    c[..., 2][~(red | green)] = full
    q, h = f32(0.25), f32(0.5)
    for jj in range(dim_y):                                      # ino:220-230 (sequential in j)
        mid = c[jj].astype(f32)
        lo = c[jj - 1].astype(f32) if jj > 0 else mid
        hi = c[jj + 1].astype(f32) if jj < dim_y - 1 else mid
        c[jj] = _uq32(q * lo + h * mid + q * hi)
    for ii in range(dim_x):                                      # ino:231-241 (sequential in i)
        mid = c[:, ii].astype(f32)
        lo = c[:, ii - 1].astype(f32) if ii > 0 else mid
        hi = c[:, ii + 1].astype(f32) if ii < dim_x - 1 else mid
        c[:, ii] = _uq32(q * lo + h * mid + q * hi)
    return np.zeros((dim_y, dim_x, 2), np.float32), c


def arduino_map(x: int, in_min: int, in_max: int, out_min: int, out_max: int) -> int:
    """Arduino map(): long arithmetic, division truncating toward zero."""
    num = (x - in_min) * (out_max - out_min)
    den = in_max - in_min
    quo = abs(num) // abs(den)
    return (quo if (num >= 0) == (den > 0) else -quo) + out_min


def touch_drags(samples, n_rows: int, n_cols: int, polling_ms: int = 10,
                cal=(200, 3700, 240, 3800)) -> np.ndarray:
    """touch_routine(), ino:63-96: `samples` is one (touched, raw_x, raw_y) per polling period.  Every
    touched sample that follows a touched sample yields a drag record {coords, velocity} with
    coords = map() of the raw reading onto [0, N_COLS] x [0, N_ROWS] (ino:77-78; the upper bound is
    inclusive in the reference — such records fall outside the grid and fs_apply_drags drops them) and
    velocity = delta_coords * 1000.f / POLLING_PERIOD nodes/s (ino:82-83).  Queue depth / dropping
    (ino:49,85) is the caller's business."""
    out, last, last_touched = [], None, False
    for touched, rx, ry in samples:
        if touched:
            cx = arduino_map(int(rx), cal[0], cal[1], 0, n_cols)
            cy = arduino_map(int(ry), cal[2], cal[3], 0, n_rows)
            if last_touched:
                vx = np.float32(np.float32(cx - last[0]) * np.float32(1000.0)) / np.float32(polling_ms)
                vy = np.float32(np.float32(cy - last[1]) * np.float32(1000.0)) / np.float32(polling_ms)
                out.append((cx & 0xFFFF, cy & 0xFFFF, vx, vy))
            last = (cx, cy)
        last_touched = bool(touched)
    return np.array(out, DRAG_DTYPE) if out else np.zeros(0, DRAG_DTYPE)


def fnv1a64(a: np.ndarray) -> int:
    """FNV-1a-64 over the raw bytes (small arrays only; pure Python loop)."""
    h = 0xcbf29ce484222325
    for b in np.ascontiguousarray(a).view(np.uint8).ravel().tolist():
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h
