"""Host-side mirror of the reference's sim-step operator surface.

Same names, argument order and meaning as the reference's free functions
(advect.h:74-76, finitediff.h:6-10, poisson.h:4-5), so parity tests read like
calls into the reference:

    advect(next_p, p, vel, dim_x, dim_y, dt, no_slip)
    calculate_divergence(div, v, dim_x, dim_y, dx)
    subtract_gradient(v, p, dim_x, dim_y, dx)
    poisson_solve(p, div, dim_x, dim_y, dx, iters, omega)

Arrays are caller-owned, dense, in the reference layout (ij = dim_x*j + i).
torch CUDA tensors go through the device-pointer entry points (``fs_*``,
asynchronous on the context's stream); numpy arrays go through the host-pointer
drop-ins (``fsh_*``: copy in, same kernels, copy out).  ``advect`` dispatches on
the payload type like the reference's template: float32 -> Vector2<float>,
uint32 -> Vector3<UQ32>.

torch is used for device memory and streams only; no torch op is on the path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import HaloCopy, Tile, check

DRAG_DTYPE = np.dtype([("cx", "<u2"), ("cy", "<u2"), ("vx", "<f4"), ("vy", "<f4")])  # ino:45-48


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


def _ptr(a, dtype_name: str, n_elems: int | None = None) -> tuple[int, bool]:
    """(address, on_device).  Validates dtype / contiguity / device."""
    if a is None:
        return 0, False
    if _is_torch(a):
        import torch
        # torch.uint32 / torch.uint16 only exist from torch 2.3 on: look them up lazily
        want = {"float32": torch.float32, "uint32": getattr(torch, "uint32", None), "int32": torch.int32,
                "uint16": getattr(torch, "uint16", None), "int16": torch.int16}[dtype_name]
        alt = {"uint32": torch.int32, "uint16": torch.int16}.get(dtype_name)
        if a.dtype != want and a.dtype != alt:
            raise TypeError(f"expected {dtype_name} tensor, got {a.dtype}")
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous (dense reference layout)")
        if n_elems is not None and a.numel() < n_elems:
            raise ValueError(f"tensor has {a.numel()} elements, need {n_elems}")
        if not a.is_cuda:
            return a.data_ptr(), False
        return a.data_ptr(), True
    if not isinstance(a, np.ndarray):
        raise TypeError(f"unsupported array type {type(a)}")
    if a.dtype != np.dtype(dtype_name):
        raise TypeError(f"expected {dtype_name} array, got {a.dtype}")
    if not a.flags.c_contiguous:
        raise ValueError("array must be C-contiguous (dense reference layout)")
    if n_elems is not None and a.size < n_elems:
        raise ValueError(f"array has {a.size} elements, need {n_elems}")
    return a.ctypes.data, False


def _payload(a) -> str:
    if _is_torch(a):
        import torch
        return "vec2f" if a.dtype == torch.float32 else "rgb"
    return "vec2f" if a.dtype == np.float32 else "rgb"


def _drags(drags):
    if drags is None:
        return None, 0
    d = np.ascontiguousarray(drags, DRAG_DTYPE)
    return d, len(d)


class Context:
    """fs_ctx: device + stream + scratch.  ``stream`` is a ``torch.cuda.Stream``,
    a raw ``cudaStream_t`` integer or None (legacy default stream)."""

    def __init__(self, device: int = 0, stream=None):
        self._L = _lib.lib()
        handle = C.c_void_p()
        s = 0
        if stream is not None:
            s = getattr(stream, "cuda_stream", stream)
        check(self._L.fs_ctx_create(C.byref(handle), int(device), C.c_void_p(s)), "fs_ctx_create")
        self._h = handle
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fs_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._L.fs_ctx_synchronize(self._h), "fs_ctx_synchronize")

    def set_option(self, name: str, value: int):
        check(self._L.fs_ctx_set_option(self._h, name.encode(), int(value)), f"set_option({name})")

    def get_option(self, name: str) -> int:
        v = C.c_int()
        check(self._L.fs_ctx_get_option(self._h, name.encode(), C.byref(v)), f"get_option({name})")
        return v.value

    @property
    def launch_count(self) -> int:
        return int(self._L.fs_ctx_launch_count(self._h))

    # --- the reference's operator surface ------------------------------------------
    def advect(self, next_p, p, vel, dim_x, dim_y, dt, no_slip):
        n = dim_x * dim_y
        kind = _payload(p)
        dt_name, nc = ("float32", 2) if kind == "vec2f" else ("uint32", 3)
        a_out, dev0 = _ptr(next_p, dt_name, n * nc)
        a_in, dev1 = _ptr(p, dt_name, n * nc)
        a_vel, dev2 = _ptr(vel, "float32", n * 2)
        if not (dev0 == dev1 == dev2):
            raise ValueError("advect: arrays must all be on the host or all on the device")
        pre = "fs_" if dev0 else "fsh_"
        fn = getattr(self._L, pre + ("advect_vec2f" if kind == "vec2f" else "advect_rgb_uq32"))
        check(fn(a_out, a_in, a_vel, dim_x, dim_y, dt, int(bool(no_slip)), self._h), fn.__name__)

    def calculate_divergence(self, div, v, dim_x, dim_y, dx):
        n = dim_x * dim_y
        a_div, d0 = _ptr(div, "float32", n)
        a_v, d1 = _ptr(v, "float32", 2 * n)
        if d0 != d1:
            raise ValueError("calculate_divergence: mixed host/device arrays")
        fn = self._L.fs_calculate_divergence if d0 else self._L.fsh_calculate_divergence
        check(fn(a_div, a_v, dim_x, dim_y, dx, self._h), "calculate_divergence")

    def subtract_gradient(self, v, p, dim_x, dim_y, dx):
        n = dim_x * dim_y
        a_v, d0 = _ptr(v, "float32", 2 * n)
        a_p, d1 = _ptr(p, "float32", n)
        if d0 != d1:
            raise ValueError("subtract_gradient: mixed host/device arrays")
        fn = self._L.fs_subtract_gradient if d0 else self._L.fsh_subtract_gradient
        check(fn(a_v, a_p, dim_x, dim_y, dx, self._h), "subtract_gradient")

    def poisson_solve(self, p, div, dim_x, dim_y, dx, iters, omega):
        n = dim_x * dim_y
        a_p, d0 = _ptr(p, "float32", n)
        a_d, d1 = _ptr(div, "float32", n)
        if d0 != d1:
            raise ValueError("poisson_solve: mixed host/device arrays")
        fn = self._L.fs_poisson_solve if d0 else self._L.fsh_poisson_solve
        check(fn(a_p, a_d, dim_x, dim_y, dx, iters, omega, self._h), "poisson_solve")

    # --- the rest of loop() ------------------------------------------------------------
    def sor_half_sweep(self, p, div, dim_x, dim_y, dx, omega, parity):
        n = dim_x * dim_y
        a_p, d0 = _ptr(p, "float32", n)
        a_d, d1 = _ptr(div, "float32", n)
        if not (d0 and d1):
            raise ValueError("sor_half_sweep: device tensors only")
        check(self._L.fs_sor_half_sweep(a_p, a_d, dim_x, dim_y, dx, omega, parity, self._h),
              "fs_sor_half_sweep")

    def poisson_residual(self, p, div, dim_x, dim_y, dx) -> tuple[float, float]:
        """(max |r|, ||r||_2) of r = gs(p) - p; device tensors."""
        n = dim_x * dim_y
        a_p, d0 = _ptr(p, "float32", n)
        a_d, d1 = _ptr(div, "float32", n)
        if not (d0 and d1):
            raise ValueError("poisson_residual: device tensors only")
        m, l2 = C.c_float(), C.c_double()
        check(self._L.fs_poisson_residual(C.byref(m), C.byref(l2), a_p, a_d, dim_x, dim_y, dx, self._h),
              "fs_poisson_residual")
        return m.value, l2.value

    def apply_drags(self, v, drags, dim_x, dim_y):
        a_v, d0 = _ptr(v, "float32", 2 * dim_x * dim_y)
        if not d0:
            raise ValueError("apply_drags: v must be a device tensor")
        d, n = _drags(drags)
        check(self._L.fs_apply_drags(a_v, d.ctypes.data if n else None, n, dim_x, dim_y, self._h),
              "fs_apply_drags")

    def step(self, v, c, drags, dim_x, dim_y, dt, dx, iters, omega, p_out=None, div_out=None):
        """loop() body (ino:249-289), in place on v and c."""
        n = dim_x * dim_y
        a_v, d0 = _ptr(v, "float32", 2 * n)
        a_c, d1 = _ptr(c, "uint32", 3 * n)
        a_p, d2 = _ptr(p_out, "float32", n)
        a_d, d3 = _ptr(div_out, "float32", n)
        if d0 != d1 or (p_out is not None and d2 != d0) or (div_out is not None and d3 != d0):
            raise ValueError("step: mixed host/device arrays")
        d, nd = _drags(drags)
        fn = self._L.fs_step if d0 else self._L.fsh_step
        check(fn(a_v, a_c, d.ctypes.data if nd else None, nd, dim_x, dim_y, dt, dx, iters, omega,
                 a_p or None, a_d or None, self._h), "step")

    def advect_drags_divergence(self, v_out, div, v_in, drags, dim_x, dim_y, dt, dx):
        """ino:253 + 264-269 + 274 in one pass (device tensors): v_out = forced advected velocity, div = its divergence."""
        n = dim_x * dim_y
        a_o, d0 = _ptr(v_out, "float32", 2 * n)
        a_d, d1 = _ptr(div, "float32", n)
        a_i, d2 = _ptr(v_in, "float32", 2 * n)
        if not (d0 and d1 and d2):
            raise ValueError("advect_drags_divergence: device tensors only")
        d, nd = _drags(drags)
        check(self._L.fs_advect_drags_divergence(a_o, a_d, a_i, d.ctypes.data if nd else None, nd, dim_x, dim_y, dt, dx,
                                                 self._h), "fs_advect_drags_divergence")

    def advect_rgb_frame(self, next_c, frame, c, vel, dim_x, dim_y, dt, no_slip=False):
        """Dye advect + the 4x RGB565 frame of the advected dye in one kernel (device tensors)."""
        n = dim_x * dim_y
        check(self._L.fs_advect_rgb_frame(_ptr(next_c, "uint32", 3 * n)[0], _ptr(frame, "uint16", 16 * (dim_x - 1) * (dim_y - 1))[0],
                                          _ptr(c, "uint32", 3 * n)[0], _ptr(vel, "float32", 2 * n)[0], dim_x, dim_y, dt,
                                          int(bool(no_slip)), self._h), "fs_advect_rgb_frame")

    def step_frame(self, v, c_in, c_out, frame, drags, dim_x, dim_y, dt, dx, iters, omega):
        """loop() + draw_routine's arithmetic: the step, and the RGB565 frame of its new dye (device tensors)."""
        n = dim_x * dim_y
        d, nd = _drags(drags)
        check(self._L.fs_step_frame(_ptr(v, "float32", 2 * n)[0], _ptr(c_in, "uint32", 3 * n)[0], _ptr(c_out, "uint32", 3 * n)[0],
                                    _ptr(frame, "uint16", 16 * (dim_x - 1) * (dim_y - 1))[0], d.ctypes.data if nd else None, nd,
                                    dim_x, dim_y, dt, dx, iters, omega, None, None, self._h), "fs_step_frame")

    def step_pingpong(self, v, c_in, c_out, drags, dim_x, dim_y, dt, dx, iters, omega, p_out=None, div_out=None):
        """loop() body with the dye going c_in -> c_out (the caller swaps, ino:286); device tensors."""
        n = dim_x * dim_y
        a_v, d0 = _ptr(v, "float32", 2 * n)
        a_ci, d1 = _ptr(c_in, "uint32", 3 * n)
        a_co, d2 = _ptr(c_out, "uint32", 3 * n)
        a_p, _ = _ptr(p_out, "float32", n)
        a_d, _ = _ptr(div_out, "float32", n)
        if not (d0 and d1 and d2):
            raise ValueError("step_pingpong: device tensors only")
        d, nd = _drags(drags)
        check(self._L.fs_step_pingpong(a_v, a_ci, a_co, d.ctypes.data if nd else None, nd, dim_x, dim_y, dt, dx,
                                       iters, omega, a_p or None, a_d or None, self._h), "fs_step_pingpong")

    def upscale4_rgb565(self, out, c, dim_x, dim_y):
        a_o, d0 = _ptr(out, "uint16", 16 * (dim_x - 1) * (dim_y - 1))
        a_c, d1 = _ptr(c, "uint32", 3 * dim_x * dim_y)
        if d0 != d1:
            raise ValueError("upscale4_rgb565: mixed host/device arrays")
        fn = self._L.fs_upscale4_rgb565 if d0 else self._L.fsh_upscale4_rgb565
        check(fn(a_o, a_c, dim_x, dim_y, self._h), "upscale4_rgb565")

    def ensemble_step(self, v, c, batch, dim_x, dim_y, dt, dx, iters, omega, n_steps=1,
                      drags=None, drag_counts=None, max_drags=0):
        n = batch * dim_x * dim_y
        a_v, d0 = _ptr(v, "float32", 2 * n)
        a_c, d1 = _ptr(c, "uint32", 3 * n)
        if not (d0 and d1):
            raise ValueError("ensemble_step: device tensors only")
        dptr = cptr = None
        if max_drags > 0:
            d = np.ascontiguousarray(drags, DRAG_DTYPE)
            cnt = np.ascontiguousarray(drag_counts, np.int32)
            assert d.size == n_steps * batch * max_drags and cnt.size == n_steps * batch
            dptr, cptr = d.ctypes.data, cnt.ctypes.data
        check(self._L.fs_ensemble_step(a_v, a_c, dptr, cptr, max_drags, batch, dim_x, dim_y, dt, dx,
                                       iters, omega, n_steps, self._h), "fs_ensemble_step")

    def ensemble_step_dev(self, v, c, drags_dev, counts_dev, max_drags, batch, dim_x, dim_y, dt, dx, iters, omega,
                          n_steps=1):
        """fs_ensemble_step with device-resident drag records (int32 view of [n_steps, batch, max_drags, 3] words)."""
        n = batch * dim_x * dim_y
        a_v, _ = _ptr(v, "float32", 2 * n)
        a_c, _ = _ptr(c, "uint32", 3 * n)
        check(self._L.fs_ensemble_step_dev(a_v, a_c, drags_dev.data_ptr(), counts_dev.data_ptr(), max_drags, batch, dim_x,
                                           dim_y, dt, dx, iters, omega, n_steps, self._h), "fs_ensemble_step_dev")

    def init_color_wheel(self, v, c, batch, dim_x, dim_y):
        """setup(), ino:196-241, on the device (device tensors, `batch` grids back to back)."""
        n = batch * dim_x * dim_y
        a_v, d0 = _ptr(v, "float32", 2 * n)
        a_c, d1 = _ptr(c, "uint32", 3 * n)
        if not (d0 and d1):
            raise ValueError("init_color_wheel: device tensors only")
        check(self._L.fs_init_color_wheel(a_v, a_c, batch, dim_x, dim_y, self._h), "fs_init_color_wheel")

    def touch_to_drags(self, drags_out, counts_out, samples, n_samples, batch, max_drags, n_rows, n_cols):
        """touch_routine(), ino:63-96, on the device.  drags_out: int32 device tensor [batch, max_drags, 3] (12-byte
        records), counts_out: int32 [batch], samples: int32 [batch, n_samples, 3]."""
        check(self._L.fs_touch_to_drags(drags_out.data_ptr(), counts_out.data_ptr(), samples.data_ptr(), n_samples, batch,
                                        max_drags, n_rows, n_cols, None, self._h), "fs_touch_to_drags")

    # --- decomposed grids ----------------------------------------------------------------
    def tile_advect(self, next_p, p, vel, tile: Tile, dt, no_slip):
        kind = _payload(p)
        dt_name = "float32" if kind == "vec2f" else "uint32"
        fn = self._L.fs_tile_advect_vec2f if kind == "vec2f" else self._L.fs_tile_advect_rgb_uq32
        check(fn(_ptr(next_p, dt_name)[0], _ptr(p, dt_name)[0], _ptr(vel, "float32")[0],
                 C.byref(tile), dt, int(bool(no_slip)), self._h), "fs_tile_advect")

    def tile_calculate_divergence(self, div, v, tile: Tile, dx):
        check(self._L.fs_tile_calculate_divergence(_ptr(div, "float32")[0], _ptr(v, "float32")[0],
                                                   C.byref(tile), dx, self._h),
              "fs_tile_calculate_divergence")

    def tile_subtract_gradient(self, v, p, tile: Tile, dx):
        check(self._L.fs_tile_subtract_gradient(_ptr(v, "float32")[0], _ptr(p, "float32")[0],
                                                C.byref(tile), dx, self._h),
              "fs_tile_subtract_gradient")

    def tile_sor_sweeps(self, p_out, p_in, div, tile: Tile, dx, omega, first_parity, n_half):
        check(self._L.fs_tile_sor_sweeps(_ptr(p_out, "float32")[0],
                                         _ptr(p_in, "float32")[0] or None,
                                         _ptr(div, "float32")[0], C.byref(tile), dx, omega,
                                         first_parity, n_half, self._h), "fs_tile_sor_sweeps")

    def tile_apply_drags(self, v, drags, tile: Tile):
        d, n = _drags(drags)
        check(self._L.fs_tile_apply_drags(_ptr(v, "float32")[0], d.ctypes.data if n else None, n,
                                          C.byref(tile), self._h), "fs_tile_apply_drags")

    def tile_check(self):
        check(self._L.fs_tile_check(self._h), "fs_tile_check")

    def tile_max_displacement(self, vel, tile: Tile, dt) -> int:
        out = C.c_int()
        check(self._L.fs_tile_max_displacement(C.byref(out), _ptr(vel, "float32")[0], C.byref(tile),
                                               dt, self._h), "fs_tile_max_displacement")
        return out.value


    # --- halo exchange over NVLink peer memory ---------------------------------------------------
    def arena_alloc(self, nbytes: int) -> int:
        out = C.c_void_p()
        check(self._L.fs_arena_alloc(C.byref(out), nbytes, self._h), "fs_arena_alloc")
        return out.value

    def arena_free(self, ptr: int):
        check(self._L.fs_arena_free(C.c_void_p(ptr), self._h), "fs_arena_free")

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        check(self._L.fs_ipc_export(C.c_void_p(ptr), buf, self._h), "fs_ipc_export")
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        out = C.c_void_p()
        check(self._L.fs_ipc_open(C.byref(out), C.create_string_buffer(handle, 64), self._h), "fs_ipc_open")
        return out.value

    def ipc_close(self, ptr: int):
        check(self._L.fs_ipc_close(C.c_void_p(ptr), self._h), "fs_ipc_close")

    def halo_exchange(self, copies, signal_flags, wait_flags, seq: int):
        """copies: [(src, dst, src_pitch, dst_pitch, row_bytes, rows)], flags: lists of addresses."""
        n, k = len(copies), len(signal_flags)
        arr = (HaloCopy * max(n, 1))(*[HaloCopy(*c) for c in copies])
        sig = (C.c_void_p * max(k, 1))(*signal_flags)
        wai = (C.c_void_p * max(k, 1))(*wait_flags)
        check(self._L.fs_halo_exchange(arr, n, sig, wai, k, seq, self._h), "fs_halo_exchange")

    def set_stream(self, stream):
        s = getattr(stream, "cuda_stream", stream) or 0
        check(self._L.fs_ctx_set_stream(self._h, C.c_void_p(s)), "fs_ctx_set_stream")


class Sim:
    """fs_sim: a device-resident simulation stepped by one CUDA-graph launch per loop() body, with the sketch's
    colour hand-off (ino:285-288) as a double-buffered asynchronous frame stream into pinned host memory."""

    def __init__(self, ctx: Context, dim_x: int, dim_y: int, dt, dx, iters: int, omega, frame: bool = False):
        self.ctx, self._L = ctx, ctx._L
        self.dim_x, self.dim_y = dim_x, dim_y
        h = C.c_void_p()
        check(self._L.fs_sim_create(C.byref(h), dim_x, dim_y, dt, dx, iters, omega, int(bool(frame)), ctx._h), "fs_sim_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None):
                self._L.fs_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def upload(self, v, c):
        check(self._L.fs_sim_upload(self._h, _ptr(v, "float32")[0], _ptr(c, "uint32")[0]), "fs_sim_upload")

    def download(self):
        n = self.dim_x * self.dim_y
        v = np.empty((self.dim_y, self.dim_x, 2), np.float32)
        c = np.empty((self.dim_y, self.dim_x, 3), np.uint32)
        p, d = np.empty((self.dim_y, self.dim_x), np.float32), np.empty((self.dim_y, self.dim_x), np.float32)
        check(self._L.fs_sim_download(self._h, v.ctypes.data, c.ctypes.data, p.ctypes.data, d.ctypes.data), "fs_sim_download")
        assert v.size == 2 * n
        return v, c, p, d

    def step(self, drags=None) -> bool:
        """False when both frame slots still wait for the consumer (FS_ERR_WOULD_BLOCK); True after a step."""
        d, nd = _drags(drags)
        code = self._L.fs_sim_step(self._h, d.ctypes.data if nd else None, nd)
        if code == _lib.FS_ERR_WOULD_BLOCK:
            return False
        check(code, "fs_sim_step")
        return True

    def acquire_frame(self):
        """The oldest unconsumed frame as a numpy VIEW of the pinned host buffer (valid until release_frame), or None."""
        ptr, rows, cols = C.c_void_p(), C.c_int(), C.c_int()
        code = self._L.fs_sim_acquire_frame(self._h, C.byref(ptr), C.byref(rows), C.byref(cols))
        if code == _lib.FS_ERR_WOULD_BLOCK:
            return None
        check(code, "fs_sim_acquire_frame")
        buf = (C.c_uint16 * (rows.value * cols.value)).from_address(ptr.value)
        return np.frombuffer(buf, np.uint16).reshape(rows.value, cols.value)

    def release_frame(self):
        check(self._L.fs_sim_release_frame(self._h), "fs_sim_release_frame")

    @property
    def stats(self) -> dict:
        a, b, c = C.c_ulonglong(), C.c_ulonglong(), C.c_ulonglong()
        check(self._L.fs_sim_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "fs_sim_stats")
        return {"steps": a.value, "graph_launches": b.value, "eager_steps": c.value}


# ---- module-level functions with the reference's exact names -------------------------
_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def _ctx_for(a, ctx):
    if ctx is not None:
        return ctx
    dev = a.device.index if _is_torch(a) and a.is_cuda else 0
    return default_context(dev or 0)


def advect(next_p, p, vel, dim_x, dim_y, dt, no_slip, ctx: Context | None = None):
    """advect<T,U>, advect.h:74-85."""
    _ctx_for(p, ctx).advect(next_p, p, vel, dim_x, dim_y, dt, no_slip)


def calculate_divergence(div, v, dim_x, dim_y, dx, ctx: Context | None = None):
    """calculate_divergence, finitediff.cpp:33-39."""
    _ctx_for(v, ctx).calculate_divergence(div, v, dim_x, dim_y, dx)


def subtract_gradient(v, p, dim_x, dim_y, dx, ctx: Context | None = None):
    """subtract_gradient, finitediff.cpp:75-82 (v in place)."""
    _ctx_for(v, ctx).subtract_gradient(v, p, dim_x, dim_y, dx)


def poisson_solve(p, div, dim_x, dim_y, dx, iters, omega, ctx: Context | None = None):
    """poisson_solve, poisson.cpp:114-125."""
    _ctx_for(p, ctx).poisson_solve(p, div, dim_x, dim_y, dx, iters, omega)
