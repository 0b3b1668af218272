"""TEST INFRASTRUCTURE — ctypes loaders for the CPU checkers.

  * ``Oracle``  -> oracle/libfluid_oracle.so  (plain-C restatement, fluid_oracle.c)
  * ``Ref``     -> oracle/_ref/libfluid_ref.so (the reference's own sources, ref_shim.cpp)

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``esp32-fluid-simulation_b200``) never does; it fails loudly without its CUDA
library instead of falling back to anything here.

Both classes expose the same method names over numpy arrays in the reference
layout: velocity ``float32[dim_y, dim_x, 2]``, dye
``uint32[dim_y, dim_x, 3]``, scalars ``float32[dim_y, dim_x]`` — i.e. node (i,j)
at ``ij = dim_x*j + i`` (operations.h:7-9).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libfluid_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libfluid_ref.so")
REFERENCE_SRC = "/root/reference/ESP32-fluid-simulation"

DRAG_DTYPE = np.dtype([("cx", "<u2"), ("cy", "<u2"), ("vx", "<f4"), ("vy", "<f4")])
assert DRAG_DTYPE.itemsize == 12  # struct drag, ino:45-48


def build(ref: bool | None = None) -> None:
    """Compile the checkers (building the checker is not using it)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir(REFERENCE_SRC)
    if ref:
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _f32(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _u32(a):
    assert a.dtype == np.uint32 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _dims(a):
    return int(a.shape[1]), int(a.shape[0])  # dim_x (fast), dim_y (slow)


class OracleTile(C.Structure):
    """oracle_tile — same fields as fs_tile (include/fluid_b200.h)."""
    _fields_ = [(n, C.c_int) for n in
                ("gdim_x", "gdim_y", "ox", "oy", "nx", "ny", "x0", "y0", "x1", "y1")]


class _Base:
    prefix = ""
    path = ""

    def __init__(self):
        if not os.path.exists(self.path):
            raise FileNotFoundError(f"{self.path} not built; run `make -C oracle`")
        self.lib = C.CDLL(self.path)
        L, p = self.lib, self.prefix
        F, U, I, f = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int, C.c_float
        self._fn = {}
        for name, args, res in [
            ("advect_vec2f", [F, F, F, I, I, f, I], None),
            ("advect_rgb_uq32", [U, U, F, I, I, f, I], None),
            ("sample_vec2f", [F, F, f, f, I, I, I], None),
            ("sample_rgb_uq32", [U, U, f, f, I, I, I], None),
            ("uq32_from_float", [f], C.c_uint32),
            ("uq32_to_float", [C.c_uint32], f),
            ("calculate_divergence", [F, F, I, I, f], None),
            ("subtract_gradient", [F, F, I, I, f], None),
            ("poisson_solve", [F, F, I, I, f, I, f], None),
            ("apply_drags", [F, C.c_void_p, I, I, I], None),
            ("step", [F, U, C.c_void_p, I, I, I, f, f, I, f, F, F], None),
        ]:
            fn = getattr(L, p + name)
            fn.argtypes, fn.restype = args, res
            self._fn[name] = fn

    # --- operators (same names as the reference's free functions) -------------
    def advect_vec2f(self, p, vel, dt, no_slip=True):
        dx_, dy_ = _dims(p)
        out = np.empty_like(p)
        self._fn["advect_vec2f"](_f32(out), _f32(p), _f32(vel), dx_, dy_, dt, int(no_slip))
        return out

    def advect_rgb_uq32(self, c, vel, dt, no_slip=False):
        dx_, dy_ = _dims(c)
        out = np.empty_like(c)
        self._fn["advect_rgb_uq32"](_u32(out), _u32(c), _f32(vel), dx_, dy_, dt, int(no_slip))
        return out

    def sample_vec2f(self, p, i, j, no_slip):
        dx_, dy_ = _dims(p)
        out = np.empty(2, np.float32)
        self._fn["sample_vec2f"](_f32(out), _f32(p), i, j, dx_, dy_, int(no_slip))
        return out

    def sample_rgb_uq32(self, c, i, j, no_slip):
        dx_, dy_ = _dims(c)
        out = np.empty(3, np.uint32)
        self._fn["sample_rgb_uq32"](_u32(out), _u32(c), i, j, dx_, dy_, int(no_slip))
        return out

    def uq32_from_float(self, x):
        return int(self._fn["uq32_from_float"](x))

    def uq32_to_float(self, raw):
        return float(self._fn["uq32_to_float"](raw))

    def calculate_divergence(self, v, dx=1.0):
        dx_, dy_ = _dims(v)
        out = np.empty(v.shape[:2], np.float32)
        self._fn["calculate_divergence"](_f32(out), _f32(v), dx_, dy_, dx)
        return out

    def subtract_gradient(self, v, p, dx=1.0):
        """In place on ``v`` (finitediff.cpp:80); also returns it."""
        dx_, dy_ = _dims(v)
        self._fn["subtract_gradient"](_f32(v), _f32(p), dx_, dy_, dx)
        return v

    def poisson_solve(self, div, dx=1.0, iters=10, omega=1.96, p=None):
        dx_, dy_ = _dims(div)
        if p is None:
            p = np.full(div.shape, 7.0, np.float32)  # initial contents must be ignored
        self._fn["poisson_solve"](_f32(p), _f32(div), dx_, dy_, dx, iters, omega)
        return p

    def apply_drags(self, v, drags):
        dx_, dy_ = _dims(v)
        drags = np.ascontiguousarray(drags, DRAG_DTYPE)
        self._fn["apply_drags"](_f32(v), drags.ctypes.data, len(drags), dx_, dy_)
        return v

    def step(self, v, c, drags=None, dt=1 / 30.0, dx=1.0, iters=10, omega=1.96,
             want_fields=False):
        """One loop() (ino:249-289), in place on v and c."""
        dx_, dy_ = _dims(v)
        if drags is None:
            drags = np.zeros(0, DRAG_DTYPE)
        drags = np.ascontiguousarray(drags, DRAG_DTYPE)
        p = d = None
        pp = dp = None
        if want_fields:
            p = np.empty(v.shape[:2], np.float32)
            d = np.empty(v.shape[:2], np.float32)
            pp, dp = _f32(p), _f32(d)
        self._fn["step"](_f32(v), _u32(c), drags.ctypes.data, len(drags), dx_, dy_,
                         dt, dx, iters, omega, pp, dp)
        return (v, c, p, d) if want_fields else (v, c)


class Oracle(_Base):
    """Plain-C restatement (oracle/fluid_oracle.c)."""
    prefix = "oracle_"
    path = ORACLE_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        F, U, I, f = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int, C.c_float
        L.oracle_sor_half_sweep.argtypes = [F, F, I, I, f, f, I]
        L.oracle_sor_half_sweep.restype = None
        L.oracle_upscale4_rgb565.argtypes = [C.POINTER(C.c_uint16), U, I, I]
        L.oracle_upscale4_rgb565.restype = None
        L.oracle_init_color_wheel.argtypes = [F, U, I, I]
        L.oracle_init_color_wheel.restype = None
        L.oracle_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_fnv1a64.restype = C.c_uint64

    # --- window ("tile") operators, used as the CPU backend of the decomposed path's tests ---
    def _tile_sigs(self):
        if getattr(self, "_tiles_bound", False):
            return
        L = self.lib
        F, U, I, f = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int, C.c_float
        T = C.POINTER(OracleTile)
        L.oracle_tile_advect_vec2f.argtypes = [F, F, F, T, f, I]
        L.oracle_tile_advect_vec2f.restype = I
        L.oracle_tile_advect_rgb_uq32.argtypes = [U, U, F, T, f, I]
        L.oracle_tile_advect_rgb_uq32.restype = I
        L.oracle_tile_calculate_divergence.argtypes = [F, F, T, f]
        L.oracle_tile_subtract_gradient.argtypes = [F, F, T, f]
        L.oracle_tile_sor_sweeps.argtypes = [F, F, F, T, f, f, I, I]
        L.oracle_tile_apply_drags.argtypes = [F, C.c_void_p, I, T]
        for n in ("calculate_divergence", "subtract_gradient", "sor_sweeps", "apply_drags"):
            getattr(L, "oracle_tile_" + n).restype = None
        self._tiles_bound = True

    @staticmethod
    def _t(tile):
        return OracleTile(*[getattr(tile, n) for n, _ in OracleTile._fields_])

    def tile_advect(self, next_p, p, vel, tile, dt, no_slip):
        """Returns True if a backtrace left the window (FS_ERR_HALO_OVERRUN in the product)."""
        self._tile_sigs()
        t = self._t(tile)
        if p.dtype == np.float32:
            return bool(self.lib.oracle_tile_advect_vec2f(_f32(next_p), _f32(p), _f32(vel), C.byref(t), dt, int(no_slip)))
        return bool(self.lib.oracle_tile_advect_rgb_uq32(_u32(next_p), _u32(p), _f32(vel), C.byref(t), dt, int(no_slip)))

    def tile_calculate_divergence(self, div, v, tile, dx):
        self._tile_sigs()
        self.lib.oracle_tile_calculate_divergence(_f32(div), _f32(v), C.byref(self._t(tile)), dx)

    def tile_subtract_gradient(self, v, p, tile, dx):
        self._tile_sigs()
        self.lib.oracle_tile_subtract_gradient(_f32(v), _f32(p), C.byref(self._t(tile)), dx)

    def tile_sor_sweeps(self, p_out, p_in, div, tile, dx, omega, first_parity, n_half):
        self._tile_sigs()
        self.lib.oracle_tile_sor_sweeps(_f32(p_out), _f32(p_in) if p_in is not None else None, _f32(div),
                                        C.byref(self._t(tile)), dx, omega, first_parity, n_half)

    def tile_apply_drags(self, v, drags, tile):
        self._tile_sigs()
        drags = np.ascontiguousarray(drags, DRAG_DTYPE)
        self.lib.oracle_tile_apply_drags(_f32(v), drags.ctypes.data, len(drags), C.byref(self._t(tile)))

    def sor_half_sweep(self, p, div, dx, omega, parity):
        dx_, dy_ = _dims(p)
        self.lib.oracle_sor_half_sweep(_f32(p), _f32(div), dx_, dy_, dx, omega, parity)
        return p

    def upscale4_rgb565(self, c):
        dx_, dy_ = _dims(c)
        out = np.empty(((dx_ - 1) * 4, (dy_ - 1) * 4), np.uint16)
        self.lib.oracle_upscale4_rgb565(out.ctypes.data_as(C.POINTER(C.c_uint16)),
                                        _u32(c), dx_, dy_)
        return out

    def init_color_wheel(self, dim_x, dim_y):
        v = np.empty((dim_y, dim_x, 2), np.float32)
        c = np.empty((dim_y, dim_x, 3), np.uint32)
        self.lib.oracle_init_color_wheel(_f32(v), _u32(c), dim_x, dim_y)
        return v, c

    def fnv1a64(self, a) -> int:
        a = np.ascontiguousarray(a)
        return int(self.lib.oracle_fnv1a64(a.ctypes.data, a.nbytes))


class Ref(_Base):
    """The reference's own sources behind ref_shim.cpp (oracle/_ref)."""
    prefix = "ref_"
    path = REF_SO

    def __init__(self):
        super().__init__()
        L = self.lib
        F, U, I = C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.c_int
        L.ref_saturates.restype = C.c_int
        L.ref_fill_rand.argtypes = [F, U, I, I, C.c_uint]
        L.ref_fill_rand.restype = None

        # the sketch itself (ESP32-fluid-simulation.ino, compiled unmodified by ino_shim.cpp)
        self.has_ino = hasattr(L, "ref_ino_loop")
        if self.has_ino:
            L.ref_ino_setup.argtypes = [F, U, I, I]
            L.ref_ino_loop.argtypes = [F, U, C.c_void_p, I, I, I]
            L.ref_ino_draw.argtypes = [C.POINTER(C.c_uint16), U, I, I]
            L.ref_ino_touch.argtypes = [C.c_void_p, I, C.POINTER(C.c_int), I, I, I]
            L.ref_ino_touch.restype = I
            for n in ("ref_ino_setup", "ref_ino_loop", "ref_ino_draw"):
                getattr(L, n).restype = None

    def saturates(self) -> bool:
        return bool(self.lib.ref_saturates())

    # --- the sketch's own routines -------------------------------------------------------------
    def ino_setup(self, dim_x, dim_y):
        """setup(), ino:194-246: (velocity, dye) initial condition."""
        v = np.empty((dim_y, dim_x, 2), np.float32)
        c = np.empty((dim_y, dim_x, 3), np.uint32)
        self.lib.ref_ino_setup(_f32(v), _u32(c), dim_x, dim_y)
        return v, c

    def ino_loop(self, v, c, drags=None):
        """loop(), ino:249-289, in place (K=10, omega=1.96, dt=1/30 are the sketch's literals)."""
        dx_, dy_ = _dims(v)
        drags = np.ascontiguousarray(drags if drags is not None else np.zeros(0, DRAG_DTYPE), DRAG_DTYPE)
        assert all(int(d["cy"]) < dx_ and int(d["cx"]) < dy_ for d in drags), "ino:266-268 writes out of bounds"
        self.lib.ref_ino_loop(_f32(v), _u32(c), drags.ctypes.data, len(drags), dx_, dy_)
        return v, c

    def ino_draw(self, c):
        """draw_routine(), ino:99-191: the RGB565 frame, (dim_x-1)*4 rows x (dim_y-1)*4 columns."""
        dx_, dy_ = _dims(c)
        out = np.empty(((dx_ - 1) * 4, (dy_ - 1) * 4), np.uint16)
        self.lib.ref_ino_draw(out.ctypes.data_as(C.POINTER(C.c_uint16)), _u32(c), dx_, dy_)
        return out

    def ino_touch(self, samples, dim_x=61, dim_y=81):
        """touch_routine(), ino:63-96: samples = [(touched, raw_x, raw_y)] per 10 ms poll -> drag records
        (at most 10: the queue depth)."""
        s = np.ascontiguousarray(samples, np.int32).reshape(-1, 3)
        out = np.zeros(16, DRAG_DTYPE)
        n = self.lib.ref_ino_touch(out.ctypes.data, len(out), s.ctypes.data_as(C.POINTER(C.c_int)), len(s),
                                   dim_x, dim_y)
        return out[:n].copy()

    def fill_rand(self, dim_x, dim_y, seed=1):
        v = np.empty((dim_y, dim_x, 2), np.float32)
        c = np.empty((dim_y, dim_x, 3), np.uint32)
        self.lib.ref_fill_rand(_f32(v), _u32(c), dim_x, dim_y, seed)
        return v, c


class Checker(Oracle):
    """What the parity tests compare against: the reference's OWN compiled code (oracle/_ref) for
    every operator it exports — advect, divergence, gradient, poisson_solve, drags, the loop() order
    — whenever that library is present and this host saturates float->uint32 like CUDA; the
    plain-C port (pinned against it in tests/test_oracle_vs_ref.py) for the window operators and
    wherever oracle/_ref is unavailable.  ``kind`` says which one answered."""

    def __init__(self):
        super().__init__()
        self.kind = "port"
        self.ref = None
        if have_ref():
            try:
                r = Ref()
            except OSError:
                r = None
            if r is not None and r.saturates():
                self.ref = r
                self._fn.update(r._fn)     # same signatures, ref_* instead of oracle_*
                self.kind = "reference"

    # ino:116-177 and ino:196-241 from the compiled sketch where available
    def upscale4_rgb565(self, c):
        if self.ref is not None and self.ref.has_ino:
            return self.ref.ino_draw(c)
        return super().upscale4_rgb565(c)

    def init_color_wheel(self, dim_x, dim_y):
        if self.ref is not None and self.ref.has_ino:
            return self.ref.ino_setup(dim_x, dim_y)
        return super().init_color_wheel(dim_x, dim_y)
