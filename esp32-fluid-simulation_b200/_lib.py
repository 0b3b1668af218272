"""ctypes binding of include/fluid_b200.h.  Fails loudly when the CUDA library is
missing: there is no CPU path in this package."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH


class FluidError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: error {code} ({error_string(code)})")
        self.code = code


FS_OK = 0
FS_ERR_INVALID_ARG = -1
FS_ERR_NO_CONTEXT = -2
FS_ERR_UNSUPPORTED = -3
FS_ERR_HALO_OVERRUN = -4
FS_ERR_HALO_TIMEOUT = -5
FS_ERR_WOULD_BLOCK = -6


class HaloCopy(C.Structure):
    """fs_halo_copy: one pitched 2-D strip pushed into a neighbour's ghost region."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_pitch", C.c_longlong),
                ("dst_pitch", C.c_longlong), ("row_bytes", C.c_int), ("rows", C.c_int)]


class Tile(C.Structure):
    """fs_tile: a rank's padded local window of a global grid."""
    _fields_ = [(n, C.c_int) for n in
                ("gdim_x", "gdim_y", "ox", "oy", "nx", "ny", "x0", "y0", "x1", "y1")]


class DistConfig(C.Structure):
    """fs_dist_config"""
    _fields_ = [(n, C.c_int) for n in ("gdim_x", "gdim_y", "world", "rank", "px", "py", "ghost", "advect_halo",
                                      "iters")] + [(n, C.c_float) for n in ("dt", "dx", "omega")] + [("frame", C.c_int)]


class DistInfo(C.Structure):
    """fs_dist_info_t"""
    _fields_ = [(n, C.c_int) for n in ("px", "py", "n_neighbours", "sor_passes", "sor_t", "div_ring",
                                      "velocity_halo", "dye_halo", "exchanges_per_step")] + \
               [("exchanges", C.c_ulonglong), ("arena_bytes", C.c_size_t), ("phase_ms", C.c_float * 5)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing — build it with `python __graft_entry__.py` "
            "(or esp32-fluid-simulation_b200/build.py). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, I, f, u64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
    TP = C.POINTER(Tile)
    sigs = {
        "fs_ctx_create": ([C.POINTER(vp), I, vp], I),
        "fs_ctx_destroy": ([vp], I),
        "fs_ctx_synchronize": ([vp], I),
        "fs_ctx_set_option": ([vp, C.c_char_p, I], I),
        "fs_ctx_get_option": ([vp, C.c_char_p, C.POINTER(I)], I),
        "fs_ctx_launch_count": ([vp], u64),
        "fs_version": ([], C.c_char_p),
        "fs_error_string": ([I], C.c_char_p),
        "fs_host_alloc": ([C.POINTER(vp), C.c_size_t], I),
        "fs_host_free": ([vp], I),
        "fs_advect_vec2f": ([vp, vp, vp, I, I, f, I, vp], I),
        "fs_advect_rgb_uq32": ([vp, vp, vp, I, I, f, I, vp], I),
        "fs_calculate_divergence": ([vp, vp, I, I, f, vp], I),
        "fs_subtract_gradient": ([vp, vp, I, I, f, vp], I),
        "fs_poisson_solve": ([vp, vp, I, I, f, I, f, vp], I),
        "fs_sor_half_sweep": ([vp, vp, I, I, f, f, I, vp], I),
        "fs_apply_drags": ([vp, vp, I, I, I, vp], I),
        "fs_poisson_residual": ([C.POINTER(f), C.POINTER(C.c_double), vp, vp, I, I, f, vp], I),
        "fs_step": ([vp, vp, vp, I, I, I, f, f, I, f, vp, vp, vp], I),
        "fs_advect_drags_divergence": ([vp, vp, vp, vp, I, I, I, f, f, vp], I),
        "fs_advect_rgb_frame": ([vp, vp, vp, vp, I, I, f, I, vp], I),
        "fs_step_frame": ([vp, vp, vp, vp, vp, I, I, I, f, f, I, f, vp, vp, vp], I),
        "fs_step_pingpong": ([vp, vp, vp, vp, I, I, I, f, f, I, f, vp, vp, vp], I),
        "fs_upscale4_rgb565": ([vp, vp, I, I, vp], I),
        "fs_ensemble_step": ([vp, vp, vp, vp, I, I, I, I, f, f, I, f, I, vp], I),
        "fs_ensemble_step_dev": ([vp, vp, vp, vp, I, I, I, I, f, f, I, f, I, vp], I),
        "fs_init_color_wheel": ([vp, vp, I, I, I, vp], I),
        "fs_touch_to_drags": ([vp, vp, vp, I, I, I, I, I, vp, vp], I),
        "fsh_advect_vec2f": ([vp, vp, vp, I, I, f, I, vp], I),
        "fsh_advect_rgb_uq32": ([vp, vp, vp, I, I, f, I, vp], I),
        "fsh_calculate_divergence": ([vp, vp, I, I, f, vp], I),
        "fsh_subtract_gradient": ([vp, vp, I, I, f, vp], I),
        "fsh_poisson_solve": ([vp, vp, I, I, f, I, f, vp], I),
        "fsh_step": ([vp, vp, vp, I, I, I, f, f, I, f, vp, vp, vp], I),
        "fsh_upscale4_rgb565": ([vp, vp, I, I, vp], I),
        "fs_tile_advect_vec2f": ([vp, vp, vp, TP, f, I, vp], I),
        "fs_tile_advect_rgb_uq32": ([vp, vp, vp, TP, f, I, vp], I),
        "fs_tile_calculate_divergence": ([vp, vp, TP, f, vp], I),
        "fs_tile_subtract_gradient": ([vp, vp, TP, f, vp], I),
        "fs_tile_sor_sweeps": ([vp, vp, vp, TP, f, f, I, I, vp], I),
        "fs_tile_apply_drags": ([vp, vp, I, TP, vp], I),
        "fs_tile_check": ([vp], I),
        "fs_tile_max_displacement": ([C.POINTER(I), vp, TP, f, vp], I),
        "fs_arena_alloc": ([C.POINTER(vp), C.c_size_t, vp], I),
        "fs_arena_free": ([vp, vp], I),
        "fs_ipc_export": ([vp, C.c_char_p, vp], I),
        "fs_ipc_open": ([C.POINTER(vp), C.c_char_p, vp], I),
        "fs_ipc_close": ([vp, vp], I),
        "fs_halo_exchange": ([C.POINTER(HaloCopy), I, C.POINTER(vp), C.POINTER(vp), I, C.c_ulonglong, vp], I),
        "fs_ctx_set_stream": ([vp, vp], I),
        "fs_sim_create": ([C.POINTER(vp), I, I, f, f, I, f, I, vp], I),
        "fs_sim_destroy": ([vp], I),
        "fs_sim_upload": ([vp, vp, vp], I),
        "fs_sim_download": ([vp, vp, vp, vp, vp], I),
        "fs_sim_step": ([vp, vp, I], I),
        "fs_sim_acquire_frame": ([vp, C.POINTER(vp), C.POINTER(I), C.POINTER(I)], I),
        "fs_sim_release_frame": ([vp], I),
        "fs_sim_stats": ([vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)], I),
        "fs_dist_create": ([C.POINTER(vp), C.POINTER(DistConfig), vp], I),
        "fs_dist_destroy": ([vp], I),
        "fs_dist_window": ([vp, TP], I),
        "fs_dist_info": ([vp, C.POINTER(DistInfo)], I),
        "fs_dist_ipc_handle": ([vp, C.c_char_p], I),
        "fs_dist_connect": ([vp, C.c_char_p], I),
        "fs_dist_connect_local": ([vp, C.POINTER(vp)], I),
        "fs_dist_upload": ([vp, vp, vp], I),
        "fs_dist_download": ([vp, vp, vp, vp, vp], I),
        "fs_dist_device_fields": ([vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)], I),
        "fs_dist_frame": ([vp, C.POINTER(vp), C.POINTER(I), C.POINTER(I)], I),
        "fs_dist_step": ([vp, vp, I], I),
        "fs_dist_check": ([vp], I),
    }
    for name, (args, res) in sigs.items():
        fn = getattr(L, name)  # AttributeError if the library does not export the header's symbol
        fn.argtypes, fn.restype = args, res
    _lib = L
    return L


def error_string(code: int) -> str:
    try:
        return lib().fs_error_string(code).decode()
    except Exception:  # library missing: still give the number
        return "?"


def check(code: int, what: str) -> None:
    if code != FS_OK:
        raise FluidError(code, what)
