"""The C-ABI library loads on a GPU-less box and exports every symbol that
include/fluid_b200.h declares; the product never falls back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fluid_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fsh?_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    for want in ("fs_advect_vec2f", "fs_advect_rgb_uq32", "fs_calculate_divergence",
                 "fs_subtract_gradient", "fs_poisson_solve", "fs_apply_drags", "fs_step",
                 "fs_upscale4_rgb565", "fs_ensemble_step", "fsh_step", "fs_tile_sor_sweeps"):
        assert want in syms


def test_library_exports_every_declared_symbol(built):
    import esp32_fluid_simulation_b200 as fb
    lib = ctypes.CDLL(fb.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/fluid_b200.h but not exported: {missing}"


def test_binding_covers_every_declared_symbol(built):
    from esp32_fluid_simulation_b200 import _lib
    L = _lib.lib()
    unbound = [s for s in declared_symbols() if getattr(L, s).argtypes is None and s not in ("fs_version",)]
    assert not unbound, f"no ctypes signature for: {unbound}"


def test_no_cpu_fallback(built):
    """Without a device, creating a context fails with a CUDA error; nothing computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import esp32_fluid_simulation_b200 as fb
    with pytest.raises(fb.FluidError) as e:
        fb.Context(0)
    assert e.value.code > 0  # a cudaError_t, not a silent success


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "esp32-fluid-simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libfluid_oracle" not in src and "libfluid_ref" not in src, f


def test_argument_validation_without_context(built):
    from esp32_fluid_simulation_b200 import _lib
    L = _lib.lib()
    assert L.fs_poisson_solve(None, None, 8, 8, 1.0, 1, 1.0, None) == _lib.FS_ERR_NO_CONTEXT
    assert L.fs_ctx_destroy(None) == _lib.FS_ERR_NO_CONTEXT
    assert b"sm_100a" in L.fs_version()
    assert L.fs_error_string(_lib.FS_ERR_HALO_OVERRUN) is not None


def test_no_contracted_packed_fma_in_sass(built):
    """The reference's results need every multiply and add rounded separately (SURVEY.md hard part 1).
    A single FFMA / FFMA2 in the library would mean a rounding step of the reference is skipped
    (ptxas even contracts explicit mul.rn.f32x2 + add.rn.f32x2, so this is checked on the SASS)."""
    import subprocess
    import esp32_fluid_simulation_b200 as fb
    sass = subprocess.run(["cuobjdump", "-sass", fb.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert " FFMA2 " not in sass
    # per kernel; the only FFMAs allowed are library-routine internals, not contractions of reference
    # arithmetic: the Newton steps inside the IEEE-correct float DIVISION of touch_to_drags_kernel (ino:82-83:
    # delta * 1000.f / POLLING_PERIOD) and the double-precision atan2 of init_wheel_kernel (ino:209)
    allowed = ("touch_to_drags_kernel", "init_wheel_kernel")
    fused, fn = [], ""
    for l in sass.splitlines():
        if "Function :" in l:
            fn = l.split("Function :")[1].strip()
        elif (" FFMA " in l or " FFMA." in l) and not any(a in fn for a in allowed):
            fused.append((fn, l.strip()))
    assert not fused, fused[:5]


def test_header_is_valid_c99_and_layouts_match_the_reference(tmp_path):
    """The boundary is a C ABI: the header must compile as plain C, and the POD layouts must be the
    reference's (Vector2<float> 8 B, Vector3<UQ32> 12 B, struct drag 12 B)."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "fluid_b200.h"\n'
                   'int main(void){ fs_tile t; fs_halo_copy h; (void)t; (void)h;\n'
                   '  return sizeof(fs_vec2f)==8 && sizeof(fs_rgb_uq32)==12 && sizeof(fs_drag)==12 ? 0 : 1; }\n')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0
