"""The host C++ harness (harness/fluid_harness.cpp): the reference's loop() order through a
table of operator pointers, each bindable to the CPU checker library or to the CUDA library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "harness", "fluid_harness")
ORACLE_SO = os.path.join(ROOT, "oracle", "libfluid_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfluid_ref.so")


def run(*args):
    r = subprocess.run([HARNESS, *args], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout + r.stderr


def hashes(out):
    return [l.split("cpu")[1].split() for l in out.splitlines() if "fnv1a64" in l and " cpu " in l][0]


def test_harness_cpu_only_reference_and_oracle_agree(built):
    rc, out_o = run("--cpu-lib", ORACLE_SO, "--cpu-prefix", "oracle_", "--steps", "6", "--gpu-ops", "none")
    assert rc == 0 and "MATCH" in out_o, out_o
    if os.path.exists(REF_SO):
        rc, out_r = run("--cpu-lib", REF_SO, "--cpu-prefix", "ref_", "--steps", "6", "--gpu-ops", "none")
        assert rc == 0, out_r
        assert hashes(out_r) == hashes(out_o)


def test_operator_mirror_header_compiles_against_reference_shaped_types(built, tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('''
#include <cstdint>
#include "fluid_ops.hpp"
template <class T> struct Vector2 { T x, y; };
template <class T> struct Vector3 { T x, y, z; };
struct UQ32 { uint32_t raw; };
void loop_body(Vector2<float>* v, Vector2<float>* v_tmp, Vector3<UQ32>* c, Vector3<UQ32>* c_tmp, float* div, float* p) {
    using namespace fluid_b200;
    fluid_b200::advect(v_tmp, v, v, 61, 81, 1 / 30.0f, true);
    fluid_b200::calculate_divergence(div, v_tmp, 61, 81, 1);
    fluid_b200::poisson_solve(p, div, 61, 81, 1, 10, 1.96);
    fluid_b200::subtract_gradient(v_tmp, p, 61, 81, 1);
    fluid_b200::advect(c_tmp, c, v_tmp, 61, 81, 1 / 30.0f, false);
}
''')
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)], check=True)


@pytest.mark.gpu
@pytest.mark.parametrize("ops", ["all", "poisson", "advect_v,advect_c", "divergence,gradient"])
def test_harness_swaps_operators_per_call(built, ops):
    import esp32_fluid_simulation_b200 as fb
    cpu = (REF_SO, "ref_") if os.path.exists(REF_SO) else (ORACLE_SO, "oracle_")
    for dims in (("61", "81"), ("256", "192")):
        rc, out = run("--cpu-lib", cpu[0], "--cpu-prefix", cpu[1], "--gpu-lib", fb.LIB_PATH, "--gpu-ops", ops,
                      "--dim-x", dims[0], "--dim-y", dims[1], "--steps", "8")
        assert rc == 0 and "MATCH: bit-identical" in out, out


@pytest.mark.gpu
@pytest.mark.parametrize("world,dims", [(2, ("512", "256")), (4, ("512", "384")), (8, ("1024", "768"))])
def test_harness_drives_the_decomposed_step_from_cpp(built, world, dims):
    """fs_dist_* lets the host C++ harness run the multi-GPU step (here: `world` ranks emulated on
    device 0, each with a context-owned stream) — bit-identical to the CPU library's whole-grid run."""
    import esp32_fluid_simulation_b200 as fb
    cpu = (REF_SO, "ref_") if os.path.exists(REF_SO) else (ORACLE_SO, "oracle_")
    rc, out = run("--cpu-lib", cpu[0], "--cpu-prefix", cpu[1], "--gpu-lib", fb.LIB_PATH, "--gpu-ops", "poisson",
                  "--dim-x", dims[0], "--dim-y", dims[1], "--steps", "3", "--iters", "20", "--drags", "6",
                  "--decomposed", str(world))
    assert rc == 0 and "MATCH: bit-identical" in out and f"({world} ranks)" in out, out
