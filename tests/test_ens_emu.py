"""The register-tiled ensemble kernel's own source (csrc/ensemble_reg.cuh) compiled for the HOST and run on
CPU threads (tests/emu/), bit-compared with the oracle: thread mapping, mailbox protocol, masks, walls and
arithmetic are checked in the CPU tier; the GPU tier (test_gpu_parity.py::test_ensemble_step) then checks
the same source as compiled by nvcc."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libens_emu.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed: cannot compile the kernel source for the host")
    src = os.path.join(EMU_DIR, "ens_emu.cpp")
    deps = [src, os.path.join(EMU_DIR, "cuda_host_shim.h")] + [
        os.path.join(ROOT, "esp32-fluid-simulation_b200", "csrc", f)
        for f in ("ensemble_reg.cuh", "ensemble_common.cuh", "advect.cuh", "sor.cuh", "fs_common.cuh")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(d) > os.path.getmtime(EMU_SO) for d in deps):
        os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
        subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                        "-I", CUDA_INC, "-o", EMU_SO, src], check=True, cwd=EMU_DIR)
    lib = ctypes.CDLL(EMU_SO)
    vp, I, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.ens_emu_step.argtypes = [vp, vp, vp, vp, I, I, I, I, f, f, I, f, I, I, I, I, I, I]
    lib.ens_emu_step.restype = I
    return lib


def run_case(emu, port, shape, iters, batch, n_steps, R, dye_smem, nblocks=2, threads=0, vmax=90.0, pipe=0):
    from esp32_fluid_simulation_b200 import synth
    dim_x, dim_y = shape
    max_drags = 4
    v = np.stack([synth.velocity(dim_x, dim_y, seed=100 + b, vmax=vmax) for b in range(batch)])
    c = np.stack([synth.dye(dim_x, dim_y, seed=200 + b, n_splats=6) for b in range(batch)])
    drags = np.zeros((n_steps, batch, max_drags), synth.DRAG_DTYPE)
    counts = np.zeros((n_steps, batch), np.int32)
    for s in range(n_steps):
        for b in range(batch):
            k = (b + s) % (max_drags + 1)
            counts[s, b] = k
            drags[s, b, :k] = synth.drags(dim_x, dim_y, s * 1000 + b, n=max_drags, vmax=300.0)[:k]
    gv, gc = v.copy(), c.copy()
    rc = emu.ens_emu_step(gv.ctypes.data, gc.ctypes.data, drags.ctypes.data, counts.ctypes.data, max_drags, batch,
                          dim_x, dim_y, synth.DT, 1.0, iters, 1.96, n_steps, R, int(dye_smem), nblocks, threads, pipe)
    assert rc == 0
    for b in range(batch):
        ov, oc = v[b].copy(), c[b].copy()
        for s in range(n_steps):
            ov, oc = port.step(ov, oc, drags[s, b, :counts[s, b]], synth.DT, 1.0, iters, 1.96)
        assert np.array_equal(gv[b].view(np.uint32), ov.view(np.uint32)), f"grid {b} velocity"
        assert np.array_equal(gc[b], oc), f"grid {b} dye"


@pytest.mark.parametrize("shape,iters,batch,n_steps,R,dye_smem", [
    ((16, 12), 3, 3, 2, 2, True),       # blocks tile the grid exactly
    ((16, 12), 3, 2, 2, 4, False),
    ((16, 12), 4, 2, 1, 6, True),
    ((16, 16), 2, 1, 2, 8, True),
    ((13, 11), 3, 2, 2, 2, True),       # ragged: last column group has 1 valid column, last strip 1 valid row
    ((13, 11), 3, 2, 2, 4, True),
    ((14, 9), 5, 2, 2, 6, False),
    ((15, 10), 2, 2, 1, 8, True),
    ((7, 3), 4, 3, 2, 4, True),         # one strip
    ((2, 2), 3, 2, 1, 2, True),         # every node a corner
    ((3, 2), 3, 2, 2, 6, True),
    ((21, 17), 0, 2, 2, 4, True),       # no SOR iterations
    ((21, 17), 1, 2, 2, 4, False),      # only the zero-start half-sweep pair
])
def test_ens_reg_source_on_cpu_threads(emu, port, shape, iters, batch, n_steps, R, dye_smem):
    run_case(emu, port, shape, iters, batch, n_steps, R, dye_smem)


def test_ens_reg_source_reference_shape(emu, port):
    """The reference's own grid (61x81, K=10) and the benchmark shape (80x60), one step each, more threads
    than blocks in the advects (threads=...)."""
    run_case(emu, port, (61, 81), 10, 1, 1, 6, True, nblocks=1)
    run_case(emu, port, (80, 60), 10, 1, 1, 4, True, nblocks=1, threads=320)


def test_ens_reg_fast_velocities(emu, port):
    """CFL >> 1: most backtraces leave the grid, so the general sample() redo path carries the step."""
    run_case(emu, port, (13, 11), 3, 2, 2, 4, True, vmax=900.0)


@pytest.mark.parametrize("shape,iters,batch,n_steps,nblocks", [
    ((16, 12), 3, 5, 1, 2),      # one step per call: every step is a grid's last (dye in place, both prefetches in it)
    ((16, 12), 3, 5, 3, 2),      # ping-pong steps, then the in-place one
    ((16, 12), 2, 4, 2, 4),      # one grid per CTA: nothing to prefetch
    ((13, 11), 3, 3, 2, 1),      # ragged, one CTA walks all grids
    ((80, 60), 2, 2, 1, 1),      # the benchmark shape with the launcher's CTA size
])
def test_ens_reg_pipelined_flow_on_cpu_threads(emu, port, shape, iters, batch, n_steps, nblocks):
    """The dye-resident R = 2 kernel's PIPELINED flow (in-place dye advect in a grid's last step, next grid's
    state copied in under it) with the bulk copies replaced by cooperative ones at the same program points."""
    run_case(emu, port, shape, iters, batch, n_steps, 2, True, nblocks=nblocks, pipe=1)


def test_ens_reg_source_under_address_sanitizer():
    """The kernel source once more, compiled with -fsanitize=address: the emulator's 'shared memory' is a heap block of
    exactly the size the launcher allocates and the state arrays are exact-size numpy arrays, so a shared-memory or
    global access outside them — an off-by-one in the congruent-copy plans, say — is reported.  Runs in a subprocess
    (the sanitizer runtime has to be preloaded)."""
    import sys
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not installed")
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not libasan or not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan not installed")
    so = os.path.join(EMU_DIR, "_build", "libens_emu_asan.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
                    "-fsanitize=address", "-I", CUDA_INC, "-o", so, os.path.join(EMU_DIR, "ens_emu.cpp")],
                   check=True, cwd=EMU_DIR)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "run_cases.py"), so], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0 and "EMU_CASES_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "AddressSanitizer" not in r.stderr


def test_ens_reg_source_random_configurations(emu):
    """A short run of the randomised soak (tests/emu/fuzz.py; the long runs are done by hand)."""
    import sys
    r = subprocess.run([sys.executable, os.path.join(EMU_DIR, "fuzz.py"), EMU_SO, "7", "12"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "FUZZ_OK" in r.stdout, (r.stdout[-1000:], r.stderr[-3000:])
