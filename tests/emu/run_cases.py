"""Runs a few ensemble-kernel emulation cases against the oracle with the emulator library given on the command
line; used by tests/test_ens_emu.py to run the kernel source under AddressSanitizer in a subprocess
(LD_PRELOAD=libasan.so): the 'shared memory' is a heap block and the state arrays are exact-size numpy arrays, so
any access outside them is reported.  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from esp32_fluid_simulation_b200 import synth  # noqa: E402
from oracle import Oracle  # noqa: E402


def main():
    lib = ctypes.CDLL(sys.argv[1])
    vp, I, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.ens_emu_step.argtypes = [vp, vp, vp, vp, I, I, I, I, f, f, I, f, I, I, I, I, I, I]
    lib.ens_emu_step.restype = I
    port = Oracle()
    # (dim_x, dim_y, iters, batch, n_steps, R, dye_smem, nblocks, pipe)
    cases = [(13, 11, 3, 3, 2, 2, 1, 1, 1), (13, 11, 3, 3, 1, 2, 1, 1, 1), (13, 11, 2, 3, 2, 2, 1, 1, 0),
             (14, 9, 2, 3, 2, 4, 1, 2, 0), (16, 12, 2, 3, 1, 2, 1, 1, 1), (15, 10, 2, 2, 2, 6, 0, 1, 0),
             (2, 2, 2, 3, 2, 2, 1, 1, 1), (7, 3, 2, 3, 1, 2, 1, 1, 1)]
    for dim_x, dim_y, iters, batch, n_steps, R, dye_smem, nblocks, pipe in cases:
        max_drags = 3
        v = np.stack([synth.velocity(dim_x, dim_y, seed=100 + b, vmax=90.0) for b in range(batch)])
        c = np.stack([synth.dye(dim_x, dim_y, seed=200 + b, n_splats=4) for b in range(batch)])
        drags = np.zeros((n_steps, batch, max_drags), synth.DRAG_DTYPE)
        counts = np.zeros((n_steps, batch), np.int32)
        for s in range(n_steps):
            for b in range(batch):
                k = (b + s) % (max_drags + 1)
                counts[s, b] = k
                drags[s, b, :k] = synth.drags(dim_x, dim_y, s * 1000 + b, n=max_drags, vmax=300.0)[:k]
        gv, gc = v.copy(), c.copy()
        rc = lib.ens_emu_step(gv.ctypes.data, gc.ctypes.data, drags.ctypes.data, counts.ctypes.data, max_drags, batch,
                              dim_x, dim_y, synth.DT, 1.0, iters, 1.96, n_steps, R, dye_smem, nblocks, 0, pipe)
        assert rc == 0, rc
        for b in range(batch):
            ov, oc = v[b].copy(), c[b].copy()
            for s in range(n_steps):
                ov, oc = port.step(ov, oc, drags[s, b, :counts[s, b]], synth.DT, 1.0, iters, 1.96)
            assert np.array_equal(gv[b].view(np.uint32), ov.view(np.uint32)), (dim_x, dim_y, b, "velocity")
            assert np.array_equal(gc[b], oc), (dim_x, dim_y, b, "dye")
    print("EMU_CASES_OK", len(cases))


if __name__ == "__main__":
    main()
