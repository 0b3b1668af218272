"""The CUDA window ("tile") kernels under a real decomposition: N ranks emulated as threads on
cuda:0 (same DecomposedSim + CudaTileOps as the multi-GPU path, in-process communicator),
bit-compared with the whole-grid oracle."""
import numpy as np
import pytest

from conftest import assert_bit_equal
from dist_util import gather_owned, run_threaded

pytestmark = pytest.mark.gpu
DT = np.float32(1 / 30.0)


@pytest.mark.parametrize("sor", [0, 1], ids=["half-sweeps", "blocked"])
@pytest.mark.parametrize("world,gx,gy,iters,sor_t,ghost", [(2, 256, 192, 10, 2, 16), (4, 512, 384, 13, 4, 32),
                                                           (8, 1024, 768, 20, 6, 64), (4, 200, 136, 9, 3, 24)])
def test_decomposed_cuda_run_matches_oracle(oracle, world, gx, gy, iters, sor_t, ghost, sor):
    from esp32_fluid_simulation_b200 import synth
    from esp32_fluid_simulation_b200.dist import CudaTileOps, DecomposedSim, Decomposition
    v0 = synth.velocity(gx, gy, vmax=150.0)
    c0 = synth.dye(gx, gy)
    steps = 3

    def make(rank, comm):
        dec = Decomposition(gx, gy, world, rank, ghost=ghost)
        ops = CudaTileOps(0)
        ops.ctx.set_option("sor", sor)
        sim = DecomposedSim(dec, ops, comm, iters, sor_t, DT)
        w = dec.window
        sim.load(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
        return sim

    drags = [synth.drags(gx, gy, s, n=8, vmax=400.0) for s in range(steps)]
    sims = run_threaded(world, make, steps, lambda s: drags[s])
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc, op, od = oracle.step(ov, oc, drags[s], DT, 1.0, iters, 1.96, want_fields=True)
    assert_bit_equal(gather_owned(sims, "v", gx, gy, 2, np.float32), ov, "velocity")
    assert_bit_equal(gather_owned(sims, "c", gx, gy, 3, np.uint32), oc, "dye")
    assert_bit_equal(gather_owned(sims, "p_last", gx, gy, 0, np.float32), op, "pressure")


def test_tile_overrun_is_reported(ctx):
    """A backtrace that leaves the window raises FS_ERR_HALO_OVERRUN instead of reading garbage."""
    import torch

    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200._lib import FS_ERR_HALO_OVERRUN
    t = fb.Tile(256, 64, 64, 0, 64, 64, 8, 0, 56, 64)        # window columns 64..127 of a 256-wide grid
    v = torch.full((64, 64, 2), 3000.0, device="cuda")        # 100 nodes per step
    out = torch.empty_like(v)
    ctx.tile_advect(out, v, v, t, DT, True)
    with pytest.raises(fb.FluidError) as e:
        ctx.tile_check()
    assert e.value.code == FS_ERR_HALO_OVERRUN
    ctx.tile_check()                                          # flag is cleared by the read
    assert ctx.tile_max_displacement(v, t, DT) == 102
