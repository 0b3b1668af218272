"""The CUDA window ("tile") kernels under a real decomposition: N ranks emulated as threads on
cuda:0 (same DecomposedSim + CudaTileOps as the multi-GPU path, in-process communicator),
bit-compared with the whole-grid oracle."""
import numpy as np
import pytest
import torch

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


# ---- fs_dist_*: the decomposed step in C++ with the SOR passes fused with their halo exchange -----------
# N ranks of one process on ONE device, each with its own context and stream; "peer" arenas are plain
# pointers (fs_dist_connect_local).  The flag protocol is the real one: rank A's persistent SOR kernel
# spins on flags that rank B's kernel sets, so every rank's persistent grid is capped at SMs/N
# ("sor_grid_limit") to keep all of them resident.

def _native_run(world, gx, gy, iters, sor_t, ghost, halo, steps, v0, c0, drags, grid=None, fuse=1, frame=False):
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200.dist import NativeDist
    props = torch.cuda.get_device_properties(0)
    ctxs, sims = [], []
    for r in range(world):
        ctx = fb.Context(0, torch.cuda.Stream())
        ctx.set_option("sor_t", sor_t)
        ctx.set_option("fuse", fuse)
        ctx.set_option("halo_timeout_ms", 20000)
        ctx.set_option("sor_grid_limit", max(1, props.multi_processor_count // world))
        ctxs.append(ctx)
        sims.append(NativeDist(ctx, gx, gy, world, r, iters, ghost=ghost, advect_halo=halo, grid=grid, frame=frame))
    for s in sims:
        s.connect_local(sims)
    for s in sims:
        w = s.window
        s.upload(np.ascontiguousarray(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx]),
                 np.ascontiguousarray(c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx]))
    torch.cuda.synchronize()
    for k in range(steps):
        for s in sims:                         # asynchronous: every rank's step is only enqueued here
            s.step(drags[k])
    gv, gc = np.zeros_like(v0), np.zeros_like(c0)
    gp, gd = np.zeros(v0.shape[:2], np.float32), np.zeros(v0.shape[:2], np.float32)
    errors = []
    for s in sims:
        try:
            s.check()
        except Exception as e:  # noqa: BLE001 — look at every rank, close everything, then report
            errors.append(e)
    if errors:
        for s in sims:
            s.close()
        raise errors[0]
    for s in sims:
        out = s.download()
        w = s.window
        ys, xs = slice(w.oy + w.y0, w.oy + w.y1), slice(w.ox + w.x0, w.ox + w.x1)
        gv[ys, xs], gc[ys, xs], gp[ys, xs], gd[ys, xs] = out["v"], out["c"], out["p"], out["d"]
    info = sims[0].info
    if frame:                                  # assemble the global RGB565 frame from the ranks' parts
        img = np.zeros(((gx - 1) * 4, (gy - 1) * 4), np.uint16)
        for s in sims:
            ptr, rows, cols = s.frame()
            from esp32_fluid_simulation_b200.dist import _device_view
            part = torch.as_tensor(_device_view(ptr, (rows, cols), "<i2"), device="cuda").cpu().numpy().view(np.uint16)
            w = s.window
            r0, c0_ = 4 * (w.ox + w.x0), 4 * (w.oy + w.y0)
            img[r0:r0 + rows, c0_:c0_ + cols] = part
        info["frame"] = img
    for s in sims:
        s.close()
    return gv, gc, gp, gd, info


@pytest.mark.parametrize("fuse", [1, 0], ids=["fused-advect-div", "one-kernel-per-operator"])
@pytest.mark.parametrize("world,grid,gx,gy,iters,sor_t,ghost,halo", [
    (2, None, 256, 192, 10, 2, 32, 20),          # 1x2, three passes
    (4, None, 512, 384, 13, 4, 32, 16),          # 2x2, corners
    (8, None, 1024, 768, 20, 6, 32, 12),         # 2x4
    (4, None, 520, 392, 9, 3, 24, 12),           # rectangles of unequal size
    (2, (2, 1), 512, 256, 50, 8, 64, 40),        # T=8, 7 passes, A=40, split along x
    (2, None, 512, 512, 50, 6, 64, 40),          # the bench's plan: K=50, T=6, remainder folded into the first pass (8 + 7 x 6)
    (4, (1, 4), 256, 1024, 5, 8, 32, 16),        # a single pass per step: no SOR hand-shake at all
    (1, None, 300, 200, 11, 4, 0, 0),            # one rank: same kernels, no neighbours
])
def test_native_decomposed_step_matches_oracle(oracle, world, grid, gx, gy, iters, sor_t, ghost, halo, fuse):
    from esp32_fluid_simulation_b200 import synth
    v0 = synth.velocity(gx, gy, vmax=150.0)
    c0 = synth.dye(gx, gy)
    steps = 3
    drags = [synth.drags(gx, gy, s, n=8, vmax=float(30 * (halo - 6)) if halo else 400.0) for s in range(steps)]
    gv, gc, gp, gd, info = _native_run(world, gx, gy, iters, sor_t, ghost, halo, steps, v0, c0, drags, grid, fuse)
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc, op, od = oracle.step(ov, oc, drags[s], DT, 1.0, iters, 1.96, want_fields=True)
    assert_bit_equal(gv, ov, "velocity")
    assert_bit_equal(gc, oc, "dye")
    assert_bit_equal(gp, op, "pressure")
    assert_bit_equal(gd, od, "divergence")
    if world > 1:
        passes = -(-iters // sor_t)
        rem = iters - (passes - 1) * sor_t
        if passes > 1 and rem <= 2 and sor_t + rem <= 8 and info["velocity_halo"] <= ghost and info["sor_passes"] == passes - 1:
            passes -= 1                                          # the remainder was folded into the first pass
        # hand-shakes: one per SOR pass but the last, + the velocity and the dye exchange on their side streams
        assert info["exchanges_per_step"] == passes + 1 and info["sor_passes"] == passes
        assert info["exchanges"] == steps * (passes + 1) + 1    # + the fresh state's velocity + dye halo


def test_native_decomposed_overrun_is_reported():
    """A backtrace beyond the static advect halo raises FS_ERR_HALO_OVERRUN instead of reading stale ghosts."""
    import esp32_fluid_simulation_b200 as fb
    from esp32_fluid_simulation_b200 import synth
    from esp32_fluid_simulation_b200._lib import FS_ERR_HALO_OVERRUN
    gx, gy = 256, 192
    v0, c0 = synth.velocity(gx, gy, vmax=60.0), synth.dye(gx, gy)
    v0[100, 10] = (0.0, 1500.0)                  # 50 nodes per step across the cut at y = 96; halo = 12
    with pytest.raises(fb.FluidError) as e:
        _native_run(2, gx, gy, 6, 3, 32, 12, 1, v0, c0, [None])
    assert e.value.code == FS_ERR_HALO_OVERRUN


@pytest.mark.parametrize("world,grid,gx,gy,iters,sor_t,ghost,halo", [
    (2, None, 256, 192, 10, 2, 32, 16),
    (4, None, 512, 384, 13, 4, 32, 12),
    (8, None, 1024, 768, 20, 6, 32, 8),
    (1, None, 300, 200, 11, 4, 0, 0),
])
def test_native_decomposed_frame(oracle, world, grid, gx, gy, iters, sor_t, ghost, halo):
    """fs_dist with frame=1: every rank renders the cells that start in its rectangle inside its dye advect (the far
    corners of its last cells are the neighbour's nodes, recomputed locally) — assembled == the reference's frame."""
    from esp32_fluid_simulation_b200 import synth
    v0, c0 = synth.velocity(gx, gy, vmax=150.0), synth.dye(gx, gy)
    steps = 2
    drags = [synth.drags(gx, gy, s, n=8, vmax=float(30 * (halo - 6)) if halo else 400.0) for s in range(steps)]
    gv, gc, gp, gd, info = _native_run(world, gx, gy, iters, sor_t, ghost, halo, steps, v0, c0, drags, grid, frame=True)
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc = oracle.step(ov, oc, drags[s], DT, 1.0, iters, 1.96)
    assert_bit_equal(gv, ov, "velocity")
    assert_bit_equal(gc, oc, "dye")
    assert_bit_equal(info["frame"], oracle.upscale4_rgb565(oc), "RGB565 frame")
