"""Host-side logic of the decomposed path on CPU: geometry, halo exchange sequencing (threads
and a real 2-process gloo group), bit-identity of a decomposed run with the whole-grid oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_bit_equal
from dist_util import OracleTileOps, gather_owned, run_threaded

DT = np.float32(1 / 30.0)


def test_process_grid_and_split():
    from esp32_fluid_simulation_b200.dist import Decomposition, process_grid, split
    assert [process_grid(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    for n, parts in [(4096, 2), (16384, 4), (100, 2), (75, 4), (24576, 2)]:
        cuts = [split(n, parts, k) for k in range(parts)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(lo % 4 == 0 for lo, _ in cuts)
    d = Decomposition(16384, 16384, 8, 5, ghost=64)            # BASELINE.json configs[3]
    assert (d.px, d.py, d.rx, d.ry) == (2, 4, 1, 2)
    w = d.window
    assert (w.x1 - w.x0, w.y1 - w.y0) == (8192, 4096) and w.nx % 4 == 0
    assert w.ox == 8192 - 64 and w.oy == 8192 - 64 and w.ny == 4096 + 128 and w.nx == 8192 + 64
    d = Decomposition(24576, 32768, 8, 0, ghost=64)            # configs[4]: 12288 x 8192 tiles
    assert (d.window.x1 - d.window.x0, d.window.y1 - d.window.y0) == (12288, 8192)


def test_neighbour_lists_mirror():
    from esp32_fluid_simulation_b200.dist import Decomposition
    world = 8
    decs = [Decomposition(64, 128, world, r, ghost=8) for r in range(world)]
    for d in decs:
        for peer, dx, dy in d.neighbours():
            assert (d.rank, -dx, -dy) in decs[peer].neighbours()
            ys, xs = d.send_slices(dx, dy, 5)
            yr, xr = decs[peer].recv_slices(-dx, -dy, 5)
            assert (ys.stop - ys.start, xs.stop - xs.start) == (yr.stop - yr.start, xr.stop - xr.start)
            # the strip I send is, in global coordinates, exactly the ghost strip the peer fills
            assert d.window.ox + xs.start == decs[peer].window.ox + xr.start
            assert d.window.oy + ys.start == decs[peer].window.oy + yr.start


def _inputs(gx, gy, seed, vmax):
    rng = np.random.default_rng(seed)
    v = ((rng.random((gy, gx, 2), np.float32) - np.float32(0.5)) * np.float32(2 * vmax)).astype(np.float32)
    c = rng.integers(0, 2 ** 32, (gy, gx, 3), dtype=np.uint32)
    return v, c


def _drags(gx, gy, step, n=6):
    from esp32_fluid_simulation_b200 import synth
    return synth.drags(gx, gy, step, n=n, vmax=300.0)


@pytest.mark.parametrize("world,gx,gy,iters,sor_t", [(2, 48, 64, 10, 2), (4, 64, 56, 7, 3), (8, 64, 128, 10, 4),
                                                     (4, 100, 75, 5, 1), (1, 40, 30, 10, 4)])
def test_decomposed_run_is_bit_identical_to_whole_grid(oracle, world, gx, gy, iters, sor_t):
    from esp32_fluid_simulation_b200.dist import DecomposedSim, Decomposition
    v0, c0 = _inputs(gx, gy, world, 120.0)
    steps = 4

    def make(rank, comm):
        dec = Decomposition(gx, gy, world, rank, ghost=16)
        sim = DecomposedSim(dec, OracleTileOps(oracle), comm, iters, sor_t, DT)
        w = dec.window
        sim.load(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
        return sim

    sims = run_threaded(world, make, steps, lambda s: _drags(gx, gy, s))
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc, op, od = oracle.step(ov, oc, _drags(gx, gy, s), DT, 1.0, iters, 1.96, want_fields=True)
    assert_bit_equal(gather_owned(sims, "v", gx, gy, 2, np.float32), ov, "velocity")
    assert_bit_equal(gather_owned(sims, "c", gx, gy, 3, np.uint32), oc, "dye")
    assert_bit_equal(gather_owned(sims, "p_last", gx, gy, 0, np.float32), op, "pressure")
    if world > 1:
        assert sims[0].exchanges > 0


def test_halo_wider_than_ghost_is_an_error(oracle):
    from esp32_fluid_simulation_b200.dist import DecomposedSim, Decomposition
    gx, gy, world = 64, 64, 2
    v0, c0 = _inputs(gx, gy, 9, 2000.0)            # ~67 nodes per step: more than ghost=16

    def make(rank, comm):
        dec = Decomposition(gx, gy, world, rank, ghost=16)
        sim = DecomposedSim(dec, OracleTileOps(oracle), comm, 4, 2, DT)
        w = dec.window
        sim.load(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
        return sim

    with pytest.raises(RuntimeError, match="larger ghost"):
        run_threaded(world, make, 1, lambda s: None)


@pytest.mark.parametrize("world", [2, 4])
def test_static_full_ghost_halo_needs_no_agreement_and_detects_overruns(oracle, world):
    """static_halo == ghost: no per-step all-reduce, still bit-identical; and a backtrace that leaves the
    window is reported by check() instead of passing silently."""
    from esp32_fluid_simulation_b200.dist import DecomposedSim, Decomposition
    gx, gy, iters, sor_t, ghost, steps = 64, 64, 6, 2, 16, 3
    v0, c0 = _inputs(gx, gy, 5, 60.0)                  # <= 2 nodes per step, drags below add <= 10

    def run(vfield, expect_overrun):
        calls = []

        def make(rank, comm):
            dec = Decomposition(gx, gy, world, rank, ghost=ghost)
            comm_all_max = comm.all_max
            comm.all_max = lambda v: calls.append(v) or comm_all_max(v)
            sim = DecomposedSim(dec, OracleTileOps(oracle), comm, iters, sor_t, DT, static_halo=ghost)
            w = dec.window
            sim.load(vfield[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
            return sim

        sims = run_threaded(world, make, steps, lambda s: _drags(gx, gy, s))
        assert not calls, "static halo must not all-reduce"
        hits = 0
        for sim in sims:
            try:
                sim.check()
            except RuntimeError:
                hits += 1
        assert (hits > 0) == expect_overrun
        return sims

    sims = run(v0, expect_overrun=False)
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc = oracle.step(ov, oc, _drags(gx, gy, s), DT, 1.0, iters, 1.96)
    assert_bit_equal(gather_owned(sims, "v", gx, gy, 2, np.float32), ov, "velocity")
    assert_bit_equal(gather_owned(sims, "c", gx, gy, 3, np.uint32), oc, "dye")
    fast = v0.copy()
    fast[gy // 2, :, 1] = 1500.0                       # 50 nodes per step across the cut between ranks
    run(fast, expect_overrun=True)


def _gloo_worker(rank, world, port, gx, gy, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from esp32_fluid_simulation_b200.dist import DecomposedSim, Decomposition, TorchComm
    from oracle import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        v0, c0 = _inputs(gx, gy, 77, 100.0)
        dec = Decomposition(gx, gy, world, rank, ghost=12)
        sim = DecomposedSim(dec, OracleTileOps(Oracle()), TorchComm("cpu"), 6, 2, DT)
        w = dec.window
        sim.load(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
        for s in range(3):
            sim.step(_drags(gx, gy, s))
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), v=sim.owned(sim.v), c=sim.owned(sim.c),
                 box=np.array([dec.gx0, dec.gx1, dec.gy0, dec.gy1]))
    finally:
        dist.destroy_process_group()


def test_two_process_gloo_group(oracle, tmp_path):
    """world_size 2 over real torch.distributed (gloo) — the same TorchComm the NCCL path uses."""
    import socket

    import torch.multiprocessing as mp
    gx, gy, world = 40, 48, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_gloo_worker, args=(world, port, gx, gy, str(tmp_path)), nprocs=world, join=True)
    v0, c0 = _inputs(gx, gy, 77, 100.0)
    ov, oc = v0.copy(), c0.copy()
    for s in range(3):
        ov, oc = oracle.step(ov, oc, _drags(gx, gy, s), DT, 1.0, 6, 1.96)
    gv, gc = np.zeros_like(ov), np.zeros_like(oc)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        x0, x1, y0, y1 = z["box"]
        gv[y0:y1, x0:x1], gc[y0:y1, x0:x1] = z["v"], z["c"]
    assert_bit_equal(gv, ov, "velocity (gloo)")
    assert_bit_equal(gc, oc, "dye (gloo)")


@pytest.mark.parametrize("schedule", ["even", "rank0-slow", "last-slow", "alternate-slow", "jitter"])
@pytest.mark.parametrize("world,iters,sor_t", [(2, 6, 2), (4, 7, 3), (8, 4, 4)])
def test_one_sided_halo_protocol_is_schedule_independent(oracle, world, iters, sor_t, schedule):
    """The peer-memory exchange (csrc/halo.cu) stores into the neighbours' windows at the SENDER's
    program point.  Emulated here with threads; ranks are slowed down to let neighbours run ahead as far
    as the flags allow.  Whatever the schedule, the result must be the whole-grid oracle's, bit for bit —
    i.e. no step of the sequence reads or rewrites ghosts that a neighbour may already be overwriting."""
    from dist_util import OneSidedComm, OneSidedOps, OneSidedWorld, run_one_sided
    from esp32_fluid_simulation_b200.dist import DecomposedSim, Decomposition
    gx, gy, ghost, steps = (48, 64, 16, 3) if world == 2 else (64, 64, 16, 3) if world == 4 else (64, 128, 16, 2)
    v0, c0 = _inputs(gx, gy, 40 + world, 60.0)
    slow = {"even": {}, "rank0-slow": {0: 0.004}, "last-slow": {world - 1: 0.004},
            "alternate-slow": {r: 0.003 for r in range(0, world, 2)}, "jitter": {}}[schedule]
    osw = OneSidedWorld(world, slow=slow, jitter=0.003 if schedule == "jitter" else 0.0, seed=world)
    decs = [Decomposition(gx, gy, world, r, ghost=ghost) for r in range(world)]

    def make(rank):
        sim = DecomposedSim(decs[rank], OneSidedOps(oracle, osw, rank), OneSidedComm(osw, rank, decs), iters, sor_t,
                            DT, static_halo=ghost)
        w = decs[rank].window
        sim.load(v0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx], c0[w.oy:w.oy + w.ny, w.ox:w.ox + w.nx])
        return sim

    sims = run_one_sided(world, make, steps, lambda s: _drags(gx, gy, s, n=4), osw)
    ov, oc = v0.copy(), c0.copy()
    for s in range(steps):
        ov, oc, op, od = oracle.step(ov, oc, _drags(gx, gy, s, n=4), DT, 1.0, iters, 1.96, want_fields=True)
    assert_bit_equal(gather_owned(sims, "v", gx, gy, 2, np.float32), ov, "velocity")
    assert_bit_equal(gather_owned(sims, "c", gx, gy, 3, np.uint32), oc, "dye")
    assert_bit_equal(gather_owned(sims, "p_last", gx, gy, 0, np.float32), op, "pressure")
