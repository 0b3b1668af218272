"""Real multi-GPU runs of the decomposed path (needs >= 2 visible GPUs; skipped otherwise): one
process per GPU, NCCL only for bootstrap, halos exchanged by peer-memory stores over NVLink
(fs_halo_exchange) — and the NCCL send/recv path for comparison.  Results are bit-compared with
the whole-grid oracle."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import assert_bit_equal

pytestmark = pytest.mark.gpu
DT = np.float32(1 / 30.0)


def _worker(rank, world, port, gx, gy, iters, sor_t, ghost, steps, mode, static_halo, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import torch
    import torch.distributed as dist
    from esp32_fluid_simulation_b200 import synth
    from esp32_fluid_simulation_b200.dist import (ArenaTileOps, CudaTileOps, DecomposedSim, Decomposition, PeerComm,
                                                  TorchComm, max_window_nodes)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        dec = Decomposition(gx, gy, world, rank, ghost=ghost)
        if mode == "peer":
            ops = ArenaTileOps(rank, max_window_nodes(gx, gy, world, ghost))
            comm = PeerComm(ops, world, rank, gx, gy, ghost)
        else:
            ops = CudaTileOps(rank)
            comm = TorchComm(torch.device("cuda", rank))
        sim = DecomposedSim(dec, ops, comm, iters, sor_t, DT, static_halo=static_halo)
        w = dec.window
        sim.load(synth.velocity(gx, gy, vmax=150.0, window=(w.ox, w.oy, w.nx, w.ny)),
                 synth.dye(gx, gy, window=(w.ox, w.oy, w.nx, w.ny)))
        for s in range(steps):
            sim.step(synth.drags(gx, gy, s, n=8, vmax=400.0))
        torch.cuda.synchronize()
        if static_halo is not None:
            sim.check()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), v=sim.owned(sim.v), c=sim.owned(sim.c),
                 p=sim.owned(sim.p_last), box=np.array([dec.gx0, dec.gx1, dec.gy0, dec.gy1]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("mode,static_halo", [("peer", None), ("peer", 24), ("nccl", None)])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_decomposed_run_matches_oracle(oracle, tmp_path, world, mode, static_halo):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from esp32_fluid_simulation_b200 import synth
    gx, gy, iters, sor_t, ghost, steps = 1024, 768, 20, 6, 32, 3
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, gx, gy, iters, sor_t, ghost, steps, mode, static_halo, str(tmp_path)),
             nprocs=world, join=True)
    ov, oc = synth.velocity(gx, gy, vmax=150.0), synth.dye(gx, gy)
    for s in range(steps):
        ov, oc, op, od = oracle.step(ov, oc, synth.drags(gx, gy, s, n=8, vmax=400.0), DT, 1.0, iters, 1.96,
                                     want_fields=True)
    gv, gc, gp = np.zeros_like(ov), np.zeros_like(oc), np.zeros_like(op)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        x0, x1, y0, y1 = z["box"]
        gv[y0:y1, x0:x1], gc[y0:y1, x0:x1], gp[y0:y1, x0:x1] = z["v"], z["c"], z["p"]
    assert_bit_equal(gv, ov, f"velocity ({mode}, {world} GPUs)")
    assert_bit_equal(gc, oc, f"dye ({mode}, {world} GPUs)")
    assert_bit_equal(gp, op, f"pressure ({mode}, {world} GPUs)")
