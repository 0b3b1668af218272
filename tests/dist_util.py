"""Test helpers for the decomposed path: a CPU compute backend over the oracle's window
operators, and an in-process communicator that lets N 'ranks' run as threads."""
import queue
import threading

import numpy as np
import torch


class OracleTileOps:
    """CPU backend of DecomposedSim: torch CPU tensors (so gloo can move them), compute by the
    oracle's oracle_tile_* functions.  Test infrastructure only."""

    def __init__(self, oracle):
        self.o = oracle

    @staticmethod
    def _np(t):
        a = t.numpy()
        return a.view(np.uint32) if a.dtype == np.int32 else a

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype={"float32": torch.float32, "uint32": torch.int32}[dtype])

    def zero(self, t):
        t.zero_()

    def upload(self, dst, a):
        self._np(dst)[...] = a

    def download(self, t):
        return self._np(t).copy()

    def tile_check(self):
        assert not getattr(self, "overrun", False), "advect backtrace left the window"

    def tile_advect(self, next_p, p, vel, w, dt, no_slip):
        hit = self.o.tile_advect(self._np(next_p), self._np(p), self._np(vel), w, dt, no_slip)
        self.overrun = hit
        self.overrun_sticky = getattr(self, "overrun_sticky", False) or hit

    def tile_check_now(self):
        """Same contract as fs_tile_check: report (and clear) any overrun since the last check."""
        hit, self.overrun_sticky = getattr(self, "overrun_sticky", False), False
        if hit:
            raise RuntimeError("advect backtrace left the window (FS_ERR_HALO_OVERRUN)")

    def tile_apply_drags(self, v, drags, w):
        self.o.tile_apply_drags(self._np(v), drags, w)

    def tile_calculate_divergence(self, div, v, w, dx):
        self.o.tile_calculate_divergence(self._np(div), self._np(v), w, dx)

    def tile_subtract_gradient(self, v, p, w, dx):
        self.o.tile_subtract_gradient(self._np(v), self._np(p), w, dx)

    def tile_sor_sweeps(self, p_out, p_in, div, w, dx, omega, first_parity, n_half):
        self.o.tile_sor_sweeps(self._np(p_out), None if p_in is None else self._np(p_in), self._np(div),
                               w, dx, omega, first_parity, n_half)

    def max_displacement(self, vel, w, dt):
        a = np.abs(self._np(vel)[w.y0:w.y1, w.x0:w.x1])
        m = float(a.max()) if a.size else 0.0
        return int(m * abs(float(dt))) + 2       # same rule as fs_tile_max_displacement


class ThreadWorld:
    """N ranks as threads of one process."""

    def __init__(self, world):
        self.world = world
        self.q = {(s, d): queue.Queue() for s in range(world) for d in range(world)}
        self.barrier = threading.Barrier(world)
        self.slots = [0] * world

    def comm(self, rank):
        return ThreadComm(self, rank)


class ThreadComm:
    def __init__(self, world, rank):
        self.w, self.rank = world, rank

    def exchange(self, sends, recvs):
        for peer, view in sends:
            self.w.q[(self.rank, peer)].put(view.clone())
        for peer, field, ys, xs in recvs:
            field[ys, xs] = self.w.q[(peer, self.rank)].get(timeout=120)

    def all_max(self, value):
        self.w.slots[self.rank] = int(value)
        self.w.barrier.wait(timeout=120)
        m = max(self.w.slots)
        self.w.barrier.wait(timeout=120)
        return m


def run_threaded(world, make_sim, n_steps, drags_for_step):
    """Run `world` DecomposedSim instances in lockstep threads; returns the list of sims."""
    tw = ThreadWorld(world)
    sims = [make_sim(r, tw.comm(r)) for r in range(world)]
    errors = []

    def body(sim):
        try:
            for s in range(n_steps):
                sim.step(drags_for_step(s))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            tw.barrier.abort()

    threads = [threading.Thread(target=body, args=(s,)) for s in sims]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return sims


def gather_owned(sims, field_name, gdim_x, gdim_y, channels, dtype):
    shape = (gdim_y, gdim_x, channels) if channels else (gdim_y, gdim_x)
    out = np.zeros(shape, dtype)
    for sim in sims:
        d = sim.dec
        out[d.gy0:d.gy1, d.gx0:d.gx1] = sim.owned(getattr(sim, field_name))
    return out
