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


# ---------------------------------------------------------------------------------------------------
# One-sided emulation of the peer-memory halo protocol (csrc/halo.cu) on CPU threads.
#
# ThreadComm above has two-sided semantics: data lands in the receiver's ghosts when the RECEIVER
# reaches the exchange.  The real PeerComm is one-sided: a rank stores its strips straight into the
# neighbour's window the moment IT reaches the exchange, then signals a flag and waits for the
# neighbours' flags.  A neighbour that is still busy with earlier kernels gets its ghosts overwritten
# under its feet — harmless only if the step sequence never has it reading (or later rewriting) those
# cells at that point.  This emulation reproduces exactly that, and lets a test slow individual ranks
# down to provoke every "neighbour runs ahead" ordering.
# ---------------------------------------------------------------------------------------------------

class OneSidedWorld:
    def __init__(self, world, slow=None, jitter=0.0, seed=0):
        import random
        self.world = world
        self.ops = [None] * world                     # rank -> OneSidedOps (registers its buffers)
        self.flags = [dict() for _ in range(world)]   # rank -> {direction: last sequence number seen}
        self.cv = threading.Condition()
        self.slow = slow or {}                        # rank -> seconds of delay before every compute op
        self.jitter, self.rng = jitter, random.Random(seed)
        self.failed = False

    def pause(self, rank):
        d = self.slow.get(rank, 0.0)
        if self.jitter:
            with self.cv:
                d += self.rng.random() * self.jitter
        if d:
            import time
            time.sleep(d)


class OneSidedOps(OracleTileOps):
    """OracleTileOps whose buffers are registered by creation order (= the arena offset of the real
    implementation, identical on every rank) and whose compute ops can be delayed."""

    def __init__(self, oracle, world: OneSidedWorld, rank: int):
        super().__init__(oracle)
        self.w, self.rank, self.bufs = world, rank, []
        world.ops[rank] = self

    def empty(self, shape, dtype):
        t = super().empty(shape, dtype)
        self.bufs.append(t)
        return t

    def key(self, tensor):
        for k, b in enumerate(self.bufs):
            if b is tensor:
                return k
        raise KeyError("field is not an arena buffer")

    def _delayed(name):
        def f(self, *a, **kw):
            self.w.pause(self.rank)
            return getattr(OracleTileOps, name)(self, *a, **kw)
        return f

    tile_advect = _delayed("tile_advect")
    tile_apply_drags = _delayed("tile_apply_drags")
    tile_calculate_divergence = _delayed("tile_calculate_divergence")
    tile_subtract_gradient = _delayed("tile_subtract_gradient")

    def tile_sor_sweeps(self, p_out, p_in, div, w, dx, omega, first_parity, n_half):
        """Same WRITE SET as the CUDA blocked kernel: only the compute rectangle of p_out is written.
        (The oracle's window sweep seeds and relaxes the rectangle grown by n_half inside p_out, i.e. it
        writes ghost cells — fine for two-sided exchanges, fatal for one-sided ones: a neighbour's strip
        that arrived early would be overwritten.  The peer path therefore requires the blocked kernel.)"""
        self.w.pause(self.rank)
        tmp = torch.zeros_like(p_out)
        OracleTileOps.tile_sor_sweeps(self, tmp, p_in, div, w, dx, omega, first_parity, n_half)
        p_out[w.y0:w.y1, w.x0:w.x1] = tmp[w.y0:w.y1, w.x0:w.x1]


class OneSidedComm:
    def __init__(self, world: OneSidedWorld, rank: int, decs):
        self.w, self.rank, self.decs, self.seq = world, rank, decs, 0

    def exchange_fields(self, dec, fields, width):
        ops = self.w.ops[self.rank]
        self.seq += 1
        for peer, dx, dy in dec.neighbours():
            ys, xs = dec.send_slices(dx, dy, width)
            yr, xr = self.decs[peer].recv_slices(-dx, -dy, width)
            for f in fields:
                # the store into the neighbour's window happens NOW, whatever the neighbour is doing
                self.w.ops[peer].bufs[ops.key(f)][yr, xr] = f[ys, xs].clone()
            with self.w.cv:
                self.w.flags[peer][(-dx, -dy)] = self.seq
                self.w.cv.notify_all()
        with self.w.cv:
            for _, dx, dy in dec.neighbours():
                if not self.w.cv.wait_for(lambda: self.w.flags[self.rank].get((dx, dy), 0) >= self.seq or self.w.failed,
                                          timeout=120):
                    raise TimeoutError("halo flag never arrived")
                if self.w.failed:
                    raise RuntimeError("another rank failed")

    def all_max(self, value):
        raise AssertionError("the one-sided protocol is used with a static halo: no all-reduce")


def run_one_sided(world, make_sim, n_steps, drags_for_step, osw: OneSidedWorld):
    sims = [make_sim(r) for r in range(world)]
    errors = []

    def body(sim):
        try:
            for s in range(n_steps):
                sim.step(drags_for_step(s))
        except Exception as e:  # noqa: BLE001
            errors.append(e)
            with osw.cv:
                osw.failed = True
                osw.cv.notify_all()

    threads = [threading.Thread(target=body, args=(s,)) for s in sims]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return sims
