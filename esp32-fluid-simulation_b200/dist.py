"""2-D block decomposition of one big grid over the GPUs of a node.

One process per GPU (torch.distributed).  Rank r owns a rectangle of the global
grid and holds a padded local WINDOW of every field: its rectangle plus `ghost`
nodes towards every neighbouring rank (none towards a domain wall).  All sim
kernels run through the window ("tile") entry points of the C ABI, which
evaluate wall rules, red/black parity and advect coordinates in GLOBAL
coordinates — so a decomposed run is bit-identical to the single-GPU run (and to
the reference).

Exchanges per step (each one batched send/recv to <= 8 neighbours, corners
included, over NCCL on NVLink):
  1. forced velocity, width H+1, so the divergence can be formed on the rectangle
     grown by H (H = half-sweeps fused per SOR pass);
  2. pressure, width H, after every SOR pass but the last; width 1 after the last;
  3. projected velocity + dye, width = this step's max displacement, before the dye
     advect.  The same velocity halo serves the NEXT step's velocity advect.
One scalar all-reduce(max) per step sizes exchange 3.

The host logic here is backend-agnostic: `CudaTileOps` drives the CUDA library,
tests drive the same `DecomposedSim` with a CPU backend over gloo.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# ---------------------------------------------------------------------------------------
# geometry (pure python; no torch)
# ---------------------------------------------------------------------------------------


def process_grid(world: int) -> tuple[int, int]:
    """(Px, Py): ranks along dim_x (fast axis) and dim_y.  1->1x1, 2->1x2, 4->2x2, 8->2x4."""
    px = 1
    while (px * 2) * (px * 2) <= world and world % (px * 2) == 0:
        px *= 2
    return px, world // px


def split(n: int, parts: int, k: int, align: int = 4) -> tuple[int, int]:
    """[lo, hi) of part k when n nodes are cut into `parts` nearly equal pieces whose
    boundaries are multiples of `align` (keeps window rows 16-byte aligned)."""
    def cut(i):
        if i <= 0:
            return 0
        if i >= parts:
            return n
        c = (n * i) // parts
        return (c // align) * align
    return cut(k), cut(k + 1)


@dataclass
class Window:
    """One rank's window; mirrors fs_tile."""
    gdim_x: int
    gdim_y: int
    ox: int
    oy: int
    nx: int
    ny: int
    x0: int
    y0: int
    x1: int
    y1: int

    def rect(self, grow: int = 0) -> "Window":
        """Same window, compute rectangle grown by `grow` (clipped to the window)."""
        return Window(self.gdim_x, self.gdim_y, self.ox, self.oy, self.nx, self.ny,
                      max(self.x0 - grow, 0), max(self.y0 - grow, 0),
                      min(self.x1 + grow, self.nx), min(self.y1 + grow, self.ny))


class Decomposition:
    def __init__(self, gdim_x: int, gdim_y: int, world: int, rank: int, ghost: int = 64,
                 grid: tuple[int, int] | None = None):
        self.px, self.py = grid or process_grid(world)
        assert self.px * self.py == world, (self.px, self.py, world)
        self.world, self.rank, self.ghost = world, rank, ghost
        self.gdim_x, self.gdim_y = gdim_x, gdim_y
        self.rx, self.ry = rank % self.px, rank // self.px
        self.gx0, self.gx1 = split(gdim_x, self.px, self.rx)
        self.gy0, self.gy1 = split(gdim_y, self.py, self.ry)
        gl = ghost if self.rx > 0 else 0
        gr = ghost if self.rx < self.px - 1 else 0
        gd = ghost if self.ry > 0 else 0
        gu = ghost if self.ry < self.py - 1 else 0
        self.window = Window(gdim_x, gdim_y, self.gx0 - gl, self.gy0 - gd,
                             (self.gx1 - self.gx0) + gl + gr, (self.gy1 - self.gy0) + gd + gu,
                             gl, gd, gl + (self.gx1 - self.gx0), gd + (self.gy1 - self.gy0))
        # the same verdict on EVERY rank (a rank-local check would let some ranks go on into a collective or a
        # halo wait while others raise): a strip is cut out of the sender's rectangle, so the ghost width must
        # not exceed the narrowest rectangle of the decomposition along a cut direction
        if world > 1 and ghost > 0:
            ext_x = min(b - a for a, b in (split(gdim_x, self.px, k) for k in range(self.px))) if self.px > 1 else ghost
            ext_y = min(b - a for a, b in (split(gdim_y, self.py, k) for k in range(self.py))) if self.py > 1 else ghost
            if ghost > min(ext_x, ext_y):
                raise ValueError(f"ghost={ghost} is wider than the narrowest rectangle of the {self.px}x{self.py} "
                                 f"decomposition of {gdim_x}x{gdim_y} ({min(ext_x, ext_y)} nodes)")

    def rank_of(self, rx: int, ry: int) -> int | None:
        if 0 <= rx < self.px and 0 <= ry < self.py:
            return ry * self.px + rx
        return None

    def neighbours(self):
        """[(peer_rank, dx, dy)] for the up-to-8 surrounding ranks, in an order that is the
        mirror image on the peer (so batched send/recv pairs line up)."""
        out = []
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == 0 and dy == 0:
                    continue
                peer = self.rank_of(self.rx + dx, self.ry + dy)
                if peer is not None:
                    out.append((peer, dx, dy))
        return out

    def send_slices(self, dx: int, dy: int, width: int):
        """Slices (rows, cols) of the window that neighbour (dx,dy) needs: the strip of MY
        rectangle adjacent to it, `width` deep."""
        w = self.window
        if (dx and width > w.x1 - w.x0) or (dy and width > w.y1 - w.y0):
            raise ValueError(f"strip of {width} nodes does not fit this rank's {w.x1 - w.x0}x{w.y1 - w.y0} rectangle")
        xs = {-1: slice(w.x0, w.x0 + width), 0: slice(w.x0, w.x1), 1: slice(w.x1 - width, w.x1)}[dx]
        ys = {-1: slice(w.y0, w.y0 + width), 0: slice(w.y0, w.y1), 1: slice(w.y1 - width, w.y1)}[dy]
        return ys, xs

    def recv_slices(self, dx: int, dy: int, width: int):
        """Slices of MY ghost region filled by neighbour (dx,dy)."""
        w = self.window
        xs = {-1: slice(w.x0 - width, w.x0), 0: slice(w.x0, w.x1), 1: slice(w.x1, w.x1 + width)}[dx]
        ys = {-1: slice(w.y0 - width, w.y0), 0: slice(w.y0, w.y1), 1: slice(w.y1, w.y1 + width)}[dy]
        return ys, xs


# ---------------------------------------------------------------------------------------
# the decomposed step
# ---------------------------------------------------------------------------------------

class DecomposedSim:
    """State windows + step sequencing for one rank.

    `ops` is the compute backend (tile_* methods with the signatures of
    ops.Context; + `empty(shape, dtype)`, `max_displacement`).  `comm` wraps
    torch.distributed (isend/irecv batch + all-reduce max).
    """

    def __init__(self, dec: Decomposition, ops, comm, iters: int, sor_t: int,
                 dt=np.float32(1 / 30.0), dx=np.float32(1.0), omega=np.float32(1.96),
                 static_halo: int | None = None):
        """static_halo: refresh the WHOLE ghost (static_halo == ghost) for the advects instead of agreeing on
        the step's max displacement (saves one all-reduce + host sync per step).  The kernels raise the
        device status flag (read by `check()`) when a backtrace leaves the window.  None = size every
        advect halo exactly (always correct, one host sync per step)."""
        self.dec, self.ops, self.comm = dec, ops, comm
        self.static_halo = static_halo
        if static_halo is not None and dec.world > 1 and static_halo != dec.ghost:
            # a backtrace that lands between a narrower static halo and the ghost edge would read stale ghosts
            # UNNOTICED (the kernels only flag reads outside the window); fs_dist (NativeDist) passes the valid
            # rectangle to the kernels and may use any width
            raise ValueError("static_halo must equal the ghost width (or be None: exact halos agreed every step)")
        self.iters, self.sor_t = iters, max(1, sor_t)
        self.dt, self.dx, self.omega = dt, dx, omega
        w = dec.window
        need = 2 * min(self.sor_t, max(iters, 1)) + 1
        if dec.world > 1 and dec.ghost < need:
            raise ValueError(f"ghost={dec.ghost} is narrower than the {need} nodes one SOR pass needs")
        self.v = ops.empty((w.ny, w.nx, 2), "float32")
        self.v2 = ops.empty((w.ny, w.nx, 2), "float32")
        self.c = ops.empty((w.ny, w.nx, 3), "uint32")
        self.c2 = ops.empty((w.ny, w.nx, 3), "uint32")
        self.div = ops.empty((w.ny, w.nx), "float32")
        self.p = ops.empty((w.ny, w.nx), "float32")
        self.p2 = ops.empty((w.ny, w.nx), "float32")
        self.v_halo = 0          # ghost width of self.v that is currently valid
        self.need_halo = None    # halo the next advect needs (nodes), None = unknown
        self.exchanges = 0

    # -- state --------------------------------------------------------------------------
    def load(self, v_window: np.ndarray, c_window: np.ndarray):
        """Fill the windows (rectangle + whatever ghosts the caller has; ghosts are refreshed
        by exchange before they are read)."""
        self.ops.upload(self.v, v_window)
        self.ops.upload(self.c, c_window)
        self.v_halo = 0
        self.need_halo = None

    def owned(self, field) -> np.ndarray:
        w = self.dec.window
        return self.ops.download(field)[w.y0:w.y1, w.x0:w.x1]

    # -- communication ------------------------------------------------------------------
    def exchange(self, fields, width: int):
        """Refresh `width` ghost nodes of every field in `fields` from the 8 neighbours."""
        if self.dec.world == 1 or width <= 0:
            return
        if width > self.dec.ghost:
            raise RuntimeError(
                f"halo of {width} nodes needed but windows carry ghost={self.dec.ghost}; "
                "re-create the simulation with a larger ghost")
        if hasattr(self.comm, "exchange_fields"):       # peer-memory path: one kernel
            self.comm.exchange_fields(self.dec, fields, width)
            self.exchanges += 1
            return
        sends, recvs = [], []
        for peer, dx, dy in self.dec.neighbours():
            for f in fields:
                ys, xs = self.dec.send_slices(dx, dy, width)
                sends.append((peer, f[ys, xs]))
                ys, xs = self.dec.recv_slices(dx, dy, width)
                recvs.append((peer, f, ys, xs))
        self.comm.exchange(sends, recvs)
        self.exchanges += 1

    def _agree_halo(self, vel) -> int:
        if self.static_halo is not None:
            return self.static_halo if self.dec.world > 1 else 0
        local = self.ops.max_displacement(vel, self.dec.window, self.dt)
        return self.comm.all_max(local)

    def check(self):
        """Raise if any advect since the last check read outside its window (static_halo too small)."""
        self.ops.tile_check_now()

    # -- one loop() body (ino:249-289) ------------------------------------------------------
    def step(self, drags):
        ops, w = self.ops, self.dec.window
        T, iters = self.sor_t, self.iters
        # advect velocity (ino:253): needs v valid `need_halo` nodes around the rectangle
        if self.need_halo is None:
            self.need_halo = self._agree_halo(self.v)
        if self.v_halo < self.need_halo:
            self.exchange([self.v], self.need_halo)
            self.v_halo = self.need_halo
        ops.tile_advect(self.v2, self.v, self.v, w, self.dt, True)
        if self.static_halo is None:
            ops.tile_check()
        # drags (ino:264-269): every rank applies the records that land in its rectangle
        if drags is not None and len(drags):
            ops.tile_apply_drags(self.v2, drags, w)
        # divergence on the rectangle grown by H so each SOR pass can recompute its halo
        passes = [min(T, iters - k) for k in range(0, iters, T)] if iters > 0 else []
        H0 = 2 * passes[0] if passes else 0
        self.exchange([self.v2], H0 + 1)
        ops.tile_calculate_divergence(self.div, self.v2, w.rect(H0), self.dx)          # ino:274
        # SOR (ino:275): p starts at zero; ping-pong p/p2, exchange H ghosts between passes
        src, dst = None, self.p
        if not passes:
            ops.zero(self.p)
        for k, t in enumerate(passes):
            ops.tile_sor_sweeps(dst, src, self.div, w, self.dx, self.omega, 0, 2 * t)
            nxt = 2 * passes[k + 1] if k + 1 < len(passes) else 1
            self.exchange([dst], nxt)
            src, dst = dst, (self.p2 if dst is self.p else self.p)
        p_final = src if src is not None else self.p
        ops.tile_subtract_gradient(self.v2, p_final, w, self.dx)                       # ino:276
        self.p_last = p_final
        # dye advect (ino:282) with the projected velocity; its halo also serves the next step
        h = self._agree_halo(self.v2)
        self.exchange([self.v2, self.c], h)
        ops.tile_advect(self.c2, self.c, self.v2, w, self.dt, False)
        if self.static_halo is None:
            ops.tile_check()
        self.v, self.v2 = self.v2, self.v
        self.c, self.c2 = self.c2, self.c
        self.v_halo, self.need_halo = h, h


# ---------------------------------------------------------------------------------------
# torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests)
# ---------------------------------------------------------------------------------------

class TorchComm:
    def __init__(self, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device

    def exchange(self, sends, recvs):
        torch, dist = self.torch, self.dist
        ops, landing = [], []
        for peer, view in sends:
            ops.append(dist.P2POp(dist.isend, view.contiguous(), peer))
        for peer, field, ys, xs in recvs:
            buf = torch.empty_like(field[ys, xs])
            ops.append(dist.P2POp(dist.irecv, buf, peer))
            landing.append((field, ys, xs, buf))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for field, ys, xs, buf in landing:
            field[ys, xs] = buf

    def all_max(self, value: int) -> int:
        t = self.torch.tensor([int(value)], dtype=self.torch.int32, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())


class CudaTileOps:
    """Backend of DecomposedSim over the CUDA library (tile entry points of the C ABI).
    The context runs on torch's CURRENT stream so NCCL ops order with the kernels."""

    def __init__(self, device: int):
        import torch

        from . import ops as _ops
        from ._lib import Tile
        self.torch, self.Tile = torch, Tile
        self.device = torch.device("cuda", device)
        self.ctx = _ops.Context(device, torch.cuda.current_stream(self.device))

    def _tile(self, w: Window):
        return self.Tile(w.gdim_x, w.gdim_y, w.ox, w.oy, w.nx, w.ny, w.x0, w.y0, w.x1, w.y1)

    def empty(self, shape, dtype):
        tdt = {"float32": self.torch.float32, "uint32": self.torch.int32}[dtype]
        return self.torch.zeros(shape, dtype=tdt, device=self.device)

    def zero(self, t):
        t.zero_()

    def upload(self, dst, a: np.ndarray):
        if a.dtype == np.uint32:
            a = a.view(np.int32)
        dst.copy_(self.torch.from_numpy(np.ascontiguousarray(a)))

    def download(self, t) -> np.ndarray:
        a = t.cpu().numpy()
        return a.view(np.uint32) if a.dtype == np.int32 else a

    def tile_advect(self, next_p, p, vel, w, dt, no_slip):
        self.ctx.tile_advect(next_p, p, vel, self._tile(w), dt, no_slip)

    def tile_check(self):
        pass  # the halo is sized from an agreed max displacement; an overrun cannot happen

    def tile_check_now(self):
        self.ctx.tile_check()

    def tile_apply_drags(self, v, drags, w):
        self.ctx.tile_apply_drags(v, drags, self._tile(w))

    def tile_calculate_divergence(self, div, v, w, dx):
        self.ctx.tile_calculate_divergence(div, v, self._tile(w), dx)

    def tile_subtract_gradient(self, v, p, w, dx):
        self.ctx.tile_subtract_gradient(v, p, self._tile(w), dx)

    def tile_sor_sweeps(self, p_out, p_in, div, w, dx, omega, first_parity, n_half):
        self.ctx.tile_sor_sweeps(p_out, p_in, div, self._tile(w), dx, omega, first_parity, n_half)

    def max_displacement(self, vel, w, dt) -> int:
        return self.ctx.tile_max_displacement(vel, self._tile(w), dt)


class ArenaTileOps(CudaTileOps):
    """CudaTileOps whose field windows live in one cudaMalloc'ed arena that the neighbouring ranks
    map through CUDA IPC.  Every rank lays its arena out identically (fields sized for the largest
    window of the decomposition), so a field's offset is the same everywhere."""
    FLAG_BYTES = 256          # 9 uint64 flag slots (one per direction), padded

    def __init__(self, device: int, max_nodes: int, n_vec2=2, n_rgb=2, n_scalar=3):
        super().__init__(device)
        self.max_nodes = max_nodes
        per = lambda b: (max_nodes * b + 255) // 256 * 256
        self.nbytes = self.FLAG_BYTES + n_vec2 * per(8) + n_rgb * per(12) + n_scalar * per(4)
        self.base = self.ctx.arena_alloc(self.nbytes)
        self._cursor = self.FLAG_BYTES

    class _View:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                             "version": 2}

    def empty(self, shape, dtype):
        itemsize = 4
        n = int(np.prod(shape))
        ptr = self.base + self._cursor
        self._cursor += (self.max_nodes * itemsize * (shape[2] if len(shape) == 3 else 1) + 255) // 256 * 256
        assert self._cursor <= self.nbytes, "arena exhausted"
        assert n * itemsize <= self.max_nodes * 12
        view = self._View(ptr, shape, "<f4" if dtype == "float32" else "<i4")
        return self.torch.as_tensor(view, device=self.device)

    def flag_address(self, base: int, dx: int, dy: int) -> int:
        return base + 8 * ((dy + 1) * 3 + (dx + 1))


class PeerComm:
    """Halo exchange by direct stores into the neighbours' windows over NVLink (fs_halo_exchange):
    one kernel per exchange pushes every strip, signals every neighbour and waits for theirs.
    torch.distributed is used once, to ship the 64-byte IPC handles."""

    def __init__(self, ops: ArenaTileOps, world: int, rank: int, gdim_x: int, gdim_y: int, ghost: int,
                 grid=None, peer_bases: dict | None = None, handles: list | None = None):
        self.ops, self.rank, self.world = ops, rank, world
        if ops.ctx.get_option("sor") != 1:
            # the one-launch-per-half-sweep solver seeds and relaxes the rectangle grown by H inside
            # p_out, i.e. it WRITES ghost cells; a neighbour's strip that was stored early would be lost
            raise ValueError("peer-memory halos need the blocked SOR kernel (option sor=1): it is the "
                             "only solver that never writes ghost cells")
        self.decs = [Decomposition(gdim_x, gdim_y, world, r, ghost, grid) for r in range(world)]
        self.seq = 0
        me = self.decs[rank]
        if peer_bases is None:
            if handles is None:
                import torch.distributed as dist
                handles = [None] * world
                dist.all_gather_object(handles, ops.ctx.ipc_export(ops.base))
            peer_bases = {peer: ops.ctx.ipc_open(handles[peer]) for peer, _, _ in me.neighbours()}
            self._opened = list(peer_bases.values())
        self.peer_bases = peer_bases
        self.signal = [ops.flag_address(peer_bases[peer], -dx, -dy) for peer, dx, dy in me.neighbours()]
        self.wait = [ops.flag_address(ops.base, dx, dy) for _, dx, dy in me.neighbours()]

    def exchange_fields(self, dec, fields, width):
        copies = []
        w = dec.window
        for peer, dx, dy in dec.neighbours():
            pw = self.decs[peer].window
            ys, xs = dec.send_slices(dx, dy, width)
            yr, xr = self.decs[peer].recv_slices(-dx, -dy, width)
            for f in fields:
                es = f.element_size() * (f.shape[2] if f.dim() == 3 else 1)
                off = f.data_ptr() - self.ops.base
                src = f.data_ptr() + (ys.start * w.nx + xs.start) * es
                dst = self.peer_bases[peer] + off + (yr.start * pw.nx + xr.start) * es
                copies.append((src, dst, w.nx * es, pw.nx * es, (xs.stop - xs.start) * es, ys.stop - ys.start))
        self.seq += 1
        self.ops.ctx.halo_exchange(copies, self.signal, self.wait, self.seq)

    def all_max(self, value: int) -> int:
        import torch
        import torch.distributed as dist
        t = torch.tensor([int(value)], dtype=torch.int32, device=self.ops.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())

    def close(self):
        for p in getattr(self, "_opened", []):
            self.ops.ctx.ipc_close(p)


class NativeDist:
    """One rank of a decomposed grid behind the C ABI (fs_dist_*, csrc/dist.cu): the whole decomposed
    loop() body — fused advect+drags+divergence on the grown rectangle, blocked SOR passes fused with
    their NVLink halo exchange, gradient-subtract, one velocity+dye exchange kernel, dye advect — is
    sequenced in C++; Python only ships the IPC handles and the drag records."""

    def __init__(self, ctx, gdim_x: int, gdim_y: int, world: int, rank: int, iters: int, ghost: int = 64,
                 advect_halo: int = 40, grid: tuple[int, int] | None = None,
                 dt=np.float32(1 / 30.0), dx=np.float32(1.0), omega=np.float32(1.96), frame: bool = False):
        import ctypes as C

        from . import _lib
        self._C, self._lib, self.ctx = C, _lib, ctx
        self._L = _lib.lib()
        px, py = grid or (0, 0)
        cfg = _lib.DistConfig(gdim_x, gdim_y, world, rank, px, py, ghost, advect_halo, iters, float(dt), float(dx),
                              float(omega), int(bool(frame)))
        h = C.c_void_p()
        _lib.check(self._L.fs_dist_create(C.byref(h), C.byref(cfg), ctx._h), "fs_dist_create")
        self._h = h
        t = _lib.Tile()
        _lib.check(self._L.fs_dist_window(self._h, C.byref(t)), "fs_dist_window")
        self.window = Window(t.gdim_x, t.gdim_y, t.ox, t.oy, t.nx, t.ny, t.x0, t.y0, t.x1, t.y1)
        self.world, self.rank = world, rank

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None):      # the context must outlive its fs_dist (else: leak, not crash)
                self._L.fs_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def info(self) -> dict:
        i = self._lib.DistInfo()
        self._lib.check(self._L.fs_dist_info(self._h, self._C.byref(i)), "fs_dist_info")
        out = {n: getattr(i, n) for n, _ in i._fields_}
        out["phase_ms"] = dict(zip(("advect_drags_div", "sor_with_fused_exchanges", "gradient", "wait_for_dye_halo",
                                    "dye_advect"), (float(x) for x in i.phase_ms)))
        return out

    def ipc_handle(self) -> bytes:
        buf = self._C.create_string_buffer(64)
        self._lib.check(self._L.fs_dist_ipc_handle(self._h, buf), "fs_dist_ipc_handle")
        return buf.raw

    def connect(self, handles: list[bytes]):
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        self._lib.check(self._L.fs_dist_connect(self._h, blob), "fs_dist_connect")

    def connect_local(self, ranks: list["NativeDist"]):
        arr = (self._C.c_void_p * self.world)(*[r._h.value for r in ranks])
        self._lib.check(self._L.fs_dist_connect_local(self._h, arr), "fs_dist_connect_local")

    @staticmethod
    def _addr(a):
        if type(a).__module__.startswith("torch"):
            assert a.is_contiguous()
            return a.data_ptr()
        assert a.flags.c_contiguous
        return a.ctypes.data

    def upload(self, v_window, c_window):
        """Whole windows [ny, nx, 2] float32 / [ny, nx, 3] uint32 (numpy, pinned or device tensors)."""
        w = self.window
        assert tuple(v_window.shape) == (w.ny, w.nx, 2) and tuple(c_window.shape) == (w.ny, w.nx, 3)
        self._lib.check(self._L.fs_dist_upload(self._h, self._addr(v_window), self._addr(c_window)), "fs_dist_upload")

    def download(self, fields="vcpd", out=None) -> dict:
        """The owned rectangle of the current state (numpy; synchronises)."""
        w = self.window
        h, wd = w.y1 - w.y0, w.x1 - w.x0
        res = out or {}
        if "v" in fields and "v" not in res:
            res["v"] = np.empty((h, wd, 2), np.float32)
        if "c" in fields and "c" not in res:
            res["c"] = np.empty((h, wd, 3), np.uint32)
        if "p" in fields and "p" not in res:
            res["p"] = np.empty((h, wd), np.float32)
        if "d" in fields and "d" not in res:
            res["d"] = np.empty((h, wd), np.float32)
        ptr = lambda k: self._addr(res[k]) if k in fields else None   # noqa: E731
        self._lib.check(self._L.fs_dist_download(self._h, ptr("v"), ptr("c"), ptr("p"), ptr("d")), "fs_dist_download")
        return res

    def device_fields(self):
        """Addresses of the CURRENT device windows (v, c, p, div)."""
        C = self._C
        v, c, p, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._lib.check(self._L.fs_dist_device_fields(self._h, C.byref(v), C.byref(c), C.byref(p), C.byref(d)),
                        "fs_dist_device_fields")
        return v.value, c.value, p.value, d.value

    def frame(self):
        """(device address, rows, cols) of this rank's part of the RGB565 frame (created with frame=True)."""
        C = self._C
        p, r, c = C.c_void_p(), C.c_int(), C.c_int()
        self._lib.check(self._L.fs_dist_frame(self._h, C.byref(p), C.byref(r), C.byref(c)), "fs_dist_frame")
        return p.value, r.value, c.value

    def step(self, drags=None):
        from .ops import _drags
        d, n = _drags(drags)
        self._lib.check(self._L.fs_dist_step(self._h, d.ctypes.data if n else None, n), "fs_dist_step")

    def check(self):
        self._lib.check(self._L.fs_dist_check(self._h), "fs_dist_check")


def max_window_nodes(gdim_x, gdim_y, world, ghost, grid=None) -> int:
    return max(d.window.nx * d.window.ny
               for d in (Decomposition(gdim_x, gdim_y, world, r, ghost, grid) for r in range(world)))


# ---------------------------------------------------------------------------------------
# bench.py's N>1 leg
# ---------------------------------------------------------------------------------------

def _device_view(ptr, shape, typestr):
    class _V:
        pass
    v = _V()
    v.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return v


def bench_decomposed(args, tile_edge: int, iters: int, n_drags: int) -> dict:
    """bench.py's N>1 arm.  Weak scaling: every GPU owns tile_edge x tile_edge nodes of one global grid
    (or --global-grid: a fixed global grid).  Default path: fs_dist_* (the decomposed step sequenced in
    C++, SOR passes fused with their NVLink halo exchange).  FS_HALO=nccl / peer-py select the
    Python-sequenced comparison paths."""
    import os
    import sys
    import time

    import torch
    import torch.distributed as dist

    import esp32_fluid_simulation_b200 as fb

    from . import synth
    mode = os.environ.get("FS_HALO", "native")
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    px, py = process_grid(world)
    gx, gy = tile_edge * px, tile_edge * py
    scaling = "weak"
    if getattr(args, "global_grid", ""):
        gx, gy = (int(t) for t in args.global_grid.lower().split("x"))
        scaling = "strong"
    ghost, halo = 64, 40      # +-1000 nodes/s synthetic drags move a node 34 cells; A = 40 leaves D + 1 = 17 for the SOR ring
    if mode == "native":
        # every rank must end up on the same path: agree on whether CUDA IPC works everywhere
        ok, why = 1, ""
        try:
            probe = fb.Context(local_rank, torch.cuda.current_stream(dev))
            base = probe.arena_alloc(1 << 20)
            handle = probe.ipc_export(base)
        except Exception as e:  # noqa: BLE001 — e.g. CUDA IPC not permitted in this container
            ok, why, handle, base = 0, f"{type(e).__name__}: {e}", None, None
        handles = [None] * world
        dist.all_gather_object(handles, handle)
        if ok:
            try:
                peer = (rank + 1) % world
                probe.ipc_close(probe.ipc_open(handles[peer]))
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"{type(e).__name__}: {e}"
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        dist.barrier()
        if base is not None:
            probe.arena_free(base)
        if int(flag.item()) == 0:
            if rank == 0:
                print(f"[bench] peer-memory halos unavailable ({why or 'another rank failed'}); using NCCL send/recv",
                      file=sys.stderr, flush=True)
            mode = "nccl"
    if mode != "native":
        return bench_decomposed_python(args, tile_edge, iters, n_drags, mode)

    parity = None
    verify = getattr(args, "verify_fn", None)        # bench.py's untimed parity leg (it owns the CPU checker)
    if verify is not None and not getattr(args, "no_verify", False):
        parity = verify(fb, torch, dist, local_rank, world, rank, gx, gy, iters, ghost, halo, n_drags)
        if not (parity["oracle_small"] and parity["one_gpu_equals_n"]):
            if rank == 0:
                print(f"[bench] PARITY FAILURE in the decomposed path: {parity}", file=sys.stderr, flush=True)
            dist.barrier()
            raise SystemExit(4)

    ctx = fb.Context(local_rank, torch.cuda.current_stream(dev))
    sor_t = ctx.get_option("sor_t")
    want_frame = bool(getattr(args, "upscale", False))
    sim = NativeDist(ctx, gx, gy, world, rank, iters, ghost=ghost, advect_halo=halo, dt=synth.DT, dx=synth.DX,
                     omega=synth.OMEGA, frame=want_frame)
    handles = [None] * world
    dist.all_gather_object(handles, sim.ipc_handle())
    sim.connect(handles)
    w = sim.window
    hv = torch.from_numpy(synth.velocity(gx, gy, window=(w.ox, w.oy, w.nx, w.ny))).pin_memory()
    hc = torch.from_numpy(synth.dye(gx, gy, window=(w.ox, w.oy, w.nx, w.ny)).view(np.int32)).pin_memory()
    sim.upload(hv, hc)
    drags = [synth.drags(gx, gy, s, n=n_drags) for s in range(args.warmup + args.steps)]
    frame = sim.frame() if want_frame else None   # this rank's part of the 4x RGB565 frame, rendered inside the dye advect

    def one_step(k):
        sim.step(drags[k])

    for s in range(args.warmup):
        one_step(s)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    launches0, ex0 = ctx.launch_count, sim.info["exchanges"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = getattr(args, "clock_sampler", None)
    clk = sampler(local_rank).__enter__() if (sampler and rank == 0) else None
    t0 = time.perf_counter()
    e0.record()
    for s in range(args.warmup, args.warmup + args.steps):
        one_step(s)
    e1.record()
    torch.cuda.synchronize()
    info = sim.info
    # per-phase times of the last timed step, every rank (max and rank 0)
    ph = torch.tensor(list(info["phase_ms"].values()), device=dev, dtype=torch.float64)
    ph_max = ph.clone()
    dist.all_reduce(ph_max, op=dist.ReduceOp.MAX)
    phases = {"rank0_ms": dict(zip(info["phase_ms"].keys(), [float(x) for x in ph.tolist()])),
              "max_over_ranks_ms": dict(zip(info["phase_ms"].keys(), [float(x) for x in ph_max.tolist()]))}
    launches, exchanges = ctx.launch_count - launches0, info["exchanges"] - ex0
    wall_ms = (time.perf_counter() - t0) * 1e3
    pv, pc, pp, pd = sim.device_fields()
    f_v = torch.as_tensor(_device_view(pv, (w.ny, w.nx, 2), "<f4"), device=dev)
    f_d = torch.as_tensor(_device_view(pd, (w.ny, w.nx), "<f4"), device=dev)
    f_p = torch.as_tensor(_device_view(pp, (w.ny, w.nx), "<f4"), device=dev)
    f_p2 = torch.empty_like(f_p)
    tile = fb.Tile(w.gdim_x, w.gdim_y, w.ox, w.oy, w.nx, w.ny, w.x0, w.y0, w.x1, w.y1)
    if clk is not None:
        t_busy = time.time() + 0.3          # keep the device busy so the 100 ms sampler sees loaded clocks
        while time.time() < t_busy:
            ctx.tile_calculate_divergence(f_p2, f_v, tile, synth.DX)
        torch.cuda.synchronize()
        clk.__exit__(None, None, None)
    clocks = clk.summary() if clk is not None else None
    sim.check()                  # no advect of the timed region left its halo, no hand-shake timed out
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), wall_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item()) / args.steps
    nodes = gx * gy
    value = nodes / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the SOR passes), per GPU, local compute only ----
    roofline = None
    try:
        plan = [min(sor_t, iters - k) for k in range(0, iters, sor_t)]
        if len(plan) > 1 and info["sor_passes"] == len(plan) - 1:    # remainder folded into the first pass (dist.cu)
            plan = [plan[0] + plan[-1]] + plan[1:-1]

        def local_solve():
            src, dst = None, f_p
            for tt in plan:
                ctx.tile_sor_sweeps(dst, src, f_d, tile, synth.DX, synth.OMEGA, 0, 2 * tt)
                src, dst = dst, (f_p2 if dst is f_p else f_p)

        for _ in range(2):
            local_solve()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        r0.record()
        for _ in range(reps):
            local_solve()
        r1.record()
        torch.cuda.synchronize()
        sor_ms = r0.elapsed_time(r1) / reps
        own = (w.x1 - w.x0) * (w.y1 - w.y0)
        peak = 6650.0
        try:
            import json as _json
            peak = float(_json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass
        achieved = 12.0 * own * iters / (sor_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "ms": sor_ms, "launches_per_solve": len(plan),
                    "kernel": "sor_blocked_tma_kernel on this rank's window (rank 0; the passes of one solve without "
                              "the halo exchange fused in); algorithmic 12 B/node-iteration",
                    "sor_share_of_step": sor_ms / ms}
    except Exception as e:  # noqa: BLE001 — never let the extra measurement take the bench line down
        roofline = {"error": f"{type(e).__name__}: {e}"}

    # ---- e2e: every rank's state starts and ends in pinned HOST memory, every step ----
    own_h, own_w = w.y1 - w.y0, w.x1 - w.x0
    ov = torch.empty((own_h, own_w, 2), dtype=torch.float32).pin_memory()
    oc = torch.empty((own_h, own_w, 3), dtype=torch.int32).pin_memory()
    e2e_steps = 3

    def e2e_step(k):
        sim.upload(hv, hc)                          # H2D: this rank's window of the state
        one_step(k % len(drags))
        self_out = {"v": ov.numpy(), "c": oc.numpy().view(np.uint32)}
        sim.download("vc", out=self_out)            # D2H: the rectangle this rank owns (synchronises)

    e2e_step(0)
    dist.barrier()
    t1 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    dist.barrier()
    te = torch.tensor([(time.perf_counter() - t1) / e2e_steps], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    sim.check()
    e2e_s = float(te.item())
    own_nodes = own_h * own_w
    e2e = {"value": nodes / e2e_s / 1e6, "unit": "Mcell-steps/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "h2d_bytes_per_step": int(w.nx * w.ny * 20 + 12 * n_drags) * world,
           "d2h_bytes_per_step": int(own_nodes * 20) * world,
           "api": "fs_dist_upload (pinned host window -> device) + fs_dist_step + fs_dist_download (owned rectangle -> "
                  "pinned host), every rank"}
    result = {
        "metric": "Mcell-steps/s (advect+project, 50 SOR iters) at 4096^2", "value": value,
        "unit": "Mcell-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32+uq32", "data": "synthetic",
        "config": {"workload": f"{gx}x{gy} grid block-decomposed {px}x{py}, {own_w}x{own_h} nodes per GPU, "
                               f"{iters} SOR iterations, velocity + dye advection"
                               + (", 4x RGB565 frame every step (rendered inside the dye advect)" if frame is not None else ""),
                   "grid": [gx, gy], "process_grid": [px, py], "ghost": ghost, "sor_t": sor_t,
                   "halo": "fs_dist (C++): SOR passes fused with their NVLink peer-store halo exchange; one exchange "
                           "kernel for velocity + dye",
                   "static_advect_halo": halo, "velocity_halo": info["velocity_halo"], "dye_halo": info["dye_halo"],
                   "div_ring": info["div_ring"],
                   "halo_exchanges_per_step": exchanges / args.steps,
                   "l2": "per-GPU state exceeds the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the compute stream, max over ranks"},
        "wall_ms_per_step_max": float(t[1].item()) / args.steps,
        "gpu_launches": int(launches),
        "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": None, "parity": parity,
        "phases_last_step": phases,
    }
    sim.close()
    del sim
    # ---- BASELINE.json configs[3] / [4] beside the headline (only on the driver's default weak-scaling run) ----
    if not getattr(args, "global_grid", "") and not want_frame and not getattr(args, "no_extra", False):
        extra = {}
        try:
            torch.cuda.empty_cache()
            extra["config3_16384x16384_k50"] = measure_extra_config(fb, torch, dist, ctx, dev, world, rank, 16384, 16384, 50,
                                                                    False, ghost, halo, n_drags, steps=10, warmup=3)
            if world == 8:
                torch.cuda.empty_cache()
                extra["config4_24576x32768_k100_frame"] = measure_extra_config(
                    fb, torch, dist, ctx, dev, world, rank, 24576, 32768, 100, True, ghost, halo, n_drags, steps=5, warmup=2)
        except Exception as e:  # noqa: BLE001 — an extra must never take the headline down (every rank reaches the barrier below)
            extra["error"] = f"{type(e).__name__}: {e}"
        result["extra"] = extra
    return result


def measure_extra_config(fb, torch, dist, ctx, dev, world, rank, gx, gy, iters, frame, ghost, halo, n_drags, steps, warmup):
    """One more decomposed configuration, timed like the headline (CUDA events, max over ranks), with its
    same-size single-GPU denominator: this rank's rectangle stepped as a whole grid on this GPU alone."""
    from . import synth
    sim = NativeDist(ctx, gx, gy, world, rank, iters, ghost=ghost, advect_halo=halo, dt=synth.DT, dx=synth.DX,
                     omega=synth.OMEGA, frame=frame)
    handles = [None] * world
    dist.all_gather_object(handles, sim.ipc_handle())
    sim.connect(handles)
    w = sim.window
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    v = (torch.rand(w.ny, w.nx, 2, device=dev, generator=g) - 0.5) * 120.0      # +-60 nodes/s like synth.velocity
    c = torch.randint(0, 2 ** 31 - 1, (w.ny, w.nx, 3), device=dev, dtype=torch.int32, generator=g)
    sim.upload(v, c)
    del v, c
    drags = [synth.drags(gx, gy, s, n=n_drags) for s in range(warmup + steps)]
    for s in range(warmup):
        sim.step(drags[s])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(warmup, warmup + steps):
        sim.step(drags[s])
    e1.record()
    torch.cuda.synchronize()
    sim.check()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    info = sim.info
    own_w, own_h = w.x1 - w.x0, w.y1 - w.y0
    sim.close()
    del sim
    torch.cuda.empty_cache()
    # the denominator: the same rectangle as a whole grid on one GPU (fs_step_pingpong / fs_step_frame)
    g = torch.Generator(device=dev).manual_seed(7)
    v1 = (torch.rand(own_h, own_w, 2, device=dev, generator=g) - 0.5) * 120.0
    c1 = torch.randint(0, 2 ** 31 - 1, (own_h, own_w, 3), device=dev, dtype=torch.int32, generator=g)
    c2 = torch.empty_like(c1)
    fr = torch.empty((own_w - 1) * 4, (own_h - 1) * 4, dtype=torch.int16, device=dev) if frame else None
    dyes = [c1, c2]

    def one(k):
        if frame:
            ctx.step_frame(v1, dyes[k & 1], dyes[(k & 1) ^ 1], fr, drags[k % len(drags)], own_w, own_h, synth.DT, synth.DX, iters,
                           synth.OMEGA)
        else:
            ctx.step_pingpong(v1, dyes[k & 1], dyes[(k & 1) ^ 1], drags[k % len(drags)], own_w, own_h, synth.DT, synth.DX,
                              iters, synth.OMEGA)

    for k in range(warmup):
        one(k)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for k in range(warmup, warmup + steps):
        one(k)
    s1.record()
    torch.cuda.synchronize()
    t1 = torch.tensor([s0.elapsed_time(s1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t1, op=dist.ReduceOp.MAX)
    ms, ms1 = float(t.item()), float(t1.item())
    del v1, c1, c2, fr
    torch.cuda.empty_cache()
    return {"grid": [gx, gy], "sor_iters": iters, "frame": bool(frame), "n_gpus": world, "nodes_per_gpu": [own_w, own_h],
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "mcell_steps_per_s": gx * gy / (ms * 1e-3) / 1e6,
            "one_gpu_same_rectangle_ms": ms1, "efficiency_vs_one_gpu_same_rectangle": ms1 / ms,
            "exchanges_per_step": info["exchanges_per_step"], "data": "synthetic (device-generated uniform fields)"}


def bench_decomposed_python(args, tile_edge: int, iters: int, n_drags: int, mode: str) -> dict:
    """The Python-sequenced decomposed step (DecomposedSim): NCCL send/recv halos (mode "nccl", kept for
    comparison) or one stand-alone peer-memory exchange kernel per exchange (mode "peer-py")."""
    import os
    import sys
    import time

    import torch
    import torch.distributed as dist

    from . import synth
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    px, py = process_grid(world)
    gx, gy = tile_edge * px, tile_edge * py
    scaling = "weak"
    if getattr(args, "global_grid", ""):
        gx, gy = (int(t) for t in args.global_grid.lower().split("x"))
        scaling = "strong"
    dev = torch.device("cuda", local_rank)
    ghost = 64
    mode = "peer" if mode == "peer-py" else mode
    dec = Decomposition(gx, gy, world, rank, ghost=ghost)
    ops = comm = None
    if mode == "peer":
        # every rank must end up on the same path: agree on whether the IPC mapping worked everywhere
        ok, why, handle = 1, "", None
        try:
            ops = ArenaTileOps(local_rank, max_window_nodes(gx, gy, world, ghost))
            handle = ops.ctx.ipc_export(ops.base)
        except Exception as e:  # noqa: BLE001 — e.g. CUDA IPC not permitted in this container
            ok, why = 0, f"{type(e).__name__}: {e}"
        handles = [None] * world
        dist.all_gather_object(handles, handle)     # every rank takes part, whatever happened above
        if ok and all(h is not None for h in handles):
            try:
                comm = PeerComm(ops, world, rank, gx, gy, ghost, handles=handles)
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"{type(e).__name__}: {e}"
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if rank == 0:
                print(f"[bench] peer-memory halos unavailable ({why or 'another rank failed'}); using NCCL send/recv",
                      file=sys.stderr, flush=True)
            mode, ops, comm = "nccl", None, None
    if mode == "peer":
        static_halo = ghost  # refresh the WHOLE ghost every time: a backtrace is then either served from fresh
                             # data or leaves the window and raises the overrun flag (the +-1000 nodes/s
                             # synthetic drags move a node 34 cells; ghost = 64)
    else:
        ops = CudaTileOps(local_rank)
        comm = TorchComm(dev)
        static_halo = None
    sor_t = ops.ctx.get_option("sor_t")
    sim = DecomposedSim(dec, ops, comm, iters, sor_t, synth.DT, synth.DX, synth.OMEGA, static_halo=static_halo)
    w = dec.window
    sim.load(synth.velocity(gx, gy, window=(w.ox, w.oy, w.nx, w.ny)),
             synth.dye(gx, gy, window=(w.ox, w.oy, w.nx, w.ny)))
    drags = [synth.drags(gx, gy, s, n=n_drags) for s in range(args.warmup + args.steps)]
    frame = None
    if getattr(args, "upscale", False):
        # the rank's part of the 4x RGB565 frame (ino:116-177): upscale the window, ghosts included
        frame = torch.empty((w.nx - 1) * 4, (w.ny - 1) * 4, dtype=torch.int16, device=dev)

    def one_step(k):
        sim.step(drags[k])
        if frame is not None:
            ops.ctx.upscale4_rgb565(frame, sim.c, w.nx, w.ny)

    for s in range(args.warmup):
        one_step(s)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    launches0, ex0 = ops.ctx.launch_count, sim.exchanges
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = getattr(args, "clock_sampler", None)
    clk = sampler(local_rank).__enter__() if (sampler and rank == 0) else None
    t0 = time.perf_counter()
    e0.record()
    for s in range(args.warmup, args.warmup + args.steps):
        one_step(s)
    e1.record()
    torch.cuda.synchronize()
    launches, exchanges = ops.ctx.launch_count - launches0, sim.exchanges - ex0
    wall_ms = (time.perf_counter() - t0) * 1e3
    if clk is not None:
        t_busy = time.time() + 0.3          # keep the device busy so the 100 ms sampler sees loaded clocks
        while time.time() < t_busy:
            ops.ctx.tile_calculate_divergence(sim.div, sim.v, ops._tile(w), synth.DX)
        torch.cuda.synchronize()
        clk.__exit__(None, None, None)
    clocks = clk.summary() if clk is not None else None
    if static_halo is not None:
        sim.check()              # no advect of the timed region read outside its window
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1), wall_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0].item()) / args.steps
    nodes = gx * gy
    value = nodes / (ms * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (the SOR passes), per GPU, local compute only ----
    roofline = None
    try:
        T = sim.sor_t
        plan = [min(T, iters - k) for k in range(0, iters, T)]
        tile = w

        def local_solve():
            src, dst = None, sim.p
            for tt in plan:
                ops.tile_sor_sweeps(dst, src, sim.div, tile, synth.DX, synth.OMEGA, 0, 2 * tt)
                src, dst = dst, (sim.p2 if dst is sim.p else sim.p)

        for _ in range(2):
            local_solve()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        r0.record()
        for _ in range(reps):
            local_solve()
        r1.record()
        torch.cuda.synchronize()
        sor_ms = r0.elapsed_time(r1) / reps
        own = (w.x1 - w.x0) * (w.y1 - w.y0)
        peak = 6650.0
        try:
            import json as _json
            peak = float(_json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass
        achieved = 12.0 * own * iters / (sor_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "ms": sor_ms, "launches_per_solve": len(plan),
                    "kernel": "sor_blocked_tma_kernel on this rank's window (rank 0; the passes of one solve without "
                              "the halo exchanges between them); algorithmic 12 B/node-iteration",
                    "sor_share_of_step": sor_ms / ms}
    except Exception as e:  # noqa: BLE001 — never let the extra measurement take the bench line down
        roofline = {"error": f"{type(e).__name__}: {e}"}

    # ---- e2e: every rank's state starts and ends in pinned HOST memory, every step ----
    hv = torch.from_numpy(synth.velocity(gx, gy, window=(w.ox, w.oy, w.nx, w.ny))).pin_memory()
    hc = torch.from_numpy(synth.dye(gx, gy, window=(w.ox, w.oy, w.nx, w.ny)).view(np.int32)).pin_memory()
    ov = torch.empty((w.y1 - w.y0, w.x1 - w.x0, 2), dtype=torch.float32).pin_memory()
    oc = torch.empty((w.y1 - w.y0, w.x1 - w.x0, 3), dtype=torch.int32).pin_memory()
    e2e_steps = 3

    def e2e_step(k):
        sim.v.copy_(hv, non_blocking=True)          # H2D: this rank's window of the state
        sim.c.copy_(hc, non_blocking=True)
        sim.v_halo, sim.need_halo = 0, None if static_halo is None else static_halo
        one_step(k % len(drags))
        ov.copy_(sim.v[w.y0:w.y1, w.x0:w.x1], non_blocking=True)   # D2H: the rectangle this rank owns
        oc.copy_(sim.c[w.y0:w.y1, w.x0:w.x1], non_blocking=True)
        torch.cuda.synchronize()

    e2e_step(0)
    dist.barrier()
    t1 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    dist.barrier()
    te = torch.tensor([(time.perf_counter() - t1) / e2e_steps], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    own_nodes = (w.x1 - w.x0) * (w.y1 - w.y0)
    e2e = {"value": nodes / e2e_s / 1e6, "unit": "Mcell-steps/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "h2d_bytes_per_step": int(w.nx * w.ny * 20 + 12 * n_drags) * world,
           "d2h_bytes_per_step": int(own_nodes * 20) * world,
           "api": "DecomposedSim: pinned host window -> device, step, owned rectangle -> pinned host (every rank)"}
    return {
        "metric": "Mcell-steps/s (advect+project, 50 SOR iters) at 4096^2", "value": value,
        "unit": "Mcell-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32+uq32", "data": "synthetic",
        "config": {"workload": f"{gx}x{gy} grid block-decomposed {px}x{py}, {w.x1 - w.x0}x{w.y1 - w.y0} nodes per GPU, "
                               f"{iters} SOR iterations, velocity + dye advection"
                               + (", 4x RGB565 frame every step" if frame is not None else ""),
                   "grid": [gx, gy], "process_grid": [px, py], "ghost": dec.ghost, "sor_t": sor_t, "halo": mode, "static_advect_halo": static_halo,
                   "halo_exchanges_per_step": exchanges / args.steps,
                   "l2": "per-GPU state exceeds the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the compute stream, max over ranks"},
        "wall_ms_per_step_max": float(t[1].item()) / args.steps,
        "gpu_launches": int(launches),
        "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": None,
    }
