"""Schedule-independence of fs_dist_step's hand-shakes (csrc/dist.cu), checked on an abstract model.

Since the velocity and dye halo exchanges moved to side streams, a rank runs THREE in-order streams whose kernels
wait for other ranks' flags: the main stream (advect+div, SOR passes with their fused pressure hand-shakes,
gradient, dye advect), the velocity exchange stream and the dye exchange stream, tied together by events.  This
test models exactly that sequence with one Python thread per (rank, stream): fields carry VERSION numbers
instead of data, every kernel asserts that what it reads has the version the sequential algorithm would see,
and every push asserts that nobody is still reading the ghosts it overwrites.  Ranks and streams are slowed
down adversarially.  The model has teeth: dropping one of the two event waits makes it fail (tested).

It is a model of the ORDERING argument in DESIGN.md 5, not of the arithmetic — the arithmetic of the same
sequence is checked bit for bit on the GPU (tests/test_gpu_dist.py, one device; tests/test_gpu_multi.py and
bench.py's parity leg on real GPUs)."""
import random
import threading
import time

import pytest


class Violation(AssertionError):
    pass


class World:
    def __init__(self, px, py, steps, passes, slow, jitter, seed, skip_wait=None, side_delay=0.0, frame=False):
        self.px, self.py, self.n = px, py, px * py
        self.steps, self.passes = steps, passes
        self.slow, self.jitter = slow, jitter
        self.skip_wait = skip_wait                  # "dye" / "velocity" / "frame": leave out that event wait (negative test)
        self.frame = frame                          # the dye advect also renders the frame: its far corners start from
                                                    # velocity GHOSTS of the projected velocity
        # extra delay before every side-stream push (a busy copy path): seconds, or {"v": .., "c": ..} per stream
        self.side_delay = side_delay if isinstance(side_delay, dict) else {"v": side_delay, "c": side_delay}
        self.rng = random.Random(seed)
        self.lock = threading.Lock()
        self.cv = threading.Condition(self.lock)
        self.errors = []
        self.ranks = [Rank(self, r) for r in range(self.n)]

    def neighbours(self, r):
        rx, ry = r % self.px, r // self.px
        out = []
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if (dx or dy) and 0 <= rx + dx < self.px and 0 <= ry + dy < self.py:
                    out.append((ry + dy) * self.px + rx + dx)
        return out

    def pause(self, r, scale=1.0):
        t = self.slow.get(r, 0.0) * scale
        if self.jitter:
            with self.lock:
                t += self.rng.random() * self.jitter
        if t:
            time.sleep(t)

    def fail(self, msg):
        with self.cv:
            self.errors.append(msg)
            self.cv.notify_all()
        raise Violation(msg)

    def wait_until(self, pred, what):
        deadline = time.time() + 20.0
        with self.cv:
            while not pred():
                if self.errors:
                    raise Violation("aborted: " + self.errors[0])
                if not self.cv.wait(timeout=0.05) and time.time() > deadline:
                    self.errors.append("dead-lock: " + what)
                    self.cv.notify_all()
                    raise Violation("dead-lock: " + what)


class Rank:
    def __init__(self, w, r):
        self.w, self.r = w, r
        self.nb = w.neighbours(r)
        # versions: own[field][buf], ghost[field][buf][neighbour]; V[0] / C[0] hold the uploaded state (version 0)
        self.own = {"v": [0, -1], "c": [0, -1]}
        self.ghost = {f: [{q: -1 for q in self.nb}, {q: -1 for q in self.nb}] for f in ("v", "c")}
        self.reading = {"v": [0, 0], "c": [0, 0]}   # kernels in flight that read the GHOSTS of that buffer
        self.flags = [{q: 0 for q in self.nb} for _ in range(3)]   # flag sets 0 (main), 1 (velocity), 2 (dye)
        self.events = set()
        self.seq0 = 0

    # ---- primitives -------------------------------------------------------------------------------------------
    def record(self, name):
        with self.w.cv:
            self.events.add(name)
            self.w.cv.notify_all()

    def wait_event(self, name):
        self.w.wait_until(lambda: name in self.events, f"rank {self.r} waits for event {name}")

    def signal(self, fset, seq):
        with self.w.cv:
            for q in self.nb:
                self.w.ranks[q].flags[fset][self.r] = seq
            self.w.cv.notify_all()

    def wait_flags(self, fset, seq):
        self.w.wait_until(lambda: all(self.flags[fset][q] >= seq for q in self.nb),
                          f"rank {self.r} waits for flag set {fset} seq {seq}")

    def push(self, field, buf, version, expect_old):
        """store this rank's strips of (field, buf) into every neighbour's ghosts"""
        for q in self.nb:
            o = self.w.ranks[q]
            with self.w.lock:
                if o.reading[field][buf]:
                    self.w.errors.append(f"rank {self.r} pushes {field}[{buf}] v{version} into rank {q} while it reads those ghosts")
                old = o.ghost[field][buf][self.r]
                if old not in expect_old:
                    self.w.errors.append(f"rank {self.r} overwrites {field}[{buf}] ghosts of rank {q}: version {old}, expected one of {expect_old}")
                o.ghost[field][buf][self.r] = version
            self.w.pause(self.r, 0.2)
        if self.w.errors:
            self.w.fail(self.w.errors[0])

    def read_with_ghosts(self, field, buf, version, what):
        with self.w.lock:
            bad = [(q, g) for q, g in self.ghost[field][buf].items() if g != version]
            if self.own[field][buf] != version or bad:
                self.w.errors.append(f"rank {self.r} step {version}: {what} reads {field}[{buf}] own v{self.own[field][buf]}, ghosts {bad}")
            self.reading[field][buf] += 1
        if self.w.errors:
            self.w.fail(self.w.errors[0])
        self.w.pause(self.r)                         # the kernel runs for a while
        with self.w.lock:
            self.reading[field][buf] -= 1

    # ---- the three streams of fs_dist_step ---------------------------------------------------------------------------
    def main_stream(self):
        w = self.w
        for s in range(w.steps):
            cur, nxt = s % 2, (s + 1) % 2
            if s == 0:                               # fresh state: velocity + dye ghosts by an exchange on this stream
                self.seq0 += 1
                self.push("v", cur, 0, (-1,))
                self.push("c", cur, 0, (-1,))
                self.signal(0, self.seq0)
                self.wait_flags(0, self.seq0)
            elif w.skip_wait != "velocity":
                self.wait_event(("vx", s - 1))
            self.read_with_ghosts("v", cur, s, "advect+div")          # 1. (writes V[nxt] on the owned rectangle only)
            base = self.seq0
            for k in range(w.passes):                                  # 2. SOR passes with fused pressure hand-shakes
                if k >= 1:
                    self.wait_flags(0, base + k)
                w.pause(self.r, 0.5)
                if k < w.passes - 1:
                    self.signal(0, base + k + 1)
            self.seq0 = base + max(w.passes - 1, 0)
            with w.lock:
                self.own["v"][nxt] = s + 1                             # 3. gradient: the projected velocity
            self.record(("grad", s))
            if s >= 1 and w.skip_wait != "dye":
                self.wait_event(("cx", s - 1))
            if w.frame:
                if w.skip_wait != "frame":
                    self.wait_event(("vx", s))                         # (dist.cu: the main stream waits here with a frame)
                self.read_with_ghosts("v", nxt, s + 1, "frame corners of the dye advect")
            self.read_with_ghosts("c", cur, s, "dye advect")          # 5.
            with w.lock:
                self.own["c"][nxt] = s + 1
            self.record(("dye", s))

    def velocity_stream(self):
        for s in range(self.w.steps):
            nxt = (s + 1) % 2
            self.wait_event(("grad", s))
            self.w.pause(self.r, 0.3)
            time.sleep(self.w.side_delay["v"])
            self.push("v", nxt, s + 1, (s - 1, -1))                   # the neighbours read these ghosts last in step s - 1
            self.signal(1, s + 1)
            self.wait_flags(1, s + 1)
            self.record(("vx", s))

    def dye_stream(self):
        for s in range(self.w.steps):
            nxt = (s + 1) % 2
            self.wait_event(("dye", s))
            self.w.pause(self.r, 0.3)
            time.sleep(self.w.side_delay["c"])
            self.push("c", nxt, s + 1, (s - 1, -1))
            self.signal(2, s + 1)
            self.wait_flags(2, s + 1)
            self.record(("cx", s))


def run_model(px, py, steps=6, passes=3, slow=None, jitter=0.0, seed=1, skip_wait=None, side_delay=0.0, frame=False):
    w = World(px, py, steps, passes, slow or {}, jitter, seed, skip_wait, side_delay, frame)
    threads = []

    def guard(fn):
        def body():
            try:
                fn()
            except Violation:
                pass
            except Exception as e:  # noqa: BLE001
                with w.cv:
                    w.errors.append(f"{type(e).__name__}: {e}")
                    w.cv.notify_all()
        return body

    for rk in w.ranks:
        for fn in (rk.main_stream, rk.velocity_stream, rk.dye_stream):
            threads.append(threading.Thread(target=guard(fn), daemon=True))
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=60)
    assert not any(t.is_alive() for t in threads), "model hung"
    return w


@pytest.mark.parametrize("schedule", ["even", "rank0-slow", "last-slow", "alternate-slow", "jitter", "slow-side-streams"])
@pytest.mark.parametrize("frame", [False, True], ids=["", "frame"])
@pytest.mark.parametrize("px,py,passes", [(1, 2, 3), (2, 2, 1), (2, 4, 4)])
def test_side_stream_exchanges_are_schedule_independent(px, py, passes, schedule, frame):
    n = px * py
    slow = {"even": {}, "rank0-slow": {0: 0.004}, "last-slow": {n - 1: 0.004},
            "alternate-slow": {r: 0.003 for r in range(0, n, 2)}, "jitter": {}, "slow-side-streams": {}}[schedule]
    w = run_model(px, py, steps=6, passes=passes, slow=slow, jitter=0.003 if schedule == "jitter" else 0.0, seed=n,
                  side_delay=0.01 if schedule == "slow-side-streams" else 0.0, frame=frame)
    assert not w.errors, w.errors[:3]
    for rk in w.ranks:                               # every rank got through all steps
        assert rk.own["v"][6 % 2] == 6 and rk.own["c"][6 % 2] == 6


@pytest.mark.parametrize("skip", ["dye", "velocity", "frame"])
def test_model_catches_a_missing_event_wait(skip):
    """Without the main stream's wait for the previous step's exchange, a slow neighbour's ghosts are read stale."""
    late = {"v": 0.03 if skip in ("velocity", "frame") else 0.0, "c": 0.03 if skip == "dye" else 0.0}   # those pushes arrive late
    w = run_model(2, 2, steps=4, passes=2, side_delay=late, skip_wait=skip, frame=skip == "frame")
    assert w.errors and "reads" in w.errors[0], w.errors[:2]
    w = run_model(2, 2, steps=4, passes=2, side_delay=late, frame=skip == "frame")  # with the waits: late, but correct
    assert not w.errors, w.errors[:2]
