"""Generates the committed golden fixtures from the REFERENCE ITSELF.

Runs oracle/_ref/libfluid_ref.so (the unmodified reference sources behind
oracle/ref_shim.cpp, built by oracle/Makefile from /root/reference) on seeded
inputs and stores inputs + outputs as small .npz files next to this script.
/root/reference does not exist on the GPU box, so the fixtures are what travels.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import DRAG_DTYPE, Ref, build  # noqa: E402

DT = np.float32(1 / 30.0)


def rand_inputs(rng, dim_x, dim_y, vmax):
    v = ((rng.random((dim_y, dim_x, 2), np.float32) - np.float32(0.5)) * np.float32(2 * vmax))
    c = rng.integers(0, 2 ** 32, (dim_y, dim_x, 3), dtype=np.uint32)
    return v.astype(np.float32), c


def rand_drags(rng, dim_x, dim_y, n):
    d = np.zeros(n, DRAG_DTYPE)
    d["cx"] = rng.integers(0, dim_y, n)
    d["cy"] = rng.integers(0, dim_x, n)
    d["vx"] = rng.normal(0, 400, n)
    d["vy"] = rng.normal(0, 400, n)
    return d


def main():
    build(ref=True)
    r = Ref()
    assert r.saturates(), "host lacks AVX-512F: float->uint32 would wrap, fixtures would be wrong"
    rng = np.random.default_rng(20261017)

    # (1) SURVEY.md §8c fact 6: 20-step 61x81 regression from glibc rand() inputs
    v0, c0 = r.fill_rand(61, 81, 1)
    v, c = v0.copy(), c0.copy()
    for _ in range(20):
        v, c, p, d = r.step(v, c, None, DT, 1.0, 10, 1.96, want_fields=True)
    np.savez_compressed(os.path.join(HERE, "regress20_61x81.npz"), v0=v0, c0=c0, v=v, c=c, p=p, d=d)

    # (2) per-operator outputs on several shapes (incl. degenerate), large velocities
    ops = {}
    for dim_x, dim_y in [(2, 2), (3, 2), (2, 5), (7, 3), (5, 4), (33, 17), (80, 60)]:
        v, c = rand_inputs(rng, dim_x, dim_y, 150.0)
        key = f"{dim_x}x{dim_y}"
        ops[key + "_v"] = v
        ops[key + "_c"] = c
        for ns in (0, 1):
            ops[key + f"_advv_ns{ns}"] = r.advect_vec2f(v, v, DT, ns)
            ops[key + f"_advc_ns{ns}"] = r.advect_rgb_uq32(c, v, DT, ns)
        div = r.calculate_divergence(v, 1.0)
        ops[key + "_div"] = div
        ops[key + "_p_k10"] = r.poisson_solve(div, 1.0, 10, 1.96)
        ops[key + "_p_k1_w1"] = r.poisson_solve(div, 1.0, 1, 1.0)
        ops[key + "_p_dx2"] = r.poisson_solve(div, 2.0, 3, 1.5)
        ops[key + "_grad"] = r.subtract_gradient(v.copy(), ops[key + "_p_k10"], 1.0)
        ops[key + "_grad_dx2"] = r.subtract_gradient(v.copy(), ops[key + "_p_k10"], 2.0)
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **ops)

    # (3) full steps with drags (5 steps, 33x17 and 80x60)
    steps = {}
    for dim_x, dim_y in [(33, 17), (80, 60)]:
        v, c = rand_inputs(rng, dim_x, dim_y, 90.0)
        dr = rand_drags(rng, dim_x, dim_y, 6)
        key = f"{dim_x}x{dim_y}"
        steps[key + "_v0"], steps[key + "_c0"], steps[key + "_drags"] = v.copy(), c.copy(), dr
        for _ in range(5):
            v, c, p, d = r.step(v, c, dr, DT, 1.0, 10, 1.96, want_fields=True)
        steps[key + "_v"], steps[key + "_c"], steps[key + "_p"], steps[key + "_d"] = v, c, p, d
    np.savez_compressed(os.path.join(HERE, "steps_drags.npz"), **steps)

    # (4) sample() edge semantics on a 5x4 grid
    v, c = rand_inputs(rng, 5, 4, 10.0)
    pts = []
    for o in (0.0, 0.1, 0.25, 0.49, 0.5, 0.75, 3.0):
        pts += [(-o, 1.3), (4 + o, 1.3), (2.6, -o), (2.6, 3 + o), (-o, -o), (4 + o, 3 + o),
                (-o, 3 + o), (4 + o, -o)]
    pts += [(4.0, 1.5), (3.999, 1.5), (1.5, 3.0), (1.5, 2.999), (0.0, 0.0), (2.25, 1.75)]
    pts = np.array(pts, np.float32)
    sv = np.stack([np.stack([r.sample_vec2f(v, float(a), float(b), ns) for a, b in pts]) for ns in (0, 1)])
    sc = np.stack([np.stack([r.sample_rgb_uq32(c, float(a), float(b), ns) for a, b in pts]) for ns in (0, 1)])
    xs = np.array([0.0, 0.49, 0.5, 1.5, 305419896.0, 4294967040.0, 4294967296.0, 1e20], np.float32)
    uq = np.array([r.uq32_from_float(float(x)) for x in xs], np.uint32)
    raws = np.array([0, 1, 0x12345678, 0xFFFFFF7F, 0xFFFFFF80, 0xFFFFFFFF], np.uint32)
    fl = np.array([r.uq32_to_float(int(x)) for x in raws], np.float32)
    np.savez_compressed(os.path.join(HERE, "sample_edges.npz"), v=v, c=c, pts=pts, sv=sv, sc=sc,
                        uq_in=xs, uq_out=uq, raw_in=raws, raw_out=fl)
    # (5) the SKETCH ITSELF (ESP32-fluid-simulation.ino compiled unmodified by oracle/ino_shim.cpp):
    # setup() initial condition, draw_routine() frames, touch_routine() drag records, loop() steps
    ino = {}
    for dim_x, dim_y in [(61, 81), (5, 4), (16, 9)]:
        _, c = r.ino_setup(dim_x, dim_y)
        ino[f"wheel_{dim_x}x{dim_y}"] = c
    for dim_x, dim_y in [(2, 2), (5, 4), (33, 17), (61, 81)]:
        _, c = rand_inputs(rng, dim_x, dim_y, 1.0)
        c[0, 0] = 0xFFFFFFFF                                    # saturating corner
        ino[f"frame_{dim_x}x{dim_y}_c"] = c
        ino[f"frame_{dim_x}x{dim_y}"] = r.ino_draw(c)
    script = [(1, 1950, 2020), (1, 1993, 2020), (0, 0, 0), (1, 500, 500), (1, 500, 559), (1, 3700, 3800),
              (1, 100, 100), (1, 230, 260), (0, 9, 9), (0, 9, 9), (1, 2000, 2000)]
    script += [(1, int(x), int(y)) for x, y in rng.integers(150, 3900, (14, 2))]   # > 10 records: the queue drops
    ino["touch_script"] = np.array(script, np.int32)
    ino["touch_drags"] = r.ino_touch(script)
    v, c = rand_inputs(rng, 61, 81, 90.0)
    dr = rand_drags(rng, 61, 81, 7)
    ino["loop_v0"], ino["loop_c0"], ino["loop_drags"] = v.copy(), c.copy(), dr
    for _ in range(3):
        r.ino_loop(v, c, dr)
    ino["loop_v"], ino["loop_c"] = v, c
    np.savez_compressed(os.path.join(HERE, "ino_sketch.npz"), **ino)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
