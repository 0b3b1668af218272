"""Oracle restatement vs the reference's own sources (oracle/_ref), live, on
random inputs.  Runs only where the reference library was built (this
container); the GPU box relies on the committed fixtures instead."""
import numpy as np
import pytest

from conftest import assert_bit_equal

DT = np.float32(1 / 30.0)


@pytest.mark.parametrize("shape", [(2, 2), (3, 7), (16, 9), (61, 81), (80, 60), (257, 129)])
@pytest.mark.parametrize("vmax", [5.0, 300.0])
def test_step_matches_reference(port, ref, shape, vmax):
    from oracle import DRAG_DTYPE
    dim_x, dim_y = shape
    rng = np.random.default_rng(dim_x * 1000 + dim_y + int(vmax))
    v = ((rng.random((dim_y, dim_x, 2), np.float32) - np.float32(0.5)) * np.float32(2 * vmax))
    c = rng.integers(0, 2 ** 32, (dim_y, dim_x, 3), dtype=np.uint32)
    dr = np.zeros(4, DRAG_DTYPE)
    dr["cx"], dr["cy"] = rng.integers(0, dim_y, 4), rng.integers(0, dim_x, 4)
    dr["vx"], dr["vy"] = rng.normal(0, 500, 4), rng.normal(0, 500, 4)
    va, ca, vb, cb = v.copy(), c.copy(), v.copy(), c.copy()
    for _ in range(4):
        va, ca, pa, da = ref.step(va, ca, dr, DT, 1.0, 10, 1.96, want_fields=True)
        vb, cb, pb, db = port.step(vb, cb, dr, DT, 1.0, 10, 1.96, want_fields=True)
    for name, a, b in (("v", va, vb), ("c", ca, cb), ("p", pa, pb), ("d", da, db)):
        assert_bit_equal(b, a, name)


@pytest.mark.parametrize("no_slip", [0, 1])
def test_advect_both_payloads_both_wall_rules(port, ref, no_slip):
    rng = np.random.default_rng(11 + no_slip)
    v = (rng.normal(0, 120, (40, 50, 2))).astype(np.float32)
    c = rng.integers(0, 2 ** 32, (40, 50, 3), dtype=np.uint32)
    assert_bit_equal(port.advect_vec2f(v, v, DT, no_slip), ref.advect_vec2f(v, v, DT, no_slip), "v")
    assert_bit_equal(port.advect_rgb_uq32(c, v, DT, no_slip), ref.advect_rgb_uq32(c, v, DT, no_slip), "c")


def test_random_sample_points(port, ref):
    rng = np.random.default_rng(5)
    v = rng.normal(0, 1, (6, 7, 2)).astype(np.float32)
    c = rng.integers(0, 2 ** 32, (6, 7, 3), dtype=np.uint32)
    for _ in range(2000):
        i, j = (float(np.float32(x)) for x in rng.uniform(-2, 9, 2))
        for ns in (0, 1):
            assert_bit_equal(port.sample_vec2f(v, i, j, ns), ref.sample_vec2f(v, i, j, ns), f"({i},{j})")
            assert_bit_equal(port.sample_rgb_uq32(c, i, j, ns), ref.sample_rgb_uq32(c, i, j, ns), f"({i},{j})")
