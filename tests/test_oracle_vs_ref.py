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


# ---- the sketch itself: ESP32-fluid-simulation.ino compiled unmodified (oracle/ino_shim.cpp) ----------

INO_SHAPES = [(61, 81), (2, 2), (5, 4), (33, 100), (200, 67)]


@pytest.mark.parametrize("shape", INO_SHAPES)
def test_upscale_restatement_matches_compiled_draw_routine(port, ref, shape):
    """ino:116-177 (a17): oracle_upscale4_rgb565 vs draw_routine() itself."""
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    c = rng.integers(0, 2 ** 32, (shape[1], shape[0], 3), dtype=np.uint32)
    c[0, 0] = 0xFFFFFFFF
    assert_bit_equal(port.upscale4_rgb565(c), ref.ino_draw(c), "RGB565 frame")


@pytest.mark.parametrize("shape", INO_SHAPES)
def test_initial_condition_restatements_match_compiled_setup(port, ref, shape):
    """ino:196-241: oracle_init_color_wheel and synth.color_wheel vs setup() itself."""
    from esp32_fluid_simulation_b200 import synth
    v, c = ref.ino_setup(*shape)
    pv, pc = port.init_color_wheel(*shape)
    assert_bit_equal(pv, v, "velocity (port)")
    assert_bit_equal(pc, c, "dye (port)")
    sv, sc = synth.color_wheel(*shape)
    assert_bit_equal(sv, v, "velocity (synth)")
    assert_bit_equal(sc, c, "dye (synth)")


@pytest.mark.parametrize("shape", [(61, 81), (5, 4), (80, 60), (130, 70)])
def test_step_order_matches_compiled_loop(port, ref, shape):
    """ino:249-289 (a15, a16): oracle_step and ref_shim's restated ORDER vs loop() itself, drags queued."""
    from oracle import DRAG_DTYPE
    rng = np.random.default_rng(shape[0])
    v = ((rng.random((shape[1], shape[0], 2), np.float32) - np.float32(0.5)) * np.float32(180)).astype(np.float32)
    c = rng.integers(0, 2 ** 32, (shape[1], shape[0], 3), dtype=np.uint32)
    dr = np.zeros(9, DRAG_DTYPE)
    dr["cx"], dr["cy"] = rng.integers(0, shape[1], 9), rng.integers(0, shape[0], 9)
    dr["cx"][8], dr["cy"][8] = dr["cx"][0], dr["cy"][0]          # duplicate node: the later record wins
    dr["vx"], dr["vy"] = rng.normal(0, 300, 9), rng.normal(0, 300, 9)
    a, b, s = (v.copy(), c.copy()), (v.copy(), c.copy()), (v.copy(), c.copy())
    for _ in range(4):
        ref.ino_loop(a[0], a[1], dr)
        port.step(b[0], b[1], dr, DT, 1.0, 10, 1.96)
        ref.step(s[0], s[1], dr, DT, 1.0, 10, 1.96)
    assert_bit_equal(b[0], a[0], "velocity (port vs loop())")
    assert_bit_equal(b[1], a[1], "dye (port vs loop())")
    assert_bit_equal(s[0], a[0], "velocity (ref_step vs loop())")
    assert_bit_equal(s[1], a[1], "dye (ref_step vs loop())")


def test_touch_restatement_matches_compiled_touch_routine(ref):
    """ino:63-96: synth.touch_drags vs touch_routine() itself (map(), finite difference, uint16 wrap, queue depth)."""
    from esp32_fluid_simulation_b200 import synth
    rng = np.random.default_rng(5)
    script = [(1, 1950, 2020), (1, 1993, 2020), (0, 0, 0), (1, 500, 500), (1, 500, 559), (1, 3700, 3800),
              (1, 100, 100), (1, 230, 260), (0, 1, 1), (1, 2000, 2000)]
    assert_bit_equal(ref.ino_touch(script).view(np.uint8), synth.touch_drags(script, 61, 81).view(np.uint8), "drags")
    long_script = [(1, int(x), int(y)) for x, y in rng.integers(100, 4000, (30, 2))]
    got, want = ref.ino_touch(long_script), synth.touch_drags(long_script, 61, 81)
    assert len(want) == 29 and len(got) == 10                        # xQueueSend(..., 0) drops when 10 are waiting
    assert_bit_equal(got.view(np.uint8), want[:10].view(np.uint8), "first 10 drags")
