"""Pins the oracle (oracle/fluid_oracle.c) against fixtures generated from the
reference's own sources (tests/golden/make_golden.py) and against the known-
answer facts of SURVEY.md §8(c).  CPU only."""
import numpy as np
import pytest

from conftest import assert_bit_equal

DT = np.float32(1 / 30.0)
SHAPES = [(2, 2), (3, 2), (2, 5), (7, 3), (5, 4), (33, 17), (80, 60)]


def test_regression_20_steps_61x81(port, golden):
    g = golden("regress20_61x81.npz")
    v, c = g["v0"].copy(), g["c0"].copy()
    for _ in range(20):
        v, c, p, d = port.step(v, c, None, DT, 1.0, 10, 1.96, want_fields=True)
    for name, got in (("v", v), ("c", c), ("p", p), ("d", d)):
        assert_bit_equal(got, g[name], name)


@pytest.mark.parametrize("shape", SHAPES)
def test_operators_match_reference_fixtures(port, golden, shape):
    g = golden("ops_small.npz")
    k = f"{shape[0]}x{shape[1]}"
    v, c = g[k + "_v"], g[k + "_c"]
    for ns in (0, 1):
        assert_bit_equal(port.advect_vec2f(v, v, DT, ns), g[k + f"_advv_ns{ns}"], f"advect v ns={ns}")
        assert_bit_equal(port.advect_rgb_uq32(c, v, DT, ns), g[k + f"_advc_ns{ns}"], f"advect c ns={ns}")
    div = port.calculate_divergence(v, 1.0)
    assert_bit_equal(div, g[k + "_div"], "div")
    p10 = port.poisson_solve(div, 1.0, 10, 1.96)
    assert_bit_equal(p10, g[k + "_p_k10"], "p k=10")
    assert_bit_equal(port.poisson_solve(div, 1.0, 1, 1.0), g[k + "_p_k1_w1"], "p k=1 w=1")
    assert_bit_equal(port.poisson_solve(div, 2.0, 3, 1.5), g[k + "_p_dx2"], "p dx=2")
    assert_bit_equal(port.subtract_gradient(v.copy(), p10, 1.0), g[k + "_grad"], "grad")
    assert_bit_equal(port.subtract_gradient(v.copy(), p10, 2.0), g[k + "_grad_dx2"], "grad dx=2")


@pytest.mark.parametrize("shape", [(33, 17), (80, 60)])
def test_steps_with_drags(port, golden, shape):
    g = golden("steps_drags.npz")
    k = f"{shape[0]}x{shape[1]}"
    v, c, dr = g[k + "_v0"].copy(), g[k + "_c0"].copy(), g[k + "_drags"]
    for _ in range(5):
        v, c, p, d = port.step(v, c, dr, DT, 1.0, 10, 1.96, want_fields=True)
    for name, got in (("v", v), ("c", c), ("p", p), ("d", d)):
        assert_bit_equal(got, g[k + "_" + name], name)


def test_sample_edge_semantics(port, golden):
    g = golden("sample_edges.npz")
    v, c, pts = g["v"], g["c"], g["pts"]
    for ns in (0, 1):
        sv = np.stack([port.sample_vec2f(v, float(a), float(b), ns) for a, b in pts])
        sc = np.stack([port.sample_rgb_uq32(c, float(a), float(b), ns) for a, b in pts])
        assert_bit_equal(sv, g["sv"][ns], f"sample vec2 ns={ns}")
        assert_bit_equal(sc, g["sc"][ns], f"sample rgb ns={ns}")
    assert [port.uq32_from_float(float(x)) for x in g["uq_in"]] == g["uq_out"].tolist()
    assert_bit_equal(np.array([port.uq32_to_float(int(r)) for r in g["raw_in"]], np.float32),
                     g["raw_out"], "uq32->float")


# ---- SURVEY.md §8(c) known-answer facts ------------------------------------------------

def test_fact1_sweep_order_5x4(port):
    d = np.ones((4, 5), np.float32)
    p = port.poisson_solve(d, 1.0, 1, 1.0, p=np.full((4, 5), 7.0, np.float32))
    assert p[0, 0] == np.float32(-0.5)                       # even colour sees zeros first
    assert abs(p[0, 1] - (-0.6944)) < 1e-4
    assert abs(p[0, 2] - (-0.3333)) < 1e-4
    assert p[1, 1] == np.float32(-0.25)
    assert abs(p[1, 2] - (-0.5208)) < 1e-4


def test_fact2_half_sweep_order(port):
    rng = np.random.default_rng(2)
    for dim_x, dim_y in [(61, 81), (80, 60), (7, 3), (2, 2)]:
        d = rng.normal(0, 5, (dim_y, dim_x)).astype(np.float32)
        want = port.poisson_solve(d, 1.0, 10, 1.96)
        p = np.zeros_like(d)
        q = np.zeros_like(d)
        for _ in range(10):
            port.sor_half_sweep(p, d, 1.0, 1.96, 0)
            port.sor_half_sweep(p, d, 1.0, 1.96, 1)
            port.sor_half_sweep(q, d, 1.0, 1.96, 1)
            port.sor_half_sweep(q, d, 1.0, 1.96, 0)
        assert_bit_equal(p, want, "even-then-odd")
        assert not np.array_equal(q, want), "odd-then-even must differ"


def test_fact3_zero_velocity(port):
    rng = np.random.default_rng(3)
    v = rng.normal(0, 3, (4, 5, 2)).astype(np.float32)
    z = np.zeros_like(v)
    assert_bit_equal(port.advect_vec2f(v, z, DT, 1), v, "velocity advect under zero velocity")
    c = np.full((4, 5, 3), 0x12345678, np.uint32)
    c[1, 2] = 0xFFFFFF7F
    out = port.advect_rgb_uq32(c, z, DT, 0)
    assert out[0, 0, 0] == 0x12345680                        # rounds to 24 significant bits
    assert out[1, 2, 0] == 0xFFFFFF00
    assert out[3, 4, 0] == 0x12345678                        # corner-copy path is exact
    assert (out[:3, :, :].ravel() != 0x12345678).all()


def test_fact4_no_slip_discount(port):
    v = np.ones((4, 5, 2), np.float32)
    for o, want in [(0.1, 0.8), (0.25, 0.5), (0.49, 0.02), (0.5, 0.0), (3.0, 0.0)]:
        got = port.sample_vec2f(v, -o, 1.5, 1)[0]
        assert abs(got - want) < 1e-6, (o, got)
        assert port.sample_vec2f(v, -o, 1.5, 0)[0] == 1.0  # free-slip: clamped edge value
    assert abs(port.sample_vec2f(v, -0.1, -0.1, 1)[0] - 0.64) < 1e-6   # corner: both axes
    v = np.random.default_rng(4).normal(0, 1, (4, 5, 2)).astype(np.float32)
    a = port.sample_vec2f(v, 4.0, 1.5, 0)                  # i == dim_x-1 takes the edge path
    b = port.sample_vec2f(v, np.float32(3.9999998), 1.5, 0)
    assert np.allclose(a, b, atol=1e-5)


def test_fact5_uq32_saturation(port):
    assert port.uq32_from_float(4294967040.0) == 0xFFFFFF00
    assert port.uq32_from_float(4294967296.0) == 0xFFFFFFFF
    assert port.uq32_from_float(-3.0) == 0
    assert port.uq32_to_float(0xFFFFFFFF) == 4294967296.0
    c = np.full((4, 5, 3), 0xFFFFFFFF, np.uint32)
    v = np.full((4, 5, 2), 7.7, np.float32)
    assert (port.advect_rgb_uq32(c, v, DT, 0) == 0xFFFFFFFF).all()


def test_drags_semantics(port):
    from oracle import DRAG_DTYPE
    v = np.zeros((4, 5, 2), np.float32)
    d = np.zeros(4, DRAG_DTYPE)
    d[0] = (2, 3, 10.0, 20.0)      # cx=2 (j), cy=3 (i)
    d[1] = (2, 3, 11.0, 21.0)      # same node: the later record wins
    d[2] = (4, 0, 1.0, 2.0)        # cx == dim_y: out of range, dropped
    d[3] = (0, 5, 1.0, 2.0)        # cy == dim_x: out of range, dropped
    port.apply_drags(v, d)
    assert v[2, 3].tolist() == [21.0, 11.0]                   # SET, x/y swapped
    assert np.count_nonzero(v) == 2


def test_upscale_known_answers(port):
    c = np.zeros((3, 2, 3), np.uint32)                        # dim_x=2, dim_y=3
    c[..., 0] = 0xF8000000
    c[..., 1] = 0xFC000000
    c[..., 2] = 0xF8000000
    img = port.upscale4_rgb565(c)
    assert img.shape == (4, 8)
    assert (img == 0xFFFF).all()
    c[:] = 0
    c[0, 0, 0] = 0x80000000                                  # red ramp from node (0,0)
    img = port.upscale4_rgb565(c)
    px = lambda w: ((int(w) & 0xFF) << 8 | int(w) >> 8)      # undo the byte swap
    assert px(img[0, 0]) >> 11 == 0x10 and px(img[0, 1]) >> 11 == 0x0C
    assert px(img[1, 0]) >> 11 == 0x0C and px(img[3, 7]) == 0


def test_color_wheel_shape(port):
    v, c = port.init_color_wheel(61, 81)
    assert not v.any()
    assert c.max() == 0xFFFFFFFF and c.shape == (81, 61, 3)
    assert (c.sum(axis=2, dtype=np.uint64) > 0).all()


def _upscale_numpy(c):
    """Second, independent restatement of ino:116-177 (vectorised float32 numpy) to cross-check the C
    oracle: the draw routine lives in the .ino and cannot be compiled off-device, so this arithmetic is
    pinned by two restatements rather than by reference code."""
    f32 = np.float32
    c = c.astype(np.float32)                                   # UQ32 -> float, uq32.h:15
    dim_y, dim_x = c.shape[:2]
    c11, c21 = c[:-1, :-1], c[:-1, 1:]                          # (i, j), (i+1, j)   [array index = (j, i)]
    c12, c22 = c[1:, :-1], c[1:, 1:]                            # (i, j+1), (i+1, j+1)
    quarter = f32(0.25)

    def ramp(a, b):                                             # c, c+=dc, ... accumulated (ino:133-152)
        d = (b - a) * quarter
        out = [a]
        for _ in range(3):
            out.append(out[-1] + d)
        return out

    left, right = ramp(c11, c21), ramp(c12, c22)                # 4 sub-rows ii along the fast axis i
    img = np.zeros(((dim_x - 1) * 4, (dim_y - 1) * 4), np.uint16)
    for ii in range(4):
        a, b = left[ii], right[ii]
        d = (b - a) * quarter                                   # ino:157
        vals = [a]
        for _ in range(3):
            vals.append(vals[-1] + d)                           # ino:160
        for jj in range(4):
            q = vals[jj] + f32(0.5)                             # uq32.h:13
            raw = np.where(q >= f32(4294967296.0), np.uint64(0xFFFFFFFF), q.astype(np.uint64)).astype(np.uint64)
            w = ((raw[..., 0] & 0xF8000000) >> 16) | ((raw[..., 1] & 0xFC000000) >> 21) | ((raw[..., 2] & 0xF8000000) >> 27)
            w = w.astype(np.uint16)
            w = ((w << 8) | (w >> 8)).astype(np.uint16)        # bswap16, ino:173
            img[ii::4, jj::4] = w.T                            # image row = 4i+ii, column = 4j+jj
    return img


@pytest.mark.parametrize("shape", [(2, 2), (5, 4), (61, 81), (33, 100)])
def test_upscale_oracle_matches_independent_numpy_restatement(port, shape):
    dim_x, dim_y = shape
    rng = np.random.default_rng(31)
    c = rng.integers(0, 2 ** 32, (dim_y, dim_x, 3), dtype=np.uint32)
    c[0, 0] = 0xFFFFFFFF                                       # saturating corner
    assert np.array_equal(port.upscale4_rgb565(c), _upscale_numpy(c))


# ---- fixtures generated from the compiled sketch (ESP32-fluid-simulation.ino, oracle/ino_shim.cpp) ------

def test_sketch_fixtures_initial_condition(port, golden):
    from esp32_fluid_simulation_b200 import synth
    g = golden("ino_sketch.npz")
    for key in [k for k in g.files if k.startswith("wheel_")]:
        dim_x, dim_y = (int(t) for t in key[6:].split("x"))
        assert_bit_equal(port.init_color_wheel(dim_x, dim_y)[1], g[key], f"port {key}")
        assert_bit_equal(synth.color_wheel(dim_x, dim_y)[1], g[key], f"synth {key}")


def test_sketch_fixtures_frames(port, golden):
    g = golden("ino_sketch.npz")
    keys = [k for k in g.files if k.startswith("frame_") and k.endswith("_c")]
    assert len(keys) == 4
    for key in keys:
        assert_bit_equal(port.upscale4_rgb565(g[key]), g[key[:-2]], key[:-2])


def test_sketch_fixtures_touch_and_loop(port, golden):
    from esp32_fluid_simulation_b200 import synth
    g = golden("ino_sketch.npz")
    want = g["touch_drags"]
    got = synth.touch_drags([tuple(r) for r in g["touch_script"]], 61, 81)
    assert len(want) == 10 and len(got) > 10                        # the sketch's queue holds 10 (ino:49)
    assert_bit_equal(got[:10].view(np.uint8), want.view(np.uint8), "drag records")
    v, c = g["loop_v0"].copy(), g["loop_c0"].copy()
    for _ in range(3):
        port.step(v, c, g["loop_drags"], DT, 1.0, 10, 1.96)
    assert_bit_equal(v, g["loop_v"], "velocity after 3 loop()s")
    assert_bit_equal(c, g["loop_c"], "dye after 3 loop()s")
