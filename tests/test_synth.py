"""Seeded synthetic inputs and the host-side restatements of the reference's own inputs."""
import numpy as np
import pytest


def test_generators_are_windowable_and_deterministic():
    from esp32_fluid_simulation_b200 import synth
    v = synth.velocity(300, 200)
    c = synth.dye(300, 200)
    assert np.array_equal(synth.velocity(300, 200, window=(17, 33, 40, 50)), v[33:83, 17:57])
    assert np.array_equal(synth.dye(300, 200, window=(17, 33, 40, 50)), c[33:83, 17:57])
    assert abs(v).max() <= 60.0 and c.max() <= synth.DYE_CAP
    d = synth.drags(61, 81, 5)
    assert (d["cy"] < 61).all() and (d["cx"] < 81).all() and np.array_equal(d, synth.drags(61, 81, 5))
    assert synth.dye(64, 64, saturate=True).max() == 0xFFFFFFFF


@pytest.mark.parametrize("shape", [(61, 81), (5, 4), (16, 9)])
def test_color_wheel_matches_the_oracle_restatement(oracle, shape):
    """ino:196-241 restated twice (C in the oracle, numpy here); the sketch itself cannot compile off-device."""
    from esp32_fluid_simulation_b200 import synth
    v, c = synth.color_wheel(*shape)
    ov, oc = oracle.init_color_wheel(*shape)
    assert np.array_equal(v, ov) and np.array_equal(c, oc)
    if shape == (61, 81):
        assert c.max() == 0xFFFFFFFF and (c.sum(axis=2, dtype=np.uint64) > 0).all()


def test_arduino_map_and_touch_drags():
    from esp32_fluid_simulation_b200 import synth
    assert synth.arduino_map(200, 200, 3700, 0, 81) == 0
    assert synth.arduino_map(3700, 200, 3700, 0, 81) == 81         # inclusive upper bound: off the grid
    assert synth.arduino_map(1950, 200, 3700, 0, 81) == 40
    assert synth.arduino_map(100, 200, 3700, 0, 81) == -2          # truncation toward zero, like C
    samples = [(1, 1950, 2020), (1, 1993, 2020), (0, 0, 0), (1, 500, 500), (1, 500, 559), (1, 3700, 3800)]
    d = synth.touch_drags(samples, 61, 81)
    assert len(d) == 3                                             # a drag needs two consecutive touched samples
    assert (d[0]["cx"], d[0]["cy"]) == (41, 30) and d[0]["vx"] == np.float32(100.0) and d[0]["vy"] == 0.0
    assert d[1]["vy"] == np.float32(100.0)
    assert d[2]["cx"] == 81 and d[2]["cy"] == 61                   # the reference's out-of-range case (ino:77-78)


def test_out_of_range_touch_is_dropped_by_the_step(oracle):
    from esp32_fluid_simulation_b200 import synth
    v = np.zeros((81, 61, 2), np.float32)
    d = synth.touch_drags([(1, 3650, 3750), (1, 3700, 3800)], 61, 81)
    before = v.copy()
    oracle.apply_drags(v, d)
    assert np.array_equal(v, before)
