import numpy as np


def test_state_dump_round_trip(tmp_path, oracle):
    from esp32_fluid_simulation_b200 import io, synth
    v, c = synth.velocity(61, 81), synth.dye(61, 81)
    v, c, p, d = oracle.step(v, c, None, synth.DT, 1.0, 10, 1.96, want_fields=True)
    io.dump_state(str(tmp_path), v, c, p, d, step=1)
    s = io.load_state(str(tmp_path))
    assert s["params"]["dim_x"] == 61 and s["params"]["dim_y"] == 81 and s["params"]["step"] == 1
    for name, a in (("velocity", v), ("color", c), ("pressure", p), ("divergence", d)):
        assert s[name].dtype == a.dtype and np.array_equal(s[name].view(np.uint32), a.view(np.uint32))
    assert (tmp_path / "sim_color.arr").stat().st_size == 61 * 81 * 12
    # restart from the checkpoint == continuing the run
    v2, c2 = oracle.step(s["velocity"].copy(), s["color"].copy(), None, synth.DT, 1.0, 10, 1.96)
    v3, c3 = oracle.step(v.copy(), c.copy(), None, synth.DT, 1.0, 10, 1.96)
    assert np.array_equal(v2.view(np.uint32), v3.view(np.uint32)) and np.array_equal(c2, c3)
