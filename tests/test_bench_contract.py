"""The committed end-of-round bench lines carry every key the driver's contract names (bench.py itself needs a GPU;
this keeps the line's shape from drifting unnoticed)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_single_gpu_line_has_the_contract_keys():
    d = _line("r02_bench_n1_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is consistent with the timing it was derived from
    nodes = d["config"]["grid"][0] * d["config"]["grid"][1]
    assert abs(d["value"] - nodes / (d["ms_per_step"] * 1e-3) / 1e6) / d["value"] < 1e-6


@pytest.mark.parametrize("n", [2, 4, 8])
def test_multi_gpu_lines_carry_the_parity_leg(n):
    d = _line(f"r02_bench_n{n}_final.json")
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert d["parity"]["oracle_small"] is True and d["parity"]["one_gpu_equals_n"] is True
    assert "phases_last_step" in d and "config3_16384x16384_k50" in d.get("extra", {})
