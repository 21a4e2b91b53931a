"""The committed bench lines (profiles/r1_bench_*.json, produced by bench.py on a B200) carry every key of the
measurement contract: metric/value/unit, timing, e2e with byte counts, gpu_launches, clocks, roofline, cpu_baseline."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks"}


def _load(name):
    f = PROFILES / name
    if not f.exists():
        pytest.skip(f"{name} not committed yet")
    return json.loads(f.read_text())


def test_our_single_gpu_line():
    d = _load("r1_bench_ours.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "ours" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1


def test_reference_arm_line():
    d = _load("r1_bench_reference.json")
    assert d["impl"] == "reference" and d["metric"] == _load("r1_bench_ours.json")["metric"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"kind", "cores", "sample", "value"} <= set(d["cpu_baseline"])
    assert d["config"]["workload"] == _load("r1_bench_ours.json")["config"]["workload"]


@pytest.mark.parametrize("name,n", [("r1_bench_ours_2gpu.json", 2), ("r1_bench_ours_4gpu.json", 4)])
def test_multi_gpu_lines(name, n):
    d = _load(name)
    assert d["n_gpus"] == n and d["scaling"] == "strong" and d["impl"] == "ours"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["gpu_launches"] > 0 and "parallelism" in d["config"]
