"""The JSON line bench.py prints, checked on the committed lines of the final build (profiles/r2/): every key of the driver's
contract is present and self-consistent.  CPU only (the lines were produced on B200s; nothing is run here)."""
import json
import os

import pytest

from conftest import ROOT

LINES = ["bench_n1_with_filon.json", "bench_n2_weak_final.json", "bench_n8_weak.json"]


def last_line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", "r2", name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", LINES)
def test_contract_keys(name):
    l = last_line(name)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "cpu_baseline" if l["n_gpus"] == 1 else "e2e", "e2e", "gpu_launches", "clocks"):
        assert key in l, key
    assert l["metric"] == "kmode_hierarchy_solves_per_s" and l["higher_is_better"] is True and l["dtype"] == "f64"
    assert l["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert l["warmup"] >= 3 and l["gpu_launches"] > 0
    assert "workload" in l["config"] and "model" not in l["config"]
    r = l["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = l["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["unit"] == l["unit"]
    assert e["value"] != l["value"]                      # measured separately, not a copy of the device-timed value
    c = l["clocks"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_value_is_units_over_time():
    l = last_line("bench_n1_with_filon.json")
    assert abs(l["value"] - 2000 * l["n_gpus"] / (1e-3 * l["ms_per_step"])) < 1e-6 * l["value"]
    cb = l["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"]


def test_weak_scaling_line_is_whole_job_throughput():
    one, eight = last_line("bench_n1_with_filon.json"), last_line("bench_n8_weak.json")
    assert eight["scaling"] == "weak" and eight["n_gpus"] == 8
    assert 0.85 < eight["value"] / (8 * one["value"]) < 1.1
    s = eight["strong_scaling"]
    assert s["n_gpus"] == 8 and s["speedup_vs_one_gpu"] > 3 and s["max_rel_diff_tt_ee_vs_single_gpu"] < 1e-4


def test_strong_scaling_c4_line():
    """BASELINE configs[3] (bench.py --scaling strong): total work fixed, value = 10^4 modes / time per spectrum set."""
    one, eight = last_line("bench_n1_strong_c4.json"), last_line("bench_n8_strong_c4.json")
    assert eight["scaling"] == "strong" and one["scaling"] == "strong"
    assert abs(eight["value"] - 1e4 / (1e-3 * eight["ms_per_step"])) < 1e-6 * eight["value"]
    assert eight["value"] / one["value"] > 4
