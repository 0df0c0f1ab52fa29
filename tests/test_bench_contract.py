"""CPU-only: the bench lines committed under profiles/ carry every key the measurement contract asks for (a regression guard
for bench.py's JSON; the numbers themselves come from B200 runs)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "clocks", "e2e", "gpu_launches"]


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def _latest(name):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r01?_{name}.json")))
    if not files:
        pytest.skip(f"no profiles/*_{name}.json")
    return files[-1]


def test_default_bench_line_has_every_contract_key():
    d = _line(_latest("bench"))
    for k in BASE + ["roofline", "cpu_baseline"]:
        assert k in d, k
    assert d["metric"] == "stitched equirect frames/sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    e = d["e2e"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in e, k
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _line(_latest("bench_reference"))
    assert d["impl"] == "reference" and d["metric"] == "stitched equirect frames/sec" and d["unit"] == "frames/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]


def test_multi_gpu_lines_are_whole_job_rates():
    one = _line(_latest("bench"))["value"]
    for name, n in (("bench_n2_replicas", 2), ("bench_n8_replicas", 8)):
        d = _line(_latest(name))
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert 0.7 * n * one < d["value"] < 1.1 * n * one, (name, d["value"], one)
