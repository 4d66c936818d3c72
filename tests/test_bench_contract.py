"""CPU: the committed bench lines (profiles/r01d_*) carry every key of the bench contract — a guard against edits to
bench.py that would silently drop one (the driver parses exactly one JSON line per arm)."""
import json
from pathlib import Path

P = Path(__file__).resolve().parents[1] / "profiles"


def _line(name):
    return json.loads((P / name).read_text().strip().splitlines()[-1])


def test_own_arm_line_has_the_contract_keys():
    d = _line("r01d_bench_n1.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "tokens/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "bf16"
    assert d["data"] == "synthetic" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert d["roofline"]["bound"] in ("hbm", "tensor") and abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value is consistent with ms_per_step and the workload's S (one request of S rows per GPU per step)
    assert abs(d["value"] - d["n_gpus"] * d["config"]["S"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    # the "next" rows report their own rooflines
    assert d["decode"]["bound"] == "hbm" and 0 < d["decode"]["frac"] < 1.2
    assert d["frontend"]["bit_exact_vs_pil_host_path"] is True
    assert set(d["stages"]["roofline_by_stage"]) == {"vit", "projector", "decoder"}


def test_reference_arm_line_has_the_contract_keys():
    r, d = _line("r01d_bench_reference.json"), _line("r01d_bench_n1.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"]
    assert r["config"]["workload"] == d["config"]["workload"] and r["higher_is_better"] == d["higher_is_better"]
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["value"] == r["value"] and r["cpu_baseline"]["kind"] in ("reference", "port")
    assert r["gpu_launches"] == 0
