"""The bench.py output contract, checked on the JSON lines the last GPU run committed under profiles/ (no GPU needed):
one line per arm with the keys the driver and the judge read."""
import json
from pathlib import Path

import pytest

PROFILES = Path(__file__).resolve().parents[1] / "profiles"
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "roofline", "clocks", "gpu_launches"}


def _line(name):
    p = PROFILES / name
    if not p.exists():
        pytest.skip(f"{name} not committed")
    lines = [l for l in p.read_text().splitlines() if l.strip()]
    assert len(lines) == 1, "bench.py prints exactly one line on stdout"
    return json.loads(lines[0])


@pytest.mark.parametrize("name", ["r1b_bench_ours.json", "r1b_bench_ours_box2.json"])
def test_our_arm_line(name):
    d = _line(name)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["unit"] == "iterations/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["config"]["workload"] == "netflix" and d["config"]["f"] == 100 and d["config"]["nnz"] == 99072112
    assert d["warmup"] >= 3 and d["value"] == pytest.approx(1e3 / d["ms_per_step"], rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["traffic"] is not None and 0 < r["frac"] < 1
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # uploads and downloads are inside the e2e region
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))


@pytest.mark.parametrize("name", ["r1b_bench_reference.json", "r1b_bench_reference_box2.json"])
def test_reference_arm_line(name):
    d = _line(name)
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    ours = _line("r1b_bench_ours.json")
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == ours[k]
    assert d["config"]["workload"] == ours["config"]["workload"] and d["config"]["nnz"] == ours["config"]["nnz"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]


def test_two_gpu_line_is_whole_job_throughput():
    d = _line("r1b_bench_2gpu_rows.json")
    one = _line("r1b_bench_ours.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["nnz"] == one["config"]["nnz"]
    assert one["value"] < d["value"] < 2.2 * one["value"]
    assert "e2e" in d and d["e2e"]["h2d_bytes_per_step"] > 0


# ---- round 2: both arms print the same `config`, the true warm-up count, per-launch roofline entries --------------------
def test_round2_lines_same_config_in_both_arms():
    ours, ref = _line("r2_bench_ours.json"), _line("r2_bench_reference.json")
    assert ref["impl"] == "reference" and "impl" not in ours
    assert ours["config"] == ref["config"]                                   # the driver's same_config check
    assert ours["steps"] == ref["steps"] and ours["warmup"] == ref["warmup"] >= 3
    for k in ("metric", "unit", "higher_is_better", "scaling", "data"):
        assert ours[k] == ref[k]
    assert ours["dtype"].startswith("f32") and "split-fp16" in ours["dtype"]
    r = ours["roofline"]
    assert r["bound"] == "hbm" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and r["traffic"]
    for side in ("x_side", "theta_side"):
        l = r["launches"][side]
        assert {"ms", "algorithmic_gbs", "frac_hbm", "useful_tflops", "frac_tensor", "dram_bytes", "tensor_pipe_pct", "bound"} <= set(l)
    assert r["launches"]["theta_side"]["bound"].startswith("tensor+l2") and r["launches"]["x_side"]["bound"] == "hbm"
    c = ours["cpu_baseline"]
    assert c["cores"] == 1 and c["kind"] == "port" and c["all_cores"]["cores"] >= 1 and c["all_cores"]["value"] > c["value"]
    e = ours["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < ours["value"]
    assert ours["gpu_launches"] > 0 and ref["gpu_launches"] == 0
    assert ours["value"] > 20 * ref["value"]


def test_round2_two_gpu_line():
    d, one = _line("r2e_bench_2gpu_rows.json"), _line("r2_bench_ours.json")
    assert d["n_gpus"] == 2 and d["config"] == one["config"] and d["scaling"] == "strong"
    assert 1.8 * one["value"] < d["value"] < 2.2 * one["value"]
    assert d["e2e"]["gpus"] == 2 and d["e2e"]["value"] > 0 and d["e2e"]["d2h_bytes_per_step"] * d["steps"] == (d["config"]["m"] + d["config"]["n"]) * d["config"]["f"] * 4


def test_round2_final_lines():
    """The last both-arms run of round 2 (tools/gpu_r2_z.sh): same contract, and the theta-side launch is labelled with the unit
    the ncu capture shows saturated (profiles/traffic.json: shared-memory data pipe 97 %)."""
    ours, ref = _line("r2z_bench_ours.json"), _line("r2z_bench_reference.json")
    assert ref["impl"] == "reference" and "impl" not in ours and ours["config"] == ref["config"]
    assert ours["steps"] == ref["steps"] and ours["warmup"] == ref["warmup"] >= 3
    r = ours["roofline"]
    assert r["bound"] == "hbm" and r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and r["traffic"]
    th, x = r["launches"]["theta_side"], r["launches"]["x_side"]
    assert th["bound"].startswith("shared-memory data pipe") and sum(th["smem_data_pipe_pct"].values()) > 90
    assert x["bound"] == "hbm" and x["dram_bytes"] > 10 * th["dram_bytes"]
    assert ours["clocks"]["samples"] > 0 and ours["clocks"]["sm_mhz"] > 0
    assert ours["e2e"]["value"] < ours["value"] and ours["gpu_launches"] > 0
    assert ref["cpu_baseline"]["kind"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0
