"""The bench.py output contract, checked on the committed B200 lines (profiles/) and on the argument parser:
keys the driver reads, algorithmic-FLOP bookkeeping, and that the reference arm's config matches ours."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline")


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name,n", [("r01_bench_final_n1.json", 1), ("r01_bench_final_n2.json", 2),
                                    ("r01_bench_final_n4.json", 4), ("r01_bench_final_n8.json", 8)])
def test_committed_bench_lines_follow_the_contract(name, n):
    d = _line(name)
    for k in REQUIRED:
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "weak" and d["higher_is_better"] is True
    assert d["config"]["global_batch"] == n * d["config"]["batch_per_gpu"] and "workload" in d["config"]
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 100 * d["steps"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 9e7 and e["d2h_bytes_per_step"] == 4 and e["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # 759.0 MFLOP per window (lines written before the numerator was pinned to it carry the
    # launches' own count, 1.1 % higher)
    assert abs(r["alg_flops_per_step"] - 759.0e6 * d["config"]["batch_per_gpu"]) < 0.02 * r["alg_flops_per_step"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0


def test_reference_arm_line_matches_our_config():
    ours, ref = _line("r01_bench_final_n1.json"), _line("r01_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"]["workload"] == ours["config"]["workload"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0
    assert ref["cpu_baseline"]["value"] == ref["value"]


@pytest.mark.parametrize("name,n", [("r02_bench_n1.json", 1), ("r02_bench_n2.json", 2), ("r02_bench_n4.json", 4),
                                    ("r02_bench_n8.json", 8)])
def test_round2_bench_lines_follow_the_contract(name, n):
    """Round-2 additions: sustained (>= 3 s) figure, burst peak, the static-traffic label, the per-kernel HBM table, and —
    on one GPU — the same-box comparators (reference on the CPU, unmodified reference eager on the GPU)."""
    d = _line(name)
    for k in REQUIRED + ("sustained",):
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert d["config"]["global_batch"] == n * d["config"]["batch_per_gpu"] and d["config"]["bench_config"] == "2"
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["sustained"]["seconds"] >= 3.0 and d["sustained"]["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["alg_flops_per_step"] - 759.0e6 * d["config"]["batch_per_gpu"]) < 1e-6 * r["alg_flops_per_step"]
    assert "burst" in r["peak_source"] and r["frac_of_sustained_peak"] > r["frac"]
    assert r["traffic"] is None or "STATIC" in r["traffic_note"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 9e7 and e["d2h_bytes_per_step"] == 4 and e["value"] > 0
    if n == 1:
        names = {row["kernel"] for row in d["hbm_kernels"]}
        assert {"bnact_fwd", "bnact_bwd_apply", "optim_step"} <= names
        c = d["cpu_baseline"]
        assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0
        ge = d["gpu_eager_baseline"]
        assert ge["value"] > 0 and d["vs_gpu_eager"] > 50


@pytest.mark.parametrize("name,cfg", [("r02_bench_config3_n1.json", "3"), ("r02_bench_config5_n1.json", "5"),
                                      ("r02_bench_config5w201_n1.json", "5w201"), ("r02_bench_config5_n2.json", "5"),
                                      ("r02_bench_config5_n4.json", "5"), ("r02_bench_config5_n8.json", "5")])
def test_round2_other_configs(name, cfg):
    d = _line(name)
    assert d["config"]["bench_config"] == cfg and d["value"] > 0 and d["roofline"]["frac"] > 0.2
    assert ("bf16" in d["dtype"]) == (cfg == "3")


def test_round2_reference_arm_is_the_staged_reference():
    ours, ref = _line("r02_bench_n1.json"), _line("r02_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["metric"] == ours["metric"] and ref["unit"] == ours["unit"]
    assert ref["config"]["workload"] == ours["config"]["workload"]
    assert ref["config"]["batch_per_gpu"] == ours["config"]["batch_per_gpu"]  # same B = 2048 (round 1 ran 128)
    assert ref["cpu_baseline"]["kind"] == "reference" and ref["cpu_baseline"]["value"] == ref["value"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_cli_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--config", "--no-resident"):
        assert flag in out.stdout
