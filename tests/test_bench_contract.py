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


def test_bench_cli_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout
