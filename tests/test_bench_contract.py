"""The JSON line bench.py printed on the B200 (kept under profiles/) carries every key of the driver's contract;
`--impl reference` and the argument parser work without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_line(path):
    with open(os.path.join(ROOT, path)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize("path,n", [("profiles/r1_bench_n1.json", 1), ("profiles/r1_bench_n2_fused.json", 2),
                                    ("profiles/r1_bench_n8_fused.json", 8)])
def test_recorded_bench_lines_follow_the_contract(path, n):
    d = last_line(path)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["vs_baseline"] is None and d["warmup"] >= 3
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"})
    if n == 1:
        c = d["cpu_baseline"]
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert d["e2e"]["h2d_bytes_per_step"] == 2 * 8192 * 8192 * 8 and d["e2e"]["d2h_bytes_per_step"] == 8192 * 8192 * 8
        assert d["e2e"]["value"] < d["value"]          # copies inside the timed region: never the device-only number
        assert d["scaling"] == "weak"
    else:
        assert d["scaling"] == "strong" and d["parity"]["panelled_step_equals_single_launch_bits"] is True


def test_reference_arm_runs_on_the_host_cores():
    """`bench.py --impl reference` times the oracle port on the host (no GPU needed) and prints the same line shape."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--size", "512"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
