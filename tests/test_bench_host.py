"""bench.py on the host: the CPU arm's JSON line (the contract the driver parses) and the clock-sample parser.
No GPU involved: `--impl reference` times the oracle port and must never load the CUDA library."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-sample-particles", "8"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["metric"].startswith("particle-steps/sec") and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and abs(d["value"] - d["cpu_baseline"]["value"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_clock_sample_parser():
    sys.path.insert(0, ROOT)
    import bench
    cs = bench.ClockSampler(0)
    good = "0, 1890, 1965, 912.3, 0x0000000000000004, Not Active, Not Active, Not Active, Active"
    bad = "garbage line"
    capped = "0, 1710, 1965, 998.0, 0x4, Not Active, Active, Not Active, Active"
    sm, smax, reasons = cs._parse([good, bad, capped])
    assert sm == [1890.0, 1710.0] and smax == [1965.0, 1965.0]
    assert reasons == {"sw_power_cap", "hw_thermal_slowdown"}
    assert cs.stop()["samples"] == 0          # never started: reports that nvidia-smi was unavailable
