"""bench.py contract on the CPU: the reference arm prints the agreed JSON line; the product arm refuses to run
without a GPU (no CPU fallback); the clock sampler degrades to an explicit "unavailable" record."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=600)


def test_reference_arm_prints_the_contract_line():
    out = run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encode+decode Mpixel/s" and d["unit"] == "Mpixel/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert "workload" in d["config"] and d["config"]["levels"] == 4
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = run("--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu")
    assert out.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in out.stdout.splitlines())


def test_clock_sampler_reports_unavailability():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    s.start()
    r = s.stop()
    assert r["sm_mhz"] is None and r["reasons"]
