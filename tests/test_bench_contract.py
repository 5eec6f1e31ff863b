"""not gpu: the JSON contract of bench.py's reference arm (`--impl reference`: the oracle port of the reference forward +
decode on the host cores; runs without a GPU) and the argument surface the driver uses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py prints ONE JSON line, got %d" % len(lines)
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1", "--ref-crops", "1")
    assert d["impl"] == "reference"
    assert d["metric"].startswith("crops/sec at 384x288") and d["unit"] == "crops/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("BUCTD-CoAM-W48 384x288") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None                     # BASELINE.md holds no published number for this metric


def test_native_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--quick"], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0, "the product path must not fall back to the CPU"
    assert not any(l.startswith("{") for l in out.stdout.splitlines())
