"""bench.py contract checks that need no GPU: the reference arms print one JSON line with the keys the driver reads, and
the native arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, env=env,
                          timeout=600)


def test_reference_arm_train_line():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-clouds", "2")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clouds/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("point clouds/sec") and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE.json configs[1]" in line["config"]["workload"]


def test_reference_arm_decode_line():
    p = _run("--impl", "reference", "--workload", "decode", "--decode-b", "2", "--decode-n", "128")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "dendrograms/s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == 1 and line["e2e"]["h2d_bytes_per_step"] == 0


def test_native_arm_has_no_cpu_fallback():
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
