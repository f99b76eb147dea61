"""-m "not gpu": the bench contract that can be checked without a GPU -- the reference arm (`bench.py --impl reference`: the
reference's own compiled objects, or the port, on the host cores) prints ONE JSON line with the agreed keys; under a multi-rank
launch only rank 0 runs it; without a CUDA device the GPU arm refuses loudly instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "0.2", "--substeps", "20"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b["impl"] == "reference" and b["metric"].startswith("M particle-contact-updates/sec") and b["unit"] == "M pair-updates/s"
    assert b["higher_is_better"] is True and b["dtype"] == "f64" and b["vs_baseline"] is None and b["value"] > 0
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = b["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == b["value"] and "particles" in cb["sample"]
    assert b["config"]["workload"].startswith("configs[2]") and b["config"]["config_index"] == 2
    assert b["gpu_launches"] == 0


def test_reference_arm_runs_on_rank_0_only():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--size", "0.2"], timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not [l for l in r.stdout.splitlines() if l.startswith("{")]
