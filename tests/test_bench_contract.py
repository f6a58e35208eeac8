"""bench.py contract on a box without a GPU: the reference arm prints ONE JSON line with the keys
the driver reads, on all host cores; the product arm fails loudly (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    # torchrun sets OMP_NUM_THREADS=1 for its workers: the arm must ignore it (round-1 finding)
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", env={"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "connect+collide edge checks/sec"
    assert j["unit"] == "checks/s" and j["higher_is_better"] is True and j["value"] > 0
    assert j["n_gpus"] == 1 and j["steps"] == 1 and j["ms_per_step"] > 0
    assert j["dtype"] == "f64" and j["data"] == "synthetic" and j["vs_baseline"] is None
    assert "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == j["value"] and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) == j["cores_available"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "1")
    assert p.returncode != 0 and not p.stdout.strip()
