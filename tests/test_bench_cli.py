"""bench.py's two arms from the command line, at toy sizes, without a GPU: the reference arm (the
oracle port on host cores) prints exactly one JSON line with the contract's keys; the CUDA arm
refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=600, env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--n", "3000", "--ref-n", "3000", "--nq", "200", "--efc", "40", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["config"]["recall_at_10"] >= 0.95 and j["steps"] == 2 and j["warmup"] == 1


def test_reference_arm_build_is_time_bounded():
    """--ref-build-seconds: the sequential CPU build stops taking rows when its time is up (checked per 20k rows) and the line
    says how many rows the index holds."""
    r = _run(["--impl", "reference", "--n", "50000", "--ref-n", "50000", "--nq", "100", "--efc", "20", "--M", "6", "--steps", "1",
              "--warmup", "1", "--ref-build-seconds", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert j["config"]["index_rows"] == 20000 and j["same_config"] is False


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_cuda_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--n", "1000", "--nq", "10", "--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
