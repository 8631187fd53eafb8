"""examples/benchmark_c.c: the recipe of benchmark/benchmark.ml:76-99 in plain C over include/hnsw_b200.h —
the boundary has to be usable from the language the OCaml stubs are written in, with no torch or Python
in the process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ocaml-hnsw_b200")


def _compile(tmp_path):
    exe = str(tmp_path / "benchmark_c")
    subprocess.run(["gcc", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "benchmark_c.c"), "-L" + PKG, "-lhnsw_b200",
                    "-Wl,-rpath," + PKG, "-lm", "-o", exe], check=True, capture_output=True, text=True)
    return exe


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_c_driver_links_and_fails_loudly_without_a_gpu(tmp_path):
    exe = _compile(tmp_path)
    deps = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libhnsw_b200.so" in deps and "torch" not in deps and "python" not in deps.lower()
    if _has_gpu():
        pytest.skip("a CUDA device is present: the failure path cannot be observed")
    r = subprocess.run([exe, "500", "16"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3                                   # status from the library, not a crash, not a CPU answer
    assert "no CPU fallback" in r.stderr and "recall" not in r.stdout


@pytest.mark.gpu
def test_c_driver_runs_the_benchmark_recipe(tmp_path):
    exe = _compile(tmp_path)
    r = subprocess.run([exe, "20000", "64", "200", "16", "100", "10", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    recall = float(r.stdout.strip().splitlines()[-1].split()[1])
    assert recall >= 0.9
    assert "max_layer" in r.stdout and "distance evaluations per query" in r.stdout


@pytest.mark.gpu
def test_c_driver_sharded_without_python(tmp_path):
    """--shards 4: the multi-GPU entry points (hnswb200_sharded_*) from plain C, four shards on device 0; with
    --gpus N the same binary spreads them over N GPUs."""
    exe = _compile(tmp_path)
    r = subprocess.run([exe, "--shards", "4", "20000", "64", "200", "16", "100", "10", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "4 shards on device 0" in r.stdout
    assert float(r.stdout.strip().splitlines()[-1].split()[1]) >= 0.9
