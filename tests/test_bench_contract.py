"""CPU: the driver-facing contract of bench.py that does not need a GPU -- the `--impl reference` arm prints one JSON
line with the agreed keys (it times the CPU restatement: the Rust crate cannot be built here), and the product arm
fails loudly without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "2", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-evals/s" and d["unit"] == "particle-evals/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] >= 3 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "C2" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3", env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box this is the real benchmark; tests/ -m gpu and the driver cover it
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0 and r.stdout.strip() == ""
