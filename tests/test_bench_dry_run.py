"""CPU: the HOST logic of bench.py's product arm, end to end, against a stand-in for the kernel library.

The product arm needs a GPU, so on this side of the fence its bookkeeping (argument handling, the timed region's control
flow, the JSON line and where it is written) would otherwise only ever run on the GPU box.  Here `eqsolver_b200` and the
two torch.cuda calls the single-rank path makes are replaced by fakes that return plausible numbers; nothing is measured
and nothing under oracle/ or the real package is touched.  What is checked is the driver-facing contract: exit code 0,
exactly one line on stdout, valid JSON with the agreed keys, library chatter on fd 1 diverted to stderr."""
import json
import os
import subprocess
import sys

from conftest import ROOT

DRIVER = r'''
import os, sys, types, runpy
import numpy as np
import torch

torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
if int(os.environ.get("WORLD_SIZE", "1")) > 1:       # the multi-rank path: same calls, over gloo and host tensors
    import torch.distributed as dist
    _init, _tensor = dist.init_process_group, torch.tensor
    dist.init_process_group = lambda backend=None, device_id=None, **kw: _init(backend="gloo", **kw)
    torch.tensor = lambda data, device=None, **kw: _tensor(data, **kw)


class Report:
    def __init__(self, iters, evals):
        self.iters_run, self.evals, self.init_ms, self.loop_ms = iters, evals, 0.2, 0.25 * iters


class Run:
    def __init__(self, n):
        self.n, self.total, self.last = n, 0, 0
    def step(self, k=1, *a):
        os.write(1, b"library chatter on fd 1\n")      # what NCCL does with NCCL_DEBUG set
        self.total += k; self.last = k
    def sync(self): pass
    def last_step_ms(self): return 0.25 * self.last, self.last
    def gbest(self): return np.zeros(self.n), 17.5, self.total, 0, 0
    def state(self): return np.zeros(self.n), np.ones(self.n), self.total, 1
    def close(self): pass


class Builder:
    def __init__(self, kind): self.kind, self.iter_max, self.count, self.n = kind, 50, 100, 1
    def __getattr__(self, name):
        if not name.startswith("with_"): raise AttributeError(name)
        def setter(v=None):
            if name == "with_iter_max": self.iter_max = v
            if name in ("with_particle_count", "with_sample_size", "with_sample_count"): self.count = v
            return self
        return setter
    def start(self, x0): return Run(len(x0))
    def solve(self, x0):
        self.last_report = Report(self.iter_max, 1 + self.count * (1 + self.iter_max))
        return np.zeros(len(x0))
    def integrate_with_seed(self, lo, hi, seed):
        self.last_report = Report(1, self.count)
        return types.SimpleNamespace(mean=0.7468, variance=0.04)


class Kind:
    def __init__(self, kind): self.kind = kind
    def new(self, *a): return Builder(self.kind)


class Names:
    def __getattr__(self, name): return name


E = types.ModuleType("eqsolver_b200")
E.ParticleSwarm, E.CrossEntropy, E.MonteCarlo = Kind("pso"), Kind("ce"), Kind("mc")
E.Objective = E.PsoMode = E.MeanDivisor = Names()
ED = types.ModuleType("eqsolver_b200.distributed")
ED.env_rank_world = lambda: (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                             int(os.environ.get("LOCAL_RANK", "0")))
ED.init_context = lambda local: types.SimpleNamespace(measure_pattern=lambda n, N: 6200.0)
E.distributed = ED
sys.modules["eqsolver_b200"], sys.modules["eqsolver_b200.distributed"] = E, ED

sys.argv = ["bench.py"] + sys.argv[1:]
runpy.run_path(os.path.join(ROOT, "bench.py"), run_name="__main__")
'''


def _dry_run(*args):
    code = f"ROOT = {ROOT!r}\n" + DRIVER
    return subprocess.run([sys.executable, "-c", code, *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_product_arm_prints_exactly_one_json_line():
    r = _dry_run("--steps", "40", "--warmup", "3", "--cpu-budget", "0.5")
    assert r.returncode == 0, r.stderr[-3000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, r.stdout[:2000]            # the fake library's chatter on fd 1 must not be here ...
    assert "library chatter on fd 1" in r.stderr       # ... but on stderr
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "also"):
        assert key in d, key
    assert d["metric"] == d["unit"] == "particle-evals/s" and d["n_gpus"] == 1 and d["steps"] == 40 and d["warmup"] == 3
    assert d["dtype"] == "f64" and d["scaling"] == "weak" and d["vs_baseline"] is None and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["value"] > 0
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and d["roofline"]["bound"] == "hbm"
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-12
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["gpu_launches"] == 40 and d["value"] > 0 and d["ms_per_step"] > 0
    assert set(d["also"]) == {"c2_sequential_exact", "c3_cross_entropy", "monte_carlo_1d"}


def test_product_arm_flags():
    r = _dry_run("--steps", "5", "--warmup", "1", "--no-cpu-baseline", "--no-also", "--workload", "c4")
    assert r.returncode == 0, r.stderr[-3000:]
    (line,) = r.stdout.splitlines()
    d = json.loads(line)
    assert d["warmup"] == 3 and d["steps"] == 5        # W >= 3 is enforced
    assert "cpu_baseline" not in d and "also" not in d and d["config"]["n"] == 256


def test_product_arm_two_ranks_over_gloo(tmp_path):
    """The N > 1 control flow (collective decisions, max / sum over ranks, the ranks parting before rank 0 writes its line)
    as two processes over gloo: rank 0 alone prints, both exit 0, nobody is left in a barrier."""
    script = tmp_path / "bench_dry_run_driver.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + DRIVER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29567", str(script), "--gpus", "2", "--steps", "12", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["steps"] == 12 and d["gpu_launches"] == 24           # summed over the ranks
    assert d["config"]["particles_total"] == 2 * d["config"]["particles_per_gpu"]
    assert "cpu_baseline" not in d and "also" not in d                                  # N = 1 only
