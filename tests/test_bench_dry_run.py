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

os.environ["EQS_BENCH_TENSOR_DEVICE"] = "cpu"
torch.cuda.set_device = lambda d: None
torch.cuda.synchronize = lambda *a: None
if int(os.environ.get("WORLD_SIZE", "1")) > 1:       # the multi-rank path: same calls, over gloo and host tensors
    import torch.distributed as dist
    _init = dist.init_process_group
    dist.init_process_group = lambda backend=None, device_id=None, **kw: _init(backend="gloo", **kw)


def shard_range(count, world, rank):
    base, rem = divmod(count, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


class Context:
    def __init__(self, device=0, rank=0, world=1): self.rank, self.world = rank, world
    def measure_pattern(self, n, N): return 6200.0
    def measure_fp64(self): return 34.0


class Report:
    def __init__(self, iters, evals):
        self.iters_run, self.evals, self.init_ms, self.loop_ms = iters, evals, 0.2, 0.25 * iters


class Run:
    def __init__(self, b, n):
        self.b, self.n, self.total, self.last = b, n, 0, 0
        ctx = b.ctx or Context()
        self.first, self.local = shard_range(b.count, ctx.world, ctx.rank)
        self.idx = np.arange(self.first, self.first + self.local, dtype=np.int64)
        if b.mode == "SEQUENTIAL_EXACT" and ctx.world > 1:   # tiles of 256 particles dealt round-robin over the ranks
            allp = np.arange(b.count, dtype=np.int64)
            self.idx = allp[(allp // 256) % ctx.world == ctx.rank]
            self.local = self.idx.size
    def local_indices(self): return self.idx
    def step(self, k=1, *a):
        os.write(1, b"library chatter on fd 1\n")      # what NCCL does with NCCL_DEBUG set
        self.total += k; self.last = k
    def sync(self): pass
    def last_step_ms(self): return 0.25 * self.last, self.last
    def gbest(self): return np.zeros(self.n), 17.5, self.total, 0, 3
    def ring(self): return np.full(50, np.inf), 0
    def pbest_costs(self): return 0.5 * self.idx
    def state(self): return np.zeros(self.n), np.ones(self.n), self.total, 1
    def elites(self):
        idx = np.arange(self.first, self.first + self.local, dtype=np.int64)
        return idx[idx < self.b.k]
    def close(self): pass


class Builder:
    def __init__(self, kind): self.kind, self.iter_max, self.count, self.k, self.ctx, self.mode = kind, 50, 100, 10, None, None
    def __getattr__(self, name):
        if not name.startswith("with_"): raise AttributeError(name)
        def setter(v=None):
            if name == "with_iter_max": self.iter_max = v
            if name in ("with_particle_count", "with_sample_size", "with_sample_count"): self.count = v
            if name == "with_importance_selection_size": self.k = v
            if name == "with_context": self.ctx = v
            if name == "with_mode": self.mode = v
            return self
        return setter
    def start(self, x0): return Run(self, len(x0))
    def solve(self, x0):
        passes = self.iter_max - 1 if self.kind == "ce" else self.iter_max
        self.last_report = Report(passes, 1 + self.count * (1 + passes))
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
E.Context, E.shard_range = Context, shard_range
ED = types.ModuleType("eqsolver_b200.distributed")
ED.env_rank_world = lambda: (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
                             int(os.environ.get("LOCAL_RANK", "0")))
ED.init_context = lambda local: Context(local, *ED.env_rank_world()[:2])
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
    assert set(d["also"]) == {"c4", "c3", "c5", "c1", "c2_sequential_exact", "monte_carlo_1d"}
    for wl, bound in (("c4", "hbm"), ("c3", "fp64"), ("c5", "fp64")):      # every configuration with its own roofline
        rf = d["also"][wl]["roofline"]
        assert rf["bound"] == bound and rf["peak"] > 0 and d["also"][wl]["value"] > 0
        assert rf["frac"] is None or abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    assert d["fp64_peak"]["value"] == 34.0
    # e2e counts the K passes only: N x K evaluations, not N x (K + 1) + 1
    assert abs(d["e2e"]["value"] * d["e2e"]["seconds_per_solve"] - (1 << 20) * 40) < 1.0
    assert "parity_check" not in d                                          # N > 1 only


def test_product_arm_flags():
    r = _dry_run("--steps", "5", "--warmup", "1", "--no-cpu-baseline", "--no-also", "--workload", "c4")
    assert r.returncode == 0, r.stderr[-3000:]
    (line,) = r.stdout.splitlines()
    d = json.loads(line)
    assert d["warmup"] == 3 and d["steps"] == 5        # W >= 3 is enforced
    assert "cpu_baseline" not in d and "also" not in d and d["config"]["n"] == 256


def test_product_arm_cross_entropy_workload():
    r = _dry_run("--steps", "6", "--warmup", "3", "--no-cpu-baseline", "--no-also", "--workload", "c3")
    assert r.returncode == 0, r.stderr[-3000:]
    (line,) = r.stdout.splitlines()
    d = json.loads(line)
    assert d["metric"] == d["unit"] == "sample-evals/s" and d["config"]["n"] == 64 and d["roofline"]["bound"] == "fp64"
    assert d["e2e"]["passes_per_solve"] == 6 and d["gpu_launches"] == 6


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
    assert "cpu_baseline" not in d                                                      # N = 1 only
    # the sharded-vs-unsharded check ran for both halves of the metric and is in the line
    assert d["parity_check"] is True and set(d["parity_detail"]) == {"headline", "c3", "c2_sequential_exact"}
    assert d["parity_detail"]["headline"]["pbest_cost_checksums_equal"] and d["parity_detail"]["c3"]["elite_set_equal"]
    assert d["parity_detail"]["c2_sequential_exact"]["index_sets_equal"] and d["parity_detail"]["c2_sequential_exact"]["ok"]
    assert set(d["also"]) >= {"c4", "c3", "c5", "c1", "c2_sequential_exact", "c2_sequential_exact_sharded"}   # every N carries the other configs


def test_parity_failure_takes_every_rank_out(tmp_path):
    """A sharded result that differs from the unsharded run: rank 0 still prints its line (parity_check false), and EVERY
    rank exits non-zero -- nobody hangs in a barrier."""
    script = tmp_path / "bench_dry_run_driver.py"
    broken = DRIVER.replace("def pbest_costs(self): return 0.5 * self.idx",
                            "def pbest_costs(self): return 0.5 * self.idx + (self.b.ctx.world > 1)")
    assert broken != DRIVER
    script.write_text(f"ROOT = {ROOT!r}\n" + broken)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29568", str(script), "--gpus", "2", "--steps", "5", "--warmup", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["parity_check"] is False and d["parity_detail"]["headline"]["pbest_cost_checksums_equal"] is False
    assert "parity_check FAILED" in r.stderr
