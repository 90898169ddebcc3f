"""CPU: properties of the COMPILED hot loops that DESIGN.md and profiles/r02_sweep.md claim, read from the SASS of the
objects the build produced (cuobjdump, no GPU): which loop is the hot one, that it is ONE basic block, what is not in it
(no CALL, no local-memory traffic, no uniform-register refills), that the population moves with Blackwell's 256-bit
accesses, and upper bounds on the instruction counts so that a regression shows up where it is introduced."""
import collections
import os
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
import sass_mix  # noqa: E402

FP64 = sass_mix.FP64
ENDS = sass_mix.ENDS


@pytest.fixture(scope="module")
def objects():
    from eqsolver_b200 import build as kbuild
    kbuild.build()   # no-op when the library is up to date; the objects sit next to it in build/
    objs = {name: os.path.join(kbuild.OBJ_DIR, name + ".o") for name in ("pso", "ce")}
    for p in objs.values():
        if not os.path.exists(p):
            pytest.skip(f"{p} missing (library built elsewhere)")
    return objs


def hot_loop(obj, pattern, min_fp64=100):
    """The body of one trip of the kernel's inner loop: of the loops with at least `min_fp64` FP64 instructions, the one made
    of the fewest basic blocks (the scalar tail and slow-path loops are full of branches), then the one with most FP64.
    Returns (body, FP64 count, opcode histogram, number of basic blocks)."""
    _, ins = sass_mix.kernel_sass(obj, pattern)
    best = None
    for s, e in sass_mix.loops(ins):
        body = [(ad, op, t) for ad, op, t in ins if s <= ad <= e]
        f = sum(1 for _, op, _ in body if op.startswith(FP64))
        blocks = 1 + sum(1 for _, op, _ in body[:-1] if op.split(".")[0] in ENDS)
        if f >= min_fp64 and (best is None or (blocks, -f) < (best[2], -best[1])):
            best = (body, f, blocks)
    assert best is not None, f"no loop with {min_fp64} FP64 instructions in {pattern}"
    body, f, blocks = best
    return body, f, collections.Counter(op.split(".")[0] for _, op, _ in body), blocks


@pytest.mark.parametrize("objective,max_instr,max_fp64", [(1, 640, 170), (3, 760, 245)])   # Rosenbrock (C2), Ackley (C4)
def test_step_kernel_trip_is_one_block_without_calls_or_local_memory(objects, objective, max_instr, max_fp64):
    body, f, hist, blocks = hot_loop(objects["pso"], rf"pso_step_kernel.*ObjectiveILi{objective}EEELb0", min_fp64=150)  # not the 1-group tail loop
    assert blocks == 1, blocks
    # one trip = one load batch of 2 groups = 8 dimensions: 6 x 256-bit loads (x, v, pbest), 4 x 256-bit stores (x, v)
    loads = [t for _, op, t in body if op.startswith("LDG")]
    stores = [t for _, op, t in body if op.startswith("STG")]
    assert len(loads) == 6 and all(".256" in t for t in loads), loads
    assert len(stores) == 4 and all(".256" in t for t in stores), stores
    for banned in ("CALL", "STL", "LDL", "R2UR", "BSSY", "BSYNC"):
        assert hist.get(banned, 0) == 0, (banned, hist)
    assert hist["IMAD"] >= 160 and hist["LOP3"] >= 160        # 8 Philox4x32-10 blocks: 160 IMAD.WIDE + 160 LOP3
    assert 500 <= len(body) <= max_instr and f <= max_fp64, (len(body), f)


@pytest.mark.parametrize("objective,max_instr,max_fp64,max_blocks", [(2, 370, 165, 1), (1, 300, 125, 2)])   # Rastrigin (C3), Rosenbrock (C5: `d > 0` in its feed)
def test_sample_kernel_trip_stays_in_registers(objects, objective, max_instr, max_fp64, max_blocks):
    body, f, hist, blocks = hot_loop(objects["ce"], rf"ce_sample_kernel.*ObjectiveILi{objective}EEELb0")
    assert blocks <= max_blocks, blocks
    # 4 coordinates = 2 Philox blocks; nothing but shared-memory reads (mu / sigma, the logarithm's tables, the cosine's
    # coefficients) besides arithmetic: no global or local memory access, no call, no uniform-register refill
    for banned in ("CALL", "STL", "LDL", "LDG", "STG", "BSSY", "BSYNC", "ATOMG", "RED"):
        assert hist.get(banned, 0) == 0, (banned, hist)
    assert hist.get("R2UR", 0) <= 6, hist                       # uniform registers suffice: no spill / refill traffic
    assert hist["IMAD"] >= 40 and hist["LOP3"] >= 40 and hist.get("LDS", 0) >= 6
    assert 250 <= len(body) <= max_instr and f <= max_fp64, (len(body), f)


def test_kernels_are_sm_100a_code(objects):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", objects["pso"]], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:500]
