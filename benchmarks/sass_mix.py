#!/usr/bin/env python3
"""Static instruction budget of a kernel's hot loop, from the SASS of an object file (no GPU needed).

    python benchmarks/sass_mix.py build/ce.o 'ce_sample_kernel.*ObjectiveILi2EEELb0'
    python benchmarks/sass_mix.py build/pso.o 'pso_step_kernel.*ObjectiveILi3EEELb0' --loop 2

Finds the kernel whose mangled name matches the regular expression, lists its loops (backward branches, largest first)
and prints, for the chosen loop: the basic blocks it is made of (a block ends at a branch, a call or a reconvergence
instruction; rarely-taken out-of-line blocks show up as short blocks that end in CALL), the opcode histogram, the FP64
share and the issue-side cycle budget between "FP64 x 2, the rest hidden" and "everything + FP64" (DESIGN.md section 4).
This is how the per-trip figures quoted in DESIGN.md and profiles/r01_sweep.md were obtained.
"""
import argparse
import collections
import re
import subprocess
import sys

FP64 = ("DFMA", "DMUL", "DADD", "DSETP")
ENDS = ("BRA", "CALL", "BSYNC", "BSSY", "EXIT", "RET", "WARPSYNC")


def kernel_sass(obj: str, pattern: str):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    name, ins, keep = None, [], False
    rx = re.compile(pattern)
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if keep and ins:
                break
            keep = bool(rx.search(m.group(1)))
            name = m.group(1) if keep else name
            continue
        if not keep:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            text = m.group(2).strip()
            op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
            ins.append((int(m.group(1), 16), op, text))
    if not ins:
        sys.exit(f"no kernel matching {pattern!r} in {obj}")
    return name, ins


def loops(ins):
    found = []
    for addr, op, text in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)\s*$", text)
            if m and int(m.group(1), 16) <= addr:
                found.append((int(m.group(1), 16), addr))
    return sorted(set(found), key=lambda se: se[0] - se[1])  # largest span first


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel", help="regular expression on the mangled kernel name")
    ap.add_argument("--loop", type=int, default=0, help="which loop, 0 = the largest (see the list printed first)")
    a = ap.parse_args()
    name, ins = kernel_sass(a.obj, a.kernel)
    print(f"{name}: {len(ins)} instructions")
    ls = loops(ins)
    for i, (s, e) in enumerate(ls[:8]):
        print(f"  loop {i}: 0x{s:04x} .. 0x{e:04x}  ({(e - s) // 16 + 1} instructions)")
    if not ls:
        return
    s, e = ls[min(a.loop, len(ls) - 1)]
    body = [(ad, op, t) for ad, op, t in ins if s <= ad <= e]
    blocks, cur = [], []
    for ad, op, t in body:
        cur.append((ad, op, t))
        if op.split(".")[0] in ENDS:
            blocks.append(cur)
            cur = []
    if cur:
        blocks.append(cur)
    print(f"loop 0x{s:04x} .. 0x{e:04x}: {len(body)} instructions in {len(blocks)} blocks")
    for b in blocks:
        f = sum(1 for _, op, _ in b if op.startswith(FP64))
        print(f"  0x{b[0][0]:04x} .. 0x{b[-1][0]:04x}: {len(b):4d} instructions, {f:3d} FP64, ends in {b[-1][2][:40]}")
    hist = collections.Counter(op.split(".")[0] for _, op, _ in body)
    f = sum(v for k, v in hist.items() if k in FP64)
    print("  " + "  ".join(f"{k} {v}" for k, v in hist.most_common(16)))
    print(f"  FP64 {f} of {len(body)} ({100.0 * f / len(body):.0f} %); issue-side cycles per trip and scheduler between {2 * f} "
          f"(FP64 pipe alone) and {len(body) + f} (every other instruction a full slot), out-of-line blocks included")


if __name__ == "__main__":
    main()
