#!/usr/bin/env python
"""Static view of the hot loop without a GPU: disassembles libbhray.so (cuobjdump), takes the largest straight-line run of
FP32 arithmetic of a trace kernel — the quiet step of hot_iteration — and prints the three per-step costs the kernel runs
against: warp instructions, FMA-pipe units (a packed FFMA2/FMUL2/FADD2 is two), register operand reads
(two per clock).  The executed-count version of the same numbers comes from tools/ncu_summary.py on a real capture.

    python tools/sass_hot_loop.py [kernel-name-substring] [library] [--list]
      default kernel: fus12trace_kernelILi1ELb0ELi4ELb1E  (FUSED, Cash-Karp, tile mode, 4 CTAs/SM, hole at origin)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import reg_source_reads                    # noqa: E402

FP = ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD")
BREAKS = ("BRA", "CALL", "RET", "EXIT", "BSYNC", "WARPSYNC", "VOTE", "VOTEU")
TAIL_MARKS = ("STL", "LDL")          # argument traffic of the out-of-line literal tail: not part of the quiet step


def kernel_sass(lib, needle):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
    start = next(i for i, l in enumerate(out) if "Function :" in l and needle in l)
    end = next((i for i in range(start + 1, len(out)) if "Function :" in out[i]), len(out))
    ins = []
    for l in out[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    return out[start].split("Function :")[1].strip(), ins


def opcode(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    needle = args[0] if len(args) > 0 else "fus12trace_kernelILi1ELb0ELi4ELb1E"
    lib = args[1] if len(args) > 1 else os.path.join(ROOT, "bhusie_b200", "lib", "libbhray.so")
    name, ins = kernel_sass(lib, needle)
    # straight-line runs: split at unpredicated control flow; predicated branches (the never-taken ones) stay inside
    runs, cur = [], []
    for addr, text in ins:
        if opcode(text) in TAIL_MARKS:
            runs.append(cur)
            cur = []
            continue
        cur.append((addr, text))
        if opcode(text) in BREAKS and not text.startswith("@"):
            runs.append(cur)
            cur = []
    runs.append(cur)
    best = max(runs, key=lambda r: sum(opcode(t) in FP for _, t in r))
    # the run may start in the loop's pre-header (the register moves that set up the first iteration): a backward branch that
    # lands inside it marks where the loop body — what is executed per step — begins
    lo, hi = int(best[0][0], 16), int(best[-1][0], 16)
    heads = []
    for addr, text in ins:
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", text)
        if m and lo < int(m.group(1), 16) <= hi and int(addr, 16) > hi:
            heads.append(int(m.group(1), 16))
    preheader = 0
    if heads:
        head = max(heads)
        preheader = sum(1 for a, _ in best if int(a, 16) < head)
        best = [(a, t) for a, t in best if int(a, 16) >= head]
    ops = collections.Counter(opcode(t) for _, t in best)
    units = sum(c * (2 if op in ("FFMA2", "FMUL2", "FADD2") else 1) for op, c in ops.items() if op in FP + ("IMAD", "HFMA2"))
    reads, prev = 0, {}
    for _, t in best:
        n, prev = reg_source_reads(t, prev)
        reads += n
    print(name)
    print(f"quiet step: {best[0][0]}..{best[-1][0]}  {len(best)} instructions, {units} FMA-pipe units, {reads} register operand reads "
          f"({reads / 2:.0f} cycles at 2 per clock)")
    print("  " + "  ".join(f"{op} {c}" for op, c in ops.most_common()))
    if preheader:
        print(f"  ({preheader} instructions before the loop head, executed once per hot phase, are not counted)")
    if "--list" in sys.argv:
        for addr, text in best:
            print(f"    /*{addr}*/  {text}")


if __name__ == "__main__":
    main()
