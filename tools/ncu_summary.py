#!/usr/bin/env python
"""Summarises an .ncu-rep (raw + source pages) into a small JSON + text: duration, DRAM bytes, issue
utilisation, pipe utilisation, stall reasons, opcode mix per ray-step, hot-loop footprint.
usage: ncu_summary.py REPORT.ncu-rep RAY_STEPS [OUT_PREFIX]"""
import collections
import csv
import io
import json
import subprocess
import sys


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


_NO_DEST = ("ST", "STS", "STL", "STG", "RED", "BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "NOP", "YIELD", "BAR",
            "MEMBAR", "DEPBAR", "BREAK", "R2UR", "UBLKCP", "SYNCS", "ERRBAR", "CCTL", "FENCE")
_WIDE_SRC = ("DMUL", "DFMA", "DADD", "DSETP", "DMNMX")


def reg_source_reads(text, reuse_prev):
    """(number of 32-bit register source operands this SASS line reads from the register file, reuse state for the next line)"""
    import re
    t = text.strip().rstrip(";").strip()
    if t.startswith("@"):
        t = t.split(None, 1)[1] if " " in t else ""
    if not t:
        return 0, {}
    parts = t.split(None, 1)
    op = parts[0]
    base = op.split(".")[0]
    toks = [x.strip() for x in parts[1].split(",")] if len(parts) > 1 else []
    toks = [x for x in toks if not re.match(r"^!?U?P[T0-9]$", x)]           # predicates are not register-file operands
    if base not in _NO_DEST and toks:
        toks = toks[1:]                                                       # destination
    wide_all = base in _WIDE_SRC or op.startswith("F2F.F32.F64")
    n, now, seen = 0, {}, set()
    for slot, x in enumerate(toks):
        for m in re.finditer(r"(?<![U\w])R(\d+)((?:\.\w+)*)", x):
            idx, suf = int(m.group(1)), m.group(2)
            width = 2 if (".F32x2" in suf or ".64" in suf or wide_all) else 1
            flagged = ".reuse" in suf
            regs = tuple(range(idx, idx + width))
            if flagged:
                now[slot] = regs
            if reuse_prev.get(slot) == regs:
                continue                                                      # served by the operand reuse cache
            for g in regs:
                if g not in seen:
                    seen.add(g)
                    n += 1
    return n, now


def main():
    rep, steps = sys.argv[1], float(sys.argv[2])
    prefix = sys.argv[3] if len(sys.argv) > 3 else None
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}

    def g(k):
        try:
            return float(m[k][0].replace(",", ""))
        except Exception:
            return None
    unit = lambda k: m.get(k, ("", ""))[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    dur = g("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[unit("gpu__time_duration.sum")]
    rd = g("dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1.0)
    wr = g("dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1.0)
    s = {
        "kernel": m.get("Kernel Name", ("?",))[0], "duration_ms": dur * 1e3,
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
        "dram_throughput_pct": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": g("launch__registers_per_thread"), "grid": g("launch__grid_size"), "block": g("launch__block_size"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "sm_throughput_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "pipe_tensor_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "warp_inst_executed": g("smsp__inst_executed.sum"),
        "threads_per_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "sm_clock_ghz": g("sm__cycles_elapsed.max.per_second"),
        "ray_steps": steps,
    }
    s["warp_inst_per_warp_step"] = s["warp_inst_executed"] / (steps / 32.0)
    s["gsteps_per_s_under_ncu"] = steps / dur / 1e9
    stalls = {}
    for k in m:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = g(k)
    s["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -(kv[1] or 0))[:8])

    src = page(rep, "source")
    h2 = src[1]
    ia, isrc, ismp = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples")
    ops, smp = collections.Counter(), collections.Counter()
    mx = 0
    rows = [r for r in src[2:] if len(r) > ia and r[ia].isdigit()]
    for r in rows:
        t = r[isrc].strip().split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += int(r[ia]); smp[op] += int(r[ismp]); mx = max(mx, int(r[ia]))
    # Register-file operand traffic: 32-bit register source operands read per warp instruction, weighted by execution
    # count.  tools/ubench/fma_pipe.cu measures the SM sub-partition at 2 such operands per lane per clock (FFMA with
    # three distinct registers: 1.53 clk; FFMA2 with three distinct pairs: 3.03 clk), so reads / 2 is a cycle bound.
    # `.reuse` hits (same register, same slot, in the instruction right after one that flagged it) are not counted.
    reads = 0.0
    reuse_prev = {}
    for r in rows:
        n, reuse_prev = reg_source_reads(r[isrc], reuse_prev)
        reads += n * int(r[ia])
    s["reg_operand_reads_per_warp_step"] = round(reads / (steps / 32.0), 1)
    s["reg_operand_cycles_per_warp_step"] = round(reads / (steps / 32.0) / 2.0, 1)
    s["cycles_per_warp_step_measured"] = round(dur * (s["sm_clock_ghz"] or 0) * 1e9 * (148 * 4) / (steps / 32.0), 1)
    # FMA-pipe occupancy in scalar-instruction units: a packed FP32 instruction holds both 16-lane sub-pipes for two
    # cycles, a scalar one holds one of them for two (tools/ubench/fma_pipe.cu), so the pipe does one unit per clock.
    units = sum(c * (2 if op in ("FFMA2", "FMUL2", "FADD2") else 1) for op, c in ops.items()
                if op in ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "IMAD", "HFMA2"))
    s["fma_pipe_units_per_warp_step"] = round(units / (steps / 32.0), 1)
    s["opcode_per_warp_step"] = {op: round(c / (steps / 32.0), 2) for op, c in ops.most_common(18)}
    s["stall_samples_by_opcode"] = dict(smp.most_common(10))
    s["hot_loop_sass_lines"] = sum(1 for r in rows if int(r[ia]) > 0.5 * mx)
    s["hot_loop_bytes"] = 16 * s["hot_loop_sass_lines"]
    s["sass_lines_total"] = len(rows)
    txt = json.dumps(s, indent=1)
    print(txt)
    if prefix:
        open(prefix + ".json", "w").write(txt + "\n")


if __name__ == "__main__":
    main()
