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
