#!/usr/bin/env python
"""Turns the captures of tools/capture_constants.sh into profiles/trace_kernel_dram.json — the per-warp-step constants of the
hot kernel that bench.py's roofline block uses (DRAM bytes per launch, warp instructions / FMA-pipe units / register operand
reads per warp ray-step), stamped with the hash of the kernel source they were measured on.

    python tools/update_constants.py gpurun_out/r2_const [--steps N] [--round r2_NN]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAY_STEPS_HEADLINE = 1945737757          # ray-steps of one 3840x2160 headline frame (oracle counter == device counter, tests)


def main():
    prefix = sys.argv[1]
    steps = RAY_STEPS_HEADLINE
    tag = os.path.basename(prefix)
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    if "--round" in sys.argv:
        tag = sys.argv[sys.argv.index("--round") + 1]
    srchash = open(prefix + ".srchash").read().strip()
    out_path = os.path.join(ROOT, "profiles", "trace_kernel_dram.json")
    out = {}
    for mode in ("fused", "literal"):
        rep = f"{prefix}_{mode}.ncu-rep"
        if not os.path.exists(rep):
            continue
        summary_prefix = os.path.join(ROOT, "profiles", f"{tag}_{mode}_rk_4k")
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, str(steps), summary_prefix], check=True,
                       stdout=subprocess.DEVNULL)
        s = json.load(open(summary_prefix + ".json"))
        out[mode] = {"source_hash": srchash, "dram_bytes_per_launch": s["dram_bytes_per_launch"],
                     "warp_inst_per_warp_step": round(s["warp_inst_per_warp_step"], 2),
                     "fma_pipe_units_per_warp_step": s["fma_pipe_units_per_warp_step"],
                     "reg_operand_reads_per_warp_step": s["reg_operand_reads_per_warp_step"],
                     "issue_active_pct": s["issue_active_pct"], "threads_per_inst": s["threads_per_inst"],
                     "kernel": s["kernel"], "duration_ms_under_ncu": s["duration_ms"],
                     "source": f"profiles/{tag}_{mode}_rk_4k.json (ncu --set full on `bench.py --steps 1 --warmup 1 --numeric-mode {mode}`)"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
