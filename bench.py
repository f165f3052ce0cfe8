#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 geodesic ray pass (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full frame of the hot path: the workload BASELINE.json's metric is quoted on —
3840x2160, adaptive Cash–Karp RK, accretion disk + relativity sphere + the 99 970-triangle
lucy.obj BVH, every pixel traced (single level), default camera — through RayPipeline.pass().
Prints ONE JSON line (rank 0).  Metric: Mray-steps/s (one ray-step = one next_ray_rk call,
ray.wgsl:528); fps = 1000 / ms_per_step is reported beside it.

  value      whole-job ray-steps/s with the scene resident in HBM (device-timed, max over ranks)
  e2e        same metric through the public API with HOST buffers: per step the ModelUniform blob
             (48 MB — the reference re-sends it every frame, array_buffer.rs:71-79) goes H2D from
             pinned memory, the pass runs, and the RGBA32F frame lands in pinned host memory (zero-copy stores over PCIe
             while the kernel traces; --e2e-chunks k for banded cudaMemcpyAsync instead)
  roofline   HBM bound per BASELINE.md §4: algorithmic bytes (256 B per ray-step + ...) / kernel time
  cpu_baseline / --impl reference: the CPU restatement of the reference pass (oracle, strict libm
             flavour, OpenMP on all host cores) on a bounded sample of the same workload.  The
             reference itself (Rust + wgpu/lavapipe) cannot run here.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mray-steps/sec"
UNIT = "Mray-steps/s"
WORKLOAD = "C3: 3840x2160 single-level adaptive-RK (Cash-Karp), disk + relativity sphere + lucy.obj 99970-tri BVH, default camera (0,0,-19)"
CPU_SAMPLE_RES = (1920, 1080)     # cpu_baseline sample: same scene and camera at 1/4 area
REF_STEP_RES = (960, 540)         # --impl reference: each step is a 1/16-area frame of the same scene


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", {}


def algorithmic_bytes(stats: dict, n_px_out: int) -> int:
    """BASELINE.md §4 / SURVEY.md §8d single-level formula."""
    return (256 * stats["ray_steps"] + 16 * n_px_out + 64 * stats["node_visits"] + 124 * stats["tri_tests"]
            + 16 * stats["tex_samples"])


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_scene(need_mesh: bool = True):
    from bhusie_b200 import assets, pipelines as P
    tex, tex_src = assets.load_textures()
    if not need_mesh:
        return tex, tex_src, None, "none", {}
    if assets.have_lucy():
        blob, info = P.load_obj_model(assets.lucy_path())
        mesh_src = "lucy.obj"
    else:
        blob, info = P.model_from_arrays(*assets.uv_sphere())
        mesh_src = "synthetic uv_sphere 224x224 (lucy.obj not staged)"
    return tex, tex_src, blob, mesh_src, info


def cpu_oracle_rate(tex, blob, res, nthreads=0, repeats=1):
    """Times the oracle (strict flavour) on one frame of `res`; returns (steps/s, steps, seconds, threads)."""
    from bhusie_b200 import uniforms as U
    from oracle import oracle as O
    sc = O.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    cam, hole = U.Camera().uniform(), U.BlackHole().uniform()
    det = U.RayDetails(integration_method=1, model_count=1 if blob is not None else 0).uniform()
    threads = nthreads or O.max_threads()
    best, steps = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.ray_pass(sc, res[0], res[1], cam, hole, det, flavour="strict", nthreads=threads)
        dt = time.perf_counter() - t0
        steps = r.counters["steps"]
        best = dt if best is None else min(best, dt)
    return steps / best, steps, best, threads


def run_reference(args):
    """--impl reference: the CPU restatement of the reference ray pass on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    tex, tex_src, blob, mesh_src, _ = load_scene()
    from oracle import oracle as O
    O.build()
    threads = O.max_threads()
    for _ in range(args.warmup):
        cpu_oracle_rate(tex, blob, REF_STEP_RES, threads)
    t_total, steps_total = 0.0, 0
    for _ in range(args.steps):
        rate, steps, dt, _ = cpu_oracle_rate(tex, blob, REF_STEP_RES, threads)
        t_total += dt
        steps_total += steps
    value = steps_total / t_total / 1e6
    sample = f"{REF_STEP_RES[0]}x{REF_STEP_RES[1]} frame of the same scene/camera per step (1/16 of the 3840x2160 area)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": f"synthetic camera path; textures={tex_src}; mesh={mesh_src}",
        "config": {"workload": WORKLOAD, "reference_arm": "CPU restatement of ray.wgsl (oracle/bh_oracle.c, strict libm flavour, OpenMP); "
                   "the Rust/wgpu reference cannot be built or run here (no cargo, no Vulkan ICD)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps_extrapolated_3840x2160": None,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bhusie_b200 import pipelines as P, uniforms as U
    from bhusie_b200.multi import TiledFrame

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world != args.gpus:
        print(f"bench.py --gpus {args.gpus} must be launched with torch.distributed.run --nproc-per-node {args.gpus}", file=sys.stderr)
        return 2
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device — the ray pass has no CPU fallback (use --impl reference for the CPU arm)", file=sys.stderr)
        return 3
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H = args.width, args.height
    tex, tex_src, blob, mesh_src, mesh_info = load_scene()
    mode = {"literal": P.NUMERIC_LITERAL, "fused": P.NUMERIC_FUSED}[args.numeric_mode]
    ctx = P.Context(local_rank, numeric_mode=mode)
    ctx.set_textures(tex)
    ctx.upload_models(blob)
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=1, model_count=1)

    frame = TiledFrame(ctx, W, H, rank, world, band_rows=args.band_rows, exchange=args.exchange)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.zero_()                       # L2 flush (256 MiB > 126 MB L2), inside the timed region
        frame.render(cam, hole, det, stream)        # local bands + (N>1) NCCL gather to rank 0

    # ---------------- device-resident timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # sampled from the warm-up on: at N=8 the timed region itself is < 100 ms
    for _ in range(args.warmup):
        step_device()
    barrier()
    n_warm_samples = len(sampler.rows)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = [], []
    e0.record(stream)
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        frame.render_local(cam, hole, det, stream)
        b.record(stream)
        frame.gather(stream)
        k0.append(a); k1.append(b)
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in zip(k0, k1)]))
    if rank == 0:
        if len(sampler.rows) - n_warm_samples >= 3:
            sampler.rows = sampler.rows[n_warm_samples:]      # enough samples inside the timed region proper
            window = "timed region"
        else:
            window = "warm-up + timed region (timed region shorter than 3 sampling periods)"
        clocks = sampler.stop()
        clocks["window"] = window
    else:
        clocks = None
    stats = frame.pipeline.stats()
    t = torch.tensor([elapsed_ms, kernel_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([stats[k] for k in ("ray_steps", "node_visits", "tri_tests", "tex_samples", "px_traced")], dtype=torch.float64, device="cuda")
    s_local = s.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    elapsed_ms, kernel_ms_max = float(t[0]), float(t[1])
    total = dict(zip(("ray_steps", "node_visits", "tri_tests", "tex_samples", "px_traced"), (int(x) for x in s.tolist())))
    ms_per_step = elapsed_ms / args.steps
    value = total["ray_steps"] / (ms_per_step * 1e-3) / 1e6

    # ---------------- end-to-end timing: host buffers, H2D model blob + pass + gather + D2H frame
    pinned_model = torch.from_numpy(blob).pin_memory()
    host_frame = torch.empty((H, W, 4), dtype=torch.float32).pin_memory() if rank == 0 else None

    def step_e2e():
        flush.zero_()
        ctx.upload_models_async(pinned_model.data_ptr(), pinned_model.numel(), stream)
        if world == 1:
            # public API: pass + read-back, the D2H of each row band overlapping the tracing of the next
            frame.pipeline.pass_to_host(cam, hole, det, host_frame.data_ptr(), args.e2e_chunks, stream)
            frame.pipeline.sync()
        else:
            frame.render(cam, hole, det, stream)
            if rank == 0:
                host_frame.copy_(frame.frame_tensor(), non_blocking=True)
            frame.consumed(stream)
            stream.synchronize()

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e_value = total["ray_steps"] * args.steps / e2e_s / 1e6
    checksum = float(host_frame[::97, ::89].double().sum()) if rank == 0 else 0.0

    if rank == 0:
        peak, peak_src, peak_json = peaks()
        # roofline for the dominant kernel (trace_kernel) on THIS rank's launch
        local = dict(zip(("ray_steps", "node_visits", "tri_tests", "tex_samples", "px_traced"), (int(x) for x in s_local.tolist())))
        n_px_local = frame.pipeline.local_rows * W
        abytes = algorithmic_bytes(local, n_px_local)
        achieved = abytes / (kernel_ms * 1e-3) / 1e9
        traffic = None
        pj = {}
        # dram__bytes_read.sum + dram__bytes_write.sum, warp instructions, FMA-pipe units and register operand reads per
        # warp-step: one `ncu --set full` capture of this kernel on this workload (profiles/trace_kernel_dram.json)
        prof = os.path.join(ROOT, "profiles", "trace_kernel_dram.json")
        if os.path.exists(prof) and (W, H) == (3840, 2160) and world == 1:
            try:
                with open(prof) as f:
                    pj = json.load(f).get(args.numeric_mode, {})
                traffic = pj.get("dram_bytes_per_launch")
            except Exception:
                traffic, pj = None, {}
        inst_per_step = pj.get("warp_inst_per_warp_step")
        fp32_peak_tflops = 148 * 128 * 2 * (peak_json.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        flops = 400.0 * local["ray_steps"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "fps": 1000.0 / ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": f"synthetic (default camera/black hole uniforms); textures={tex_src}; mesh={mesh_src}",
            "config": {"workload": WORKLOAD if (W, H) == (3840, 2160) else f"{W}x{H} variant of: {WORKLOAD}", "width": W, "height": H,
                       "integrator": "cash-karp-rk", "step_size": 0.15, "max_iterations": 2000, "triangles": mesh_info.get("triangle_count"),
                       "bvh_nodes": mesh_info.get("nodes_used"), "tiling": f"cyclic bands of {frame.band_rows} rows over {world} rank(s)",
                       "exchange": {"p2p": "ray kernel stores finished pixels directly into rank 0's frame over NVLink (CUDA IPC peer memory) + one 4-byte all-reduce",
                                    "nccl": "NCCL gather of compact band buffers to rank 0 + de-interleave copy", "none": "single GPU"}[frame.exchange],
                       "l2_flush": "256 MiB memset before every step, inside the timed region",
                       "ray_steps_per_frame": total["ray_steps"],
                       "numerics": ("FUSED: explicit fma contraction + reciprocal-multiply, det-math transcendentals (bit-exact vs oracle 'fused')"
                                    if mode == P.NUMERIC_FUSED else
                                    "LITERAL: one IEEE f32 op per WGSL node, det-math transcendentals (bit-exact vs oracle 'contract')")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pinned_model.numel() + 196),
                    "d2h_bytes_per_step": int(W * H * 16), "ms_per_step": 1000.0 * e2e_s / args.steps, "fps": args.steps / e2e_s,
                    "frame_checksum": checksum,
                    "path": ("bh_ctx_upload_models_async + bh_ray_pipeline_pass_to_host("
                             + ("zero-copy: pixel stores land in the pinned host frame over PCIe during the pass" if args.e2e_chunks == 0
                                else f"{args.e2e_chunks} bands, D2H overlapped") + ") + sync"
                             if world == 1 else f"upload_models_async + tiled pass ({frame.exchange} exchange) + D2H of the assembled frame")},
            "gpu_launches": int(args.steps * 1),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "kernel": "trace_kernel<1,false,4,origin>", "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": abytes,
                         "note": "ray state is register-resident: algorithmic bytes (256 B/ray-step, SURVEY §8d) are not DRAM traffic; "
                                 "frac > 1 means the kernel beats the streamed-state HBM formulation; roofline_issue / roofline_fma_pipe / roofline_regfile are the limits it runs against"},
            "roofline_fp32": {"bound": "fp32-alu", "achieved": flops / (kernel_ms * 1e-3) / 1e12, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                              "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp32_peak_tflops, "flops_per_ray_step": 400,
                              "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (non-tensor FP32, BASELINE.md §2)"},
        }
        if inst_per_step:
            # The three limits the hot loop actually runs against, all per SM sub-partition and clock: one warp instruction
            # issued, one FMA-pipe unit (a packed FFMA2/FMUL2/FADD2 is two), two 32-bit register operands per lane read
            # (tools/ubench/fma_pipe.cu measures the last two).  Per-warp-step counts come from the committed ncu capture of
            # this kernel (profiles/trace_kernel_dram.json); time and clock are measured live.
            sm_clock_hz = ((clocks or {}).get("sm_mhz") or peak_json.get("sm_max_mhz", 1965.0)) * 1e6
            slots = 148 * 4 * sm_clock_hz
            warp_steps_per_s = (local["ray_steps"] / 32.0) / (kernel_ms * 1e-3)
            line["roofline_issue"] = {"bound": "warp-instruction issue", "achieved": inst_per_step * warp_steps_per_s / 1e9, "peak": slots / 1e9,
                                      "unit": "G warp-inst/s", "frac": inst_per_step * warp_steps_per_s / slots,
                                      "warp_inst_per_warp_ray_step": inst_per_step,
                                      "peak_source": "148 SM x 4 schedulers x SM clock sampled during the run"}
            if pj.get("fma_pipe_units_per_warp_step"):
                u = pj["fma_pipe_units_per_warp_step"]
                line["roofline_fma_pipe"] = {"bound": "FP32 FMA pipe", "achieved": u * warp_steps_per_s / 1e9, "peak": slots / 1e9,
                                             "unit": "G pipe-units/s", "frac": u * warp_steps_per_s / slots, "units_per_warp_ray_step": u,
                                             "peak_source": "1 scalar FP32 warp instruction (or half a packed one) per sub-partition per clock"}
            if pj.get("reg_operand_reads_per_warp_step"):
                r = pj["reg_operand_reads_per_warp_step"]
                line["roofline_regfile"] = {"bound": "register-file operand bandwidth", "achieved": r * warp_steps_per_s / 1e9, "peak": 2 * slots / 1e9,
                                            "unit": "G operand reads/s", "frac": r * warp_steps_per_s / (2 * slots), "reads_per_warp_ray_step": r,
                                            "peak_source": "2 x 32-bit register source operands per lane per clock per sub-partition (measured)"}
        # CPU baseline beside it (N=1 only): bounded sample of the same workload
        if world == 1 and not args.no_cpu_baseline:
            try:
                rate, steps_c, dt, threads = cpu_oracle_rate(tex, blob, CPU_SAMPLE_RES)
                line["cpu_baseline"] = {"value": rate / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{CPU_SAMPLE_RES[0]}x{CPU_SAMPLE_RES[1]} frame of the same scene/camera "
                                                  f"({steps_c} ray-steps, {dt:.1f} s); oracle strict flavour (glibc libm), OpenMP",
                                        "fps_extrapolated_3840x2160": rate / total["ray_steps"]}
            except Exception as ex:   # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        print(json.dumps(line))
    frame.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--band-rows", type=int, default=8)
    ap.add_argument("--numeric-mode", default="fused", choices=["fused", "literal"])
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="N=1 e2e read-back: 0 = zero-copy (kernel stores pixels into the pinned host frame over PCIe while tracing), "
                         "k>0 = k row bands with overlapped cudaMemcpyAsync")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = kernels store straight into rank 0's frame over NVLink (CUDA IPC); nccl = gather of band buffers")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--allow-short-warmup", action="store_true", help="profiling runs only (ncu); numbers from such runs are not bench values")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.allow_short_warmup:
        args.warmup = 3
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
