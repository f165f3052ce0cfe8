#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 geodesic ray pass (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full frame of the hot path: the workload BASELINE.json's metric is quoted on —
3840x2160, adaptive Cash–Karp RK, accretion disk + relativity sphere + the 99 970-triangle
lucy.obj BVH, every pixel traced (single level), default camera — through RayPipeline.pass().
Prints ONE JSON line (rank 0).  Metric: Mray-steps/s (one ray-step = one next_ray_rk call,
ray.wgsl:528); fps = 1000 / ms_per_step is reported beside it.

  value          whole-job ray-steps/s with the scene resident in HBM (device-timed, max over ranks), FUSED numeric mode
  value_literal  the same measurement in LITERAL numeric mode (one IEEE op per WGSL node), ms_per_step_literal beside it
  e2e            same metric through the public API with HOST buffers: per step the ModelUniform blob
                 (48 MB — the reference re-sends it every frame, array_buffer.rs:71-79) goes H2D from
                 pinned memory, the pass runs, and the RGBA32F frame lands in page-locked host memory: the kernel's pixel
                 stores go there directly over PCIe while it traces.  N > 1: ONE host frame in POSIX shared memory that
                 every rank maps, so each rank uses its own PCIe link.  e2e_header_only: the variant that sends the 16-byte
                 model header instead of the blob (bh_ctx_set_model_header)
  roofline       the limit the hot kernel actually runs against — warp-instruction issue — from the committed ncu
                 capture of THIS kernel source (profiles/trace_kernel_dram.json, refused when its source hash is stale);
                 achieved_dram_gbs = measured DRAM traffic / kernel time; roofline_hbm_streamed = the SURVEY §8d
                 accounting (256 B per ray-step as if the state were streamed through HBM)
  parity         (N = 1) 48 rows of the timed 4K frame against the oracle: bit-exact vs the mode's flavour, outlier
                 fraction / max-abs vs the strict libm flavour, share of outliers the float64 shadow marks ill-conditioned
  pyramid        (N = 1) the reference's own frame: 4-level adaptive grid 72x41 -> 1918x1081 + sky (+ post chain), Euler and RK
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
STAT_KEYS = ("ray_steps", "node_visits", "tri_tests", "tex_samples", "px_traced")


def host_threads() -> int:
    """Cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not
    inherit that."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


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
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")][1:])

    def stop(self, t0: float | None = None, t1: float | None = None) -> dict:
        """Median SM clock and throttle reasons of the samples taken in [t0, t1] (wall clock); of all samples if fewer than 3 are."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows, window = self.rows, "whole run (warm-up + timed regions)"
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[0] <= t1]
            if len(inside) >= 3:
                rows, window = inside, "timed region"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def load_scene():
    """Textures + mesh through the PRODUCT's host code (libbhray.so's OBJ loader / BVH builder)."""
    from bhusie_b200 import assets, pipelines as P
    tex, tex_src = assets.load_textures()
    if assets.have_lucy():
        blob, info = P.load_obj_model(assets.lucy_path())
        mesh_src = "lucy.obj"
    else:
        blob, info = P.model_from_arrays(*assets.uv_sphere())
        mesh_src = "synthetic uv_sphere 224x224 (lucy.obj not staged)"
    return tex, tex_src, blob, mesh_src, info


def load_scene_oracle():
    """The same scene built with the ORACLE's own OBJ loader / BVH builder: the reference arm never maps libbhray.so."""
    from bhusie_b200 import assets                  # pure Python (PNG decode, synthetic fallbacks): loads no native library
    from oracle import oracle as O
    tex, tex_src = assets.load_textures()
    if assets.have_lucy():
        blob, _ = O.load_obj(assets.lucy_path())
        mesh_src = "lucy.obj"
    else:
        pts, nrm, tris = assets.uv_sphere()
        blob = O.new_model_blob()
        v = O.blob_views(blob)
        v["points"][: len(pts), :3] = pts
        v["normals"][: len(nrm), :3] = nrm
        v["triangles"][: len(tris)] = tris
        v["position"][:] = (-10.0, 0.0, 30.0)
        v["visible"][0] = 1
        hdr = blob[:48].view(np.int32)
        hdr[8], hdr[10] = len(pts), len(tris)
        O.build_bvh(blob, len(tris))
        mesh_src = "synthetic uv_sphere 224x224 (lucy.obj not staged)"
    return tex, tex_src, blob, mesh_src


def cpu_oracle_rate(tex, blob, res, nthreads, repeats=1):
    """Times the oracle (strict flavour) on one frame of `res`; returns (steps/s, steps, seconds, threads)."""
    from bhusie_b200 import uniforms as U
    from oracle import oracle as O
    sc = O.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    cam, hole = U.Camera().uniform(), U.BlackHole().uniform()
    det = U.RayDetails(integration_method=1, model_count=1 if blob is not None else 0).uniform()
    best, steps = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.ray_pass(sc, res[0], res[1], cam, hole, det, flavour="strict", nthreads=nthreads)
        dt = time.perf_counter() - t0
        steps = r.counters["steps"]
        best = dt if best is None else min(best, dt)
    return steps / best, steps, best, nthreads


def run_reference(args):
    """--impl reference: the CPU restatement of the reference ray pass on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)          # before libgomp initialises (the oracle library is not loaded yet)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import oracle as O
    O.build()
    tex, tex_src, blob, mesh_src = load_scene_oracle()
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
                   "the Rust/wgpu reference cannot be built or run here (no cargo, no Vulkan ICD)", "sample": sample,
                   "mesh_loader": "oracle's own bho_load_obj / bho_build_bvh (libbhray.so is not loaded by this arm)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "fps_extrapolated_3840x2160": None,
    }
    print(json.dumps(line))
    return 0


def kernel_constants(numeric_mode: str):
    """Per-warp-step counts of the hot kernel from the committed ncu capture — only when it was taken from THIS source."""
    from bhusie_b200 import build as B
    prof = os.path.join(ROOT, "profiles", "trace_kernel_dram.json")
    try:
        with open(prof) as f:
            pj = json.load(f).get(numeric_mode, {})
    except Exception:
        return {}, "profiles/trace_kernel_dram.json unreadable"
    have, want = pj.get("source_hash"), B.kernel_source_hash()
    if have != want:
        return {}, f"stale: capture is of kernel source {have}, this build is {want} (re-run tools/capture_constants.sh)"
    return pj, "current"


def parity_block(ctx, P, U, tex, blob, W, H, cam, hole, det, n_rows=48):
    """48 rows of the headline frame, device vs oracle (the oracle is the checker here, never the thing measured)."""
    from oracle import oracle as O
    osc = O.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    rows = sorted(set(np.linspace(0, H - 1, n_rows - 8).astype(int).tolist() + [H // 2 + d for d in (-60, -25, -8, -1, 0, 7, 24, 59)]))
    camb, holeb, detb = cam.uniform(), hole.uniform(), det.uniform()
    neutral = {}
    for fl, pert in (("strict", 0.0), ("shadow", 0.0), ("probe", O.SHADOW_PERTURBATION)):
        neutral[fl] = np.stack([O.ray_pass(osc, W, H, camb, holeb, detb, rows=(y, y + 1), flavour="strict" if fl == "strict" else "shadow",
                                           perturb=pert).rgba[y] for y in rows])
    out = {"rows": len(rows), "pixels": len(rows) * W, "tolerance": 1e-4,
           "how": "rows of the timed 3840x2160 frame; oracle flavours computed on the host for the same rows; ill-conditioned = the strict f32 "
                  "evaluation is off its float64 shadow by > tol, or a 1e-6 rad rotation of the camera ray moves the float64 result by > tol; "
                  "whole-frame figures for every BASELINE config: profiles/r2_parity_report.json"}
    mode0 = ctx.numeric_mode
    for mode, name in ((P.NUMERIC_FUSED, "fused"), (P.NUMERIC_LITERAL, "literal")):
        ctx.set_numeric_mode(mode)
        rp = P.RayPipeline(ctx, W, H, aux=P.AUX_HIT | P.AUX_STEPS)
        rp.pass_(cam, hole, det)
        dev = rp.read()
        rp.close()
        fl = P.ORACLE_FLAVOUR_OF_MODE[mode]
        exact = True
        for y in rows:
            o = O.ray_pass(osc, W, H, camb, holeb, detb, rows=(y, y + 1), flavour=fl)
            exact = exact and bool(np.array_equal(dev["rgba"][y].view(np.uint32), o.rgba[y].view(np.uint32))
                                   and np.array_equal(dev["hit"][y], o.hit[y]) and np.array_equal(dev["steps"][y], o.steps[y]))
        rep = O.parity_report(dev["rgba"][rows], neutral["strict"], neutral["shadow"], 1e-4, neutral["probe"])
        out[name] = {"bit_exact_vs_oracle_flavour": exact, "oracle_flavour": fl, "vs_strict_outlier_frac": rep["outlier_frac"],
                     "vs_strict_max_abs": rep["max_abs"], "vs_strict_rms": rep["rms"],
                     "outliers_ill_conditioned_share": rep["outliers_ill_conditioned_share"],
                     "outlier_frac_well_conditioned": rep["outlier_frac_well_conditioned"]}
    ctx.set_numeric_mode(mode0)
    return out


def pyramid_block(ctx, P, U, torch, stream, reps=20):
    """The reference's own frame (mod.rs:170-216, 406-431): 4 levels 72x41 -> 1918x1081, sky resolve, post chain."""
    from bhusie_b200.post import PostChain
    cam, hole = U.Camera(), U.BlackHole()
    out = {"what": "4-level adaptive grid 72x41->1918x1081 (x3-2) + Rgba16Float sky resolve; post = bloom x10 + mix + ACES + FXAA; "
                   "device-timed, mean of %d frames; 4k = base 143x81 -> 3835x2161" % reps}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for name, base in (("reference_1918x1081", (72, 41)), ("4k_3835x2161", (143, 81))):
        for method, mname in ((0, "euler"), (1, "rk")):
            det = U.RayDetails(integration_method=method, model_count=1)
            pyr = P.RayPyramid(ctx, base=base)
            chain = PostChain(ctx, pyr.sky)
            ms_ray = timed(lambda: pyr.pass_(cam, hole, det, stream))
            ms_all = timed(lambda: (pyr.pass_(cam, hole, det, stream), chain.pass_(stream)))
            steps = sum(rp.stats()["ray_steps"] for rp in pyr.levels)
            last = pyr.levels[-1].stats()
            out[f"{name}_{mname}"] = {"ms_ray_levels_plus_sky": ms_ray, "ms_frame_with_post_chain": ms_all, "fps": 1000.0 / ms_all,
                                      "ray_steps": steps, "gsteps_per_s": steps / ms_ray / 1e6,
                                      "last_level_px_traced": last["px_traced"], "last_level_px_interp": last["px_interp"]}
            # the same frame with TWO frames in flight (a renderer's usual double buffering: its own pyramid, post chain and stream per
            # frame slot): frame k+1's latency-bound coarse levels run under frame k's last level.  Throughput, not latency.
            pyr_b = P.RayPyramid(ctx, base=base)
            chain_b = PostChain(ctx, pyr_b.sky)
            stream_b = torch.cuda.Stream()
            slots = ((pyr, chain, stream), (pyr_b, chain_b, stream_b))

            def frames(n):
                for k in range(n):
                    py, ch, st = slots[k & 1]
                    py.pass_(cam, hole, det, st)
                    ch.pass_(st)
            frames(4)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            stream_b.wait_event(e0)
            frames(reps)
            stream.wait_stream(stream_b)
            e1.record(stream)
            torch.cuda.synchronize()
            out[f"{name}_{mname}"]["ms_frame_with_post_chain_two_in_flight"] = e0.elapsed_time(e1) / reps
            chain_b.close(); pyr_b.close()
            chain.close(); pyr.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bhusie_b200 import build as B, pipelines as P, uniforms as U
    from bhusie_b200.multi import HostTiledFrame, TiledFrame

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
        host_group = dist.new_group(backend="gloo")       # host-side barrier: ranks can wait without a kernel spinning on their GPU

    W, H = args.width, args.height
    tex, tex_src, blob, mesh_src, mesh_info = load_scene()
    mode = {"literal": P.NUMERIC_LITERAL, "fused": P.NUMERIC_FUSED}[args.numeric_mode]
    ctx = P.Context(local_rank, numeric_mode=mode)
    ctx.set_textures(tex)
    ctx.upload_models(blob)
    cam, hole = U.Camera(), U.BlackHole()
    det = U.RayDetails(integration_method=1, model_count=1)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                 # nvidia-smi takes a few hundred ms to deliver its first sample: started well before the timed region
    frame = TiledFrame(ctx, W, H, rank, world, band_rows=args.band_rows, exchange=args.exchange)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(fr=frame):
        flush.zero_()                       # L2 flush (256 MiB > 126 MB L2)
        fr.render(cam, hole, det, stream)   # local bands; N > 1: peer stores into rank 0's frame + flag wait (or NCCL gather)
        fr.consumed(stream)

    def timed_steps(fr, steps, warmup):
        """(ms summed over `steps` step intervals, the same including the L2 flushes between them, mean kernel ms of this rank) —
        device events on the launching stream.  A step's interval opens after the flush that precedes it (the flush is the
        timing hygiene BETWEEN iterations, not part of the pass) and closes after the frame is complete on rank 0."""
        for _ in range(warmup):
            step_device(fr)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1, s0, s1 = [], [], [], []
        e0.record(stream)
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sa, sb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sa.record(stream)
            fr.render_local(cam, hole, det, stream, events=(a, b))      # events around the ray kernel proper (after any flag wait)
            fr.gather(stream)
            fr.resolve_sky(stream)
            sb.record(stream)
            fr.consumed(stream)
            k0.append(a); k1.append(b); s0.append(sa); s1.append(sb)
        e1.record(stream)
        barrier()
        ctx.check_async()
        kernel = float(np.mean([a.elapsed_time(b) for a, b in zip(k0, k1)])) if k0 else 0.0
        return float(sum(a.elapsed_time(b) for a, b in zip(s0, s1))), e0.elapsed_time(e1), kernel

    # ---------------- N > 1: the assembled frame must be bit-identical to a single-GPU render (in the warm-up, not timed)
    exchange_bit_identical = None
    if world > 1:
        step_device()
        barrier()
        if rank == 0:
            single = P.RayPipeline(ctx, W, H)
            single.pass_(cam, hole, det, stream)
            ref = torch.as_tensor(_DevPtr(single.output_ptr, (H, W, 4)), device="cuda")
            torch.cuda.synchronize()
            exchange_bit_identical = bool(torch.equal(frame.frame_tensor().view(torch.int32), ref.view(torch.int32)))
            del ref
            single.close()
        barrier()

    # ---------------- device-resident timing (FUSED unless --numeric-mode literal)
    for _ in range(args.warmup):
        step_device()
    barrier()
    t_wall0 = time.time()
    elapsed_ms, elapsed_flush_ms, kernel_ms = timed_steps(frame, args.steps, 0)
    t_wall1 = time.time()
    stats = frame.pipeline.stats()
    t = torch.tensor([elapsed_ms, kernel_ms, elapsed_flush_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([stats[k] for k in STAT_KEYS], dtype=torch.float64, device="cuda")
    s_local = s.clone()
    per_rank = None
    if world > 1:
        # every rank's own share: ray-steps, mean kernel time, elapsed time
        mine = torch.tensor([float(stats["ray_steps"]), kernel_ms, elapsed_ms], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ray_steps": [int(x[0]) for x in allr], "kernel_ms": [float(x[1]) for x in allr], "elapsed_ms": [float(x[2]) for x in allr]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    elapsed_ms, elapsed_flush_ms = float(t[0]), float(t[2])
    total = dict(zip(STAT_KEYS, (int(x) for x in s.tolist())))
    ms_per_step = elapsed_ms / args.steps
    value = total["ray_steps"] / (ms_per_step * 1e-3) / 1e6

    # ---------------- the other numeric mode, same loop (LITERAL: one IEEE op per WGSL node; the mode whose outliers vs libm are 10x rarer)
    other = P.NUMERIC_LITERAL if mode == P.NUMERIC_FUSED else P.NUMERIC_FUSED
    ctx.set_numeric_mode(other)
    lit_steps = max(3, args.steps // 2)
    lit_ms, _, _ = timed_steps(frame, lit_steps, 3)
    lt = torch.tensor([lit_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(lt, op=dist.ReduceOp.MAX)
    ms_other = float(lt[0]) / lit_steps
    ctx.set_numeric_mode(mode)
    other_name = "literal" if other == P.NUMERIC_LITERAL else "fused"
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---------------- end-to-end timing: host buffers, H2D model blob + pass with the pixels stored straight into the host frame.
    # TWO frames in flight (the renderer's usual double buffering): frame k+1's inputs go up while frame k is traced, each frame on
    # its own context (model buffer), stream and host frame; a frame's buffers are reused only after it was finished and read.
    # `e2e_serial` is the same loop with ONE frame in flight (upload, pass, wait — what round 1 reported as e2e).
    pinned_model = torch.from_numpy(blob).pin_memory()
    ctx_b = P.Context(local_rank, numeric_mode=mode)
    ctx_b.set_textures(tex)
    ctx_b.upload_models(blob)
    stream_b = torch.cuda.Stream()
    lanes = []                                            # [context, stream, host frame (N=1) | HostTiledFrame (N>1), ray pipeline]
    e2e_error = None
    try:
        for i, (c, st) in enumerate(((ctx, stream), (ctx_b, stream_b))):
            if world == 1:
                hf = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
                rp = frame.pipeline if i == 0 else P.RayPipeline(c, W, H)
                lanes.append([c, st, hf, rp])
            else:
                hf = HostTiledFrame(c, W, H, rank, world, band_rows=args.band_rows,
                                    name=f"/bhframe_{os.environ.get('MASTER_PORT', '0')}_{W}x{H}_{i}")
                lanes.append([c, st, hf, hf.pipeline])
    except RuntimeError as ex:        # every rank gets the same verdict (HostTiledFrame agrees on it collectively): no host frame, no e2e
        e2e_error = str(ex)[:400]
    probe = [0.0]

    def e2e_enqueue(lane, header_only):
        c, st, hf, rp = lane
        with torch.cuda.stream(st):
            flush.zero_()
        if header_only:
            c.set_model_header(0, (-10.0, 0.0, 30.0), 1)          # what the UI can change per frame (ui/model_settings.rs:39-48): 16 bytes
        else:
            c.upload_models_async(pinned_model.data_ptr(), pinned_model.numel(), st)
        if world == 1:
            rp.pass_to_host(cam, hole, det, hf.data_ptr(), args.e2e_chunks, st)
        else:
            hf.enqueue(cam, hole, det, st)                        # own bands -> the shared host frame over this rank's own PCIe link

    def e2e_finish(lane):
        c, st, hf, rp = lane
        if world == 1:
            rp.sync()
            probe[0] += float(hf[H // 2, W // 2, 0])              # the step's result, read on the host
        else:
            hf.finish()
            if rank == 0:
                probe[0] += float(hf.frame_array()[H // 2, W // 2, 0])
            hf.consumed()

    def timed_e2e(header_only, in_flight):
        def loop(n):
            pending = []
            for k in range(n):
                lane = lanes[k % in_flight]
                e2e_enqueue(lane, header_only)
                pending.append(lane)
                if len(pending) == in_flight:
                    e2e_finish(pending.pop(0))
            while pending:
                e2e_finish(pending.pop(0))
        loop(max(2, min(args.warmup, 4)))
        barrier()
        t0 = time.perf_counter()
        loop(args.steps)
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te[0])

    e2e_serial_s = e2e_hdr_s = e2e_s = e2e_value = checksum = e2e_lanes_equal = e2e_matches_device = None
    if e2e_error is None:
        e2e_serial_s = timed_e2e(False, 1)
        e2e_hdr_s = timed_e2e(True, 1)
        e2e_s = timed_e2e(False, 2)
        e2e_value = total["ray_steps"] * args.steps / e2e_s / 1e6
        if world > 1:
            step_device()                   # the device frame in the headline numeric mode again, to compare the host frame with
            barrier()
        if rank == 0:
            hf = lanes[0][2].numpy() if world == 1 else lanes[0][2].frame_array()
            hf_b = lanes[1][2].numpy() if world == 1 else lanes[1][2].frame_array()
            checksum = float(hf[::97, ::89].astype(np.float64).sum())
            e2e_lanes_equal = bool(np.array_equal(np.ascontiguousarray(hf).view(np.uint32), np.ascontiguousarray(hf_b).view(np.uint32)))
            if world > 1:
                dev_frame = frame.frame_tensor().cpu().numpy()
                e2e_matches_device = bool(np.array_equal(dev_frame.view(np.uint32), np.ascontiguousarray(hf).view(np.uint32)))
                del dev_frame
    barrier()

    # ---------------- C4 (BASELINE configs[3]): 7680x4320 over the same ranks
    c4 = None
    if world > 1 and (args.c4 == "on" or (args.c4 == "auto" and world == 8)):
        f8 = TiledFrame(ctx, 7680, 4320, rank, world, band_rows=args.band_rows, exchange=args.exchange)
        c4_steps = max(3, min(args.steps, 8))
        c4_ms, _, _ = timed_steps(f8, c4_steps, 3)
        st8 = f8.pipeline.stats()
        v8 = torch.tensor([c4_ms, float(st8["ray_steps"])], dtype=torch.float64, device="cuda")
        vmax = v8.clone()
        if world > 1:
            dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(v8, op=dist.ReduceOp.SUM)
        c4 = {"workload": "C4: 7680x4320 single-level adaptive-RK, full scene, cyclic row bands over %d GPUs" % world,
              "ms_per_step": float(vmax[0]) / c4_steps, "fps": 1000.0 * c4_steps / float(vmax[0]), "ray_steps_per_frame": int(v8[1]),
              "value": float(v8[1]) / (float(vmax[0]) / c4_steps * 1e-3) / 1e6, "unit": UNIT, "steps": c4_steps}
        f8.close()
        barrier()

    # ---------------- bh_frame_multi (one process, one host thread, all N devices) checked by rank 0 while the others idle
    frame_multi = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0 and not args.no_frame_multi:
            try:
                frame_multi = frame_multi_block(P, U, torch, ctx, tex, blob, world, W, H, args.band_rows, cam, hole, det)
            except Exception as ex:               # a diagnostic, not the measurement: never lose the line over it
                frame_multi = {"error": str(ex)[:300]}
        dist.barrier(group=host_group)            # the other ranks wait on the host: their GPUs are idle for rank 0's contexts
    barrier()

    if rank == 0:
        peak, peak_src, peak_json = peaks()
        # roofline for the dominant kernel (trace_kernel) on THIS rank's launch
        local = dict(zip(STAT_KEYS, (int(x) for x in s_local.tolist())))
        n_px_local = frame.pipeline.local_rows * W
        abytes = algorithmic_bytes(local, n_px_local)
        hbm_achieved = abytes / (kernel_ms * 1e-3) / 1e9
        pj, pj_state = kernel_constants(args.numeric_mode) if (W, H) == (3840, 2160) else ({}, "constants are for 3840x2160")
        traffic = pj.get("dram_bytes_per_launch") if world == 1 else None
        inst_per_step = pj.get("warp_inst_per_warp_step")
        fp32_peak_tflops = 148 * 128 * 2 * (peak_json.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
        flops = 400.0 * local["ray_steps"]
        sm_clock_hz = ((clocks or {}).get("sm_mhz") or peak_json.get("sm_max_mhz", 1965.0)) * 1e6
        slots = 148 * 4 * sm_clock_hz
        warp_steps_per_s = (local["ray_steps"] / 32.0) / (kernel_ms * 1e-3)
        hbm_block = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "traffic": traffic,
                     "achieved_dram_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None,
                     "peak_source": peak_src, "kernel": "trace_kernel<1,false,4,origin>", "kernel_ms": kernel_ms,
                     "algorithmic_bytes_per_launch": abytes,
                     "note": "SURVEY §8d accounting: 256 B per ray-step as if the 128-byte ray state were streamed through HBM.  The state is "
                             "register-resident, so these are NOT DRAM bytes (achieved_dram_gbs is the measured DRAM rate) and frac > 1 "
                             "only says the kernel beats any streamed formulation (25.5 G ray-steps/s ceiling)"}
        if inst_per_step:
            # DRAM traffic is ~1e-3 of the algorithmic bytes: the binding limit is warp-instruction issue (1 per SM sub-partition
            # per clock), with the FP32 FMA pipe and register-operand bandwidth within a few % of it
            roofline = {"bound": "issue", "achieved": inst_per_step * warp_steps_per_s / 1e9, "peak": slots / 1e9, "unit": "G warp-inst/s",
                        "frac": inst_per_step * warp_steps_per_s / slots, "traffic": traffic,
                        "achieved_dram_gbs": (traffic / (kernel_ms * 1e-3) / 1e9) if traffic else None, "hbm_peak_gbs": peak,
                        "warp_inst_per_warp_ray_step": inst_per_step, "kernel": "trace_kernel<1,false,4,origin>", "kernel_ms": kernel_ms,
                        "constants": f"profiles/trace_kernel_dram.json ({pj_state}; kernel source {pj.get('source_hash')}): ncu --set full capture of this kernel; "
                                     "time and SM clock measured live",
                        "peak_source": "148 SM x 4 schedulers x SM clock sampled during the run",
                        "why_not_hbm": "measured DRAM traffic per launch is the output frame (+ ~10 % scene reads): < 0.2 % of HBM bandwidth"}
        else:
            roofline = dict(hbm_block)
            roofline["constants"] = pj_state
        if e2e_error is not None:
            e2e_block = e2e_serial_block = e2e_hdr_block = {"value": None, "unit": UNIT, "error": "no host frame: " + e2e_error}
        else:
            e2e_block = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(pinned_model.numel() + 196) * world,
                        "d2h_bytes_per_step": int(W * H * 16), "ms_per_step": 1000.0 * e2e_s / args.steps, "fps": args.steps / e2e_s,
                        "frames_in_flight": 2, "frame_checksum": checksum, "both_host_frames_equal": e2e_lanes_equal,
                        "path": ("two frames in flight, each on its own context / stream / host frame: "
                                 + ("bh_ctx_upload_models_async + bh_ray_pipeline_pass_to_host("
                                    + ("zero-copy: pixel stores land in the pinned host frame over PCIe during the pass" if args.e2e_chunks == 0
                                       else f"{args.e2e_chunks} bands, D2H overlapped") + ") + sync + host read of the frame"
                                    if world == 1 else
                                    "per rank bh_ctx_upload_models_async (48 MB over its own PCIe link) + bh_ray_pipeline_pass_to_host_frame (its bands "
                                    "stored straight into ONE page-locked host frame in POSIX shared memory, bh_host_frame) + host-side flags")
                                 + "; frame k+1's upload runs under frame k's pass, buffers are reused only after the frame was finished and read")}
            e2e_serial_block = {"value": total["ray_steps"] * args.steps / e2e_serial_s / 1e6, "unit": UNIT, "ms_per_step": 1000.0 * e2e_serial_s / args.steps,
                               "frames_in_flight": 1, "path": "same calls, one frame at a time: upload, pass, wait, read (round 1's e2e)"}
            e2e_hdr_block = {"value": total["ray_steps"] * args.steps / e2e_hdr_s / 1e6, "unit": UNIT, "ms_per_step": 1000.0 * e2e_hdr_s / args.steps,
                                    "h2d_bytes_per_step": 16 * world + 196 * world, "d2h_bytes_per_step": int(W * H * 16),
                                    "frames_in_flight": 1,
                                    "path": "bh_ctx_set_model_header (position + visible, the fields the UI edits) instead of re-sending the 48 MB ModelUniform"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "ms_per_step_incl_flush": elapsed_flush_ms / args.steps, "fps": 1000.0 / ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": f"synthetic (default camera/black hole uniforms); textures={tex_src}; mesh={mesh_src}",
            "numeric_mode": args.numeric_mode,
            f"value_{other_name}": total["ray_steps"] / (ms_other * 1e-3) / 1e6, f"ms_per_step_{other_name}": ms_other,
            "config": {"workload": WORKLOAD if (W, H) == (3840, 2160) else f"{W}x{H} variant of: {WORKLOAD}", "width": W, "height": H,
                       "integrator": "cash-karp-rk", "step_size": 0.15, "max_iterations": 2000, "triangles": mesh_info.get("triangle_count"),
                       "bvh_nodes": mesh_info.get("nodes_used"), "tiling": f"cyclic bands of {frame.band_rows} rows over {world} rank(s)",
                       "exchange": {"p2p": "ray kernel stores finished pixels directly into rank 0's frame over NVLink (CUDA IPC peer memory); ordering by "
                                           "stream-ordered flag stores / waits in that memory (no collective in the frame loop)",
                                    "nccl": "NCCL gather of compact band buffers to rank 0 + de-interleave copy", "none": "single GPU"}[frame.exchange],
                       "l2_flush": "256 MiB memset before every step, between the per-step event pairs that are summed (ms_per_step_incl_flush: one event "
                                   "pair around all steps, flushes included)",
                       "ray_steps_per_frame": total["ray_steps"],
                       "numerics": ("FUSED: explicit fma contraction + reciprocal-multiply, det-math transcendentals (bit-exact vs oracle 'fused')"
                                    if mode == P.NUMERIC_FUSED else
                                    "LITERAL: one IEEE f32 op per WGSL node, det-math transcendentals (bit-exact vs oracle 'contract')"),
                       "kernel_source_hash": B.kernel_source_hash()},
            "e2e": e2e_block, "e2e_serial": e2e_serial_block, "e2e_header_only": e2e_hdr_block,
            "gpu_launches": int(args.steps * (1 if world == 1 else 2)),
            "gpu_launches_note": "rank 0, timed region: trace_kernel per step" + ("" if world == 1 else " + the flag-wait kernel (other ranks: wait + trace + signal)"),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_hbm_streamed": hbm_block,
            "roofline_fp32": {"bound": "fp32-alu", "achieved": flops / (kernel_ms * 1e-3) / 1e12, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                              "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp32_peak_tflops, "flops_per_ray_step": 400,
                              "peak_source": "148 SM x 128 lanes x 2 x sm_max_mhz (non-tensor FP32, BASELINE.md §2)"},
        }
        if world > 1:
            line["per_rank"] = per_rank
            line["exchange_bit_identical"] = exchange_bit_identical
            line["e2e"]["host_frame_equals_device_frame"] = e2e_matches_device
        if c4 is not None:
            line["c4_8k"] = c4
        if frame_multi is not None:
            line["frame_multi"] = frame_multi
        if inst_per_step:
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
        if world == 1 and not args.no_extras:
            try:
                line["parity"] = parity_block(ctx, P, U, tex, blob, W, H, cam, hole, det)
            except Exception as ex:
                line["parity"] = {"error": str(ex)[:300]}
            try:
                line["pyramid"] = pyramid_block(ctx, P, U, torch, stream)
            except Exception as ex:
                line["pyramid"] = {"error": str(ex)[:300]}
        # CPU baseline beside it (N=1 only): bounded sample of the same workload
        if world == 1 and not args.no_cpu_baseline:
            try:
                rate, steps_c, dt, threads = cpu_oracle_rate(tex, blob, CPU_SAMPLE_RES, host_threads())
                line["cpu_baseline"] = {"value": rate / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
                                        "sample": f"{CPU_SAMPLE_RES[0]}x{CPU_SAMPLE_RES[1]} frame of the same scene/camera "
                                                  f"({steps_c} ray-steps, {dt:.1f} s); oracle strict flavour (glibc libm), OpenMP",
                                        "fps_extrapolated_3840x2160": rate / total["ray_steps"]}
            except Exception as ex:   # the oracle is a checker; its absence must not hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        print(json.dumps(line))
    for i, lane in enumerate(lanes):
        if world > 1:
            lane[2].close()
        elif i > 0:
            lane[3].close()
    ctx_b.close()
    frame.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


class _DevPtr:
    """__cuda_array_interface__ carrier: lets torch view device memory the library owns."""

    def __init__(self, ptr: int, shape: tuple):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2, "strides": None}


def frame_multi_block(P, U, torch, ctx0, tex, blob, world, W, H, band_rows, cam, hole, det, reps=5):
    """bh_frame_multi on all `world` devices from this one process and thread: single-level 4K frame and the reference's
    4-level adaptive grid, both compared bit for bit with single-GPU renders, plus the frame time."""
    ctxs = [ctx0]
    for d in range(1, world):
        c = P.Context(d, numeric_mode=ctx0.numeric_mode)
        c.set_textures(tex)
        c.upload_models(blob)
        ctxs.append(c)
    out = {"devices": world, "driver": "one process, one host thread; peer stores into device 0's frame, CUDA events; no torch / NCCL on the path"}
    try:
        fm = P.FrameMulti(ctxs, base=(W, H), iters=1, band_rows=band_rows, sky_format=None)
        fm.pass_(cam, hole, det)
        got = fm.read(sky=False)["rgba"]
        single = P.RayPipeline(ctx0, W, H)
        single.pass_(cam, hole, det)
        ref = single.read(aux=False)["rgba"]
        single.close()
        out["single_level_bit_identical"] = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        for _ in range(3):
            fm.pass_(cam, hole, det)
        fm.sync()
        t0 = time.perf_counter()
        for _ in range(reps):
            fm.pass_(cam, hole, det)
        fm.sync()
        out["single_level_ms_per_frame_host_clock"] = 1000.0 * (time.perf_counter() - t0) / reps
        out["single_level_ms_last_frame_device"] = fm.stats()["elapsed_ms"]
        fm.close()
        # the reference's own frame: coarse levels replicated, last level tiled, sky on device 0
        dete = U.RayDetails(integration_method=0, model_count=1)
        fp = P.FrameMulti(ctxs, base=(72, 41), iters=4, band_rows=band_rows, sky_format=P.SKY_RGBA16F)
        fp.pass_(cam, hole, dete)
        got = fp.read(rgba=False)["sky"]
        pyr = P.RayPyramid(ctx0)
        pyr.pass_(cam, hole, dete)
        ref = pyr.sky.read()
        pyr.close()
        out["pyramid_sky_bit_identical"] = bool(np.array_equal(got.view(np.uint16), ref.view(np.uint16)))
        for _ in range(3):
            fp.pass_(cam, hole, dete)
        fp.sync()
        t0 = time.perf_counter()
        for _ in range(reps * 2):
            fp.pass_(cam, hole, dete)
        fp.sync()
        out["pyramid_1918x1081_euler_ms_per_frame_host_clock"] = 1000.0 * (time.perf_counter() - t0) / (reps * 2)
        fp.close()
    finally:
        for c in ctxs[1:]:
            c.close()
    return out


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
    ap.add_argument("--c4", default="auto", choices=["auto", "on", "off"], help="N>1: also time BASELINE configs[3] (7680x4320); auto = at N=8")
    ap.add_argument("--no-frame-multi", action="store_true", help="N>1: skip rank 0's bh_frame_multi check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N=1: skip the parity and pyramid blocks (profiling runs)")
    ap.add_argument("--allow-short-warmup", action="store_true", help="profiling runs only (ncu); numbers from such runs are not bench values")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.allow_short_warmup:
        args.warmup = 3
    sys.exit(run_reference(args) if args.impl == "reference" else run_ours(args))


if __name__ == "__main__":
    main()
