#!/usr/bin/env python
"""Round measurements beyond the headline bench (BASELINE.json configs C1, C2, C5 + per-kernel numbers):
   * reference pyramid frame (72x41 -> 1918x1081, 4 levels + sky resolve): ms/frame, fps, per-level stats
   * 4K pyramid (143x81 base -> 3835x2161)
   * single-level C1 (256x256 Euler) and C2 (1920x1080 RK, no mesh)
   * sky_kernel and classify_kernel effective GB/s (HBM-bound kernels)
   * C5 step sweep at 3840x2160: max_iterations in {64..2048}, R_rel = 20 and 1000
Writes gpurun_out/sweeps.json.  Device-event timing, 3 warm-up + 5 timed passes each."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402


def timed(fn, stream, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    out = {}
    tex, src = assets.load_textures()
    blob, info = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    ctx = P.Context(0)
    ctx.set_textures(tex)
    ctx.upload_models(blob)
    s = torch.cuda.current_stream()
    cam, hole = U.Camera(), U.BlackHole()
    out["assets"] = {"textures": src, "mesh": "lucy.obj" if assets.have_lucy() else "uv_sphere"}

    # ---- pyramids
    for name, base, iters in (("pyramid_reference_1918x1081", (72, 41), 4), ("pyramid_4k_3835x2161", (143, 81), 4)):
        for method, mname in ((1, "rk"), (0, "euler")):
            det = U.RayDetails(integration_method=method, model_count=1)
            pyr = P.RayPyramid(ctx, base=base, iters=iters)
            ms = timed(lambda: pyr.pass_(cam, hole, det, s), s)
            levels = []
            for rp in pyr.levels:
                st = rp.stats()
                levels.append({"size": [rp.width, rp.height], **{k: st[k] for k in ("ray_steps", "px_traced", "px_copied", "px_interp")}})
            steps = sum(l["ray_steps"] for l in levels)
            out[f"{name}_{mname}"] = {"ms_per_frame": ms, "fps": 1000.0 / ms, "ray_steps": steps, "gsteps_per_s": steps / ms / 1e6,
                                      "levels": levels, "sizes": pyr.sizes}
            print(name, mname, out[f"{name}_{mname}"]["ms_per_frame"], out[f"{name}_{mname}"]["fps"], flush=True)
            # per-kernel: last level classify+trace vs sky
            last = pyr.levels[-1]
            ms_last = timed(lambda: last.pass_(cam, hole, det, s), s)
            ms_sky = timed(lambda: pyr.sky.pass_(s), s)
            npx = last.width * last.height
            out[f"{name}_{mname}"]["last_level_ms"] = ms_last
            out[f"{name}_{mname}"]["sky_ms"] = ms_sky
            out[f"{name}_{mname}"]["sky_gbs_algorithmic"] = npx * 24 / ms_sky / 1e6      # 16 B in + 8 B out per pixel
            pyr.close()

    # ---- single-level configs
    for name, (w, h), method, mc in (("c1_256x256_euler_nomesh", (256, 256), 0, 0), ("c2_1920x1080_rk_nomesh", (1920, 1080), 1, 0),
                                      ("c3_3840x2160_rk_mesh", (3840, 2160), 1, 1), ("c3_3840x2160_rk_mesh_cam45", (3840, 2160), 1, 1)):
        c = U.Camera(position=(0, 0, -45)) if name.endswith("cam45") else cam
        det = U.RayDetails(integration_method=method, model_count=mc)
        rp = P.RayPipeline(ctx, w, h)
        ms = timed(lambda: rp.pass_(c, hole, det, s), s)
        st = rp.stats()
        out[name] = {"ms": ms, "fps": 1000.0 / ms, "ray_steps": st["ray_steps"], "gsteps_per_s": st["ray_steps"] / ms / 1e6, "stats": st}
        print(name, out[name]["ms"], out[name]["gsteps_per_s"], flush=True)
        rp.close()

    # ---- C5 sweep
    sweep = []
    rp = P.RayPipeline(ctx, 3840, 2160)
    for rrel in (20.0, 1000.0):
        for mi in (64, 128, 256, 512, 1024, 2048):
            det = U.RayDetails(integration_method=1, model_count=1, max_iterations=mi)
            hl = U.BlackHole(relativity_sphere_radius=rrel)
            ms = timed(lambda: rp.pass_(cam, hl, det, s), s, warm=2, reps=3)
            st = rp.stats(strict=False)
            row = {"relativity_radius": rrel, "max_iterations": mi, "ms": ms, "ray_steps": st["ray_steps"],
                   "gsteps_per_s": st["ray_steps"] / ms / 1e6, "algorithmic_gbs": 256 * st["ray_steps"] / ms / 1e6, "rk_reject": st["rk_reject"]}
            sweep.append(row)
            print("sweep", row, flush=True)
    rp.close()
    out["c5_sweep_3840x2160_rk"] = sweep
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweeps.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
