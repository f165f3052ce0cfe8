#!/usr/bin/env python
"""Whole-frame timing of what the reference renders every frame (mod.rs:406-431): 4-level ray pyramid + sky resolve +
bloom x10 + mix + ACES + FXAA, at 1918x1081 (and a 4K pyramid).  Device events; writes gpurun_out/frame.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402
from bhusie_b200.post import PostChain                           # noqa: E402


def timed(fn, s, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        fn()
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    tex, src = assets.load_textures()
    blob, _ = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    ctx = P.Context(0)
    ctx.set_textures(tex); ctx.upload_models(blob)
    s = torch.cuda.current_stream()
    cam, hole = U.Camera(), U.BlackHole()
    out = {}
    for name, base in (("reference_1918x1081", (72, 41)), ("4k_3835x2161", (143, 81))):
        for method, mname in ((0, "euler_default"), (1, "rk")):
            det = U.RayDetails(integration_method=method, model_count=1)
            pyr = P.RayPyramid(ctx, base=base)
            chain = PostChain(ctx, pyr.sky)
            ms_ray = timed(lambda: pyr.pass_(cam, hole, det, s), s)
            ms_post = timed(lambda: chain.pass_(s), s)
            ms_all = timed(lambda: (pyr.pass_(cam, hole, det, s), chain.pass_(s)), s)
            stages = {}
            for i, bp in enumerate(chain.blooms):
                stages[f"bloom{i}_{'down' if i < 5 else 'up'}_{bp.width}x{bp.height}"] = timed(lambda: bp.pass_(s), s, 2, 20)
            stages["mix"] = timed(lambda: chain.mix.pass_(chain.mix_details, s), s, 2, 20)
            stages["hdr"] = timed(lambda: chain.hdr.pass_(s), s, 2, 20)
            stages["fxaa"] = timed(lambda: chain.fxaa.pass_(chain.fxaa_details, s), s, 2, 20)
            w, h = pyr.levels[-1].width, pyr.levels[-1].height
            npx = w * h
            gbs = {"mix": npx * 24 / stages["mix"] / 1e6, "hdr": npx * 16 / stages["hdr"] / 1e6, "fxaa_min": npx * 12 / stages["fxaa"] / 1e6}
            out[f"{name}_{mname}"] = {"ms_ray_pyramid_plus_sky": ms_ray, "ms_post_chain": ms_post, "ms_frame": ms_all, "fps": 1000.0 / ms_all,
                                      "post_stage_ms": stages, "post_algorithmic_gbs": gbs, "size": [w, h]}
            print(name, mname, {k: round(v, 4) for k, v in out[f"{name}_{mname}"].items() if isinstance(v, float)}, flush=True)
            chain.close(); pyr.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "frame.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
