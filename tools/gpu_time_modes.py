#!/usr/bin/env python
"""Times the single-level RK pass in both numeric modes at a given size (device events, 3 warm-up, 5 timed)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
    tex, src = assets.load_textures()
    blob, info = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    res = {}
    modes = ((P.NUMERIC_LITERAL, "literal"), (P.NUMERIC_FUSED, "fused"))
    if os.environ.get("BH_TIME_FUSED_ONLY"):
        modes = modes[1:]
    for mode, name in modes:
        ctx = P.Context(0, numeric_mode=mode)
        ctx.set_textures(tex)
        ctx.upload_models(blob)
        for method, mname in ((1, "rk"), (0, "euler")):
            cam, hole, det = U.Camera(), U.BlackHole(), U.RayDetails(integration_method=method, model_count=1)
            rp = P.RayPipeline(ctx, w, h)
            s = torch.cuda.current_stream()
            for _ in range(3):
                rp.pass_(cam, hole, det, s)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(5):
                rp.pass_(cam, hole, det, s)
            e1.record(s)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            st = rp.stats()
            res[f"{name}_{mname}"] = {"ms": ms, "gsteps_per_s": st["ray_steps"] / ms / 1e6, "steps": st["ray_steps"]}
            print(name, mname, res[f"{name}_{mname}"], flush=True)
            rp.close()
        ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("BH_TIME_OUT", "time_modes.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
