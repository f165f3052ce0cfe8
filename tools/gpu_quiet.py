#!/usr/bin/env python
"""Steady-state rate of the hot loop: 3840x2160 RK, no mesh, R_rel = 1000, max_iterations = 512 (every ray runs 512 steps,
almost all of them 'quiet'), beside the default scene.  Device events, 2 warm-up + 3 timed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402
from gpu_sweeps import timed                                     # noqa: E402


def main():
    tex, src = assets.load_textures()
    ctx = P.Context(0)
    ctx.set_textures(tex)
    blob, info = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    ctx.upload_models(blob)
    s = torch.cuda.current_stream()
    rp = P.RayPipeline(ctx, 3840, 2160)
    res = {}
    for name, method, rrel, mi, mc in (("rk_quiet_R1000_512", 1, 1000.0, 512, 0), ("euler_quiet_R1000_512", 0, 1000.0, 512, 0),
                                       ("rk_default_nomesh", 1, 20.0, 2000, 0), ("rk_default_mesh", 1, 20.0, 2000, 1)):
        det = U.RayDetails(integration_method=method, model_count=mc, max_iterations=mi)
        hl = U.BlackHole(relativity_sphere_radius=rrel)
        ms = timed(lambda: rp.pass_(U.Camera(), hl, det, s), s, warm=2, reps=3)
        st = rp.stats(strict=False)
        res[name] = {"ms": ms, "gsteps_per_s": st["ray_steps"] / ms / 1e6, "ray_steps": st["ray_steps"]}
        print(name, res[name], flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("BH_TIME_OUT", "quiet.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
