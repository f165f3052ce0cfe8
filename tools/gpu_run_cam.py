#!/usr/bin/env python
"""gpu_run_cam.py W H Z [REPS]: fused RK pass with the camera at (0,0,Z) (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bhusie_b200 import assets, pipelines as P, uniforms as U
w, h, z = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
tex, src = assets.load_textures()
blob, info = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
ctx = P.Context(0)
ctx.set_textures(tex); ctx.upload_models(blob)
rp = P.RayPipeline(ctx, w, h)
for _ in range(reps):
    rp.pass_(U.Camera(position=(0, 0, z)), U.BlackHole(), U.RayDetails(integration_method=1, model_count=1))
torch.cuda.synchronize()
print(rp.stats())
