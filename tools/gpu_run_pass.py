#!/usr/bin/env python
"""Runs a few single-level passes (for ncu): gpu_run_pass.py MODE METHOD W H [REPS] [MODEL_COUNT]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402

mode = {"literal": P.NUMERIC_LITERAL, "fused": P.NUMERIC_FUSED}[sys.argv[1]]
method = {"euler": 0, "rk": 1}[sys.argv[2]]
w, h = int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
mc = int(sys.argv[6]) if len(sys.argv) > 6 else 1
tex, src = assets.load_textures()
blob, info = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
ctx = P.Context(0, numeric_mode=mode)
ctx.set_textures(tex)
ctx.upload_models(blob)
rp = P.RayPipeline(ctx, w, h)
for _ in range(reps):
    rp.pass_(U.Camera(), U.BlackHole(), U.RayDetails(integration_method=method, model_count=mc))
torch.cuda.synchronize()
print(rp.stats())
