#!/usr/bin/env python
"""One reference frame (4-level ray pyramid 72x41 -> 1918x1081 + sky resolve + post chain), three times, for a per-kernel
launch list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/gpu_pyramid_once.py [rk|euler]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402
from bhusie_b200.post import PostChain                           # noqa: E402

method = 0 if (len(sys.argv) > 1 and sys.argv[1] == "euler") else 1
tex, src = assets.load_textures()
blob, _ = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
ctx = P.Context(0)
ctx.set_textures(tex); ctx.upload_models(blob)
s = torch.cuda.current_stream()
det = U.RayDetails(integration_method=method, model_count=1)
pyr = P.RayPyramid(ctx, base=(72, 41))
chain = PostChain(ctx, pyr.sky)
for _ in range(3):
    pyr.pass_(U.Camera(), U.BlackHole(), det, s)
    chain.pass_(s)
torch.cuda.synchronize()
print("done", [rp.stats(strict=False)["px_traced"] for rp in pyr.levels])
