#!/usr/bin/env python
"""How much of a small level's latency is cold instruction fetch?  Level 0 (72x41, tile mode) and level 1 (queue mode) of the
reference frame timed (a) back to back with themselves, (b) inside the whole frame (other kernels in between)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402
from bhusie_b200.post import PostChain                           # noqa: E402

tex, _ = assets.load_textures()
blob, _ = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
ctx = P.Context(0)
ctx.set_textures(tex); ctx.upload_models(blob)
s = torch.cuda.current_stream()
cam, hole = U.Camera(), U.BlackHole()


def ev():
    return torch.cuda.Event(enable_timing=True)


for method, name in ((0, "euler"), (1, "rk")):
    det = U.RayDetails(integration_method=method, model_count=1)
    pyr = P.RayPyramid(ctx, base=(72, 41))
    chain = PostChain(ctx, pyr.sky)
    for _ in range(3):
        pyr.pass_(cam, hole, det, s); chain.pass_(s)
    torch.cuda.synchronize()
    for li in (0, 1, 2):
        rp = pyr.levels[li]
        # (a) the level alone, back to back
        a, b = ev(), ev()
        for _ in range(3):
            rp.pass_(cam, hole, det, s)
        a.record(s)
        for _ in range(20):
            rp.pass_(cam, hole, det, s)
        b.record(s)
        torch.cuda.synchronize()
        alone = a.elapsed_time(b) / 20
        # (b) inside whole frames
        tot = 0.0
        for _ in range(10):
            for lj, q in enumerate(pyr.levels):
                if lj == li:
                    a, b = ev(), ev()
                    a.record(s); q.pass_(cam, hole, det, s); b.record(s)
                else:
                    q.pass_(cam, hole, det, s)
            pyr.sky.pass_(s); chain.pass_(s)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        print(f"{name} level {li}: alone back-to-back {alone:.4f} ms, inside the frame {tot / 10:.4f} ms", flush=True)
    chain.close(); pyr.close()
