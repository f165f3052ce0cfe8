#!/usr/bin/env python
"""Queue-mode scheduling sweep of trace_kernel (env BH_TUNE = park,serve_div,refill_div): the reference pyramid, per level.
   python tools/gpu_tune.py TAG [tune ...]   -> gpurun_out/TAG_tune.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch
    from bhusie_b200 import assets, pipelines as P, uniforms as U
    from tools.gpu_quick import timed
    tex, src = assets.load_textures()
    blob, _ = P.load_obj_model(assets.lucy_path())
    ctx = P.Context(0)
    ctx.set_textures(tex); ctx.upload_models(blob)
    s = torch.cuda.current_stream()
    cam, hole = U.Camera(), U.BlackHole()
    out = {}
    for name, base in (("ref", (72, 41)), ("4k", (143, 81))):
        for method, mname in ((1, "rk"), (0, "euler")):
            det = U.RayDetails(integration_method=method, model_count=1)
            pyr = P.RayPyramid(ctx, base=base)
            ms = timed(lambda: pyr.pass_(cam, hole, det, s), s, 3, 12)
            lv = [timed(lambda: rp.pass_(cam, hole, det, s), s, 2, 8) for rp in pyr.levels]
            out[f"{name}_{mname}"] = {"ms": ms, "levels": lv}
            pyr.close()
    print("RESULT " + json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        return child()
    tag = sys.argv[1]
    tunes = sys.argv[2:] or ["0,0,0", "0,0,2", "0,0,4", "0,2,2", "0,4,4", "0,4,1", "1,4,4", "1,2,2"]
    res = {}
    for t in tunes:
        env = dict(os.environ, BH_TUNE=t)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        res[t] = json.loads(line[0][7:]) if line else {"error": r.stderr[-500:]}
        if line:
            d = res[t]
            print(t, " ".join(f"{k}: {v['ms']:.3f} [" + " ".join(f"{x:.3f}" for x in v["levels"]) + "]" for k, v in d.items()), flush=True)
        else:
            print(t, "ERROR", r.stderr[-300:], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"{tag}_tune.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
