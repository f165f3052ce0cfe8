#!/usr/bin/env python
"""Quick device-event timings of the cases the round-2 kernel work targets -> gpurun_out/<tag>_quick.json
   headline 4K (RK, Euler; fused + literal), camera outside the sphere (cam45), the reference pyramid per level, 4K pyramid."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                     # noqa: E402
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402


def timed(fn, s, warm=3, reps=8):
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
    tag = sys.argv[1] if len(sys.argv) > 1 else "quick"
    tex, src = assets.load_textures()
    blob, _ = P.load_obj_model(assets.lucy_path()) if assets.have_lucy() else P.model_from_arrays(*assets.uv_sphere())
    ctx = P.Context(0)
    ctx.set_textures(tex); ctx.upload_models(blob)
    s = torch.cuda.current_stream()
    hole = U.BlackHole()
    out = {}
    rp = P.RayPipeline(ctx, 3840, 2160)
    for name, cam, method, mode in (("c3_rk_fused", U.Camera(), 1, P.NUMERIC_FUSED), ("c3_euler_fused", U.Camera(), 0, P.NUMERIC_FUSED),
                                    ("c3_rk_literal", U.Camera(), 1, P.NUMERIC_LITERAL), ("c3_cam45_rk_fused", U.Camera(position=(0, 0, -45)), 1, P.NUMERIC_FUSED),
                                    ("c3_cam45_euler_fused", U.Camera(position=(0, 0, -45)), 0, P.NUMERIC_FUSED)):
        ctx.set_numeric_mode(mode)
        det = U.RayDetails(integration_method=method, model_count=1)
        ms = timed(lambda: rp.pass_(cam, hole, det, s), s, 2, 5)
        st = rp.stats()
        out[name] = {"ms": ms, "gsteps_per_s": st["ray_steps"] / ms / 1e6, "ray_steps": st["ray_steps"], "node_visits": st["node_visits"]}
        print(name, out[name], flush=True)
    rp.close()
    ctx.set_numeric_mode(P.NUMERIC_FUSED)
    cam = U.Camera()
    # one rank's share of a tiled frame, on one device: where does the 1 -> 8 GPU kernel-time overhead come from?
    det = U.RayDetails(integration_method=1, model_count=1)
    for world in (2, 4, 8, 16):
        t = P.RayPipeline(ctx, 3840, 2160)
        t.set_tiling(8, 1, world)
        ms = timed(lambda: t.pass_(cam, hole, det, s), s, 2, 8)
        st = t.stats()
        out[f"tiled_rank1_of_{world}"] = {"ms": ms, "ms_ideal": out["c3_rk_fused"]["ms"] * st["ray_steps"] / out["c3_rk_fused"]["ray_steps"],
                                          "ray_steps": st["ray_steps"], "local_rows": t.local_rows}
        print(f"tiled 1 of {world}", out[f"tiled_rank1_of_{world}"], flush=True)
        t.close()
    # the same with a cold L2 (what bench.py's flush before every step does): events around the pass only
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for world in (1, 8):
        t = P.RayPipeline(ctx, 3840, 2160)
        if world > 1:
            t.set_tiling(8, 1, world)
        tot = 0.0
        for _ in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); t.pass_(cam, hole, det, s); b.record(s)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        out[f"cold_l2_rank1_of_{world}"] = {"ms": tot / 6}
        print(f"cold L2, 1 of {world}", out[f"cold_l2_rank1_of_{world}"], flush=True)
        t.close()
    small = P.RayPipeline(ctx, 3840, 272)          # the same pixel count as one rank of 8, as a plain frame of its own
    ms = timed(lambda: small.pass_(cam, hole, det, s), s, 2, 8)
    out["plain_3840x272"] = {"ms": ms, "ray_steps": small.stats()["ray_steps"], "gsteps_per_s": small.stats()["ray_steps"] / ms / 1e6}
    print("plain 3840x272", out["plain_3840x272"], flush=True)
    small.close()
    for name, base in (("pyramid_1918x1081", (72, 41)), ("pyramid_3835x2161", (143, 81))):
        for method, mname in ((0, "euler"), (1, "rk")):
            det = U.RayDetails(integration_method=method, model_count=1)
            pyr = P.RayPyramid(ctx, base=base)
            ms = timed(lambda: pyr.pass_(cam, hole, det, s), s, 3, 20)
            lv = []
            prev_ok = True
            for rp in pyr.levels:
                lv.append({"size": [rp.width, rp.height], "ms": timed(lambda: rp.pass_(cam, hole, det, s), s, 2, 10), **{k: rp.stats()[k] for k in ("ray_steps", "px_traced")}})
            out[f"{name}_{mname}"] = {"ms_levels_plus_sky": ms, "levels": lv, "ms_sky": timed(lambda: pyr.sky.pass_(s), s, 2, 10)}
            print(name, mname, round(ms, 4), [round(l["ms"], 4) for l in lv], flush=True)
            pyr.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{tag}_quick.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
