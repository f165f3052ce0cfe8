#!/usr/bin/env python
"""Quick GPU bring-up: parity of the CUDA pass vs the oracle (contract flavour: bit-exact expected;
strict flavour: <=1e-4 expected) on small frames, plus a first timing.  Writes gpurun_out/quick.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bhusie_b200 import assets, pipelines as P, uniforms as U   # noqa: E402
from oracle import oracle as O                                   # noqa: E402  (checker only)

out = {}


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def compare(name, dev, ora, strict=None):
    r = {}
    same = bits(dev["rgba"]) == bits(ora.rgba)
    r["rgba_bit_exact_frac"] = float(same.mean())
    d = np.abs(dev["rgba"].astype(np.float64) - ora.rgba.astype(np.float64))
    r["rgba_maxabs_vs_contract"] = float(np.nanmax(d))
    if "hit" in dev:
        r["hit_equal_frac"] = float((dev["hit"] == ora.hit).mean())
        r["hit_pixels"] = int((ora.hit >= 0).sum())
    if "steps" in dev:
        r["steps_equal_frac"] = float((dev["steps"] == ora.steps).mean())
    if "cls" in dev:
        r["class_equal_frac"] = float((dev["cls"] == ora.cls).mean())
    if strict is not None:
        ds = np.abs(dev["rgba"].astype(np.float64) - strict.rgba.astype(np.float64))
        r["rgba_maxabs_vs_strict"] = float(np.nanmax(ds))
        r["frac_gt_1e-4_vs_strict"] = float((ds > 1e-4).any(axis=2).mean())
        r["hit_equal_frac_vs_strict"] = float((dev["hit"] == strict.hit).mean()) if "hit" in dev else None
        r["steps_equal_frac_vs_strict"] = float((dev["steps"] == strict.steps).mean()) if "steps" in dev else None
    out[name] = r
    print(name, json.dumps(r))


def main():
    tex, src = assets.load_textures()
    out["textures"] = src
    ctx = P.Context(0)
    ctx.set_textures(tex)
    osc = O.OracleScene(tex["color"], tex["disk"], tex["sky"])
    cam, hole = U.Camera(), U.BlackHole()

    # det-math probe
    rng = np.random.default_rng(1)
    n = 200000
    probes = {
        "pow": (rng.uniform(0, 50, n), rng.uniform(-3, 3, n)),
        "pow5": (rng.uniform(0, 100, n), None), "pow4": (rng.uniform(0, 1, n), None),
        "sin": (rng.uniform(-50, 50, n), None), "cos": (rng.uniform(-50, 50, n), None),
        "tan": (rng.uniform(-1.5, 1.5, n), None),
        "atan2": (rng.normal(size=n), rng.normal(size=n)), "acos": (rng.uniform(-1.01, 1.01, n), None),
    }
    pm = {}
    for fn, (a, b) in probes.items():
        a = a.astype(np.float32)
        b = None if b is None else b.astype(np.float32)
        d = ctx.math_probe(fn, a, b)
        o = O.math_array(fn, a, b, flavour="contract")
        pm[fn] = int((bits(d) != bits(o)).sum() - (np.isnan(d) & np.isnan(o)).sum())
    out["detmath_mismatches"] = pm
    print("detmath mismatches", pm)

    aux = P.AUX_HIT | P.AUX_STEPS | P.AUX_CLASS
    # C1: 256x256 Euler, disk only
    for name, (w, h), method in (("c1_euler_256", (256, 256), 0), ("rk_192x108", (192, 108), 1)):
        det = U.RayDetails(integration_method=method)
        rp = P.RayPipeline(ctx, w, h, aux=aux)
        rp.pass_(cam, hole, det)
        dev = rp.read()
        st = rp.stats()
        ora = O.ray_pass(osc, w, h, cam.uniform(), hole.uniform(), det.uniform(), flavour="contract")
        stri = O.ray_pass(osc, w, h, cam.uniform(), hole.uniform(), det.uniform(), flavour="strict")
        compare(name, dev, ora, stri)
        out[name]["stats"] = st
        out[name]["oracle_counters"] = ora.counters
        rp.close()

    # lucy / synthetic mesh, two cameras
    if assets.have_lucy():
        blob, info = P.load_obj_model(assets.lucy_path())
        out["mesh"] = {"source": "lucy.obj", **info}
    else:
        blob, info = P.model_from_arrays(*assets.uv_sphere())
        out["mesh"] = {"source": "uv_sphere", **info}
    ctx.upload_models(blob)
    osc.models = blob
    for name, campos in (("mesh_cam19", (0, 0, -19)), ("mesh_cam45", (0, 0, -45)), ("mesh_cam_side", (-30, 5, 30))):
        c = U.Camera(position=campos, forward=(0.6, -0.1, 0.0) if name == "mesh_cam_side" else (0, 0, 1))
        det = U.RayDetails(integration_method=1, model_count=1)
        w, h = 320, 180
        rp = P.RayPipeline(ctx, w, h, aux=aux)
        rp.pass_(c, hole, det)
        dev = rp.read()
        st = rp.stats()
        ora = O.ray_pass(osc, w, h, c.uniform(), hole.uniform(), det.uniform(), flavour="contract")
        stri = O.ray_pass(osc, w, h, c.uniform(), hole.uniform(), det.uniform(), flavour="strict")
        compare(name, dev, ora, stri)
        out[name]["stats"] = st
        out[name]["oracle_counters"] = ora.counters
        rp.close()

    # pyramid 3 levels + sky
    pyr = P.RayPyramid(ctx, base=(24, 14), iters=3, aux=aux, sky_format=P.SKY_RGBA32F)
    det = U.RayDetails(integration_method=1, model_count=1)
    pyr.pass_(cam, hole, det)
    prev = None
    for li, rp in enumerate(pyr.levels):
        dev = rp.read()
        ora = O.ray_pass(osc, rp.width, rp.height, cam.uniform(), hole.uniform(), det.uniform(), prev=prev, flavour="contract")
        compare(f"pyramid_level{li}_{rp.width}x{rp.height}", dev, ora)
        out[f"pyramid_level{li}_{rp.width}x{rp.height}"]["stats"] = rp.stats()
        prev = ora.rgba
    sky_dev = pyr.sky.read()
    o32, o16, _ = O.sky_pass(osc, prev, flavour="contract")
    out["sky_f32_bit_exact_frac"] = float((bits(sky_dev) == bits(o32)).mean())
    pyr16 = P.SkyPipeline(ctx, pyr.levels[-1], P.SKY_RGBA16F)
    pyr16.pass_()
    s16 = pyr16.read()
    out["sky_f16_bit_exact_frac"] = float((s16.view(np.uint16) == o16).mean())
    print("sky", out["sky_f32_bit_exact_frac"], out["sky_f16_bit_exact_frac"])

    # first timing: 1920x1080 RK single level, no mesh / with mesh
    import torch
    for name, mc in (("time_1080p_rk_nomesh", 0), ("time_1080p_rk_mesh", 1)):
        det = U.RayDetails(integration_method=1, model_count=mc)
        rp = P.RayPipeline(ctx, 1920, 1080)
        s = torch.cuda.current_stream()
        for _ in range(2):
            rp.pass_(cam, hole, det, s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3):
            rp.pass_(cam, hole, det, s)
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        st = rp.stats()
        out[name] = {"ms": ms, "steps": st["ray_steps"], "gsteps_per_s": st["ray_steps"] / ms / 1e6, "stats": st}
        print(name, out[name])
        rp.close()

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "quick.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
