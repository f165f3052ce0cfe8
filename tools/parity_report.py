#!/usr/bin/env python
"""SURVEY §8c parity protocol, evaluated on the CPU: for every BASELINE.json configuration the oracle flavour each kernel
numeric mode is bit-identical with ("contract" = LITERAL, "fused" = FUSED; the GPU tests assert that identity) is compared
with the neutral strict (glibc libm) flavour, and the float64 shadow classifies the outliers.

    python tools/parity_report.py [--out profiles/r2_parity_report.json] [--quick]

Prints / writes per configuration and mode: outlier_frac (pixels with a channel beyond 1e-4), max_abs, rms, the share of the
outliers that the shadow marks ill-conditioned, the outlier fraction left on well-conditioned pixels, and the agreement of hit
indices and step counts.  The thresholds in tests/test_gpu_parity.py are 2x these measured values.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bhusie_b200 import assets, uniforms as U          # noqa: E402  (pure Python: no CUDA library is loaded)
from oracle import oracle as O                         # noqa: E402


def report(name, sc, w, h, cam, hole, det, prev=None, out=None):
    res = {}
    for f in O.FLAVOURS:
        t = time.time()
        res[f] = O.ray_pass(sc, w, h, cam, hole, det, prev=prev, flavour=f)
        print(f"  {name} {w}x{h} {f}: {time.time() - t:.1f} s, {res[f].counters['steps']} steps", flush=True)
    probe = O.ray_pass(sc, w, h, cam, hole, det, prev=prev, flavour="shadow", perturb=O.SHADOW_PERTURBATION)
    entry = {"size": [w, h], "ray_steps_strict": res["strict"].counters["steps"]}
    for mode, f in (("literal", "contract"), ("fused", "fused")):
        r = O.parity_report(res[f].rgba, res["strict"].rgba, res["shadow"].rgba, 1e-4, probe.rgba)
        r["hit_index_equal_frac"] = float((res[f].hit == res["strict"].hit).mean())
        r["step_count_equal_frac"] = float((res[f].steps == res["strict"].steps).mean())
        r["class_equal_frac"] = float((res[f].cls == res["strict"].cls).mean())
        entry[mode] = r
        print(f"  {name} {mode}: outlier_frac {r['outlier_frac']:.3e} max_abs {r['max_abs']:.3g} ill-conditioned share "
              f"{r['outliers_ill_conditioned_share']:.3f} well-conditioned outlier_frac {r['outlier_frac_well_conditioned']:.3e} "
              f"hit== {r['hit_index_equal_frac']:.6f} steps== {r['step_count_equal_frac']:.6f}", flush=True)
    if out is not None:
        out[name] = entry
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity_report.json"))
    ap.add_argument("--quick", action="store_true", help="quarter-area frames")
    args = ap.parse_args()
    O.build()
    tex, tex_src = assets.load_textures()
    if assets.have_lucy():
        blob, info = O.load_obj(assets.lucy_path())
        mesh = "lucy.obj"
    else:
        raise SystemExit("lucy.obj not staged (tools/stage_assets.py)")
    sc = O.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    cam, hole = U.Camera().uniform(), U.BlackHole().uniform()
    q = 2 if args.quick else 1
    out = {"tolerance": 1e-4, "textures": tex_src, "mesh": mesh, "threads": O.max_threads(),
           "note": "flavour vs strict on the CPU; the kernel's LITERAL / FUSED modes are bit-identical with contract / fused (tests/test_gpu_fullsize.py)"}
    # C1: 256x256 Euler, disk only
    report("C1", sc, 256, 256, cam, hole, U.RayDetails(integration_method=0, model_count=0).uniform(), out=out)
    # C2(i): 1920x1080 RK single level, disk
    report("C2i", sc, 1920 // q, 1080 // q, cam, hole, U.RayDetails(integration_method=1, model_count=0).uniform(), out=out)
    # C2(ii): the reference's adaptive grid, every level with the strict flavour's previous level as input
    det = U.RayDetails(integration_method=1, model_count=0).uniform()
    prev = None
    for (w, h) in U.pyramid_levels(iters=4 if not args.quick else 3):
        res = report(f"C2ii_level_{w}x{h}", sc, w, h, cam, hole, det, prev=prev, out=out)
        prev = res["strict"].rgba
    # C3: 3840x2160 RK, disk + sphere + lucy BVH; and the camera outside R_rel
    det = U.RayDetails(integration_method=1, model_count=1).uniform()
    report("C3", sc, 3840 // q, 2160 // q, cam, hole, det, out=out)
    report("C3_cam45_quarter", sc, 1920 // q, 1080 // q, U.Camera(position=(0, 0, -45)).uniform(), hole, det, out=out)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
