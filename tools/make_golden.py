#!/usr/bin/env python
"""Generates tests/golden/*: frozen uniform bytes and golden oracle frames.

The reference ships no golden vectors for this path, so these pin the ORACLE
against itself over time (a regression guard), not the reference.  Run from the repo root in the
build container:  python tools/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bhusie_b200 import assets, uniforms as U   # noqa: E402
from oracle import oracle as O                  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def small_scene():
    tex = assets.small_textures()
    pts, nrm, tris = assets.uv_sphere(12, 16, radius=4.0)
    blob = O.new_model_blob()
    v = O.blob_views(blob)
    v["points"][: len(pts), :3] = pts
    v["normals"][: len(nrm), :3] = nrm
    v["triangles"][: len(tris)] = tris
    v["position"][:] = (-10.0, 0.0, 30.0)
    v["visible"][0] = 1
    O.build_bvh(blob, len(tris))
    return tex, blob


def main():
    os.makedirs(GOLD, exist_ok=True)
    uni = {
        "camera_default": U.Camera().uniform().hex(),
        "black_hole_default": U.BlackHole().uniform().hex(),
        "ray_details_default": U.RayDetails().uniform().hex(),
    }
    with open(os.path.join(GOLD, "default_uniforms.json"), "w") as f:
        json.dump(uni, f, indent=1)

    tex, blob = small_scene()
    sc = O.OracleScene(tex["color"], tex["disk"], tex["sky"], blob)
    cam, hole = U.Camera().uniform(), U.BlackHole().uniform()
    gold = {"small_scene": {}, "reference_assets": {}}
    cases = {
        "euler_64x36": (64, 36, U.RayDetails(integration_method=0, model_count=1), U.Camera()),
        "rk_64x36": (64, 36, U.RayDetails(integration_method=1, model_count=1), U.Camera()),
        "rk_outside_64x36": (64, 36, U.RayDetails(integration_method=1, model_count=1), U.Camera(position=(0, 0, -45))),
    }
    for name, (w, h, det, camera) in cases.items():
        rc = O.ray_pass(sc, w, h, camera.uniform(), hole, det.uniform(), flavour="contract")
        rs = O.ray_pass(sc, w, h, camera.uniform(), hole, det.uniform(), flavour="strict")
        np.save(os.path.join(GOLD, f"small_{name}_strict_rgba.npy"), rs.rgba)
        gold["small_scene"][name] = {
            "contract_rgba_sha256": sha(rc.rgba), "contract_hit_sha256": sha(rc.hit), "contract_steps_sha256": sha(rc.steps),
            "steps_total": rc.counters["steps"], "tri_pixels": int((rc.hit >= 0).sum()),
        }
    # 2-level pyramid class map
    prev = O.ray_pass(sc, 32, 18, cam, hole, U.RayDetails(integration_method=1, model_count=1).uniform(), flavour="contract")
    lvl1 = O.ray_pass(sc, 94, 52, cam, hole, U.RayDetails(integration_method=1, model_count=1).uniform(), prev=prev.rgba, flavour="contract")
    gold["small_scene"]["pyramid_32x18_to_94x52"] = {
        "contract_rgba_sha256": sha(lvl1.rgba), "class_sha256": sha(lvl1.cls),
        "px_traced": lvl1.counters["px_traced"], "px_copied": lvl1.counters["px_copied"], "px_interp": lvl1.counters["px_interp"],
    }
    det_i = U.RayDetails(integration_method=1, model_count=1, angle_division_threshold=0.08).uniform()
    lvl1i = O.ray_pass(sc, 94, 52, cam, hole, det_i, prev=prev.rgba, flavour="contract")
    gold["small_scene"]["pyramid_32x18_to_94x52_thr0.08"] = {
        "contract_rgba_sha256": sha(lvl1i.rgba), "class_sha256": sha(lvl1i.cls),
        "px_traced": lvl1i.counters["px_traced"], "px_copied": lvl1i.counters["px_copied"], "px_interp": lvl1i.counters["px_interp"],
    }
    if assets.have_reference_assets():
        rtex, _ = assets.load_textures()
        rblob = None
        if assets.have_lucy():
            rblob, info = O.load_obj(assets.lucy_path())
            gold["reference_assets"]["lucy"] = {**info, "blob_sha256": sha(rblob)}
        rsc = O.OracleScene(rtex["color"], rtex["disk"], rtex["sky"], rblob)
        mc = 1 if rblob is not None else 0
        for name, (w, h, det, camera) in {
            "c1_euler_256x256": (256, 256, U.RayDetails(integration_method=0), U.Camera()),
            "rk_mesh_192x108": (192, 108, U.RayDetails(integration_method=1, model_count=mc), U.Camera()),
            "rk_mesh_outside_192x108": (192, 108, U.RayDetails(integration_method=1, model_count=mc), U.Camera(position=(0, 0, -45))),
        }.items():
            rc = O.ray_pass(rsc, w, h, camera.uniform(), hole, det.uniform(), flavour="contract")
            gold["reference_assets"][name] = {
                "contract_rgba_sha256": sha(rc.rgba), "contract_hit_sha256": sha(rc.hit), "steps_total": rc.counters["steps"],
                "tri_pixels": int((rc.hit >= 0).sum()), "stack_overflow": rc.counters["stack_overflow"], "rk_reject": rc.counters["rk_reject"],
            }
    with open(os.path.join(GOLD, "oracle_frames.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    main()
